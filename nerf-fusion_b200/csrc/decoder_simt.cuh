// FP32 CUDA-core decoder engine: exact-precision path (and the A/B reference for the tcgen05 engine).
//
// One thread owns one query; a CTA of DEC_T threads walks the 4 hidden layers forward and 3 backward with the
// activations of its 128 queries resident in shared memory ([feature][query], conflict-free) and each layer's
// weights staged through shared memory in 64-output chunks read as warp-uniform broadcasts.
// Math restated from network/di_decoder.py:55-86; the reverse pass is SURVEY.md Appendix B.
#pragma once
#include "decoder_common.cuh"

namespace dfb {

constexpr int DEC_T = 128;          // queries per CTA tile
constexpr int DEC_IN = 32;          // 29 latent + 3 xyz
constexpr int DEC_H = 128;

// ---- decoder blob layout (floats); see nerf-fusion_b200/weights.py -------------------------------
// dense() operand layout for a layer with n_in reduction rows and n_out outputs:
//   chunks of 64 outputs (last may be 32): [chunk][n_in][cw]
constexpr int DB_F0 = 0;                         // fwd L0: 32 -> 128      (4096)
constexpr int DB_F1 = DB_F0 + 32 * 128;          // fwd L1: 128 -> 128     (16384)
constexpr int DB_F2 = DB_F1 + 128 * 128;         // fwd L2: 128 -> 96      (12288)
constexpr int DB_F3 = DB_F2 + 128 * 96;          // fwd L3: [h2(96); x(32)] -> 128 (16384)
constexpr int DB_B3 = DB_F3 + 128 * 128;         // bwd through L3: delta3(128) -> delta2(96)   (12288)
constexpr int DB_B2 = DB_B3 + 128 * 96;          // bwd through L2: delta2(96)  -> delta1(128)  (12288)
constexpr int DB_B1 = DB_B2 + 96 * 128;          // bwd through L1: delta1(128) -> delta0(128)  (16384)
constexpr int DB_SMALL = DB_B1 + 128 * 128;      // small block, DS_* offsets below
constexpr int DS_B0 = 0, DS_B1 = 128, DS_B2 = 256, DS_B3 = 352;   // biases 128,128,96,128
constexpr int DS_W4 = 480, DS_WU = 608;                           // heads (128 each)
constexpr int DS_W3X = 736;                                       // W3[o][125..127] as [128][3]
constexpr int DS_W0X = DS_W3X + 384;                              // W0[o][29..31]   as [128][3]
constexpr int DS_B4 = DS_W0X + 384;                               // b4, bu
constexpr int DS_SIZE = DS_B4 + 8;                                // 1512 (padded)
constexpr int DB_TOTAL = DB_SMALL + DS_SIZE;

struct DecSmem {
  float actA[DEC_H * DEC_T];
  float actB[DEC_H * DEC_T];
  float x0[DEC_IN * DEC_T];
  float wstage[DEC_H * 64];
  uint32_t masks[4 * 4 * DEC_T];
  float small_[DS_SIZE];
};

// out[j] = epi( bias[j] + sum_i W[i][j] * in[i] ),  i over n1 rows of in1 followed by n2 rows of in2.
// FWD epilogue: relu + record mask bits;  BWD epilogue: multiply by recorded mask.
template <bool FWD>
__device__ __forceinline__ void dense(DecSmem& S, const float* __restrict__ in1, int n1, const float* __restrict__ in2,
                                      int n2, const float* __restrict__ wblob, const float* __restrict__ bias, int n_out,
                                      float* __restrict__ out, int mask_layer) {
  const int tid = threadIdx.x;
  const int n_in = n1 + n2;
  for (int o0 = 0; o0 < n_out; o0 += 64) {
    const int cw = min(64, n_out - o0);
    __syncthreads();   // previous consumers of wstage / producers of `in` are done
    {
      const float4* src = reinterpret_cast<const float4*>(wblob + (size_t)o0 * n_in);
      float4* dst = reinterpret_cast<float4*>(S.wstage);
      const int n4 = n_in * cw / 4;
      for (int t = tid; t < n4; t += DEC_T) dst[t] = __ldg(src + t);
    }
    __syncthreads();
    for (int p = 0; p < cw; p += 16) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = bias ? bias[o0 + p + j] : 0.f;
      const float* wrow = S.wstage + p;
#pragma unroll 4
      for (int k = 0; k < n1; ++k) {
        const float a = in1[k * DEC_T + tid];
        const float4* w = reinterpret_cast<const float4*>(wrow + k * cw);
        const float4 w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
        acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]); acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
        acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]); acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
        acc[8] = fmaf(a, w2.x, acc[8]); acc[9] = fmaf(a, w2.y, acc[9]); acc[10] = fmaf(a, w2.z, acc[10]); acc[11] = fmaf(a, w2.w, acc[11]);
        acc[12] = fmaf(a, w3.x, acc[12]); acc[13] = fmaf(a, w3.y, acc[13]); acc[14] = fmaf(a, w3.z, acc[14]); acc[15] = fmaf(a, w3.w, acc[15]);
      }
#pragma unroll 4
      for (int k = 0; k < n2; ++k) {
        const float a = in2[k * DEC_T + tid];
        const float4* w = reinterpret_cast<const float4*>(wrow + (n1 + k) * cw);
        const float4 w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
        acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]); acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
        acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]); acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
        acc[8] = fmaf(a, w2.x, acc[8]); acc[9] = fmaf(a, w2.y, acc[9]); acc[10] = fmaf(a, w2.z, acc[10]); acc[11] = fmaf(a, w2.w, acc[11]);
        acc[12] = fmaf(a, w3.x, acc[12]); acc[13] = fmaf(a, w3.y, acc[13]); acc[14] = fmaf(a, w3.z, acc[14]); acc[15] = fmaf(a, w3.w, acc[15]);
      }
      const int o = o0 + p;                       // 16-aligned
      uint32_t* mword = &S.masks[(mask_layer * 4 + (o >> 5)) * DEC_T + tid];
      const int sh = o & 31;                      // 0 or 16
      if (FWD) {
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const bool pos = acc[j] > 0.f;
          bits |= (pos ? 1u : 0u) << j;
          out[(o + j) * DEC_T + tid] = pos ? acc[j] : 0.f;
        }
        if (sh == 0) *mword = bits; else *mword |= bits << 16;
      } else {
        const uint32_t bits = (*mword) >> sh;
#pragma unroll
        for (int j = 0; j < 16; ++j) out[(o + j) * DEC_T + tid] = ((bits >> j) & 1u) ? acc[j] : 0.f;
      }
    }
  }
  __syncthreads();
}

// Forward for the tile; x0 must already hold the 32 inputs of this thread's query.  Returns z (pre-tanh) and u
// (pre-softplus); h3 stays in S.actB.
__device__ __forceinline__ void decoder_forward_tile(DecSmem& S, const float* __restrict__ blob, float& z, float& u) {
  const float* sm = S.small_;
  dense<true>(S, S.x0, 32, nullptr, 0, blob + DB_F0, sm + DS_B0, 128, S.actA, 0);
  dense<true>(S, S.actA, 128, nullptr, 0, blob + DB_F1, sm + DS_B1, 128, S.actB, 1);
  dense<true>(S, S.actB, 128, nullptr, 0, blob + DB_F2, sm + DS_B2, 96, S.actA, 2);
  dense<true>(S, S.actA, 96, S.x0, 32, blob + DB_F3, sm + DS_B3, 128, S.actB, 3);
  const int tid = threadIdx.x;
  float zz = sm[DS_B4], uu = sm[DS_B4 + 1];
#pragma unroll 8
  for (int o = 0; o < 128; ++o) {
    const float h = S.actB[o * DEC_T + tid];
    zz = fmaf(sm[DS_W4 + o], h, zz);
    uu = fmaf(sm[DS_WU + o], h, uu);
  }
  z = zz; u = uu;
}

// Reverse pass: seeds on z and u -> gradient w.r.t. the 3 xyz inputs (network units).
__device__ __forceinline__ void decoder_backward_tile(DecSmem& S, const float* __restrict__ blob, float seed_z, float seed_u,
                                                      float g[3]) {
  const float* sm = S.small_;
  const int tid = threadIdx.x;
  float gx = 0.f, gy = 0.f, gz = 0.f;
  // delta3 = (seed_z W4 + seed_u Wu) * [a3 > 0]  -> actA
#pragma unroll 4
  for (int o = 0; o < 128; ++o) {
    const uint32_t bit = (S.masks[(3 * 4 + (o >> 5)) * DEC_T + tid] >> (o & 31)) & 1u;
    const float d = bit ? fmaf(seed_z, sm[DS_W4 + o], seed_u * sm[DS_WU + o]) : 0.f;
    S.actA[o * DEC_T + tid] = d;
    gx = fmaf(sm[DS_W3X + 3 * o + 0], d, gx);
    gy = fmaf(sm[DS_W3X + 3 * o + 1], d, gy);
    gz = fmaf(sm[DS_W3X + 3 * o + 2], d, gz);
  }
  dense<false>(S, S.actA, 128, nullptr, 0, blob + DB_B3, nullptr, 96, S.actB, 2);    // delta2
  dense<false>(S, S.actB, 96, nullptr, 0, blob + DB_B2, nullptr, 128, S.actA, 1);    // delta1
  dense<false>(S, S.actA, 128, nullptr, 0, blob + DB_B1, nullptr, 128, S.actB, 0);   // delta0
#pragma unroll 4
  for (int o = 0; o < 128; ++o) {
    const float d = S.actB[o * DEC_T + tid];
    gx = fmaf(sm[DS_W0X + 3 * o + 0], d, gx);
    gy = fmaf(sm[DS_W0X + 3 * o + 1], d, gy);
    gz = fmaf(sm[DS_W0X + 3 * o + 2], d, gz);
  }
  g[0] = gx; g[1] = gy; g[2] = gz;
}


// Reverse pass with the gradient w.r.t. ALL 32 network inputs (29 latent + 3 xyz), for the latent optimiser
// (map.py:81-113 back-propagates into latent_vecs_unique): g_in[k] = sum_o delta0[o] W0[o][k] + sum_o delta3[o] W3[o][96 + k].
// The two products read the forward operands transposed (F0 / F3 are [k][o] in chunks of 64 outputs); warp-uniform
// addresses, i.e. broadcast loads.  Not a hot path (the optimiser is disabled by the target config).
__device__ __forceinline__ void decoder_backward_inputs_tile(DecSmem& S, const float* __restrict__ blob, float seed_z, float seed_u,
                                                             float g_in[DEC_IN]) {
  const float* sm = S.small_;
  const int tid = threadIdx.x;
#pragma unroll
  for (int k = 0; k < DEC_IN; ++k) g_in[k] = 0.f;
#pragma unroll 1
  for (int o = 0; o < 128; ++o) {                 // delta3 = (seed_z W4 + seed_u Wu) * [a3 > 0]  -> actA
    const uint32_t bit = (S.masks[(3 * 4 + (o >> 5)) * DEC_T + tid] >> (o & 31)) & 1u;
    const float d = bit ? fmaf(seed_z, sm[DS_W4 + o], seed_u * sm[DS_WU + o]) : 0.f;
    S.actA[o * DEC_T + tid] = d;
    const float* w = blob + DB_F3 + (o >> 6) * (128 * 64) + (o & 63);       // F3[(96 + k)][o]
#pragma unroll
    for (int k = 0; k < DEC_IN; ++k) g_in[k] = fmaf(__ldg(w + (96 + k) * 64), d, g_in[k]);
  }
  dense<false>(S, S.actA, 128, nullptr, 0, blob + DB_B3, nullptr, 96, S.actB, 2);    // delta2
  dense<false>(S, S.actB, 96, nullptr, 0, blob + DB_B2, nullptr, 128, S.actA, 1);    // delta1
  dense<false>(S, S.actA, 128, nullptr, 0, blob + DB_B1, nullptr, 128, S.actB, 0);   // delta0
#pragma unroll 1
  for (int o = 0; o < 128; ++o) {
    const float d = S.actB[o * DEC_T + tid];
    const float* w = blob + DB_F0 + (o >> 6) * (32 * 64) + (o & 63);        // F0[k][o]
#pragma unroll
    for (int k = 0; k < DEC_IN; ++k) g_in[k] = fmaf(__ldg(w + k * 64), d, g_in[k]);
  }
}

__device__ __forceinline__ void decoder_load_small(DecSmem& S, const float* __restrict__ blob) {
  for (int t = threadIdx.x; t < DS_SIZE; t += DEC_T) S.small_[t] = __ldg(blob + DB_SMALL + t);
  __syncthreads();
}

}  // namespace dfb
