// tcgen05 encoder engine (sm_100a): the per-sample encoder MLP 6 -> 32 -> 64 -> 256 -> 29 (network/di_encoder.py:26-30,
// BatchNorm folded on the host) on the tensor cores, fused with the scatter-add into the per-voxel accumulators
// (map.py:446-449, indexing.cu:59-71).
//
// Same construction as decoder_tc.cu: FP16 weight images resident in shared memory (bulk-TMA loads), ACTIVATIONS RESIDENT
// IN TENSOR MEMORY (every layer is D[tmem] = A[tmem] * B[smem], the epilogue warps turn the FP32 accumulator into the
// next layer's A operand with tcgen05.ld / tcgen05.st), and every product computed as three FP16 MMAs
//     A W ~= A_hi W_hi + A_lo W_hi + A_hi W_lo      (hi = fp16(x), lo = fp16(x - hi))
// so the latents agree with an FP32 evaluation to ~1e-6 instead of the 2e-4 of single-FP16 operands.
//
// Tensor-memory plan (224 of 256 columns, so TWO CTAs -- two tiles -- are resident per SM; the images take 110 KB):
//   X = [  0,128)  accumulator of layers 0, 1 and of one 128-column HALF of layer 2; the layer-2 epilogue writes the
//                  layer-3 operand IN PLACE: a thread reads 16 FP32 columns and stores 8 hi + 8 lo packed columns over them
//   Y = [128,192)  A operand of layers 0, 1, 2 (per 16-element K step: 8 hi columns, then 8 lo columns)
//   Z = [192,224)  layer-3 accumulator (the 29 latents), accumulated over the two K halves
// Layer order per tile: L0, L1, L2a, L3a, L2b, L3b.
#include <algorithm>

#include "tc_common.cuh"

namespace dfb {
namespace etc {
using namespace tcp;

constexpr int NPART = 2;           // threads per row
constexpr int CTA_T = T * NPART;   // one tile per CTA, two CTAs per SM
// ---- blob (bytes): FP16 SWIZZLE_128B K-major images; packed by weights.pack_encoder_tc -----------------------------
constexpr int IMG_W0 = 0;          // [ 32 rows x 64]: cols 0..5 = hi(W0), cols 16..21 = lo(W0)
constexpr int IMG_W1 = 4096;       // [ 64 rows x 64]: cols 0..31 = hi(W1), cols 32..63 = lo(W1)
constexpr int IMG_W2H = 12288;     // [256 rows x 64]
constexpr int IMG_W2L = 45056;
constexpr int IMG_W3H = 77824;     // 4 blocks x [32 rows x 64]: rows 0..28 = W3
constexpr int IMG_W3L = 94208;
constexpr int IMG_END = 110592;
constexpr int ES_B0 = 0, ES_B1 = 32, ES_B2 = 96, ES_B3 = 352;   // FP32 biases (floats)
constexpr int SMALL_BYTES = 2048;
constexpr int BLOB_BYTES = IMG_END + SMALL_BYTES;
static_assert(BLOB_BYTES == 112640, "integrate.cu ENC_TC_BLOB_BYTES / weights.pack_encoder_tc");
// ---- shared memory -------------------------------------------------------------------------------------------------
constexpr int SM_SMALL = IMG_END;
constexpr int SM_BAR = BLOB_BYTES;
constexpr int SM_TOTAL = SM_BAR + 64;
constexpr int SM_ALLOC = SM_TOTAL + 1024;
static_assert(2 * (SM_ALLOC + 1024) <= 233472, "two CTAs per SM");
// ---- tensor memory -------------------------------------------------------------------------------------------------
constexpr int TM_X = 0, TM_Y = 128, TM_Z = 192;
constexpr int TMEM_COLS = 256;

struct Sample {   // same 32-byte record as integrate.cu
  int slot;
  float rel[3];
  float nrm[3];
  int pad;
};

struct Ctx {
  uint8_t* sm;
  uint32_t sa, tmem, mma_bar, wbar, phase;
  int row, part;
};

__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// Issuing thread only.  A: KSTEPS K steps from tensor memory at a_tmem, 16 columns apart (8 hi columns, then 8 lo); A_LO picks
// the lo half.  B: K-major image at b_addr, B_ROWS rows per 64-column block, rows [N0, N0 + N), starting at K step B_S0.
template <int KSTEPS, int N, int B_ROWS, int B_S0, bool A_LO, int N0 = 0>
__device__ __forceinline__ void issue(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_addr, bool accumulate_first) {
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(T >> 4) << 24);
  const uint64_t bd0 = smem_desc(b_addr + N0 * 128, 16, 1024);
#pragma unroll
  for (int s = 0; s < KSTEPS; ++s) {
    const int sb = s + B_S0;
    const uint64_t bd = bd0 + (uint64_t)(((sb >> 2) * (B_ROWS * 128) + (sb & 3) * 32) >> 4);
    mma_f16_ts(d_tmem, a_tmem + 16u * s + (A_LO ? 8u : 0u), bd, idesc, (s > 0 || accumulate_first) ? 1u : 0u);
  }
}
// A_hi W_hi + A_lo W_hi + A_hi W_lo with the hi / lo weight images side by side in one 64-column block (lo at K step LO_S0)
template <int KSTEPS, int N, int B_ROWS, int LO_S0>
__device__ __forceinline__ void issue3_packed(uint32_t d, uint32_t a, uint32_t img) {
  issue<KSTEPS, N, B_ROWS, 0, false>(d, a, img, false);
  issue<KSTEPS, N, B_ROWS, 0, true>(d, a, img, true);
  issue<KSTEPS, N, B_ROWS, LO_S0, false>(d, a, img, true);
}
// same with separate hi / lo images, rows [N0, N0 + N) of both, K steps from B_S0
template <int KSTEPS, int N, int B_ROWS, int B_S0, int N0>
__device__ __forceinline__ void issue3(uint32_t d, uint32_t a, uint32_t img_hi, uint32_t img_lo, bool accumulate_first) {
  issue<KSTEPS, N, B_ROWS, B_S0, false, N0>(d, a, img_hi, accumulate_first);
  issue<KSTEPS, N, B_ROWS, B_S0, true, N0>(d, a, img_hi, true);
  issue<KSTEPS, N, B_ROWS, B_S0, false, N0>(d, a, img_lo, true);
}

#define ETC_LAYER(ISSUE_STMT)                         \
  do {                                                \
    tmem_st_wait();                                   \
    tc_fence_before();                                \
    __syncthreads();                                  \
    if (threadIdx.x == 0) {                           \
      tc_fence_after();                               \
      ISSUE_STMT;                                     \
      mma_commit(c.mma_bar);                          \
    }                                                 \
    mbar_wait(c.mma_bar, c.phase);                    \
    c.phase ^= 1u;                                    \
    tc_fence_after();                                 \
  } while (0)

__device__ __forceinline__ uint32_t lane_base(const Ctx& c) { return c.tmem + ((uint32_t)(c.row & ~31) << 16); }

// 16 values of this thread's row -> 8 hi + 8 lo packed columns at dst .. dst + 15
__device__ __forceinline__ void store_step(const Ctx& c, uint32_t dst_col, const float* h) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) split_h2(h[2 * q], h[2 * q + 1], hi[q], lo[q]);
  const uint32_t tb = lane_base(c) + dst_col;
  tmem_st8(tb, hi);
  tmem_st8(tb + 8, lo);
}

// hidden-layer epilogue: relu(D + b) of NCOLS accumulator columns at X -> A operand K steps at dst (NCOLS / 16 steps);
// this thread takes every NPART-th step.  In place when dst == TM_X (a step only overwrites the columns it was read from).
template <int NCOLS>
__device__ __forceinline__ void epi_hidden(Ctx& c, int bias_off, uint32_t dst) {
  const float* sm = reinterpret_cast<const float*>(c.sm + SM_SMALL);
  const uint32_t tb = lane_base(c) + TM_X;
#pragma unroll
  for (int q = 0; q < NCOLS / 16 / NPART; ++q) {
    const int step = NPART * q + c.part;
    float v[16];
    tmem_ld16(tb + 16 * step, v);
    const float4* b4 = reinterpret_cast<const float4*>(sm + bias_off + 16 * step);
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
      const float4 b = b4[i4];
      v[4 * i4] = fmaxf(v[4 * i4] + b.x, 0.f); v[4 * i4 + 1] = fmaxf(v[4 * i4 + 1] + b.y, 0.f);
      v[4 * i4 + 2] = fmaxf(v[4 * i4 + 2] + b.z, 0.f); v[4 * i4 + 3] = fmaxf(v[4 * i4 + 3] + b.w, 0.f);
    }
    store_step(c, dst + 16 * step, v);
  }
}

extern __shared__ unsigned char etc_smem_raw[];

__device__ __forceinline__ void prologue(Ctx& c, const void* blob) {
  const uint32_t raw = smem_u32(etc_smem_raw);
  const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
  c.sm = etc_smem_raw + pad;
  c.sa = raw + pad;
  c.phase = 0;
  c.part = threadIdx.x / T;
  c.row = threadIdx.x % T;
  c.wbar = c.sa + SM_BAR;
  const uint32_t slot = c.sa + SM_BAR + 48;
  c.mma_bar = c.sa + SM_BAR + 8;
  if (threadIdx.x == 0) {
    mbar_init(c.wbar, 1);
    mbar_init(c.mma_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_alloc(slot, TMEM_COLS);
  if (threadIdx.x == 0) {
    mbar_expect_tx(c.wbar, BLOB_BYTES);
    const char* src = reinterpret_cast<const char*>(blob);
    for (int off = 0; off < BLOB_BYTES; off += 8192) bulk_g2s(c.sa + off, src + off, (uint32_t)min(8192, BLOB_BYTES - off), c.wbar);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *reinterpret_cast<volatile uint32_t*>(c.sm + SM_BAR + 48);
}

// one tile: in[6] of this row known to both threads of the row.  Leaves the outputs (+bias) of columns
// [16 part, 16 part + 16) in out16 (29 are latents).
__device__ __forceinline__ void encode_tile(Ctx& c, const float in[6], float out16[16]) {
  const float* sm = reinterpret_cast<const float*>(c.sm + SM_SMALL);
  if (c.part == 0) {                                              // layer-0 operand: one K step (6 inputs, 10 zeros)
    float h[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) h[k] = k < 6 ? in[k] : 0.f;
    store_step(c, TM_Y, h);
  }
  mbar_wait(c.wbar, 0);                                           // completes once; later calls return immediately
  const uint32_t X = c.tmem + TM_X, Y = c.tmem + TM_Y, Z = c.tmem + TM_Z;
  ETC_LAYER((issue3_packed<1, 32, 32, 1>(X, Y, c.sa + IMG_W0)));                                        // 6 -> 32
  epi_hidden<32>(c, ES_B0, TM_Y);
  ETC_LAYER((issue3_packed<2, 64, 64, 2>(X, Y, c.sa + IMG_W1)));                                        // 32 -> 64
  epi_hidden<64>(c, ES_B1, TM_Y);
  ETC_LAYER((issue3<4, 128, 256, 0, 0>(X, Y, c.sa + IMG_W2H, c.sa + IMG_W2L, false)));                  // 64 -> 256, outputs 0..127
  epi_hidden<128>(c, ES_B2, TM_X);                                                                      //   in place: layer-3 operand, K 0..127
  ETC_LAYER((issue3<8, 32, 32, 0, 0>(Z, X, c.sa + IMG_W3H, c.sa + IMG_W3L, false)));                    // 256 -> 29, K 0..127
  ETC_LAYER((issue3<4, 128, 256, 0, 128>(X, Y, c.sa + IMG_W2H, c.sa + IMG_W2L, false)));                // 64 -> 256, outputs 128..255
  epi_hidden<128>(c, ES_B2 + 128, TM_X);
  ETC_LAYER((issue3<8, 32, 32, 8, 0>(Z, X, c.sa + IMG_W3H, c.sa + IMG_W3L, true)));                     // 256 -> 29, K 128..255
  tmem_ld16(lane_base(c) + TM_Z + 16 * c.part, out16);
#pragma unroll
  for (int i = 0; i < 16; ++i) out16[i] += sm[ES_B3 + 16 * c.part + i];
}

// Input forms: explicit rows x6 (SCATTER = false), one contiguous sample array with its length on the device (m_dev), or --
// for the sharded map -- N_SEG fixed-capacity segments of a receive buffer (segment s = rows [s seg_cap, s seg_cap +
// seg_counts[s])), consumed in place: no compaction pass between the exchange and the encoder.
constexpr int MAX_SEG = 8;
template <bool SCATTER>
__global__ void __launch_bounds__(CTA_T, 2) encoder_kernel(const Sample* __restrict__ samples, const float* __restrict__ x6, int m_host,
                                                           const int* __restrict__ m_dev, const int* __restrict__ seg_counts, int n_seg,
                                                           int seg_cap, const void* __restrict__ blob, float* __restrict__ out) {
  Ctx c;
  prologue(c, blob);
  int cum[MAX_SEG + 1];                                  // tiles before segment s
  cum[0] = 0;
  if (n_seg > 0) {
#pragma unroll
    for (int sgi = 0; sgi < MAX_SEG; ++sgi) cum[sgi + 1] = cum[sgi] + (sgi < n_seg ? (min(seg_counts[sgi], seg_cap) + T - 1) / T : 0);
  } else {
    const int m = m_dev ? *m_dev : m_host;
#pragma unroll
    for (int sgi = 0; sgi < MAX_SEG; ++sgi) cum[sgi + 1] = (m + T - 1) / T;
  }
  const int m0 = n_seg > 0 ? 0 : (m_dev ? *m_dev : m_host);
  for (int tile = blockIdx.x; tile < cum[MAX_SEG]; tile += gridDim.x) {
    int seg = 0;
#pragma unroll
    for (int sgi = 1; sgi < MAX_SEG; ++sgi) seg += (n_seg > 0 && tile >= cum[sgi]) ? 1 : 0;
    int base = 0;
#pragma unroll
    for (int sgi = 0; sgi < MAX_SEG; ++sgi) base = (sgi == seg) ? cum[sgi] : base;
    const int il = (tile - base) * T + c.row;                                  // row inside the segment (or the whole input)
    const int m = n_seg > 0 ? min(seg_counts[seg], seg_cap) : m0;
    const bool valid = il < m;
    const long long i = n_seg > 0 ? (long long)seg * seg_cap + il : il;
    float in[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int slot = 0;
    if (valid) {
      if (SCATTER) {
        const Sample s = samples[i];
        in[0] = s.rel[0]; in[1] = s.rel[1]; in[2] = s.rel[2]; in[3] = s.nrm[0]; in[4] = s.nrm[1]; in[5] = s.nrm[2];
        slot = s.slot;
      } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) in[k] = x6[(size_t)i * 6 + k];
      }
    }
    float o[16];
    encode_tile(c, in, o);
    if (valid && slot >= 0) {
      // SCATTER: `out` is the map's fixed-point accumulator array (common.cuh acc_add: order-independent sums)
      float* dst = out + (size_t)i * DFB_LATENT_DIM + 16 * c.part;
      long long* adst = reinterpret_cast<long long*>(out) + (size_t)slot * DFB_LATENT_DIM + 16 * c.part;
#pragma unroll
      for (int l = 0; l < 16; ++l) {
        if (16 * c.part + l < DFB_LATENT_DIM) {
          if (SCATTER) acc_add(adst + l, o[l]); else dst[l] = o[l];
        }
      }
    }
    // the next tile's layer-0 operand goes to Y, which layer 2b read last: its MMA has completed (waited above); Z is read
    // by this thread only
  }
  mbar_wait(c.wbar, 0);                       // a CTA without tiles must not exit with its bulk copies in flight
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(c.tmem, TMEM_COLS);
}

}  // namespace etc

static int enc_grid(long long m) { return (int)std::min<long long>(div_up(m, etc::T), 2LL * sm_count()); }

int tc_encoder_scatter(const void* samples, const int* m_dev, int m_max, const void* tc_blob, long long* acc, cudaStream_t s) {
  DFB_CUDA(cudaFuncSetAttribute(etc::encoder_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, etc::SM_ALLOC));
  etc::encoder_kernel<true><<<enc_grid(m_max), etc::CTA_T, etc::SM_ALLOC, s>>>(reinterpret_cast<const etc::Sample*>(samples), nullptr, 0, m_dev,
                                                                            nullptr, 0, 0, tc_blob, reinterpret_cast<float*>(acc));
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

// sharded map: the samples are the segments of the exchange's receive buffer (sharded.cu)
int tc_encoder_scatter_segments(const void* samples, const int* seg_counts, int n_seg, int seg_cap, const void* tc_blob, long long* acc,
                                cudaStream_t s) {
  if (n_seg < 1 || n_seg > etc::MAX_SEG) { set_error("encoder: 1..%d segments", etc::MAX_SEG); return DFB_E_INVALID; }
  DFB_CUDA(cudaFuncSetAttribute(etc::encoder_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, etc::SM_ALLOC));
  etc::encoder_kernel<true><<<enc_grid((long long)n_seg * seg_cap), etc::CTA_T, etc::SM_ALLOC, s>>>(
      reinterpret_cast<const etc::Sample*>(samples), nullptr, 0, nullptr, seg_counts, n_seg, seg_cap, tc_blob, reinterpret_cast<float*>(acc));
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int tc_encoder_explicit(const float* x6, int m, const void* tc_blob, float* out, cudaStream_t s) {
  DFB_CUDA(cudaFuncSetAttribute(etc::encoder_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, etc::SM_ALLOC));
  etc::encoder_kernel<false><<<enc_grid(m), etc::CTA_T, etc::SM_ALLOC, s>>>(nullptr, x6, m, nullptr, nullptr, 0, 0, tc_blob, out);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

}  // namespace dfb
