// tcgen05 decoder engine (sm_100a): the 5-layer decoder MLP, its reverse pass and the tracker reduction with the
// weights RESIDENT in shared memory (one bulk-TMA load per CTA), activations as FP16 K-major SWIZZLE_128B tiles,
// FP32 accumulators in TMEM, one elected thread issuing tcgen05.mma, NPART x 128 threads per tile (NPART threads per
// query row = TMEM lane, each owning 128 / NPART accumulator columns) running the epilogues (bias/ReLU/mask -> FP16 ->
// swizzled smem) between layers, GROUPS tiles in flight per CTA.  The file also holds gn_eval_kernel: one Gauss-Newton
// evaluation (SDF term on these tiles + photometric pixels + block reduction + last-block step) in a single launch.
//
//   forward   D[128 x N] = A[128 x K] * W^T      A: activations (K-major), B: weight image (K-major)
//   reverse   D[128 x K] = delta[128 x N] * W    A: deltas (K-major),      B: the SAME weight image read MN-major
//
// so one FP16 image per layer serves both passes (120 KB for the whole network).  Inputs are split hi+lo into the
// two halves of the 64-wide input block (weights duplicated) so positions/latents enter with ~22 significant bits.
// Math: network/di_decoder.py:55-86; reverse pass: SURVEY.md Appendix B.
#include <algorithm>

#include "decoder_common.cuh"
#include "tc_common.cuh"
#include "decoder_simt.cuh"   // DS_* offsets of the FP32 "small" block (biases, heads, xyz taps)
#include "photometric.cuh"
#include "gn_step.cuh"

namespace dfb {
namespace tc {
using namespace tcp;

constexpr int GROUPS = 2;       // tiles in flight per CTA (each with its own 4 x NPART warps, smem tiles, TMEM columns, barrier)
constexpr int NPART = 2;        // threads per row: part p owns accumulator columns [CW p, CW p + CW) (warps 4p..4p+3 of the group; same TMEM
                                // lanes).  Measured on B200 at 2^24 queries: NPART=2 1.42 G q/s, NPART=4 1.14 G q/s (barrier + redundant
                                // front-end cost outweighs the extra warps)
constexpr int CW = 128 / NPART; // accumulator columns per thread
constexpr int GT = T * NPART;   // threads per tile group
constexpr int CTA_T = GT * GROUPS;
// ---- blob (bytes): FP16 swizzled images + FP32 small block; packed by weights.pack_decoder_tc -------------------
constexpr int IMG_W0 = 0;          // [128 rows x 64]: cols 0..31 = W0, cols 32..63 = W0 again (lo halves of the input)
constexpr int IMG_W1 = 16384;      // 2 blocks x [128 x 64]
constexpr int IMG_W2 = 49152;      // 2 blocks x [ 96 x 64]
constexpr int IMG_W3A = 73728;     // 2 blocks x [128 x 64]: W3[:, 0:96], cols 96..127 zero
constexpr int IMG_W3B = 106496;    // [128 x 64]: W3[:, 96:128] twice (hi, lo)
constexpr int IMG_END = 122880;
constexpr int SMALL_BYTES = 6144;
constexpr int BLOB_BYTES = IMG_END + SMALL_BYTES;
// ---- shared memory (bytes from the 1024-aligned base) ------------------------------------------------------------
constexpr int SM_SMALL = IMG_END;
constexpr int SM_A = SM_SMALL + SMALL_BYTES;   // 129024 = 126 * 1024: activation tile, 2 blocks x [128 x 64] fp16
constexpr int SM_XOFF = 32768;                 // input tile [128 x 64] fp16 (hi | lo), after the activation tile
constexpr int SM_TILE_BYTES = 32768 + 16384;   // per tile group
constexpr int SM_BAR = SM_A + GROUPS * SM_TILE_BYTES;   // mbarriers + TMEM slot
constexpr int SM_TOTAL = SM_BAR + 64;
constexpr int SM_ALLOC = SM_TOTAL + 1024;      // slack for manual 1024-byte alignment
constexpr int TMEM_COLS = 128 * GROUPS;

struct Ctx {
  uint8_t* sm;        // 1024-aligned base (generic)
  uint32_t sa;        // same, shared-window address
  uint32_t tmem;      // TMEM address of this group's accumulator (lane 0)
  uint32_t tmem_base; // allocation base
  int row, part, grp; // row within the tile (= TMEM lane), column part owned by this thread, tile group
  uint32_t a_off, x_off;   // byte offsets of this group's activation / input tiles
  uint32_t mma_bar;   // shared address of the "MMA done" barrier
  uint32_t wbar;      // shared address of the "weights landed" barrier (completes once)
  uint32_t phase;     // its parity
  uint32_t mask[4][CW / 32];   // ReLU masks of this thread's columns, per layer
};

// Issuing thread only.  K-major A (activation tile at a_addr, blocks of 16 KB), B = weight image at b_addr with B_ROWS rows
// per 64-column block.  fwd: B read K-major; bwd: the same image read MN-major (k-step s = image rows 16s..16s+15).
// Everything but the two base addresses is a compile-time constant, so each step is two 64-bit adds and one tcgen05.mma.
template <int KSTEPS, int N, bool BWD, int B_ROWS>
__device__ __forceinline__ void issue(const Ctx& c, uint32_t a_addr, uint32_t b_addr, bool accumulate_first) {
  constexpr uint32_t idesc = (1u << 4) | ((BWD ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(T >> 4) << 24);
  const uint64_t ad0 = smem_desc(a_addr, 16, 1024);
  const uint64_t bd0 = BWD ? smem_desc(b_addr, (uint32_t)B_ROWS * 128u, 1024) : smem_desc(b_addr, 16, 1024);
#pragma unroll
  for (int s = 0; s < KSTEPS; ++s) {
    const uint64_t ad = ad0 + (uint64_t)(((s >> 2) * 16384 + (s & 3) * 32) >> 4);
    const uint64_t bd = bd0 + (uint64_t)((BWD ? s * 2048 : (s >> 2) * (B_ROWS * 128) + (s & 3) * 32) >> 4);
    mma_f16(c.tmem, ad, bd, idesc, (s > 0 || accumulate_first) ? 1u : 0u);
  }
}

#ifdef DFB_TC_PROFILE
__device__ unsigned long long g_prof[256];
__device__ int g_prof_n;
#define PROF_MARK(c)                                                                         \
  do {                                                                                       \
    if (blockIdx.x == 0 && (c).grp == 0 && (c).row == 0 && (c).part == 0) {                  \
      int k_ = g_prof_n; if (k_ < 256) { g_prof[k_] = clock64(); g_prof_n = k_ + 1; }        \
    }                                                                                        \
  } while (0)
#else
#define PROF_MARK(c) do {} while (0)
#endif

// Tile schedule of a persistent CTA: full rounds give every (CTA, group) slot one tile; the last, partial round hands its
// tiles out one per CTA first (group 0 of every CTA, then group 1), so that a CTA whose sibling group has nothing left
// runs its tile with the SM's issue slots and the tensor pipe to itself.  Returns -1 when the group is done.
__device__ __forceinline__ long long tile_of(const Ctx& c, long long n, long long rnd) {
  const long long tiles = (n + T - 1) / T, slots = (long long)gridDim.x * GROUPS;
  const long long full = tiles / slots;
  if (rnd < full) return rnd * slots + (long long)blockIdx.x * GROUPS + c.grp;
  if (rnd > full) return -1;
  const long long j = (long long)c.grp * gridDim.x + blockIdx.x;
  return j < tiles - full * slots ? full * slots + j : -1;
}

__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(GT) : "memory"); }

// all threads of the group: make the tile writes visible to the tensor core, then thread 0 issues; everyone waits for completion
#define TC_LAYER(ISSUE_STMTS)                         \
  do {                                                \
    PROF_MARK(c);                                     \
    fence_proxy_async();                              \
    tc_fence_before();                                \
    group_sync(c.grp);                                \
    PROF_MARK(c);                                     \
    if (c.row == 0 && c.part == 0) {                  \
      tc_fence_after();                               \
      ISSUE_STMTS;                                    \
      mma_commit(c.mma_bar);                          \
    }                                                 \
    PROF_MARK(c);                                     \
    mbar_wait(c.mma_bar, c.phase);                    \
    c.phase ^= 1u;                                    \
    tc_fence_after();                                 \
    PROF_MARK(c);                                     \
  } while (0)

// store 32 consecutive columns [col0, col0+32) of this thread's row into a swizzled K-major tile
__device__ __forceinline__ void store_cols32(uint8_t* tile, int row, int col0, const float* h) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = (col0 >> 3) + q;   // 16-byte chunk index along K
    uint4 v;
    v.x = pack_h2(h[8 * q + 0], h[8 * q + 1]); v.y = pack_h2(h[8 * q + 2], h[8 * q + 3]);
    v.z = pack_h2(h[8 * q + 4], h[8 * q + 5]); v.w = pack_h2(h[8 * q + 6], h[8 * q + 7]);
    *reinterpret_cast<uint4*>(tile + (chunk >> 3) * 16384 + row * 128 + (((chunk & 7) ^ (row & 7)) << 4)) = v;
  }
}

// hidden-layer forward epilogue for this thread's column half: a = D + b, record [a > 0], relu(a) -> FP16 activation tile
template <int NCOLS>
__device__ __forceinline__ void epi_fwd(Ctx& c, int layer, int bias_off) {
  const float* sm = reinterpret_cast<const float*>(c.sm + SM_SMALL);
  const int row = c.row;
  const uint32_t tbase = c.tmem + ((uint32_t)(row & ~31) << 16);
#pragma unroll
  for (int jj = 0; jj < CW / 32; ++jj) {
    const int col0 = CW * c.part + 32 * jj;
    if (col0 < NCOLS) {                     // warp-uniform
      float v[32];
      tmem_ld32(tbase + col0, v);
      uint32_t m = 0;
      const float4* b4 = reinterpret_cast<const float4*>(sm + bias_off + col0);
#pragma unroll
      for (int i4 = 0; i4 < 8; ++i4) {
        const float4 b = b4[i4];
        const float a0 = v[4 * i4] + b.x, a1 = v[4 * i4 + 1] + b.y, a2 = v[4 * i4 + 2] + b.z, a3 = v[4 * i4 + 3] + b.w;
        m |= ((a0 > 0.f ? 1u : 0u) | (a1 > 0.f ? 2u : 0u) | (a2 > 0.f ? 4u : 0u) | (a3 > 0.f ? 8u : 0u)) << (4 * i4);
        v[4 * i4] = fmaxf(a0, 0.f); v[4 * i4 + 1] = fmaxf(a1, 0.f); v[4 * i4 + 2] = fmaxf(a2, 0.f); v[4 * i4 + 3] = fmaxf(a3, 0.f);
      }
      c.mask[layer][jj] = m;
      store_cols32(c.sm + c.a_off, row, col0, v);
    }
  }
}

// reverse epilogue: delta = D * mask[layer] -> FP16 tile
template <int NCOLS>
__device__ __forceinline__ void epi_bwd(Ctx& c, int layer) {
  const int row = c.row;
  const uint32_t tbase = c.tmem + ((uint32_t)(row & ~31) << 16);
#pragma unroll
  for (int jj = 0; jj < CW / 32; ++jj) {
    const int col0 = CW * c.part + 32 * jj;
    if (col0 < NCOLS) {
      float v[32];
      tmem_ld32(tbase + col0, v);
      const uint32_t m = c.mask[layer][jj];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ((m >> i) & 1u) ? v[i] : 0.f;
      store_cols32(c.sm + c.a_off, row, col0, v);
    }
  }
}

// the two threads of a row (column halves) combine partial sums through the (currently idle) input tile
__device__ __forceinline__ void exchange(Ctx& c, float* vals, int nvals) {
  float* ex = reinterpret_cast<float*>(c.sm + c.x_off);
  for (int k = 0; k < nvals; ++k) ex[(c.part * T + c.row) * 4 + k] = vals[k];
  group_sync(c.grp);
  for (int k = 0; k < nvals; ++k) {
    float sacc = ex[c.row * 4 + k];
#pragma unroll
    for (int p = 1; p < NPART; ++p) sacc += ex[(p * T + c.row) * 4 + k];   // fixed order: all threads of the row get the same bits
    vals[k] = sacc;
  }
  group_sync(c.grp);
}

// Write this row's query (29 latent + 3 rel) into the input tile [hi(32) | lo(32)]: each of the NPART threads of the row
// stores XC = 64 / NPART consecutive columns (FP16 hi parts in columns 0..31, x - hi in 32..63).
constexpr int XC = 64 / NPART;
__device__ __forceinline__ void store_input(Ctx& c, const float* __restrict__ latent_row, const float rel[3], bool valid) {
  const int col_begin = XC * c.part;
  const bool lo = col_begin >= 32;
  const int k0 = col_begin & 31;
  float h[XC];
#pragma unroll
  for (int i = 0; i < XC; ++i) {
    const int k = k0 + i;
    float x = 0.f;
    if (valid) x = k < DFB_LATENT_DIM ? __ldg(latent_row + k) : (k == DFB_LATENT_DIM ? rel[0] : (k == DFB_LATENT_DIM + 1 ? rel[1] : rel[2]));
    const float hi = __half2float(__float2half_rn(x));
    h[i] = lo ? x - hi : hi;
  }
#pragma unroll
  for (int q = 0; q < XC / 8; ++q) {
    const int chunk = (col_begin >> 3) + q;
    uint4 v;
    v.x = pack_h2(h[8 * q + 0], h[8 * q + 1]); v.y = pack_h2(h[8 * q + 2], h[8 * q + 3]);
    v.z = pack_h2(h[8 * q + 4], h[8 * q + 5]); v.w = pack_h2(h[8 * q + 6], h[8 * q + 7]);
    *reinterpret_cast<uint4*>(c.sm + c.x_off + c.row * 128 + (((chunk & 7) ^ (c.row & 7)) << 4)) = v;
  }
}

// forward pass of the tile; returns pre-activation heads z (sdf) and u (std)
__device__ __forceinline__ void forward(Ctx& c, float& z, float& u) {
  const float* sm = reinterpret_cast<const float*>(c.sm + SM_SMALL);
  mbar_wait(c.wbar, 0);   // phase 0 completes once (bulk load of the network); later calls return immediately
  TC_LAYER((issue<4, 128, false, 128>(c, c.sa + c.x_off, c.sa + IMG_W0, false)));
  epi_fwd<128>(c, 0, DS_B0);
  TC_LAYER((issue<8, 128, false, 128>(c, c.sa + c.a_off, c.sa + IMG_W1, false)));
  epi_fwd<128>(c, 1, DS_B1);
  TC_LAYER((issue<8, 96, false, 96>(c, c.sa + c.a_off, c.sa + IMG_W2, false)));
  epi_fwd<96>(c, 2, DS_B2);
  TC_LAYER((issue<8, 128, false, 128>(c, c.sa + c.a_off, c.sa + IMG_W3A, false));
           (issue<4, 128, false, 128>(c, c.sa + c.x_off, c.sa + IMG_W3B, true)));
  // heads in FP32 straight from the accumulator (h3 never leaves TMEM/registers); each thread sums its 64 columns
  const int row = c.row;
  const uint32_t tbase = c.tmem + ((uint32_t)(row & ~31) << 16);
  float zu[2] = {0.f, 0.f};
  float z1 = 0.f, u1 = 0.f;
#pragma unroll
  for (int jj = 0; jj < CW / 32; ++jj) {
    const int col0 = CW * c.part + 32 * jj;
    float v[32];
    tmem_ld32(tbase + col0, v);
    uint32_t m = 0;
    const float4* b4 = reinterpret_cast<const float4*>(sm + DS_B3 + col0);
    const float4* w4 = reinterpret_cast<const float4*>(sm + DS_W4 + col0);
    const float4* wu = reinterpret_cast<const float4*>(sm + DS_WU + col0);
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
      const float4 b = b4[i4], wz = w4[i4], wv = wu[i4];
      const float a0 = v[4 * i4] + b.x, a1 = v[4 * i4 + 1] + b.y, a2 = v[4 * i4 + 2] + b.z, a3 = v[4 * i4 + 3] + b.w;
      m |= ((a0 > 0.f ? 1u : 0u) | (a1 > 0.f ? 2u : 0u) | (a2 > 0.f ? 4u : 0u) | (a3 > 0.f ? 8u : 0u)) << (4 * i4);
      const float h0 = fmaxf(a0, 0.f), h1 = fmaxf(a1, 0.f), h2 = fmaxf(a2, 0.f), h3 = fmaxf(a3, 0.f);
      zu[0] = fmaf(wz.x, h0, zu[0]); z1 = fmaf(wz.y, h1, z1); zu[0] = fmaf(wz.z, h2, zu[0]); z1 = fmaf(wz.w, h3, z1);
      zu[1] = fmaf(wv.x, h0, zu[1]); u1 = fmaf(wv.y, h1, u1); zu[1] = fmaf(wv.z, h2, zu[1]); u1 = fmaf(wv.w, h3, u1);
    }
    c.mask[3][jj] = m;
  }
  zu[0] += z1; zu[1] += u1;
  exchange(c, zu, 2);
  z = zu[0] + sm[DS_B4]; u = zu[1] + sm[DS_B4 + 1];
}

// reverse pass: seeds on z and u -> d/d(xyz) in network units (both threads of a row return the full sum)
// HAS_U = false: the caller's seed on the std head is identically zero (the tracker detaches std, tracker.py:191), so its
// weights are neither loaded nor multiplied
template <bool HAS_U = true>
__device__ __forceinline__ void backward(Ctx& c, float seed_z, float seed_u, float g[3]) {
  const float* sm = reinterpret_cast<const float*>(c.sm + SM_SMALL);
  const int row = c.row;
  float ga[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int jj = 0; jj < CW / 32; ++jj) {
    const int col0 = CW * c.part + 32 * jj;
    float d[32];
    const uint32_t m = c.mask[3][jj];
    const float4* w4 = reinterpret_cast<const float4*>(sm + DS_W4 + col0);
    const float4* wu = reinterpret_cast<const float4*>(sm + DS_WU + col0);
    const float4* t4 = reinterpret_cast<const float4*>(sm + DS_W3X + 3 * col0);
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
      const float4 wz = w4[i4];
      const float4 wv = HAS_U ? wu[i4] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 ta = t4[3 * i4], tb = t4[3 * i4 + 1], tc_ = t4[3 * i4 + 2];     // taps of 4 consecutive units, 3 floats each
      const uint32_t mm = m >> (4 * i4);
      const float d0 = (mm & 1u) ? (HAS_U ? fmaf(seed_z, wz.x, seed_u * wv.x) : seed_z * wz.x) : 0.f;
      const float d1 = (mm & 2u) ? (HAS_U ? fmaf(seed_z, wz.y, seed_u * wv.y) : seed_z * wz.y) : 0.f;
      const float d2 = (mm & 4u) ? (HAS_U ? fmaf(seed_z, wz.z, seed_u * wv.z) : seed_z * wz.z) : 0.f;
      const float d3 = (mm & 8u) ? (HAS_U ? fmaf(seed_z, wz.w, seed_u * wv.w) : seed_z * wz.w) : 0.f;
      d[4 * i4] = d0; d[4 * i4 + 1] = d1; d[4 * i4 + 2] = d2; d[4 * i4 + 3] = d3;
      ga[0] = fmaf(ta.x, d0, ga[0]); ga[1] = fmaf(ta.y, d0, ga[1]); ga[2] = fmaf(ta.z, d0, ga[2]);
      ga[0] = fmaf(ta.w, d1, ga[0]); ga[1] = fmaf(tb.x, d1, ga[1]); ga[2] = fmaf(tb.y, d1, ga[2]);
      ga[0] = fmaf(tb.z, d2, ga[0]); ga[1] = fmaf(tb.w, d2, ga[1]); ga[2] = fmaf(tc_.x, d2, ga[2]);
      ga[0] = fmaf(tc_.y, d3, ga[0]); ga[1] = fmaf(tc_.z, d3, ga[1]); ga[2] = fmaf(tc_.w, d3, ga[2]);
    }
    store_cols32(c.sm + c.a_off, row, col0, d);
  }
  TC_LAYER((issue<8, 128, true, 128>(c, c.sa + c.a_off, c.sa + IMG_W3A, false)));   // delta2 = delta3 * W3[:, :96] (cols 96.. are 0)
  epi_bwd<96>(c, 2);
  TC_LAYER((issue<6, 128, true, 96>(c, c.sa + c.a_off, c.sa + IMG_W2, false)));     // delta1 = delta2 * W2
  epi_bwd<128>(c, 1);
  TC_LAYER((issue<8, 128, true, 128>(c, c.sa + c.a_off, c.sa + IMG_W1, false)));    // delta0 = delta1 * W1 (masked below)
  const uint32_t tbase = c.tmem + ((uint32_t)(row & ~31) << 16);
#pragma unroll
  for (int jj = 0; jj < CW / 32; ++jj) {
    const int col0 = CW * c.part + 32 * jj;
    float v[32];
    tmem_ld32(tbase + col0, v);
    const uint32_t m = c.mask[0][jj];
    const float4* t4 = reinterpret_cast<const float4*>(sm + DS_W0X + 3 * col0);
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
      const float4 ta = t4[3 * i4], tb = t4[3 * i4 + 1], tc_ = t4[3 * i4 + 2];
      const uint32_t mm = m >> (4 * i4);
      const float d0 = (mm & 1u) ? v[4 * i4] : 0.f, d1 = (mm & 2u) ? v[4 * i4 + 1] : 0.f;
      const float d2 = (mm & 4u) ? v[4 * i4 + 2] : 0.f, d3 = (mm & 8u) ? v[4 * i4 + 3] : 0.f;
      ga[0] = fmaf(ta.x, d0, ga[0]); ga[1] = fmaf(ta.y, d0, ga[1]); ga[2] = fmaf(ta.z, d0, ga[2]);
      ga[0] = fmaf(ta.w, d1, ga[0]); ga[1] = fmaf(tb.x, d1, ga[1]); ga[2] = fmaf(tb.y, d1, ga[2]);
      ga[0] = fmaf(tb.z, d2, ga[0]); ga[1] = fmaf(tb.w, d2, ga[1]); ga[2] = fmaf(tc_.x, d2, ga[2]);
      ga[0] = fmaf(tc_.y, d3, ga[0]); ga[1] = fmaf(tc_.z, d3, ga[1]); ga[2] = fmaf(tc_.w, d3, ga[2]);
    }
  }
  exchange(c, ga, 3);
  g[0] = ga[0]; g[1] = ga[1]; g[2] = ga[2];
}

extern __shared__ unsigned char tc_smem_raw[];

// CTA prologue: align, barriers, TMEM, one bulk load of the whole network
__device__ __forceinline__ void prologue(Ctx& c, const void* blob) {
  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
  c.sm = tc_smem_raw + pad;
  c.sa = raw + pad;
  c.phase = 0;
  c.grp = threadIdx.x / GT;
  c.part = (threadIdx.x % GT) / T;
  c.row = threadIdx.x % T;
  c.a_off = SM_A + c.grp * SM_TILE_BYTES;
  c.x_off = c.a_off + SM_XOFF;
  const uint32_t wbar = c.sa + SM_BAR, slot = c.sa + SM_BAR + 48;
  c.mma_bar = c.sa + SM_BAR + 8 + 8 * c.grp;
  if (threadIdx.x == 0) {
    mbar_init(wbar, 1);
    for (int g = 0; g < GROUPS; ++g) mbar_init(c.sa + SM_BAR + 8 + 8 * g, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_alloc(slot, TMEM_COLS);
  if (threadIdx.x == 0) {
    mbar_expect_tx(wbar, BLOB_BYTES);
    const char* src = reinterpret_cast<const char*>(blob);
    for (int off = 0; off < BLOB_BYTES; off += 16128) {   // 8 chunks of 16128 B (multiple of 16)
      const int n = min(16128, BLOB_BYTES - off);
      bulk_g2s(c.sa + off, src + off, (uint32_t)n, wbar);
    }
  }
  // zero the activation / input tiles once (unused K columns must be finite)
  for (int i = threadIdx.x; i < GROUPS * SM_TILE_BYTES / 16; i += CTA_T) reinterpret_cast<uint4*>(c.sm + SM_A)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem_base = *reinterpret_cast<volatile uint32_t*>(c.sm + SM_BAR + 48);
  c.tmem = c.tmem_base + 128u * c.grp;
  c.wbar = wbar;      // waited on in forward(): the weight load overlaps the first tile's map lookups and input staging
}

__device__ __forceinline__ void epilogue_free(Ctx& c) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(c.tmem_base, TMEM_COLS);
}

// The NPART threads of a row split the 29 Gauss-Newton sums (21 upper-triangle J J^T, 6 J r, energy, count): part p keeps
// packed entries [8p, 8p+8) in acc[8].
__device__ __forceinline__ void hg_accumulate_case(float* acc8, const float* J, float r, float w, bool with_J, int which) {
  switch (which) {   // static indices inside each case
    case 0:
      if (with_J) { acc8[0] += w * J[0] * J[0]; acc8[1] += w * J[0] * J[1]; acc8[2] += w * J[0] * J[2]; acc8[3] += w * J[0] * J[3];
                    acc8[4] += w * J[0] * J[4]; acc8[5] += w * J[0] * J[5]; acc8[6] += w * J[1] * J[1]; acc8[7] += w * J[1] * J[2]; }
      break;
    case 1:
      if (with_J) { acc8[0] += w * J[1] * J[3]; acc8[1] += w * J[1] * J[4]; acc8[2] += w * J[1] * J[5]; acc8[3] += w * J[2] * J[2];
                    acc8[4] += w * J[2] * J[3]; acc8[5] += w * J[2] * J[4]; acc8[6] += w * J[2] * J[5]; acc8[7] += w * J[3] * J[3]; }
      break;
    case 2:
      if (with_J) { acc8[0] += w * J[3] * J[4]; acc8[1] += w * J[3] * J[5]; acc8[2] += w * J[4] * J[4]; acc8[3] += w * J[4] * J[5];
                    acc8[4] += w * J[5] * J[5]; acc8[5] += w * r * J[0]; acc8[6] += w * r * J[1]; acc8[7] += w * r * J[2]; }
      break;
    default:
      if (with_J) { acc8[0] += w * r * J[3]; acc8[1] += w * r * J[4]; acc8[2] += w * r * J[5]; }
      acc8[3] += w * r * r;
      acc8[4] += 1.0f;
      break;
  }
}

constexpr int HG_PER_THREAD = 32 / NPART;   // packed entries kept by one thread (4 / NPART cases of 8)
__device__ __forceinline__ void hg_accumulate_part(float* acc, const float* J, float r, float w, bool with_J, int part) {
#pragma unroll
  for (int sub = 0; sub < 4 / NPART; ++sub) hg_accumulate_case(acc + 8 * sub, J, r, w, with_J, part * (4 / NPART) + sub);
}

// CTA reduction of the per-part partial sums into the packed 29 doubles (scratch: idle tile buffers)
__device__ __forceinline__ void block_reduce_parts(const float* acc8, int part, double* out, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < HG_PER_THREAD; ++k) {
    const double d = warp_sum((double)acc8[k]);
    if (lane == 0) red[w * HG_PER_THREAD + k] = d;
  }
  __syncthreads();
  if (threadIdx.x < 29) {
    const int p = threadIdx.x / HG_PER_THREAD, k = threadIdx.x % HG_PER_THREAD;
    double sacc = 0.0;
    for (int ww = 0; ww < CTA_T / 32; ++ww)
      if (((ww * 32) % GT) / T == p) sacc += red[ww * HG_PER_THREAD + k];
    if (sacc != 0.0) atomicAdd(&out[threadIdx.x], sacc);
  }
}

// ------------------------------------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(CTA_T, 1) explicit_kernel(const float* __restrict__ x, int n, const void* __restrict__ blob,
                                                        float* __restrict__ sdf, float* __restrict__ std_) {
  Ctx c;
  prologue(c, blob);
  for (long long rnd = 0, tile; (tile = tile_of(c, n, rnd)) >= 0; ++rnd) {
    const int i = (int)(tile * T) + c.row;
    {
      const float* xr = x + (size_t)(i < n ? i : 0) * 32;
      const float rel3[3] = {i < n ? xr[29] : 0.f, i < n ? xr[30] : 0.f, i < n ? xr[31] : 0.f};
      store_input(c, xr, rel3, i < n);
    }
    float z, u;
    forward(c, z, u);
    if (i < n && c.part == 0) { sdf[i] = tanhf(z); std_[i] = 0.05f + 0.5f * softplus_torch(u); }
  }
  epilogue_free(c);
}

__global__ void __launch_bounds__(CTA_T, 1) get_sdf_kernel(MapDev M, const float* __restrict__ xyz, int n, const int64_t* __restrict__ indexer,
                                                       const float* __restrict__ latents, const float* __restrict__ obs_count,
                                                       const void* __restrict__ blob, float* __restrict__ sdf, float* __restrict__ std_,
                                                       uint8_t* __restrict__ valid_out, const float* __restrict__ g_sdf,
                                                       const float* __restrict__ g_std, float* __restrict__ grad_xyz) {
  Ctx c;
  prologue(c, blob);
  for (long long rnd = 0, tile; (tile = tile_of(c, n, rnd)) >= 0; ++rnd) {
    const int i = (int)(tile * T) + c.row;
    bool valid = false;
    long long slot = 0;
    float rel[3] = {0.f, 0.f, 0.f};
    if (i < n) valid = map_lookup(M, xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], indexer, obs_count, slot, rel);
    store_input(c, latents + (valid ? slot : 0) * DFB_LATENT_DIM, rel, valid);
    float z, u;
    forward(c, z, u);
    const float s = tanhf(z), sd = 0.05f + 0.5f * softplus_torch(u);
    if (i < n && c.part == 0) {
      valid_out[i] = valid ? 1 : 0;
      if (sdf) sdf[i] = valid ? s : 0.f;
      if (std_) std_[i] = valid ? sd : 0.f;
    }
    if (grad_xyz) {
      float gs = 0.f, gu = 0.f;
      if (valid) {
        gs = g_sdf ? g_sdf[i] * (1.0f - s * s) : 0.f;
        gu = g_std ? g_std[i] * 0.5f * softplus_grad(u) : 0.f;
      }
      float g[3];
      backward(c, gs, gu, g);
      if (i < n && c.part == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) grad_xyz[3 * (size_t)i + a] = valid ? div_vs(g[a], M.vs, M.inv_vs, M.div_mode) : 0.f;
      }
    }
  }
  epilogue_free(c);
}

// SDF term over this group's tiles: transform -> map lookup -> decoder forward -> reverse pass -> Jacobian -> robust weight ->
// per-thread partial sums (acc[HG_PER_THREAD]).  tracker.py:179-223.
__device__ __forceinline__ void sdf_tiles(Ctx& c, const MapDev& M, const PoseDev& P, const float* __restrict__ obs, int n,
                                          const int64_t* __restrict__ indexer, const float* __restrict__ latents,
                                          const float* __restrict__ obs_count, int robust, float robust_k, int with_J, float* acc) {
  for (long long rnd = 0, tile; (tile = tile_of(c, n, rnd)) >= 0; ++rnd) {
    const int i = (int)(tile * T) + c.row;
    PROF_MARK(c);                                // tile start
    bool valid = false;
    long long slot = 0;
    float rel[3] = {0.f, 0.f, 0.f}, pc[3] = {0.f, 0.f, 0.f};
    if (i < n) {
      pc[0] = obs[3 * (size_t)i]; pc[1] = obs[3 * (size_t)i + 1]; pc[2] = obs[3 * (size_t)i + 2];
      float pw[3];
      xform(P.Rt, P.tt, pc[0], pc[1], pc[2], pw);
      valid = map_lookup(M, pw[0], pw[1], pw[2], indexer, obs_count, slot, rel);
    }
    store_input(c, latents + (valid ? slot : 0) * DFB_LATENT_DIM, rel, valid);
    PROF_MARK(c);                                // lookup + input staged
    float z, u;
    forward(c, z, u);
    const float s = tanhf(z), sd = 0.05f + 0.5f * softplus_torch(u);
    const float r = s / sd;
    PROF_MARK(c);
    float J[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (with_J) {
      float g[3];
      backward<false>(c, valid ? (1.0f - s * s) / sd : 0.f, 0.f, g);
      if (valid) {
        float gw[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) gw[a] = div_vs(g[a], M.vs, M.inv_vs, M.div_mode);
        sdf_jacobian(P, gw, pc, J);
      }
    }
    if (valid) hg_accumulate_part(acc, J, r, robust_w(r, robust, robust_k), with_J != 0, c.part);
    PROF_MARK(c);                                // tile end
  }
}

__global__ void __launch_bounds__(CTA_T, 1) sdf_hg_kernel(MapDev M, PoseDev P, const float* __restrict__ obs, int n,
                                                      const int64_t* __restrict__ indexer, const float* __restrict__ latents,
                                                      const float* __restrict__ obs_count, const void* __restrict__ blob, int robust,
                                                      float robust_k, int with_J, double* __restrict__ packed) {
  Ctx c;
#ifdef DFB_TC_PROFILE
  c.grp = threadIdx.x / GT; c.part = (threadIdx.x % GT) / T; c.row = threadIdx.x % T;
#endif
  PROF_MARK(c);                                  // kernel start
  prologue(c, blob);
  PROF_MARK(c);                                  // prologue done
  float acc[HG_PER_THREAD];
#pragma unroll
  for (int k = 0; k < HG_PER_THREAD; ++k) acc[k] = 0.f;
  sdf_tiles(c, M, P, obs, n, indexer, latents, obs_count, robust, robust_k, with_J, acc);
  epilogue_free(c);
  PROF_MARK(c);
  block_reduce_parts(acc, c.part, packed, reinterpret_cast<double*>(c.sm + SM_A));
  PROF_MARK(c);                                  // kernel end
}

// One Gauss-Newton evaluation in ONE launch (gauss_newton.cu): the SDF term on the tensor cores; the photometric term's
// pixels are handed out in chunks to whichever tile group has run out of tiles (in the last round half of the groups
// have, and their SM's tensor pipe is busy with the sibling group anyway); both terms' sums go to GnShared with one
// atomic per value per block; the last block to finish runs the step (solve, pose update, record for the host).
constexpr int RGB_PIX = 2;                       // pixels per thread per chunk (chunk = GT * RGB_PIX pixels)
__global__ void __launch_bounds__(CTA_T, 1) gn_eval_kernel(MapDev M, const float* __restrict__ obs, int n, const int* __restrict__ n_dev,
                                                       const int64_t* __restrict__ indexer,
                                                       const float* __restrict__ latents, const float* __restrict__ obs_count,
                                                       const void* __restrict__ blob, int robust, float robust_k, int with_J, RgbDev R,
                                                       GnShared* gs, int gi, gn::StepArgs sa) {
  if (gs->done[gi]) {                            // a launch queued ahead of a group that has ended: only the record is owed
    if (blockIdx.x == 0 && threadIdx.x < 32) gn::skip_record(gs, sa);
    return;
  }
  const PoseDev P = *reinterpret_cast<const PoseDev*>(gs->pose_sdf);
  if (n_dev) n = min(n, max(*n_dev, 0));         // row count still on the device (dfb_preprocess_frame's output)
  Ctx c;
#ifdef DFB_TC_PROFILE
  c.grp = threadIdx.x / GT; c.part = (threadIdx.x % GT) / T; c.row = threadIdx.x % T;
#endif
  PROF_MARK(c);                                  // kernel start
  prologue(c, blob);
  PROF_MARK(c);                                  // prologue done
  float acc[HG_PER_THREAD];
#pragma unroll
  for (int k = 0; k < HG_PER_THREAD; ++k) acc[k] = 0.f;
  sdf_tiles(c, M, P, obs, n, indexer, latents, obs_count, robust, robust_k, with_J, acc);
  PROF_MARK(c);                                  // tiles done
  float racc[29];
#pragma unroll
  for (int k = 0; k < 29; ++k) racc[k] = 0.f;
  if (R.on) {
#pragma unroll
    for (int i = 0; i < 9; ++i) R.P.k[i] = gs->krk[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) R.P.kt[i] = gs->kt[i];
    const int tg = threadIdx.x % GT, npx = R.H * R.W;
    volatile int* slotp = reinterpret_cast<volatile int*>(c.sm + SM_BAR + 32 + 4 * c.grp);
    for (;;) {
      if (tg == 0) *slotp = atomicAdd(&gs->rgb_cursor, 1);
      group_sync(c.grp);
      const int base = *slotp * (GT * RGB_PIX);
      group_sync(c.grp);
      if (base >= npx) break;
#pragma unroll
      for (int e = 0; e < RGB_PIX; ++e) {
        const int i = base + e * GT + tg;
        if (i < npx) {
          const int v = i / R.W, u = i - v * R.W;
          float f, J[6];
          if (rgb_pixel(R.prev_I, R.prev_D, R.cur_I, R.cur_D, R.cur_G, R.H, R.W, R.P, v, u, with_J != 0, f, J)) {
            if (with_J) {
#pragma unroll
              for (int a = 0; a < 6; ++a) J[a] = -J[a];   // tracker.py:162
            }
            hg_accumulate(racc, J, f, robust_w(f, R.robust, R.robust_k), with_J != 0);
          }
        }
      }
    }
  }
  PROF_MARK(c);                                  // pixels done
  epilogue_free(c);                              // block-wide barrier; the tile buffers are scratch from here on
  PROF_MARK(c);                                  // all groups done
  // Block sums through shared memory (the tile buffers are idle): every thread writes its partial sums column-wise, then 8
  // threads per value add their share in float64 and meet with three shuffles.  (A shuffle-only reduction of 45 doubles
  // per thread is bound by the SM's one-warp-per-clock shuffle unit: measured 6.5 us here.)
  {
    constexpr int NS = HG_PER_THREAD * NPART;                     // 32 SDF rows (29 used); a row holds the GROUPS * T threads of one part
    constexpr int LDS_ = GROUPS * T + 8, LDR = CTA_T + 8;         // padded rows: the 4 values a warp reads land in distinct banks
    static_assert((NS * LDS_ + 29 * LDR) * 4 + gn::STEP_SCRATCH_BYTES <= GROUPS * SM_TILE_BYTES, "reduction scratch exceeds the tile buffers");
    float* bs = reinterpret_cast<float*>(c.sm + SM_A);
    float* br = bs + NS * LDS_;
#pragma unroll
    for (int k = 0; k < HG_PER_THREAD; ++k) bs[(c.part * HG_PER_THREAD + k) * LDS_ + c.grp * T + c.row] = acc[k];
    if (R.on) {
#pragma unroll
      for (int k = 0; k < 29; ++k) br[k * LDR + threadIdx.x] = racc[k];
    }
    __syncthreads();
    const int v = threadIdx.x >> 3, seg = threadIdx.x & 7;       // value, 1/8 of its row
    double sacc = 0.0;
    if (v < NS) {
      const float* rowp = bs + v * LDS_ + seg;
      double s1 = 0.0;
#pragma unroll 8
      for (int j = 0; j < GROUPS * T / 8; j += 2) { sacc += (double)rowp[8 * j]; s1 += (double)rowp[8 * j + 8]; }
      sacc += s1;
    } else if (v < NS + 29 && R.on) {
      const float* rowp = br + (v - NS) * LDR + seg;
      double s1 = 0.0;
#pragma unroll 8
      for (int j = 0; j < CTA_T / 8; j += 2) { sacc += (double)rowp[8 * j]; s1 += (double)rowp[8 * j + 8]; }
      sacc += s1;
    }
    sacc += __shfl_xor_sync(0xffffffffu, sacc, 1); sacc += __shfl_xor_sync(0xffffffffu, sacc, 2); sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
    if (seg == 0 && sacc != 0.0) {
      if (v < 29) atomicAdd(&gs->sums[0][v], sacc);
      else if (v >= NS && v < NS + 29) atomicAdd(&gs->sums[1][v - NS], sacc);
    }
  }
  __threadfence();
  __syncthreads();
  PROF_MARK(c);                                  // sums out
  gn::tail_step(gs, sa, c.sm + GROUPS * SM_TILE_BYTES + SM_A - gn::STEP_SCRATCH_BYTES, reinterpret_cast<int*>(c.sm + SM_BAR + 56));
  PROF_MARK(c);                                  // kernel end (block 0; the step runs in whichever block finishes last)
}

__global__ void __launch_bounds__(CTA_T, 1) cube_low_kernel(const float* __restrict__ latents, const int64_t* __restrict__ occ, int B, int r,
                                                        float vsize, float a, const void* __restrict__ blob, float* __restrict__ low_sdf,
                                                        float* __restrict__ low_std) {
  Ctx c;
  prologue(c, blob);
  const long long r3 = (long long)r * r * r, n = (long long)B * r3;
  for (long long rnd = 0, tile; (tile = tile_of(c, n, rnd)) >= 0; ++rnd) {
    const long long i = tile * T + c.row;
    const bool valid = i < n;
    float rel[3] = {0.f, 0.f, 0.f};
    long long slot = 0;
    if (valid) {
      const int b = (int)(i / r3), cc = (int)(i - (long long)b * r3);
      rel[0] = lattice(cc / (r * r), vsize, a); rel[1] = lattice((cc / r) % r, vsize, a); rel[2] = lattice(cc % r, vsize, a);
      slot = occ[b];
    }
    store_input(c, latents + slot * DFB_LATENT_DIM, rel, valid);
    float z, u;
    forward(c, z, u);
    if (valid && c.part == 0) { low_sdf[i] = tanhf(z); low_std[i] = 0.05f + 0.5f * softplus_torch(u); }
  }
  epilogue_free(c);
}

__global__ void __launch_bounds__(CTA_T, 1) cube_refine_kernel(const float* __restrict__ latents, const int64_t* __restrict__ occ, int r, float vsize,
                                                           float a, const void* __restrict__ blob, const int* __restrict__ refine_count,
                                                           const long long* __restrict__ refine_list, float* __restrict__ cube_sdf,
                                                           float* __restrict__ cube_std) {
  Ctx c;
  prologue(c, blob);
  const int R = 2 * r;
  const long long R3 = (long long)R * R * R;
  const int n = *refine_count;
  for (long long rnd = 0, tile; (tile = tile_of(c, n, rnd)) >= 0; ++rnd) {
    const int t = (int)(tile * T) + c.row;
    const bool valid = t < n;
    float rel[3] = {0.f, 0.f, 0.f};
    long long slot = 0, i = 0;
    if (valid) {
      i = refine_list[t];
      const int b = (int)(i / R3), cc = (int)(i - (long long)b * R3);
      rel[0] = lattice(cc / (R * R), vsize, a); rel[1] = lattice((cc / R) % R, vsize, a); rel[2] = lattice(cc % R, vsize, a);
      slot = occ[b];
    }
    store_input(c, latents + slot * DFB_LATENT_DIM, rel, valid);
    float z, u;
    forward(c, z, u);
    if (valid && c.part == 0) { cube_sdf[i] = -tanhf(z); cube_std[i] = 0.05f + 0.5f * softplus_torch(u); }
  }
  epilogue_free(c);
}

template <typename K>
static int prep(K kernel) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_ALLOC);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(tc): %s", cudaGetErrorString(e)); return DFB_E_CUDA; }
  return DFB_OK;
}
static int grid_for(long long n) { return (int)std::min<long long>(div_up(n, T), (long long)sm_count()); }

}  // namespace tc

#ifdef DFB_TC_PROFILE
extern "C" int dfb_debug_read_step_prof(unsigned long long* h_out) {   // marks of the last step run inside gn_eval_kernel
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h_out, gn::g_step_prof, sizeof(unsigned long long) * 16);
  return 0;
}
extern "C" int dfb_debug_read_prof(unsigned long long* h_out, int* h_n) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h_n, tc::g_prof_n, sizeof(int));
  cudaMemcpyFromSymbol(h_out, tc::g_prof, sizeof(unsigned long long) * 256);
  int zero = 0;
  cudaMemcpyToSymbol(tc::g_prof_n, &zero, sizeof(int));
  return 0;
}
#endif

int tc_decoder_explicit(const float* x, int n, const void* blob, float* sdf, float* std_, cudaStream_t s) {
  int rc = tc::prep(tc::explicit_kernel);
  if (rc) return rc;
  tc::explicit_kernel<<<tc::grid_for(n), tc::CTA_T, tc::SM_ALLOC, s>>>(x, n, blob, sdf, std_);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int tc_get_sdf(const MapDev& M, const float* xyz, int n, const int64_t* indexer, const float* latents, const float* obs_count,
               const void* blob, float* sdf, float* std_, uint8_t* valid, const float* g_sdf, const float* g_std, float* grad_xyz,
               cudaStream_t s) {
  int rc = tc::prep(tc::get_sdf_kernel);
  if (rc) return rc;
  tc::get_sdf_kernel<<<tc::grid_for(n), tc::CTA_T, tc::SM_ALLOC, s>>>(M, xyz, n, indexer, latents, obs_count, blob, sdf, std_, valid, g_sdf,
                                                                    g_std, grad_xyz);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int tc_gn_eval(const MapDev& M, const float* obs, int n, const int* n_dev, const int64_t* indexer, const float* latents, const float* obs_count,
               const void* blob, int robust, float robust_k, int with_J, const RgbDev& R, GnShared* gs, int gi, const gn::StepArgs& sa,
               cudaStream_t s) {
  int rc = tc::prep(tc::gn_eval_kernel);
  if (rc) return rc;
  const int grid = R.on ? sm_count() : std::max(1, tc::grid_for(n));     // every SM takes photometric chunks
  tc::gn_eval_kernel<<<grid, tc::CTA_T, tc::SM_ALLOC, s>>>(M, obs, n, n_dev, indexer, latents, obs_count, blob, robust, robust_k, with_J, R, gs, gi, sa);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int tc_sdf_hg(const MapDev& M, const PoseDev& P, const float* obs, int n, const int64_t* indexer, const float* latents,
              const float* obs_count, const void* blob, int robust, float robust_k, int with_J, double* packed, cudaStream_t s) {
  int rc = tc::prep(tc::sdf_hg_kernel);
  if (rc) return rc;
  tc::sdf_hg_kernel<<<tc::grid_for(n), tc::CTA_T, tc::SM_ALLOC, s>>>(M, P, obs, n, indexer, latents, obs_count, blob, robust, robust_k, with_J,
                                                                   packed);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int tc_cube_low(const float* latents, const int64_t* occ, int B, int r, float vsize, float a, const void* blob, float* low_sdf,
                float* low_std, cudaStream_t s) {
  int rc = tc::prep(tc::cube_low_kernel);
  if (rc) return rc;
  tc::cube_low_kernel<<<tc::grid_for((long long)B * r * r * r), tc::CTA_T, tc::SM_ALLOC, s>>>(latents, occ, B, r, vsize, a, blob, low_sdf, low_std);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int tc_cube_refine(const float* latents, const int64_t* occ, int r, float vsize, float a, const void* blob, const int* refine_count,
                   const long long* refine_list, float* cube_sdf, float* cube_std, cudaStream_t s) {
  int rc = tc::prep(tc::cube_refine_kernel);
  if (rc) return rc;
  tc::cube_refine_kernel<<<sm_count(), tc::CTA_T, tc::SM_ALLOC, s>>>(latents, occ, r, vsize, a, blob, refine_count, refine_list, cube_sdf, cube_std);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

}  // namespace dfb
