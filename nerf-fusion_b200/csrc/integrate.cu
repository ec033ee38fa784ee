// Stages 2+3: integrate_keyframe (system/map.py:341-453, do_optimize=False) as two asynchronous phases with one
// 4-byte host read between them (the number of new voxels, so the host can grow its arrays like map.py:263-285).
//
// The reference does this with ~40 torch ops, >=5 torch.unique sorts, a 2 MB + 0.5 MB full-grid memset per call and
// per-element global atomics from 29-thread blocks (indexing.cu:99-106).  Here:
//   * per-cell counts and the allocation set live in persistent zero-invariant grids (int32 count, bitmap);
//   * "sorted unique new voxel ids -> consecutive slots" (map.py:384-388, 310-319) is a popcount prefix sum over
//     the bitmap, which yields ascending-id slot numbering without any sort;
//   * (point, offset) samples are appended with warp-aggregated atomics, the encoder MLP runs on the compact list
//     and its outputs are scatter-added per voxel; a final kernel applies the running mean (map.py:449-452).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace dfb {

struct MapI {
  int nx, ny, nz;
  float bx, by, bz, vs, inv_vs;
  int div_mode, prune_min;
  float enc_th;
};

static MapI to_mapi(const dfb_map_params* p) {
  MapI m;
  m.nx = p->nx; m.ny = p->ny; m.nz = p->nz;
  m.bx = p->bound_min[0]; m.by = p->bound_min[1]; m.bz = p->bound_min[2];
  m.vs = p->voxel_size; m.inv_vs = 1.0f / p->voxel_size;
  m.div_mode = p->div_mode; m.prune_min = p->prune_min_vox_obs; m.enc_th = p->encoder_count_th;
  return m;
}

struct Sample {          // 32 bytes
  int slot;
  float rel[3];
  float nrm[3];
  int pad;
};

struct IntegrateWs {
  float* xn;        // n*3
  int* gid;         // n   (-1 = outside the grid)
  int* word_rank;   // n_words + 1
  int* bsums;
  int* counters;    // [0] n_samples, [1] n_touched
  Sample* samples;  // 8n
};

static void integrate_ws_layout(Arena& a, int n, long long n_cells, IntegrateWs* w) {
  const long long n_words = (n_cells + 31) / 32;
  w->xn = a.take<float>((size_t)n * 3 + 4);
  w->gid = a.take<int>((size_t)n + 1);
  w->word_rank = a.take<int>((size_t)n_words + 2);
  w->bsums = a.take<int>((size_t)n_words / 2048 + 8);
  w->counters = a.take<int>(8);
  w->samples = a.take<Sample>((size_t)n * 8 + 1);
}

__device__ __forceinline__ int lin_id(const MapI& M, int x, int y, int z) { return z + M.nz * y + M.nz * M.ny * x; }

// map.py:367-370
__global__ void __launch_bounds__(256) ik_ids_kernel(MapI M, const float* __restrict__ xyz, int n, float* __restrict__ xn,
                                                     int* __restrict__ gid, int* __restrict__ grid_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = div_vs(__fsub_rn(xyz[3 * (size_t)i], M.bx), M.vs, M.inv_vs, M.div_mode);
  const float y = div_vs(__fsub_rn(xyz[3 * (size_t)i + 1], M.by), M.vs, M.inv_vs, M.div_mode);
  const float z = div_vs(__fsub_rn(xyz[3 * (size_t)i + 2], M.bz), M.vs, M.inv_vs, M.div_mode);
  xn[3 * (size_t)i] = x; xn[3 * (size_t)i + 1] = y; xn[3 * (size_t)i + 2] = z;
  const float cx = ceilf(x) - 1.f, cy = ceilf(y) - 1.f, cz = ceilf(z) - 1.f;
  int g = -1;
  if (cx >= 0.f && cx < (float)M.nx && cy >= 0.f && cy < (float)M.ny && cz >= 0.f && cz < (float)M.nz) {
    g = lin_id(M, (int)cx, (int)cy, (int)cz);
    atomicAdd(&grid_count[g], 1);
  }
  gid[i] = g;
}

// map.py:373-388 (mask and allocation set).  A voxel enters the set iff it is unallocated and is the home voxel of a
// kept point or a clamped face neighbour of such a voxel.
__global__ void __launch_bounds__(256) ik_mask_mark_kernel(MapI M, int n, const int* __restrict__ gid,
                                                           const int* __restrict__ grid_count, const int64_t* __restrict__ indexer,
                                                           uint8_t* __restrict__ mask, uint32_t* __restrict__ bits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int g = gid[i];
  const bool keep = g >= 0 && (M.prune_min <= 0 || grid_count[g] > M.prune_min);
  mask[i] = keep ? 1 : 0;
  if (!keep || indexer[g] != -1) return;
  // (the bit of g may already have been set by a NEIGHBOUR's dilation, so it says nothing about g's own neighbours:
  //  every kept point marks all seven cells; the plain read filters most of the redundant atomics)
  if (!(bits[g >> 5] & (1u << (g & 31)))) atomicOr(&bits[g >> 5], 1u << (g & 31));
  const int x = g / (M.ny * M.nz), y = (g / M.nz) % M.ny, z = g % M.nz;
  const int nb[6] = {lin_id(M, max(x - 1, 0), y, z), lin_id(M, min(x + 1, M.nx - 1), y, z),
                     lin_id(M, x, max(y - 1, 0), z), lin_id(M, x, min(y + 1, M.ny - 1), z),
                     lin_id(M, x, y, max(z - 1, 0)), lin_id(M, x, y, min(z + 1, M.nz - 1))};
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (indexer[nb[k]] == -1) atomicOr(&bits[nb[k] >> 5], 1u << (nb[k] & 31));
}

__global__ void __launch_bounds__(256) ik_reset_counts_kernel(int n, const int* __restrict__ gid, int* __restrict__ grid_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int g = gid[i];
  if (g >= 0) grid_count[g] = 0;
}

// map.py:310-319 + 283: slots n_occupied.. in ascending voxel id order.
__global__ void __launch_bounds__(256) ik_assign_kernel(int n_words, uint32_t* __restrict__ bits, const int* __restrict__ word_rank,
                                                        long long n_occupied, long long capacity, int64_t* __restrict__ indexer,
                                                        int64_t* __restrict__ pos, float* __restrict__ latents,
                                                        float* __restrict__ obs_count) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  uint32_t word = bits[w];
  if (!word) return;
  long long slot = n_occupied + word_rank[w];
  while (word) {
    const int b = __ffs(word) - 1;
    word &= word - 1;
    if (slot < capacity) {
      const long long id = (long long)w * 32 + b;
      indexer[id] = slot;
      pos[slot] = id;
      obs_count[slot] = 0.f;
      for (int l = 0; l < DFB_LATENT_DIM; ++l) latents[slot * DFB_LATENT_DIM + l] = 0.f;
    }
    ++slot;
  }
  bits[w] = 0u;
}

__device__ __forceinline__ bool is_candidate(const MapI& M, int id, const int64_t* __restrict__ indexer,
                                             const float* __restrict__ obs_count, long long& slot) {
  slot = indexer[id];
  return slot >= 0 && obs_count[slot] < M.enc_th;          // map.py:410-412
}

// map.py:390-436: focus prune + 8 offset samples, appended with one atomic per warp per offset.
__global__ void __launch_bounds__(256) ik_samples_kernel(MapI M, int n, const float* __restrict__ xn, const int* __restrict__ gid,
                                                         const uint8_t* __restrict__ mask, const float* __restrict__ normal,
                                                         const int64_t* __restrict__ indexer, const float* __restrict__ obs_count,
                                                         int* __restrict__ acc_n, int* __restrict__ touched, int* __restrict__ counters,
                                                         Sample* __restrict__ samples) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool focus = false;
  float x = 0.f, y = 0.f, z = 0.f;
  if (i < n && mask[i]) {
    const int g = gid[i];
    const int gx = g / (M.ny * M.nz), gy = (g / M.nz) % M.ny, gz = g % M.nz;
    long long s;
    // home voxel in dilate6(candidates)  <=>  home or one of its in-range face neighbours is a candidate
    focus = is_candidate(M, g, indexer, obs_count, s) ||
            (gx > 0 && is_candidate(M, lin_id(M, gx - 1, gy, gz), indexer, obs_count, s)) ||
            (gx < M.nx - 1 && is_candidate(M, lin_id(M, gx + 1, gy, gz), indexer, obs_count, s)) ||
            (gy > 0 && is_candidate(M, lin_id(M, gx, gy - 1, gz), indexer, obs_count, s)) ||
            (gy < M.ny - 1 && is_candidate(M, lin_id(M, gx, gy + 1, gz), indexer, obs_count, s)) ||
            (gz > 0 && is_candidate(M, lin_id(M, gx, gy, gz - 1), indexer, obs_count, s)) ||
            (gz < M.nz - 1 && is_candidate(M, lin_id(M, gx, gy, gz + 1), indexer, obs_count, s));
    x = xn[3 * (size_t)i]; y = xn[3 * (size_t)i + 1]; z = xn[3 * (size_t)i + 2];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {                               // map.py:186-189 order
    const float ox = (k & 4) ? 0.5f : -0.5f, oy = (k & 2) ? 0.5f : -0.5f, oz = (k & 1) ? 0.5f : -0.5f;
    bool take = false;
    long long slot = -1;
    float rx = 0.f, ry = 0.f, rz = 0.f;
    if (focus) {
      // ceil(xn + off) - 1, clamped (map.py:423-425)
      const float cx = fminf(fmaxf(ceilf(__fadd_rn(x, ox)) - 1.f, 0.f), (float)(M.nx - 1));
      const float cy = fminf(fmaxf(ceilf(__fadd_rn(y, oy)) - 1.f, 0.f), (float)(M.ny - 1));
      const float cz = fminf(fmaxf(ceilf(__fadd_rn(z, oz)) - 1.f, 0.f), (float)(M.nz - 1));
      take = is_candidate(M, lin_id(M, (int)cx, (int)cy, (int)cz), indexer, obs_count, slot);
      rx = __fsub_rn(__fsub_rn(x, cx), 0.5f);                 // map.py:426
      ry = __fsub_rn(__fsub_rn(y, cy), 0.5f);
      rz = __fsub_rn(__fsub_rn(z, cz), 0.5f);
    }
    const int p = warp_append(&counters[0], take);
    if (take) {
      Sample sm;
      sm.slot = (int)slot;
      sm.rel[0] = rx; sm.rel[1] = ry; sm.rel[2] = rz;
      sm.nrm[0] = normal[3 * (size_t)i]; sm.nrm[1] = normal[3 * (size_t)i + 1]; sm.nrm[2] = normal[3 * (size_t)i + 2];
      sm.pad = 0;
      samples[p] = sm;
      if (atomicAdd(&acc_n[slot], 1) == 0) touched[atomicAdd(&counters[1], 1)] = (int)slot;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// encoder MLP (di_encoder.py:26-30; BN folded on the host), one thread per sample
// ------------------------------------------------------------------------------------------------
constexpr int ENC_T = 128;
// blob layout (floats)
constexpr int EB_W0 = 0;                     // [32][8]  (6 inputs padded to 8)
constexpr int EB_B0 = EB_W0 + 32 * 8;        // [32]
constexpr int EB_W1 = EB_B0 + 32;            // [64][32]
constexpr int EB_B1 = EB_W1 + 64 * 32;       // [64]
constexpr int EB_W2 = EB_B1 + 64;            // [256/8][64][8]
constexpr int EB_B2 = EB_W2 + 256 * 64;      // [256]
constexpr int EB_W3 = EB_B2 + 256;           // [256][32] (29 outputs padded to 32)
constexpr int EB_B3 = EB_W3 + 256 * 32;      // [32]
constexpr int EB_TOTAL = EB_B3 + 32;

struct EncSmem {
  float w[EB_TOTAL];
  float h1[64 * ENC_T];
};

__device__ __forceinline__ void encoder_one(const EncSmem& S, const float in[6], float* __restrict__ h1col, float out[32]) {
  const float* W = S.w;
  float h0[32];
#pragma unroll
  for (int o = 0; o < 32; ++o) {
    const float4 a = *reinterpret_cast<const float4*>(W + EB_W0 + o * 8);
    const float4 b = *reinterpret_cast<const float4*>(W + EB_W0 + o * 8 + 4);
    float v = W[EB_B0 + o];
    v = fmaf(a.x, in[0], v); v = fmaf(a.y, in[1], v); v = fmaf(a.z, in[2], v);
    v = fmaf(a.w, in[3], v); v = fmaf(b.x, in[4], v); v = fmaf(b.y, in[5], v);
    h0[o] = fmaxf(v, 0.f);
  }
#pragma unroll 2
  for (int o = 0; o < 64; ++o) {
    float v = W[EB_B1 + o];
    const float4* wr = reinterpret_cast<const float4*>(W + EB_W1 + o * 32);
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
      const float4 w4 = wr[k4];
      v = fmaf(w4.x, h0[4 * k4], v); v = fmaf(w4.y, h0[4 * k4 + 1], v);
      v = fmaf(w4.z, h0[4 * k4 + 2], v); v = fmaf(w4.w, h0[4 * k4 + 3], v);
    }
    h1col[o * ENC_T] = fmaxf(v, 0.f);
  }
#pragma unroll
  for (int l = 0; l < 32; ++l) out[l] = W[EB_B3 + l];
  for (int j0 = 0; j0 < 256; j0 += 8) {
    float a[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) a[jj] = W[EB_B2 + j0 + jj];
    const float* w2 = W + EB_W2 + j0 * 64;          // [64][8] block of this group of 8 hidden units
#pragma unroll 8
    for (int k = 0; k < 64; ++k) {
      const float h = h1col[k * ENC_T];
      const float4 wa = *reinterpret_cast<const float4*>(w2 + k * 8);
      const float4 wb = *reinterpret_cast<const float4*>(w2 + k * 8 + 4);
      a[0] = fmaf(wa.x, h, a[0]); a[1] = fmaf(wa.y, h, a[1]); a[2] = fmaf(wa.z, h, a[2]); a[3] = fmaf(wa.w, h, a[3]);
      a[4] = fmaf(wb.x, h, a[4]); a[5] = fmaf(wb.y, h, a[5]); a[6] = fmaf(wb.z, h, a[6]); a[7] = fmaf(wb.w, h, a[7]);
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const float aj = fmaxf(a[jj], 0.f);
      const float4* w3 = reinterpret_cast<const float4*>(W + EB_W3 + (j0 + jj) * 32);
#pragma unroll
      for (int l4 = 0; l4 < 8; ++l4) {
        const float4 w4 = w3[l4];
        out[4 * l4] = fmaf(w4.x, aj, out[4 * l4]); out[4 * l4 + 1] = fmaf(w4.y, aj, out[4 * l4 + 1]);
        out[4 * l4 + 2] = fmaf(w4.z, aj, out[4 * l4 + 2]); out[4 * l4 + 3] = fmaf(w4.w, aj, out[4 * l4 + 3]);
      }
    }
  }
}

extern __shared__ __align__(16) unsigned char enc_smem_raw[];

__device__ __forceinline__ void encoder_load(EncSmem& S, const float* __restrict__ blob) {
  const float4* src = reinterpret_cast<const float4*>(blob);
  float4* dst = reinterpret_cast<float4*>(S.w);
  for (int t = threadIdx.x; t < EB_TOTAL / 4; t += ENC_T) dst[t] = __ldg(src + t);
  __syncthreads();
}

// samples -> encoder -> scatter-add into acc (map.py:446-449 + indexing.cu:59-71)
__global__ void __launch_bounds__(ENC_T, 1) encoder_scatter_kernel(const Sample* __restrict__ samples, const int* __restrict__ counters,
                                                                   const float* __restrict__ blob, long long* __restrict__ acc) {
  EncSmem& S = *reinterpret_cast<EncSmem*>(enc_smem_raw);
  encoder_load(S, blob);
  const int m = counters[0];
  for (int i = blockIdx.x * ENC_T + threadIdx.x; i < m; i += gridDim.x * ENC_T) {
    const Sample sm = samples[i];
    const float in[6] = {sm.rel[0], sm.rel[1], sm.rel[2], sm.nrm[0], sm.nrm[1], sm.nrm[2]};
    float out[32];
    encoder_one(S, in, S.h1 + threadIdx.x, out);
    long long* dst = acc + (size_t)sm.slot * DFB_LATENT_DIM;
#pragma unroll
    for (int l = 0; l < DFB_LATENT_DIM; ++l) acc_add(dst + l, out[l]);
  }
}

__global__ void __launch_bounds__(ENC_T, 1) encoder_explicit_kernel(const float* __restrict__ x, int m, const float* __restrict__ blob,
                                                                    float* __restrict__ y) {
  EncSmem& S = *reinterpret_cast<EncSmem*>(enc_smem_raw);
  encoder_load(S, blob);
  for (int i = blockIdx.x * ENC_T + threadIdx.x; i < m; i += gridDim.x * ENC_T) {
    float in[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) in[k] = x[(size_t)i * 6 + k];
    float out[32];
    encoder_one(S, in, S.h1 + threadIdx.x, out);
#pragma unroll
    for (int l = 0; l < DFB_LATENT_DIM; ++l) y[(size_t)i * DFB_LATENT_DIM + l] = out[l];
  }
}

// map.py:450-453: one warp per touched slot
__global__ void __launch_bounds__(256) ik_finalize_kernel(const int* __restrict__ counters, const int* __restrict__ touched,
                                                          long long* __restrict__ acc, int* __restrict__ acc_n, float* __restrict__ latents,
                                                          float* __restrict__ obs_count, uint8_t* __restrict__ updated, int* __restrict__ stats,
                                                          int n_new) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nt = counters[1];
  if (blockIdx.x == 0 && threadIdx.x == 0) { stats[0] = counters[0]; stats[1] = nt; stats[2] = n_new; stats[3] = 0; }
  for (int t = warp; t < nt; t += nwarps) {
    const int slot = touched[t];
    const float cnt = obs_count[slot];
    const float add = (float)acc_n[slot];                      // pcounts.float() (map.py:440)
    const float cnt_new = __fadd_rn(cnt, add);
    __syncwarp();
    if (lane < DFB_LATENT_DIM) {
      float* lp = latents + (size_t)slot * DFB_LATENT_DIM + lane;
      long long* ap = acc + (size_t)slot * DFB_LATENT_DIM + lane;
      const float sum = __fadd_rn(acc_read(*ap), __fmul_rn(*lp, cnt));   // :450
      *lp = __fdiv_rn(sum, cnt_new);                           // :452
      *ap = 0;
    }
    __syncwarp();
    if (lane == 0) { obs_count[slot] = cnt_new; acc_n[slot] = 0; updated[slot] = 1; }
  }
}

}  // namespace dfb

namespace dfb {
// tcgen05 encoder engine (encoder_tc.cu)
int tc_encoder_scatter(const void* samples, const int* m_dev, int m_max, const void* tc_blob, long long* acc, cudaStream_t s);
int tc_encoder_explicit(const float* x6, int m, const void* tc_blob, float* out, cudaStream_t s);
}  // namespace dfb

using namespace dfb;

// 1 = tcgen05 engine (default; hi+lo split FP16 operands everywhere, latents within ~2e-6 relative of the reference;
// tolerance 1e-3), 0 = FP32 CUDA cores (latents within 6e-7).
// Env DFB_ENCODER_ENGINE or dfb_set_encoder_engine().
static int g_enc_engine = -1;
static int encoder_engine() {
  if (g_enc_engine < 0) {
    const char* e = getenv("DFB_ENCODER_ENGINE");
    g_enc_engine = e ? (atoi(e) ? 1 : 0) : 1;
  }
  return g_enc_engine;
}
constexpr int ENC_TC_BLOB_BYTES = 110592 + 2048;   // encoder_tc.cu: etc::BLOB_BYTES

namespace dfb { size_t encoder_tc_blob_offset_floats() { return EB_TOTAL; } }   // used by sharded.cu

extern "C" {

size_t dfb_encoder_blob_floats(void) { return EB_TOTAL + ENC_TC_BLOB_BYTES / 4; }
int dfb_set_encoder_engine(int engine) { g_enc_engine = engine ? 1 : 0; return DFB_OK; }
int dfb_get_encoder_engine(void) { return encoder_engine(); }

size_t dfb_integrate_ws_bytes(int n, int64_t n_cells) {
  Arena a(nullptr, 0);
  IntegrateWs w;
  integrate_ws_layout(a, n, n_cells, &w);
  return a.off + 256;
}

int dfb_integrate_plan(const dfb_map_params* h_params, const float* xyz, const float* normal, int n, const int64_t* indexer,
                       int32_t* grid_count, uint32_t* grid_bits, uint8_t* unq_mask, int32_t* d_n_new, void* ws, size_t ws_bytes,
                       void* stream) {
  (void)normal;
  DFB_CHECK_ARG(h_params && n >= 0 && d_n_new, "integrate_plan");
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) { DFB_CUDA(cudaMemsetAsync(d_n_new, 0, sizeof(int32_t), s)); return DFB_OK; }
  DFB_CHECK_ARG(xyz && indexer && grid_count && grid_bits && unq_mask && ws, "integrate_plan: null pointer");
  const MapI M = to_mapi(h_params);
  const long long G = (long long)M.nx * M.ny * M.nz;
  DFB_CHECK_ARG(G > 0 && G < (1ll << 31), "integrate_plan: grid too large for 32-bit cell ids");
  Arena a(ws, ws_bytes);
  IntegrateWs w;
  integrate_ws_layout(a, n, G, &w);
  if (!a.ok()) { set_error("workspace too small: need %zu", a.off); return DFB_E_WORKSPACE; }
  const int n_words = (int)((G + 31) / 32);
  ik_ids_kernel<<<div_up(n, 256), 256, 0, s>>>(M, xyz, n, w.xn, w.gid, grid_count);
  ik_mask_mark_kernel<<<div_up(n, 256), 256, 0, s>>>(M, n, w.gid, grid_count, indexer, unq_mask, grid_bits);
  DFB_LAUNCH_CHECK();
  int rc = exclusive_scan_popc(grid_bits, w.word_rank, n_words, w.bsums, d_n_new, s);
  if (rc) return rc;
  ik_reset_counts_kernel<<<div_up(n, 256), 256, 0, s>>>(n, w.gid, grid_count);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_integrate_commit(const dfb_map_params* h_params, const float* normal, const uint8_t* unq_mask, int n, int64_t* indexer,
                         float* latent_vecs, int64_t* latent_vecs_pos, float* voxel_obs_count, uint8_t* updated_flag,
                         int64_t n_occupied, int64_t capacity, int32_t n_new, uint32_t* grid_bits, int64_t* acc, int32_t* acc_n,
                         int32_t* touched, const float* encoder_blob, int32_t* d_stats, void* ws, size_t ws_bytes, void* stream) {
  DFB_CHECK_ARG(h_params && n >= 0 && d_stats, "integrate_commit");
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) { DFB_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(int32_t) * 4, s)); return DFB_OK; }
  DFB_CHECK_ARG(normal && unq_mask && indexer && latent_vecs && latent_vecs_pos && voxel_obs_count && updated_flag && grid_bits &&
                    acc && acc_n && touched && encoder_blob && ws, "integrate_commit: null pointer");
  if (n_occupied + n_new > capacity) { set_error("integrate_commit: capacity %lld < %lld", (long long)capacity, (long long)(n_occupied + n_new)); return DFB_E_CAPACITY; }
  const MapI M = to_mapi(h_params);
  const long long G = (long long)M.nx * M.ny * M.nz;
  Arena a(ws, ws_bytes);
  IntegrateWs w;
  integrate_ws_layout(a, n, G, &w);
  if (!a.ok()) { set_error("workspace too small: need %zu", a.off); return DFB_E_WORKSPACE; }
  const int n_words = (int)((G + 31) / 32);
  ik_assign_kernel<<<div_up(n_words, 256), 256, 0, s>>>(n_words, grid_bits, w.word_rank, n_occupied, capacity, indexer,
                                                       latent_vecs_pos, latent_vecs, voxel_obs_count);
  DFB_CUDA(cudaMemsetAsync(w.counters, 0, sizeof(int) * 8, s));
  ik_samples_kernel<<<div_up(n, 256), 256, 0, s>>>(M, n, w.xn, w.gid, unq_mask, normal, indexer, voxel_obs_count, acc_n, touched,
                                                  w.counters, w.samples);
  DFB_LAUNCH_CHECK();
  if (encoder_engine() == 1) {
    int rc = tc_encoder_scatter(w.samples, w.counters, n * 8, encoder_blob + EB_TOTAL, reinterpret_cast<long long*>(acc), s);
    if (rc) return rc;
  } else {
    DFB_CUDA(cudaFuncSetAttribute(encoder_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmem)));
    // the sample count lives on the device: persistent grid sized for the worst case (8 n samples), capped at the SM count
    const int enc_grid = std::min(sm_count(), div_up((long long)n * 8, ENC_T));
    encoder_scatter_kernel<<<enc_grid, ENC_T, sizeof(EncSmem), s>>>(w.samples, w.counters, encoder_blob, reinterpret_cast<long long*>(acc));
  }
  ik_finalize_kernel<<<std::min(2 * sm_count(), div_up((long long)n * 32, 256)), 256, 0, s>>>(w.counters, touched, reinterpret_cast<long long*>(acc), acc_n, latent_vecs,
                                                                                        voxel_obs_count, updated_flag, d_stats, n_new);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_encoder_forward(const float* x, int m, const float* encoder_blob, float* out, void* stream) {
  DFB_CHECK_ARG(m >= 0, "encoder_forward");
  if (m == 0) return DFB_OK;
  DFB_CHECK_ARG(x && encoder_blob && out, "encoder_forward: null pointer");
  if (encoder_engine() == 1) return tc_encoder_explicit(x, m, encoder_blob + EB_TOTAL, out, (cudaStream_t)stream);
  DFB_CUDA(cudaFuncSetAttribute(encoder_explicit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmem)));
  encoder_explicit_kernel<<<std::min(sm_count(), div_up(m, ENC_T)), ENC_T, sizeof(EncSmem), (cudaStream_t)stream>>>(x, m, encoder_blob, out);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

}  // extern "C"
