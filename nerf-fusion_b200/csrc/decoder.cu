// Stage 4 (decoder queries) and the sampling half of stage 5 (decode_cubes), FP32 CUDA-core engine.
// Reference: system/map.py:560-580 (get_sdf), :637-688 (do_meshing sampling), system/tracker.py:179-223
// (compute_sdf_Hg), network/di_decoder.py:55-86, network/utility.py:61-126,129-149.
#include <algorithm>

#include <cstdlib>

#include "decoder_common.cuh"
#include "decoder_simt.cuh"
#include "photometric.cuh"
#include "gn_step.cuh"

namespace dfb {

__device__ __forceinline__ void load_query(DecSmem& S, const float* __restrict__ latent_row, const float rel[3], bool valid) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int k = 0; k < DFB_LATENT_DIM; ++k) S.x0[k * DEC_T + tid] = valid ? __ldg(latent_row + k) : 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) S.x0[(DFB_LATENT_DIM + k) * DEC_T + tid] = valid ? rel[k] : 0.f;
}

extern __shared__ __align__(16) unsigned char dec_smem_raw[];

// ------------------------------------------------------------------------------------------------
// explicit rows: forward_model(network_input)  (utility.py:61)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DEC_T, 1) decoder_explicit_kernel(const float* __restrict__ x, int n, const float* __restrict__ blob,
                                                                    float* __restrict__ sdf, float* __restrict__ std_) {
  DecSmem& S = *reinterpret_cast<DecSmem*>(dec_smem_raw);
  decoder_load_small(S, blob);
  const int tid = threadIdx.x;
  for (int base = blockIdx.x * DEC_T; base < n; base += gridDim.x * DEC_T) {
    const int i = base + tid;
    for (int k = 0; k < DEC_IN; ++k) S.x0[k * DEC_T + tid] = i < n ? x[(size_t)i * DEC_IN + k] : 0.f;
    float z, u;
    decoder_forward_tile(S, blob, z, u);
    if (i < n) {
      sdf[i] = tanhf(z);
      std_[i] = 0.05f + 0.5f * softplus_torch(u);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// get_sdf (+ optional vector-Jacobian product for autograd)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DEC_T, 1) get_sdf_kernel(MapDev M, const float* __restrict__ xyz, int n,
                                                           const int64_t* __restrict__ indexer, const float* __restrict__ latents,
                                                           const float* __restrict__ obs_count, const float* __restrict__ blob,
                                                           float* __restrict__ sdf, float* __restrict__ std_, uint8_t* __restrict__ valid_out,
                                                           const float* __restrict__ g_sdf, const float* __restrict__ g_std,
                                                           float* __restrict__ grad_xyz) {
  DecSmem& S = *reinterpret_cast<DecSmem*>(dec_smem_raw);
  decoder_load_small(S, blob);
  const int tid = threadIdx.x;
  for (int base = blockIdx.x * DEC_T; base < n; base += gridDim.x * DEC_T) {
    const int i = base + tid;
    bool valid = false;
    long long slot = -1;
    float rel[3] = {0.f, 0.f, 0.f};
    if (i < n) valid = map_lookup(M, xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], indexer, obs_count, slot, rel);
    load_query(S, latents + (valid ? slot : 0) * DFB_LATENT_DIM, rel, valid);
    float z, u;
    decoder_forward_tile(S, blob, z, u);
    const float s = tanhf(z);
    const float sd = 0.05f + 0.5f * softplus_torch(u);
    if (i < n) {
      valid_out[i] = valid ? 1 : 0;
      if (sdf) sdf[i] = valid ? s : 0.f;
      if (std_) std_[i] = valid ? sd : 0.f;
    }
    if (grad_xyz) {   // uniform across the grid
      float gs = 0.f, gu = 0.f;
      if (valid) {
        gs = g_sdf ? g_sdf[i] * (1.0f - s * s) : 0.f;
        // d softplus(u)/du = sigmoid(u) (1 past the threshold)
        const float dsp = softplus_grad(u);
        gu = g_std ? g_std[i] * 0.5f * dsp : 0.f;
      }
      float g[3];
      decoder_backward_tile(S, blob, gs, gu, g);
      if (i < n) {
#pragma unroll
        for (int a = 0; a < 3; ++a) grad_xyz[3 * (size_t)i + a] = valid ? div_vs(g[a], M.vs, M.inv_vs, M.div_mode) : 0.f;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// compute_sdf_Hg fused
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DEC_T, 1) sdf_hg_kernel(MapDev M, PoseDev P, const float* __restrict__ obs, int n,
                                                          const int64_t* __restrict__ indexer, const float* __restrict__ latents,
                                                          const float* __restrict__ obs_count, const float* __restrict__ blob,
                                                          int robust, float robust_k, int with_J, double* __restrict__ packed,
                                                          const GnShared* __restrict__ gs, int gi, const int* __restrict__ n_dev) {
  if (gs) {
    if (gs->done[gi]) return;
    P = *reinterpret_cast<const PoseDev*>(gs->pose_sdf);
  }
  if (n_dev) n = min(n, max(*n_dev, 0));
  DecSmem& S = *reinterpret_cast<DecSmem*>(dec_smem_raw);
  decoder_load_small(S, blob);
  const int tid = threadIdx.x;
  float acc[29];
#pragma unroll
  for (int k = 0; k < 29; ++k) acc[k] = 0.f;
  for (int base = blockIdx.x * DEC_T; base < n; base += gridDim.x * DEC_T) {
    const int i = base + tid;
    bool valid = false;
    long long slot = -1;
    float rel[3] = {0.f, 0.f, 0.f}, pc[3] = {0.f, 0.f, 0.f};
    if (i < n) {
      pc[0] = obs[3 * (size_t)i]; pc[1] = obs[3 * (size_t)i + 1]; pc[2] = obs[3 * (size_t)i + 2];
      float pw[3];
      xform(P.Rt, P.tt, pc[0], pc[1], pc[2], pw);
      valid = map_lookup(M, pw[0], pw[1], pw[2], indexer, obs_count, slot, rel);
    }
    load_query(S, latents + (valid ? slot : 0) * DFB_LATENT_DIM, rel, valid);
    float z, u;
    decoder_forward_tile(S, blob, z, u);
    const float s = tanhf(z);
    const float sd = 0.05f + 0.5f * softplus_torch(u);
    const float r = s / sd;                                   // tracker.py:191
    float J[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (with_J) {
      float g[3];
      decoder_backward_tile(S, blob, valid ? (1.0f - s * s) / sd : 0.f, 0.f, g);
      if (valid) {
        float gw[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) gw[a] = div_vs(g[a], M.vs, M.inv_vs, M.div_mode);
        sdf_jacobian(P, gw, pc, J);
      }
    }
    if (valid) hg_accumulate(acc, J, r, robust_w(r, robust, robust_k), with_J != 0);
  }
  block_reduce_atomic<29, DEC_T>(acc, packed);
}


// ------------------------------------------------------------------------------------------------
// latent optimiser (map.py:29-113 OptimizeProcess.do_optimize): one Adam iteration = latent_grad_kernel + latent_adam_kernel
// ------------------------------------------------------------------------------------------------
// Sample i: latent row inv[i] of `latents`, relative position xyz[i], target gt[i].  Loss (map.py:88-103): negative Gaussian
// log-likelihood of clamp(gt, +-0.2) under N(clamp(sdf, +-0.2), std), summed and divided by the sample count; its gradient
// w.r.t. the 29 latent inputs is scatter-added into grad[inv[i]].
__global__ void __launch_bounds__(DEC_T, 1) latent_grad_kernel(const float* __restrict__ latents, const int64_t* __restrict__ inv,
                                                               const float* __restrict__ xyz, const float* __restrict__ gt, int n,
                                                               const float* __restrict__ blob, float inv_n, float* __restrict__ grad) {
  DecSmem& S = *reinterpret_cast<DecSmem*>(dec_smem_raw);
  decoder_load_small(S, blob);
  const int tid = threadIdx.x;
  for (int base = blockIdx.x * DEC_T; base < n; base += gridDim.x * DEC_T) {
    const int i = base + tid;
    const bool valid = i < n;
    const long long row = valid ? inv[i] : 0;
    float rel[3] = {0.f, 0.f, 0.f};
    if (valid) { rel[0] = xyz[3 * (size_t)i]; rel[1] = xyz[3 * (size_t)i + 1]; rel[2] = xyz[3 * (size_t)i + 2]; }
    load_query(S, latents + row * DFB_LATENT_DIM, rel, valid);
    float z, u;
    decoder_forward_tile(S, blob, z, u);
    const float s = tanhf(z), sd = 0.05f + 0.5f * softplus_torch(u);
    float seed_z = 0.f, seed_u = 0.f;
    if (valid) {
      const float g = fminf(fmaxf(gt[i], -0.2f), 0.2f), pd = fminf(fmaxf(s, -0.2f), 0.2f);
      const float d = g - pd;
      // -log N(g; pd, sd) = d^2 / (2 sd^2) + log sd + const
      const float dl_dpd = (s >= -0.2f && s <= 0.2f) ? -d / (sd * sd) * inv_n : 0.f;        // clamp passes the gradient inside [min, max]
      const float dl_dsd = (-d * d / (sd * sd * sd) + 1.0f / sd) * inv_n;
      seed_z = dl_dpd * (1.0f - s * s);
      seed_u = dl_dsd * 0.5f * softplus_grad(u);
    }
    float g_in[DEC_IN];
    decoder_backward_inputs_tile(S, blob, seed_z, seed_u, g_in);
    if (valid) {
#pragma unroll
      for (int k = 0; k < DFB_LATENT_DIM; ++k) atomicAdd(grad + row * DFB_LATENT_DIM + k, g_in[k]);
    }
  }
}

// torch.optim.Adam (single tensor, no weight decay, amsgrad off) on (n_rows, 29) latents, one warp per row; adds the code
// regulariser's gradient reg_scale * x / ||x|| (map.py:98-101) and clears grad for the next iteration.
__global__ void __launch_bounds__(256) latent_adam_kernel(float* __restrict__ latents, float* __restrict__ grad, float* __restrict__ m1,
                                                          float* __restrict__ m2, int n_rows, float reg_scale, float beta1, float beta2,
                                                          float step_size, float bc2_sqrt, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const size_t i = (size_t)row * DFB_LATENT_DIM + lane;
  const bool on = lane < DFB_LATENT_DIM;
  const float x = on ? latents[i] : 0.f;
  float g = on ? grad[i] : 0.f;
  if (reg_scale != 0.f) {
    const float nrm = sqrtf(warp_sum(x * x));
    if (nrm > 0.f) g += reg_scale * x / nrm;
  }
  if (on) {
    const float m = m1[i] + (1.0f - beta1) * (g - m1[i]);                 // exp_avg.lerp_(grad, 1 - beta1)
    const float v = m2[i] * beta2 + (1.0f - beta2) * g * g;              // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    m1[i] = m; m2[i] = v;
    latents[i] = x - step_size * (m / (sqrtf(v) / bc2_sqrt + eps));      // param.addcdiv_(exp_avg, denom, value = -step_size)
    grad[i] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// decode_cubes
// ------------------------------------------------------------------------------------------------
// low-resolution pass: query id = b * r^3 + cell
__global__ void __launch_bounds__(DEC_T, 1) cube_low_kernel(const float* __restrict__ latents, const int64_t* __restrict__ occ,
                                                            int B, int r, float vsize, float a, const float* __restrict__ blob,
                                                            float* __restrict__ low_sdf, float* __restrict__ low_std) {
  DecSmem& S = *reinterpret_cast<DecSmem*>(dec_smem_raw);
  decoder_load_small(S, blob);
  const int tid = threadIdx.x;
  const long long r3 = (long long)r * r * r;
  const long long n = (long long)B * r3;
  for (long long base = (long long)blockIdx.x * DEC_T; base < n; base += (long long)gridDim.x * DEC_T) {
    const long long i = base + tid;
    const bool valid = i < n;
    float rel[3] = {0.f, 0.f, 0.f};
    long long slot = 0;
    if (valid) {
      const int b = (int)(i / r3);
      const int c = (int)(i - (long long)b * r3);
      rel[0] = lattice(c / (r * r), vsize, a);
      rel[1] = lattice((c / r) % r, vsize, a);
      rel[2] = lattice(c % r, vsize, a);
      slot = occ[b];
    }
    load_query(S, latents + slot * DFB_LATENT_DIM, rel, valid);
    float z, u;
    decoder_forward_tile(S, blob, z, u);
    if (valid) {
      low_sdf[i] = tanhf(z);
      low_std[i] = 0.05f + 0.5f * softplus_torch(u);
    }
  }
}

// trilinear x2 with align_corners=True (F.interpolate, map.py:659-664), negate (map.py:688), and list the
// |sdf| < band samples for exact re-decoding (map.py:668).
constexpr int UPS_T = 1024;
__global__ void __launch_bounds__(UPS_T) cube_upsample_kernel(const float* __restrict__ low_sdf, const float* __restrict__ low_std,
                                                            int B, int r, float band, float* __restrict__ cube_sdf,
                                                            float* __restrict__ cube_std, int* __restrict__ refine_count,
                                                            long long* __restrict__ refine_list) {
  const int R = 2 * r;
  const long long R3 = (long long)R * R * R;
  const long long n = (long long)B * R3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool hit = false;
  if (i < n) {
    const int b = (int)(i / R3);
    const int c = (int)(i - (long long)b * R3);
    const int ix = c / (R * R), iy = (c / R) % R, iz = c % R;
    const float scale = (float)(r - 1) / (float)(R - 1);        // area_pixel_compute_scale, align_corners
    const float fx = scale * ix, fy = scale * iy, fz = scale * iz;
    const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
    const int x1 = x0 + (x0 < r - 1 ? 1 : 0), y1 = y0 + (y0 < r - 1 ? 1 : 0), z1 = z0 + (z0 < r - 1 ? 1 : 0);
    const float lx1 = fx - x0, ly1 = fy - y0, lz1 = fz - z0;
    const float lx0 = 1.f - lx1, ly0 = 1.f - ly1, lz0 = 1.f - lz1;
    const float* ls = low_sdf + (long long)b * r * r * r;
    const float* ld = low_std + (long long)b * r * r * r;
#define AT(p, X, Y, Z) p[((X) * r + (Y)) * r + (Z)]
    const float s = lx0 * (ly0 * (lz0 * AT(ls, x0, y0, z0) + lz1 * AT(ls, x0, y0, z1)) + ly1 * (lz0 * AT(ls, x0, y1, z0) + lz1 * AT(ls, x0, y1, z1))) +
                    lx1 * (ly0 * (lz0 * AT(ls, x1, y0, z0) + lz1 * AT(ls, x1, y0, z1)) + ly1 * (lz0 * AT(ls, x1, y1, z0) + lz1 * AT(ls, x1, y1, z1)));
    const float d = lx0 * (ly0 * (lz0 * AT(ld, x0, y0, z0) + lz1 * AT(ld, x0, y0, z1)) + ly1 * (lz0 * AT(ld, x0, y1, z0) + lz1 * AT(ld, x0, y1, z1))) +
                    lx1 * (ly0 * (lz0 * AT(ld, x1, y0, z0) + lz1 * AT(ld, x1, y0, z1)) + ly1 * (lz0 * AT(ld, x1, y1, z0) + lz1 * AT(ld, x1, y1, z1)));
#undef AT
    cube_sdf[i] = -s;
    cube_std[i] = d;
    hit = fabsf(s) < band;
  }
  // One atomic per BLOCK: with ~45 % of the samples inside the band a warp-level append is one atomic per warp on a single
  // address -- 25 M of them for 200 k voxels at r = 8, i.e. the whole 12 ms the kernel took (its traffic is worth ~1.5 ms).
  __shared__ int s_cnt[UPS_T / 32];
  __shared__ int s_base;
  const unsigned bal = __ballot_sync(0xffffffffu, hit);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s_cnt[w] = __popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int k = 0; k < UPS_T / 32; ++k) { const int t = s_cnt[k]; s_cnt[k] = run; run += t; }
    s_base = run > 0 ? atomicAdd(refine_count, run) : 0;
  }
  __syncthreads();
  if (hit) refine_list[s_base + s_cnt[w] + __popc(bal & ((1u << lane) - 1u))] = i;
}

__global__ void __launch_bounds__(DEC_T, 1) cube_refine_kernel(const float* __restrict__ latents, const int64_t* __restrict__ occ,
                                                               int r, float vsize, float a, const float* __restrict__ blob,
                                                               const int* __restrict__ refine_count, const long long* __restrict__ refine_list,
                                                               float* __restrict__ cube_sdf, float* __restrict__ cube_std) {
  DecSmem& S = *reinterpret_cast<DecSmem*>(dec_smem_raw);
  decoder_load_small(S, blob);
  const int tid = threadIdx.x;
  const int R = 2 * r;
  const long long R3 = (long long)R * R * R;
  const int n = *refine_count;
  for (int base = blockIdx.x * DEC_T; base < n; base += gridDim.x * DEC_T) {
    const int t = base + tid;
    const bool valid = t < n;
    float rel[3] = {0.f, 0.f, 0.f};
    long long slot = 0, i = 0;
    if (valid) {
      i = refine_list[t];
      const int b = (int)(i / R3);
      const int c = (int)(i - (long long)b * R3);
      rel[0] = lattice(c / (R * R), vsize, a);
      rel[1] = lattice((c / R) % R, vsize, a);
      rel[2] = lattice(c % R, vsize, a);
      slot = occ[b];
    }
    load_query(S, latents + slot * DFB_LATENT_DIM, rel, valid);
    float z, u;
    decoder_forward_tile(S, blob, z, u);
    if (valid) {
      cube_sdf[i] = -tanhf(z);
      cube_std[i] = 0.05f + 0.5f * softplus_torch(u);
    }
  }
}

template <typename K>
static int set_dec_smem(K kernel) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecSmem));
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DFB_E_CUDA; }
  return DFB_OK;
}

static int dec_grid(long long n) { return (int)std::min<long long>(div_up(n, DEC_T), (long long)sm_count()); }

}  // namespace dfb

using namespace dfb;

// engine selection: 1 = tcgen05 FP16 (default), 0 = FP32 CUDA cores.  Env DFB_DECODER_ENGINE or dfb_set_decoder_engine().
static int g_engine = -1;
int dfb::decoder_engine() {
  if (g_engine < 0) {
    const char* e = getenv("DFB_DECODER_ENGINE");
    g_engine = e ? atoi(e) : 1;
  }
  return g_engine;
}
constexpr int TC_BLOB_BYTES = 196608 + 6144;   // decoder_tc.cu: tc::BLOB_BYTES
static inline const void* tc_part(const float* blob) { return blob + DB_TOTAL; }

extern "C" {

size_t dfb_decoder_blob_floats(void) { return DB_TOTAL + TC_BLOB_BYTES / 4; }
int dfb_set_decoder_engine(int engine) { g_engine = engine ? 1 : 0; return DFB_OK; }
int dfb_get_decoder_engine(void) { return decoder_engine(); }

int dfb_decoder_forward(const float* x, int n, const float* decoder_blob, float* sdf, float* std_, void* stream) {
  DFB_CHECK_ARG(n >= 0, "decoder_forward");
  if (n == 0) return DFB_OK;
  DFB_CHECK_ARG(x && decoder_blob && sdf && std_, "decoder_forward: null pointer");
  if (decoder_engine() == 1) return tc_decoder_explicit(x, n, tc_part(decoder_blob), sdf, std_, (cudaStream_t)stream);
  int rc = set_dec_smem(decoder_explicit_kernel);
  if (rc) return rc;
  decoder_explicit_kernel<<<dec_grid(n), DEC_T, sizeof(DecSmem), (cudaStream_t)stream>>>(x, n, decoder_blob, sdf, std_);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_get_sdf(const dfb_map_params* h_params, const float* xyz, int n, const int64_t* indexer, const float* latent_vecs,
                const float* voxel_obs_count, const float* decoder_blob, float* sdf, float* std_, uint8_t* valid,
                const float* g_sdf, const float* g_std, float* grad_xyz, void* stream) {
  DFB_CHECK_ARG(h_params && n >= 0, "get_sdf");
  if (n == 0) return DFB_OK;
  DFB_CHECK_ARG(xyz && indexer && latent_vecs && voxel_obs_count && decoder_blob && valid, "get_sdf: null pointer");
  if (decoder_engine() == 1)
    return tc_get_sdf(to_dev(h_params), xyz, n, indexer, latent_vecs, voxel_obs_count, tc_part(decoder_blob), sdf, std_, valid, g_sdf,
                      g_std, grad_xyz, (cudaStream_t)stream);
  int rc = set_dec_smem(get_sdf_kernel);
  if (rc) return rc;
  get_sdf_kernel<<<dec_grid(n), DEC_T, sizeof(DecSmem), (cudaStream_t)stream>>>(to_dev(h_params), xyz, n, indexer, latent_vecs,
                                                                                voxel_obs_count, decoder_blob, sdf, std_, valid,
                                                                                g_sdf, g_std, grad_xyz);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}


int dfb_latent_adam_step(float* latents, int n_rows, const int64_t* inv, const float* rel_xyz, const float* gt_sdf, int n_samples,
                         const float* decoder_blob, float* grad, float* adam_m, float* adam_v, int step, float lr, float reg_scale,
                         void* stream) {
  DFB_CHECK_ARG(n_rows >= 0 && n_samples >= 0 && step >= 1, "latent_adam_step");
  if (n_rows == 0 || n_samples == 0) return DFB_OK;
  DFB_CHECK_ARG(latents && inv && rel_xyz && gt_sdf && decoder_blob && grad && adam_m && adam_v, "latent_adam_step: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = set_dec_smem(latent_grad_kernel);
  if (rc) return rc;
  latent_grad_kernel<<<dec_grid(n_samples), DEC_T, sizeof(DecSmem), s>>>(latents, inv, rel_xyz, gt_sdf, n_samples, decoder_blob,
                                                                        1.0f / (float)n_samples, grad);
  const double b1 = 0.9, b2 = 0.999;                       // torch.optim.Adam defaults (map.py:84)
  const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
  latent_adam_kernel<<<div_up((long long)n_rows * 32, 256), 256, 0, s>>>(latents, grad, adam_m, adam_v, n_rows, reg_scale, (float)b1, (float)b2,
                                                                        (float)(lr / bc1), (float)sqrt(bc2), 1e-8f);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_sdf_hg(const dfb_map_params* h_params, const float* obs_xyz, int n, const float* h_pose, const int64_t* indexer,
               const float* latent_vecs, const float* voxel_obs_count, const float* decoder_blob, int robust, float robust_k,
               int compute_J, double* out44, void* stream) {
  DFB_CHECK_ARG(h_params && h_pose && out44 && n >= 0, "sdf_hg");
  cudaStream_t s = (cudaStream_t)stream;
  double* packed = out44 + 44;                       // caller provides 80 doubles
  DFB_CUDA(cudaMemsetAsync(out44, 0, sizeof(double) * 80, s));
  if (n == 0) return DFB_OK;
  DFB_CHECK_ARG(obs_xyz && indexer && latent_vecs && voxel_obs_count && decoder_blob, "sdf_hg: null pointer");
  const PoseDev P = to_pose(h_pose);
  if (decoder_engine() == 1) {
    int rc = tc_sdf_hg(to_dev(h_params), P, obs_xyz, n, indexer, latent_vecs, voxel_obs_count, tc_part(decoder_blob), robust, robust_k,
                       compute_J, packed, s);
    if (rc) return rc;
    launch_hg_expand(packed, out44, s);
    DFB_LAUNCH_CHECK();
    return DFB_OK;
  }
  int rc = set_dec_smem(sdf_hg_kernel);
  if (rc) return rc;
  sdf_hg_kernel<<<dec_grid(n), DEC_T, sizeof(DecSmem), s>>>(to_dev(h_params), P, obs_xyz, n, indexer, latent_vecs, voxel_obs_count,
                                                            decoder_blob, robust, robust_k, compute_J, packed, nullptr, 0, nullptr);
  launch_hg_expand(packed, out44, s);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

size_t dfb_decode_cubes_ws_bytes(int B, int r) {
  Arena a(nullptr, 0);
  size_t r3 = (size_t)r * r * r;
  a.take<float>((size_t)B * r3 + 1); a.take<float>((size_t)B * r3 + 1); a.take<int>(4);
  a.take<long long>((size_t)B * r3 * 8 + 1);
  return a.off + 256;
}

int dfb_decode_cubes(const float* latent_vecs, const int64_t* occ, int B, int r, float refine_band, const float* decoder_blob,
                     float* cube_sdf, float* cube_std, void* ws, size_t ws_bytes, void* stream) {
  DFB_CHECK_ARG(B >= 0 && r >= 2 && r <= 64, "decode_cubes");
  if (B == 0) return DFB_OK;
  DFB_CHECK_ARG(latent_vecs && occ && decoder_blob && cube_sdf && cube_std && ws, "decode_cubes: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  Arena a(ws, ws_bytes);
  const size_t r3 = (size_t)r * r * r;
  float* low_sdf = a.take<float>((size_t)B * r3 + 1);
  float* low_std = a.take<float>((size_t)B * r3 + 1);
  int* refine_count = a.take<int>(4);
  long long* refine_list = a.take<long long>((size_t)B * r3 * 8 + 1);
  if (!a.ok()) { set_error("workspace too small: need %zu", a.off); return DFB_E_WORKSPACE; }
  // map.py:641-647: a = -(r//2)/r, b = 1 + ((r-1)//2)/r in Python doubles; vsize = (b-a)/(res-1) cast to fp32
  const double sa = -(double)(r / 2) * (1.0 / r);
  const double sb = 1.0 + (double)((r - 1) / 2) * (1.0 / r);
  const float a32 = (float)sa;
  const float v_low = (float)((sb - sa) / (r - 1));
  const float v_high = (float)((sb - sa) / (2 * r - 1));
  DFB_CUDA(cudaMemsetAsync(refine_count, 0, sizeof(int) * 4, s));
  const long long nh_tc = (long long)B * r3 * 8;
  if (decoder_engine() == 1) {
    int rc = tc_cube_low(latent_vecs, occ, B, r, v_low, a32, tc_part(decoder_blob), low_sdf, low_std, s);
    if (rc) return rc;
    cube_upsample_kernel<<<div_up(nh_tc, UPS_T), UPS_T, 0, s>>>(low_sdf, low_std, B, r, refine_band, cube_sdf, cube_std, refine_count, refine_list);
    DFB_LAUNCH_CHECK();
    return tc_cube_refine(latent_vecs, occ, r, v_high, a32, tc_part(decoder_blob), refine_count, refine_list, cube_sdf, cube_std, s);
  }
  int rc = set_dec_smem(cube_low_kernel);
  if (rc) return rc;
  rc = set_dec_smem(cube_refine_kernel);
  if (rc) return rc;
  cube_low_kernel<<<dec_grid((long long)B * r3), DEC_T, sizeof(DecSmem), s>>>(latent_vecs, occ, B, r, v_low, a32, decoder_blob, low_sdf, low_std);
  const long long nh = (long long)B * r3 * 8;
  cube_upsample_kernel<<<div_up(nh, UPS_T), UPS_T, 0, s>>>(low_sdf, low_std, B, r, refine_band, cube_sdf, cube_std, refine_count, refine_list);
  // the refine count lives on the device: launch a persistent grid that reads it
  cube_refine_kernel<<<sm_count(), DEC_T, sizeof(DecSmem), s>>>(latent_vecs, occ, r, v_high, a32, decoder_blob, refine_count, refine_list,
                                                                cube_sdf, cube_std);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

}  // extern "C"

// Gauss-Newton evaluations launched by gauss_newton.cu.
namespace dfb {
// tcgen05 engine: SDF term + optional photometric term + step in one launch (decoder_tc.cu gn_eval_kernel)
int launch_sdf_rgb_gn(const dfb_map_params* h_params, const float* obs_xyz, int n, const int32_t* n_dev, const int64_t* indexer, const float* latent_vecs,
                      const float* voxel_obs_count, const float* decoder_blob, int sdf_robust, float sdf_robust_k, int compute_J,
                      const dfb_rgb_level* L, const float* intr4, float min_grad_scale, float max_depth_delta, int rgb_robust,
                      float rgb_robust_k, GnShared* gs, int gi, const gn::StepArgs* sa, cudaStream_t s) {
  RgbDev R = {};
  if (L) {
    R.prev_I = L->prev_I; R.prev_D = L->prev_D; R.cur_I = L->cur_I; R.cur_D = L->cur_D; R.cur_G = L->cur_G; R.H = L->H; R.W = L->W;
    R.P.fx = intr4[0]; R.P.fy = intr4[1]; R.P.cx = intr4[2]; R.P.cy = intr4[3];
    R.P.min_grad_scale = min_grad_scale; R.P.max_depth_delta = max_depth_delta;
    R.robust = rgb_robust; R.robust_k = rgb_robust_k; R.on = 1;
  }
  return tc_gn_eval(to_dev(h_params), obs_xyz, n, n_dev, indexer, latent_vecs, voxel_obs_count, tc_part(decoder_blob), sdf_robust, sdf_robust_k,
                    compute_J, R, gs, gi, *sa, s);
}

// FP32 CUDA-core engine: the SDF term alone (pose and "group finished" flag from `gs`, sums into gs->sums[0])
int launch_sdf_hg_gn(const dfb_map_params* h_params, const float* obs_xyz, int n, const int32_t* n_dev, const int64_t* indexer, const float* latent_vecs,
                     const float* voxel_obs_count, const float* decoder_blob, int robust, float robust_k, int compute_J, GnShared* gs, int gi,
                     cudaStream_t s) {
  if (n == 0) return DFB_OK;
  const PoseDev P = {};
  int rc = set_dec_smem(sdf_hg_kernel);
  if (rc) return rc;
  sdf_hg_kernel<<<dec_grid(n), DEC_T, sizeof(DecSmem), s>>>(to_dev(h_params), P, obs_xyz, n, indexer, latent_vecs, voxel_obs_count, decoder_blob,
                                                            robust, robust_k, compute_J, gs->sums[0], gs, gi, n_dev);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}
}  // namespace dfb
