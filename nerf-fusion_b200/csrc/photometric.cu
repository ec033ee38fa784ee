// Photometric term: Sobel gradients, per-pixel residual/Jacobian (drop-in ops) and the fused H/g reduction.
// Reference: system/ext/imgproc/photometric.cu:3-138, system/tracker.py:136-177.
#include "common.cuh"
#include "photometric.cuh"
#include "gn_step.cuh"

namespace dfb {

__global__ void __launch_bounds__(256) gradient_xy_kernel(const float* __restrict__ I, int H, int W, float* __restrict__ g) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  int v = i / W, u = i - v * W;
  if (v < 1 || v > H - 2 || u < 1 || u > W - 2) {
    g[2 * i] = g[2 * i + 1] = CUDART_NAN_F;
    return;
  }
  const float* r0 = I + (size_t)(v - 1) * W + u;
  const float* r1 = I + (size_t)v * W + u;
  const float* r2 = I + (size_t)(v + 1) * W + u;
  float u_d1 = r0[1] - r0[-1], u_d2 = r1[1] - r1[-1], u_d3 = r2[1] - r2[-1];
  g[2 * i] = (u_d1 + 2 * u_d2 + u_d3) / 8.0f;
  float v_d1 = r2[-1] - r0[-1], v_d2 = r2[0] - r0[0], v_d3 = r2[1] - r0[1];
  g[2 * i + 1] = (v_d1 + 2 * v_d2 + v_d3) / 8.0f;
}

__global__ void __launch_bounds__(256) rgb_odometry_kernel(const float* prev_I, const float* prev_D, const float* cur_I,
                                                           const float* cur_D, const float* dIdxy, int H, int W, RgbParams P,
                                                           float* __restrict__ f_out, float* __restrict__ J_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  int v = i / W, u = i - v * W;
  float f, J[6];
  bool ok = rgb_pixel(prev_I, prev_D, cur_I, cur_D, dIdxy, H, W, P, v, u, J_out != nullptr, f, J);
  f_out[i] = ok ? f : CUDART_NAN_F;
  if (J_out && ok) {           // the reference leaves J uninitialised where f is NaN; we do the same (never read)
#pragma unroll
    for (int a = 0; a < 6; ++a) J_out[6 * (size_t)i + a] = J[a];
  }
}

constexpr int HG_T = 256;
constexpr int HG_PIX_PER_THREAD = 4;

__global__ void __launch_bounds__(HG_T) rgb_hg_kernel(const float* prev_I, const float* prev_D, const float* cur_I,
                                                      const float* cur_D, const float* dIdxy, int H, int W, RgbParams P,
                                                      int robust, float robust_k, int with_J, double* out29,
                                                      const GnShared* __restrict__ gs, int gi) {
  if (gs) {                                      // device-resident Gauss-Newton (gauss_newton.cu)
    if (gs->done[gi]) return;
#pragma unroll
    for (int i = 0; i < 9; ++i) P.k[i] = gs->krk[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) P.kt[i] = gs->kt[i];
  }
  float acc[29];
#pragma unroll
  for (int k = 0; k < 29; ++k) acc[k] = 0.f;
  int base = blockIdx.x * (HG_T * HG_PIX_PER_THREAD) + threadIdx.x;
#pragma unroll
  for (int e = 0; e < HG_PIX_PER_THREAD; ++e) {
    int i = base + e * HG_T;
    if (i < H * W) {
      int v = i / W, u = i - v * W;
      float f, J[6];
      if (rgb_pixel(prev_I, prev_D, cur_I, cur_D, dIdxy, H, W, P, v, u, with_J != 0, f, J)) {
        if (with_J) {
#pragma unroll
          for (int a = 0; a < 6; ++a) J[a] = -J[a];   // tracker.py:162
        }
        hg_accumulate(acc, J, f, robust_w(f, robust, robust_k), with_J != 0);
      }
    }
  }
  block_reduce_atomic<29, HG_T>(acc, out29);
}

}  // namespace dfb

using namespace dfb;

static void fill_rgb_params(RgbParams& P, const float* intr, const float* k, const float* kt, float mgs, float mdd) {
  for (int i = 0; i < 9; ++i) P.k[i] = k[i];
  for (int i = 0; i < 3; ++i) P.kt[i] = kt[i];
  P.fx = intr[0]; P.fy = intr[1]; P.cx = intr[2]; P.cy = intr[3];
  P.min_grad_scale = mgs; P.max_depth_delta = mdd;
}

extern "C" {

int dfb_gradient_xy(const float* intensity, int H, int W, float* grad, void* stream) {
  DFB_CHECK_ARG(H >= 0 && W >= 0, "gradient_xy");
  if (H * W == 0) return DFB_OK;
  DFB_CHECK_ARG(intensity && grad, "gradient_xy: null pointer");
  gradient_xy_kernel<<<div_up((long long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(intensity, H, W, grad);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_rgb_odometry(const float* prev_I, const float* prev_D, const float* cur_I, const float* cur_D,
                     const float* cur_dIdxy, int H, int W, const float* h_intr, const float* h_krkinv,
                     const float* h_kt, float min_grad_scale, float max_depth_delta, float* f, float* J, void* stream) {
  DFB_CHECK_ARG(H > 0 && W > 0 && prev_I && prev_D && cur_I && cur_D && cur_dIdxy && h_intr && h_krkinv && h_kt && f,
                "rgb_odometry");
  RgbParams P;
  fill_rgb_params(P, h_intr, h_krkinv, h_kt, min_grad_scale, max_depth_delta);
  rgb_odometry_kernel<<<div_up((long long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(prev_I, prev_D, cur_I, cur_D,
                                                                                      cur_dIdxy, H, W, P, f, J);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_rgb_hg(const float* prev_I, const float* prev_D, const float* cur_I, const float* cur_D, const float* cur_dIdxy,
               int H, int W, const float* h_intr, const float* h_krkinv, const float* h_kt, float min_grad_scale,
               float max_depth_delta, int robust, float robust_k, int compute_J, double* out44, void* stream) {
  DFB_CHECK_ARG(H > 0 && W > 0 && prev_I && prev_D && cur_I && cur_D && cur_dIdxy && h_intr && h_krkinv && h_kt && out44,
                "rgb_hg");
  cudaStream_t s = (cudaStream_t)stream;
  RgbParams P;
  fill_rgb_params(P, h_intr, h_krkinv, h_kt, min_grad_scale, max_depth_delta);
  // out44 doubles as scratch: packed sums live in out44[44..72] (caller provides 80 doubles)
  double* packed = out44 + 44;
  DFB_CUDA(cudaMemsetAsync(out44, 0, sizeof(double) * 80, s));
  rgb_hg_kernel<<<div_up((long long)H * W, HG_T * HG_PIX_PER_THREAD), HG_T, 0, s>>>(
      prev_I, prev_D, cur_I, cur_D, cur_dIdxy, H, W, P, robust, robust_k, compute_J, packed, nullptr, 0);
  launch_hg_expand(packed, out44, s);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

}  // extern "C"

// Photometric-only evaluation (first group of fusion-lr-kt.yaml): pixels -> block sums -> atomics, and the last block to
// finish runs the Gauss-Newton step.  PIX pixels per thread (1 for the small pyramid levels: latency matters, not occupancy).
namespace dfb {
template <int PIX>
__global__ void __launch_bounds__(HG_T) rgb_step_gn_kernel(const float* prev_I, const float* prev_D, const float* cur_I, const float* cur_D,
                                                           const float* dIdxy, int H, int W, RgbParams P, int robust, float robust_k,
                                                           int with_J, GnShared* gs, int gi, gn::StepArgs sa) {
  __shared__ __align__(16) unsigned char scratch[gn::STEP_SCRATCH_BYTES];
  __shared__ int flag;
  if (gs->done[gi]) {
    if (blockIdx.x == 0 && threadIdx.x < 32) gn::skip_record(gs, sa);
    return;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) P.k[i] = gs->krk[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) P.kt[i] = gs->kt[i];
  float acc[29];
#pragma unroll
  for (int k = 0; k < 29; ++k) acc[k] = 0.f;
  const int base = blockIdx.x * (HG_T * PIX) + threadIdx.x;
#pragma unroll
  for (int e = 0; e < PIX; ++e) {
    const int i = base + e * HG_T;
    if (i < H * W) {
      const int v = i / W, u = i - v * W;
      float f, J[6];
      if (rgb_pixel(prev_I, prev_D, cur_I, cur_D, dIdxy, H, W, P, v, u, with_J != 0, f, J)) {
        if (with_J) {
#pragma unroll
          for (int a = 0; a < 6; ++a) J[a] = -J[a];   // tracker.py:162
        }
        hg_accumulate(acc, J, f, robust_w(f, robust, robust_k), with_J != 0);
      }
    }
  }
  block_reduce_atomic<29, HG_T>(acc, gs->sums[1]);
  __threadfence();
  __syncthreads();
  gn::tail_step(gs, sa, scratch, &flag);
}

int launch_rgb_step_gn(const dfb_rgb_level* L, const float* intr4, float min_grad_scale, float max_depth_delta, int robust, float robust_k,
                       int compute_J, GnShared* gs, int gi, const gn::StepArgs* sa, cudaStream_t s) {
  RgbParams P = {};
  P.fx = intr4[0]; P.fy = intr4[1]; P.cx = intr4[2]; P.cy = intr4[3];
  P.min_grad_scale = min_grad_scale; P.max_depth_delta = max_depth_delta;
  const long long npx = (long long)L->H * L->W;
  if (npx >= 4LL * HG_T * 148)
    rgb_step_gn_kernel<4><<<div_up(npx, HG_T * 4), HG_T, 0, s>>>(L->prev_I, L->prev_D, L->cur_I, L->cur_D, L->cur_G, L->H, L->W, P, robust,
                                                                robust_k, compute_J, gs, gi, *sa);
  else
    rgb_step_gn_kernel<1><<<div_up(npx, HG_T), HG_T, 0, s>>>(L->prev_I, L->prev_D, L->cur_I, L->cur_D, L->cur_G, L->H, L->W, P, robust,
                                                            robust_k, compute_J, gs, gi, *sa);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}
}  // namespace dfb

// Photometric term of one device-resident Gauss-Newton evaluation: K R K^-1 and K t come from `gs`, sums go to gs->sums[1].
namespace dfb {
int launch_rgb_hg_gn(const dfb_rgb_level* L, const float* intr4, float min_grad_scale, float max_depth_delta, int robust, float robust_k,
                     int compute_J, GnShared* gs, int gi, cudaStream_t s) {
  RgbParams P = {};
  P.fx = intr4[0]; P.fy = intr4[1]; P.cx = intr4[2]; P.cy = intr4[3];
  P.min_grad_scale = min_grad_scale; P.max_depth_delta = max_depth_delta;
  rgb_hg_kernel<<<div_up((long long)L->H * L->W, HG_T * HG_PIX_PER_THREAD), HG_T, 0, s>>>(
      L->prev_I, L->prev_D, L->cur_I, L->cur_D, L->cur_G, L->H, L->W, P, robust, robust_k, compute_J, gs->sums[1], gs, gi);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}
}  // namespace dfb
