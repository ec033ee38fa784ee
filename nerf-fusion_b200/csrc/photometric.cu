// Photometric term: Sobel gradients, per-pixel residual/Jacobian (drop-in ops) and the fused H/g reduction.
// Reference: system/ext/imgproc/photometric.cu:3-138, system/tracker.py:136-177.
#include <cstdlib>
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

// ------------------------------------------------------------------------------------------------
// Image half of the tracker front end in two launches (tracker.py:42-57, 84; main.py:56-57): depth clipping, intensity,
// the 3-level pyramid (bilinear align_corners=True for intensity, nearest for depth) and the Sobel gradients of all levels.
// The reference runs these as ~13 torch kernels; the arithmetic below reproduces torch's CUDA kernels bit for bit (rounding
// sequences established on the GPU by tools/aten_formula_probe.py and pinned by test_frame_images_equals_torch):
//   mean over 3 channels  = ((r + b) + g) * (1/3)                      (4-way unrolled reduction, then the mean factor)
//   bilinear              = fma(h0, fma(w0, a, w1 * b), h1 * fma(w0, c, w1 * d)),  source index = dst * (in - 1) / (out - 1)
//   nearest               = src[min(floor(dst * in / out), in - 1)]
struct PyrDims { int H[3], W[3]; float sh[2], sw[2], nh[2], nw[2]; };   // bilinear / nearest scales of level l -> l + 1

__device__ __forceinline__ float px_intensity(const float* __restrict__ rgb, int W, int v, int u) {
  const float* p = rgb + 3 * ((size_t)v * W + u);
  return __fmul_rn(__fadd_rn(__fadd_rn(p[0], p[2]), p[1]), 1.0f / 3.0f);
}
__device__ __forceinline__ float px_cut(const float* __restrict__ depth, int W, int v, int u, float cut_near, float cut_far, bool cut) {
  const float d = depth[(size_t)v * W + u];
  return (cut && (d < cut_near || d > cut_far)) ? CUDART_NAN_F : d;
}
// bilinear sample of a (Hin x Win) image given by the functor `at(v, u)` at output pixel (v2, u2)
template <typename F>
__device__ __forceinline__ float bilinear_ac(F at, int Hin, int Win, float sh, float sw, int v2, int u2) {
  const float h1r = __fmul_rn(sh, (float)v2), w1r = __fmul_rn(sw, (float)u2);
  const int h1 = (int)h1r, w1 = (int)w1r;
  const int h1p = h1 < Hin - 1 ? 1 : 0, w1p = w1 < Win - 1 ? 1 : 0;
  const float h1l = __fsub_rn(h1r, (float)h1), w1l = __fsub_rn(w1r, (float)w1);
  const float h0l = __fsub_rn(1.0f, h1l), w0l = __fsub_rn(1.0f, w1l);
  const float top = __fmaf_rn(w0l, at(h1, w1), __fmul_rn(w1l, at(h1, w1 + w1p)));
  const float bot = __fmaf_rn(w0l, at(h1 + h1p, w1), __fmul_rn(w1l, at(h1 + h1p, w1 + w1p)));
  return __fmaf_rn(h0l, top, __fmul_rn(h1l, bot));
}
__device__ __forceinline__ int nearest_src(float scale, int dst, int n_in) { return min((int)floorf(__fmul_rn((float)dst, scale)), n_in - 1); }

__global__ void __launch_bounds__(256) frame_images_kernel(const float* __restrict__ rgb, const float* __restrict__ depth, PyrDims P, float cut_near,
                                                           float cut_far, int cut, float* __restrict__ I0, float* __restrict__ D0,
                                                           float* __restrict__ I1, float* __restrict__ D1, float* __restrict__ I2,
                                                           float* __restrict__ D2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int H0 = P.H[0], W0 = P.W[0], H1 = P.H[1], W1 = P.W[1], H2 = P.H[2], W2 = P.W[2];
  auto i0 = [&](int v, int u) { return px_intensity(rgb, W0, v, u); };
  auto i1 = [&](int v, int u) { return bilinear_ac(i0, H0, W0, P.sh[0], P.sw[0], v, u); };
  if (i < H0 * W0) {
    const int v = i / W0, u = i - v * W0;
    I0[i] = i0(v, u);
    D0[i] = px_cut(depth, W0, v, u, cut_near, cut_far, cut != 0);
  }
  if (i < H1 * W1) {
    const int v = i / W1, u = i - v * W1;
    I1[i] = i1(v, u);
    D1[i] = px_cut(depth, W0, nearest_src(P.nh[0], v, H0), nearest_src(P.nw[0], u, W0), cut_near, cut_far, cut != 0);
  }
  if (i < H2 * W2) {
    const int v = i / W2, u = i - v * W2;
    I2[i] = bilinear_ac(i1, H1, W1, P.sh[1], P.sw[1], v, u);
    const int v1 = nearest_src(P.nh[1], v, H1), u1 = nearest_src(P.nw[1], u, W1);
    D2[i] = px_cut(depth, W0, nearest_src(P.nh[0], v1, H0), nearest_src(P.nw[0], u1, W0), cut_near, cut_far, cut != 0);
  }
}

// gradient_xy of the three levels in one launch (same arithmetic as gradient_xy_kernel)
__global__ void __launch_bounds__(256) gradient3_kernel(const float* __restrict__ I0, const float* __restrict__ I1, const float* __restrict__ I2,
                                                        PyrDims P, float* __restrict__ G0, float* __restrict__ G1, float* __restrict__ G2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n0 = P.H[0] * P.W[0], n1 = P.H[1] * P.W[1], n2 = P.H[2] * P.W[2];
  const float* I; float* g; int H, W;
  if (i < n0) { I = I0; g = G0; H = P.H[0]; W = P.W[0]; }
  else if (i < n0 + n1) { i -= n0; I = I1; g = G1; H = P.H[1]; W = P.W[1]; }
  else if (i < n0 + n1 + n2) { i -= n0 + n1; I = I2; g = G2; H = P.H[2]; W = P.W[2]; }
  else return;
  const int v = i / W, u = i - v * W;
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

int dfb_frame_images(const float* rgb, const float* depth, int H, int W, float cut_near, float cut_far, int cut, float* I0, float* D0,
                     float* I1, float* D1, float* I2, float* D2, float* G0, float* G1, float* G2, void* stream) {
  DFB_CHECK_ARG(H >= 8 && W >= 8 && rgb && depth && I0 && D0 && I1 && D1 && I2 && D2 && G0 && G1 && G2, "frame_images");
  cudaStream_t s = (cudaStream_t)stream;
  PyrDims P;
  P.H[0] = H; P.W[0] = W; P.H[1] = H / 2; P.W[1] = W / 2; P.H[2] = P.H[1] / 2; P.W[2] = P.W[1] / 2;
  for (int l = 0; l < 2; ++l) {
    // torch: area_pixel_compute_scale(align_corners=True) = float(in - 1) / (out - 1); nearest: float(in) / out
    P.sh[l] = P.H[l + 1] > 1 ? (float)(P.H[l] - 1) / (float)(P.H[l + 1] - 1) : 0.f;
    P.sw[l] = P.W[l + 1] > 1 ? (float)(P.W[l] - 1) / (float)(P.W[l + 1] - 1) : 0.f;
    P.nh[l] = (float)P.H[l] / (float)P.H[l + 1];
    P.nw[l] = (float)P.W[l] / (float)P.W[l + 1];
  }
  frame_images_kernel<<<div_up((long long)H * W, 256), 256, 0, s>>>(rgb, depth, P, cut_near, cut_far, cut, I0, D0, I1, D1, I2, D2);
  const long long ng = (long long)H * W + (long long)P.H[1] * P.W[1] + (long long)P.H[2] * P.W[2];
  gradient3_kernel<<<div_up(ng, 256), 256, 0, s>>>(I0, I1, I2, P, G0, G1, G2);
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
  // programmatic dependent launch (like gn_eval_kernel): the blocks are placed while the previous evaluation's last block still
  // runs its step; everything this kernel reads from `gs` comes after the dependency wait
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
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
  static const bool no_pdl = getenv("DFB_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(HG_T); cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = no_pdl ? 0 : 1;
  cudaError_t e;
  if (npx >= 4LL * HG_T * 148) {
    cfg.gridDim = dim3(div_up(npx, HG_T * 4));
    e = cudaLaunchKernelEx(&cfg, rgb_step_gn_kernel<4>, L->prev_I, L->prev_D, L->cur_I, L->cur_D, L->cur_G, L->H, L->W, P, robust, robust_k,
                           compute_J, gs, gi, *sa);
  } else {
    cfg.gridDim = dim3(div_up(npx, HG_T));
    e = cudaLaunchKernelEx(&cfg, rgb_step_gn_kernel<1>, L->prev_I, L->prev_D, L->cur_I, L->cur_D, L->cur_G, L->H, L->W, P, robust, robust_k,
                           compute_J, gs, gi, *sa);
  }
  if (e != cudaSuccess) { set_error("rgb_step_gn launch: %s", cudaGetErrorString(e)); return DFB_E_CUDA; }
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
