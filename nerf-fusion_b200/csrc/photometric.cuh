// Per-pixel photometric residual / Jacobian shared by photometric.cu and the fused Gauss-Newton evaluation kernel.
// Reference: system/ext/imgproc/photometric.cu:24-77.
#pragma once
#include "common.cuh"

namespace dfb {

struct RgbParams {
  float k[9];
  float kt[3];
  float fx, fy, cx, cy;
  float min_grad_scale, max_depth_delta;
};

// photometric.cu:24-77 for one pixel.  Returns validity; f and J[6] filled when valid.
__device__ __forceinline__ bool rgb_pixel(const float* __restrict__ prev_I, const float* __restrict__ prev_D,
                                          const float* __restrict__ cur_I, const float* __restrict__ cur_D,
                                          const float* __restrict__ dIdxy, int H, int W, const RgbParams& P, int v, int u,
                                          bool want_J, float& f, float* J) {
  int i = v * W + u;
  float dI_dx = dIdxy[2 * i], dI_dy = dIdxy[2 * i + 1];
  float mTwo = (dI_dx * dI_dx) + (dI_dy * dI_dy);
  if (mTwo < P.min_grad_scale || isnan(mTwo)) return false;
  float d1 = cur_D[i];
  if (isnan(d1)) return false;
  float warpped_d1 = d1 * (P.k[6] * u + P.k[7] * v + P.k[8]) + P.kt[2];
  int u0 = __float2int_rn((d1 * (P.k[0] * u + P.k[1] * v + P.k[2]) + P.kt[0]) / warpped_d1);
  int v0 = __float2int_rn((d1 * (P.k[3] * u + P.k[4] * v + P.k[5]) + P.kt[1]) / warpped_d1);
  if (!(u0 >= 0 && u0 < W && v0 >= 0 && v0 < H)) return false;
  float d0 = prev_D[v0 * W + u0];
  if (!(!isnan(d0) && fabsf(warpped_d1 - d0) <= P.max_depth_delta && d0 > 0.0f)) return false;
  f = cur_I[i] - prev_I[v0 * W + u0];
  if (want_J) {
    float Gx = d0 * (u0 - P.cx) / P.fx, Gy = d0 * (v0 - P.cy) / P.fy, Gz = d0;
    float p0 = dI_dx * P.fx / Gz;
    float p1 = dI_dy * P.fy / Gz;
    float p2 = -(p0 * Gx + p1 * Gy) / Gz;
    J[0] = p0; J[1] = p1; J[2] = p2;
    J[3] = -Gz * p1 + Gy * p2;
    J[4] = Gz * p0 - Gx * p2;
    J[5] = -Gy * p0 + Gx * p1;
  }
  return true;
}

// one pyramid level + parameters of the photometric term, as passed to the fused evaluation kernel
struct RgbDev {
  const float *prev_I, *prev_D, *cur_I, *cur_D, *cur_G;
  int H, W;
  RgbParams P;             // k / kt are filled on the device from GnShared
  int robust;
  float robust_k;
  int on;
};

}  // namespace dfb
