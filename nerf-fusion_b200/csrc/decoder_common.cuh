// Front-end pieces shared by the FP32 and tcgen05 decoder engines: map lookup, pose transform, cube lattice.
#pragma once
#include "common.cuh"

namespace dfb {

struct MapDev {
  int nx, ny, nz;
  float bx, by, bz, vs, inv_vs;
  int div_mode;
  float ignore_th;
};

static inline MapDev to_dev(const dfb_map_params* p) {
  MapDev m;
  m.nx = p->nx; m.ny = p->ny; m.nz = p->nz;
  m.bx = p->bound_min[0]; m.by = p->bound_min[1]; m.bz = p->bound_min[2];
  m.vs = p->voxel_size; m.inv_vs = 1.0f / p->voxel_size;
  m.div_mode = p->div_mode; m.ignore_th = p->ignore_count_th;
  return m;
}

// map.py:566-576.  The reference does not bounds-check (map.py:313); out-of-grid points are reported invalid here.
__device__ __forceinline__ bool map_lookup(const MapDev& M, float px, float py, float pz, const int64_t* __restrict__ indexer,
                                           const float* __restrict__ obs_count, long long& slot, float rel[3]) {
  const float xn = div_vs(__fsub_rn(px, M.bx), M.vs, M.inv_vs, M.div_mode);
  const float yn = div_vs(__fsub_rn(py, M.by), M.vs, M.inv_vs, M.div_mode);
  const float zn = div_vs(__fsub_rn(pz, M.bz), M.vs, M.inv_vs, M.div_mode);
  const float cx = ceilf(xn) - 1.0f, cy = ceilf(yn) - 1.0f, cz = ceilf(zn) - 1.0f;   // exact in fp32 for grid-sized values
  if (!(cx >= 0.f && cx < (float)M.nx && cy >= 0.f && cy < (float)M.ny && cz >= 0.f && cz < (float)M.nz)) return false;
  const long long lin = (long long)cz + (long long)M.nz * (long long)cy + (long long)M.nz * M.ny * (long long)cx;
  slot = indexer[lin];
  if (slot < 0) return false;
  if (!(obs_count[slot] > M.ignore_th)) return false;
  rel[0] = __fsub_rn(__fsub_rn(xn, cx), 0.5f);
  rel[1] = __fsub_rn(__fsub_rn(yn, cy), 0.5f);
  rel[2] = __fsub_rn(__fsub_rn(zn, cz), 0.5f);
  return true;
}

struct PoseDev {
  float Rt[9], tt[3];   // total = last o delta
  float Rd[9], td[3];   // delta
  float Rl[9];          // last
};
static_assert(sizeof(PoseDev) == 33 * sizeof(float), "GnShared::pose_sdf holds a PoseDev image");

static inline PoseDev to_pose(const float* h_pose) {
  PoseDev P;
  for (int i = 0; i < 9; ++i) { P.Rt[i] = h_pose[i]; P.Rd[i] = h_pose[12 + i]; P.Rl[i] = h_pose[24 + i]; }
  for (int i = 0; i < 3; ++i) { P.tt[i] = h_pose[9 + i]; P.td[i] = h_pose[21 + i]; }
  return P;
}

__device__ __forceinline__ void xform(const float* R, const float* t, float x, float y, float z, float o[3]) {
  // other @ R^T + t  (motion_util.py:323-328)
#pragma unroll
  for (int j = 0; j < 3; ++j) o[j] = fmaf(z, R[3 * j + 2], fmaf(y, R[3 * j + 1], x * R[3 * j])) + t[j];
}

// tracker.py:199-205: J = [grad @ R_last^T, cross(T_delta p, .)]
__device__ __forceinline__ void sdf_jacobian(const PoseDev& P, const float gw[3], const float pc[3], float J[6]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) J[j] = fmaf(gw[2], P.Rl[3 * j + 2], fmaf(gw[1], P.Rl[3 * j + 1], gw[0] * P.Rl[3 * j]));
  float d[3];
  xform(P.Rd, P.td, pc[0], pc[1], pc[2], d);
  J[3] = d[1] * J[2] - d[2] * J[1];
  J[4] = d[2] * J[0] - d[0] * J[2];
  J[5] = d[0] * J[1] - d[1] * J[0];
}

// lattice coordinate of utility.get_samples minus the 0.5 network offset, fp32 op order of
// `(idx * vsize + a) - 0.5` (utility.py:143-147, map.py:646-647)
__device__ __forceinline__ float lattice(int i, float vsize, float a) {
  return __fsub_rn(__fadd_rn(__fmul_rn((float)i, vsize), a), 0.5f);
}

__device__ __forceinline__ float softplus_torch(float u) {   // F.softplus, beta 1, threshold 20
  return u > 20.f ? u : log1pf(expf(u));
}
__device__ __forceinline__ float softplus_grad(float u) { return u > 20.f ? 1.0f : 1.0f / (1.0f + expf(-u)); }

// which engine the decoder entry points use: 0 = FP32 CUDA cores, 1 = tcgen05 FP16 (default)
int decoder_engine();

// tcgen05 engine launchers (decoder_tc.cu); same contracts as the FP32 kernels in decoder.cu
int tc_decoder_explicit(const float* x, int n, const void* tc_blob, float* sdf, float* std_, cudaStream_t s);
int tc_get_sdf(const MapDev& M, const float* xyz, int n, const int64_t* indexer, const float* latents, const float* obs_count,
               const void* tc_blob, float* sdf, float* std_, uint8_t* valid, const float* g_sdf, const float* g_std,
               float* grad_xyz, cudaStream_t s);
int tc_sdf_hg(const MapDev& M, const PoseDev& P, const float* obs, int n, const int64_t* indexer, const float* latents,
              const float* obs_count, const void* blob, int robust, float robust_k, int with_J, double* packed, cudaStream_t s);
struct RgbDev;
namespace gn { struct StepArgs; }
int tc_gn_eval(const MapDev& M, const float* obs, int n, const int* n_dev, const int64_t* indexer, const float* latents, const float* obs_count,
               const void* blob, int robust, float robust_k, int with_J, const RgbDev& R, GnShared* gs, int gi, const gn::StepArgs& sa,
               cudaStream_t s);
int tc_cube_low(const float* latents, const int64_t* occ, int B, int r, float vsize, float a, const void* tc_blob, float* low_sdf,
                float* low_std, cudaStream_t s);
int tc_cube_refine(const float* latents, const int64_t* occ, int r, float vsize, float a, const void* tc_blob,
                   const int* refine_count, const long long* refine_list, float* cube_sdf, float* cube_std, cudaStream_t s);

}  // namespace dfb
