// Shared helpers for the difusion_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/difusion_b200.h"

namespace dfb {

void set_error(const char* fmt, ...);

#define DFB_CHECK_ARG(cond, msg)                  \
  do {                                            \
    if (!(cond)) {                                \
      dfb::set_error("invalid argument: %s", msg); \
      return DFB_E_INVALID;                       \
    }                                             \
  } while (0)

#define DFB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      dfb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DFB_E_CUDA;                                                                 \
    }                                                                                    \
  } while (0)

#define DFB_LAUNCH_CHECK()                                                                \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      dfb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DFB_E_CUDA;                                                                  \
    }                                                                                     \
  } while (0)

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Bump allocator over the caller's workspace.
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return off <= cap; }
};

int sm_count();

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Warp-aggregated append: every lane with `pred` gets a unique slot from *counter with ONE atomic per warp.
__device__ __forceinline__ int warp_append(int* counter, bool pred) {
  unsigned m = __ballot_sync(0xffffffffu, pred);
  int lane = threadIdx.x & 31;
  int base = 0;
  if (m) {
    int leader = __ffs(m) - 1;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
  }
  return base + __popc(m & ((1u << lane) - 1u));
}

// order-preserving float <-> uint map for atomicMin/Max on floats
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// x / vs with the selected torch semantics, never contracted.
__device__ __forceinline__ float div_vs(float x, float vs, float inv_vs, int div_mode) {
  return div_mode == DFB_DIV_IEEE ? __fdiv_rn(x, vs) : __fmul_rn(x, inv_vs);
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of int32 (3 kernels, any n).  block_sums: ceil(n/2048)+1 ints of scratch.
// total (sum of all) is written to *total if non-null.
// ------------------------------------------------------------------------------------------------
int exclusive_scan_i32(const int* in, int* out, int n, int* block_sums, int* total, cudaStream_t s, const int* n_dev = nullptr);
// same over popcounts of 32-bit words
int exclusive_scan_popc(const uint32_t* words, int* out, int n, int* block_sums, int* total, cudaStream_t s, const int* n_dev = nullptr);
// zero the first *n_dev (<= n_max) 32-bit words of p (16-byte aligned) without a host-side length
void zero_words_dev(void* p, int n_max, const int* n_dev, cudaStream_t s);

// ------------------------------------------------------------------------------------------------
// Gauss-Newton reduction helpers shared by the SDF and photometric terms
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float robust_w(float x, int kind, float k) {
  if (kind == 1) { float a = fabsf(x); return a > k ? k / a : 1.0f; }                       // tracker.py:61-66
  if (kind == 2) { float t = x / k; float s = 1.f - t * t; return fabsf(x) <= k ? s * s : 0.f; }  // :67-71
  return 1.0f;
}

// Block-level reduction of NV doubles per thread into out[] with one atomic per value per block.
template <int NV, int THREADS>
__device__ __forceinline__ void block_reduce_atomic(const float* vals, double* out) {
  __shared__ double red[THREADS / 32][NV];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double d = warp_sum((double)vals[k]);
    if (lane == 0) red[w][k] = d;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int ww = 0; ww < THREADS / 32; ++ww) s += red[ww][threadIdx.x];
    if (s != 0.0) atomicAdd(&out[threadIdx.x], s);
  }
}

// Accumulate sum w J J^T (upper triangle, 21), sum w r J (6), sum w r^2, count into acc[29].
// Per-voxel encoder accumulators (I6, map.py:446-449) are 64-bit FIXED-POINT sums (units of 2^-34): integer addition is
// associative, so the scatter-add gives the same bits whatever order the atomics land in -- run to run, and for any sharding of
// the samples over GPUs -- where float atomics (the reference's scatter, indexing.cu:59-71) leave the map different in its last
// bits every run.  One addend is rounded to 5.8e-11 absolute (its FP32 ulp is larger from |v| = 2^-10 up); the sum itself is
// exact, i.e. more accurate than a float running sum; range +-5e8.
constexpr float DFB_ACC_SCALE = 17179869184.0f;             // 2^34
__device__ __forceinline__ void acc_add(long long* p, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)__float2ll_rn(v * DFB_ACC_SCALE));
}
__device__ __forceinline__ float acc_read(long long v) { return __double2float_rn((double)v * (1.0 / 17179869184.0)); }

__device__ __forceinline__ void hg_accumulate(float* acc, const float* J, float r, float w, bool with_J) {
  if (with_J) {
    int t = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = a; b < 6; ++b) acc[t++] += w * J[a] * J[b];
#pragma unroll
    for (int a = 0; a < 6; ++a) acc[21 + a] += w * r * J[a];
  }
  acc[27] += w * r * r;
  acc[28] += 1.0f;
}


// expands the packed 29 doubles (21 upper-tri, 6, 1, 1) into the public 44-double layout
void launch_hg_expand(const double* packed, double* out44, cudaStream_t s);

// ---- device-resident Gauss-Newton state (gauss_newton.cu) ------------------------------------------------------------
// The term kernels of a Gauss-Newton evaluation read their pose from here and add their packed sums here; the
// one-warp step that follows them (gn_step.cuh) consumes the sums, solves, updates the pose and publishes the pose blocks
// of the next evaluation, so the host never has to read H, g back between iterations.
struct GnShared {
  double sums[2][32];        // packed 29 sums of the SDF term [0] and the photometric term [1]; zero between evaluations
  double delta[12], last_delta[12], last[12];   // poses: R row-major (9), t (3)
  double last_energy;
  double intr[4];            // fx, fy, cx, cy
  double kinv[4];            // 1/fx, 1/fy, -cx/fx, -cy/fy (entries of K^-1)
  float pose_sdf[36];        // PoseDev image of the next SDF evaluation (33 floats used)
  float krk[12], kt[4];      // K R K^-1 (9 used), K t (3 used) of the next photometric evaluation
  int done[8];               // group i finished (converged, rolled back or failed): its remaining launches return at once
  int error;                 // 1 = singular normal equations
  int ticket;                // blocks of the evaluation's last term kernel that have finished (last-block-done step)
  int rgb_cursor;            // next chunk of photometric pixels (work stealing inside the fused evaluation kernel)
  int pad_;
};
int launch_sdf_hg_gn(const dfb_map_params* h_params, const float* obs_xyz, int n, const int32_t* n_dev, const int64_t* indexer, const float* latent_vecs,
                     const float* voxel_obs_count, const float* decoder_blob, int robust, float robust_k, int compute_J, GnShared* gs, int gi,
                     cudaStream_t s);                                                                        // decoder.cu
int launch_rgb_hg_gn(const dfb_rgb_level* L, const float* intr4, float min_grad_scale, float max_depth_delta, int robust, float robust_k,
                     int compute_J, GnShared* gs, int gi, cudaStream_t s);                                   // photometric.cu
namespace gn { struct StepArgs; }
// fused evaluation (tcgen05 engine): SDF term + optional photometric term (L may be null) + step, one launch      // decoder.cu / decoder_tc.cu
int launch_sdf_rgb_gn(const dfb_map_params* h_params, const float* obs_xyz, int n, const int32_t* n_dev, const int64_t* indexer, const float* latent_vecs,
                      const float* voxel_obs_count, const float* decoder_blob, int sdf_robust, float sdf_robust_k, int compute_J,
                      const dfb_rgb_level* L, const float* intr4, float min_grad_scale, float max_depth_delta, int rgb_robust,
                      float rgb_robust_k, GnShared* gs, int gi, const gn::StepArgs* sa, cudaStream_t s);
// photometric term + step, one launch                                                                                // photometric.cu
int launch_rgb_step_gn(const dfb_rgb_level* L, const float* intr4, float min_grad_scale, float max_depth_delta, int robust, float robust_k,
                       int compute_J, GnShared* gs, int gi, const gn::StepArgs* sa, cudaStream_t s);
// Written by the step into pinned host memory, polled by the driver.  The first 16 bytes go out as ONE vector store
// (seq and check bracket the payload, so the host accepts the header only when both 8-byte halves have landed); delta is
// written -- and fenced system-wide -- before the header, and only by the step that ends its group (GN_HAS_DELTA).
struct alignas(16) GnRecord {
  int seq;                   // sequence number of the evaluation
  int flags;                 // GN_EXECUTED | GN_BROKE | GN_ERROR | GN_HAS_DELTA
  float cnt0;                // valid count of the SDF term (exact below 2^24)
  int check;                 // seq ^ GN_CHECK
  double delta[12];          // pose after this step (valid when GN_HAS_DELTA)
};
constexpr int GN_EXECUTED = 1, GN_BROKE = 2, GN_ERROR = 4, GN_HAS_DELTA = 8, GN_CHECK = 0x5a5a5a5a;

}  // namespace dfb
