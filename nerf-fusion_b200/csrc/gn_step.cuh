// The Gauss-Newton step (tracker.py:240-281) as a warp-level device function, so it can run as the tail of the last
// term kernel of an evaluation (last-block-done) as well as in its own small kernel.  Pose algebra restated from
// utils/motion_util.py:205-228 (from_twist), :275-279 (inv, dot).
#pragma once
#include "common.cuh"

namespace dfb {
namespace gn {

struct StepArgs {
  GnRecord* ring;        // device view of the pinned record ring
  int seq, gi, step, n_it, use_sdf, use_rgb;
  double rgb_weight;
};
constexpr int STEP_SCRATCH_BYTES = 2048;   // shared-memory scratch the step needs (8-byte aligned)

#ifdef DFB_TC_PROFILE
static __device__ unsigned long long g_step_prof[16];
#define STEP_MARK(i) do { if ((threadIdx.x & 31) == 0) g_step_prof[(a.step & 1) * 8 + (i)] = clock64(); } while (0)   /* even / odd steps kept apart */
#else
#define STEP_MARK(i) do {} while (0)
#endif

// All of this runs in one device thread (gn_step_kernel); loops have constant bounds and are unrolled so poses and the
// 6x7 elimination tableau stay in registers (no local-memory traffic on the critical path between two evaluations).
struct Pose {   // x -> R x + t, float64
  double R[9], t[3];
};

__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ __forceinline__ void mat3_vec(const double* A, const double* v, double* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
__device__ __forceinline__ Pose compose(const Pose& a, const Pose& b) {   // a o b  (Isometry.dot)
  Pose c;
  mat3_mul(a.R, b.R, c.R);
  double rt[3];
  mat3_vec(a.R, b.t, rt);
#pragma unroll
  for (int i = 0; i < 3; ++i) c.t[i] = rt[i] + a.t[i];
  return c;
}

// The reference stores rotations as unit quaternions (pyquaternion normalises on every rotation_matrix access), which
// re-orthonormalises the pose each iteration; do the same round trip.
__device__ __forceinline__ void renormalise(double* R) {
  // matrix -> quaternion -> unit quaternion -> matrix; divisions by a common value are one reciprocal and multiplies
  // (the step sits between two evaluations: its latency is paid once per Gauss-Newton iteration)
  double q0, q1, q2, q3;
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    const double s = sqrt(tr + 1.0) * 2, is = 1.0 / s;
    q0 = 0.25 * s; q1 = (R[7] - R[5]) * is; q2 = (R[2] - R[6]) * is; q3 = (R[3] - R[1]) * is;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2, is = 1.0 / s;
    q0 = (R[7] - R[5]) * is; q1 = 0.25 * s; q2 = (R[1] + R[3]) * is; q3 = (R[2] + R[6]) * is;
  } else if (R[4] > R[8]) {
    const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2, is = 1.0 / s;
    q0 = (R[2] - R[6]) * is; q1 = (R[1] + R[3]) * is; q2 = 0.25 * s; q3 = (R[5] + R[7]) * is;
  } else {
    const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2, is = 1.0 / s;
    q0 = (R[3] - R[1]) * is; q1 = (R[2] + R[6]) * is; q2 = (R[5] + R[7]) * is; q3 = 0.25 * s;
  }
  const double in = rsqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  const double w = q0 * in, x = q1 * in, y = q2 * in, z = q3 * in;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}

__device__ __forceinline__ Pose from_twist(const double* xi) {   // motion_util.py:205-228
  Pose p;
  const double* rho = xi;
  const double* phi = xi + 3;
  const double angle = sqrt(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
  const double Wd[9] = {0, -phi[2], phi[1], phi[2], 0, -phi[0], -phi[1], phi[0], 0};
  double J[9];
  if (fabs(angle) <= 1e-8) {   // np.isclose(angle, 0.)
#pragma unroll
    for (int i = 0; i < 9; ++i) { p.R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + Wd[i]; J[i] = ((i % 4 == 0) ? 1.0 : 0.0) + 0.5 * Wd[i]; }
  } else {
    const double ia = 1.0 / angle;
    const double ax[3] = {phi[0] * ia, phi[1] * ia, phi[2] * ia};
    double s, c;
    sincos(angle, &s, &c);
    const double Wa[9] = {0, -ax[2], ax[1], ax[2], 0, -ax[0], -ax[1], ax[0], 0};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double I = (i == j) ? 1.0 : 0.0, oo = ax[i] * ax[j];
        p.R[3 * i + j] = c * I + (1 - c) * oo + s * Wa[3 * i + j];
        J[3 * i + j] = (s * ia) * I + (1 - s * ia) * oo + ((1 - c) * ia) * Wa[3 * i + j];
      }
  }
  // (pyquaternion would normalise the quaternion of this matrix here; Rodrigues' formula already gives a rotation to
  // rounding, and the composed pose is renormalised by the caller)
  mat3_vec(J, rho, p.t);
  return p;
}

// np.linalg.solve(H, -g): LU with partial pivoting, float64, by one warp.  Lane j < 7 holds column j of the 6 x 7 tableau
// [H | -g] in six registers (lanes 7.. shadow lane 6); the pivot choice and the five row factors of a column come from
// lane c by shuffle, every lane updates its own column.  Same operations per element as the sequential elimination;
// the solution ends up on every lane.  Returns false when singular / non-finite.
__device__ __forceinline__ bool solve6_warp(const double* H, const double* g, double* x) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int j = lane < 7 ? lane : 6;
  double a[6], ipiv[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) a[r] = (j < 6) ? H[6 * r + j] : -g[r];
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    double best = fabs(a[c]);
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      const double v = fabs(a[r]);
      if (v > best) { best = v; piv = r; }
    }
    piv = __shfl_sync(FULL, piv, c);
    best = __shfl_sync(FULL, best, c);
    if (!(best > 0.0)) ok = false;
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {                    // swap rows c and piv: selects on registers, no indexing
      const bool sw = (piv == r);
      const double ac = a[c], ar = a[r];
      a[c] = sw ? ar : ac;
      a[r] = sw ? ac : ar;
    }
    // one reciprocal per pivot, then multiplies (what LAPACK's dgetf2 does: DSCAL by 1 / A(j,j)); a float64 division costs
    // several hundred cycles here and the step sits between two evaluations
    const double ipc = 1.0 / __shfl_sync(FULL, a[c], c);
    ipiv[c] = ipc;
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      const double f = __shfl_sync(FULL, a[r], c) * ipc;
      a[r] -= f * a[c];
    }
  }
#pragma unroll
  for (int r = 5; r >= 0; --r) {
    double sacc = __shfl_sync(FULL, a[r], 6);
#pragma unroll
    for (int jj = r + 1; jj < 6; ++jj) sacc -= __shfl_sync(FULL, a[r], jj) * x[jj];
    x[r] = sacc * ipiv[r];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
    if (!(fabs(x[i]) <= 1.79769313486231570e308)) ok = false;   // non-finite
  return ok;
}

// float images of the current pose for the two term kernels (same casts as the host-side wrappers: tracker.py:145-147, :201)
__device__ __forceinline__ void publish_pose(GnShared* gs) {
  Pose last, delta;
  #pragma unroll
  for (int i = 0; i < 9; ++i) { last.R[i] = gs->last[i]; delta.R[i] = gs->delta[i]; }
  #pragma unroll
  for (int i = 0; i < 3; ++i) { last.t[i] = gs->last[9 + i]; delta.t[i] = gs->delta[9 + i]; }
  const Pose total = compose(last, delta);
  float* hp = gs->pose_sdf;                                   // PoseDev: Rt(9) tt(3) Rd(9) td(3) Rl(9)
  #pragma unroll
  for (int i = 0; i < 9; ++i) { hp[i] = (float)total.R[i]; hp[12 + i] = (float)delta.R[i]; hp[24 + i] = (float)last.R[i]; }
  #pragma unroll
  for (int i = 0; i < 3; ++i) { hp[9 + i] = (float)total.t[i]; hp[21 + i] = (float)delta.t[i]; }
  const double fx = gs->intr[0], fy = gs->intr[1], cx = gs->intr[2], cy = gs->intr[3];
  const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  const double Kinv[9] = {gs->kinv[0], 0, gs->kinv[2], 0, gs->kinv[1], gs->kinv[3], 0, 0, 1};   // 1/fx, 1/fy, -cx/fx, -cy/fy (init kernel)
  double KR[9], KRK[9], Kt[3];
  mat3_mul(K, delta.R, KR); mat3_mul(KR, Kinv, KRK); mat3_vec(K, delta.t, Kt);
  #pragma unroll
  for (int i = 0; i < 9; ++i) gs->krk[i] = (float)KRK[i];
  #pragma unroll
  for (int i = 0; i < 3; ++i) gs->kt[i] = (float)Kt[i];
}


// Record for the host (pinned memory).  When the step ends its group the pose goes out first and is fenced system-wide;
// the 16-byte header is one vector store (see GnRecord).
__device__ __forceinline__ void write_record(GnRecord* rec, int seq, const double* delta, int flags, double cnt0) {
  const int lane = threadIdx.x & 31;
  if (flags & GN_HAS_DELTA) {
    if (lane < 12) rec->delta[lane] = delta[lane];
    __threadfence_system();
    __syncwarp();
  }
  if (lane == 0) {
    const int c0 = __float_as_int((float)cnt0), chk = seq ^ GN_CHECK;
    asm volatile("st.volatile.global.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(rec), "r"(seq), "r"(flags), "r"(c0), "r"(chk) : "memory");
  }
}

// A launch of a finished group still owes the host its record (one warp).
__device__ __forceinline__ void skip_record(const GnShared* gs, const StepArgs& a) {
  write_record(a.ring + (a.seq & 3), a.seq, nullptr, gs->error ? GN_ERROR : 0, 0.0);
}

// One Gauss-Newton step for group a.gi, iteration a.step (a.step == a.n_it is the evaluation-only pass, i_iter = -1).
// Called by ONE full warp after every term kernel block of the evaluation has added its sums (stream order, or the
// last-block-done ticket).  The state is staged through `scratch` (shared memory: one round of global loads and one
// of stores, all lanes), the 36 + 6 normal-equation entries are scaled and summed lane-parallel, lane 0 runs the float64
// solve and pose update out of registers.  Also re-arms the per-evaluation counters (ticket, rgb_cursor, sums).
static __device__ __noinline__ void step_warp(GnShared* gs, const StepArgs a, void* scratch) {
  static_assert(sizeof(GnShared) % 8 == 0, "GnShared is copied as doubles");
  static_assert(sizeof(GnShared) + 44 * 8 + 16 <= STEP_SCRATCH_BYTES, "scratch too small");
  constexpr int ND = sizeof(GnShared) / 8;
  GnShared& sh = *reinterpret_cast<GnShared*>(scratch);
  double* Hs = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + sizeof(GnShared));
  double* gsv = Hs + 36;
  int* flags = reinterpret_cast<int*>(gsv + 8);
  const int lane = threadIdx.x & 31;
  STEP_MARK(0);
  {
    // L2 loads (the sums were produced by atomics of other blocks), all in flight before the first shared-memory store:
    // interleaved with the stores the compiler must keep them in order (possible aliasing) and pays one L2 latency each
    const double* src = reinterpret_cast<const double*>(gs);
    double* dst = reinterpret_cast<double*>(&sh);
    constexpr int PER_LANE = (ND + 31) / 32;
    double v[PER_LANE];
#pragma unroll
    for (int k = 0; k < PER_LANE; ++k) v[k] = (lane + 32 * k < ND) ? __ldcg(src + lane + 32 * k) : 0.0;
#pragma unroll
    for (int k = 0; k < PER_LANE; ++k)
      if (lane + 32 * k < ND) dst[lane + 32 * k] = v[k];
  }
  __syncwarp();
  STEP_MARK(1);
  const int gi = a.gi;
  const bool run = !sh.done[gi];
  const bool no_grad = (a.step == a.n_it);
  double cnt0 = 0.0, cnt1 = 0.0;
  if (lane == 0) flags[1] = 0;
  if (run) {
    cnt0 = a.use_sdf ? sh.sums[0][28] : 0.0;
    cnt1 = a.use_rgb ? sh.sums[1][28] : 0.0;
    const double scale0 = 1.0 / cnt0, scale1 = a.rgb_weight / cnt1;   // tracker.py:215 / :170 (inf/NaN when nothing is valid, like 1/0 there)
    if (!no_grad) {
      for (int e = lane; e < 42; e += 32) {                          // SDF term first, then the photometric term (tracker.py:248-262)
        int idx;
        if (e < 36) {
          const int r = e / 6, c = e % 6, lo = r < c ? r : c, hi = r < c ? c : r;
          idx = lo * 6 - lo * (lo - 1) / 2 + (hi - lo);
        } else {
          idx = 21 + (e - 36);
        }
        double v = 0.0;
        if (a.use_sdf) v += sh.sums[0][idx] * scale0;
        if (a.use_rgb) v += sh.sums[1][idx] * scale1;
        if (e < 36) Hs[e] = v; else gsv[e - 36] = v;
      }
    }
    __syncwarp();
    STEP_MARK(2);
    // decision (uniform over the warp: every lane reads the same shared values)
    double energy = 0.0;
    if (a.use_sdf) energy += sh.sums[0][27] * scale0;
    if (a.use_rgb) energy += sh.sums[1][27] * scale1;
    const double last_energy = a.step == 0 ? CUDART_INF : sh.last_energy;
    const bool rollback = energy > last_energy;                     // tracker.py:269-271: roll back, leave the group
    double xi[6];
    bool solved = true;
    if (!rollback && !no_grad) solved = solve6_warp(Hs, gsv, xi);   // all lanes
    STEP_MARK(3);
    __syncwarp();
    if (lane == 0) {
      int broke = 0;
      if (rollback) {
#pragma unroll
        for (int i = 0; i < 12; ++i) sh.delta[i] = sh.last_delta[i];
        sh.done[gi] = 1;
        broke = 1;
      } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) sh.last_delta[i] = sh.delta[i];
        sh.last_energy = energy;
        if (!no_grad) {
          if (!solved) {
            sh.error = 1;
#pragma unroll
            for (int i = 0; i < 8; ++i) sh.done[i] = 1;
          } else {
            Pose d;
#pragma unroll
            for (int i = 0; i < 9; ++i) d.R[i] = sh.delta[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) d.t[i] = sh.delta[9 + i];
            Pose nd = compose(from_twist(xi), d);                     // tracker.py:277-278
            renormalise(nd.R);
#pragma unroll
            for (int i = 0; i < 9; ++i) sh.delta[i] = nd.R[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) sh.delta[9 + i] = nd.t[i];
            STEP_MARK(4);
          }
        } else {
          sh.done[gi] = 1;                                            // the evaluation-only pass closes the group
        }
      }
      publish_pose(&sh);
      STEP_MARK(5);
      flags[1] = broke;
    }
    __syncwarp();
    for (int i = lane; i < 64; i += 32) reinterpret_cast<double*>(sh.sums)[i] = 0.0;   // zero-invariant for the next evaluation
  }
  if (lane == 0) { sh.ticket = 0; sh.rgb_cursor = 0; }
  __syncwarp();
  {
    double* dst = reinterpret_cast<double*>(gs);
    const double* src = reinterpret_cast<const double*>(&sh);
    for (int i = lane; i < ND; i += 32) dst[i] = src[i];
  }
  STEP_MARK(6);
  const int rflags = (run ? GN_EXECUTED : 0) | (flags[1] ? GN_BROKE : 0) | (sh.error ? GN_ERROR : 0) |
                     ((run && sh.done[gi]) ? GN_HAS_DELTA : 0);        // the step that ends its group carries the pose
  write_record(a.ring + (a.seq & 3), a.seq, sh.delta, rflags, cnt0);
  STEP_MARK(7);
}

// Tail of a term kernel: the last block of the grid to arrive runs the step (classic threadfence reduction).  Every
// thread that issued atomics into gs->sums must have executed __threadfence() before the block-wide barrier that
// precedes this call.  `flag` is a shared-memory int.
__device__ __forceinline__ void tail_step(GnShared* gs, const StepArgs& a, void* scratch, int* flag) {
  if (threadIdx.x == 0) {
    __threadfence();
    *flag = (atomicAdd(&gs->ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (*flag && threadIdx.x < 32) {
    __threadfence();
    step_warp(gs, a, scratch);
  }
}

}  // namespace gn
}  // namespace dfb
