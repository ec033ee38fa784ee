// Native Gauss-Newton driver: system/tracker.py:225-288 (gauss_newton) with the two fused terms, device-resident.
// One C call runs the whole pose solve of a frame.  An evaluation ("slot") is up to three launches on the caller's stream:
//   [SDF term kernel] [photometric term kernel] [gn_step_kernel <<<1, 32>>>]
// The term kernels read their pose from GnShared and add their packed sums there; the step kernel scales and sums the
// terms, applies the accept / rollback rule, solves the 6x6 system (float64 LU, partial pivoting), composes the SE(3)
// update, publishes the pose blocks of the next evaluation and writes a small record into pinned host memory.  The host
// keeps ONE slot of look-ahead: slot k+1 is enqueued before the record of slot k is read, so the GPU never waits for
// the host between evaluations; when slot k ends its group (energy rose -> rollback, tracker.py:269-271) the launches
// of slot k+1 see done[group] and return immediately.
// Pose algebra restated from utils/motion_util.py:205-228 (from_twist), :275-279 (inv, dot).
#include <math.h>
#include <string.h>
#include <chrono>
#include <vector>

#include "common.cuh"

namespace {

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
  double q0, q1, q2, q3;
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    const double s = sqrt(tr + 1.0) * 2;
    q0 = 0.25 * s; q1 = (R[7] - R[5]) / s; q2 = (R[2] - R[6]) / s; q3 = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
    q0 = (R[7] - R[5]) / s; q1 = 0.25 * s; q2 = (R[1] + R[3]) / s; q3 = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
    q0 = (R[2] - R[6]) / s; q1 = (R[1] + R[3]) / s; q2 = 0.25 * s; q3 = (R[5] + R[7]) / s;
  } else {
    const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
    q0 = (R[3] - R[1]) / s; q1 = (R[2] + R[6]) / s; q2 = (R[5] + R[7]) / s; q3 = 0.25 * s;
  }
  const double n = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  const double w = q0 / n, x = q1 / n, y = q2 / n, z = q3 / n;
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
    const double ax[3] = {phi[0] / angle, phi[1] / angle, phi[2] / angle};
    double s, c;
    sincos(angle, &s, &c);
    const double Wa[9] = {0, -ax[2], ax[1], ax[2], 0, -ax[0], -ax[1], ax[0], 0};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double I = (i == j) ? 1.0 : 0.0, oo = ax[i] * ax[j];
        p.R[3 * i + j] = c * I + (1 - c) * oo + s * Wa[3 * i + j];
        J[3 * i + j] = (s / angle) * I + (1 - s / angle) * oo + ((1 - c) / angle) * Wa[3 * i + j];
      }
  }
  renormalise(p.R);
  mat3_vec(J, rho, p.t);
  return p;
}

// np.linalg.solve(H, -g): LU with partial pivoting, float64.  Returns false when singular / non-finite.
__device__ __forceinline__ bool solve6(const double* H, const double* g, double* x) {
  double A[42];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j < 6; ++j) A[7 * (i) + (j)] = H[6 * i + j];
    A[7 * (i) + (6)] = -g[i];
  }
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    double best = fabs(A[7 * (c) + (c)]);
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      const double v = fabs(A[7 * (r) + (c)]);
      if (v > best) { best = v; piv = r; }
    }
    if (!(best > 0.0)) ok = false;
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      if (piv == r) {
#pragma unroll
        for (int j = 0; j < 7; ++j) { const double t = A[7 * (c) + (j)]; A[7 * (c) + (j)] = A[7 * (r) + (j)]; A[7 * (r) + (j)] = t; }
      }
    }
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      const double f = A[7 * (r) + (c)] / A[7 * (c) + (c)];
#pragma unroll
      for (int j = c; j < 7; ++j) A[7 * (r) + (j)] -= f * A[7 * (c) + (j)];
    }
  }
#pragma unroll
  for (int r = 5; r >= 0; --r) {
    double sacc = A[7 * (r) + (6)];
#pragma unroll
    for (int j = r + 1; j < 6; ++j) sacc -= A[7 * (r) + (j)] * x[j];
    x[r] = sacc / A[7 * (r) + (r)];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
    if (!(fabs(x[i]) <= 1.79769313486231570e308)) ok = false;   // non-finite
  return ok;
}

// float images of the current pose for the two term kernels (same casts as the host-side wrappers: tracker.py:145-147, :201)
__device__ void publish_pose(dfb::GnShared* gs) {
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
  const double Kinv[9] = {1 / fx, 0, -cx / fx, 0, 1 / fy, -cy / fy, 0, 0, 1};
  double KR[9], KRK[9], Kt[3];
  mat3_mul(K, delta.R, KR); mat3_mul(KR, Kinv, KRK); mat3_vec(K, delta.t, Kt);
  #pragma unroll
  for (int i = 0; i < 9; ++i) gs->krk[i] = (float)KRK[i];
  #pragma unroll
  for (int i = 0; i < 3; ++i) gs->kt[i] = (float)Kt[i];
}

struct GnInit {
  double last[12], delta[12], intr[4];
};

__global__ void gn_init_kernel(dfb::GnShared* gs, GnInit in) {
  const int t = threadIdx.x;
  if (t < 64) reinterpret_cast<double*>(gs->sums)[t] = 0.0;
  if (t < 12) { gs->last[t] = in.last[t]; gs->delta[t] = in.delta[t]; gs->last_delta[t] = in.delta[t]; }
  if (t < 4) gs->intr[t] = in.intr[t];
  if (t < 8) gs->done[t] = 0;
  if (t == 0) { gs->error = 0; gs->last_energy = CUDART_INF; }
  __syncthreads();
  if (t == 0) publish_pose(gs);
}

// One Gauss-Newton step (the body of the loop at tracker.py:240-281) for group gi, iteration `step` (step == n_it is the
// evaluation-only pass, i_iter = -1).  One warp: the state is staged through shared memory (one round of global loads
// and one of stores, all lanes), the 36 + 6 normal-equation entries are scaled and summed lane-parallel, and lane 0 runs
// the float64 solve and pose update out of registers.
__global__ void __launch_bounds__(32) gn_step_kernel(dfb::GnShared* __restrict__ gs, dfb::GnRecord* __restrict__ ring, int seq, int gi, int step,
                                                     int n_it, int use_sdf, int use_rgb, double rgb_weight) {
  __shared__ dfb::GnShared sh;
  __shared__ double Hs[36], gsv[6];
  __shared__ int flags[2];                       // executed, broke
  static_assert(sizeof(dfb::GnShared) % 8 == 0, "GnShared is copied as doubles");
  constexpr int ND = sizeof(dfb::GnShared) / 8;
  const int lane = threadIdx.x;
  dfb::GnRecord* rec = ring + (seq & 3);
  {
    const double* src = reinterpret_cast<const double*>(gs);
    double* dst = reinterpret_cast<double*>(&sh);
    for (int i = lane; i < ND; i += 32) dst[i] = src[i];
  }
  __syncwarp();
  const bool run = !sh.done[gi];
  const bool no_grad = (step == n_it);
  double cnt0 = 0.0, cnt1 = 0.0;
  if (run) {
    cnt0 = use_sdf ? sh.sums[0][28] : 0.0;
    cnt1 = use_rgb ? sh.sums[1][28] : 0.0;
    const double scale0 = 1.0 / cnt0, scale1 = rgb_weight / cnt1;   // tracker.py:215 / :170 (inf/NaN when nothing is valid, like 1/0 there)
    if (!no_grad) {
      for (int e = lane; e < 42; e += 32) {                          // SDF term first, then the photometric term (tracker.py:248-262)
        int idx;
        if (e < 36) {
          const int a = e / 6, b = e % 6, lo = a < b ? a : b, hi = a < b ? b : a;
          idx = lo * 6 - lo * (lo - 1) / 2 + (hi - lo);
        } else {
          idx = 21 + (e - 36);
        }
        double v = 0.0;
        if (use_sdf) v += sh.sums[0][idx] * scale0;
        if (use_rgb) v += sh.sums[1][idx] * scale1;
        if (e < 36) Hs[e] = v; else gsv[e - 36] = v;
      }
    }
    __syncwarp();
    if (lane == 0) {
      double energy = 0.0;
      if (use_sdf) energy += sh.sums[0][27] * scale0;
      if (use_rgb) energy += sh.sums[1][27] * scale1;
      int broke = 0;
      const double last_energy = step == 0 ? CUDART_INF : sh.last_energy;
      if (energy > last_energy) {                                     // tracker.py:269-271: roll back, leave the group
#pragma unroll
        for (int i = 0; i < 12; ++i) sh.delta[i] = sh.last_delta[i];
        sh.done[gi] = 1;
        broke = 1;
      } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) sh.last_delta[i] = sh.delta[i];
        sh.last_energy = energy;
        if (!no_grad) {
          double xi[6];
          if (!solve6(Hs, gsv, xi)) {
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
          }
        } else {
          sh.done[gi] = 1;                                            // the evaluation-only pass closes the group
        }
      }
      publish_pose(&sh);
      flags[1] = broke;
    }
    __syncwarp();
    for (int i = lane; i < 64; i += 32) reinterpret_cast<double*>(sh.sums)[i] = 0.0;   // zero-invariant for the next evaluation
    __syncwarp();
    double* dst = reinterpret_cast<double*>(gs);
    const double* src = reinterpret_cast<const double*>(&sh);
    for (int i = lane; i < ND; i += 32) dst[i] = src[i];
  }
  // record for the host (pinned memory): payload from all lanes, then a system-wide fence, then the sequence number
  if (lane < 12) rec->delta[lane] = sh.delta[lane];
  if (lane == 12) { rec->executed = run ? 1 : 0; rec->broke = run ? flags[1] : 0; rec->error = sh.error; }
  if (lane == 13) { rec->cnt[0] = cnt0; rec->cnt[1] = cnt1; }
  __threadfence_system();                        // every lane: its payload stores are visible system-wide ...
  __syncwarp();                                  // ... before lane 0 publishes the sequence number
  if (lane == 0) *reinterpret_cast<volatile int*>(&rec->seq) = seq;
}

}  // namespace

extern "C" int dfb_gauss_newton(const dfb_map_params* h_params, const dfb_gn_config* h_cfg, const float* obs_xyz, int n,
                                const int64_t* indexer, const float* latent_vecs, const float* voxel_obs_count,
                                const float* decoder_blob, const dfb_rgb_level* h_levels, const double* h_intr,
                                const double* h_last_pose, double* h_delta_pose, double* d_scratch, double* h_pinned,
                                int32_t* h_stats, void* stream) {
  using dfb::GnRecord;
  using dfb::GnShared;
  static_assert(sizeof(GnShared) <= DFB_GN_SCRATCH_DOUBLES * sizeof(double), "d_scratch too small");
  static_assert(4 * sizeof(GnRecord) <= DFB_GN_PINNED_DOUBLES * sizeof(double), "h_pinned too small");
  // optional kernel timing (CUDA events on the launching stream around every SDF-term launch), reported through
  // h_stats[4..6]: enabled when h_stats[4] == 0x54494d45 ('TIME') on entry
  const bool timing = h_stats && h_stats[4] == 0x54494d45;
  DFB_CHECK_ARG(h_params && h_cfg && h_last_pose && h_delta_pose && d_scratch && h_pinned && h_stats, "gauss_newton");
  DFB_CHECK_ARG(h_cfg->n_groups >= 0 && h_cfg->n_groups <= 8, "gauss_newton: n_groups must be in [0, 8]");
  cudaStream_t s = (cudaStream_t)stream;
  GnShared* gs = reinterpret_cast<GnShared*>(d_scratch);
  GnRecord* ring = nullptr;                                        // device view of the pinned ring
  DFB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ring), h_pinned, 0));
  volatile GnRecord* hring = reinterpret_cast<volatile GnRecord*>(h_pinned);
  for (int i = 0; i < 4; ++i) hring[i].seq = 0;

  static std::vector<cudaEvent_t> events;                           // pairs, grown on demand (timing only)
  auto event = [&](size_t i) -> cudaEvent_t {
    while (events.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); events.push_back(e); }
    return events[i];
  };

  GnInit init;
  memcpy(init.last, h_last_pose, sizeof(double) * 12); memcpy(init.delta, h_delta_pose, sizeof(double) * 12);
  for (int i = 0; i < 4; ++i) init.intr[i] = h_intr ? h_intr[i] : (i < 2 ? 1.0 : 0.0);
  gn_init_kernel<<<1, 64, 0, s>>>(gs, init);
  DFB_LAUNCH_CHECK();
  const float intr4[4] = {(float)init.intr[0], (float)init.intr[1], (float)init.intr[2], (float)init.intr[3]};

  static int seq_base = 0;                                          // records carry a process-unique, non-zero sequence number
  struct Slot { int seq, gi, step, sdf_event; };
  int n_sdf = 0, n_rgb = 0, n_steps = 0, i_iter = 0, n_events = 0, error = 0;
  double sdf_ms = 0.0, sdf_q_j = 0.0, sdf_q_nj = 0.0;
  double delta_out[12];
  memcpy(delta_out, h_delta_pose, sizeof(delta_out));

  auto enqueue = [&](int gi, int step, Slot& out) -> int {
    const int n_it = h_cfg->n_iter[gi];
    const bool no_grad = (step == n_it);
    const int lvl = h_cfg->rgb_level[gi];
    out.gi = gi; out.step = step; out.sdf_event = -1;
    if (++seq_base == 0 || seq_base == INT32_MAX) seq_base = 1;
    out.seq = seq_base;
    if (h_cfg->use_sdf[gi]) {
      if (timing) { out.sdf_event = n_events; cudaEventRecord(event(2 * n_events), s); }
      int rc = dfb::launch_sdf_hg_gn(h_params, obs_xyz, n, indexer, latent_vecs, voxel_obs_count, decoder_blob, h_cfg->sdf_robust,
                                     h_cfg->sdf_robust_k, no_grad ? 0 : 1, gs, gi, s);
      if (rc) return rc;
      if (timing) { cudaEventRecord(event(2 * n_events + 1), s); ++n_events; }
    }
    if (lvl >= 0) {
      int rc = dfb::launch_rgb_hg_gn(&h_levels[lvl], intr4, h_cfg->rgb_min_grad_scale, h_cfg->rgb_max_depth_delta, h_cfg->rgb_robust,
                                     h_cfg->rgb_robust_k, no_grad ? 0 : 1, gs, gi, s);
      if (rc) return rc;
    }
    gn_step_kernel<<<1, 32, 0, s>>>(gs, ring, out.seq, gi, step, n_it, h_cfg->use_sdf[gi] ? 1 : 0, lvl >= 0 ? 1 : 0, (double)h_cfg->rgb_weight);
    DFB_LAUNCH_CHECK();
    return DFB_OK;
  };
  // wait for the record of a slot (spin on pinned memory; the stream is polled for errors now and then)
  auto wait = [&](const Slot& sl, GnRecord& r) -> int {
    volatile GnRecord* rec = hring + (sl.seq & 3);
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spin = 0;; ++spin) {
      if (rec->seq == sl.seq) { __atomic_thread_fence(__ATOMIC_ACQUIRE); break; }
      if ((spin & 0xfff) == 0xfff) {
        cudaError_t e = cudaStreamQuery(s);
        if (e != cudaSuccess && e != cudaErrorNotReady) { dfb::set_error("gauss_newton: %s", cudaGetErrorString(e)); return DFB_E_CUDA; }
        if (e == cudaSuccess && rec->seq != sl.seq) {               // stream drained without the record: cannot happen unless a launch failed
          if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2)) { dfb::set_error("gauss_newton: step record never arrived"); return DFB_E_CUDA; }
        }
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) { dfb::set_error("gauss_newton: timed out waiting for the device"); return DFB_E_CUDA; }
      }
    }
    r.executed = rec->executed; r.broke = rec->broke; r.error = rec->error;
    r.cnt[0] = rec->cnt[0]; r.cnt[1] = rec->cnt[1];
    for (int i = 0; i < 12; ++i) r.delta[i] = rec->delta[i];
    return DFB_OK;
  };
  auto account = [&](const Slot& sl, const GnRecord& r) {
    if (!r.executed) return;
    const int n_it = h_cfg->n_iter[sl.gi];
    const bool no_grad = (sl.step == n_it);
    i_iter = no_grad ? -1 : sl.step;
    ++n_steps;
    if (h_cfg->use_sdf[sl.gi]) {
      ++n_sdf;
      if (timing && sl.sdf_event >= 0) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, event(2 * sl.sdf_event), event(2 * sl.sdf_event + 1));
        sdf_ms += ms;
        (no_grad ? sdf_q_nj : sdf_q_j) += r.cnt[0];
      }
    }
    if (h_cfg->rgb_level[sl.gi] >= 0) ++n_rgb;
    memcpy(delta_out, r.delta, sizeof(delta_out));
    if (r.error) error = 1;
  };

  int rc = DFB_OK;
  for (int gi = 0; gi < h_cfg->n_groups && rc == DFB_OK && !error; ++gi) {
    const int n_it = h_cfg->n_iter[gi];
    DFB_CHECK_ARG(n_it >= 0 && n_it < 100000, "gauss_newton: n_iter out of range");
    if (h_cfg->rgb_level[gi] >= 0) DFB_CHECK_ARG(h_levels && h_intr && h_cfg->rgb_level[gi] < 3, "gauss_newton: rgb term needs pyramid levels and intrinsics");
    Slot cur, next;
    rc = enqueue(gi, 0, cur);
    bool have_next = false;
    for (int step = 0; rc == DFB_OK; ++step) {
      have_next = false;
      if (step + 1 <= n_it) {                                       // look-ahead: the next evaluation is queued before this one is read
        rc = enqueue(gi, step + 1, next);
        if (rc) break;
        have_next = true;
      }
      GnRecord r;
      rc = wait(cur, r);
      if (rc) break;
      account(cur, r);
      const bool group_over = r.broke || r.error || step == n_it;
      if (group_over) {
        if (have_next) {                                            // already queued: it returns at once on the device; drain its record
          rc = wait(next, r);
          if (rc == DFB_OK) account(next, r);
        }
        break;
      }
      cur = next;
    }
  }
  if (rc) { cudaStreamSynchronize(s); return rc; }
  if (error) { dfb::set_error("gauss_newton: singular normal equations"); h_stats[3] = 1; return DFB_E_INVALID; }
  memcpy(h_delta_pose, delta_out, sizeof(double) * 12);
  h_stats[0] = i_iter; h_stats[1] = n_sdf; h_stats[2] = n_rgb; h_stats[3] = 0;
  if (timing) {
    h_stats[4] = (int32_t)(sdf_ms * 1e3);        // microseconds spent in the SDF-term kernels
    h_stats[5] = (int32_t)sdf_q_j;               // valid queries evaluated with the reverse pass
    h_stats[6] = (int32_t)sdf_q_nj;              // valid queries evaluated forward-only
  }
  h_stats[7] = n_steps;                          // evaluations executed (= step kernels that did work)
  return DFB_OK;
}
