// Native Gauss-Newton driver: system/tracker.py:225-288 (gauss_newton) with the two fused terms, device-resident.
// One C call runs the whole pose solve of a frame.  An evaluation ("slot") is ONE launch on the caller's stream with the
// tcgen05 engine (decoder_tc.cu gn_eval_kernel: SDF tiles + photometric pixels + the step as the tail of the last
// block), one launch for photometric-only groups (photometric.cu rgb_step_gn_kernel), and
//   [SDF term kernel] [photometric term kernel] [gn_step_kernel <<<1, 32>>>]
// with the FP32 engine.  The term code reads its pose from GnShared and adds its packed sums there; the step
// (gn_step.cuh) scales and sums the terms, applies the accept / rollback rule, solves the 6x6 system (float64 LU,
// partial pivoting), composes the SE(3) update, publishes the pose blocks of the next evaluation and writes a small
// record into pinned host memory.  The host keeps ONE slot of look-ahead: slot k+1 is enqueued before the record of
// slot k is read, so the GPU never waits for the host between evaluations; when slot k ends its group (energy rose ->
// rollback, tracker.py:269-271) the launch of slot k+1 sees done[group] and returns immediately.
// Pose algebra restated from utils/motion_util.py:205-228 (from_twist), :275-279 (inv, dot).
#include <math.h>
#include <string.h>
#include <chrono>
#include <vector>

#include "common.cuh"
#include "gn_step.cuh"

namespace {
using namespace dfb::gn;

struct GnInit {
  double last[12], delta[12], intr[4];
};

__global__ void gn_init_kernel(dfb::GnShared* gs, GnInit in) {
  const int t = threadIdx.x;
  if (t < 64) reinterpret_cast<double*>(gs->sums)[t] = 0.0;
  if (t < 12) { gs->last[t] = in.last[t]; gs->delta[t] = in.delta[t]; gs->last_delta[t] = in.delta[t]; }
  if (t < 4) gs->intr[t] = in.intr[t];
  if (t == 0) { gs->kinv[0] = 1 / in.intr[0]; gs->kinv[1] = 1 / in.intr[1]; gs->kinv[2] = -in.intr[2] / in.intr[0]; gs->kinv[3] = -in.intr[3] / in.intr[1]; }
  if (t < 8) gs->done[t] = 0;
  if (t == 0) { gs->error = 0; gs->last_energy = CUDART_INF; gs->ticket = 0; gs->rgb_cursor = 0; gs->pad_ = 0; }
  __syncthreads();
  if (t == 0) publish_pose(gs);
}

// Stand-alone step kernel (used after the FP32-engine SDF term, whose kernel has no fused tail).
__global__ void __launch_bounds__(32) gn_step_kernel(dfb::GnShared* gs, StepArgs a) {
  __shared__ __align__(16) unsigned char scratch[STEP_SCRATCH_BYTES];
  step_warp(gs, a, scratch);
}

}  // namespace

extern "C" int dfb_gauss_newton(const dfb_map_params* h_params, const dfb_gn_config* h_cfg, const float* obs_xyz, int n,
                                const int32_t* d_n, const int64_t* indexer, const float* latent_vecs, const float* voxel_obs_count,
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
  for (int i = 0; i < 4; ++i) { hring[i].seq = 0; hring[i].check = 0; }

  static thread_local std::vector<cudaEvent_t> events;                           // pairs, grown on demand (timing only)
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

  const bool tc_engine = dfb_get_decoder_engine() == 1;
  static thread_local int seq_base = 0;                             // (per host thread: two trackers may solve concurrently)                                          // records carry a process-unique, non-zero sequence number
  struct Slot { int seq, gi, step, sdf_event; };
  int n_sdf = 0, n_rgb = 0, n_launches = 1 /* gn_init_kernel */, i_iter = 0, n_events = 0, error = 0;
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
    const int use_sdf = h_cfg->use_sdf[gi] ? 1 : 0, use_rgb = lvl >= 0 ? 1 : 0;
    const StepArgs sa = {ring, out.seq, gi, step, n_it, use_sdf, use_rgb, (double)h_cfg->rgb_weight};
    if (use_sdf && tc_engine) {
      // ONE launch per evaluation: SDF tiles on the tensor cores, photometric pixels work-stolen by idle tile groups,
      // and the step as the tail of the last block (decoder_tc.cu)
      if (timing) { out.sdf_event = n_events; cudaEventRecord(event(2 * n_events), s); }
      int rc = dfb::launch_sdf_rgb_gn(h_params, obs_xyz, n, d_n, indexer, latent_vecs, voxel_obs_count, decoder_blob, h_cfg->sdf_robust,
                                      h_cfg->sdf_robust_k, no_grad ? 0 : 1, use_rgb ? &h_levels[lvl] : nullptr, intr4,
                                      h_cfg->rgb_min_grad_scale, h_cfg->rgb_max_depth_delta, h_cfg->rgb_robust, h_cfg->rgb_robust_k, gs, gi,
                                      &sa, s);
      if (rc) return rc;
      ++n_launches;
      if (timing) { cudaEventRecord(event(2 * n_events + 1), s); ++n_events; }
      return DFB_OK;
    }
    if (!use_sdf && use_rgb) {                   // photometric-only group: one launch, step as the tail of the last block
      ++n_launches;
      return dfb::launch_rgb_step_gn(&h_levels[lvl], intr4, h_cfg->rgb_min_grad_scale, h_cfg->rgb_max_depth_delta, h_cfg->rgb_robust,
                                     h_cfg->rgb_robust_k, no_grad ? 0 : 1, gs, gi, &sa, s);
    }
    n_launches += (use_sdf && n > 0 ? 1 : 0) + use_rgb + 1;
    if (use_sdf) {                               // FP32 CUDA-core engine: term kernels, then the stand-alone step kernel
      if (timing) { out.sdf_event = n_events; cudaEventRecord(event(2 * n_events), s); }
      int rc = dfb::launch_sdf_hg_gn(h_params, obs_xyz, n, d_n, indexer, latent_vecs, voxel_obs_count, decoder_blob, h_cfg->sdf_robust,
                                     h_cfg->sdf_robust_k, no_grad ? 0 : 1, gs, gi, s);
      if (rc) return rc;
      if (timing) { cudaEventRecord(event(2 * n_events + 1), s); ++n_events; }
    }
    if (use_rgb) {
      int rc = dfb::launch_rgb_hg_gn(&h_levels[lvl], intr4, h_cfg->rgb_min_grad_scale, h_cfg->rgb_max_depth_delta, h_cfg->rgb_robust,
                                     h_cfg->rgb_robust_k, no_grad ? 0 : 1, gs, gi, s);
      if (rc) return rc;
    }
    gn_step_kernel<<<1, 32, 0, s>>>(gs, sa);
    DFB_LAUNCH_CHECK();
    return DFB_OK;
  };
  // wait for the record of a slot (spin on pinned memory; the stream is polled for errors now and then)
  struct Rec { bool executed, broke, error, has_delta; double cnt0; double delta[12]; };
  auto wait = [&](const Slot& sl, Rec& r) -> int {
    volatile GnRecord* rec = hring + (sl.seq & 3);
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spin = 0;; ++spin) {
      if (rec->seq == sl.seq && rec->check == (sl.seq ^ dfb::GN_CHECK)) { __atomic_thread_fence(__ATOMIC_ACQUIRE); break; }
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();                     // be polite to the sibling hyper-thread while spinning
#endif
      if ((spin & 0xfff) == 0xfff) {
        cudaError_t e = cudaStreamQuery(s);
        if (e != cudaSuccess && e != cudaErrorNotReady) { dfb::set_error("gauss_newton: %s", cudaGetErrorString(e)); return DFB_E_CUDA; }
        if (e == cudaSuccess && rec->seq != sl.seq) {               // stream drained without the record: cannot happen unless a launch failed
          if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2)) { dfb::set_error("gauss_newton: step record never arrived"); return DFB_E_CUDA; }
        }
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) { dfb::set_error("gauss_newton: timed out waiting for the device"); return DFB_E_CUDA; }
      }
    }
    const int flags = rec->flags;
    r.executed = flags & dfb::GN_EXECUTED; r.broke = flags & dfb::GN_BROKE; r.error = flags & dfb::GN_ERROR;
    r.has_delta = flags & dfb::GN_HAS_DELTA;
    r.cnt0 = (double)rec->cnt0;
    if (r.has_delta)
      for (int i = 0; i < 12; ++i) r.delta[i] = rec->delta[i];
    return DFB_OK;
  };
  auto account = [&](const Slot& sl, const Rec& r) {
    if (r.error) error = 1;
    if (!r.executed) return;
    const int n_it = h_cfg->n_iter[sl.gi];
    const bool no_grad = (sl.step == n_it);
    i_iter = no_grad ? -1 : sl.step;
    if (h_cfg->use_sdf[sl.gi]) {
      ++n_sdf;
      if (timing && sl.sdf_event >= 0) {
        float ms = 0.f;
        cudaEventSynchronize(event(2 * sl.sdf_event + 1));   // the record is written by the kernel's last block, just before it exits
        cudaEventElapsedTime(&ms, event(2 * sl.sdf_event), event(2 * sl.sdf_event + 1));
        sdf_ms += ms;
        (no_grad ? sdf_q_nj : sdf_q_j) += r.cnt0;
      }
    }
    if (h_cfg->rgb_level[sl.gi] >= 0) ++n_rgb;
    if (r.has_delta) memcpy(delta_out, r.delta, sizeof(delta_out));   // the step that ends a group carries the pose
  };

  int rc = DFB_OK;
  // A look-ahead launch that was queued behind the evaluation that ended its group returns at once on the device; its record
  // is drained only AFTER the next group's first evaluation has been enqueued, so the GPU goes from one group to the next
  // without waiting for a host round trip (ring: at most the skipped slot + two slots of the next group are outstanding).
  Slot skipped; bool have_skipped = false;
  auto drain_skipped = [&]() -> int {
    if (!have_skipped) return DFB_OK;
    have_skipped = false;
    Rec r;
    int rc2 = wait(skipped, r);
    if (rc2 == DFB_OK) account(skipped, r);
    return rc2;
  };
  for (int gi = 0; gi < h_cfg->n_groups && rc == DFB_OK && !error; ++gi) {
    const int n_it = h_cfg->n_iter[gi];
    DFB_CHECK_ARG(n_it >= 0 && n_it < 100000, "gauss_newton: n_iter out of range");
    if (h_cfg->rgb_level[gi] >= 0) DFB_CHECK_ARG(h_levels && h_intr && h_cfg->rgb_level[gi] < 3, "gauss_newton: rgb term needs pyramid levels and intrinsics");
    Slot cur, next;
    rc = enqueue(gi, 0, cur);
    if (rc == DFB_OK) rc = drain_skipped();
    bool have_next = false;
    for (int step = 0; rc == DFB_OK; ++step) {
      have_next = false;
      if (step + 1 <= n_it) {                                       // look-ahead: the next evaluation is queued before this one is read
        rc = enqueue(gi, step + 1, next);
        if (rc) break;
        have_next = true;
      }
      Rec r;
      rc = wait(cur, r);
      if (rc) break;
      account(cur, r);
      const bool group_over = r.broke || r.error || step == n_it;
      if (group_over) {
        if (have_next) { skipped = next; have_skipped = true; }     // already queued: it returns at once on the device
        break;
      }
      cur = next;
    }
  }
  if (rc == DFB_OK) rc = drain_skipped();
  if (rc) { cudaStreamSynchronize(s); return rc; }
  if (error) { dfb::set_error("gauss_newton: singular normal equations"); h_stats[3] = 1; return DFB_E_INVALID; }
  memcpy(h_delta_pose, delta_out, sizeof(double) * 12);
  h_stats[0] = i_iter; h_stats[1] = n_sdf; h_stats[2] = n_rgb; h_stats[3] = 0;
  if (timing) {
    h_stats[4] = (int32_t)(sdf_ms * 1e3);        // microseconds spent in the SDF-term kernels
    h_stats[5] = (int32_t)sdf_q_j;               // valid queries evaluated with the reverse pass
    h_stats[6] = (int32_t)sdf_q_nj;              // valid queries evaluated forward-only
  }
  h_stats[7] = n_launches;                       // kernels launched by this call (including look-ahead launches that returned at once)
  return DFB_OK;
}
