// Native Gauss-Newton driver: system/tracker.py:225-288 (gauss_newton) with the two fused terms, so one C call runs
// the whole pose solve of a frame: per evaluation one kernel (+ a 352-byte read back), the 6x6 solve, the SE(3)
// update and the accept / rollback logic in float64 on the host -- no Python, no torch ops, no extra syncs.
// Pose algebra restated from utils/motion_util.py:205-228 (from_twist), :275-279 (inv, dot).
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace {

struct Pose {   // x -> R x + t, float64
  double R[9], t[3];
};

void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
void mat3_vec(const double* A, const double* v, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
Pose compose(const Pose& a, const Pose& b) {   // a o b  (Isometry.dot)
  Pose c;
  mat3_mul(a.R, b.R, c.R);
  double rt[3];
  mat3_vec(a.R, b.t, rt);
  for (int i = 0; i < 3; ++i) c.t[i] = rt[i] + a.t[i];
  return c;
}

// The reference stores rotations as unit quaternions (pyquaternion normalises on every rotation_matrix access), which
// re-orthonormalises the pose each iteration; do the same round trip.
void renormalise(double* R) {
  double q[4];
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    double s = sqrt(tr + 1.0) * 2;
    q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
    q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
    q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
  } else {
    double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
    q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
  }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}

Pose from_twist(const double* xi) {   // motion_util.py:205-228
  Pose p;
  const double* rho = xi;
  const double* phi = xi + 3;
  const double angle = sqrt(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
  double Wd[9] = {0, -phi[2], phi[1], phi[2], 0, -phi[0], -phi[1], phi[0], 0};
  double J[9];
  if (fabs(angle) <= 1e-8) {   // np.isclose(angle, 0.)
    for (int i = 0; i < 9; ++i) { p.R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + Wd[i]; J[i] = ((i % 4 == 0) ? 1.0 : 0.0) + 0.5 * Wd[i]; }
  } else {
    const double ax[3] = {phi[0] / angle, phi[1] / angle, phi[2] / angle};
    const double s = sin(angle), c = cos(angle);
    const double Wa[9] = {0, -ax[2], ax[1], ax[2], 0, -ax[0], -ax[1], ax[0], 0};
    for (int i = 0; i < 3; ++i)
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

// np.linalg.solve(H, -g): LU with partial pivoting, float64.  Returns false when singular.
bool solve6(const double* H, const double* g, double* x) {
  double A[6][7];
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j) A[i][j] = H[6 * i + j];
    A[i][6] = -g[i];
  }
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    for (int r = c + 1; r < 6; ++r)
      if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
    if (!(fabs(A[piv][c]) > 0.0)) return false;
    if (piv != c)
      for (int j = 0; j < 7; ++j) { double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
    for (int r = c + 1; r < 6; ++r) {
      const double f = A[r][c] / A[c][c];
      for (int j = c; j < 7; ++j) A[r][j] -= f * A[c][j];
    }
  }
  for (int r = 5; r >= 0; --r) {
    double s = A[r][6];
    for (int j = r + 1; j < 6; ++j) s -= A[r][j] * x[j];
    x[r] = s / A[r][r];
  }
  for (int i = 0; i < 6; ++i)
    if (!isfinite(x[i])) return false;
  return true;
}

}  // namespace

extern "C" int dfb_gauss_newton(const dfb_map_params* h_params, const dfb_gn_config* h_cfg, const float* obs_xyz, int n,
                                const int64_t* indexer, const float* latent_vecs, const float* voxel_obs_count,
                                const float* decoder_blob, const dfb_rgb_level* h_levels, const double* h_intr,
                                const double* h_last_pose, double* h_delta_pose, double* d_scratch80, double* h_pinned44,
                                int32_t* h_stats, void* stream) {
  // optional kernel timing (CUDA events on the launching stream around every SDF term), reported through h_stats[4..7]:
  // enabled when h_stats[4] == 0x54494d45 ('TIME') on entry
  const bool timing = h_stats && h_stats[4] == 0x54494d45;
  static cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double sdf_ms = 0.0, sdf_q_j = 0.0, sdf_q_nj = 0.0;
  if (timing && !ev0) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); }
  DFB_CHECK_ARG(h_params && h_cfg && h_last_pose && h_delta_pose && d_scratch80 && h_pinned44 && h_stats, "gauss_newton");
  DFB_CHECK_ARG(h_cfg->n_groups >= 0 && h_cfg->n_groups <= 8, "gauss_newton: n_groups must be in [0, 8]");
  cudaStream_t s = (cudaStream_t)stream;
  Pose last, delta, last_delta;
  memcpy(last.R, h_last_pose, sizeof(double) * 9); memcpy(last.t, h_last_pose + 9, sizeof(double) * 3);
  memcpy(delta.R, h_delta_pose, sizeof(double) * 9); memcpy(delta.t, h_delta_pose + 9, sizeof(double) * 3);
  last_delta = delta;
  int n_sdf = 0, n_rgb = 0, i_iter = 0;
  const double fx = h_intr ? h_intr[0] : 1, fy = h_intr ? h_intr[1] : 1, cx = h_intr ? h_intr[2] : 0, cy = h_intr ? h_intr[3] : 0;
  const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  const double Kinv[9] = {1 / fx, 0, -cx / fx, 0, 1 / fy, -cy / fy, 0, 0, 1};

  for (int gi = 0; gi < h_cfg->n_groups; ++gi) {
    double last_energy = INFINITY;
    const int n_it = h_cfg->n_iter[gi];
    for (int step = 0; step <= n_it; ++step) {          // n_it iterations + one evaluation-only pass (i_iter = -1)
      i_iter = step < n_it ? step : -1;
      const bool no_grad = (i_iter == -1);
      double H[36], g[6], energy = 0.0;
      memset(H, 0, sizeof(H)); memset(g, 0, sizeof(g));
      for (int term = 0; term < 2; ++term) {
        int rc = DFB_OK;
        double scale_mul = 1.0;
        if (term == 0) {
          if (!h_cfg->use_sdf[gi]) continue;
          const Pose total = compose(last, delta);
          float hp[33];
          for (int i = 0; i < 9; ++i) { hp[i] = (float)total.R[i]; hp[12 + i] = (float)delta.R[i]; hp[24 + i] = (float)last.R[i]; }
          for (int i = 0; i < 3; ++i) { hp[9 + i] = (float)total.t[i]; hp[21 + i] = (float)delta.t[i]; }
          if (timing) cudaEventRecord(ev0, s);
          rc = dfb_sdf_hg(h_params, obs_xyz, n, hp, indexer, latent_vecs, voxel_obs_count, decoder_blob, h_cfg->sdf_robust,
                          h_cfg->sdf_robust_k, no_grad ? 0 : 1, d_scratch80, stream);
          if (timing) cudaEventRecord(ev1, s);
          ++n_sdf;
        } else {
          const int lvl = h_cfg->rgb_level[gi];
          if (lvl < 0) continue;
          DFB_CHECK_ARG(h_levels && h_intr && lvl < 3, "gauss_newton: rgb term needs pyramid levels and intrinsics");
          double KR[9], KRK[9], Kt[3];
          mat3_mul(K, delta.R, KR); mat3_mul(KR, Kinv, KRK); mat3_vec(K, delta.t, Kt);
          float intr[4] = {(float)fx, (float)fy, (float)cx, (float)cy}, krk[9], kt[3];
          for (int i = 0; i < 9; ++i) krk[i] = (float)KRK[i];
          for (int i = 0; i < 3; ++i) kt[i] = (float)Kt[i];
          const dfb_rgb_level& L = h_levels[lvl];
          rc = dfb_rgb_hg(L.prev_I, L.prev_D, L.cur_I, L.cur_D, L.cur_G, L.H, L.W, intr, krk, kt, h_cfg->rgb_min_grad_scale,
                          h_cfg->rgb_max_depth_delta, h_cfg->rgb_robust, h_cfg->rgb_robust_k, no_grad ? 0 : 1, d_scratch80, stream);
          scale_mul = h_cfg->rgb_weight;
          ++n_rgb;
        }
        if (rc) return rc;
        DFB_CUDA(cudaMemcpyAsync(h_pinned44, d_scratch80, sizeof(double) * 44, cudaMemcpyDeviceToHost, s));
        DFB_CUDA(cudaStreamSynchronize(s));
        const double cnt = h_pinned44[43];
        if (timing && term == 0) {
          float ms = 0.f;
          cudaEventElapsedTime(&ms, ev0, ev1);
          sdf_ms += ms;
          (no_grad ? sdf_q_nj : sdf_q_j) += cnt;
        }
        const double scale = scale_mul / cnt;                 // tracker.py:215 / :170 (inf/NaN when nothing is valid, like 1/0 there)
        energy += h_pinned44[42] * scale;
        if (!no_grad) {
          for (int i = 0; i < 36; ++i) H[i] += h_pinned44[i] * scale;
          for (int i = 0; i < 6; ++i) g[i] += h_pinned44[36 + i] * scale;
        }
      }
      if (energy > last_energy) {                             // tracker.py:269-271
        delta = last_delta;
        break;
      }
      last_delta = delta;
      last_energy = energy;
      if (!no_grad) {
        double xi[6];
        if (!solve6(H, g, xi)) { dfb::set_error("gauss_newton: singular normal equations"); h_stats[3] = 1; return DFB_E_INVALID; }
        delta = compose(from_twist(xi), delta);               // tracker.py:277-278
        renormalise(delta.R);
      }
    }
  }
  memcpy(h_delta_pose, delta.R, sizeof(double) * 9); memcpy(h_delta_pose + 9, delta.t, sizeof(double) * 3);
  h_stats[0] = i_iter; h_stats[1] = n_sdf; h_stats[2] = n_rgb; h_stats[3] = 0;
  if (timing) {
    h_stats[4] = (int32_t)(sdf_ms * 1e3);        // microseconds spent in the SDF-term launches (memset + kernel + expand)
    h_stats[5] = (int32_t)sdf_q_j;               // valid queries evaluated with the reverse pass
    h_stats[6] = (int32_t)sdf_q_nj;              // valid queries evaluated forward-only
  }
  return DFB_OK;
}
