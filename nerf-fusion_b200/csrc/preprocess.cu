// Stage 1: depth unprojection, uniform-grid exact neighbour search (radius-outlier filter, PCA normals),
// 2 cm box filter / scatter_mean, groupby_sum.
//
// The reference builds a FLANN kd-tree TWICE per frame with thrust sorts and a host sync per tree level
// (cuda_kdtree.cu:723-856) and keeps per-query heaps in global memory (cuda_kdtree.cu:130-214).  Both of its
// queries are radius-bounded (pcproc.cu:104,120), so an exact answer only needs the points of the 27 cells
// around the query in a uniform grid with cell >= radius: one counting sort, no host sync, candidates read
// coalesced from the cell-sorted copy, result set in registers.
#include <cstdlib>
#include "common.cuh"

namespace dfb {

// ------------------------------------------------------------------------------------------------
// unproject (imgproc.cu:5-23).  x -> column (coalesced); the reference maps threadIdx.x to rows.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unproject_kernel(const float* __restrict__ depth, int H, int W, float fx,
                                                        float fy, float cx, float cy, float* __restrict__ pc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  int v = i / W, u = i - v * W;
  float d = depth[i];
  float x, y, z;
  if (!isnan(d)) {
    x = ((float)u - cx) / fx * d;   // same expression shape as imgproc.cu:17-19 (IEEE divide)
    y = ((float)v - cy) / fy * d;
    z = d;
  } else {
    x = y = z = CUDART_NAN_F;
  }
  pc[3 * i + 0] = x;
  pc[3 * i + 1] = y;
  pc[3 * i + 2] = z;
}

// Isometry @ points (utils/motion_util.py:323-328: other @ R^T + t in fp32).  A 3x3 transform does not need cuBLAS -- and
// must not use it in the frame loop: a point count cuBLAS has not seen before can select a kernel that is not loaded
// yet, and the lazy module load was measured at 20-36 ms in the middle of a 1.1 ms frame.
struct Rt9 { float r[9], t[3]; };
__global__ void __launch_bounds__(256) transform_points_kernel(const float* __restrict__ xyz, int n, Rt9 R, bool add_t, float* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float acc = x * R.r[3 * a];
    acc = fmaf(y, R.r[3 * a + 1], acc);
    acc = fmaf(z, R.r[3 * a + 2], acc);
    out[3 * (size_t)i + a] = add_t ? __fadd_rn(acc, R.t[a]) : acc;
  }
}

// Frame ingest (dataset/production/icl_nuim.py:110-114 + main.py:56-57): the raw 16-bit depth PNG and 8-bit colour image are
// uploaded as they are (5 bytes/pixel instead of 16) and converted here: depth = raw / scale, colour = raw / 255, optional
// BGR -> RGB swap (cv2.cvtColor in the reference), optional clipping to NaN outside [cut_min, cut_max].
__global__ void __launch_bounds__(256) ingest_kernel(const uint16_t* __restrict__ depth_raw, const uint8_t* __restrict__ color_raw, int n,
                                                     float scale, float inv_scale, int div_mode, float cut_min, float cut_max, int bgr,
                                                     float* __restrict__ depth_out, float* __restrict__ rgb_out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  if (depth_raw) {
    float d = div_vs((float)depth_raw[i], scale, inv_scale, div_mode);
    if (cut_min < cut_max && (d < cut_min || d > cut_max)) d = CUDART_NAN_F;
    depth_out[i] = d;
  }
  if (color_raw) {
    const float c0 = (float)color_raw[3 * i], c1 = (float)color_raw[3 * i + 1], c2 = (float)color_raw[3 * i + 2];
    const float inv255 = 1.0f / 255.0f;
    rgb_out[3 * i + 0] = div_vs(bgr ? c2 : c0, 255.0f, inv255, div_mode);
    rgb_out[3 * i + 1] = div_vs(c1, 255.0f, inv255, div_mode);
    rgb_out[3 * i + 2] = div_vs(bgr ? c0 : c2, 255.0f, inv255, div_mode);
  }
}

// ------------------------------------------------------------------------------------------------
// uniform grid
// ------------------------------------------------------------------------------------------------
constexpr int GRID_CAP = 1 << 22;        // max cells (workspace sizing)
constexpr int GRID_CAP_COARSE = 1 << 20; // cell budget of the radius-count grid (cell = radius)
constexpr int NORMAL_SUB = 4;            // the kNN grid of estimate_normals uses cells of radius / NORMAL_SUB

struct GridParams {
  float ox, oy, oz, cell, inv_cell;
  int nx, ny, nz, ncell;
  int ncell1;          // ncell + 1: device-side length of the count / start arrays
};

struct GridWs {
  unsigned* bbox;      // 6: min xyz, max xyz (ordered-uint encoding)
  GridParams* gp;
  int* cell_of;        // n
  int* cell_count;     // GRID_CAP (+1)
  int* cell_start;     // GRID_CAP + 1
  int* block_sums;
  float4* sorted;      // n   (x, y, z, original index bits)
};

static size_t grid_ws_layout(Arena& a, int n, GridWs* w) {
  w->bbox = a.take<unsigned>(8);
  w->gp = a.take<GridParams>(1);
  w->cell_of = a.take<int>(n > 0 ? n : 1);
  w->cell_count = a.take<int>(GRID_CAP + 1);
  w->cell_start = a.take<int>(GRID_CAP + 1);
  w->block_sums = a.take<int>(GRID_CAP / 2048 + 8);
  w->sorted = a.take<float4>(n > 0 ? n : 1);
  return a.off;
}

// bbox[6] counts the occupied cells of the grid being built (grid_count_kernel); bbox[7] keeps the count of the previous
// build in the same workspace (the density hint of grid_params_kernel)
__global__ void bbox_init_kernel(unsigned* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) bbox[threadIdx.x] = 0u;
  else if (threadIdx.x == 6) { bbox[7] = bbox[6]; bbox[6] = 0u; }
}

__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ p, int n, int stride, unsigned* bbox, const int* __restrict__ n_dev) {
  if (n_dev) n = *n_dev;
  float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float v = p[(size_t)i * stride + a];
      if (isfinite(v)) { mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  }
  // warp results meet in shared memory: six global atomics per block instead of six per warp (the atomics all hit the
  // same six words, so their number is the kernel's latency)
  __shared__ float s_mn[8][3], s_mx[8][3];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { s_mn[w][a] = mn[a]; s_mx[w][a] = mx[a]; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int a = threadIdx.x;
    float lo = s_mn[0][a], hi = s_mx[0][a];
#pragma unroll
    for (int ww = 1; ww < 8; ++ww) { lo = fminf(lo, s_mn[ww][a]); hi = fmaxf(hi, s_mx[ww][a]); }
    if (lo <= hi) {
      atomicMin(&bbox[a], f2ord(lo));
      atomicMax(&bbox[3 + a], f2ord(hi));
    }
  }
}

// hint_n (optional): the PREVIOUS grid built in this workspace held *hint_n points in bbox[7] occupied cells of edge gp->cell.
// For a surface-like cloud that gives the point spacing s = cell / sqrt(points per occupied cell), and the kNN search wants
// cells of about the 16-neighbour radius (2.26 s): with much smaller cells a query walks dozens of nearly empty cells (at
// 3.6 m a 640x480 depth frame has 15 mm spacing: 85 % of the queries needed two shells of 25 mm cells, 5 % three or four),
// with much larger ones it scans hundreds of candidates.  The result of the search does not depend on the cell size.
__global__ void grid_params_kernel(const unsigned* bbox, float radius, int cap, GridParams* gp, const int* hint_n, float cell_max) {
  if (hint_n) {
    const float occ = (float)bbox[7], n = (float)*hint_n;
    if (occ > 0.f && n > 0.f) {
      const float spacing = gp->cell * rsqrtf(n / occ);
      radius = fminf(fmaxf(3.0f * spacing, radius), cell_max);
    }
  }
  float lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = ord2f(bbox[a]);
    hi[a] = ord2f(bbox[3 + a]);
    if (!(lo[a] <= hi[a])) { lo[a] = 0.f; hi[a] = 0.f; }
  }
  float cell = radius * 1.001f;
  int nx, ny, nz;
  for (int it = 0; it < 64; ++it) {
    nx = (int)floorf((hi[0] - lo[0]) / cell) + 1;
    ny = (int)floorf((hi[1] - lo[1]) / cell) + 1;
    nz = (int)floorf((hi[2] - lo[2]) / cell) + 1;
    if ((double)nx * ny * nz <= (double)cap) break;
    cell *= 1.25f;
  }
  gp->ox = lo[0]; gp->oy = lo[1]; gp->oz = lo[2];
  gp->cell = cell; gp->inv_cell = 1.0f / cell;
  gp->nx = nx; gp->ny = ny; gp->nz = nz; gp->ncell = nx * ny * nz; gp->ncell1 = gp->ncell + 1;
}

__device__ __forceinline__ int3 cell_coord(const GridParams& g, float x, float y, float z) {
  int cx = (int)floorf((x - g.ox) * g.inv_cell), cy = (int)floorf((y - g.oy) * g.inv_cell),
      cz = (int)floorf((z - g.oz) * g.inv_cell);
  cx = min(max(cx, 0), g.nx - 1); cy = min(max(cy, 0), g.ny - 1); cz = min(max(cz, 0), g.nz - 1);
  return make_int3(cx, cy, cz);
}

__global__ void __launch_bounds__(256) grid_count_kernel(const float* __restrict__ pc4, int n, const GridParams* gpp,
                                                         int* cell_of, int* cell_count, const int* __restrict__ n_dev, unsigned* bbox) {
  if (n_dev) n = *n_dev;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  GridParams g = *gpp;
  float4 p = reinterpret_cast<const float4*>(pc4)[i];
  if (isnan(p.x)) { cell_of[i] = -1; return; }      // an invalid row of an uncompacted cloud (fused preprocessing): in no cell
  int3 c = cell_coord(g, p.x, p.y, p.z);
  int id = (c.x * g.ny + c.y) * g.nz + c.z;
  cell_of[i] = id;
  if (atomicAdd(&cell_count[id], 1) == 0) atomicAdd(&bbox[6], 1u);
}

__global__ void __launch_bounds__(256) grid_scatter_kernel(const float* __restrict__ pc4, int n, const int* cell_of,
                                                           const int* cell_start, int* cursor, float4* sorted, const int* __restrict__ n_dev) {
  if (n_dev) n = *n_dev;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = reinterpret_cast<const float4*>(pc4)[i];
  int c = cell_of[i];
  if (c < 0) return;
  // `cursor` IS the cell-count array: counting it back down to zero hands out the slots of the cell (in reverse) and leaves the
  // array cleared for the next grid -- no zeroing pass between the count and the scatter
  int pos = cell_start[c] + atomicSub(&cursor[c], 1) - 1;
  sorted[pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
}

// squared distance, same expression shape as cutil_math.h:1127-1130 on (a - b) with w = 0
__device__ __forceinline__ float dist2(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = ax - bx, dy = ay - by, dz = az - bz;
  return dx * dx + dy * dy + dz * dz;
}

__global__ void __launch_bounds__(128) radius_count_kernel(const float4* __restrict__ sorted, int n,
                                                           const GridParams* gpp, const int* __restrict__ cell_start,
                                                           int nb_points, float radius, uint8_t* __restrict__ mask,
                                                           const int* __restrict__ n_dev, int* __restrict__ flag = nullptr,
                                                           uint8_t* __restrict__ dead = nullptr) {
  if (n_dev) n = *n_dev;
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  GridParams g = *gpp;
  float4 q = sorted[j];
  int3 c = cell_coord(g, q.x, q.y, q.z);
  const float r2 = radius * radius;
  int count = 0;
  for (int dx = -1; dx <= 1 && count < nb_points; ++dx) {
    int x = c.x + dx;
    if (x < 0 || x >= g.nx) continue;
    for (int dy = -1; dy <= 1 && count < nb_points; ++dy) {
      int y = c.y + dy;
      if (y < 0 || y >= g.ny) continue;
      int z0 = max(c.z - 1, 0), z1 = min(c.z + 1, g.nz - 1);
      int row = (x * g.ny + y) * g.nz;
      int b = cell_start[row + z0], e = cell_start[row + z1 + 1];   // z-neighbours are contiguous in memory
      for (int k = b; k < e; ++k) {
        float4 p = sorted[k];
        // element - query, like cuda_kdtree.cu:1019 distance.dist(elements[i], q)
        count += (dist2(p.x, p.y, p.z, q.x, q.y, q.z) < r2) ? 1 : 0;
      }
    }
  }
  mask[__float_as_int(q.w)] = count >= nb_points ? 1 : 0;
  if (flag) flag[__float_as_int(q.w)] = count >= nb_points ? 1 : 0;     // fused path: the compaction's scan input, no separate pass
  if (dead) dead[j] = count >= nb_points ? 0 : 1;                       // by position in `sorted`: lets the normals search reuse THIS grid
}

// pcproc.cu:21-96, same expression shapes (including the double-precision promotions through M_PI).
__device__ float4 sym3eig_smallest(float3 x1, float3 x2, float3 x3) {
  float4 ret;
  const float p1 = x1.y * x1.y + x1.z * x1.z + x2.z * x2.z;
  const float q = (x1.x + x2.y + x3.z) / 3.0f;
  const float p2 = (x1.x - q) * (x1.x - q) + (x2.y - q) * (x2.y - q) + (x3.z - q) * (x3.z - q) + 2 * p1;
  const float p = sqrtf(p2 / 6.0f);
  const float ip = 1.0f / p;
  const float b11 = ip * (x1.x - q), b12 = ip * x1.y, b13 = ip * x1.z;
  const float b21 = ip * x2.x, b22 = ip * (x2.y - q), b23 = ip * x2.z;
  const float b31 = ip * x3.x, b32 = ip * x3.y, b33 = ip * (x3.z - q);
  float r = b11 * b22 * b33 + b12 * b23 * b31 + b13 * b21 * b32 - b13 * b22 * b31 - b12 * b21 * b33 - b11 * b23 * b32;
  r = r / 2.0f;
  float phi;
  if (r <= -1) phi = (float)(3.14159265358979323846 / 3.0f);
  else if (r >= 1) phi = 0;
  else phi = acosf(r) / 3.0f;
  ret.w = (float)(q + 2 * p * cos(phi + (2 * 3.14159265358979323846 / 3)));
  x1.x -= ret.w; x2.y -= ret.w; x3.z -= ret.w;
  const float r12_1 = x1.y * x2.z - x1.z * x2.y, r12_2 = x1.z * x2.x - x1.x * x2.z, r12_3 = x1.x * x2.y - x1.y * x2.x;
  const float r13_1 = x1.y * x3.z - x1.z * x3.y, r13_2 = x1.z * x3.x - x1.x * x3.z, r13_3 = x1.x * x3.y - x1.y * x3.x;
  const float r23_1 = x2.y * x3.z - x2.z * x3.y, r23_2 = x2.z * x3.x - x2.x * x3.z, r23_3 = x2.x * x3.y - x2.y * x3.x;
  const float d1 = r12_1 * r12_1 + r12_2 * r12_2 + r12_3 * r12_3;
  const float d2 = r13_1 * r13_1 + r13_2 * r13_2 + r13_3 * r13_3;
  const float d3 = r23_1 * r23_1 + r23_2 * r23_2 + r23_3 * r23_3;
  float d_max = d1;
  int i_max = 0;
  if (d2 > d_max) { d_max = d2; i_max = 1; }
  if (d3 > d_max) { i_max = 2; }
  if (i_max == 0) { float s = sqrtf(d1); ret.x = r12_1 / s; ret.y = r12_2 / s; ret.z = r12_3 / s; }
  else if (i_max == 1) { float s = sqrtf(d2); ret.x = r13_1 / s; ret.y = r13_2 / s; ret.z = r13_3 / s; }
  else { float s = sqrtf(d3); ret.x = r23_1 / s; ret.y = r23_2 / s; ret.z = r23_3 / s; }
  return ret;
}

// pcproc.cu:107-158 on the sorted key list (slot 0 = the query itself): mean, covariance, smallest eigenvector, orientation
template <int K>
__device__ __forceinline__ void normal_from_keys(const unsigned long long (&key)[K], const float4 q, const float4* __restrict__ pc4, int max_nn,
                                                 float r2, float3 cam, float* __restrict__ normals, int* __restrict__ flag = nullptr) {
  int oi = __float_as_int(q.w);
  float3 mean = make_float3(0.f, 0.f, 0.f);
  float valid = 0.f;
#pragma unroll
  for (int t = 1; t < K; ++t) {
    if (t < max_nn && __uint_as_float((unsigned)(key[t] >> 32)) < r2) {
      float4 p = pc4[(unsigned)key[t]];
      mean.x += p.x; mean.y += p.y; mean.z += p.z;
      valid += 1.0f;
    }
  }
  if (valid < 5.0f) {
    normals[3 * oi + 0] = normals[3 * oi + 1] = normals[3 * oi + 2] = CUDART_NAN_F;
    if (flag) flag[oi] = 0;
    return;
  }
  mean.x /= valid; mean.y /= valid; mean.z /= valid;
  float3 c1 = make_float3(0.f, 0.f, 0.f), c2 = c1, c3 = c1;
#pragma unroll
  for (int t = 1; t < K; ++t) {
    if (t < max_nn && __uint_as_float((unsigned)(key[t] >> 32)) < r2) {
      float4 pp = pc4[(unsigned)key[t]];
      float3 pos = make_float3(pp.x - mean.x, pp.y - mean.y, pp.z - mean.z);
      c1.x += pos.x * pos.x; c1.y += pos.x * pos.y; c1.z += pos.x * pos.z;
      c2.x += pos.y * pos.x; c2.y += pos.y * pos.y; c2.z += pos.y * pos.z;
      c3.x += pos.z * pos.x; c3.y += pos.z * pos.y; c3.z += pos.z * pos.z;
    }
  }
  float4 ev = sym3eig_smallest(c1, c2, c3);
  float3 nrm = make_float3(ev.x, ev.y, ev.z);
  float3 dp = make_float3(q.x - cam.x, q.y - cam.y, q.z - cam.z);
  if (nrm.x * dp.x + nrm.y * dp.y + nrm.z * dp.z > 0.0f) { nrm.x = -nrm.x; nrm.y = -nrm.y; nrm.z = -nrm.z; }
  normals[3 * oi + 0] = nrm.x;
  normals[3 * oi + 1] = nrm.y;
  normals[3 * oi + 2] = nrm.z;
  if (flag) flag[oi] = isnan(nrm.x) ? 0 : 1;       // what flag_from_normals_kernel computed in a pass of its own
}

// Position of a query inside its grid cell (in cell units) and the distance, in cells, from the query to the slab of cells
// at offset d along one axis (0 for the query's own slab).  cell2 carries a 0.2 % safety margin for the rounding of
// cell_coord, like the `reach` test of the shell loop: the bound may only err on the small side.
struct CellFrac { float x, y, z, cell2; };
__device__ __forceinline__ CellFrac cell_frac(const GridParams& g, const float4 q, const int3 c) {
  CellFrac f;
  f.x = (q.x - g.ox) * g.inv_cell - (float)c.x; f.y = (q.y - g.oy) * g.inv_cell - (float)c.y; f.z = (q.z - g.oz) * g.inv_cell - (float)c.z;
  f.cell2 = g.cell * g.cell * 0.998f;
  return f;
}
__device__ __forceinline__ float cell_gap(int d, float frac) {
  return d > 0 ? fmaxf((float)d - frac, 0.f) : (d < 0 ? fmaxf(frac - (float)(d + 1), 0.f) : 0.f);
}

// K nearest (self included) kept sorted ascending in registers; then pcproc.cu:107-158.
// An entry is one 64-bit key: (distance bits << 32) | ORIGINAL index.  Distances are >= +0, so their bit patterns order
// like the values, and one unsigned compare implements "closer, or equally close and earlier in the input" -- the
// tie-break that makes the result independent of the atomics-defined order inside a grid cell and of the visiting
// order (equal distances are common on a pixel lattice).  Coordinates are re-read from the input array by that index.
template <int K>
__global__ void __launch_bounds__(128, 4) normals_kernel(const float4* __restrict__ sorted, const float4* __restrict__ pc4, int n,
                                                         const GridParams* gpp, const int* __restrict__ cell_start, int max_nn, float radius,
                                                         float3 cam, float* __restrict__ normals, const int* __restrict__ n_dev, int* __restrict__ flag = nullptr,
                                                         const uint8_t* __restrict__ dead = nullptr) {
  if (n_dev) n = *n_dev;
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  GridParams g = *gpp;
  float4 q = sorted[j];
  if (dead && dead[j]) {                          // a point the radius filter removed: no query, and never a candidate (below)
    if (flag) flag[__float_as_int(q.w)] = 0;
    return;
  }
  int3 c = cell_coord(g, q.x, q.y, q.z);
  unsigned long long key[K];
#pragma unroll
  for (int t = 0; t < K; ++t) key[t] = 0x7f8000007fffffffull;       // (+inf, INT_MAX)
  const float r2 = radius * radius;
  // Expanding shells of cells around the query's cell (the grid is finer than the radius).  After the block [c-R, c+R]^3
  // has been scanned every unvisited point is farther than R * cell from the query, so once the last entry that will
  // be used is closer than that (with a margin for the rounding of cell_coord) the list is final: the exact K nearest.
  const int Rmax = max(1, (int)ceilf(radius * g.inv_cell));
  const int last = min(max_nn, K) - 1;
  const CellFrac cf = cell_frac(g, q, c);
  for (int R = 1; R <= Rmax; ++R) {
    // columns nearest to the query first (0, -1, +1, -2, ...): the K-th distance tightens early and most later
    // candidates fail the cheap test
    for (int ix = 0; ix <= 2 * R; ++ix) {
      const int dx = (ix & 1) ? -((ix + 1) >> 1) : (ix >> 1);
      const int x = c.x + dx;
      if (x < 0 || x >= g.nx) continue;
      const float gx = cell_gap(dx, cf.x);
      for (int iy = 0; iy <= 2 * R; ++iy) {
        const int dy = (iy & 1) ? -((iy + 1) >> 1) : (iy >> 1);
        const int y = c.y + dy;
        if (y < 0 || y >= g.ny) continue;
        // no point of this column can be closer than its cells are: skip it when that already exceeds the K-th distance
        const float gy = cell_gap(dy, cf.y);
        if ((gx * gx + gy * gy) * cf.cell2 >= __uint_as_float((unsigned)(key[K - 1] >> 32))) continue;
        const int row = (x * g.ny + y) * g.nz;
        // R == 1 or a column on the shell's side faces: the whole z-run (contiguous in memory); interior column: only the
        // two new end cells
        const bool whole = (R == 1 || dx == -R || dx == R || dy == -R || dy == R);
        for (int sgm = 0; sgm < 2; ++sgm) {
          int za, zb;
          if (whole) {
            if (sgm) break;
            za = max(c.z - R, 0); zb = min(c.z + R, g.nz - 1);
          } else {
            za = zb = sgm ? c.z + R : c.z - R;
            if (za < 0 || za >= g.nz) continue;
          }
          const int b = cell_start[row + za], e = cell_start[row + zb + 1];
          for (int k = b; k < e; ++k) {
            if (dead && dead[k]) continue;
            const float4 p = sorted[k];
            const float d = dist2(p.x, p.y, p.z, q.x, q.y, q.z);
            // entries beyond the radius can never be used (pcproc.cu:120 breaks at the first miss), except that the
            // self entry (d = 0) must occupy slot 0; d < r2 keeps self.
            const unsigned long long nk = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
            if (d < r2 && nk < key[K - 1]) {
              key[K - 1] = nk;
#pragma unroll
              for (int t = K - 1; t > 0; --t) {
                const unsigned long long a = key[t - 1], bb = key[t];
                const bool sw = bb < a;
                key[t - 1] = sw ? bb : a;
                key[t] = sw ? a : bb;
              }
            }
          }
        }
      }
    }
    const float reach = (float)R * g.cell * 0.9999f;
    if (__uint_as_float((unsigned)(key[last] >> 32)) < reach * reach) break;
  }
  normal_from_keys<K>(key, q, pc4, max_nn, r2, cam, normals, flag);
}


// ---- estimate_normals, batched form (K = 16; the default) -------------------------------------------------------------
// Same result as normals_kernel<16> bit for bit (the exact 16 smallest (distance, index) keys, ascending), different search
// economics.  In the one-thread-per-query form nearly every candidate step pays a 16-deep insertion (~130 instructions)
// because SOME lane of the warp inserts, and lanes run different trip counts (12 of 32 lanes active, ncu).  Here
//   * the warp walks the candidate runs in lock-step (trip count = the longest run among its lanes, shorter lanes predicated
//     off), so warp votes are legal at every step;
//   * a candidate that beats the lane's current 16th key is only APPENDED to the lane's column of a shared-memory batch
//     (16 slots, conflict-free layout);
//   * when any lane's batch is full -- and at the end of a shell -- ALL lanes sort their batch (bitonic network in registers)
//     and merge it into their sorted top-16 (bitonic merge): ~750 instructions per 16 candidates instead of ~130 per candidate,
//     executed converged.
#define DFB_CE(a, b) { const bool s_ = (b) < (a); const unsigned long long t_ = s_ ? (b) : (a); (b) = s_ ? (a) : (b); (a) = t_; }
constexpr int NB_T = 128;              // threads per block
constexpr int NB_CAP = 16;             // batch slots per lane
__device__ __forceinline__ void nb_merge(unsigned long long (&top)[16], const unsigned long long* __restrict__ batch, int& cnt) {
  unsigned long long b[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) b[i] = i < cnt ? batch[i * NB_T] : 0x7f8000007fffffffull;
  cnt = 0;
  // bitonic sort, ascending
#pragma unroll
  for (int k = 2; k <= 16; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int l = i ^ j;
        if (l > i) {
          if ((i & k) == 0) { DFB_CE(b[i], b[l]); } else { DFB_CE(b[l], b[i]); }
        }
      }
    }
  }
  // the 16 smallest of top (ascending) and b (ascending): min(top[i], b[15 - i]) is bitonic; one bitonic merge sorts it
#pragma unroll
  for (int i = 0; i < 16; ++i) { const unsigned long long y = b[15 - i]; top[i] = y < top[i] ? y : top[i]; }
#pragma unroll
  for (int j = 8; j > 0; j >>= 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int l = i ^ j;
      if (l > i) DFB_CE(top[i], top[l]);
    }
  }
}

__global__ void __launch_bounds__(NB_T, 4) normals_batched_kernel(const float4* __restrict__ sorted, const float4* __restrict__ pc4, int n,
                                                                  const GridParams* gpp, const int* __restrict__ cell_start, int max_nn, float radius,
                                                                  float3 cam, float* __restrict__ normals, const int* __restrict__ n_dev, int* __restrict__ flag = nullptr,
                                                                  const uint8_t* __restrict__ dead = nullptr) {
  __shared__ unsigned long long s_batch[NB_CAP * NB_T];
  if (n_dev) n = *n_dev;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if ((j & ~31) >= n) return;                                        // whole warp past the end (warp-uniform)
  bool live = j < n;
  const GridParams g = *gpp;
  const float4 q = sorted[live ? j : n - 1];
  if (live && dead && dead[j]) {                                     // removed by the radius filter: no query, never a candidate
    if (flag) flag[__float_as_int(q.w)] = 0;
    live = false;
  }
  const int3 c = cell_coord(g, q.x, q.y, q.z);
  unsigned long long top[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) top[t] = 0x7f8000007fffffffull;       // (+inf, INT_MAX)
  unsigned long long* batch = s_batch + threadIdx.x;
  int cnt = 0;
  const float r2 = radius * radius;
  const int Rmax = max(1, (int)ceilf(radius * g.inv_cell));
  const int last = min(max_nn, 16) - 1;
  const CellFrac cf = cell_frac(g, q, c);
  bool done = !live;
  for (int R = 1; R <= Rmax; ++R) {
    for (int ix = 0; ix <= 2 * R; ++ix) {
      const int dx = (ix & 1) ? -((ix + 1) >> 1) : (ix >> 1);
      const int x = c.x + dx;
      const bool okx = !done && x >= 0 && x < g.nx;
      const float gx = cell_gap(dx, cf.x);
      for (int iy = 0; iy <= 2 * R; ++iy) {
        const int dy = (iy & 1) ? -((iy + 1) >> 1) : (iy >> 1);
        const int y = c.y + dy;
        const float gy = cell_gap(dy, cf.y);
        // (the 16th key may be stale by the unmerged batch: the bound is then only less tight)
        const bool okxy = okx && y >= 0 && y < g.ny && (gx * gx + gy * gy) * cf.cell2 < __uint_as_float((unsigned)(top[15] >> 32));
        const int row = (x * g.ny + y) * g.nz;
        const bool whole = (R == 1 || dx == -R || dx == R || dy == -R || dy == R);   // warp-uniform
        for (int sgm = 0; sgm < (whole ? 1 : 2); ++sgm) {
          int b = 0, e = 0;
          if (okxy) {
            if (whole) {
              b = cell_start[row + max(c.z - R, 0)]; e = cell_start[row + min(c.z + R, g.nz - 1) + 1];
            } else {
              const int z = sgm ? c.z + R : c.z - R;
              if (z >= 0 && z < g.nz) { b = cell_start[row + z]; e = cell_start[row + z + 1]; }
            }
          }
          const int len = e - b;
          const int L = __reduce_max_sync(0xffffffffu, len);
          for (int k = 0; k < L; ++k) {
            if (k < len && !(dead && dead[b + k])) {
              const float4 p = sorted[b + k];
              const float d = dist2(p.x, p.y, p.z, q.x, q.y, q.z);
              const unsigned long long nk = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
              if (d < r2 && nk < top[15]) { batch[cnt * NB_T] = nk; ++cnt; }
            }
            if (__any_sync(0xffffffffu, cnt == NB_CAP)) nb_merge(top, batch, cnt);
          }
        }
      }
    }
    if (__any_sync(0xffffffffu, cnt > 0)) nb_merge(top, batch, cnt);
    const float reach = (float)R * g.cell * 0.9999f;
    if (__uint_as_float((unsigned)(top[last] >> 32)) < reach * reach) done = true;
    if (__all_sync(0xffffffffu, done)) break;
  }
  if (live) normal_from_keys<16>(top, q, pc4, max_nn, r2, cam, normals, flag);
}

static bool normals_use_v1() { static const bool v = getenv("DFB_NORMALS_V1") != nullptr; return v; }   // A/B switch: one thread per query, immediate insertion

// cell: target cell edge (grown by 1.25x steps until the bounding box fits into `cap` cells)
__device__ __forceinline__ void bbox_init_inline(unsigned* bbox) {   // bbox_init_kernel's work, by threads 0..6 of a caller's block 0
  const int t = threadIdx.x;
  if (t < 3) bbox[t] = 0xffffffffu;
  else if (t < 6) bbox[t] = 0u;
  else if (t == 6) { bbox[7] = bbox[6]; bbox[6] = 0u; }
}
static int build_grid(const float* pc4, int n, float cell, int cap, GridWs& w, cudaStream_t s, const int* n_dev = nullptr,
                      const int* hint_n = nullptr, float cell_max = 0.f, bool bbox_ready = false, int* total_out = nullptr) {
  if (!bbox_ready) bbox_init_kernel<<<1, 32, 0, s>>>(w.bbox);     // (the fused path has the preceding compaction kernel do it)
  bbox_kernel<<<min(div_up(n, 256), 2 * sm_count()), 256, 0, s>>>(pc4, n, 4, w.bbox, n_dev);
  grid_params_kernel<<<1, 1, 0, s>>>(w.bbox, cell, cap, w.gp, hint_n, cell_max);
  // counts, scan and cursor reset cover the ncell + 1 cells the bounding box needs (device-side length), not the capacity
  zero_words_dev(w.cell_count, cap + 1, &w.gp->ncell1, s);
  grid_count_kernel<<<div_up(n, 256), 256, 0, s>>>(pc4, n, w.gp, w.cell_of, w.cell_count, n_dev, w.bbox);
  DFB_LAUNCH_CHECK();
  int rc = exclusive_scan_i32(w.cell_count, w.cell_start, cap + 1, w.block_sums, total_out, s, &w.gp->ncell1);   // total = rows in cells
  if (rc) return rc;
  grid_scatter_kernel<<<div_up(n, 256), 256, 0, s>>>(pc4, n, w.cell_of, w.cell_start, w.cell_count, w.sorted, n_dev);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------------
// scatter_mean / box filter: deterministic segmented mean
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seg_count_kernel(const int64_t* __restrict__ index64, const int* __restrict__ index32,
                                                        int n, int* seg_count, const int* __restrict__ n_dev) {
  if (n_dev) n = *n_dev;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int g = index64 ? (int)index64[i] : index32[i];
  atomicAdd(&seg_count[g], 1);
}

__global__ void __launch_bounds__(256) seg_fill_kernel(const int64_t* __restrict__ index64, const int* __restrict__ index32,
                                                       int n, const int* seg_start, int* cursor, int* members,
                                                       const int* __restrict__ n_dev) {
  if (n_dev) n = *n_dev;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int g = index64 ? (int)index64[i] : index32[i];
  members[seg_start[g] + atomicAdd(&cursor[g], 1)] = i;
}

// One thread per output row: visit the segment's members in ascending row order (selection by repeated minimum;
// segments are tiny) and sum sequentially -> bit-identical to a sequential CPU scatter (torch_scatter's CPU path).
template <int D>
__global__ void __launch_bounds__(128) seg_mean_kernel(const float* __restrict__ srcA, const float* __restrict__ srcB,
                                                       const int* __restrict__ seg_start, const int* __restrict__ members,
                                                       const int* __restrict__ n_out_dev, int n_out_host,
                                                       float* __restrict__ outA, float* __restrict__ outB) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  int n_out = n_out_dev ? *n_out_dev : n_out_host;
  if (g >= n_out) return;
  int b = seg_start[g], e = seg_start[g + 1];
  float sa[D], sb[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { sa[d] = 0.f; sb[d] = 0.f; }
  int last = -1;
  for (int t = b; t < e; ++t) {
    int best = 0x7fffffff;
    for (int u = b; u < e; ++u) {
      int m = members[u];
      if (m > last && m < best) best = m;
    }
    last = best;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      sa[d] += srcA[(size_t)best * D + d];
      if (srcB) sb[d] += srcB[(size_t)best * D + d];
    }
  }
  float cnt = (float)max(e - b, 1);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    outA[(size_t)g * D + d] = sa[d] / cnt;
    if (srcB) outB[(size_t)g * D + d] = sb[d] / cnt;
  }
}

// generic-D fallback (one thread per (row, column))
__global__ void __launch_bounds__(128) seg_mean_generic_kernel(const float* __restrict__ src, int D,
                                                               const int* __restrict__ seg_start,
                                                               const int* __restrict__ members, int n_out,
                                                               float* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_out * D) return;
  int g = t / D, d = t - g * D;
  int b = seg_start[g], e = seg_start[g + 1];
  float s = 0.f;
  int last = -1;
  for (int k = b; k < e; ++k) {
    int best = 0x7fffffff;
    for (int u = b; u < e; ++u) {
      int m = members[u];
      if (m > last && m < best) best = m;
    }
    last = best;
    s += src[(size_t)best * D + d];
  }
  out[t] = s / (float)max(e - b, 1);
}

// box filter front end (tracker.py:16-21)
struct BoxParams {
  float mnx, mny, mnz;
  long long nx, ny, nz;
  int n_words;
  int overflow;
  int n_zero;          // n_words + 1: words of the bitmap to clear
};
constexpr long long BOX_BITS_CAP = 1ll << 27;  // 16 MiB bitmap

__global__ void box_params_kernel(const unsigned* bbox, float voxel_size, int div_mode, BoxParams* bp) {
  // min_bound = min - vs*0.5, max_bound = max + vs*0.5 (fp32 tensors; vs*0.5 is a Python double rounded to fp32)
  float half = (float)((double)voxel_size * 0.5);
  float inv = 1.0f / voxel_size;
  float mn[3], ext[3];
  for (int a = 0; a < 3; ++a) {
    float lo = ord2f(bbox[a]), hi = ord2f(bbox[3 + a]);
    if (!(lo <= hi)) { lo = 0.f; hi = 0.f; }       // empty input (the fused path may legitimately have 0 rows)
    mn[a] = __fsub_rn(lo, half);
    float mx = __fadd_rn(hi, half);
    ext[a] = floorf(div_vs(__fsub_rn(mx, mn[a]), voxel_size, inv, div_mode));
  }
  bp->mnx = mn[0]; bp->mny = mn[1]; bp->mnz = mn[2];
  bp->nx = (long long)ext[0] + 16; bp->ny = (long long)ext[1] + 16; bp->nz = (long long)ext[2] + 16;
  long long bits = bp->nx * bp->ny * bp->nz;
  bp->overflow = bits > BOX_BITS_CAP ? 1 : 0;
  if (bp->overflow) bits = 0;
  bp->n_words = (int)((bits + 31) / 32);
  bp->n_zero = bp->n_words + 1;
}

__global__ void __launch_bounds__(256) box_key_kernel(const float* __restrict__ pts, int n, float voxel_size,
                                                      int div_mode, const BoxParams* bpp, long long* keys,
                                                      uint32_t* bits, const int* __restrict__ n_dev) {
  if (n_dev) n = *n_dev;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  BoxParams bp = *bpp;
  if (bp.overflow) return;
  float inv = 1.0f / voxel_size;
  long long cx = (long long)floorf(div_vs(__fsub_rn(pts[3 * i + 0], bp.mnx), voxel_size, inv, div_mode));
  long long cy = (long long)floorf(div_vs(__fsub_rn(pts[3 * i + 1], bp.mny), voxel_size, inv, div_mode));
  long long cz = (long long)floorf(div_vs(__fsub_rn(pts[3 * i + 2], bp.mnz), voxel_size, inv, div_mode));
  long long key = cx + cy * bp.nx + cz * bp.nx * bp.ny;   // tracker.py:20
  if (key < 0 || key >= bp.nx * bp.ny * bp.nz) key = 0;   // non-finite input row; the reference would fault
  keys[i] = key;
  atomicOr(&bits[key >> 5], 1u << (key & 31));
}

__global__ void __launch_bounds__(256) box_rank_kernel(const long long* __restrict__ keys, int n,
                                                       const uint32_t* __restrict__ bits, const int* __restrict__ word_rank,
                                                       const BoxParams* bpp, int* rank_out, const int* __restrict__ n_dev, int32_t* n_out) {
  // (key range beyond the bitmap capacity: the row count becomes -1 here, before the segmented mean reads it, so that pass does
  // nothing and the host raises; this was a one-thread kernel of its own)
  if (blockIdx.x == 0 && threadIdx.x == 0 && bpp->overflow) *n_out = -1;
  if (n_dev) n = *n_dev;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (bpp->overflow) { rank_out[i] = 0; return; }
  long long key = keys[i];
  uint32_t w = bits[key >> 5];
  rank_out[i] = word_rank[key >> 5] + __popc(w & ((1u << (key & 31)) - 1u));
}


// ------------------------------------------------------------------------------------------------
// groupby_sum (indexing.cu:59-71): one thread per element, coalesced reads, one red per element
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) groupby_sum_kernel(const float* __restrict__ values, const int64_t* __restrict__ indices,
                                                          long long total, int L, int C, float* sum, int32_t* count) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int i = (int)(t / L), l = (int)(t - (long long)i * L);
  long long g = indices[i];
  if (g < 0 || g >= C) return;
  atomicAdd(&sum[g * L + l], values[t]);
  if (l == 0) atomicAdd(&count[g], L);   // the reference bumps the count once per ELEMENT (indexing.cu:69-70), i.e. L per row
}

// ---- fused preprocessing helpers -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unproject_sub_kernel(const float* __restrict__ depth, int H, int W, int Hs, int Ws, float fx, float fy,
                                                            float cx, float cy, float* __restrict__ pc4, int* __restrict__ flag,
                                                            int nan_invalid = 0, unsigned* init_bbox = nullptr) {
  if (init_bbox && blockIdx.x == 0) bbox_init_inline(init_bbox);     // for the grid build that follows (no compaction in between)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Hs * Ws) return;
  const int v = i / Ws, u = i - v * Ws;
  const float d = depth[(size_t)(2 * v) * W + 2 * u];       // F.interpolate(scale 0.5, nearest) == depth[2v][2u]
  const bool ok = !isnan(d);
  // nan_invalid: the cloud stays uncompacted; its invalid rows are NaN so that the bounding box and the grid skip them
  float4 p = nan_invalid ? make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) { p.x = ((float)u - cx) / fx * d; p.y = ((float)v - cy) / fy * d; p.z = d; }
  reinterpret_cast<float4*>(pc4)[i] = p;
  flag[i] = ok ? 1 : 0;
}

// order-preserving compaction (the reference's boolean-mask indexing keeps row order, tracker.py:104-116)
__global__ void __launch_bounds__(256) compact4_kernel(const float* __restrict__ in4, const int* __restrict__ flag, const int* __restrict__ pos,
                                                       int n, const int* __restrict__ n_dev, float* __restrict__ out4, unsigned* init_bbox = nullptr) {
  if (init_bbox && blockIdx.x == 0) bbox_init_inline(init_bbox);     // for the grid build that follows
  if (n_dev) n = *n_dev;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  reinterpret_cast<float4*>(out4)[pos[i]] = reinterpret_cast<const float4*>(in4)[i];
}

__global__ void __launch_bounds__(256) compact_pn_kernel(const float* __restrict__ pc4, const float* __restrict__ nrm, const int* __restrict__ flag,
                                                         const int* __restrict__ pos, int n, const int* __restrict__ n_dev,
                                                         float* __restrict__ p3, float* __restrict__ n3, unsigned* init_bbox = nullptr) {
  if (init_bbox && blockIdx.x == 0) bbox_init_inline(init_bbox);     // for the box filter's bounding box
  if (n_dev) n = *n_dev;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  const int o = pos[i];
#pragma unroll
  for (int a = 0; a < 3; ++a) { p3[3 * (size_t)o + a] = pc4[4 * (size_t)i + a]; n3[3 * (size_t)o + a] = nrm[3 * (size_t)i + a]; }
}

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_unproject_depth(const float* depth, int H, int W, float fx, float fy, float cx, float cy, float* pc,
                        void* stream) {
  DFB_CHECK_ARG(H >= 0 && W >= 0, "unproject_depth");
  if (H * W == 0) return DFB_OK;
  DFB_CHECK_ARG(depth && pc, "unproject_depth: null pointer");
  unproject_kernel<<<div_up((long long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(depth, H, W, fx, fy, cx, cy, pc);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

size_t dfb_pcproc_ws_bytes(int n) {
  Arena a(nullptr, 0);
  GridWs w;
  return grid_ws_layout(a, n, &w) + 256;
}

int dfb_remove_radius_outlier(const float* pc4, int n, int nb_points, float radius, uint8_t* mask, void* ws,
                              size_t ws_bytes, void* stream) {
  DFB_CHECK_ARG(n >= 0 && nb_points > 0 && radius > 0.f, "remove_radius_outlier");
  if (n == 0) return DFB_OK;
  DFB_CHECK_ARG(pc4 && mask && ws, "remove_radius_outlier: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  Arena a(ws, ws_bytes);
  GridWs w;
  grid_ws_layout(a, n, &w);
  if (!a.ok()) { set_error("workspace too small: need %zu", a.off); return DFB_E_WORKSPACE; }
  int rc = build_grid(pc4, n, radius, GRID_CAP_COARSE, w, s);
  if (rc) return rc;
  radius_count_kernel<<<div_up(n, 128), 128, 0, s>>>(w.sorted, n, w.gp, w.cell_start, nb_points, radius, mask, nullptr);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_transform_points(const float* xyz, int n, const float* h_R, const float* h_t, float* out, void* stream) {
  DFB_CHECK_ARG(n >= 0 && h_R, "transform_points");
  if (n == 0) return DFB_OK;
  DFB_CHECK_ARG(xyz && out, "transform_points: null pointer");
  Rt9 R;
  for (int i = 0; i < 9; ++i) R.r[i] = h_R[i];
  for (int i = 0; i < 3; ++i) R.t[i] = h_t ? h_t[i] : 0.f;
  transform_points_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(xyz, n, R, h_t != nullptr, out);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_ingest_frame(const uint16_t* depth_raw, const uint8_t* color_raw, int H, int W, float depth_scale, int div_mode, float cut_min,
                     float cut_max, int bgr, float* depth_out, float* rgb_out, void* stream) {
  DFB_CHECK_ARG(H >= 0 && W >= 0 && depth_scale > 0.f, "ingest_frame");
  if (H * W == 0) return DFB_OK;
  DFB_CHECK_ARG((depth_raw == nullptr) == (depth_out == nullptr) && (color_raw == nullptr) == (rgb_out == nullptr), "ingest_frame: in/out pairs");
  ingest_kernel<<<div_up((long long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(depth_raw, color_raw, H * W, depth_scale, 1.0f / depth_scale,
                                                                                div_mode, cut_min, cut_max, bgr, depth_out, rgb_out);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_estimate_normals(const float* pc4, int n, int max_nn, float radius, const float* h_cam_xyz, float* normals,
                         void* ws, size_t ws_bytes, void* stream) {
  DFB_CHECK_ARG(n >= 0 && max_nn > 1 && max_nn <= 32 && radius > 0.f && h_cam_xyz, "estimate_normals");
  if (n == 0) return DFB_OK;
  DFB_CHECK_ARG(pc4 && normals && ws, "estimate_normals: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  Arena a(ws, ws_bytes);
  GridWs w;
  grid_ws_layout(a, n, &w);
  if (!a.ok()) { set_error("workspace too small: need %zu", a.off); return DFB_E_WORKSPACE; }
  int rc = build_grid(pc4, n, radius / NORMAL_SUB, GRID_CAP, w, s);
  if (rc) return rc;
  float3 cam = make_float3(h_cam_xyz[0], h_cam_xyz[1], h_cam_xyz[2]);
  if (max_nn <= 16 && !normals_use_v1())
    normals_batched_kernel<<<div_up(n, NB_T), NB_T, 0, s>>>(w.sorted, reinterpret_cast<const float4*>(pc4), n, w.gp, w.cell_start, max_nn, radius, cam, normals, nullptr);
  else if (max_nn <= 16)
    normals_kernel<16><<<div_up(n, 128), 128, 0, s>>>(w.sorted, reinterpret_cast<const float4*>(pc4), n, w.gp, w.cell_start, max_nn, radius, cam, normals, nullptr);
  else
    normals_kernel<32><<<div_up(n, 128), 128, 0, s>>>(w.sorted, reinterpret_cast<const float4*>(pc4), n, w.gp, w.cell_start, max_nn, radius, cam, normals, nullptr);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

size_t dfb_scatter_mean_ws_bytes(int n, int n_out) {
  Arena a(nullptr, 0);
  a.take<int>(n_out + 2); a.take<int>(n_out + 2); a.take<int>(n_out + 2); a.take<int>(n + 1);
  a.take<int>(n_out / 2048 + 8);
  return a.off + 256;
}

static int segmented_mean(const float* srcA, const float* srcB, const int64_t* index64, const int* index32, int n, int d,
                          int n_out_cap, const int* n_out_dev, float* outA, float* outB, Arena& a, cudaStream_t s,
                          const int* n_dev = nullptr) {
  int* seg_count = a.take<int>(n_out_cap + 2);
  int* seg_start = a.take<int>(n_out_cap + 2);
  int* cursor = a.take<int>(n_out_cap + 2);
  int* members = a.take<int>(n + 1);
  int* bsums = a.take<int>(n_out_cap / 2048 + 8);
  if (!a.ok()) { set_error("workspace too small: need %zu", a.off); return DFB_E_WORKSPACE; }
  DFB_CUDA(cudaMemsetAsync(seg_count, 0, sizeof(int) * (n_out_cap + 2), s));
  DFB_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * (n_out_cap + 2), s));
  seg_count_kernel<<<div_up(n, 256), 256, 0, s>>>(index64, index32, n, seg_count, n_dev);
  DFB_LAUNCH_CHECK();
  int rc = exclusive_scan_i32(seg_count, seg_start, n_out_cap + 1, bsums, nullptr, s);
  if (rc) return rc;
  seg_fill_kernel<<<div_up(n, 256), 256, 0, s>>>(index64, index32, n, seg_start, cursor, members, n_dev);
  if (d == 3)
    seg_mean_kernel<3><<<div_up(n_out_cap, 128), 128, 0, s>>>(srcA, srcB, seg_start, members, n_out_dev, n_out_cap, outA, outB);
  else {
    seg_mean_generic_kernel<<<div_up((long long)n_out_cap * d, 128), 128, 0, s>>>(srcA, d, seg_start, members, n_out_cap, outA);
    if (srcB) seg_mean_generic_kernel<<<div_up((long long)n_out_cap * d, 128), 128, 0, s>>>(srcB, d, seg_start, members, n_out_cap, outB);
  }
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_scatter_mean(const float* src, const int64_t* index, int n, int d, int n_out, float* out, void* ws,
                     size_t ws_bytes, void* stream) {
  DFB_CHECK_ARG(n >= 0 && d > 0 && n_out >= 0, "scatter_mean");
  if (n_out == 0) return DFB_OK;
  DFB_CHECK_ARG(out && ws && (n == 0 || (src && index)), "scatter_mean: null pointer");
  Arena a(ws, ws_bytes);
  return segmented_mean(src, nullptr, index, nullptr, n, d, n_out, nullptr, out, nullptr, a, (cudaStream_t)stream);
}

size_t dfb_box_filter_ws_bytes(int n) {
  Arena a(nullptr, 0);
  a.take<unsigned>(8); a.take<BoxParams>(1); a.take<long long>(n + 1); a.take<uint32_t>(BOX_BITS_CAP / 32 + 1);
  a.take<int>(BOX_BITS_CAP / 32 + 2); a.take<int>(BOX_BITS_CAP / 32 / 2048 + 8); a.take<int>(n + 1);
  return a.off + dfb_scatter_mean_ws_bytes(n, n) + 256;
}

static int box_filter_impl(const float* points, const float* normals, int n, const int* n_dev, float voxel_size, int div_mode,
                           float* out_points, float* out_normals, int32_t* d_n_out, Arena& a, cudaStream_t s, unsigned* ready_bbox = nullptr) {
  unsigned* bbox = ready_bbox ? ready_bbox : a.take<unsigned>(8);
  BoxParams* bp = a.take<BoxParams>(1);
  long long* keys = a.take<long long>(n + 1);
  const int max_words = (int)(BOX_BITS_CAP / 32);
  uint32_t* bits = a.take<uint32_t>(max_words + 1);
  int* word_rank = a.take<int>(max_words + 2);
  int* bsums = a.take<int>(max_words / 2048 + 8);
  int* rank = a.take<int>(n + 1);
  if (!a.ok()) { set_error("workspace too small: need %zu", a.off); return DFB_E_WORKSPACE; }
  if (!ready_bbox) bbox_init_kernel<<<1, 32, 0, s>>>(bbox);
  bbox_kernel<<<min(div_up(n, 256), 2 * sm_count()), 256, 0, s>>>(points, n, 3, bbox, n_dev);
  box_params_kernel<<<1, 1, 0, s>>>(bbox, voxel_size, div_mode, bp);
  zero_words_dev(bits, max_words + 1, &bp->n_zero, s);        // only the words this frame's bounding box uses
  box_key_kernel<<<div_up(n, 256), 256, 0, s>>>(points, n, voxel_size, div_mode, bp, keys, bits, n_dev);
  DFB_LAUNCH_CHECK();
  int rc = exclusive_scan_popc(bits, word_rank, max_words, bsums, d_n_out, s, &bp->n_words);
  if (rc) return rc;
  box_rank_kernel<<<div_up(n, 256), 256, 0, s>>>(keys, n, bits, word_rank, bp, rank, n_dev, d_n_out);
  DFB_LAUNCH_CHECK();
  rc = segmented_mean(points, normals, nullptr, rank, n, 3, n, d_n_out, out_points, out_normals, a, s, n_dev);
  return rc;
}

int dfb_point_box_filter(const float* points, const float* normals, int n, float voxel_size, int div_mode,
                         float* out_points, float* out_normals, int32_t* d_n_out, void* ws, size_t ws_bytes,
                         void* stream) {
  DFB_CHECK_ARG(n >= 0 && voxel_size > 0.f && d_n_out, "point_box_filter");
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) { DFB_CUDA(cudaMemsetAsync(d_n_out, 0, sizeof(int32_t), s)); return DFB_OK; }
  DFB_CHECK_ARG(points && normals && out_points && out_normals && ws, "point_box_filter: null pointer");
  Arena a(ws, ws_bytes);
  return box_filter_impl(points, normals, n, nullptr, voxel_size, div_mode, out_points, out_normals, d_n_out, a, s);
}

// ------------------------------------------------------------------------------------------------
// fused per-frame preprocessing (tracker.py:89-120): no host synchronisation, sizes stay on the device
// ------------------------------------------------------------------------------------------------
size_t dfb_preprocess_ws_bytes(int H, int W) {
  const int n = (H / 2) * (W / 2) + 16;
  Arena a(nullptr, 0);
  a.take<float>((size_t)n * 4); a.take<int>(n + 1); a.take<int>(n + 1); a.take<int>(n / 2048 + 8); a.take<int>(8);
  a.take<float>((size_t)n * 4); a.take<float>((size_t)n * 4); a.take<uint8_t>(n + 16); a.take<float>((size_t)n * 3);
  a.take<float>((size_t)n * 3); a.take<float>((size_t)n * 3); a.take<unsigned>(8); a.take<uint8_t>(n + 32);
  return a.off + dfb_pcproc_ws_bytes(n) + dfb_box_filter_ws_bytes(n) + 1024;
}

int dfb_preprocess_frame(const float* depth, int H, int W, float fx, float fy, float cx, float cy, int nb_points,
                         float outlier_radius, int max_nn, float normal_radius, const float* h_cam_xyz, float box_voxel,
                         int div_mode, float* out_points, float* out_normals, int32_t* d_n_out, void* ws, size_t ws_bytes,
                         void* stream) {
  DFB_CHECK_ARG(depth && H >= 2 && W >= 2 && out_points && out_normals && d_n_out && ws && h_cam_xyz, "preprocess_frame");
  DFB_CHECK_ARG(max_nn > 1 && max_nn <= 32 && nb_points > 0, "preprocess_frame: bad neighbour counts");
  cudaStream_t s = (cudaStream_t)stream;
  const int Hs = H / 2, Ws = W / 2, n = Hs * Ws;
  Arena a(ws, ws_bytes);
  float* pcA = a.take<float>((size_t)(n + 16) * 4);        // dense (uncompacted) unprojection
  int* flag = a.take<int>(n + 17);
  int* pos = a.take<int>(n + 17);
  int* bsums = a.take<int>((n + 16) / 2048 + 8);
  int* counts = a.take<int>(8);                             // [0] nA, [1] nB, [2] nC
  float* pcB_buf = a.take<float>((size_t)(n + 16) * 4);
  const float* pcB = pcB_buf;
  float* pcC = a.take<float>((size_t)(n + 16) * 4);
  uint8_t* mask = a.take<uint8_t>(n + 32);
  float* nrmC = a.take<float>((size_t)(n + 16) * 3);
  float* pD = a.take<float>((size_t)(n + 16) * 3);
  float* nD = a.take<float>((size_t)(n + 16) * 3);
  GridWs w;
  grid_ws_layout(a, n + 16, &w);
  if (!a.ok()) { set_error("workspace too small: need %zu", a.off); return DFB_E_WORKSPACE; }
  // One grid serves both searches when its cells (edge = the outlier radius) are no larger than half the normal radius (the
  // reference's 0.05 / 0.1 m): the cloud stays UNCOMPACTED from the unprojection to the box filter (invalid rows are NaN and in
  // no cell), the radius filter marks the points it removes by position in the sorted array, the kNN search skips them as
  // queries and as candidates -- no second grid build (7 launches) and no compaction before either stage (2 x 3 launches).
  // Row indices are pixel indices, whose order is the order of the compacted clouds: same neighbours, same tie-breaks.
  const bool one_grid = outlier_radius * 2.0f <= normal_radius * 1.0001f && getenv("DFB_TWO_GRIDS") == nullptr;
  int rc;
  // P1: nearest x0.5 subsample (tracker.py:91-93) + unproject with halved intrinsics (:97-98) + validity flags
  if (one_grid) {
    unproject_sub_kernel<<<div_up(n, 256), 256, 0, s>>>(depth, H, W, Hs, Ws, fx * 0.5f, fy * 0.5f, cx * 0.5f, cy * 0.5f, pcA, flag, 1, w.bbox);
    DFB_LAUNCH_CHECK();
    pcB = pcA;
    // P2: radius outlier filter; counts[0] = rows that are in a cell = valid rows (the scan's total)
    rc = build_grid(pcB, n, outlier_radius, GRID_CAP_COARSE, w, s, nullptr, nullptr, 0.f, true, &counts[0]);
    if (rc) return rc;
  } else {
    unproject_sub_kernel<<<div_up(n, 256), 256, 0, s>>>(depth, H, W, Hs, Ws, fx * 0.5f, fy * 0.5f, cx * 0.5f, cy * 0.5f, pcA, flag);
    DFB_LAUNCH_CHECK();
    rc = exclusive_scan_i32(flag, pos, n, bsums, &counts[0], s);
    if (rc) return rc;
    compact4_kernel<<<div_up(n, 256), 256, 0, s>>>(pcA, flag, pos, n, nullptr, pcB_buf, w.bbox);
    // P2: radius outlier filter (the kernel writes the compaction flags itself; rows past the device-side count are not scanned)
    rc = build_grid(pcB, n, outlier_radius, GRID_CAP_COARSE, w, s, &counts[0], nullptr, 0.f, true);
    if (rc) return rc;
  }
  const float3 cam = make_float3(h_cam_xyz[0], h_cam_xyz[1], h_cam_xyz[2]);
  unsigned* box_bbox = a.take<unsigned>(8);
  uint8_t* dead = a.take<uint8_t>(n + 32);
  if (!a.ok()) { set_error("workspace too small: need %zu", a.off); return DFB_E_WORKSPACE; }
  const float* pcN = pcB;                                   // the cloud the normals index
  const int* nN = &counts[0];
  radius_count_kernel<<<div_up(n, 128), 128, 0, s>>>(w.sorted, n, w.gp, w.cell_start, nb_points, outlier_radius, mask, &counts[0], flag,
                                                     one_grid ? dead : nullptr);
  DFB_LAUNCH_CHECK();
  if (!one_grid) {
    rc = exclusive_scan_i32(flag, pos, n, bsums, &counts[1], s, &counts[0]);
    if (rc) return rc;
    compact4_kernel<<<div_up(n, 256), 256, 0, s>>>(pcB, flag, pos, n, &counts[0], pcC, w.bbox);
    // P3: normals (cell size from the density the radius filter's grid just measured: grid_params_kernel)
    rc = build_grid(pcC, n, normal_radius / NORMAL_SUB, GRID_CAP, w, s, &counts[1], &counts[0], normal_radius * 0.5f, true);
    if (rc) return rc;
    pcN = pcC; nN = &counts[1];
  }
  const uint8_t* dd = one_grid ? dead : nullptr;
  if (max_nn <= 16 && !normals_use_v1())
    normals_batched_kernel<<<div_up(n, NB_T), NB_T, 0, s>>>(w.sorted, reinterpret_cast<const float4*>(pcN), n, w.gp, w.cell_start, max_nn, normal_radius, cam, nrmC, nN, flag, dd);
  else if (max_nn <= 16)
    normals_kernel<16><<<div_up(n, 128), 128, 0, s>>>(w.sorted, reinterpret_cast<const float4*>(pcN), n, w.gp, w.cell_start, max_nn, normal_radius, cam, nrmC, nN, flag, dd);
  else
    normals_kernel<32><<<div_up(n, 128), 128, 0, s>>>(w.sorted, reinterpret_cast<const float4*>(pcN), n, w.gp, w.cell_start, max_nn, normal_radius, cam, nrmC, nN, flag, dd);
  DFB_LAUNCH_CHECK();
  // (one grid: rows are pixel indices, all n of them carry a flag -- invalid pixels 0 from the unprojection)
  rc = exclusive_scan_i32(flag, pos, n, bsums, &counts[2], s, one_grid ? nullptr : nN);
  if (rc) return rc;
  compact_pn_kernel<<<div_up(n, 256), 256, 0, s>>>(pcN, nrmC, flag, pos, n, one_grid ? nullptr : nN, pD, nD, box_bbox);
  DFB_LAUNCH_CHECK();
  // P4: box filter
  return box_filter_impl(pD, nD, n, &counts[2], box_voxel, div_mode, out_points, out_normals, d_n_out, a, s, box_bbox);
}

int dfb_groupby_sum(const float* values, const int64_t* indices, int n, int L, int C, float* sum, int32_t* count,
                    void* stream) {
  DFB_CHECK_ARG(n >= 0 && L > 0 && C >= 0, "groupby_sum");
  cudaStream_t s = (cudaStream_t)stream;
  if (C == 0) return DFB_OK;
  DFB_CHECK_ARG(sum && count, "groupby_sum: null output");
  DFB_CUDA(cudaMemsetAsync(sum, 0, sizeof(float) * (size_t)C * L, s));
  DFB_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t) * (size_t)C, s));
  if (n == 0) return DFB_OK;
  long long total = (long long)n * L;
  groupby_sum_kernel<<<div_up(total, 256), 256, 0, s>>>(values, indices, total, L, C, sum, count);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

}  // extern "C"
