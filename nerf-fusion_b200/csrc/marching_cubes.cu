// Stage 5: sparse marching cubes with std-weighted cross-voxel blending.
// Reference: system/ext/marching_cubes/mc_interp_kernel.cu:7-382 (one thread per (voxel, sub-cell); each thread
// re-blends its 8 corners = 64 cube loads + 64 indexer/mapping lookups, and appends triangles with one global
// atomic each).  Here one CTA owns one voxel: the (r+1)^3 blended corner values are computed ONCE into shared
// memory (8x fewer blends), cells read them from there, and triangles are appended with one atomic per CTA batch
// (CTA-level prefix sum), so the output is ordered within a voxel.
#include <algorithm>

#include "common.cuh"
#include "mc_tables.cuh"

namespace dfb {

constexpr int MC_T = 128;

struct McArgs {
  const int64_t* indexer;
  int nx, ny, nz;
  const int64_t* valid_blocks;
  int U;
  const int32_t* mapping;
  int n_map;
  const float* cube_sdf;
  const float* cube_std;
  int B, r;
  float max_std;
  int max_tri;
  float* tri;
  int64_t* flat_id;
  float* tri_std;
  int* n_tri;
};

// mc_interp_kernel.cu:7-17: cube batch of voxel (bx, by, bz), -1 when it is outside the grid, unallocated or not in this
// extraction.  The CTA resolves its 27 neighbours once (nb[] in shared memory) instead of once per corner per neighbour.
__device__ __forceinline__ int voxel_batch(const McArgs& A, int bx, int by, int bz) {
  if (bx < 0 || by < 0 || bz < 0 || bx >= A.nx || by >= A.ny || bz >= A.nz) return -1;   // uint wrap in the reference
  const long long vec = A.indexer[((long long)bx * A.ny + by) * A.nz + bz];
  if (vec == -1 || vec >= A.n_map) return -1;
  return A.mapping[vec];
}

// mc_interp_kernel.cu:18-29; (dx, dy, dz) in {-1, 0, 1}: which neighbour of the CTA's voxel
__device__ __forceinline__ float2 query_raw(const McArgs& A, const int* nb, int dx, int dy, int dz, int ax, int ay, int az) {
  const float2 nan2 = make_float2(CUDART_NAN_F, CUDART_NAN_F);
  const int batch = nb[(dx + 1) * 9 + (dy + 1) * 3 + (dz + 1)];
  if (batch == -1) return nan2;
  const int R = 2 * A.r;
  const size_t off = (((size_t)batch * R + ax) * R + ay) * R + az;
  return make_float2(A.cube_sdf[off], A.cube_std[off]);
}

// mc_interp_kernel.cu:34-185 (STD_W_SDF variant): tent-weighted blend over the 2x2x2 nearest voxel centres, each
// weight multiplied by that voxel's predicted std; NaN if the voxel owning this half is missing.
__device__ float2 blend(const McArgs& A, const int* nb, int rx, int ry, int rz) {
  const int r = A.r;
  const int rbound = (r - 1) / 2, rstart = r / 2;
  const float rmid = r / 2.0f;
  int bm[3], rm[3], bp[3], rp[3], zero[3];
  float wm[3], wp[3];
  const int rpos[3] = {rx, ry, rz};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (rpos[a] <= rbound) {
      bm[a] = -1; rm[a] = r; bp[a] = 0; rp[a] = 0;
      wp[a] = (float)rpos[a] + rmid; wm[a] = rmid - (float)rpos[a];
      zero[a] = 1;
    } else {
      bm[a] = 0; rm[a] = 0; bp[a] = 1; rp[a] = -r;
      wp[a] = (float)rpos[a] - rmid; wm[a] = rmid + r - (float)rpos[a];
      zero[a] = 0;
    }
    wm[a] /= r; wp[a] /= r;
  }
  const int qx = rx + rstart, qy = ry + rstart, qz = rz + rstart;
  const int zero_det = zero[0] * 4 + zero[1] * 2 + zero[2];
  float tot_sdf = 0.f, tot_wsdf = 0.f, tot_std = 0.f, tot_w = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int sx = (c >> 2) & 1, sy = (c >> 1) & 1, sz = c & 1;
    const float2 v = query_raw(A, nb, sx ? bp[0] : bm[0], sy ? bp[1] : bm[1], sz ? bp[2] : bm[2],
                               qx + (sx ? rp[0] : rm[0]), qy + (sy ? rp[1] : rm[1]), qz + (sz ? rp[2] : rm[2]));
    const float w = (sx ? wp[0] : wm[0]) * (sy ? wp[1] : wm[1]) * (sz ? wp[2] : wm[2]);
    if (!isnan(v.x)) {
      tot_sdf += v.x * w * v.y; tot_wsdf += w * v.y;
      tot_std += w * v.y; tot_w += w;
    } else if (zero_det == c) {
      return make_float2(CUDART_NAN_F, CUDART_NAN_F);
    }
  }
  return make_float2(tot_sdf / tot_wsdf, tot_std / tot_w);
}

// Edge e joins corners edge_a(e) and edge_b(e); corner c sits at offset (dx, dy, dz) = bits 0, 1, 2 of corner_off(c)
// (mc_interp_kernel.cu:236-266).  The tables {0,1,2,3,4,5,6,7,0,1,2,3}, {1,2,3,0,5,6,7,4,4,5,6,7} and {0,1,3,2,4,5,7,6} are
// evaluated arithmetically: lanes of a warp look up different entries, which a __constant__ table would serialise.
__device__ __forceinline__ int edge_a(int e) { return e < 8 ? e : e - 8; }
__device__ __forceinline__ int edge_b(int e) { return e < 8 ? (e & 4) | ((e + 1) & 3) : e - 4; }
__device__ __forceinline__ int corner_off(int c) { return c ^ ((c >> 1) & 1); }

// mc_interp_kernel.cu:187-200
__device__ __forceinline__ float4 sdf_interp(float3 p1, float3 p2, float s1, float s2, float v1, float v2) {
  if (fabsf(0.0f - v1) < 1.0e-5f) return make_float4(p1.x, p1.y, p1.z, s1);
  if (fabsf(0.0f - v2) < 1.0e-5f) return make_float4(p2.x, p2.y, p2.z, s2);
  if (fabsf(v1 - v2) < 1.0e-5f) return make_float4(p1.x, p1.y, p1.z, s1);
  const float w2 = (0.0f - v1) / (v2 - v1);
  const float w1 = 1 - w2;
  return make_float4(p1.x * w1 + p2.x * w2, p1.y * w1 + p2.y * w2, p1.z * w1 + p2.z * w2, s1 * w1 + s2 * w2);
}

extern __shared__ __align__(16) unsigned char mc_smem_raw[];

__global__ void __launch_bounds__(MC_T) mc_kernel(McArgs A) {
  float2* corner = reinterpret_cast<float2*>(mc_smem_raw);
  __shared__ int s_warp[MC_T / 32];
  __shared__ int s_base;
  __shared__ int nb[27];
  const int r = A.r, r1 = r + 1;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float sbs = 1.0f / r;
  for (int lif = blockIdx.x; lif < A.U; lif += gridDim.x) {
    const long long vb = A.valid_blocks[lif];
    const int bx = (int)((vb / ((long long)A.ny * A.nz)) % A.nx), by = (int)((vb / A.nz) % A.ny), bz = (int)(vb % A.nz);
    __syncthreads();
    if (tid < 27) nb[tid] = voxel_batch(A, bx + tid / 9 - 1, by + (tid / 3) % 3 - 1, bz + tid % 3 - 1);
    __syncthreads();
    for (int c = tid; c < r1 * r1 * r1; c += MC_T) {
      const int cx = c / (r1 * r1), cy = (c / r1) % r1, cz = c % r1;
      corner[c] = blend(A, nb, cx, cy, cz);
    }
    __syncthreads();
    const int r3 = r * r * r;
    for (int cell0 = 0; cell0 < r3; cell0 += MC_T) {
      const int cell = cell0 + tid;
      int ntri = 0;
      int cube_type = 0;
      float sv[8], sd[8];
      int rx = 0, ry = 0, rz = 0;
      if (cell < r3) {
        rx = cell / (r * r); ry = (cell / r) % r; rz = cell % r;
        bool alive = true;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int o = corner_off(c);
          const float2 v = corner[((rx + (o & 1)) * r1 + (ry + ((o >> 1) & 1))) * r1 + (rz + ((o >> 2) & 1))];
          sv[c] = v.x; sd[c] = v.y;
          alive = alive && !isnan(v.x);
        }
        if (alive) {
#pragma unroll
          for (int c = 0; c < 8; ++c) cube_type |= (sv[c] < 0.f ? 1 : 0) << c;
        }
      }
      // pass 1: count surviving triangles of this cell; pass 2: write them
      uint64_t row = kMcTriTable[cube_type];
      int my_off = 0;
      // A vertex std is a convex combination of two corner stds: when no corner exceeds max_std (always, with the
      // reference's default 2000) no triangle can be dropped and the counting pass only walks the table.
      bool may_drop = false;
      if (cube_type != 0) {
        float smax = sd[0];
#pragma unroll
        for (int c = 1; c < 8; ++c) smax = fmaxf(smax, sd[c]);
        may_drop = !(smax <= A.max_std);
      }
      for (int pass = 0; pass < 2; ++pass) {
        int k = 0;
        uint64_t rr = row;
        if (pass == 0 && !may_drop) {
          while ((rr & 0xF) != 0xF) { ++k; rr >>= 12; }
        }
        while ((rr & 0xF) != 0xF) {
          float4 vp[3];
#pragma unroll
          for (int vi = 0; vi < 3; ++vi) {
            const int e = (int)((rr >> (4 * vi)) & 0xF);
            const int ca = edge_a(e), cb = edge_b(e);
            const int oa = corner_off(ca), ob = corner_off(cb);
            const int ax = rx + (oa & 1), ay = ry + ((oa >> 1) & 1), az = rz + ((oa >> 2) & 1);
            const int cx = rx + (ob & 1), cy = ry + ((ob >> 1) & 1), cz = rz + ((ob >> 2) & 1);
            const float3 pa = make_float3(bx + ax * sbs, by + ay * sbs, bz + az * sbs);
            const float3 pb = make_float3(bx + cx * sbs, by + cy * sbs, bz + cz * sbs);
            // sv[] / sd[] are indexed by a runtime corner number, i.e. they live in local memory (L1).  With r = 16 the
            // 32^3 cubes streaming through L1 evict them, and reading the corner from shared memory again is faster
            // (measured 19.0 vs 25.7 ms on 40 k voxels); for r <= 8 the local copies win (9.4 vs 14.3 ms on 200 k voxels).
            if (r > 8) {
              const float2 va = corner[(ax * r1 + ay) * r1 + az], vb2 = corner[(cx * r1 + cy) * r1 + cz];
              vp[vi] = sdf_interp(pa, pb, va.y, vb2.y, va.x, vb2.x);
            } else {
              vp[vi] = sdf_interp(pa, pb, sd[ca], sd[cb], sv[ca], sv[cb]);
            }
          }
          rr >>= 12;
          if (vp[0].w > A.max_std || vp[1].w > A.max_std || vp[2].w > A.max_std) continue;
          if (pass == 1) {
            const int t = s_base + my_off + k;
            if (t < A.max_tri) {
#pragma unroll
              for (int vi = 0; vi < 3; ++vi) {
                A.tri[(size_t)t * 9 + vi * 3 + 0] = vp[vi].x;
                A.tri[(size_t)t * 9 + vi * 3 + 1] = vp[vi].y;
                A.tri[(size_t)t * 9 + vi * 3 + 2] = vp[vi].z;
                A.tri_std[(size_t)t * 3 + vi] = vp[vi].w;
              }
              A.flat_id[t] = vb;
            }
          }
          ++k;
        }
        if (pass == 0) {
          ntri = k;
          // CTA exclusive scan of ntri, one atomic for the batch
          int inc = ntri;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
          }
          if (lane == 31) s_warp[wid] = inc;
          __syncthreads();
          if (tid == 0) {
            int run = 0;
            for (int w = 0; w < MC_T / 32; ++w) { const int t = s_warp[w]; s_warp[w] = run; run += t; }
            s_base = run > 0 ? atomicAdd(A.n_tri, run) : 0;
          }
          __syncthreads();
          my_off = s_warp[wid] + inc - ntri;
        }
      }
      __syncthreads();   // s_warp / s_base reused by the next batch
    }
  }
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_marching_cubes(const int64_t* indexer, int nx, int ny, int nz, const int64_t* valid_blocks, int U,
                                  const int32_t* vec_batch_mapping, int n_map, const float* cube_sdf, const float* cube_std,
                                  int B, int r, float max_std, int max_tri, float* tri, int64_t* flat_id, float* tri_std,
                                  int32_t* d_n_tri, void* stream) {
  DFB_CHECK_ARG(U >= 0 && r >= 1 && r <= 16 && max_tri >= 0 && d_n_tri, "marching_cubes (r must be <= 16)");
  cudaStream_t s = (cudaStream_t)stream;
  DFB_CUDA(cudaMemsetAsync(d_n_tri, 0, sizeof(int32_t), s));
  if (U == 0) return DFB_OK;
  DFB_CHECK_ARG(indexer && valid_blocks && vec_batch_mapping && cube_sdf && cube_std && (max_tri == 0 || (tri && flat_id && tri_std)),
                "marching_cubes: null pointer");
  McArgs A{indexer, nx, ny, nz, valid_blocks, U, vec_batch_mapping, n_map, cube_sdf, cube_std, B, r, max_std, max_tri,
           tri, flat_id, tri_std, d_n_tri};
  const size_t smem = sizeof(float2) * (size_t)(r + 1) * (r + 1) * (r + 1);
  DFB_CUDA(cudaFuncSetAttribute(mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks_per_sm = smem > 16384 ? 4 : 8;
  mc_kernel<<<std::min(U, sm_count() * blocks_per_sm), MC_T, smem, s>>>(A);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}
