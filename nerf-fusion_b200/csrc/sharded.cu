// Spatially sharded latent voxel map (BASELINE config 5): integrate_keyframe (system/map.py:341-453) over
// the union of all ranks' points, with the voxel id space partitioned in 8^3-voxel bricks dealt round-robin to the ranks.
//
// Every exchange of the path is fused into the kernel that produces the data: a record is written STRAIGHT INTO THE OWNER'S
// RECEIVE BUFFER through a peer pointer (NVLink / NVSwitch peer stores; CUDA-IPC mappings set up once by the host side).
// Each (source, destination) pair owns a fixed-capacity segment of the destination's buffer, so the cursor a source
// advances is LOCAL (warp-aggregated atomics on its own memory) and nothing but payload and one 4-byte count per pair
// crosses the fabric; the consumer kernels (count, allocate, resolve, encoder) read the segments in place.  Phases are
// separated by a stream-ordered barrier the host side provides (a 4-byte NCCL all-reduce): no kernel ever waits for a peer.
//
//   phase 1  route_points      points -> owner of the HOME voxel (complete per-voxel counts where the prune is decided)
//   phase 2  count / prune / allocate home voxels + 6 clamped face neighbours (remote neighbours: 4-byte id requests)
//   phase 3  allocate requested ids; broadcast the candidate-set DELTAS (new voxels, voxels that crossed encoder_count_th)
//   phase 4  apply deltas to the local candidate bitmap; build (point, offset) samples exactly like map.py:390-436 and
//            push each accepted sample to the owner of its voxel
//   phase 5  resolve id -> slot, encoder on the receive buffer (encoder_tc.cu), running mean (map.py:446-452)
//
// Per rank: indexer, latents, counts for ITS bricks only (int32 indexer over owned cells); the only grid-sized state is the
// candidate BITMAP (1 bit per cell, 16 MB at 128 M cells), kept consistent by the deltas -- the focus prune of map.py:390-399
// needs the candidacy of face neighbours that live on other ranks.  Parity with the single-GPU map is defined on
// {linear voxel id -> (latent, count)}; slot numbers are per shard and follow allocation order.
#include <algorithm>

#include "common.cuh"

namespace dfb {
int tc_encoder_scatter_segments(const void* samples, const int* seg_counts, int n_seg, int seg_cap, const void* tc_blob, long long* acc,
                                cudaStream_t s);
size_t encoder_tc_blob_offset_floats();

namespace shard {

constexpr int BRICK = 8, BRICK_CELLS = BRICK * BRICK * BRICK;
constexpr int MAXW = DFB_SHARD_MAX_WORLD;

struct PointRec {   // 32 bytes
  float xn[3];
  float nrm[3];
  int home;         // linear voxel id of the home voxel
  int keep;         // set by the owner: survives the > prune_min_vox_obs test
};
struct Sample {     // 32 bytes, same record as integrate.cu / encoder_tc.cu; `slot` carries the voxel id until resolved
  int slot;
  float rel[3];
  float nrm[3];
  int pad;
};
static_assert(sizeof(PointRec) == 32 && sizeof(Sample) == 32, "record size");

// counters (device ints) ------------------------------------------------------------------------------------------------
constexpr int C_PTS = 0, C_IDS = MAXW, C_SMP = 2 * MAXW;   // per-destination cursors of the three record channels
constexpr int C_NDELTA = 3 * MAXW;                         // entries of delta_list
constexpr int C_NTOUCH = 3 * MAXW + 1;
constexpr int C_ERR = 3 * MAXW + 2;                        // bit 0: a segment overflowed, bit 1: slot capacity exceeded
constexpr int C_NOCC = 3 * MAXW + 3;                       // n_occupied
constexpr int C_NSAMPLES = 3 * MAXW + 4;                   // samples resolved this keyframe (statistics)
constexpr int C_NPOINTS = 3 * MAXW + 5;                    // points received this keyframe (statistics)
constexpr int C_NALLOC = 3 * MAXW + 6;                     // voxels allocated this keyframe
constexpr int C_NOCC0 = 3 * MAXW + 7;                      // n_occupied when the keyframe started
constexpr int EXPANDED = 0x40000000;                       // flag in grid_count: this fresh home voxel's neighbours were emitted

struct Geo {
  int nx, ny, nz, nbx, nby, nbz, world, rank, div_mode, prune_min;
  float bx, by, bz, vs, inv_vs, enc_th;
};
static Geo geo_of(const dfb_shard* S) {
  Geo g;
  g.nx = S->nx; g.ny = S->ny; g.nz = S->nz;
  g.nbx = (S->nx + BRICK - 1) / BRICK; g.nby = (S->ny + BRICK - 1) / BRICK; g.nbz = (S->nz + BRICK - 1) / BRICK;
  g.world = S->world; g.rank = S->rank; g.div_mode = S->div_mode; g.prune_min = S->prune_min_vox_obs;
  g.bx = S->bound_min[0]; g.by = S->bound_min[1]; g.bz = S->bound_min[2]; g.vs = S->voxel_size; g.inv_vs = 1.0f / S->voxel_size;
  g.enc_th = S->encoder_count_th;
  return g;
}
__device__ __forceinline__ int lin(const Geo& g, int x, int y, int z) { return z + g.nz * y + g.nz * g.ny * x; }
__device__ __forceinline__ void unlin(const Geo& g, int id, int& x, int& y, int& z) { x = id / (g.ny * g.nz); y = (id / g.nz) % g.ny; z = id % g.nz; }
__device__ __forceinline__ int brick_of(const Geo& g, int x, int y, int z) { return (z >> 3) + g.nbz * ((y >> 3) + g.nby * (x >> 3)); }
__device__ __forceinline__ int owner_of(const Geo& g, int x, int y, int z) { return brick_of(g, x, y, z) % g.world; }
// index of an OWNED cell in this rank's brick-local arrays
__device__ __forceinline__ int local_cell(const Geo& g, int x, int y, int z) {
  return (brick_of(g, x, y, z) / g.world) * BRICK_CELLS + ((x & 7) << 6) + ((y & 7) << 3) + (z & 7);
}
__device__ __forceinline__ bool cand_bit(const uint32_t* __restrict__ bits, int id) { return (bits[id >> 5] >> (id & 31)) & 1u; }

struct Channel {            // one record channel as seen by a producer
  void* dst[MAXW];          // my segment inside destination d's receive buffer (peer pointer)
  int cap;                  // records per segment
};

// Warp-aggregated append to the segment of destination `dest` (lanes with take = false pass dest = -1).  Returns the
// record index inside the segment, or -1 (not taken / overflow).
__device__ __forceinline__ int push_slot(int* cursors, int dest, bool take, int cap, int* err) {
  const unsigned active = __ballot_sync(0xffffffffu, take);
  int pos = -1;
  if (take) {
    const unsigned peers = __match_any_sync(active, dest);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    int base = 0;
    if (lane == leader) base = atomicAdd(&cursors[dest], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    pos = base + __popc(peers & ((1u << lane) - 1u));
    if (pos >= cap) { atomicOr(err, 1); pos = -1; }
  }
  return pos;
}

// ---- phase 1 -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) route_points_kernel(Geo g, const float* __restrict__ xyz, const float* __restrict__ nrm, int n,
                                                           Channel ch, int* __restrict__ counters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  PointRec r;
  int dest = -1;
  bool take = false;
  if (i < n) {
    r.xn[0] = div_vs(__fsub_rn(xyz[3 * (size_t)i], g.bx), g.vs, g.inv_vs, g.div_mode);          // map.py:367-368
    r.xn[1] = div_vs(__fsub_rn(xyz[3 * (size_t)i + 1], g.by), g.vs, g.inv_vs, g.div_mode);
    r.xn[2] = div_vs(__fsub_rn(xyz[3 * (size_t)i + 2], g.bz), g.vs, g.inv_vs, g.div_mode);
    const float cx = ceilf(r.xn[0]) - 1.f, cy = ceilf(r.xn[1]) - 1.f, cz = ceilf(r.xn[2]) - 1.f;   // :369
    if (cx >= 0.f && cx < (float)g.nx && cy >= 0.f && cy < (float)g.ny && cz >= 0.f && cz < (float)g.nz) {   // outside the grid: dropped
      take = true;
      r.home = lin(g, (int)cx, (int)cy, (int)cz);
      dest = owner_of(g, (int)cx, (int)cy, (int)cz);
      r.nrm[0] = nrm[3 * (size_t)i]; r.nrm[1] = nrm[3 * (size_t)i + 1]; r.nrm[2] = nrm[3 * (size_t)i + 2];
      r.keep = 0;
    }
  }
  const int pos = push_slot(counters + C_PTS, dest, take, ch.cap, counters + C_ERR);
  if (pos >= 0) {
    uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<PointRec*>(ch.dst[dest]) + pos);
    const uint4* s = reinterpret_cast<const uint4*>(&r);
    d[0] = s[0]; d[1] = s[1];                                   // two 16-byte (peer) stores
  }
}

// a source tells every destination how many records it wrote into its segment there (one 4-byte peer store per pair)
struct CountPtrs { int* p[MAXW]; };
__global__ void publish_counts_kernel(const int* __restrict__ cursors, int cap, CountPtrs cp, int world) {
  const int d = threadIdx.x;
  if (d < world) { *cp.p[d] = min(cursors[d], cap); }
  __threadfence_system();
}

// ---- phase 2 -----------------------------------------------------------------------------------------------------
// blockIdx.y = source segment of my point inbox
__global__ void __launch_bounds__(256) count_home_kernel(Geo g, const PointRec* __restrict__ inbox, const int* __restrict__ seg_count, int cap,
                                                         int* __restrict__ grid_count, int* __restrict__ counters) {
  const int seg = blockIdx.y, n = min(seg_count[seg], cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int x, y, z;
    unlin(g, inbox[(size_t)seg * cap + i].home, x, y, z);
    atomicAdd(&grid_count[local_cell(g, x, y, z)], 1);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(&counters[C_NPOINTS], n);
    if (seg == 0) counters[C_NOCC0] = counters[C_NOCC];        // nothing allocates between here and prune_alloc_kernel
  }
}

// allocate an owned, unallocated cell (map.py:310-319); any number of threads may race for the same cell
__device__ __forceinline__ void alloc_owned(const Geo& g, int id, int x, int y, int z, int* __restrict__ indexer, int* __restrict__ pos,
                                            int capacity, int* __restrict__ counters, int* __restrict__ delta_list, int delta_cap) {
  const int lc = local_cell(g, x, y, z);
  if (indexer[lc] != -1) return;
  if (atomicCAS(&indexer[lc], -1, -2) != -1) return;            // somebody else is allocating it
  const int slot = atomicAdd(&counters[C_NOCC], 1);
  if (slot >= capacity) { atomicOr(&counters[C_ERR], 2); return; }
  pos[slot] = id;                                               // latents / counts of unused slots are zero already
  atomicExch(&indexer[lc], slot);
  atomicAdd(&counters[C_NALLOC], 1);
  const int k = atomicAdd(&counters[C_NDELTA], 1);              // a new voxel is a candidate (obs_count 0 < encoder_count_th)
  if (k < delta_cap) delta_list[k] = id << 1; else atomicOr(&counters[C_ERR], 1);
}

__global__ void __launch_bounds__(256) prune_alloc_kernel(Geo g, PointRec* __restrict__ inbox, const int* __restrict__ seg_count, int cap,
                                                          int* __restrict__ grid_count, int* __restrict__ indexer, int* __restrict__ pos,
                                                          int capacity, int* __restrict__ counters, int* __restrict__ delta_list, int delta_cap,
                                                          Channel ids) {
  const int seg = blockIdx.y, n = min(seg_count[seg], cap);
  const int n0 = counters[C_NOCC0];
  const int iters = (n + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
  for (int it = 0; it < iters; ++it) {                          // whole warps stay in the loop: push_slot is warp-collective
    const int i = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    bool fresh = false;
    int x = 0, y = 0, z = 0, home = 0;
    if (i < n) {
      PointRec* r = inbox + (size_t)seg * cap + i;
      home = r->home;
      unlin(g, home, x, y, z);
      const int lc = local_cell(g, x, y, z);
      const bool keep = g.prune_min <= 0 || (grid_count[lc] & ~EXPANDED) > g.prune_min;   // map.py:373-379
      r->keep = keep ? 1 : 0;
      // A home voxel of a kept point that was NOT allocated when the keyframe started is allocated together with its six
      // clamped face neighbours (:382-388) -- also when a neighbour's dilation got to it first during this kernel, which
      // is why "fresh" is decided by the slot number and not by the indexer being empty.  One thread per voxel expands.
      const int v = indexer[lc];
      if (keep && !(v >= 0 && v < n0)) {
        alloc_owned(g, home, x, y, z, indexer, pos, capacity, counters, delta_list, delta_cap);
        fresh = !(atomicOr(&grid_count[lc], EXPANDED) & EXPANDED);
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      int qx = x, qy = y, qz = z;
      if (k == 0) qx = max(x - 1, 0); else if (k == 1) qx = min(x + 1, g.nx - 1);
      else if (k == 2) qy = max(y - 1, 0); else if (k == 3) qy = min(y + 1, g.ny - 1);
      else if (k == 4) qz = max(z - 1, 0); else qz = min(z + 1, g.nz - 1);
      const int own = owner_of(g, qx, qy, qz);
      const int qid = lin(g, qx, qy, qz);
      const bool remote = fresh && own != g.rank && qid != home;
      if (fresh && own == g.rank && qid != home) alloc_owned(g, qid, qx, qy, qz, indexer, pos, capacity, counters, delta_list, delta_cap);
      const int p = push_slot(counters + C_IDS, remote ? own : -1, remote, ids.cap, counters + C_ERR);
      if (p >= 0) reinterpret_cast<int*>(ids.dst[own])[p] = qid;
    }
  }
}

// ---- phase 3 -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) alloc_requests_kernel(Geo g, const int* __restrict__ inbox, const int* __restrict__ seg_count, int cap,
                                                             int* __restrict__ indexer, int* __restrict__ pos, int capacity,
                                                             int* __restrict__ counters, int* __restrict__ delta_list, int delta_cap) {
  const int seg = blockIdx.y, n = min(seg_count[seg], cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int id = inbox[(size_t)seg * cap + i];
    int x, y, z;
    unlin(g, id, x, y, z);
    alloc_owned(g, id, x, y, z, indexer, pos, capacity, counters, delta_list, delta_cap);
  }
}

// my candidate-set deltas -> my segment of EVERY rank's delta inbox (blockIdx.y = destination)
__global__ void __launch_bounds__(256) push_deltas_kernel(const int* __restrict__ delta_list, const int* __restrict__ counters, Channel ch,
                                                          CountPtrs cp) {
  const int d = blockIdx.y, n = min(counters[C_NDELTA], ch.cap);
  int* dst = reinterpret_cast<int*>(ch.dst[d]);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = delta_list[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) *cp.p[d] = n;
  __threadfence_system();
}

// ---- phase 4 -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) apply_deltas_kernel(const int* __restrict__ inbox, const int* __restrict__ seg_count, int cap,
                                                           uint32_t* __restrict__ cand_bits) {
  const int seg = blockIdx.y, n = min(seg_count[seg], cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = inbox[(size_t)seg * cap + i], id = e >> 1;
    if (e & 1) atomicAnd(&cand_bits[id >> 5], ~(1u << (id & 31)));
    else atomicOr(&cand_bits[id >> 5], 1u << (id & 31));
  }
}

__global__ void __launch_bounds__(256) emit_samples_kernel(Geo g, const PointRec* __restrict__ inbox, const int* __restrict__ seg_count, int cap,
                                                           const uint32_t* __restrict__ cand_bits, int* __restrict__ grid_count,
                                                           int* __restrict__ counters, Channel smp) {
  const int seg = blockIdx.y, n = min(seg_count[seg], cap);
  const int iters = (n + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
  for (int it = 0; it < iters; ++it) {
    const int i = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    bool focus = false;
    PointRec r;
    if (i < n) {
      r = inbox[(size_t)seg * cap + i];
      int x, y, z;
      unlin(g, r.home, x, y, z);
      grid_count[local_cell(g, x, y, z)] = 0;                    // restore the zero invariant for the next keyframe
      if (r.keep) {
        // map.py:390-399: home voxel in dilate6(candidates) <=> home or an in-range face neighbour is a candidate
        focus = cand_bit(cand_bits, r.home) || (x > 0 && cand_bit(cand_bits, lin(g, x - 1, y, z))) ||
                (x < g.nx - 1 && cand_bit(cand_bits, lin(g, x + 1, y, z))) || (y > 0 && cand_bit(cand_bits, lin(g, x, y - 1, z))) ||
                (y < g.ny - 1 && cand_bit(cand_bits, lin(g, x, y + 1, z))) || (z > 0 && cand_bit(cand_bits, lin(g, x, y, z - 1))) ||
                (z < g.nz - 1 && cand_bit(cand_bits, lin(g, x, y, z + 1)));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {                                // map.py:186-189 order
      const float ox = (k & 4) ? 0.5f : -0.5f, oy = (k & 2) ? 0.5f : -0.5f, oz = (k & 1) ? 0.5f : -0.5f;
      bool take = false;
      int dest = -1;
      Sample sm;
      if (focus) {
        const float cx = fminf(fmaxf(ceilf(__fadd_rn(r.xn[0], ox)) - 1.f, 0.f), (float)(g.nx - 1));   // :423-425
        const float cy = fminf(fmaxf(ceilf(__fadd_rn(r.xn[1], oy)) - 1.f, 0.f), (float)(g.ny - 1));
        const float cz = fminf(fmaxf(ceilf(__fadd_rn(r.xn[2], oz)) - 1.f, 0.f), (float)(g.nz - 1));
        const int id = lin(g, (int)cx, (int)cy, (int)cz);
        take = cand_bit(cand_bits, id);                          // :429
        dest = owner_of(g, (int)cx, (int)cy, (int)cz);
        sm.slot = id;
        sm.rel[0] = __fsub_rn(__fsub_rn(r.xn[0], cx), 0.5f);     // :426
        sm.rel[1] = __fsub_rn(__fsub_rn(r.xn[1], cy), 0.5f);
        sm.rel[2] = __fsub_rn(__fsub_rn(r.xn[2], cz), 0.5f);
        sm.nrm[0] = r.nrm[0]; sm.nrm[1] = r.nrm[1]; sm.nrm[2] = r.nrm[2];
        sm.pad = 0;
      }
      const int p = push_slot(counters + C_SMP, take ? dest : -1, take, smp.cap, counters + C_ERR);
      if (p >= 0) {
        uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<Sample*>(smp.dst[dest]) + p);
        const uint4* s = reinterpret_cast<const uint4*>(&sm);
        d[0] = s[0]; d[1] = s[1];
      }
    }
  }
}

// ---- phase 5 -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resolve_slots_kernel(Geo g, Sample* __restrict__ inbox, const int* __restrict__ seg_count, int cap,
                                                            const int* __restrict__ indexer, int* __restrict__ acc_n, int* __restrict__ touched,
                                                            int* __restrict__ counters) {
  const int seg = blockIdx.y, n = min(seg_count[seg], cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Sample* s = inbox + (size_t)seg * cap + i;
    int x, y, z;
    unlin(g, s->slot, x, y, z);
    const int slot = indexer[local_cell(g, x, y, z)];
    s->slot = slot;                                              // the encoder scatters by slot; < 0 cannot happen for a candidate
    if (slot >= 0 && atomicAdd(&acc_n[slot], 1) == 0) touched[atomicAdd(&counters[C_NTOUCH], 1)] = slot;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counters[C_NSAMPLES], n);
}

// map.py:450-453: one warp per touched slot; a voxel whose count reaches encoder_count_th leaves the candidate set: that
// delta opens the next keyframe's list
__global__ void __launch_bounds__(256) finalize_kernel(Geo g, int* __restrict__ counters, const int* __restrict__ touched, long long* __restrict__ acc,
                                                       int* __restrict__ acc_n, float* __restrict__ latents, float* __restrict__ obs_count,
                                                       const int* __restrict__ pos, int* __restrict__ next_delta, int* __restrict__ n_next_delta,
                                                       int delta_cap) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nt = counters[C_NTOUCH];
  for (int t = warp; t < nt; t += nwarps) {
    const int slot = touched[t];
    const float cnt = obs_count[slot];
    const float add = (float)acc_n[slot];
    const float cnt_new = __fadd_rn(cnt, add);
    __syncwarp();
    if (lane < DFB_LATENT_DIM) {
      float* lp = latents + (size_t)slot * DFB_LATENT_DIM + lane;
      long long* ap = acc + (size_t)slot * DFB_LATENT_DIM + lane;
      *lp = __fdiv_rn(__fadd_rn(acc_read(*ap), __fmul_rn(*lp, cnt)), cnt_new);
      *ap = 0;
    }
    __syncwarp();
    if (lane == 0) {
      obs_count[slot] = cnt_new; acc_n[slot] = 0;
      if (cnt < g.enc_th && !(cnt_new < g.enc_th)) {
        const int k = atomicAdd(n_next_delta, 1);
        if (k < delta_cap) next_delta[k] = (pos[slot] << 1) | 1;
      }
    }
  }
}

// end of keyframe: the removals recorded by finalize become the head of the next keyframe's delta list; cursors re-armed
__global__ void rearm_kernel(int* __restrict__ counters, int* __restrict__ delta_list, const int* __restrict__ next_delta,
                             int* __restrict__ n_next_delta, int delta_cap, int* __restrict__ stats) {
  const int n = min(*n_next_delta, delta_cap);
  for (int i = threadIdx.x; i < n; i += blockDim.x) delta_list[i] = next_delta[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    stats[0] = counters[C_NPOINTS]; stats[1] = counters[C_NSAMPLES]; stats[2] = counters[C_NALLOC]; stats[3] = counters[C_ERR];
    stats[4] = counters[C_NOCC]; stats[5] = counters[C_NTOUCH];
    for (int d = 0; d < MAXW; ++d) { stats[8 + d] = counters[C_PTS + d]; stats[8 + MAXW + d] = counters[C_SMP + d]; }
    for (int k = 0; k < 3 * MAXW; ++k) counters[k] = 0;
    counters[C_NDELTA] = n; counters[C_NTOUCH] = 0; counters[C_NSAMPLES] = 0; counters[C_NPOINTS] = 0; counters[C_NALLOC] = 0;
    *n_next_delta = 0;
  }
}

struct FlagPtrs { int* p[MAXW]; };
__global__ void peer_barrier_kernel(FlagPtrs peers, volatile int* mine, int world, int epoch, const int* __restrict__ cursors, int cap,
                                    CountPtrs cp) {
  const int d = threadIdx.x;
  if (d < world) {
    if (cursors) *cp.p[d] = min(cursors[d], cap);   // the phase's per-pair record counts travel with the barrier
    __threadfence_system();                       // this rank's earlier peer stores are ordered before the flag
    *reinterpret_cast<volatile int*>(peers.p[d]) = epoch;
    __threadfence_system();
    const long long t0 = clock64();
    while (mine[d] < epoch) {
      __nanosleep(200);
      if (clock64() - t0 > 4000000000LL) __trap();              // ~2 s: a peer is gone; fail loudly instead of hanging
    }
  }
  __threadfence_system();
}

static Channel channel(void* const* peers, int cap, int world) {
  Channel c;
  for (int d = 0; d < MAXW; ++d) c.dst[d] = d < world ? peers[d] : nullptr;
  c.cap = cap;
  return c;
}
static CountPtrs count_ptrs(int32_t* const* p, int world) {
  CountPtrs c;
  for (int d = 0; d < MAXW; ++d) c.p[d] = d < world ? p[d] : nullptr;
  return c;
}
static dim3 seg_grid(int cap, int world) { return dim3(std::max(1, std::min(2 * sm_count(), div_up(cap, 256))), world, 1); }

}  // namespace shard
}  // namespace dfb

using namespace dfb;
using namespace dfb::shard;

#define SHARD_CHECK(S)                                                                                     \
  DFB_CHECK_ARG((S) && (S)->world >= 1 && (S)->world <= MAXW && (S)->rank >= 0 && (S)->rank < (S)->world, "shard: world / rank"); \
  DFB_CHECK_ARG((long long)(S)->nx * (S)->ny * (S)->nz < (1LL << 30), "shard: grid too large for 31-bit delta entries")

extern "C" {

int dfb_shard_counter_ints(void) { return 3 * MAXW + 8; }

int64_t dfb_shard_local_cells(int nx, int ny, int nz, int world) {
  const int64_t nb = (int64_t)((nx + BRICK - 1) / BRICK) * ((ny + BRICK - 1) / BRICK) * ((nz + BRICK - 1) / BRICK);
  return ((nb + world - 1) / world) * BRICK_CELLS;
}

int dfb_shard_phase1(const dfb_shard* S, const float* xyz, const float* normal, int n, void* stream) {
  SHARD_CHECK(S);
  DFB_CHECK_ARG(n >= 0 && n <= S->pts_cap, "shard: more points than a point segment holds");
  cudaStream_t s = (cudaStream_t)stream;
  const Geo g = geo_of(S);
  if (n > 0) {
    DFB_CHECK_ARG(xyz && normal, "shard: null input");
    route_points_kernel<<<div_up(n, 256), 256, 0, s>>>(g, xyz, normal, n, channel(S->peer_pts, S->pts_cap, S->world), S->counters);
  }
  if (!S->fuse_publish) publish_counts_kernel<<<1, 32, 0, s>>>(S->counters + C_PTS, S->pts_cap, count_ptrs(S->peer_pts_count, S->world), S->world);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_shard_phase2(const dfb_shard* S, void* stream) {
  SHARD_CHECK(S);
  cudaStream_t s = (cudaStream_t)stream;
  const Geo g = geo_of(S);
  const dim3 grid = seg_grid(S->pts_cap, S->world);
  count_home_kernel<<<grid, 256, 0, s>>>(g, reinterpret_cast<const PointRec*>(S->pts_inbox), S->pts_count, S->pts_cap, S->grid_count, S->counters);
  prune_alloc_kernel<<<grid, 256, 0, s>>>(g, reinterpret_cast<PointRec*>(S->pts_inbox), S->pts_count, S->pts_cap, S->grid_count, S->indexer_local,
                                          S->latent_vecs_pos, S->capacity, S->counters, S->delta_list, S->delta_cap,
                                          channel(S->peer_ids, S->ids_cap, S->world));
  if (!S->fuse_publish) publish_counts_kernel<<<1, 32, 0, s>>>(S->counters + C_IDS, S->ids_cap, count_ptrs(S->peer_ids_count, S->world), S->world);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_shard_phase3(const dfb_shard* S, void* stream) {
  SHARD_CHECK(S);
  cudaStream_t s = (cudaStream_t)stream;
  const Geo g = geo_of(S);
  alloc_requests_kernel<<<seg_grid(S->ids_cap, S->world), 256, 0, s>>>(g, S->ids_inbox, S->ids_count, S->ids_cap, S->indexer_local, S->latent_vecs_pos,
                                                                      S->capacity, S->counters, S->delta_list, S->delta_cap);
  push_deltas_kernel<<<dim3(std::max(1, std::min(64, div_up(S->dlt_cap, 256))), S->world), 256, 0, s>>>(
      S->delta_list, S->counters, channel(S->peer_dlt, S->dlt_cap, S->world), count_ptrs(S->peer_dlt_count, S->world));
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_shard_phase4(const dfb_shard* S, void* stream) {
  SHARD_CHECK(S);
  cudaStream_t s = (cudaStream_t)stream;
  const Geo g = geo_of(S);
  apply_deltas_kernel<<<seg_grid(S->dlt_cap, S->world), 256, 0, s>>>(S->dlt_inbox, S->dlt_count, S->dlt_cap, S->cand_bits);
  emit_samples_kernel<<<seg_grid(S->pts_cap, S->world), 256, 0, s>>>(g, reinterpret_cast<const PointRec*>(S->pts_inbox), S->pts_count, S->pts_cap,
                                                                    S->cand_bits, S->grid_count, S->counters, channel(S->peer_smp, S->smp_cap, S->world));
  if (!S->fuse_publish) publish_counts_kernel<<<1, 32, 0, s>>>(S->counters + C_SMP, S->smp_cap, count_ptrs(S->peer_smp_count, S->world), S->world);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_shard_phase5(const dfb_shard* S, const float* encoder_blob, int32_t* d_stats, void* stream) {
  SHARD_CHECK(S);
  DFB_CHECK_ARG(encoder_blob && d_stats, "shard: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const Geo g = geo_of(S);
  resolve_slots_kernel<<<seg_grid(S->smp_cap, S->world), 256, 0, s>>>(g, reinterpret_cast<Sample*>(S->smp_inbox), S->smp_count, S->smp_cap,
                                                                     S->indexer_local, S->acc_n, S->touched, S->counters);
  DFB_LAUNCH_CHECK();
  int rc = tc_encoder_scatter_segments(S->smp_inbox, S->smp_count, S->world, S->smp_cap, encoder_blob + encoder_tc_blob_offset_floats(), reinterpret_cast<long long*>(S->acc), s);
  if (rc) return rc;
  finalize_kernel<<<2 * sm_count(), 256, 0, s>>>(g, S->counters, S->touched, reinterpret_cast<long long*>(S->acc), S->acc_n, S->latent_vecs, S->voxel_obs_count, S->latent_vecs_pos,
                                               S->next_delta, S->n_next_delta, S->delta_cap);
  rearm_kernel<<<1, 256, 0, s>>>(S->counters, S->delta_list, S->next_delta, S->n_next_delta, S->delta_cap, d_stats);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_shard_barrier(const dfb_shard* S, int epoch, int after_phase, void* stream) {
  SHARD_CHECK(S);
  DFB_CHECK_ARG(S->flags && epoch > 0, "shard_barrier");
  FlagPtrs fp;
  for (int d = 0; d < MAXW; ++d) fp.p[d] = d < S->world ? S->peer_flags[d] : nullptr;
  const int* cursors = nullptr;
  int cap = 0;
  CountPtrs cp = count_ptrs(S->peer_pts_count, S->world);
  if (S->fuse_publish) {
    if (after_phase == 1) { cursors = S->counters + C_PTS; cap = S->pts_cap; }
    else if (after_phase == 2) { cursors = S->counters + C_IDS; cap = S->ids_cap; cp = count_ptrs(S->peer_ids_count, S->world); }
    else if (after_phase == 4) { cursors = S->counters + C_SMP; cap = S->smp_cap; cp = count_ptrs(S->peer_smp_count, S->world); }
  }
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(fp, S->flags, S->world, epoch, cursors, cap, cp);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

// ---- peer memory: cudaMalloc'ed buffers exported / imported with CUDA IPC (one process per GPU on one node) ---------------
int dfb_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  DFB_CHECK_ARG(ptr && handle64 && bytes > 0, "peer_alloc");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DFB_CUDA(cudaMalloc(ptr, bytes));
  DFB_CUDA(cudaMemset(*ptr, 0, bytes));
  cudaIpcMemHandle_t h;
  DFB_CUDA(cudaIpcGetMemHandle(&h, *ptr));
  memcpy(handle64, &h, 64);
  return DFB_OK;
}
int dfb_peer_open(const unsigned char* handle64, void** ptr) {
  DFB_CHECK_ARG(ptr && handle64, "peer_open");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  DFB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return DFB_OK;
}
int dfb_peer_close(void* ptr) {
  if (ptr) DFB_CUDA(cudaIpcCloseMemHandle(ptr));
  return DFB_OK;
}
int dfb_peer_free(void* ptr) {
  if (ptr) DFB_CUDA(cudaFree(ptr));
  return DFB_OK;
}

}  // extern "C"
