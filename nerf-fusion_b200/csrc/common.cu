// Error plumbing, device info and the shared scan primitive.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dfb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

// ---------------------------------------------------------------------------------------------
constexpr int SCAN_T = 256;
constexpr int SCAN_E = 8;
constexpr int SCAN_TILE = SCAN_T * SCAN_E;

template <bool POPC>
__device__ __forceinline__ int scan_load(const void* in, int i, int n) {
  if (i >= n) return 0;
  if (POPC) return __popc(((const uint32_t*)in)[i]);
  return ((const int*)in)[i];
}

// block-wide exclusive scan of one int per thread; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int warp_tot[SCAN_T / 32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    int t = lane < SCAN_T / 32 ? warp_tot[lane] : 0;
    int ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    if (lane < SCAN_T / 32) warp_tot[lane] = ti - t;  // exclusive warp offsets
    if (lane == SCAN_T / 32 - 1) *total = ti;
  }
  __syncthreads();
  int r = warp_tot[w] + inc - v;
  return r;
}

// n_dev (optional): device-side length, n is then only the upper bound the grid was sized for; tiles past the device
// length return at once, so a scan over a mostly unused capacity costs what its used part costs.
template <bool POPC>
__global__ void __launch_bounds__(SCAN_T) scan_tile_sums(const void* in, int n, int* block_sums, const int* __restrict__ n_dev) {
  __shared__ int tot;
  if (n_dev) n = min(n, *n_dev);
  if (blockIdx.x * SCAN_TILE >= n) return;
  int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_E;
  int s = 0;
#pragma unroll
  for (int e = 0; e < SCAN_E; ++e) s += scan_load<POPC>(in, base + e, n);
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// Second (and last) launch of a scan: every tile adds up the sums of the tiles before it on its own (at most a few
// thousand values, L2-resident) instead of waiting for a separate single-block scan of the tile sums; the tile holding the
// last element also writes the total.
template <bool POPC>
__global__ void __launch_bounds__(SCAN_T) scan_apply(const void* in, int* out, int n, const int* __restrict__ block_sums, int* __restrict__ total,
                                                     const int* __restrict__ n_dev) {
  __shared__ int tot, pre;
  if (n_dev) n = min(n, *n_dev);
  if (n <= 0) {
    if (total && blockIdx.x == 0 && threadIdx.x == 0) *total = 0;
    return;
  }
  if (blockIdx.x * SCAN_TILE >= n) return;
  int p = 0;
  for (int i = threadIdx.x; i < (int)blockIdx.x; i += SCAN_T) p += block_sums[i];
  block_excl_scan(p, &pre);
  __syncthreads();
  int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_E;
  int v[SCAN_E];
  int s = 0;
#pragma unroll
  for (int e = 0; e < SCAN_E; ++e) {
    v[e] = scan_load<POPC>(in, base + e, n);
    s += v[e];
  }
  int ex = block_excl_scan(s, &tot) + pre;
  if (total && threadIdx.x == 0 && (blockIdx.x + 1) * SCAN_TILE >= n) *total = pre + tot;
#pragma unroll
  for (int e = 0; e < SCAN_E; ++e) {
    if (base + e < n) out[base + e] = ex;
    ex += v[e];
  }
}

template <bool POPC>
static int scan_impl(const void* in, int* out, int n, int* block_sums, int* total, const int* n_dev, cudaStream_t s) {
  if (n <= 0) {
    if (total) DFB_CUDA(cudaMemsetAsync(total, 0, sizeof(int), s));
    return DFB_OK;
  }
  int nb = div_up(n, SCAN_TILE);
  scan_tile_sums<POPC><<<nb, SCAN_T, 0, s>>>(in, n, block_sums, n_dev);
  scan_apply<POPC><<<nb, SCAN_T, 0, s>>>(in, out, n, block_sums, total, n_dev);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int exclusive_scan_i32(const int* in, int* out, int n, int* block_sums, int* total, cudaStream_t s, const int* n_dev) {
  return scan_impl<false>(in, out, n, block_sums, total, n_dev, s);
}
int exclusive_scan_popc(const uint32_t* words, int* out, int n, int* block_sums, int* total, cudaStream_t s, const int* n_dev) {
  return scan_impl<true>(words, out, n, block_sums, total, n_dev, s);
}

// zero the first *n_dev (<= n_max) 32-bit words
__global__ void __launch_bounds__(256) zero_words_kernel(uint32_t* __restrict__ p, int n_max, const int* __restrict__ n_dev) {
  const int n = min(n_max, *n_dev);
  const int i = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (i + 3 < n) *reinterpret_cast<uint4*>(p + i) = make_uint4(0, 0, 0, 0);
  else for (int k = i; k < n; ++k) p[k] = 0;
}
void zero_words_dev(void* p, int n_max, const int* n_dev, cudaStream_t s) {
  zero_words_kernel<<<div_up(n_max, 1024), 256, 0, s>>>(reinterpret_cast<uint32_t*>(p), n_max, n_dev);
}

// expands the packed 29 doubles (21 upper-tri, 6, 1, 1) into the public 44-double layout
__global__ void hg_expand_kernel(const double* packed, double* out44) {
  int t = threadIdx.x;
  if (t < 36) {
    int a = t / 6, b = t % 6;
    int lo = a < b ? a : b, hi = a < b ? b : a;
    int idx = lo * 6 - lo * (lo - 1) / 2 + (hi - lo);
    out44[t] = packed[idx];
  } else if (t < 42) out44[t] = packed[21 + (t - 36)];
  else if (t == 42) out44[42] = packed[27];
  else if (t == 43) out44[43] = packed[28];
}


void launch_hg_expand(const double* packed, double* out44, cudaStream_t s) { hg_expand_kernel<<<1, 64, 0, s>>>(packed, out44); }

}  // namespace dfb

extern "C" {
int dfb_version(void) { return 100; }
const char* dfb_last_error(void) { return dfb::g_err; }
int dfb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  DFB_CUDA(cudaGetDevice(&dev));
  if (sm_count) DFB_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) DFB_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) DFB_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return DFB_OK;
}
}
