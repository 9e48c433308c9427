// Spatial-proximity loop-closure candidate generator (radius join over poses).
//
// Replaces detect_loop_closure_candidates of the reference's integration scripts
// (orb_slam3_integration.py:167-217 and its DROID / LeGO-LOAM twins): a KD-tree ball query
// per pose (||p_i - p_j|| <= r), drop |i - j| < min_index_gap, keep i < j.  Here: an exact
// fp64 brute-force join, one warp per pose i over j in [i + gap, n), count -> scan -> fill,
// output sorted by (i, j).  Squared distances are summed in the order scipy's cKDTree uses
// for 3-d points (((dx^2) + dy^2) + dz^2, no FMA) and compared with r*r.
#include "launch.h"

#include <algorithm>

namespace semgate {

constexpr int kSpatialWarps = 8;

__device__ __forceinline__ bool within(const double* __restrict__ p, double xi, double yi, double zi, int64_t j, double r2,
                                       double* d2_out) {
  const double dx = xi - p[3 * j], dy = yi - p[3 * j + 1], dz = zi - p[3 * j + 2];
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  *d2_out = d2;
  return d2 <= r2;
}

__global__ void __launch_bounds__(kSpatialWarps * 32)
spatial_count_kernel(const double* __restrict__ pos, int64_t n, double r2, int64_t gap, int32_t* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kSpatialWarps + (threadIdx.x >> 5);
  if (i >= n) return;
  const double xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
  int c = 0;
  double d2;
  for (int64_t j = i + gap + lane; j < n; j += 32) c += within(pos, xi, yi, zi, j, r2, &d2) ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) count[i] = c;
}

__global__ void __launch_bounds__(kSpatialWarps * 32)
spatial_fill_kernel(const double* __restrict__ pos, int64_t n, double r2, int64_t gap, const int64_t* __restrict__ row_offset,
                    int32_t* __restrict__ out_i, int32_t* __restrict__ out_j, double* __restrict__ out_dist, int64_t capacity) {
  const int lane = threadIdx.x & 31;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kSpatialWarps + (threadIdx.x >> 5);
  if (i >= n) return;
  const double xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
  int64_t o = row_offset[i];
  for (int64_t j0 = i + gap; j0 < n; j0 += 32) {       // warp-uniform trip count
    const int64_t j = j0 + lane;
    double d2 = 0.0;
    const bool hit = j < n && within(pos, xi, yi, zi, j, r2, &d2);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const int64_t w = o + __popc(m & ((1u << lane) - 1u));
      if (w < capacity) {
        out_i[w] = static_cast<int32_t>(i);
        out_j[w] = static_cast<int32_t>(j);
        if (out_dist) out_dist[w] = sqrt(d2);
      }
    }
    o += __popc(m);
  }
}

// exclusive scan of int32 counts into int64 offsets: one block per 1024 rows + a serial pass over block sums
__global__ void __launch_bounds__(1024)
rows_block_scan_kernel(const int32_t* __restrict__ count, int64_t n, int64_t* __restrict__ offset, int64_t* __restrict__ block_sum) {
  __shared__ int warp_tot[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 1024 + threadIdx.x;
  const int v = i < n ? count[i] : 0;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    const int s = warp_tot[lane];
    int sinc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sinc, o); if (lane >= o) sinc += t; }
    warp_tot[lane] = sinc - s;
    if (lane == 31) block_sum[blockIdx.x] = sinc;
  }
  __syncthreads();
  if (i < n) offset[i] = warp_tot[w] + inc - v;       // block-local; the base is added below
}

__global__ void block_sum_scan_kernel(int64_t* __restrict__ block_sum, int64_t nblocks, int64_t* __restrict__ total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int64_t run = 0;
    for (int64_t b = 0; b < nblocks; ++b) { const int64_t v = block_sum[b]; block_sum[b] = run; run += v; }
    *total = run;
  }
}

__global__ void __launch_bounds__(1024)
rows_add_base_kernel(int64_t* __restrict__ offset, int64_t n, const int64_t* __restrict__ block_base) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 1024 + threadIdx.x;
  if (i < n) offset[i] += block_base[blockIdx.x];
}

size_t spatial_workspace_bytes(int64_t n) {
  const int64_t nb = (n + 1023) / 1024 + 1;
  return static_cast<size_t>(n) * (sizeof(int32_t) + sizeof(int64_t)) + static_cast<size_t>(nb) * sizeof(int64_t) + 512;
}

int launch_spatial_count(const double* pos, int64_t n, double radius, int64_t gap, void* workspace, int64_t* total,
                         cudaStream_t st) {
  if (n <= 0) return static_cast<int>(cudaMemsetAsync(total, 0, sizeof(int64_t), st));
  if (gap < 1) gap = 1;                                  // i < j
  char* ws = static_cast<char*>(workspace);
  int64_t* offset = reinterpret_cast<int64_t*>(ws);
  int64_t* bsum = offset + n;
  const int64_t nb = (n + 1023) / 1024;
  int32_t* count = reinterpret_cast<int32_t*>(bsum + nb + 1);
  const unsigned grid = static_cast<unsigned>((n + kSpatialWarps - 1) / kSpatialWarps);
  spatial_count_kernel<<<grid, kSpatialWarps * 32, 0, st>>>(pos, n, radius * radius, gap, count);
  rows_block_scan_kernel<<<static_cast<unsigned>(nb), 1024, 0, st>>>(count, n, offset, bsum);
  block_sum_scan_kernel<<<1, 32, 0, st>>>(bsum, nb, total);
  rows_add_base_kernel<<<static_cast<unsigned>(nb), 1024, 0, st>>>(offset, n, bsum);
  return static_cast<int>(cudaGetLastError());
}

int launch_spatial_fill(const double* pos, int64_t n, double radius, int64_t gap, const void* workspace, int32_t* out_i,
                        int32_t* out_j, double* out_dist, int64_t capacity, cudaStream_t st) {
  if (n <= 0) return 0;
  if (gap < 1) gap = 1;
  const int64_t* offset = reinterpret_cast<const int64_t*>(workspace);
  const unsigned grid = static_cast<unsigned>((n + kSpatialWarps - 1) / kSpatialWarps);
  spatial_fill_kernel<<<grid, kSpatialWarps * 32, 0, st>>>(pos, n, radius * radius, gap, offset, out_i, out_j, out_dist, capacity);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace semgate
