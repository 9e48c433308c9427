// K6 — streaming query: a handful of query keyframes against the whole resident database.
//
// The online form of the path (place_recognition.py:117-163, one new keyframe per call against a
// growing database).  With <= 4 query rows the work is a GEMV: 2*N*Dpad bytes
// of bf16 database rows must cross HBM once and there is nothing for a tensor-core tile to reuse,
// so this kernel is built for bandwidth instead: every warp reads whole database rows contiguously
// (16 bytes per lane, 8 loads in flight), the query rows sit in shared memory, products accumulate
// in fp32, and the same gate / window / threshold / running top-k as the fused kernel's epilogue is
// applied per row.  The database is split by ROWS across the blocks (no 256-row tile quantisation),
// each block merges its warps' lists and writes one list per query; K3 merges the per-block lists.
#include "common.cuh"
#include "launch.h"
#include "merge.cuh"

#include <cuda_bf16.h>
#include <algorithm>

namespace semgate {

namespace {

constexpr int kSQWarps = 16;
constexpr int kSQThreads = kSQWarps * 32;
constexpr int kSQUnroll = 8;            // 16-byte loads in flight per lane

struct StreamParams {
  int Q;                 // 1..kStreamMaxQ
  int N;
  int d_pad;             // multiple of 64 -> a row is a multiple of 128 bytes
  int k;
  float threshold;
  int use_time;
  double gap;
  int max_floor_diff;
  int gate_mode;
  uint32_t db_index_offset;
  const double* q_ts;
  const double* db_ts;
  const int32_t* q_floor;
  const int32_t* db_floor;
  const __nv_bfloat16* q;
  const __nv_bfloat16* db;
  uint64_t* partial;     // [Q][gridDim.x][k]
};

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__device__ __forceinline__ float dot8(const uint4& a, const uint4& b, float acc) {
  acc = fmaf(bf16_lo(a.x), bf16_lo(b.x), acc); acc = fmaf(bf16_hi(a.x), bf16_hi(b.x), acc);
  acc = fmaf(bf16_lo(a.y), bf16_lo(b.y), acc); acc = fmaf(bf16_hi(a.y), bf16_hi(b.y), acc);
  acc = fmaf(bf16_lo(a.z), bf16_lo(b.z), acc); acc = fmaf(bf16_hi(a.z), bf16_hi(b.z), acc);
  acc = fmaf(bf16_lo(a.w), bf16_lo(b.w), acc); acc = fmaf(bf16_hi(a.w), bf16_hi(b.w), acc);
  return acc;
}

template <int NQ>
__global__ void __launch_bounds__(kSQThreads, 2)
stream_query_kernel(const StreamParams p) {
  extern __shared__ __align__(16) uint8_t sq_smem[];
  // layout: query rows bf16 [NQ][d_pad] | lists u64 [kSQWarps][NQ][k]
  uint4* qs = reinterpret_cast<uint4*>(sq_smem);
  uint64_t* lists = reinterpret_cast<uint64_t*>(sq_smem + static_cast<size_t>(NQ) * p.d_pad * 2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = p.k;
  const int vec_per_row = p.d_pad >> 3;            // uint4 (8 bf16) per row

  for (int i = threadIdx.x; i < NQ * vec_per_row; i += kSQThreads) {
    const int q = i / vec_per_row, c = i - q * vec_per_row;
    qs[i] = q < p.Q ? __ldg(reinterpret_cast<const uint4*>(p.q + static_cast<size_t>(q) * p.d_pad) + c) : make_uint4(0, 0, 0, 0);
  }
  // lane q of every warp owns the warp's list of query q
  RowList L;
  L.keys = lists + (static_cast<size_t>(warp) * NQ + (lane < NQ ? lane : 0)) * k;
  double tq = 0.0;
  int32_t qf = kFloorNone;
  const bool mask_mode = p.gate_mode == 1 && p.max_floor_diff >= 0 && p.q_floor != nullptr && p.db_floor != nullptr;
  const bool owner = lane < NQ && lane < p.Q;
  if (owner) {
    if (p.use_time) tq = p.q_ts[lane];
    if (mask_mode) qf = p.q_floor[lane];
  }
  L.reset(owner ? p.threshold : __int_as_float(0x7f800000));
  __syncthreads();

  // this block's rows, the warps take them round-robin
  const int64_t r0 = (static_cast<int64_t>(p.N) * blockIdx.x) / gridDim.x;
  const int64_t r1 = (static_cast<int64_t>(p.N) * (blockIdx.x + 1)) / gridDim.x;
  for (int64_t r = r0 + warp; r < r1; r += kSQWarps) {
    const uint4* row = reinterpret_cast<const uint4*>(p.db + r * p.d_pad);
    float acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = 0.f;
    int c = lane;
    for (; c + (kSQUnroll - 1) * 32 < vec_per_row; c += kSQUnroll * 32) {
      uint4 v[kSQUnroll];
#pragma unroll
      for (int u = 0; u < kSQUnroll; ++u) v[u] = __ldcs(row + c + u * 32);     // streamed once: evict first
#pragma unroll
      for (int u = 0; u < kSQUnroll; ++u) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[q] = dot8(v[u], qs[q * vec_per_row + c + u * 32], acc[q]);
      }
    }
    for (; c < vec_per_row; c += 32) {
      const uint4 v = __ldcs(row + c);
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc[q] = dot8(v, qs[q * vec_per_row + c], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
    }
    // lane q takes query q's score (static register indices only)
    float s = acc[0];
#pragma unroll
    for (int q = 1; q < NQ; ++q) s = lane == q ? acc[q] : s;
    if (owner && s >= L.f) {
      bool ok = true;
      if (p.use_time) ok = !time_excluded(__ldg(p.db_ts + r), tq, p.gap);
      if (ok && mask_mode) ok = floor_ok(qf, __ldg(p.db_floor + r), p.max_floor_diff);
      if (ok) L.insert(pack_key(s, static_cast<uint32_t>(r) + p.db_index_offset), k);
    }
  }
  // empty slots of the warp lists must read as 0 for the merge
  if (lane < NQ) {
    const int have = owner ? L.cnt : 0;
    for (int i = have; i < k; ++i) L.keys[i] = 0ull;
  }
  __syncthreads();
  // warp q merges the kSQWarps lists of query q and writes the block's list
  if (warp < NQ && warp < p.Q) {
    uint64_t run[2] = {0ull, 0ull};
    merge_range<8>(lists + static_cast<size_t>(warp) * k, 0, kSQWarps * k, k, static_cast<int64_t>(NQ) * k, k, lane, run, false);
    uint64_t* out = p.partial + (static_cast<size_t>(warp) * gridDim.x + blockIdx.x) * k;
    if (lane < k) out[lane] = run[0];
    if (lane + 32 < k) out[lane + 32] = run[1];
  }
}

size_t sq_smem_bytes(int nq, int d_pad, int k) {
  return static_cast<size_t>(nq) * d_pad * 2 + static_cast<size_t>(kSQWarps) * nq * k * 8;
}

int sq_grid(int sm_count) { return 2 * sm_count; }   // two co-resident blocks per SM: 128 KB of loads in flight per SM

}  // namespace

bool stream_query_fits(int64_t Q, int d_pad, int k) {
  if (Q < 1 || Q > kStreamMaxQ) return false;
  const int nq = Q <= 1 ? 1 : (Q <= 2 ? 2 : 4);
  return sq_smem_bytes(nq, d_pad, k) <= 110 * 1024;   // two blocks per SM stay resident
}

size_t stream_query_workspace_bytes(int64_t Q, int k, int sm_count) {
  return static_cast<size_t>(Q) * sq_grid(sm_count) * k * sizeof(uint64_t);
}

int stream_query_lists(int sm_count) { return sq_grid(sm_count); }

int launch_stream_query(const TopkLaunch& a, uint64_t* partial, cudaStream_t st) {
  StreamParams p{};
  p.Q = static_cast<int>(a.Q); p.N = static_cast<int>(a.N); p.d_pad = a.d_pad; p.k = a.k;
  p.threshold = a.threshold;
  p.use_time = (a.q_ts != nullptr && a.db_ts != nullptr) ? 1 : 0;
  p.gap = a.gap; p.max_floor_diff = a.max_floor_diff; p.gate_mode = a.gate_mode;
  p.db_index_offset = a.db_index_offset;
  p.q_ts = a.q_ts; p.db_ts = a.db_ts; p.q_floor = a.q_floor; p.db_floor = a.db_floor;
  p.q = static_cast<const __nv_bfloat16*>(a.q_bf16);
  p.db = static_cast<const __nv_bfloat16*>(a.db_bf16);
  p.partial = partial;
  const int nq = a.Q <= 1 ? 1 : (a.Q <= 2 ? 2 : 4);
  const size_t smem = sq_smem_bytes(nq, a.d_pad, a.k);
  const unsigned grid = static_cast<unsigned>(sq_grid(a.sm_count));
  auto launch = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kernel<<<grid, kSQThreads, smem, st>>>(p);
    return cudaGetLastError();
  };
  cudaError_t e = nq == 1 ? launch(stream_query_kernel<1>) : nq == 2 ? launch(stream_query_kernel<2>) : launch(stream_query_kernel<4>);
  return static_cast<int>(e);
}

}  // namespace semgate
