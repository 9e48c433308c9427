// HBM-bound companions of the fused kernel:
//   K1 normalize_cast : x / (||x|| + 1e-8) -> bf16, zero-padded rows   (place_recognition.py:186-187, :169-170)
//   K3 merge_topk     : merge per-run / per-GPU candidate lists, sort, decode, apply the floor flag
//                                                                       (place_recognition.py:888-899)
//   K4 compact        : padded [Q,k] lists -> flat candidates, query asc / score desc (place_recognition.py:901-909)
//   gate_candidates   : SemanticLoopClosureGate over explicit pairs     (loop_closure_gate.py:60-126)
#include "launch.h"

#include <cuda_bf16.h>
#include <algorithm>

namespace semgate {

// =========================================================================== K1
constexpr int kNormThreads = 256;
constexpr int kNormCache = 4;   // float4 per thread kept in registers (covers d <= 4096 in one pass)

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < (blockDim.x >> 5)) ? red[l] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  return t;
}

__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* dst, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b);
  __nv_bfloat162 hi = __floats2bfloat162_rn(c, d);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(dst) = u;
}

// one block per row (grid-stride over rows); VEC: 16-byte loads are legal
template <bool VEC>
__global__ void __launch_bounds__(kNormThreads)
normalize_cast_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ld, __nv_bfloat16* __restrict__ out, int d_pad) {
  __shared__ float red[32];
  const int tid = threadIdx.x;
  for (int64_t row = blockIdx.x; row < n; row += gridDim.x) {
    const float* xr = x + row * ld;
    __nv_bfloat16* orow = out + row * static_cast<int64_t>(d_pad);
    float ss = 0.f;
    if constexpr (VEC) {
      const int nv = d >> 2;                 // float4 count (d % 4 == 0)
      float4 cache[kNormCache];
#pragma unroll
      for (int c = 0; c < kNormCache; ++c) {
        const int i = c * kNormThreads + tid;
        cache[c] = i < nv ? __ldg(reinterpret_cast<const float4*>(xr) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        ss += cache[c].x * cache[c].x + cache[c].y * cache[c].y + cache[c].z * cache[c].z + cache[c].w * cache[c].w;
      }
      for (int i = kNormCache * kNormThreads + tid; i < nv; i += kNormThreads) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xr) + i);
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
      const float denom = sqrtf(block_sum(ss, red)) + 1e-8f;
#pragma unroll
      for (int c = 0; c < kNormCache; ++c) {
        const int i = c * kNormThreads + tid;
        if (i < nv) store_bf16x4(orow + 4 * i, cache[c].x / denom, cache[c].y / denom, cache[c].z / denom, cache[c].w / denom);
      }
      for (int i = kNormCache * kNormThreads + tid; i < nv; i += kNormThreads) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xr) + i);   // second touch: L2
        store_bf16x4(orow + 4 * i, v.x / denom, v.y / denom, v.z / denom, v.w / denom);
      }
      for (int i = nv + tid; i < (d_pad >> 2); i += kNormThreads) store_bf16x4(orow + 4 * i, 0.f, 0.f, 0.f, 0.f);
    } else {
      for (int i = tid; i < d; i += kNormThreads) { const float v = xr[i]; ss += v * v; }
      const float denom = sqrtf(block_sum(ss, red)) + 1e-8f;
      for (int i = tid; i < d_pad; i += kNormThreads) orow[i] = __float2bfloat16_rn(i < d ? xr[i] / denom : 0.f);
    }
  }
}

// Long rows (d > 4096, e.g. SALAD 8448-d, AnyLoc 49152-d): one 1024-thread block per row, the row
// is staged in shared memory so that DRAM is touched once (read 4*d, write 2*d_pad).
constexpr int kNormBigThreads = 1024;

__global__ void __launch_bounds__(kNormBigThreads)
normalize_cast_smem_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ld, __nv_bfloat16* __restrict__ out, int d_pad) {
  extern __shared__ __align__(16) float row_s[];
  __shared__ float red[32];
  const int tid = threadIdx.x;
  const int nv = d >> 2;
  for (int64_t row = blockIdx.x; row < n; row += gridDim.x) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * ld);
    __nv_bfloat16* orow = out + row * static_cast<int64_t>(d_pad);
    float ss = 0.f;
    int i = tid;
    for (; i + 3 * kNormBigThreads < nv; i += 4 * kNormBigThreads) {      // 4 independent 16-byte loads in flight
      const float4 a = __ldg(xr + i), b = __ldg(xr + i + kNormBigThreads), c = __ldg(xr + i + 2 * kNormBigThreads),
                   e = __ldg(xr + i + 3 * kNormBigThreads);
      reinterpret_cast<float4*>(row_s)[i] = a;
      reinterpret_cast<float4*>(row_s)[i + kNormBigThreads] = b;
      reinterpret_cast<float4*>(row_s)[i + 2 * kNormBigThreads] = c;
      reinterpret_cast<float4*>(row_s)[i + 3 * kNormBigThreads] = e;
      ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w +
            c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w + e.x * e.x + e.y * e.y + e.z * e.z + e.w * e.w;
    }
    for (; i < nv; i += kNormBigThreads) {
      const float4 a = __ldg(xr + i);
      reinterpret_cast<float4*>(row_s)[i] = a;
      ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    const float denom = sqrtf(block_sum(ss, red)) + 1e-8f;     // block_sum's barriers also publish row_s
    for (int j = tid; j < nv; j += kNormBigThreads) {
      const float4 v = reinterpret_cast<const float4*>(row_s)[j];
      store_bf16x4(orow + 4 * j, v.x / denom, v.y / denom, v.z / denom, v.w / denom);
    }
    for (int j = nv + tid; j < (d_pad >> 2); j += kNormBigThreads) store_bf16x4(orow + 4 * j, 0.f, 0.f, 0.f, 0.f);
    __syncthreads();                                            // row_s is reused by the next row
  }
}

int launch_normalize_cast(const float* x, int64_t n, int d, int64_t ld, void* out_bf16, int d_pad, cudaStream_t st) {
  if (n <= 0) return 0;
  const bool vec = (d % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const size_t row_bytes = static_cast<size_t>(d) * 4;
  // mid-length rows (8448-d = 33 KB) do better re-reading from L2 with many small blocks per SM
  if (vec && row_bytes >= 64 * 1024 && row_bytes <= 200 * 1024) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(normalize_cast_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return static_cast<int>(e);
      attr_set = true;
    }
    const int per_sm = static_cast<int>(std::max<size_t>(1, std::min<size_t>(2, (220 * 1024) / (row_bytes + 1024))));
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(n, 148 * per_sm));
    normalize_cast_smem_kernel<<<grid, kNormBigThreads, row_bytes, st>>>(x, n, d, ld, static_cast<__nv_bfloat16*>(out_bf16), d_pad);
    return static_cast<int>(cudaGetLastError());
  }
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(n, 148 * 16));
  if (vec)
    normalize_cast_kernel<true><<<grid, kNormThreads, 0, st>>>(x, n, d, ld, static_cast<__nv_bfloat16*>(out_bf16), d_pad);
  else
    normalize_cast_kernel<false><<<grid, kNormThreads, 0, st>>>(x, n, d, ld, static_cast<__nv_bfloat16*>(out_bf16), d_pad);
  return static_cast<int>(cudaGetLastError());
}

// =========================================================================== K3
// One warp per query row.  Candidates are consumed in batches of 256 keys (8 per
// lane) next to the running list (<= 64 keys, 2 per lane); k rounds of warp-wide
// arg-max extract the new running list in descending order.  Keys are unique
// (distinct database rows), 0 = empty.
constexpr int kMergeWarps = 8;
constexpr int kBatchPerLane = 8;

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
  const uint32_t hi = static_cast<uint32_t>(v >> 32);
  const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
  const uint32_t lo = hi == mh ? static_cast<uint32_t>(v) : 0u;
  const uint32_t ml = __reduce_max_sync(0xffffffffu, lo);
  return (static_cast<uint64_t>(mh) << 32) | ml;
}

__global__ void __launch_bounds__(kMergeWarps * 32)
merge_topk_kernel(const MergeLaunch a) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kMergeWarps + (threadIdx.x >> 5);
  if (row >= a.Q) return;
  const int k = a.k;
  int n_lists = a.n_lists;
  if (n_lists < 0) n_lists = sched_slots(a.sc, static_cast<int>(row / a.rows_per_mblock));
  const uint64_t* base = a.keys_in + row * a.row_stride;
  const int total = n_lists * k;

  uint64_t run[2] = {0ull, 0ull};          // running list: rank t lives in run[t>>5] of lane t&31
  if (a.seed_keys != nullptr) {            // already sorted descending: rank t -> lane t&31, slot t>>5
    const uint64_t* seed = a.seed_keys + row * k;
    if (lane < k) run[0] = seed[lane];
    if (lane + 32 < k) run[1] = seed[lane + 32];
  }
  for (int b0 = 0; b0 < total; b0 += 32 * kBatchPerLane) {
    uint64_t c[kBatchPerLane + 2];
#pragma unroll
    for (int i = 0; i < kBatchPerLane; ++i) {
      const int e = b0 + i * 32 + lane;
      c[i] = e < total ? base[static_cast<int64_t>(e / k) * a.list_stride + (e % k)] : 0ull;
    }
    c[kBatchPerLane] = run[0];
    c[kBatchPerLane + 1] = run[1];
    uint64_t nr[2] = {0ull, 0ull};
    for (int t = 0; t < k; ++t) {
      uint64_t lm = c[0];
#pragma unroll
      for (int i = 1; i < kBatchPerLane + 2; ++i) lm = c[i] > lm ? c[i] : lm;
      const uint64_t best = warp_max_u64(lm);
      if (best == 0ull) break;                           // warp-uniform
#pragma unroll
      for (int i = 0; i < kBatchPerLane + 2; ++i) if (c[i] == best) c[i] = 0ull;
      if ((t & 31) == lane) nr[t >> 5] = best;
    }
    run[0] = nr[0];
    run[1] = nr[1];
  }

  int cnt = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int t = h * 32 + lane;
    if (t < k) {
      const uint64_t key = run[h];
      const bool got = key != 0ull;
      cnt += got ? 1 : 0;
      const int64_t o = row * k + t;
      if (a.keys_out) a.keys_out[o] = key;
      const uint32_t gi = key_index(key);
      if (a.scores) a.scores[o] = got ? key_score(key) : __int_as_float(0xff800000);
      if (a.idx) a.idx[o] = got ? static_cast<int32_t>(gi) : -1;
      if (a.valid) {
        bool ok = got;
        if (got && a.max_floor_diff >= 0 && a.q_floor != nullptr && a.db_floor != nullptr)
          ok = floor_ok(a.q_floor[row], a.db_floor[static_cast<int64_t>(gi) - a.floor_index_offset], a.max_floor_diff);
        a.valid[o] = ok ? 1 : 0;
      }
    }
  }
  if (a.count) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) a.count[row] = cnt;
  }
}

int launch_merge_topk(const MergeLaunch& a, cudaStream_t st) {
  if (a.Q <= 0) return 0;
  const unsigned grid = static_cast<unsigned>((a.Q + kMergeWarps - 1) / kMergeWarps);
  merge_topk_kernel<<<grid, kMergeWarps * 32, 0, st>>>(a);
  return static_cast<int>(cudaGetLastError());
}

// =========================================================================== K4
constexpr int kScanBlock = 1024;

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem /*>=32*/, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) smem[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = lane < (blockDim.x >> 5) ? smem[lane] : 0;
    int sinc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sinc, o); if (lane >= o) sinc += t; }
    smem[lane] = sinc - s;
    if (lane == 31 && total) *total = sinc;
  }
  __syncthreads();
  const int r = smem[w] + inc - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanBlock)
compact_count_kernel(const int32_t* __restrict__ count, int64_t Q, int64_t* __restrict__ block_sums) {
  __shared__ int sm[32];
  __shared__ int tot;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kScanBlock + threadIdx.x;
  const int v = row < Q ? count[row] : 0;
  block_exclusive_scan(v, sm, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanBlock)
compact_scan_kernel(int64_t* __restrict__ block_sums, int64_t nblocks, int64_t* __restrict__ out_total) {
  // serial-in-chunks exclusive scan; nblocks = ceil(Q/1024) is small
  __shared__ int sm[32];
  __shared__ int tot;
  int64_t carry = 0;
  for (int64_t b0 = 0; b0 < nblocks; b0 += kScanBlock) {
    const int64_t i = b0 + threadIdx.x;
    const int v = i < nblocks ? static_cast<int>(block_sums[i]) : 0;
    const int ex = block_exclusive_scan(v, sm, &tot);
    if (i < nblocks) block_sums[i] = carry + ex;
    carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_total = carry;
}

__global__ void __launch_bounds__(kScanBlock)
compact_scatter_kernel(const float* __restrict__ scores, const int32_t* __restrict__ idx, const uint8_t* __restrict__ valid,
                       const int32_t* __restrict__ count, int64_t Q, int k, const int64_t* __restrict__ block_offsets,
                       int32_t* __restrict__ out_q, int32_t* __restrict__ out_m, float* __restrict__ out_s,
                       uint8_t* __restrict__ out_v) {
  __shared__ int sm[32];
  __shared__ int offs[kScanBlock];
  __shared__ int cnts[kScanBlock];
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kScanBlock;
  const int64_t row = row0 + threadIdx.x;
  const int v = row < Q ? count[row] : 0;
  const int ex = block_exclusive_scan(v, sm, nullptr);
  offs[threadIdx.x] = ex;
  cnts[threadIdx.x] = v;
  __syncthreads();
  const int64_t base = block_offsets[blockIdx.x];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = w; r < kScanBlock && row0 + r < Q; r += (kScanBlock >> 5)) {
    const int c = cnts[r];
    const int64_t o = base + offs[r];
    const int64_t src = (row0 + r) * k;
    for (int i = lane; i < c; i += 32) {
      out_q[o + i] = static_cast<int32_t>(row0 + r);
      out_m[o + i] = idx[src + i];
      out_s[o + i] = scores[src + i];
      out_v[o + i] = valid[src + i];
    }
  }
}

size_t compact_workspace_bytes(int64_t Q) {
  return static_cast<size_t>((Q + kScanBlock - 1) / kScanBlock + 1) * sizeof(int64_t);
}

int launch_compact(const float* scores, const int32_t* idx, const uint8_t* valid, const int32_t* count, int64_t Q, int k,
                   int32_t* out_q, int32_t* out_m, float* out_s, uint8_t* out_v, int64_t* out_total, void* workspace,
                   cudaStream_t st) {
  if (Q <= 0) return static_cast<int>(cudaMemsetAsync(out_total, 0, sizeof(int64_t), st));
  const int64_t nb = (Q + kScanBlock - 1) / kScanBlock;
  int64_t* bs = static_cast<int64_t*>(workspace);
  compact_count_kernel<<<static_cast<unsigned>(nb), kScanBlock, 0, st>>>(count, Q, bs);
  compact_scan_kernel<<<1, kScanBlock, 0, st>>>(bs, nb, out_total);
  compact_scatter_kernel<<<static_cast<unsigned>(nb), kScanBlock, 0, st>>>(scores, idx, valid, count, Q, k, bs, out_q, out_m,
                                                                          out_s, out_v);
  return static_cast<int>(cudaGetLastError());
}

// =========================================================================== floor gate over pairs
__global__ void __launch_bounds__(256)
gate_candidates_kernel(const int32_t* __restrict__ floors, int64_t n_floors, const int32_t* __restrict__ q_idx,
                       const int32_t* __restrict__ m_idx, int64_t M, int max_floor_diff, uint8_t* __restrict__ out_valid,
                       unsigned long long* __restrict__ counts) {
  unsigned acc = 0, rej = 0, bad = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < M;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t q = q_idx[i], m = m_idx[i];
    bool ok = false;
    if (q < 0 || q >= n_floors || m < 0 || m >= n_floors) {
      ++bad;
    } else {
      // the gate class has no None labels and no "off" switch: diff > max -> reject (loop_closure_gate.py:89-101)
      int64_t dfl = static_cast<int64_t>(__ldg(floors + q)) - static_cast<int64_t>(__ldg(floors + m));
      if (dfl < 0) dfl = -dfl;
      ok = dfl <= max_floor_diff;
      if (ok) ++acc; else ++rej;
    }
    out_valid[i] = ok ? 1 : 0;
  }
  acc = __reduce_add_sync(0xffffffffu, acc);
  rej = __reduce_add_sync(0xffffffffu, rej);
  bad = __reduce_add_sync(0xffffffffu, bad);
  __shared__ unsigned s[3];
  if (threadIdx.x < 3) s[threadIdx.x] = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s[0], acc); atomicAdd(&s[1], rej); atomicAdd(&s[2], bad); }
  __syncthreads();
  if (threadIdx.x < 3 && s[threadIdx.x]) atomicAdd(&counts[threadIdx.x], static_cast<unsigned long long>(s[threadIdx.x]));
}

int launch_gate_candidates(const int32_t* floors, int64_t n_floors, const int32_t* q_idx, const int32_t* m_idx, int64_t M,
                           int max_floor_diff, uint8_t* out_valid, unsigned long long* counts, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(counts, 0, 3 * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (M <= 0) return 0;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>((M + 255) / 256, 148 * 8));
  gate_candidates_kernel<<<grid, 256, 0, st>>>(floors, n_floors, q_idx, m_idx, M, max_floor_diff, out_valid, counts);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace semgate
