// HBM-bound companions of the fused kernel:
//   K1 normalize_cast : x / (||x|| + 1e-8) -> bf16, zero-padded rows   (place_recognition.py:186-187, :169-170)
//   K3 merge_topk     : merge per-run / per-GPU candidate lists, sort, decode, apply the floor flag
//                                                                       (place_recognition.py:888-899)
//   K4 compact        : padded [Q,k] lists -> flat candidates, query asc / score desc (place_recognition.py:901-909)
//   gate_candidates   : SemanticLoopClosureGate over explicit pairs     (loop_closure_gate.py:60-126)
#include "launch.h"
#include "merge.cuh"
#include "ptx.cuh"
#include "sortnet.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <atomic>
#include <cstdlib>

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

// Four consecutive elements of a row as fp32.  Half-precision inputs (IEEE fp16 / bf16, e.g. descriptors an extractor
// produced under autocast) are widened exactly; the element -> thread mapping and the order of the sums are the same for
// every input type, so a half-precision row gives bit for bit what its fp32 image gives.
template <bool STREAM>
__device__ __forceinline__ float4 load4(const float* row, int i) {
  return STREAM ? __ldcs(reinterpret_cast<const float4*>(row) + i) : __ldg(reinterpret_cast<const float4*>(row) + i);
}
template <bool STREAM>
__device__ __forceinline__ float4 load4(const __half* row, int i) {
  const uint2 u = STREAM ? __ldcs(reinterpret_cast<const uint2*>(row) + i) : __ldg(reinterpret_cast<const uint2*>(row) + i);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <bool STREAM>
__device__ __forceinline__ float4 load4(const __nv_bfloat16* row, int i) {
  const uint2 u = STREAM ? __ldcs(reinterpret_cast<const uint2*>(row) + i) : __ldg(reinterpret_cast<const uint2*>(row) + i);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// one block per row (grid-stride over rows); VEC: vector loads of four elements are legal
template <bool VEC, typename T>
__global__ void __launch_bounds__(kNormThreads)
normalize_cast_kernel(const T* __restrict__ x, int64_t n, int d, int64_t ld, __nv_bfloat16* __restrict__ out, int d_pad) {
  __shared__ float red[32];
  const int tid = threadIdx.x;
  for (int64_t row = blockIdx.x; row < n; row += gridDim.x) {
    const T* xr = x + row * ld;
    __nv_bfloat16* orow = out + row * static_cast<int64_t>(d_pad);
    float ss = 0.f;
    if constexpr (VEC) {
      const int nv = d >> 2;                 // float4 count (d % 4 == 0)
      float4 cache[kNormCache];
#pragma unroll
      for (int c = 0; c < kNormCache; ++c) {
        const int i = c * kNormThreads + tid;
        cache[c] = i < nv ? load4<false>(xr, i) : make_float4(0.f, 0.f, 0.f, 0.f);
        ss += cache[c].x * cache[c].x + cache[c].y * cache[c].y + cache[c].z * cache[c].z + cache[c].w * cache[c].w;
      }
      for (int i = kNormCache * kNormThreads + tid; i < nv; i += kNormThreads) {
        const float4 v = load4<false>(xr, i);
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
      const float denom = sqrtf(block_sum(ss, red)) + 1e-8f;
#pragma unroll
      for (int c = 0; c < kNormCache; ++c) {
        const int i = c * kNormThreads + tid;
        if (i < nv) store_bf16x4(orow + 4 * i, cache[c].x / denom, cache[c].y / denom, cache[c].z / denom, cache[c].w / denom);
      }
      for (int i = kNormCache * kNormThreads + tid; i < nv; i += kNormThreads) {
        const float4 v = load4<false>(xr, i);   // second touch: L2
        store_bf16x4(orow + 4 * i, v.x / denom, v.y / denom, v.z / denom, v.w / denom);
      }
      for (int i = nv + tid; i < (d_pad >> 2); i += kNormThreads) store_bf16x4(orow + 4 * i, 0.f, 0.f, 0.f, 0.f);
    } else {
      for (int i = tid; i < d; i += kNormThreads) { const float v = to_f32(xr[i]); ss += v * v; }
      const float denom = sqrtf(block_sum(ss, red)) + 1e-8f;
      for (int i = tid; i < d_pad; i += kNormThreads) orow[i] = __float2bfloat16_rn(i < d ? to_f32(xr[i]) / denom : 0.f);
    }
  }
}

// Long rows (d > 4096, e.g. SALAD 8448-d, AnyLoc 49152-d): a row is split over the CTAs of a
// thread-block cluster (<= 8), every thread keeps its 12 float4 in registers between the two
// passes, and the per-CTA partial sums of squares are exchanged through distributed shared
// memory.  DRAM is touched once (read 4*d, write 2*d_pad), several rows are in flight per SM, and
// every CTA adds the partials in rank order, so all of them divide by the same denominator.
constexpr int kNormLongCache = 12;                                   // float4 per thread
constexpr int kNormLongChunk = kNormThreads * kNormLongCache * 4;    // 12288 floats per CTA

__device__ __forceinline__ float ld_dsmem_f32(const float* local, uint32_t cta) {
  uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(local)), remote;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(cta));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kNormThreads)
normalize_cast_cluster_kernel(const T* __restrict__ x, int64_t n, int d, int64_t ld, __nv_bfloat16* __restrict__ out,
                              int d_pad, int csize) {
  __shared__ float red[32];
  __shared__ float partial[2];
  const int tid = threadIdx.x;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int64_t cluster_id = blockIdx.x / csize, n_clusters = gridDim.x / csize;
  const int nv = d >> 2;                                  // float4 per row (d % 4 == 0)
  const int v0 = static_cast<int>(rank) * (kNormLongChunk >> 2);   // this CTA's first float4 of the row
  int par = 0;
  for (int64_t row = cluster_id; row < n; row += n_clusters, par ^= 1) {
    const T* xr = x + row * ld;
    __nv_bfloat16* orow = out + row * static_cast<int64_t>(d_pad);
    float4 cache[kNormLongCache];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kNormLongCache; ++c) {
      const int i = v0 + c * kNormThreads + tid;
      cache[c] = i < nv ? load4<true>(xr, i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int c = 0; c < kNormLongCache; ++c)
      ss += cache[c].x * cache[c].x + cache[c].y * cache[c].y + cache[c].z * cache[c].z + cache[c].w * cache[c].w;
    ss = block_sum(ss, red);
    float total = ss;
    if (csize > 1) {
      if (tid == 0) partial[par] = ss;
      asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
      total = 0.f;
      for (int r = 0; r < csize; ++r) total += ld_dsmem_f32(&partial[par], static_cast<uint32_t>(r));
    }
    const float denom = sqrtf(total) + 1e-8f;
#pragma unroll
    for (int c = 0; c < kNormLongCache; ++c) {
      const int i = v0 + c * kNormThreads + tid;
      if (i < nv) store_bf16x4(orow + 4 * i, cache[c].x / denom, cache[c].y / denom, cache[c].z / denom, cache[c].w / denom);
      else if (i < (d_pad >> 2)) store_bf16x4(orow + 4 * i, 0.f, 0.f, 0.f, 0.f);
    }
  }
  // a peer may still be reading this CTA's partial sum of the last row
  if (csize > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <typename T>
static int launch_normalize_cast_t(const T* x, int64_t n, int d, int64_t ld, void* out_bf16, int d_pad, cudaStream_t st) {
  if (n <= 0) return 0;
  // four elements per vector load: 16 bytes of fp32, 8 bytes of fp16 / bf16
  const bool vec = (d % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & (4 * sizeof(T) - 1)) == 0);
  if (vec && d > 4096 && d_pad <= 8 * kNormLongChunk) {
    int csize = (d_pad + kNormLongChunk - 1) / kNormLongChunk;       // 1, 2, 3..8 -> round up to a power of two
    while (csize & (csize - 1)) ++csize;
    // one cluster per row (grid-stride only beyond 2^20 rows): residency of the clusters is the
    // hardware scheduler's business, nothing here assumes they are all co-resident
    const int64_t clusters = std::min<int64_t>(n, 1 << 20);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(clusters * csize));
    cfg.blockDim = dim3(kNormThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(csize); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return static_cast<int>(cudaLaunchKernelEx(&cfg, normalize_cast_cluster_kernel<T>, x, n, d, ld,
                                               static_cast<__nv_bfloat16*>(out_bf16), d_pad, csize));
  }
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(n, 148 * 16));
  if (vec)
    normalize_cast_kernel<true, T><<<grid, kNormThreads, 0, st>>>(x, n, d, ld, static_cast<__nv_bfloat16*>(out_bf16), d_pad);
  else
    normalize_cast_kernel<false, T><<<grid, kNormThreads, 0, st>>>(x, n, d, ld, static_cast<__nv_bfloat16*>(out_bf16), d_pad);
  return static_cast<int>(cudaGetLastError());
}

int launch_normalize_cast(const float* x, int64_t n, int d, int64_t ld, void* out_bf16, int d_pad, cudaStream_t st) {
  return launch_normalize_cast_t<float>(x, n, d, ld, out_bf16, d_pad, st);
}

// dtype: SEMGATE_DTYPE_F32 / _F16 / _BF16 (semgate.h)
int launch_normalize_cast_any(const void* x, int dtype, int64_t n, int d, int64_t ld, void* out_bf16, int d_pad, cudaStream_t st) {
  switch (dtype) {
    case 0: return launch_normalize_cast_t<float>(static_cast<const float*>(x), n, d, ld, out_bf16, d_pad, st);
    case 1: return launch_normalize_cast_t<__half>(static_cast<const __half*>(x), n, d, ld, out_bf16, d_pad, st);
    case 2: return launch_normalize_cast_t<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(x), n, d, ld, out_bf16, d_pad, st);
    default: return static_cast<int>(cudaErrorInvalidValue);
  }
}

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while the kernel before it drains; nothing that kernel wrote may be read before this returns.  A no-op otherwise.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void zero_handoff(const MergeLaunch& a) {
  if (a.zero_ptr != nullptr && blockIdx.x == 0)
    for (int i = threadIdx.x; i < a.zero_words; i += blockDim.x) a.zero_ptr[i] = 0ull;
}

template <typename Kernel, typename... Args>
static cudaError_t launch_maybe_pdl(bool pdl, Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// =========================================================================== K3
// One warp per query row.  Every lane keeps P candidate keys sorted descending in
// registers (a fixed compare-exchange network); then k rounds of "warp-wide max over
// the lanes' heads, the winning lane pops" emit the merged list in descending order.
// A batch is 32*P keys: the running list (<= 64 keys, 2 per lane) plus new ones; the
// common case (s*k <= 128 keys) is a single batch with P = 4.  Keys are unique
// (distinct database rows), 0 = empty.
constexpr int kMergeWarps = 8;
constexpr int kGatherCap = 64;     // sparse-row fast path: at most this many non-empty keys per row (two per lane)

// ROWBLOCK = false: one warp per row (8 rows per block).  ROWBLOCK = true (few rows, many lists: a
// streaming query leaves one list per block of K6): one 32-warp block per row, a three-level tree —
// every warp merges a slice of the row's keys, warps 0-3 merge eight of those lists each, warp 0
// merges the last four — so the dependent chain is 2 + 1 + 1 batches instead of one warp's dozens.
constexpr int kRowBlockWarps = 32;

template <int P, bool ROWBLOCK>
__global__ void __launch_bounds__(ROWBLOCK ? kRowBlockWarps * 32 : kMergeWarps * 32)
merge_topk_kernel(const MergeLaunch a) {
  static_assert(P == 4 || P == 8, "keys per lane");
  __shared__ uint64_t stage[ROWBLOCK ? kRowBlockWarps * kMaxK : 1];
  __shared__ uint64_t stage2[ROWBLOCK ? 4 * kMaxK : 1];
  __shared__ uint64_t gathered[ROWBLOCK ? 1 : kMergeWarps * kGatherCap];   // sparse rows: the row's non-empty keys
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = ROWBLOCK ? static_cast<int64_t>(blockIdx.x) : static_cast<int64_t>(blockIdx.x) * kMergeWarps + warp;
  griddep_wait();
  zero_handoff(a);
  // per-GPU lists read in place: OR of the peers' overflow flags (one word behind every peer's keys), so that the
  // exchange needs no collective of its own for it
  if (a.any_flag_out != nullptr && a.list_ptrs != nullptr && blockIdx.x == 0 && warp == 0) {
    uint32_t f = 0;
    for (int g = lane; g < a.n_lists; g += 32) f |= *reinterpret_cast<const volatile uint32_t*>(a.list_ptrs[g] + a.flag_offset);
    f = __reduce_or_sync(0xffffffffu, f);
    if (lane == 0) *a.any_flag_out = f;
  }
  if (a.sym_flag_copy != nullptr && a.sym_flag != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *a.sym_flag_copy = *a.sym_flag;
  if (row >= a.Q) return;
  const int k = a.k;
  int n_lists = a.n_lists;
  int64_t list_stride = a.list_stride;
  const int64_t in_row = row + a.row_offset;      // row of the input lists / query labels (a slice of the rows is merged)
  const uint64_t* base = a.keys_in + in_row * a.row_stride;
  // symmetric sweep: the row also owns a buffer of column-direction candidates, unless a buffer overflowed
  // and the full sweep redid the job (then its schedule `sc` describes the lists)
  const bool sym = a.sym_flag != nullptr && (a.sym_force || *a.sym_flag == 0u);
  if (n_lists < 0) {   // the fused kernel's partial lists: list count and offset follow the tile schedule
    const Schedule& sc = sym ? a.sc_sym : a.sc;
    const int mb = static_cast<int>(row / a.rows_per_mblock);
    n_lists = sched_slots(sc, mb);
    base = a.keys_in + sched_run_list_offset(sc, mb, sc.tab_runs != nullptr ? sc.tab_block_first[mb] : 0, row, a.rows_per_mblock, a.k);
    list_stride = sched_list_stride(sc, a.rows_per_mblock, a.k);
  }
  const int total = n_lists * k;

  uint64_t run[2] = {0ull, 0ull};
  bool have_run = false;
  if (a.seed_keys != nullptr && (!ROWBLOCK || warp == 0)) {   // one more list per row (order irrelevant here)
    const uint64_t* seed = a.seed_keys + row * k;
    if (lane < k) run[0] = seed[lane];
    if (lane + 32 < k) run[1] = seed[lane + 32];
    have_run = true;
  }
  bool done = false;
  if constexpr (!ROWBLOCK) {
    // Sparse rows first.  A run of a few tiles leaves a handful of candidates in its k slots, so a row's dozen lists
    // (and its symmetric-sweep buffer) usually hold a few dozen keys between hundreds of empty slots: gather the
    // non-empty ones (coalesced scan, ballot + prefix), and if they fit kGatherCap, rank them by counting -- every lane
    // compares its two keys with each gathered key (a shared-memory broadcast) -- and scatter by rank.  ~10 instructions
    // per 32 slots scanned plus ~8 per key, where the tournament below pays ~700 per batch of 256 slots.
    uint64_t* wg = gathered + warp * kGatherCap;
    int n = 0;
    bool fits = true;
    auto push = [&](uint64_t key) {
      const bool nz = key != 0ull;
      const uint32_t m = __ballot_sync(0xffffffffu, nz);
      const int c = __popc(m);
      if (n + c > kGatherCap) { fits = false; return; }
      if (nz) wg[n + __popc(m & ((1u << lane) - 1u))] = key;
      n += c;
    };
    if (have_run) { push(run[0]); if (fits) push(run[1]); }
    // (four loads in flight per pass: the scan is a chain of global-memory latencies otherwise)
    for (int e0 = 0; e0 < total && fits; e0 += 128) {
      uint64_t v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * 32 + lane;
        v[u] = 0ull;
        if (e < total) {
          const int g = e / k, j = e - g * k;
          v[u] = a.list_ptrs ? a.list_ptrs[g][in_row * k + j] : base[static_cast<int64_t>(g) * list_stride + j];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (fits && e0 + u * 32 < total) push(v[u]);
    }
    if (sym && fits) {
      const int extra = static_cast<int>(min(a.sym_cnt[row], static_cast<uint32_t>(a.sym_cap)));
      const uint64_t* buf = a.sym_ovf + row * a.sym_cap;
      for (int e0 = 0; e0 < extra && fits; e0 += 128) {
        uint64_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = e0 + u * 32 + lane < extra ? buf[e0 + u * 32 + lane] : 0ull;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (fits && e0 + u * 32 < extra) push(v[u]);
      }
    }
    if (fits) {
      __syncwarp();
      const uint64_t k0 = lane < n ? wg[lane] : 0ull, k1 = lane + 32 < n ? wg[lane + 32] : 0ull;
      int r0 = 0, r1 = 0;
      for (int j = 0; j < n; ++j) {
        const uint64_t x = wg[j];
        r0 += x > k0 ? 1 : 0;
        r1 += x > k1 ? 1 : 0;
      }
      __syncwarp();                                         // everybody has read the gathered keys
      if (lane < n) wg[r0] = k0;                            // keys are unique: the ranks are a permutation of 0 .. n-1
      if (lane + 32 < n) wg[r1] = k1;
      __syncwarp();
      const int keep = min(n, k);
      run[0] = lane < keep ? wg[lane] : 0ull;
      run[1] = lane + 32 < keep ? wg[lane + 32] : 0ull;
      done = true;
    }
  }
  if constexpr (!ROWBLOCK) {
    if (!done) {
      // the general case: rows with few lists take the cheaper 4-keys-per-lane network; warp-uniform choice
      if (P == 8 && total + (have_run ? 64 : 0) <= 128)
        merge_range<4>(base, 0, total, k, list_stride, k, lane, run, have_run, a.list_ptrs, in_row * k);
      else
        merge_range<P>(base, 0, total, k, list_stride, k, lane, run, have_run, a.list_ptrs, in_row * k);
      if (sym) {
        const int extra = static_cast<int>(min(a.sym_cnt[row], static_cast<uint32_t>(a.sym_cap)));
        if (extra > 0) merge_range<P>(a.sym_ovf + row * a.sym_cap, 0, extra, a.sym_cap, 0, k, lane, run, true);
      }
    }
  } else {
    const int per = (total + kRowBlockWarps - 1) / kRowBlockWarps;
    const int e0 = min(total, warp * per), e1 = min(total, e0 + per);
    merge_range<P>(base, e0, e1, k, list_stride, k, lane, run, have_run, a.list_ptrs, in_row * k);
    if (lane < k) stage[warp * k + lane] = run[0];              // k keys per warp, packed
    if (lane + 32 < k) stage[warp * k + 32 + lane] = run[1];
    __syncthreads();
    if (warp >= 4) return;
    run[0] = run[1] = 0ull;
    merge_range<P>(stage + warp * 8 * k, 0, 8 * k, k, k, k, lane, run, false);
    if (lane < k) stage2[warp * k + lane] = run[0];
    if (lane + 32 < k) stage2[warp * k + 32 + lane] = run[1];
    // only warps 0-3 are left: a named barrier over 128 threads
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp != 0) return;
    run[0] = run[1] = 0ull;
    merge_range<P>(stage2, 0, 4 * k, k, k, k, lane, run, false);
  }

  int cnt = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int t = h * 32 + lane;
    if (t < k) {
      const uint64_t key = run[h];
      const bool got = key != 0ull;
      cnt += got ? 1 : 0;
      const int64_t o = a.out_stride > 0 ? row * a.out_stride + a.out_col + t : row * k + t;
      if (a.keys_out) a.keys_out[o] = key;
      const uint32_t gi = key_index(key);
      if (a.scores) a.scores[o] = got ? key_score(key) : __int_as_float(0xff800000);
      if (a.idx) a.idx[o] = got ? static_cast<int32_t>(gi) : -1;
      if (a.valid) {
        bool ok = got;
        if (got && a.max_floor_diff >= 0 && a.q_floor != nullptr && a.db_floor != nullptr) {
          // a key from outside the label array (lists seeded from other database slices) cannot be flagged here
          const int64_t fi = static_cast<int64_t>(gi) - a.floor_index_offset;
          ok = fi >= 0 && (a.floor_n <= 0 || fi < a.floor_n) && floor_ok(a.q_floor[in_row], a.db_floor[fi], a.max_floor_diff);
        }
        a.valid[o] = ok ? 1 : 0;
      }
    }
  }
  if (a.count) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) a.count[row] = a.count_add ? a.count[row] + cnt : cnt;
  }
}

// K3, thread-per-row form on register sorting networks (k <= 32, the common case: many rows).
// The warp-per-row tournament above spends ~13 warp instructions per key (1M rows x 4 lists: 1.3 ms at boost clocks,
// 0.12 of the copy bandwidth, instruction-issue bound).  Here 32 rows share every instruction and nothing depends on a
// data-dependent index: a warp stages one list of its 32 rows in shared memory with coalesced loads (a row's list is
// k contiguous keys), every thread pulls its own row's 32 slots into registers, sorts them with Batcher's odd-even
// merge network (191 compare-exchanges, all static), folds them into its running sorted top-32 with one
// max-against-the-reverse step (the result is bitonic) and a 5-stage bitonic merge (80 compare-exchanges).  About
// 1 900 warp instructions per list for 32 rows, no shared-memory latency chain, no divergence.  (A first thread-per-row
// version kept the running list sorted in shared memory by shifting: 2.8 ms for the same problem, bound by the
// LDS -> compare -> STS chain.)  The symmetric sweep's per-keyframe candidate buffers (unsorted, up to sym_cap keys)
// are first filtered against the row's k-th key by the whole warp with coalesced loads; only the survivors, packed
// 32 at a time, go through the network.  Same inputs and outputs as merge_topk_kernel.
__device__ __forceinline__ void ce_desc(uint64_t& a, uint64_t& b) {     // afterwards a >= b
  const bool sw = a < b;
  const uint64_t hi = sw ? b : a, lo = sw ? a : b;
  a = hi; b = lo;
}

// Batcher's odd-even merge sort, 32 inputs, descending: the compare-exchange list is generated (tools/gen_sortnet.py)
// with literal indices -- a loop nest with computed bounds does not unroll and sends the keys to local memory
__device__ __forceinline__ void sort32_desc(uint64_t (&c)[32]) {
#define SEMGATE_CE(x, y) ce_desc(c[x], c[y]);
  SEMGATE_SORT32_DESC(SEMGATE_CE)
#undef SEMGATE_CE
}

// a bitonic sequence of 32 -> sorted descending
__device__ __forceinline__ void bitonic_merge32_desc(uint64_t (&c)[32]) {
#pragma unroll
  for (int stride = 16; stride >= 1; stride >>= 1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if ((i & stride) == 0) ce_desc(c[i], c[i | stride]);
    }
  }
}

template <int W>
__device__ __forceinline__ void bitonic_merge_desc(uint64_t (&c)[32]) {     // the first W slots
#pragma unroll
  for (int stride = W / 2; stride >= 1; stride >>= 1) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
      if ((i & stride) == 0) ce_desc(c[i], c[i | stride]);
    }
  }
}

// this thread's packed row (its first `cnt` of 32 slots in shared memory) -> registers -> sorting network -> folded
// into its running top-32
// Lists that are already sorted -- the per-GPU lists of a sharded sweep (each is a K3 output), a seeded list -- skip the
// sorting network (191 of the ~300 compare-exchanges per list): 31 compares per row and a warp vote decide, so any
// input is still handled.  `have_best` (warp-uniform): the first fold of a row is a copy.
__device__ __forceinline__ void fold_row(const uint64_t* __restrict__ mine, int cnt, uint64_t (&best)[32], bool& have_best) {
  uint64_t cur[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) cur[i] = i < cnt ? mine[i] : 0ull;
  bool sorted = true;
#pragma unroll
  for (int i = 0; i < 31; ++i) sorted &= cur[i] >= cur[i + 1];
  if (!__all_sync(0xffffffffu, sorted)) sort32_desc(cur);
  if (!have_best) {
#pragma unroll
    for (int i = 0; i < 32; ++i) best[i] = cur[i];
    have_best = true;
    return;
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) best[i] = best[i] > cur[31 - i] ? best[i] : cur[31 - i];    // best desc, cur reversed: bitonic
  bitonic_merge32_desc(best);
}

constexpr int kNetRows = 128;      // rows (= threads) per block
constexpr int kNetPitch = 33;      // shared-memory row pitch in keys (odd: conflict-free 64-bit accesses)

// Candidate keys reach a row from several sources: its partial lists (k slots each, mostly sparse: a run of a few tiles
// leaves a handful of candidates), a seeded list, the symmetric sweep's buffer.  Each source is staged raw by the warp
// (coalesced), every thread appends its row's non-empty keys to its pack row, and the network only runs when some row
// of the warp could not take another source: the number of network passes follows the keys, not the sources.
__global__ void __launch_bounds__(kNetRows)
merge_net_kernel(const MergeLaunch a) {
  extern __shared__ uint64_t net_smem[];
  uint64_t* raw = net_smem;                                            // [kNetRows][kNetPitch] staging
  uint64_t* pack = net_smem + static_cast<size_t>(kNetRows) * kNetPitch;   // [kNetRows][kNetPitch] packed keys per row
  const int k = a.k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  griddep_wait();
  zero_handoff(a);
  if (a.any_flag_out != nullptr && a.list_ptrs != nullptr && blockIdx.x == 0 && warp == 0) {
    uint32_t f = 0;
    for (int g = lane; g < a.n_lists; g += 32) f |= *reinterpret_cast<const volatile uint32_t*>(a.list_ptrs[g] + a.flag_offset);
    f = __reduce_or_sync(0xffffffffu, f);
    if (lane == 0) *a.any_flag_out = f;
  }
  if (a.sym_flag_copy != nullptr && a.sym_flag != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *a.sym_flag_copy = *a.sym_flag;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kNetRows;
  const int64_t wrow0 = row0 + warp * 32;                  // first row of this warp
  if (wrow0 >= a.Q) return;
  const int64_t in_row0 = row0 + a.row_offset;
  const bool sym = a.sym_flag != nullptr && (a.sym_force || *a.sym_flag == 0u);

  // where the lists of this block's rows live (uniform over the block: kNetRows divides the rows of an m-block)
  int n_lists = a.n_lists;
  const uint64_t* base0 = a.keys_in + in_row0 * a.row_stride;   // list 0 of row0
  int64_t row_pitch = a.row_stride, list_stride = a.list_stride;
  if (n_lists < 0) {
    const Schedule& sc = sym ? a.sc_sym : a.sc;
    const int mb = static_cast<int>(row0 / a.rows_per_mblock);
    n_lists = sched_slots(sc, mb);
    if (sc.tab_runs != nullptr) {
      base0 = a.keys_in + sched_run_list_offset(sc, mb, sc.tab_block_first[mb], row0, a.rows_per_mblock, k);
      row_pitch = k;
      list_stride = static_cast<int64_t>(a.rows_per_mblock) * k;
    } else {
      const int64_t rows_full = static_cast<int64_t>(sc.n_full) * sc.rm * a.rows_per_mblock;
      base0 = a.keys_in + sched_list_offset(sc, row0, a.rows_per_mblock, k);
      row_pitch = static_cast<int64_t>(row0 < rows_full ? sc.s_main : sc.s_last) * k;
      list_stride = k;
    }
  }

  const int rows_here = static_cast<int>(min(static_cast<int64_t>(32), a.Q - wrow0));
  uint64_t* wraw = raw + static_cast<size_t>(warp) * 32 * kNetPitch;
  const uint64_t* raw_mine = wraw + static_cast<size_t>(lane) * kNetPitch;
  uint64_t* mine = pack + (static_cast<size_t>(warp) * 32 + lane) * kNetPitch;
  uint64_t best[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) best[i] = 0ull;
  int cnt = 0;                                              // keys in this thread's pack row
  bool have_best = false;

  // lists: g = -1 is the seeded list (one more list per row, local rows), then the n_lists lists; each is k keys per
  // row, `pitch` keys between rows
  for (int g = a.seed_keys != nullptr ? -1 : 0; g < n_lists; ++g) {
    const uint64_t* src_row0;
    int64_t pitch = k;
    if (g < 0) src_row0 = a.seed_keys + row0 * k;
    else if (a.list_ptrs != nullptr) src_row0 = a.list_ptrs[g] + in_row0 * k;
    else { src_row0 = base0 + g * list_stride; pitch = row_pitch; }
    src_row0 += static_cast<int64_t>(warp) * 32 * pitch + lane;
    uint64_t v[32];
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) v[rr] = (rr < rows_here && lane < k) ? __ldg(src_row0 + rr * pitch) : 0ull;   // 32 loads in flight
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) wraw[rr * kNetPitch + lane] = v[rr];
    __syncwarp();
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      if (j < k) {
        const uint64_t key = raw_mine[j];
        if (key != 0ull) mine[cnt++] = key;
      }
    }
    __syncwarp();                                           // the staging rows may be refilled
    if (__any_sync(0xffffffffu, cnt > 32 - k)) {            // some row could not take another list
      fold_row(mine, cnt, best, have_best);
      cnt = 0;
    }
  }

  if (sym) {
    // Column-direction candidates of the symmetric sweep: keyframe r owns min(sym_cnt[r], sym_cap) unsorted keys, 16 at
    // a time here (two rows per load instruction).  What cannot reach the top 32 any more is dropped before packing.
    const int64_t row = wrow0 + lane;
    const int extra = lane < rows_here ? static_cast<int>(min(a.sym_cnt[row], static_cast<uint32_t>(a.sym_cap))) : 0;
    int chunks = (extra + 15) >> 4;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) chunks = max(chunks, __shfl_xor_sync(0xffffffffu, chunks, o));
    const int half = lane >> 4, l16 = lane & 15;
    for (int c = 0; c < chunks; ++c) {
      if (__any_sync(0xffffffffu, cnt > 16)) {              // room for 16 more in every row?
        fold_row(mine, cnt, best, have_best);
        cnt = 0;
      }
      uint64_t v[16];
#pragma unroll
      for (int r2 = 0; r2 < 16; ++r2) {
        const int rr = 2 * r2 + half;
        const int ex = __shfl_sync(0xffffffffu, extra, rr);
        const int e = c * 16 + l16;
        v[r2] = e < ex ? __ldg(a.sym_ovf + (wrow0 + rr) * a.sym_cap + e) : 0ull;
      }
#pragma unroll
      for (int r2 = 0; r2 < 16; ++r2) wraw[(2 * r2 + half) * kNetPitch + l16] = v[r2];
      __syncwarp();
      const uint64_t bar = best[31];                        // literal index on purpose (see fold_row): keys below the
#pragma unroll                                              // row's 32nd best so far cannot reach its top k <= 32
      for (int j = 0; j < 16; ++j) {
        const uint64_t key = raw_mine[j];
        if (key > bar) mine[cnt++] = key;
      }
      __syncwarp();
    }
  }
  if (__any_sync(0xffffffffu, cnt > 0)) fold_row(mine, cnt, best, have_best);

  // outputs: registers -> staging rows -> one row at a time by the whole warp (coalesced)
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 32; ++i) wraw[lane * kNetPitch + i] = best[i];
  __syncwarp();
  const bool gate = a.valid != nullptr && a.max_floor_diff >= 0 && a.q_floor != nullptr && a.db_floor != nullptr;
  for (int rr = 0; rr < rows_here; ++rr) {
    const int64_t orow = wrow0 + rr;
    const uint64_t key = lane < k ? wraw[rr * kNetPitch + lane] : 0ull;
    const bool got = key != 0ull;
    const int c = __popc(__ballot_sync(0xffffffffu, got));
    if (lane < k) {
      const int64_t o = a.out_stride > 0 ? orow * a.out_stride + a.out_col + lane : orow * k + lane;
      if (a.keys_out) a.keys_out[o] = key;
      const uint32_t gi = key_index(key);
      if (a.scores) a.scores[o] = got ? key_score(key) : __int_as_float(0xff800000);
      if (a.idx) a.idx[o] = got ? static_cast<int32_t>(gi) : -1;
      if (a.valid) {
        bool ok = got;
        if (got && gate) {
          // a key from outside the label array (lists seeded from other database slices) cannot be flagged here
          const int64_t fi = static_cast<int64_t>(gi) - a.floor_index_offset;
          ok = fi >= 0 && (a.floor_n <= 0 || fi < a.floor_n) &&
               floor_ok(__ldg(a.q_floor + in_row0 + warp * 32 + rr), __ldg(a.db_floor + fi), a.max_floor_diff);
        }
        a.valid[o] = ok ? 1 : 0;
      }
    }
    if (a.count && lane == 0) a.count[orow] = a.count_add ? a.count[orow] + c : c;
  }
}

// K3, thread-per-row form for DENSE lists given explicitly (n_lists >= 0: the per-GPU lists of a sharded sweep, read from
// one [G,Q,k] array or in place over NVLink through a pointer table, plus a seeded list).  Such lists are K3 outputs:
// sorted, full or nearly full, rows contiguous.  The network kernel above spends most of its ~340 warp instructions per
// row on moving keys (32 loads + 32 stores to stage a list, a data-dependent packing pass, a reload): here
//  * a warp's 32 rows of one list are ONE contiguous block of 32 k keys: a single cp.async.bulk (TMA, one instruction,
//    mbarrier completion) lands it in shared memory, two lists in flight per warp; rows are k keys apart, so with odd k
//    (25) every thread reads its own row conflict-free; even k, a ragged last warp or an unaligned source take plain
//    coalesced loads into rows pitched k | 1;
//  * no packing: a list goes from shared memory straight into registers; sorted lists (31 compares and a warp vote;
//    anything else runs the sorting network, so every input is still handled) fold into the running list with one
//    max-against-the-reverse step and the bitonic merge; with k known at compile time (KT = 25) the slots beyond k are
//    literal zeros and the compiler drops every compare-exchange that only feeds them;
//  * outputs leave as the contiguous block they are: element e of the warp's 32 k outputs is row e / k, column e % k.
constexpr int kDenseWarps = 4;
// rows k keys apart are read by 32 threads at once, 8 bytes each: conflict-free for odd k, two-way for k = 2 mod 4 (kept:
// a bulk copy needs rows unpitched); other k get an odd pitch and plain loads
__host__ __device__ constexpr int dense_pitch(int k) { return (k & 1) || (k & 3) == 2 ? k : k | 1; }

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int KT>
__global__ void __launch_bounds__(kDenseWarps * 32, 3)
merge_dense_kernel(const MergeLaunch a) {
  using namespace ptx;
  constexpr int KK = KT > 0 ? KT : 32;                       // slots that can hold a key
  constexpr int W = KK <= 8 ? 8 : KK <= 16 ? 16 : 32;        // width of the fold
  extern __shared__ __align__(128) uint8_t dense_smem[];
  const int k = KT > 0 ? KT : a.k;
  const int pitch = dense_pitch(k);                          // keys between rows in shared memory
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t stage_bytes = 32u * pitch * 8u;
  uint64_t* stage0 = reinterpret_cast<uint64_t*>(dense_smem + 128 + static_cast<size_t>(warp) * 2 * stage_bytes);
  const uint32_t bar0 = smem_u32(dense_smem) + warp * 16;    // two mbarriers per warp
  griddep_wait();
  zero_handoff(a);
  if (a.any_flag_out != nullptr && a.list_ptrs != nullptr && blockIdx.x == 0 && warp == 0) {
    uint32_t f = 0;
    for (int g = lane; g < a.n_lists; g += 32) f |= *reinterpret_cast<const volatile uint32_t*>(a.list_ptrs[g] + a.flag_offset);
    f = __reduce_or_sync(0xffffffffu, f);
    if (lane == 0) *a.any_flag_out = f;
  }
  const int64_t wrow0 = static_cast<int64_t>(blockIdx.x) * (kDenseWarps * 32) + warp * 32;    // first row of this warp
  if (wrow0 >= a.Q) return;                                  // nothing below synchronises the block
  const int64_t in_wrow0 = wrow0 + a.row_offset;
  const int rows_here = static_cast<int>(min(static_cast<int64_t>(32), a.Q - wrow0));
  const int n_keys = rows_here * k;
  if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); fence_barrier_init(); }
  __syncwarp();
  // element e = lane + 32 t of the warp's block of rows_here x k keys is (row, column) = (e / k, e % k): one division
  // here, then steps of 32
  const int step_r = 32 / k, step_c = 32 - step_r * k;
  const int r_first = lane / k, c_first = lane - r_first * k;

  const int g0 = a.seed_keys != nullptr ? -1 : 0;            // source -1 is the seeded list (local rows)
  const int n_src = a.n_lists - g0;
  auto source = [&](int s) -> const uint64_t* {
    const int g = g0 + s;
    if (g < 0) return a.seed_keys + wrow0 * k;
    if (a.list_ptrs != nullptr) return a.list_ptrs[g] + in_wrow0 * k;
    return a.keys_in + g * a.list_stride + in_wrow0 * k;
  };
  // the warp's block of source s -> stage s & 1: one bulk copy, or coalesced loads into pitched rows
  auto fetch = [&](int s) {
    const uint64_t* src = source(s);
    uint64_t* dst = stage0 + static_cast<size_t>(s & 1) * (stage_bytes / 8);
    const bool bulk = pitch == k && rows_here == 32 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;     // warp-uniform
    if (bulk) {
      if (lane == 0) {
        const uint32_t bar = bar0 + (s & 1) * 8;
        fence_proxy_async();                                 // the stage may hold keys written by plain stores
        mbar_arrive_expect_tx(bar, stage_bytes);
        bulk_load(smem_u32(dst), src, stage_bytes, bar);
      }
    } else {
      int r = r_first, c = c_first;
      for (int e = lane; e < n_keys; e += 32) {
        dst[r * pitch + c] = __ldg(src + e);
        r += step_r; c += step_c;
        if (c >= k) { c -= k; ++r; }
      }
    }
    return bulk;
  };

  uint64_t best[32];
  uint32_t bulk_mask = 0;                                    // bit s & 1: the stage was filled by a bulk copy
  uint32_t phase = 0;                                        // bit s & 1: parity of that stage's barrier
  for (int s = 0; s < n_src && s < 2; ++s) bulk_mask |= fetch(s) ? 1u << s : 0u;
  for (int s = 0; s < n_src; ++s) {
    const int st = s & 1;
    if (bulk_mask >> st & 1) { mbar_wait(bar0 + st * 8, phase >> st & 1); phase ^= 1u << st; }
    else __syncwarp();
    const uint64_t* mine = stage0 + static_cast<size_t>(st) * (stage_bytes / 8) + lane * pitch;
    uint64_t cur[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) cur[i] = (i < KK && i < k && lane < rows_here) ? mine[i] : 0ull;
    __syncwarp();                                            // every lane has read its row: the stage may be refilled
    if (s + 2 < n_src) bulk_mask = (bulk_mask & ~(1u << st)) | (fetch(s + 2) ? 1u << st : 0u);
    bool sorted = true;
#pragma unroll
    for (int i = 0; i + 1 < KK; ++i) sorted &= cur[i] >= cur[i + 1];
    if (!__all_sync(0xffffffffu, sorted)) sort32_desc(cur);
#pragma unroll
    for (int i = KK; i < 32; ++i) cur[i] = 0ull;             // zeros sort last: literal again after the network
    if (s == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) best[i] = cur[i];
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) best[i] = best[i] > cur[W - 1 - i] ? best[i] : cur[W - 1 - i];   // bitonic
      bitonic_merge_desc<W>(best);
#pragma unroll
      for (int i = KK; i < 32; ++i) best[i] = 0ull;          // only the top k go on
    }
  }
  if (n_src <= 0) {
#pragma unroll
    for (int i = 0; i < 32; ++i) best[i] = 0ull;
  }

  // outputs: registers -> the warp's first stage (all fetches have been consumed) -> one contiguous block per array
  __syncwarp();
  uint64_t* outk = stage0;
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < KK; ++i) {
    if (i < k) { outk[lane * pitch + i] = best[i]; cnt += best[i] != 0ull ? 1 : 0; }
  }
  __syncwarp();
  if (a.count && lane < rows_here) a.count[wrow0 + lane] = a.count_add ? a.count[wrow0 + lane] + cnt : cnt;
  const bool gate = a.valid != nullptr && a.max_floor_diff >= 0 && a.q_floor != nullptr && a.db_floor != nullptr;
  if constexpr (KT > 0) {
    // the common case -- a full warp of rows, plain [Q,k] outputs, k known at compile time -- unrolled: slot e = 32 j + lane of
    // the warp's block sits at a literal offset of every array, its row is a division by a constant
    if (rows_here == 32 && a.out_stride <= 0 && pitch == KT) {
      const int64_t o0 = wrow0 * KT + lane;
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        const uint64_t key = outk[32 * j + lane];
        const bool got = key != 0ull;
        const int64_t o = o0 + 32 * j;
        if (a.keys_out) a.keys_out[o] = key;
        const uint32_t gi = key_index(key);
        if (a.scores) a.scores[o] = got ? key_score(key) : __int_as_float(0xff800000);
        if (a.idx) a.idx[o] = got ? static_cast<int32_t>(gi) : -1;
        if (a.valid) {
          bool ok = got;
          if (got && gate) {
            const int64_t fi = static_cast<int64_t>(gi) - a.floor_index_offset;
            ok = fi >= 0 && (a.floor_n <= 0 || fi < a.floor_n) &&
                 floor_ok(__ldg(a.q_floor + in_wrow0 + (32 * j + lane) / KT), __ldg(a.db_floor + fi), a.max_floor_diff);
          }
          a.valid[o] = ok ? 1 : 0;
        }
      }
      return;
    }
  }
  int r = r_first, c = c_first;
  for (int e = lane; e < n_keys; e += 32) {
    const uint64_t key = outk[r * pitch + c];
    const bool got = key != 0ull;
    const int64_t o = a.out_stride > 0 ? (wrow0 + r) * a.out_stride + a.out_col + c : wrow0 * k + e;
    if (a.keys_out) a.keys_out[o] = key;
    const uint32_t gi = key_index(key);
    if (a.scores) a.scores[o] = got ? key_score(key) : __int_as_float(0xff800000);
    if (a.idx) a.idx[o] = got ? static_cast<int32_t>(gi) : -1;
    if (a.valid) {
      bool ok = got;
      if (got && gate) {
        // a key from outside the label array (lists seeded from other database slices) cannot be flagged here
        const int64_t fi = static_cast<int64_t>(gi) - a.floor_index_offset;
        ok = fi >= 0 && (a.floor_n <= 0 || fi < a.floor_n) && floor_ok(__ldg(a.q_floor + in_wrow0 + r), __ldg(a.db_floor + fi), a.max_floor_diff);
      }
      a.valid[o] = ok ? 1 : 0;
    }
    r += step_r; c += step_c;
    if (c >= k) { c -= k; ++r; }
  }
}

template <int KT>
static int launch_merge_dense(const MergeLaunch& a, cudaStream_t st) {
  const unsigned grid = static_cast<unsigned>((a.Q + kDenseWarps * 32 - 1) / (kDenseWarps * 32));
  const size_t smem = 128 + static_cast<size_t>(kDenseWarps) * 2 * 32 * dense_pitch(a.k) * sizeof(uint64_t);   // k = 25: 51.3 KB
  cudaError_t e = cudaFuncSetAttribute(merge_dense_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  return static_cast<int>(launch_maybe_pdl(a.pdl != 0, merge_dense_kernel<KT>, dim3(grid), dim3(kDenseWarps * 32), smem, st, a));
}

// process-wide switch (semgate_set_option "k3_dense"; first read: SEMGATE_K3_DENSE): A/B runs and tests of the general kernel
static std::atomic<int> g_merge_dense{-1};
void set_merge_dense(int on) { g_merge_dense.store(on != 0 ? 1 : 0); }
static bool merge_dense_enabled() {
  int v = g_merge_dense.load();
  if (v < 0) { const char* e = getenv("SEMGATE_K3_DENSE"); v = (e == nullptr || atoi(e) != 0) ? 1 : 0; g_merge_dense.store(v); }
  return v != 0;
}

int launch_merge_topk(const MergeLaunch& a, cudaStream_t st) {
  if (a.Q <= 0 && a.any_flag_out == nullptr) return 0;
  int lists = a.n_lists >= 0 ? a.n_lists : std::max(a.sc.s_main, a.sc.s_last);
  if (a.n_lists < 0 && a.sym_flag != nullptr) lists = std::max(lists, std::max(a.sc_sym.s_max, std::max(a.sc_sym.s_main, a.sc_sym.s_last)));
  const int64_t keys = static_cast<int64_t>(lists) * a.k;
  // few rows, many lists (a streaming query leaves one list per block of K6): one block per row, a tree of warps
  if (a.Q > 0 && a.Q <= 2048 && keys >= 1024 && a.sym_flag == nullptr && a.row_offset == 0) {
    return static_cast<int>(launch_maybe_pdl(a.pdl != 0, merge_topk_kernel<8, true>, dim3(static_cast<unsigned>(a.Q)), dim3(kRowBlockWarps * 32), 0, st, a));
  }
  // many rows, a few (possibly full) lists each, e.g. the per-GPU lists of a sharded 1M sweep: thread-per-row on register
  // sorting networks (1M x 4 full lists: 0.98 ms against the tournament's 1.3-1.9 ms); everything else -- in particular
  // the sparse lists of run-table sweeps -- goes to the warp-per-row kernel and its gather-and-rank fast path
  if (a.k <= 32 && a.Q >= 65536 && lists <= 8 && a.sym_flag == nullptr) {
    // lists handed over as arrays (per-GPU lists, a seeded list) are K3 outputs -- sorted, mostly full, rows k keys apart:
    // the dense form (1M x 4 sorted lists: 0.65 ms in the general network kernel); SEMGATE_K3_DENSE=0 keeps the latter
    if (merge_dense_enabled() && a.n_lists >= 1 && (a.list_ptrs != nullptr || a.row_stride == a.k))
      return a.k == 25 ? launch_merge_dense<25>(a, st) : a.k == 10 ? launch_merge_dense<10>(a, st) : a.k == 5 ? launch_merge_dense<5>(a, st) : launch_merge_dense<0>(a, st);
    const unsigned grid = static_cast<unsigned>(std::max<int64_t>(1, (a.Q + kNetRows - 1) / kNetRows));
    const size_t smem = 2 * static_cast<size_t>(kNetRows) * kNetPitch * sizeof(uint64_t);      // 67.6 KB: staging + pack rows
    cudaError_t e = cudaFuncSetAttribute(merge_net_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    return static_cast<int>(launch_maybe_pdl(a.pdl != 0, merge_net_kernel, dim3(grid), dim3(kNetRows), smem, st, a));
  }
  const unsigned grid = static_cast<unsigned>(std::max<int64_t>(1, (a.Q + kMergeWarps - 1) / kMergeWarps));
  if (keys <= (a.seed_keys ? 64 : 128))
    return static_cast<int>(launch_maybe_pdl(a.pdl != 0, merge_topk_kernel<4, false>, dim3(grid), dim3(kMergeWarps * 32), 0, st, a));
  return static_cast<int>(launch_maybe_pdl(a.pdl != 0, merge_topk_kernel<8, false>, dim3(grid), dim3(kMergeWarps * 32), 0, st, a));
}

// =========================================================================== K4
constexpr int kScanBlock = 128;     // rows per block of the count / scatter kernels (many small blocks: all SMs busy)
constexpr int kCompactMaxThreads = 512;   // K4: threads per block (a multiple of kScanBlock; the extra ones only scatter)

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

// entries of a row that are emitted: all `count[row]` of them, or only the floor-consistent ones
// (valid_only: the list handed to geometric verification, which skips cross-floor pairs anyway,
// geometric_verification.py:709)
__device__ __forceinline__ int row_emit_count(const int32_t* __restrict__ count, const uint8_t* __restrict__ valid, int64_t row,
                                              int k, bool valid_only) {
  const int c = count[row];
  if (!valid_only) return c;
  int n = 0;
  for (int i = 0; i < c; ++i) n += valid[row * k + i] ? 1 : 0;
  return n;
}

// One pass: a block takes a ticket (its tile of kScanBlock rows), scans its rows' counts, publishes the tile's total,
// looks back over the tiles before it for its offset ("decoupled look-back": a tile publishes first its own total,
// then, once known, the inclusive prefix; a later tile adds totals backwards until it meets an inclusive prefix) and
// scatters.  Tickets make the order of tiles the order in which blocks started, so every tile a block waits for is
// already running or done.  One launch where count -> scan -> scatter were three: small sweeps are launch-bound.
// state: [tiles] 64-bit words = value << 2 | flag (0 nothing yet, 1 tile total, 2 inclusive prefix), then the ticket
// counter; zeroed by the caller per launch.
// (relaxed, gpu scope: a state word carries its whole message -- value and flag in one 64-bit word -- and guards no other
// data, so neither side needs a fence: 1M x 25 full lists 0.205 -> 0.193 ms, 1M rows with one candidate each 0.082 ->
// 0.067 ms.  Measured and dropped on the same box (tools/k4_bench.py): 256 tiles per look-back round instead of 32
// (0.226 / 0.091 ms), one thread per INPUT slot with 4 or 8 loads issued before their stores and the first batch in flight
// during the look-back (0.197 - 0.27 ms): the search-per-output form below stayed the fastest at every density.)
__device__ __forceinline__ unsigned long long ld_state_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_state_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

#ifdef SEMGATE_K4_TRACE   // tools/probe/k4_trace.cu: per-tile stage stamps (globaltimer ns), rounds and spins of the look-back
__device__ unsigned long long k4_trace[8 * 65536];
#define K4_STAMP(slot) do { if (threadIdx.x == 0 && tile < 65536u) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); k4_trace[8 * tile + (slot)] = t_; } } while (0)
#define K4_NOTE(slot, v) do { if (tile < 65536u) k4_trace[8 * tile + (slot)] = (v); } while (0)
#else
#define K4_STAMP(slot) do { } while (0)
#define K4_NOTE(slot, v) do { } while (0)
#endif
__global__ void __launch_bounds__(kCompactMaxThreads)
compact_onepass_kernel(const float* __restrict__ scores, const int32_t* __restrict__ idx, const uint8_t* __restrict__ valid,
                       const int32_t* __restrict__ count, int64_t Q, int k, bool valid_only, unsigned long long* __restrict__ state,
                       int64_t q_offset, int32_t* __restrict__ out_q, int32_t* __restrict__ out_m, float* __restrict__ out_s,
                       uint8_t* __restrict__ out_v, int64_t* __restrict__ out_total) {
  __shared__ int sm[32];
  __shared__ int offs[kScanBlock];
  __shared__ int cnts[kScanBlock];
  __shared__ int tot_s;
  __shared__ unsigned tile_s;
  __shared__ long long base_s;
  const unsigned tiles = gridDim.x;
  griddep_wait();
#ifdef SEMGATE_K4_TRACE
  unsigned long long t_top; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_top));
#endif
  if (threadIdx.x == 0) tile_s = atomicAdd(reinterpret_cast<unsigned*>(state + tiles), 1u);
  __syncthreads();
  const unsigned tile = tile_s;
#ifdef SEMGATE_K4_TRACE
  if (threadIdx.x == 0) K4_NOTE(0, t_top);
#endif
  K4_STAMP(1);
  const int64_t row0 = static_cast<int64_t>(tile) * kScanBlock;
  const int64_t row = row0 + threadIdx.x;
  const bool has_row = threadIdx.x < kScanBlock && row < Q;     // threads beyond the tile's rows only help to scatter
  const int v = has_row ? row_emit_count(count, valid, row, k, valid_only) : 0;
  const int ex = block_exclusive_scan(v, sm, &tot_s);
  if (threadIdx.x < kScanBlock) {
    offs[threadIdx.x] = ex;
    cnts[threadIdx.x] = has_row ? count[row] : 0;
  }
  __syncthreads();
  K4_STAMP(2);
  if (threadIdx.x < 32) {
    // look-back by the first warp, 32 tiles at a time: lane l inspects tile j - l (a tile before tile 0 counts as an
    // inclusive prefix of 0); the nearest inclusive prefix ends the walk, everything nearer contributes its total
    const int lane = threadIdx.x;
    const unsigned long long tot = static_cast<unsigned long long>(tot_s);
    unsigned long long prefix = 0;
    if (tile > 0) {
      if (lane == 0) st_state_u64(state + tile, (tot << 2) | 1ull);
      long long j = static_cast<long long>(tile) - 1;
      uint64_t t0 = 0;
      uint32_t spins = 0;
#ifdef SEMGATE_K4_TRACE
      unsigned long long rounds = 0;
#endif
      while (true) {
#ifdef SEMGATE_K4_TRACE
        ++rounds;
#endif
        const long long at = j - lane;
        const unsigned long long sv = at >= 0 ? ld_state_u64(state + at) : 2ull;
        if (__any_sync(0xffffffffu, (sv & 3ull) == 0ull)) {   // some tile of the window has not published yet: it is
          if ((++spins & 0xFFFu) == 0) {                      // running (tickets).  Bounded like every wait here.
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
          }
          continue;
        }
        const uint32_t incl = __ballot_sync(0xffffffffu, (sv & 3ull) == 2ull);
        const int stop = incl ? __ffs(incl) - 1 : 31;         // nearest inclusive prefix in the window, if any
        unsigned long long part = lane <= stop ? (sv >> 2) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        prefix += part;
        if (incl) break;
        j -= 32;
      }
#ifdef SEMGATE_K4_TRACE
      if (lane == 0) { K4_NOTE(5, rounds); K4_NOTE(6, static_cast<unsigned long long>(spins)); }
#endif
    }
    if (lane == 0) {
      st_state_u64(state + tile, ((prefix + tot) << 2) | 2ull);
      base_s = static_cast<long long>(prefix);
      if (tile == tiles - 1) *out_total = static_cast<int64_t>(prefix + tot);
    }
  }
  __syncthreads();
  K4_STAMP(3);
  const int64_t base = base_s;
  if (!valid_only) {
    // Everything is emitted: one thread per OUTPUT element.  The block's rows start at offs[]; a binary search over
    // them finds the element's row (the last row whose offset is <= the element: rows without entries share their
    // successor's offset and never win), its position in the row follows.  Loads of different elements are
    // independent and the stores are perfectly coalesced.
    const int total = tot_s;
    for (int o = threadIdx.x; o < total; o += blockDim.x) {
      int lo = 0, hi = kScanBlock - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (offs[mid] <= o) lo = mid; else hi = mid - 1;
      }
      const int i = o - offs[lo];
      const int64_t src = (row0 + lo) * k + i;
      out_q[base + o] = static_cast<int32_t>(q_offset + row0 + lo);
      out_m[base + o] = __ldg(idx + src);
      out_s[base + o] = __ldg(scores + src);
      out_v[base + o] = __ldg(valid + src);
    }
#ifdef SEMGATE_K4_TRACE
    __syncthreads();
    K4_STAMP(4);
#endif
    return;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = w; r < kScanBlock && row0 + r < Q; r += (blockDim.x >> 5)) {
    const int c = cnts[r];
    int64_t o = base + offs[r];
    const int64_t src = (row0 + r) * k;
    for (int i0 = 0; i0 < c; i0 += 32) {
      const int i = i0 + lane;
      const bool live = i < c;
      const uint8_t vv = live ? valid[src + i] : 0;
      const bool emit = live && vv != 0;
      const uint32_t ballot = __ballot_sync(0xffffffffu, emit);
      if (emit) {
        const int64_t d = o + __popc(ballot & ((1u << lane) - 1u));   // order inside the row is kept
        out_q[d] = static_cast<int32_t>(q_offset + row0 + r);
        out_m[d] = idx[src + i];
        out_s[d] = scores[src + i];
        out_v[d] = vv;
      }
      o += __popc(ballot);
    }
  }
}

size_t compact_workspace_bytes(int64_t Q) {
  return static_cast<size_t>((Q + kScanBlock - 1) / kScanBlock + 2) * sizeof(int64_t);
}

int launch_compact(const float* scores, const int32_t* idx, const uint8_t* valid, const int32_t* count, int64_t Q, int k,
                   bool valid_only, int64_t q_offset, int32_t* out_q, int32_t* out_m, float* out_s, uint8_t* out_v,
                   int64_t* out_total, void* workspace, cudaStream_t st, bool state_zeroed, bool pdl) {
  if (Q <= 0) return static_cast<int>(cudaMemsetAsync(out_total, 0, sizeof(int64_t), st));
  const int64_t nb = (Q + kScanBlock - 1) / kScanBlock;
  unsigned long long* state = static_cast<unsigned long long*>(workspace);
  if (!state_zeroed) {      // (the one-call sweep lets K3 zero it: compact_state_words)
    cudaError_t e = cudaMemsetAsync(state, 0, static_cast<size_t>(nb + 1) * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  // A tile's scatter is a chain of dependent round trips per thread (search -> loads -> stores, 25 of them with 128 threads
  // and full lists): with few tiles (one wave or so) the kernel's duration IS that chain, so 512 threads share a tile
  // (config 1's graph step 68 -> 58 us, 20k x 1024-d 380 -> 371 us); with many tiles there are enough blocks per SM to
  // hide it and the small block wins (1M x 25 full lists: 0.175 ms with 128 threads, 0.22 ms with 512).
  static const int forced = [] { const char* e = getenv("SEMGATE_K4_THREADS"); const int v = e ? atoi(e) : 0; return (v == 128 || v == 256 || v == 512) ? v : 0; }();
  const int threads = forced ? forced : (nb <= 4 * 148 ? kCompactMaxThreads : kScanBlock);
  return static_cast<int>(launch_maybe_pdl(pdl, compact_onepass_kernel, dim3(static_cast<unsigned>(nb)), dim3(threads), 0, st, scores, idx, valid,
                                           count, Q, k, valid_only, state, q_offset, out_q, out_m, out_s, out_v, out_total));
}
int compact_state_words(int64_t Q) { return static_cast<int>((Q + kScanBlock - 1) / kScanBlock + 1); }

// =========================================================================== match statistics
// get_statistics (place_recognition.py:913-933) over a device-resident candidate list: number of
// valid candidates, sum of similarities, sum over the valid ones, in fp64.  Every block reduces a
// grid-stride slice; the block that finishes last adds the per-block partials in index order, so
// the result does not depend on scheduling.
constexpr int kStatsBlocks = 296;
constexpr int kStatsThreads = 256;

__device__ __forceinline__ double block_sum_f64(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = (l < (blockDim.x >> 5)) ? red[l] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(kStatsThreads)
candidate_stats_kernel(const float* __restrict__ sim, const uint8_t* __restrict__ valid, const int64_t* __restrict__ total_dev,
                       int64_t M_host, double* __restrict__ partials /*[grid][3]*/, unsigned* __restrict__ done,
                       double* __restrict__ out /*[4]: total, valid, sum, sum_valid*/) {
  __shared__ double red[32];
  __shared__ bool last;
  const int64_t M = total_dev ? *total_dev : M_host;
  double nv = 0.0, ss = 0.0, sv = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < M;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double s = static_cast<double>(sim[i]);
    const bool v = valid[i] != 0;
    ss += s;
    if (v) { nv += 1.0; sv += s; }
  }
  nv = block_sum_f64(nv, red);
  ss = block_sum_f64(ss, red);
  sv = block_sum_f64(sv, red);
  if (threadIdx.x == 0) {
    partials[3 * blockIdx.x] = nv; partials[3 * blockIdx.x + 1] = ss; partials[3 * blockIdx.x + 2] = sv;
    __threadfence();
    last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double a = 0.0, b = 0.0, c = 0.0;
    for (unsigned g = 0; g < gridDim.x; ++g) {
      a += __ldcg(partials + 3 * g); b += __ldcg(partials + 3 * g + 1); c += __ldcg(partials + 3 * g + 2);
    }
    out[0] = static_cast<double>(M); out[1] = a; out[2] = b; out[3] = c;
    *done = 0;   // ready for the next launch
  }
}

size_t stats_workspace_bytes() { return static_cast<size_t>(kStatsBlocks) * 3 * sizeof(double) + 256; }

int launch_candidate_stats(const float* sim, const uint8_t* valid, const int64_t* total_dev, int64_t M, void* workspace,
                           double* out, cudaStream_t st) {
  double* partials = static_cast<double*>(workspace);
  unsigned* done = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + static_cast<size_t>(kStatsBlocks) * 3 * sizeof(double));
  cudaError_t e = cudaMemsetAsync(done, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  candidate_stats_kernel<<<kStatsBlocks, kStatsThreads, 0, st>>>(sim, valid, total_dev, M, partials, done, out);
  return static_cast<int>(cudaGetLastError());
}

// =========================================================================== floor gate over pairs
// the gate class has no None labels and no "off" switch: diff > max -> reject (loop_closure_gate.py:89-101)
__device__ __forceinline__ uint32_t gate_one(const int32_t* __restrict__ floors, int64_t n_floors, int32_t q, int32_t m,
                                             int max_floor_diff, unsigned& acc, unsigned& rej, unsigned& bad) {
  if (q < 0 || q >= n_floors || m < 0 || m >= n_floors) { ++bad; return 0u; }
  int64_t dfl = static_cast<int64_t>(__ldg(floors + q)) - static_cast<int64_t>(__ldg(floors + m));
  if (dfl < 0) dfl = -dfl;
  const bool ok = dfl <= max_floor_diff;
  if (ok) ++acc; else ++rej;
  return ok ? 1u : 0u;
}
// the same against a copy of the label table in shared memory
__device__ __forceinline__ uint32_t gate_one_smem(const int32_t* tab, int n_floors, int32_t q, int32_t m, int max_floor_diff,
                                                  unsigned& acc, unsigned& rej, unsigned& bad) {
  if (static_cast<uint32_t>(q) >= static_cast<uint32_t>(n_floors) || static_cast<uint32_t>(m) >= static_cast<uint32_t>(n_floors)) { ++bad; return 0u; }
  int64_t dfl = static_cast<int64_t>(tab[q]) - static_cast<int64_t>(tab[m]);
  if (dfl < 0) dfl = -dfl;
  const bool ok = dfl <= max_floor_diff;
  if (ok) ++acc; else ++rej;
  return ok ? 1u : 0u;
}

__device__ __forceinline__ void gate_counts_out(unsigned acc, unsigned rej, unsigned bad, unsigned long long* __restrict__ counts) {
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

// VEC: four candidates per thread per step (16-byte index loads, eight label gathers in
// flight, one 4-byte store); needs 16-byte aligned index arrays and a 4-byte aligned output.
template <bool VEC>
__global__ void __launch_bounds__(256)
gate_candidates_kernel(const int32_t* __restrict__ floors, int64_t n_floors, const int32_t* __restrict__ q_idx,
                       const int32_t* __restrict__ m_idx, int64_t M, int max_floor_diff, uint8_t* __restrict__ out_valid,
                       unsigned long long* __restrict__ counts) {
  unsigned acc = 0, rej = 0, bad = 0;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
  if constexpr (VEC) {
    const int64_t M4 = M >> 2;
    for (int64_t i = tid; i < M4; i += nthreads) {
      const int4 q = __ldcs(reinterpret_cast<const int4*>(q_idx) + i);
      const int4 m = __ldcs(reinterpret_cast<const int4*>(m_idx) + i);
      uint32_t r = gate_one(floors, n_floors, q.x, m.x, max_floor_diff, acc, rej, bad);
      r |= gate_one(floors, n_floors, q.y, m.y, max_floor_diff, acc, rej, bad) << 8;
      r |= gate_one(floors, n_floors, q.z, m.z, max_floor_diff, acc, rej, bad) << 16;
      r |= gate_one(floors, n_floors, q.w, m.w, max_floor_diff, acc, rej, bad) << 24;
      __stcs(reinterpret_cast<uint32_t*>(out_valid) + i, r);
    }
    for (int64_t i = (M4 << 2) + tid; i < M; i += nthreads)
      out_valid[i] = static_cast<uint8_t>(gate_one(floors, n_floors, q_idx[i], m_idx[i], max_floor_diff, acc, rej, bad));
  } else {
    for (int64_t i = tid; i < M; i += nthreads)
      out_valid[i] = static_cast<uint8_t>(gate_one(floors, n_floors, q_idx[i], m_idx[i], max_floor_diff, acc, rej, bad));
  }
  gate_counts_out(acc, rej, bad, counts);
}

// Label table in shared memory.  Gathering labels from global memory costs a 32-byte L2 sector per 4-byte label: the kernel
// above is bound by L2 sector traffic (ncu: lts throughput 73 %, DRAM 25 %; 0.31 of the copy bandwidth).  The label table of
// a trajectory is small -- one int32 per keyframe, 77 KB for the 19 163 poses of the reference's published gate counts
// (5.1 M candidates) -- so when it fits beside the block (<= 50 K labels) and there are enough candidates to pay for one
// copy of it per block, every block keeps its own copy and the gathers become shared-memory loads: what remains is the
// 9 bytes per candidate that must cross HBM.
constexpr int kGateSmemThreads = 1024;
constexpr int kGateSmemMaxLabels = 50 * 1024;
__global__ void __launch_bounds__(kGateSmemThreads, 2)
gate_candidates_smem_kernel(const int32_t* __restrict__ floors, int n_floors, const int32_t* __restrict__ q_idx,
                            const int32_t* __restrict__ m_idx, int64_t M, int max_floor_diff, uint8_t* __restrict__ out_valid,
                            unsigned long long* __restrict__ counts) {
  extern __shared__ int32_t gate_tab[];
  for (int i = threadIdx.x; i < n_floors; i += kGateSmemThreads) gate_tab[i] = __ldg(floors + i);
  __syncthreads();
  unsigned acc = 0, rej = 0, bad = 0;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * kGateSmemThreads + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * kGateSmemThreads;
  const int64_t M4 = M >> 2;
  for (int64_t i = tid; i < M4; i += nthreads) {
    const int4 q = __ldcs(reinterpret_cast<const int4*>(q_idx) + i);
    const int4 m = __ldcs(reinterpret_cast<const int4*>(m_idx) + i);
    uint32_t r = gate_one_smem(gate_tab, n_floors, q.x, m.x, max_floor_diff, acc, rej, bad);
    r |= gate_one_smem(gate_tab, n_floors, q.y, m.y, max_floor_diff, acc, rej, bad) << 8;
    r |= gate_one_smem(gate_tab, n_floors, q.z, m.z, max_floor_diff, acc, rej, bad) << 16;
    r |= gate_one_smem(gate_tab, n_floors, q.w, m.w, max_floor_diff, acc, rej, bad) << 24;
    __stcs(reinterpret_cast<uint32_t*>(out_valid) + i, r);
  }
  for (int64_t i = (M4 << 2) + tid; i < M; i += nthreads)
    out_valid[i] = static_cast<uint8_t>(gate_one_smem(gate_tab, n_floors, q_idx[i], m_idx[i], max_floor_diff, acc, rej, bad));
  gate_counts_out(acc, rej, bad, counts);
}

int launch_gate_candidates(const int32_t* floors, int64_t n_floors, const int32_t* q_idx, const int32_t* m_idx, int64_t M,
                           int max_floor_diff, uint8_t* out_valid, unsigned long long* counts, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(counts, 0, 3 * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (M <= 0) return 0;
  const bool vec = ((reinterpret_cast<uintptr_t>(q_idx) | reinterpret_cast<uintptr_t>(m_idx)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out_valid) & 3) == 0;
  // one copy of the table per block is n_floors labels through L2; a gather costs 8 labels' worth (a sector) per label
  if (vec && n_floors > 0 && n_floors <= kGateSmemMaxLabels && M >= 10 * n_floors) {
    const size_t smem = static_cast<size_t>(n_floors) * sizeof(int32_t);
    e = cudaFuncSetAttribute(gate_candidates_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    const int per_sm = smem + 1024 <= 113 * 1024 ? 2 : 1;
    const int64_t want = (M / 4 + 4 * kGateSmemThreads - 1) / (4 * kGateSmemThreads);    // >= 4 steps per thread
    const unsigned grid = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(want, 148 * per_sm)));
    gate_candidates_smem_kernel<<<grid, kGateSmemThreads, smem, st>>>(floors, static_cast<int>(n_floors), q_idx, m_idx, M, max_floor_diff,
                                                                      out_valid, counts);
    return static_cast<int>(cudaGetLastError());
  }
  const int64_t work = vec ? (M + 3) / 4 : M;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>((work + 255) / 256, 148 * 8));
  if (vec)
    gate_candidates_kernel<true><<<grid, 256, 0, st>>>(floors, n_floors, q_idx, m_idx, M, max_floor_diff, out_valid, counts);
  else
    gate_candidates_kernel<false><<<grid, 256, 0, st>>>(floors, n_floors, q_idx, m_idx, M, max_floor_diff, out_valid, counts);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace semgate
