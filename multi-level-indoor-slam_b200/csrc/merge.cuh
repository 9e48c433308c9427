// Warp-level merge of candidate keys (shared by K3 and the streaming-query kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace semgate {

__device__ __forceinline__ void cmp_swap_desc(uint64_t& a, uint64_t& b) {
  const bool sw = a < b;
  const uint64_t hi = sw ? b : a, lo = sw ? a : b;
  a = hi; b = lo;
}

template <int P>
__device__ __forceinline__ void sort_desc(uint64_t (&c)[P]) {
  if constexpr (P == 4) {
    cmp_swap_desc(c[0], c[1]); cmp_swap_desc(c[2], c[3]);
    cmp_swap_desc(c[0], c[2]); cmp_swap_desc(c[1], c[3]);
    cmp_swap_desc(c[1], c[2]);
  } else {   // P == 8: Batcher's odd-even merge sort, 19 comparators
    cmp_swap_desc(c[0], c[1]); cmp_swap_desc(c[2], c[3]); cmp_swap_desc(c[4], c[5]); cmp_swap_desc(c[6], c[7]);
    cmp_swap_desc(c[0], c[2]); cmp_swap_desc(c[1], c[3]); cmp_swap_desc(c[4], c[6]); cmp_swap_desc(c[5], c[7]);
    cmp_swap_desc(c[1], c[2]); cmp_swap_desc(c[5], c[6]);
    cmp_swap_desc(c[0], c[4]); cmp_swap_desc(c[1], c[5]); cmp_swap_desc(c[2], c[6]); cmp_swap_desc(c[3], c[7]);
    cmp_swap_desc(c[2], c[4]); cmp_swap_desc(c[3], c[5]);
    cmp_swap_desc(c[1], c[2]); cmp_swap_desc(c[3], c[4]); cmp_swap_desc(c[5], c[6]);
  }
}

// Merge the keys base[(e / kk) * list_stride + e % kk], e in [e0, e1), into `run` (the warp's running
// top-k list: rank t lives in run[t>>5] of lane t&31).  `have_run`: run already holds keys.
// `ptrs` != nullptr: list g lives at ptrs[g] + base_off (peer GPUs' buffers), else at base + g * list_stride.
template <int P>
__device__ __forceinline__ void merge_range(const uint64_t* __restrict__ base, int e0, int e1, int kk, int64_t list_stride,
                                            int k, int lane, uint64_t (&run)[2], bool have_run,
                                            const uint64_t* const* __restrict__ ptrs = nullptr, int64_t base_off = 0) {
  int b0 = e0;
  do {
    const int nnew = have_run ? P - 2 : P;   // key slots per lane for new candidates
    uint64_t c[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const int e = b0 + i * 32 + lane;
      uint64_t v = 0ull;
      if (i < nnew && e < e1) {
        const int g = e / kk, j = e - g * kk;
        v = ptrs ? ptrs[g][base_off + j] : base[static_cast<int64_t>(g) * list_stride + j];
      }
      c[i] = v;
    }
    if (have_run) { c[P - 2] = run[0]; c[P - 1] = run[1]; }
    b0 += 32 * nnew;
    sort_desc<P>(c);
    run[0] = run[1] = 0ull;
    for (int t = 0; t < k; ++t) {
      // warp-wide max of the heads (64-bit as two 32-bit reductions)
      const uint32_t hi = static_cast<uint32_t>(c[0] >> 32), lo = static_cast<uint32_t>(c[0]);
      const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
      const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
      if ((mh | ml) == 0u) break;                        // warp-uniform: nothing left
      const bool mine = hi == mh && lo == ml;
#pragma unroll
      for (int i = 0; i + 1 < P; ++i) c[i] = mine ? c[i + 1] : c[i];
      c[P - 1] = mine ? 0ull : c[P - 1];
      const uint64_t best = (static_cast<uint64_t>(mh) << 32) | ml;
      if (t == lane) run[0] = best;                      // static register indices: no local-memory array
      if (t == lane + 32) run[1] = best;
    }
    have_run = true;
  } while (b0 < e1);
}

}  // namespace semgate
