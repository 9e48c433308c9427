// Shared definitions for the semgate kernels: candidate keys, floor-gate
// predicate, tile schedule.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace semgate {

constexpr int32_t kFloorNone = INT32_MIN;  // encodes floor_label=None (place_recognition.py:78,898)
constexpr int kMaxK = 64;

// ---- candidate key: (score, index) packed so that a larger key is a better
// candidate under the total order (score descending, index ascending).
// key == 0 is the empty slot (no finite or infinite score maps to it).
__host__ __device__ __forceinline__ uint32_t score_to_ordered(float s) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(s);
#else
  union { float f; uint32_t u; } c; c.f = s; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_score(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t pack_key(float s, uint32_t idx) {
  return (static_cast<uint64_t>(score_to_ordered(s)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - idx);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) { return ordered_to_score(static_cast<uint32_t>(key >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_index(uint64_t key) { return 0xFFFFFFFFu - static_cast<uint32_t>(key); }

// ---- floor gate (loop_closure_gate.py:89-101, place_recognition.py:897-899)
// max_floor_diff < 0: gating off; 0: strict; 1: non-strict (+-1 floor allowed).
__host__ __device__ __forceinline__ bool floor_ok(int32_t qf, int32_t mf, int max_floor_diff) {
  if (max_floor_diff < 0) return true;
  if (qf == kFloorNone || mf == kFloorNone) return true;
  int64_t d = static_cast<int64_t>(qf) - static_cast<int64_t>(mf);
  if (d < 0) d = -d;
  return d <= static_cast<int64_t>(max_floor_diff);
}

// ---- temporal exclusion (place_recognition.py:884 / :146): strict, fp64
__host__ __device__ __forceinline__ bool time_excluded(double t_db, double t_q, double gap) {
  double d = t_db - t_q;
  if (d < 0) d = -d;
  return d < gap;
}

// ---- tile schedule of the fused kernel.
// The output is tiled BM x BN.  Query blocks ("m-blocks") are processed in
// super-rows of `rm` consecutive m-blocks; inside a super-row every m-block's
// database range is split into `s` contiguous runs of n-tiles, one run per
// CTA (or CTA pair), so that rm*s <= units.  All units of a super-row stream
// the same database tiles at about the same time (L2 reuse) and each unit
// keeps one running top-k list per query row; the `s` partial lists of a row
// are merged afterwards.  The last super-row may hold fewer m-blocks and is
// split finer.
// Long databases are additionally cut into `n_panels` column panels that fit
// in L2, visited panel-major (for panel: for super-row: run): units that drift
// apart over a long sweep still find the panel's tiles in L2 instead of each
// streaming them from DRAM.  A row's list is carried from panel to panel through
// its partial-list slot in HBM.
struct Schedule {
  int mblocks;      // ceil(Q / BM)            (for pairs: counted in pair-rows of 2*BM)
  int ntiles;       // ceil(N / BN)
  int rm;           // m-blocks per full super-row
  int s_main;       // splits per m-block in full super-rows
  int n_full;       // number of full super-rows
  int r_last;       // m-blocks in the trailing partial super-row (0 if none)
  int s_last;       // splits per m-block there
  int s_max;        // max(s_main, s_last): slot stride of the partial lists
  int n_panels;     // column panels (>= 1); every panel holds >= s_max tiles
};

__host__ __device__ __forceinline__ int sched_slots(const Schedule& sc, int mb) {
  return mb < sc.n_full * sc.rm ? sc.s_main : sc.s_last;
}

// balanced contiguous split of [0, n) into s parts
__host__ __device__ __forceinline__ int split_begin(int n, int s, int j) {
  return static_cast<int>((static_cast<int64_t>(n) * j) / s);
}

// One unit's work list, identical for every warp role: f(mb, slot, nt0, nt1, carry).
struct Run {
  int mb;      // m-block (pair-row for CG = 2)
  int slot;    // which of the m-block's runs (partial-list slot)
  int nt0;     // first n-tile
  int nt1;     // one past the last n-tile
  bool carry;  // the row lists continue from an earlier panel
};

template <typename F>
__device__ __forceinline__ void for_each_run(const Schedule& sc, int unit, F&& f) {
  const int n_sr = sc.n_full + (sc.r_last > 0 ? 1 : 0);
  for (int p = 0; p < sc.n_panels; ++p) {
    const int pt0 = split_begin(sc.ntiles, sc.n_panels, p);
    const int len = split_begin(sc.ntiles, sc.n_panels, p + 1) - pt0;
    for (int sr = 0; sr < n_sr; ++sr) {
      const bool full_sr = sr < sc.n_full;
      const int r = full_sr ? sc.rm : sc.r_last;
      const int S = full_sr ? sc.s_main : sc.s_last;
      if (unit >= r * S) continue;
      Run run;
      run.mb = sr * sc.rm + unit % r;
      run.slot = unit / r;
      run.nt0 = pt0 + split_begin(len, S, run.slot);
      run.nt1 = pt0 + split_begin(len, S, run.slot + 1);
      run.carry = p > 0;
      if (run.nt0 < run.nt1) f(run);
    }
  }
}

}  // namespace semgate
