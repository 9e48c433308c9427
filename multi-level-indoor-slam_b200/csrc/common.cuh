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

// ---- running top-k list of one query row: k unsorted keys (shared memory), admission bound = threshold
// until the list is full, then its smallest key's score; insertion replaces the minimum and rescans.
struct RowList {
  uint64_t* keys;     // this thread's row in shared memory
  int cnt;
  int min_pos;
  uint64_t min_key;
  float f;            // current admission bound on the score

  __device__ __forceinline__ void reset(float thr) { cnt = 0; min_pos = 0; min_key = 0; f = thr; }

  __device__ __forceinline__ void rescan(int k) {
    uint64_t mk = keys[0];
    int mp = 0;
    for (int i = 1; i < k; ++i) {
      uint64_t v = keys[i];
      if (v < mk) { mk = v; mp = i; }
    }
    min_key = mk; min_pos = mp; f = key_score(mk);
  }

  __device__ __forceinline__ void insert(uint64_t key, int k) {
    if (cnt < k) {
      keys[cnt++] = key;
      if (cnt == k) rescan(k);
    } else if (key > min_key) {
      keys[min_pos] = key;
      rescan(k);
    }
  }
};

// ---- tile schedule of the fused kernel.
// The output is tiled BM x BN.  Query blocks ("m-blocks") are processed in
// super-rows of `rm` consecutive m-blocks; inside a super-row every m-block's
// database range is split into `s` contiguous runs of n-tiles, one run per
// CTA (or CTA pair), so that rm*s <= units.  All units of a super-row stream
// the same `s` database tiles at about the same time (each tile is fetched from
// DRAM once and served to the other rm - 1 readers by L2) and each unit keeps
// one running top-k list per query row; the `s` partial lists of a row are
// merged afterwards.  The last super-row may hold fewer m-blocks and is split
// finer.
//
// Pacing.  The L2 only helps while the units of a super-row stay close together.
// Every unit counts the chunks (`pace_kb` k-blocks of one tile) it has issued on a
// per-super-row counter array, and does not start chunk p before ALL units of the
// super-row have issued chunk p - sync_window.  The window is sized so that the
// streamed operands of the whole grid inside it fit in L2 next to the resident
// query blocks.  With very long descriptors (query blocks no longer L2-resident,
// `a_resident` = 0) both operands are streamed and the same lock-step makes every
// k-slice of every operand come from DRAM once per super-row step.
struct RunEntry { int mb, list, nt0, nt1; };   // one run of the run table: block, list number, tiles [nt0, nt1)

struct Schedule {
  int mblocks;      // ceil(Q / BM)            (for pairs: counted in pair-rows of 2*BM)
  int ntiles;       // ceil(N / BN)
  int rm;           // m-blocks per full super-row
  int s_main;       // splits per m-block in full super-rows
  int n_full;       // number of full super-rows
  int r_last;       // m-blocks in the trailing partial super-row (0 if none)
  int s_last;       // splits per m-block there
  int s_max;        // max(s_main, s_last)
  int a_resident;   // the super-row's query blocks stay in L2 across database tiles
  int sync_window;  // pacing window in chunks (0 = pacing off)
  int pace_kb;      // k-blocks per chunk
  int cpt;          // chunks per tile = ceil(kblocks / pace_kb)
  int len_main;     // ceil(ntiles / s_main): longest run of a full super-row, in tiles
  int len_last;     // ceil(ntiles / s_last)
  int sym;          // 1: the queries ARE the database (all-pairs sweep, S = S^T): only tiles on or above the
                    //    block diagonal are computed, see "Symmetric sweep" below
  int part_index;   // symmetric sweep split over `part_count` GPUs: this launch computes the super-rows that
  int part_count;   //    sched_owned() gives part_index (0 / 0 or 1: everything)
  // Run table (small symmetric sweeps).  A triangle of a few thousand tiles is coarse for the super-row formula:
  // every super-row lasts as long as its longest block.  Instead the host cuts every block's tiles into short runs
  // (super-row by super-row, column chunk by column chunk, so that runs close in the order share database tiles),
  // deals them in that order to whichever unit is free first, and uploads the result: unit u executes
  // tab_runs[tab_unit_begin[u] .. tab_unit_begin[u+1]); the lists of block b are numbered
  // tab_block_first[b] .. tab_block_first[b+1) (none for blocks another part owns).  No pacing in this mode.
  const RunEntry* tab_runs;
  const int* tab_unit_begin;
  const int* tab_block_first;
  int tab_lists;    // lists in all
  long long tab_tiles;   // tiles the table computes (this part's share of the triangle)
};

// Super-rows get shorter towards the end of a symmetric sweep; dealing them out boustrophedon
// (0 1 .. G-1 G-1 .. 1 0 0 1 ..) keeps the parts' tile counts within a fraction of a percent.
__host__ __device__ __forceinline__ bool sched_owned(const Schedule& sc, int sr) {
  if (sc.part_count <= 1) return true;
  const int g = sr % (2 * sc.part_count);
  return (g < sc.part_count ? g : 2 * sc.part_count - 1 - g) == sc.part_index;
}

// Symmetric sweep (sym = 1, CTA-pair tiles only, so that a query block and a database tile are both 256
// rows and share one numbering).  Super-row sr holds the blocks [sr*rm, sr*rm + r) and sweeps the tiles
// [sr*rm, ntiles) as a rectangle, split into min(S, length) runs; block b skips the tiles left of its
// diagonal tile (nt < b), which only shortens the first run(s).  A tile right of the diagonal feeds two
// sets of lists: the rows' (as always) and, transposed, the columns' (gated_topk.cuh, "column direction").
__host__ __device__ __forceinline__ int sched_sym_splits(const Schedule& sc, int sr) {
  const int len = sc.ntiles - sr * sc.rm;
  const int S = sr < sc.n_full ? sc.s_main : sc.s_last;
  return S < len ? S : len;
}

__host__ __device__ __forceinline__ int64_t sched_sync_counters(const Schedule& sc) {
  return (static_cast<int64_t>(sc.n_full) * sc.len_main + sc.len_last) * sc.cpt;
}

__host__ __device__ __forceinline__ int sched_slots(const Schedule& sc, int mb) {
  if (sc.tab_runs != nullptr) return sc.tab_block_first[mb + 1] - sc.tab_block_first[mb];
  if (sc.sym) return sched_owned(sc, mb / sc.rm) ? sched_sym_splits(sc, mb / sc.rm) : 0;   // other parts' rows: no lists here
  return mb < sc.n_full * sc.rm ? sc.s_main : sc.s_last;
}

// Partial lists of the fused kernel: row r keeps one k-entry list per run of its m-block, rows of
// full super-rows (s_main lists each) first, then the rows of the tail super-row (s_last lists each).
__host__ __device__ __forceinline__ int64_t sched_list_offset(const Schedule& sc, int64_t row, int rows_per_mblock, int k) {
  const int64_t rows_full = static_cast<int64_t>(sc.n_full) * sc.rm * rows_per_mblock;
  return row < rows_full ? row * sc.s_main * k : (rows_full * sc.s_main + (row - rows_full) * sc.s_last) * k;
}
__host__ __device__ __forceinline__ int64_t sched_list_keys(const Schedule& sc, int rows_per_mblock, int k) {
  if (sc.tab_runs != nullptr) return static_cast<int64_t>(sc.tab_lists) * rows_per_mblock * k;
  return sched_list_offset(sc, static_cast<int64_t>(sc.mblocks) * rows_per_mblock, rows_per_mblock, k);
}
// Where row `row` keeps the list of run (block mb, list number `slot`), and how far apart the lists of one row
// are.  Formula schedules: a row's lists are consecutive.  Run table: list-major, [list][row in block][k].
__host__ __device__ __forceinline__ int64_t sched_run_list_offset(const Schedule& sc, int mb, int slot, int64_t row,
                                                                  int rows_per_mblock, int k) {
  if (sc.tab_runs != nullptr)
    return (static_cast<int64_t>(slot) * rows_per_mblock + (row - static_cast<int64_t>(mb) * rows_per_mblock)) * k;
  return sched_list_offset(sc, row, rows_per_mblock, k) + static_cast<int64_t>(slot) * k;
}
__host__ __device__ __forceinline__ int64_t sched_list_stride(const Schedule& sc, int rows_per_mblock, int k) {
  return sc.tab_runs != nullptr ? static_cast<int64_t>(rows_per_mblock) * k : k;
}

// balanced contiguous split of [0, n) into s parts
__host__ __device__ __forceinline__ int split_begin(int n, int s, int j) {
  return static_cast<int>((static_cast<int64_t>(n) * j) / s);
}

// One unit's work list, identical for every warp role.
struct Run {
  int mb;      // m-block (pair-row for CG = 2)
  int slot;    // which of the m-block's runs (partial-list slot)
  int nt0;     // first n-tile of the run (its position in the pacing counters)
  int nt_first;  // first n-tile actually computed (symmetric sweep: max(nt0, mb); else nt0)
  int nt1;     // one past the last n-tile
  // pacing: counters of this super-row; all `units_all` units reach tile ordinal < short_len,
  // only the `units_long` units with the longer runs reach ordinal == short_len
  int64_t sync_base;
  int short_len;
  int units_all;
  int units_long;
};

// (also walked on the host by schedule_selfcheck with a host-only lambda: that instantiation is never compiled
//  for the device, so the "host function called from __host__ __device__" diagnostics do not apply)
#pragma nv_diag_suppress 20011, 20013, 20015
template <typename F>
__host__ __device__ __forceinline__ void for_each_run(const Schedule& sc, int unit, F&& f) {
  if (sc.tab_runs != nullptr) {
    for (int i = sc.tab_unit_begin[unit]; i < sc.tab_unit_begin[unit + 1]; ++i) {
      const RunEntry e = sc.tab_runs[i];
      Run run;
      run.mb = e.mb; run.slot = e.list; run.nt0 = e.nt0; run.nt_first = e.nt0; run.nt1 = e.nt1;
      run.sync_base = 0; run.short_len = 0; run.units_all = 0; run.units_long = 0;
      f(run);
    }
    return;
  }
  const int n_sr = sc.n_full + (sc.r_last > 0 ? 1 : 0);
  for (int sr = 0; sr < n_sr; ++sr) {
    if (!sched_owned(sc, sr)) continue;
    const bool full_sr = sr < sc.n_full;
    const int r = full_sr ? sc.rm : sc.r_last;
    const int lo = sc.sym ? sr * sc.rm : 0;          // first tile the super-row sweeps
    const int len = sc.ntiles - lo;
    const int S = sc.sym ? sched_sym_splits(sc, sr) : (full_sr ? sc.s_main : sc.s_last);
    if (unit >= r * S) continue;
    Run run;
    run.mb = sr * sc.rm + unit % r;
    run.slot = unit / r;
    run.nt0 = lo + split_begin(len, S, run.slot);
    run.nt1 = lo + split_begin(len, S, run.slot + 1);
    run.nt_first = (sc.sym && run.mb > run.nt0) ? (run.mb < run.nt1 ? run.mb : run.nt1) : run.nt0;
    run.sync_base = (full_sr ? static_cast<int64_t>(sr) * sc.len_main : static_cast<int64_t>(sc.n_full) * sc.len_main) * sc.cpt;
    run.short_len = len / S;
    run.units_all = r * S;
    run.units_long = r * (len % S);
    // a symmetric run may be empty (all of it left of the diagonal): it still flushes an empty list
    if (run.nt0 < run.nt1) f(run);
  }
}
#pragma nv_diag_default 20011, 20013, 20015

}  // namespace semgate
