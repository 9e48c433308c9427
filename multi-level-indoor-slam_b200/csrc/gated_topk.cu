// Host side of K2: tile schedule, TMA tensor maps, launch.
#include "gated_topk.cuh"
#include "launch.h"

#include <cudaTypedefs.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <queue>
#include <vector>

namespace semgate {

// ------------------------------------------------------------------ schedule
static long env_long(const char* name, long dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atol(e) : dflt;
}

static size_t smem_bytes_for(int cg, int stages, int kstride, bool sym = false, int sets = 1);
constexpr size_t kSmemLimit = 232448;   // 227 KB per CTA

// Schedule units (CTAs, CTA pairs, or 4-CTA clusters of two pairs) that are co-resident.  A cluster
// must sit inside one GPC, so 4-CTA clusters may not cover every SM: ask the occupancy calculator.
int topk_units(int cta_group, int sm_count) {
  if (cta_group != 4) return std::max(1, sm_count / cta_group);
  static int cached = -1;
  static std::once_flag once;
  std::call_once(once, [&] {
    int stages = kMaxStages;                       // same ring-depth rule as the launch (largest list size)
    while (stages > 2 && smem_bytes_for(2, stages, kMaxK | 1) > kSmemLimit) --stages;
    const size_t smem = smem_bytes_for(2, stages, kMaxK | 1);
    if (cudaFuncSetAttribute(gated_topk_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(4 * sm_count));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gated_topk_kernel<2, 2>, &cfg) == cudaSuccess) cached = n;
  });
  cudaGetLastError();
  return std::max(1, cached > 0 ? cached : sm_count / 4 - 5);   // conservative if the query failed
}

// makespan of a symmetric sweep in tile-times: super-row sr sweeps the tiles [sr*rm, nb) in min(S, length) runs;
// split over `parts` GPUs (sched_owned's deal) it is the busiest part's
static int64_t sym_cost(int nb, int rm, int units, int parts) {
  std::vector<int64_t> c(static_cast<size_t>(std::max(parts, 1)), 0);
  Schedule deal{};
  deal.part_count = parts;
  int sr = 0;
  for (int lo = 0; lo < nb; lo += rm, ++sr) {
    const int r = std::min(rm, nb - lo), len = nb - lo;
    const int S = std::min(len, std::max(1, units / r));
    for (int g = 0; g < std::max(parts, 1); ++g) {
      deal.part_index = g;
      if (sched_owned(deal, sr)) c[g] += (len + S - 1) / S;
    }
  }
  return *std::max_element(c.begin(), c.end());
}

Schedule make_schedule(int64_t Q, int64_t N, int d_pad, int cta_group, int sm_count, bool sym, int part_index, int part_count) {
  Schedule sc{};
  sc.sym = sym ? 1 : 0;
  sc.part_index = sym ? part_index : 0;
  sc.part_count = sym ? std::max(part_count, 1) : 1;
  const int units = topk_units(cta_group, sm_count);
  const int bm_unit = BM * cta_group;
  sc.mblocks = static_cast<int>((Q + bm_unit - 1) / bm_unit);
  sc.ntiles = static_cast<int>((N + BN - 1) / BN);
  if (sc.mblocks <= 0 || sc.ntiles <= 0) { sc.mblocks = std::max(sc.mblocks, 0); return sc; }
  // A super-row's query blocks are re-read once per database tile.  While they fit in L2
  // (`cap`; measured: the usable share of the 126 MB is about one die's half minus the streamed
  // tiles) only the database tiles come from DRAM: s * BN rows per tile-time.  Beyond that
  // both operands stream: s * BN + rm * BM rows.  Among the super-row heights whose makespan is
  // within 4 % of the best, take the one with the least DRAM traffic.
  const int64_t a_bytes = static_cast<int64_t>(bm_unit) * d_pad * 2;
  const int64_t cap = env_long("SEMGATE_RM_CAP_MB", 40) << 20;
  const int rm_hi = std::min(sc.mblocks, units);
  std::vector<int64_t> cost(rm_hi + 1, 0);
  int64_t best_cost = -1;
  for (int rm = 1; rm <= rm_hi; ++rm) {
    const int s_main = std::min(sc.ntiles, units / rm);
    const int n_full = sc.mblocks / rm;
    const int r_last = sc.mblocks % rm;
    int64_t c = static_cast<int64_t>(n_full) * ((sc.ntiles + s_main - 1) / s_main);
    if (r_last > 0) {
      const int s_last = std::min(sc.ntiles, units / r_last);
      c += (sc.ntiles + s_last - 1) / s_last;
    }
    if (sym) c = sym_cost(sc.ntiles, rm, units, sc.part_count);
    cost[rm] = c;
    if (best_cost < 0 || c < best_cost) best_cost = c;
  }
  int best_rm = 1;
  int64_t best_rows = -1;
  for (int rm = 1; rm <= rm_hi; ++rm) {
    if (cost[rm] * 100 > best_cost * 104) continue;
    const int s_main = std::min(sc.ntiles, units / rm);
    const bool resident = rm * a_bytes <= cap;
    const int64_t rows = static_cast<int64_t>(s_main) * BN + (resident ? 0 : static_cast<int64_t>(rm) * bm_unit);
    if (best_rows < 0 || rows < best_rows || (rows == best_rows && cost[rm] <= cost[best_rm])) { best_rows = rows; best_rm = rm; }
  }
  sc.rm = best_rm;
  sc.a_resident = sc.rm * a_bytes <= cap ? 1 : 0;
  sc.s_main = std::min(sc.ntiles, units / sc.rm);
  sc.n_full = sc.mblocks / sc.rm;
  sc.r_last = sc.mblocks % sc.rm;
  sc.s_last = sc.r_last > 0 ? std::min(sc.ntiles, units / sc.r_last) : 0;
  sc.s_max = std::max(sc.n_full > 0 ? sc.s_main : 0, sc.s_last);
  sc.len_main = sc.n_full > 0 ? (sc.ntiles + sc.s_main - 1) / sc.s_main : 0;
  sc.len_last = sc.r_last > 0 ? (sc.ntiles + sc.s_last - 1) / sc.s_last : 0;
  if (sym && sc.r_last > 0) {   // the tail super-row only sweeps its own r_last tiles
    const int S = std::min(sc.s_last, sc.r_last);
    sc.len_last = (sc.r_last + S - 1) / S;
  }

  // Pacing: chunks of 16 k-blocks; the window holds `SEMGATE_WINDOW_MB` of streamed operands of the
  // whole grid.  Off for short runs (nothing to drift) and with SEMGATE_WINDOW_MB=0.
  const int kblocks = d_pad / BK;
  sc.pace_kb = std::min(kblocks, 16);
  sc.cpt = (kblocks + sc.pace_kb - 1) / sc.pace_kb;
  const int64_t window_bytes = env_long("SEMGATE_WINDOW_MB", 8) << 20;
  const int64_t stream_rows = static_cast<int64_t>(sc.s_main) * BN + (sc.a_resident ? 0 : static_cast<int64_t>(sc.rm) * bm_unit);
  const int64_t chunk_bytes = stream_rows * sc.pace_kb * BK * 2;
  int window = static_cast<int>(std::min<int64_t>(64, std::max<int64_t>(2, window_bytes / std::max<int64_t>(chunk_bytes, 1))));
  const int64_t run_chunks = static_cast<int64_t>(std::max(sc.len_main, sc.len_last)) * sc.cpt;
  // one query block per super-row that stays in L2: the units share nothing that streams, nothing to pace
  const bool shares = sc.rm > 1 || !sc.a_resident;
  sc.sync_window = (window_bytes > 0 && shares && run_chunks > 4 * window) ? window : 0;
  // test knob: force a window of that many chunks wherever a run is longer than the window
  const int forced = static_cast<int>(env_long("SEMGATE_WINDOW_CHUNKS", 0));
  if (forced > 0) sc.sync_window = run_chunks > forced ? forced : 0;
  return sc;
}

size_t topk_partial_bytes(const Schedule& sc, int cta_group, int k) {
  const size_t b = static_cast<size_t>(sched_list_keys(sc, BM * cta_group, k)) * sizeof(uint64_t);
  return (b + 255) & ~static_cast<size_t>(255);
}

size_t topk_sync_bytes(const Schedule& sc) {
  return sc.sync_window > 0 ? static_cast<size_t>(sched_sync_counters(sc)) * sizeof(uint32_t) : 0;
}

size_t topk_workspace_bytes(const Schedule& sc, int cta_group, int k) {
  return topk_partial_bytes(sc, cta_group, k) + topk_sync_bytes(sc);
}

// Host-side walk of the schedule the kernel's warp roles follow (tests/test_abi.py, no GPU needed).
// Returns 0 if it is consistent, else a code naming the first broken invariant:
//   1 a tile computed twice or not at all (full: every (block, tile); symmetric: tile >= block, nothing else)
//   2 a run's list slot outside sched_slots(block)      3 two runs of a block share a list slot
//   4 a pacing counter outside the allocated array      5 arrivals on a counter differ from what waiters expect
//   6 a block without a flushed list for one of its slots
int schedule_selfcheck(const Schedule& sc, int units, int64_t* tiles_computed, int64_t* makespan_tiles, const std::vector<int>* owner_of_block) {
  const int nb = sc.mblocks, nt = sc.ntiles;
  std::vector<uint8_t> seen(static_cast<size_t>(nb) * nt, 0);
  std::vector<uint32_t> slot_used(static_cast<size_t>(nb) * std::max(sc.s_max, 1), 0);
  const int64_t n_counters = sched_sync_counters(sc);
  std::vector<int32_t> arrivals(static_cast<size_t>(std::max<int64_t>(n_counters, 1)), 0), expect(arrivals.size(), -1);
  int err = 0;
  int64_t computed = 0, makespan = 0;
  for (int u = 0; u < units && !err; ++u) {
    int64_t mine = 0;
    for_each_run(sc, u, [&](const Run& run) {
      if (err) return;
      if (run.mb < 0 || run.mb >= nb) { err = 2; return; }
      const int slot = sc.tab_runs ? run.slot - sc.tab_block_first[run.mb] : run.slot;   // list number inside the block
      if (slot < 0 || slot >= sched_slots(sc, run.mb) || slot >= sc.s_max) { err = 2; return; }
      if (slot_used[static_cast<size_t>(run.mb) * sc.s_max + slot]++) { err = 3; return; }
      if (run.nt_first < run.nt0 || run.nt_first > run.nt1) { err = 1; return; }
      for (int t = run.nt_first; t < run.nt1; ++t) {
        if (t >= nt || seen[static_cast<size_t>(run.mb) * nt + t]++) { err = 1; return; }
        ++computed; ++mine;
      }
      // every chunk ordinal of the run (computed or skipped) arrives once on the super-row's counters
      for (int64_t c = 0; sc.sync_window > 0 && c < static_cast<int64_t>(run.nt1 - run.nt0) * sc.cpt; ++c) {
        const int64_t at = run.sync_base + c;
        if (at < 0 || at >= n_counters) { err = 4; return; }
        ++arrivals[at];
        const int need = (c / sc.cpt) < run.short_len ? run.units_all : run.units_long;
        if (expect[at] >= 0 && expect[at] != need) { err = 5; return; }
        expect[at] = need;
      }
    });
    makespan = std::max(makespan, mine);
  }
  if (!err)
    for (size_t i = 0; i < arrivals.size() && !err; ++i)
      if (expect[i] >= 0 && arrivals[i] != expect[i]) err = 5;
  if (!err)
    for (int b = 0; b < nb && !err; ++b) {
      for (int t = 0; t < nt; ++t) {
        const bool mine = owner_of_block ? (*owner_of_block)[b] == sc.part_index : sched_owned(sc, b / sc.rm);
        const bool want = sc.sym ? (t >= b && mine) : true;
        if ((seen[static_cast<size_t>(b) * nt + t] != 0) != want) { err = 1; break; }
      }
      for (int sl = 0; sl < sched_slots(sc, b) && !err; ++sl)
        if (slot_used[static_cast<size_t>(b) * sc.s_max + sl] != 1) err = 6;
    }
  if (tiles_computed) *tiles_computed = computed;
  if (makespan_tiles) *makespan_tiles = makespan;
  return err;
}

// ------------------------------------------------------------------ run table (small symmetric sweeps)
bool sym_table_wanted(int64_t N) {
  const long force = env_long("SEMGATE_SYM_TABLE", -1);
  if (force == 0) return false;
  const int64_t nb = (N + BN - 1) / BN;
  return nb >= 2 && (force == 1 || nb <= 224);     // beyond ~57k keyframes the super-row formula is within a few % and paces
}

namespace {
struct Dealt { std::vector<RunEntry> runs; std::vector<int> owner; std::vector<int> lists_of_block; int64_t makespan = 0; int64_t tiles = 0; };
struct Group { int lo, hi; };   // blocks [lo, hi): a "super-row" of the table, its runs share database tiles

// Run length by remaining work ("guided self-scheduling"): long runs (few lists per keyframe for K3, few run starts)
// while there is plenty left, shorter and shorter ones towards the end so that the units finish together.  With a
// constant length of 8 the list scheduler below ends 3-5 tile-times above the ideal makespan; this way within one.
inline int guided_len(int64_t remaining, int units, int max_len) {
  const int64_t l = (2 * remaining + 3 * units - 1) / (3 * units);      // ceil(remaining / (1.5 * units))
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(max_len, l)));
}

// Cut the groups this part owns into runs (absolute column chunks per group, so that the blocks of a group cut at
// the same tiles), in the order group -> chunk -> block, and deal them in that order to the unit that is free first.
Dealt deal_runs(int nb, const std::vector<Group>& groups, const std::vector<int>& group_owner, int part, int units, int max_len) {
  Dealt d;
  d.lists_of_block.assign(nb, 0);
  using Slot = std::pair<int64_t, int>;   // (busy until, unit)
  std::priority_queue<Slot, std::vector<Slot>, std::greater<Slot>> free_at;
  for (int u = 0; u < units; ++u) free_at.push({0, u});
  int64_t remaining = 0;
  for (size_t g = 0; g < groups.size(); ++g)
    if (group_owner[g] == part)
      for (int b = groups[g].lo; b < groups[g].hi; ++b) remaining += nb - b;
  d.tiles = remaining;
  for (size_t g = 0; g < groups.size(); ++g) {
    if (group_owner[g] != part) continue;
    for (int t0g = groups[g].lo; t0g < nb;) {
      const int t1g = std::min(nb, t0g + guided_len(remaining, units, max_len));
      for (int b = groups[g].lo; b < groups[g].hi; ++b) {
        const int t0 = std::max(t0g, b);
        if (t1g <= t0) continue;
        Slot s = free_at.top();
        free_at.pop();
        d.runs.push_back(RunEntry{b, d.lists_of_block[b]++, t0, t1g});   // list number inside the block for now
        d.owner.push_back(s.second);
        s.first += t1g - t0;
        remaining -= t1g - t0;
        d.makespan = std::max(d.makespan, s.first);
        free_at.push(s);
      }
      t0g = t1g;
    }
  }
  return d;
}

// Groups of blocks for a sweep split over `parts` GPUs: contiguous, about parts * m of them with equal tile counts
// (tall at the bottom of the triangle, where a block has few tiles; at most rm_cap blocks), given to the parts largest
// first to whichever has the least so far.  Every part cuts the triangle the same way, so the ownership is a function.
void balanced_groups(int nb, int parts, int m, int rm_cap, std::vector<Group>* groups, std::vector<int>* owner) {
  groups->clear();
  const double total = 0.5 * nb * (nb + 1.0), target = total / (static_cast<double>(parts) * m);
  double acc = 0.0;
  int lo = 0;
  for (int b = 0; b < nb; ++b) {
    acc += nb - b;
    if (acc >= target * (groups->size() + 1) - 1e-9 || b + 1 - lo >= rm_cap) { groups->push_back(Group{lo, b + 1}); lo = b + 1; }
  }
  if (lo < nb) groups->push_back(Group{lo, nb});
  std::vector<int64_t> w(groups->size());
  std::vector<int> order(groups->size());
  for (size_t g = 0; g < groups->size(); ++g) {
    order[g] = static_cast<int>(g);
    const int64_t a = (*groups)[g].lo, e = (*groups)[g].hi;
    w[g] = (e - a) * nb - (e * (e - 1) - a * (a - 1)) / 2;
  }
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return w[x] > w[y]; });
  std::vector<int64_t> load(parts, 0);
  owner->assign(groups->size(), 0);
  for (int g : order) {
    const int pmin = static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin());
    (*owner)[g] = pmin;
    load[pmin] += w[g];
  }
}
}  // namespace

void build_sym_table(int64_t N, int d_pad, int sm_count, int part_index, int part_count, SymTable* out) {
  const int units = topk_units(2, sm_count);
  const int nb = static_cast<int>((N + BN - 1) / BN);
  part_count = std::max(part_count, 1);
  // Taller groups mean fewer distinct database tiles in flight (less DRAM traffic), as long as their query blocks stay
  // in L2; shorter runs mean a tighter finish but more lists per keyframe for K3 to merge.  Candidates are judged by
  // the busiest part's makespan (every part must cut the triangle the same way).
  const int64_t a_bytes = static_cast<int64_t>(BM) * 2 * d_pad * 2;
  const int64_t cap = env_long("SEMGATE_RM_CAP_MB", 40) << 20;
  const int rm_cap = static_cast<int>(std::max<int64_t>(6, std::min<int64_t>(24, cap / a_bytes)));
  int max_len = 8;
  if (env_long("SEMGATE_SYM_RUN", 0) > 0) max_len = static_cast<int>(env_long("SEMGATE_SYM_RUN", 0));   // A/B knob
  std::vector<Group> groups, best_groups;
  std::vector<int> owner, best_owner;
  int64_t best = -1;
  auto consider = [&]() {
    int64_t m = 0;
    for (int g = 0; g < part_count; ++g) m = std::max(m, deal_runs(nb, groups, owner, g, units, max_len).makespan);
    if (best < 0 || m < best) { best = m; best_groups = groups; best_owner = owner; }
  };
  if (part_count == 1) {
    const long forced_rm = env_long("SEMGATE_SYM_RM", 0);                                                   // A/B knob
    for (int rm : {24, 18, 16, 12, 10, 8, 6}) {                 // tallest first: ties go to the taller groups
      if (forced_rm > 0) rm = static_cast<int>(forced_rm);
      else if (rm != 6 && (rm > nb || rm > rm_cap)) continue;
      rm = std::max(1, std::min(rm, nb));
      groups.clear();
      for (int lo = 0; lo < nb; lo += rm) groups.push_back(Group{lo, std::min(nb, lo + rm)});
      owner.assign(groups.size(), 0);
      consider();
      if (forced_rm > 0) break;
    }
  } else {
    for (int m : {3, 4, 6}) {                                   // fewest groups first: ties go to the taller groups
      balanced_groups(nb, part_count, m, rm_cap, &groups, &owner);
      consider();
    }
  }
  Dealt d = deal_runs(nb, best_groups, best_owner, part_index, units, max_len);
  out->owner_of_block.assign(nb, 0);
  int rm_max = 1;
  for (size_t g = 0; g < best_groups.size(); ++g) {
    rm_max = std::max(rm_max, best_groups[g].hi - best_groups[g].lo);
    for (int b = best_groups[g].lo; b < best_groups[g].hi; ++b) out->owner_of_block[b] = best_owner[g];
  }
  out->block_first.assign(nb + 1, 0);
  for (int b = 0; b < nb; ++b) out->block_first[b + 1] = out->block_first[b] + d.lists_of_block[b];
  out->unit_begin.assign(units + 1, 0);
  for (int o : d.owner) ++out->unit_begin[o + 1];
  for (int u = 0; u < units; ++u) out->unit_begin[u + 1] += out->unit_begin[u];
  out->runs.resize(d.runs.size());
  std::vector<int> at(out->unit_begin.begin(), out->unit_begin.end() - 1);
  for (size_t i = 0; i < d.runs.size(); ++i) {      // stable: a unit keeps the order it was dealt
    RunEntry e = d.runs[i];
    e.list += out->block_first[e.mb];
    out->runs[at[d.owner[i]]++] = e;
  }
  out->makespan = d.makespan;
  out->tiles = d.tiles;
  Schedule& sc = out->sc;
  sc = Schedule{};
  sc.sym = 1;
  sc.mblocks = sc.ntiles = nb;
  sc.rm = rm_max;
  sc.n_full = nb / rm_max; sc.r_last = nb % rm_max;
  sc.part_index = part_index; sc.part_count = part_count;
  sc.s_max = std::max(1, *std::max_element(d.lists_of_block.begin(), d.lists_of_block.end()));
  sc.s_main = sc.s_last = sc.s_max;
  sc.a_resident = rm_max * a_bytes <= cap ? 1 : 0;
  sc.sync_window = 0;
  const int kblocks = d_pad / BK;
  sc.pace_kb = std::min(kblocks, 16);
  sc.cpt = (kblocks + sc.pace_kb - 1) / sc.pace_kb;
  sc.tab_lists = out->block_first[nb];
  sc.tab_tiles = d.tiles;
  sc.tab_runs = out->runs.data();                  // host pointers; the caller swaps in the device copies
  sc.tab_unit_begin = out->unit_begin.data();
  sc.tab_block_first = out->block_first.data();
}

// ------------------------------------------------------------------ tensor maps
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// row-major [rows, d_pad] bf16, box = [box_rows, 64 elements], 128-byte swizzle, OOB -> 0
static int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int d_pad, uint32_t box_rows) {
  auto enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(d_pad), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(d_pad) * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

static size_t smem_bytes_for(int cg, int stages, int kstride, bool sym, int sets) {
  const size_t stage = A_STAGE_BYTES + static_cast<size_t>(BN / std::min(cg, 2)) * BK * 2;
  return 1024 /*realign slack*/ + stages * stage + static_cast<size_t>(BM) * kstride * 8 +
         2 * BN * (sizeof(float) + sizeof(int32_t)) /*per-tile timestamp offsets + labels*/ +
         (sym ? 2 * BN * sizeof(float) + 64 : 0) /*column bounds + chunk minima*/ +
         (sets == 2 ? static_cast<size_t>(BM) * kstride * 8 + BM * 4 : 0) /*the second epilogue set's lists and counts*/ + 128 /*chunk stamp ranges*/ + 256 /*barriers + tmem slot*/;
}

// Symmetric sweep: candidate-buffer depth per keyframe and the layout of its state behind the pacing counters:
// [pacing counters | bound u32[N] | count u32[N] | flag (64 B)] (zeroed per launch, one memset) [buffers u64[N][cap]]
// depth: at least 4 k-lists; up to 1024 while the buffers of all keyframes stay within 1 GiB
int sym_capacity(int k, int64_t N) {
  int cap = k <= 32 ? 128 : 256;
  while (cap < 1024 && static_cast<int64_t>(2 * cap) * N * 8 <= (1ll << 30)) cap *= 2;
  return cap;
}
size_t sym_zeroed_bytes(int64_t N, size_t sync_bytes) { return ((sync_bytes + 8 * static_cast<size_t>(N) + 64) + 255) & ~static_cast<size_t>(255); }
size_t sym_state_bytes(int64_t N, int k, size_t sync_bytes) {
  return sym_zeroed_bytes(N, sync_bytes) + static_cast<size_t>(N) * sym_capacity(k, N) * sizeof(uint64_t);
}

int launch_gated_topk(const TopkLaunch& a, const Schedule& sc, uint64_t* partial, cudaStream_t st, int* launches) {
  if (a.Q <= 0 || a.N <= 0) return 0;
  const int cg = a.cta_group;
  CUtensorMap tq, tdb;
  int rc = make_tmap(&tq, a.q_bf16, a.Q, a.d_pad, BM);
  if (rc) return rc;
  rc = make_tmap(&tdb, a.db_bf16, a.N, a.d_pad, BN / cg);   // rows one CTA fetches per k-block
  if (rc) return rc;

  TopkParams p{};
  p.Q = static_cast<int>(a.Q);
  p.N = static_cast<int>(a.N);
  p.kblocks = a.d_pad / BK;
  p.k = a.k;
  p.kstride = a.k | 1;
  p.threshold = a.threshold;
  p.use_time = (a.q_ts != nullptr && a.db_ts != nullptr) ? 1 : 0;
  p.window_skip = static_cast<int>(env_long("SEMGATE_WINDOW_SKIP", 1)) != 0 ? 1 : 0;
  p.gap = a.gap;
  {   // fp32 neighbours of the window length for the pre-test: gap_lo <= gap <= gap_hi
    const float g = static_cast<float>(a.gap);
    p.gap_lo = static_cast<double>(g) > a.gap ? nextafterf(g, -INFINITY) : g;
    p.gap_hi = static_cast<double>(g) < a.gap ? nextafterf(g, INFINITY) : g;
  }
  p.max_floor_diff = a.max_floor_diff;
  p.gate_mode = a.gate_mode;
  p.db_index_offset = a.db_index_offset;
  p.q_ts = a.q_ts; p.db_ts = a.db_ts; p.q_floor = a.q_floor; p.db_floor = a.db_floor;
  p.partial = partial;
  p.sc = sc;
  p.dense = a.dense; p.dense_ld = a.dense_ld;
  if (a.dense != nullptr) p.sc.sync_window = 0;   // no workspace in dense mode
  p.sync = nullptr;
  p.run_if = a.run_if;
  p.ceil_keys = a.ceil_keys; p.ceil_stride = a.ceil_stride;
  p.clk = a.clk;
  const bool sym = sc.sym != 0;
  if (sym) {
    if (cg != 2 || a.Q != a.N || a.q_bf16 != a.db_bf16 || a.dense != nullptr || a.state == nullptr) return static_cast<int>(cudaErrorInvalidValue);
    // [pacing counters | bounds | counts | flag] zeroed together, candidate buffers behind
    char* z = static_cast<char*>(a.state);
    const size_t sync_bytes = topk_sync_bytes(sc);
    p.sync = sync_bytes ? reinterpret_cast<uint32_t*>(z) : nullptr;
    p.sym_bound = reinterpret_cast<uint32_t*>(z + sync_bytes);
    p.sym_cnt = p.sym_bound + a.N;
    p.sym_flag = p.sym_cnt + a.N;
    p.sym_cap = sym_capacity(a.k, a.N);
    p.sym_ovf = reinterpret_cast<uint64_t*>(z + sym_zeroed_bytes(a.N, sync_bytes));
    cudaError_t me = cudaMemsetAsync(z, 0, sym_zeroed_bytes(a.N, sync_bytes), st);
    if (me != cudaSuccess) return static_cast<int>(me);
    if (launches) ++*launches;
  } else if (p.sc.sync_window > 0) {
    // the pacing counters live behind the partial lists in the caller's workspace (or where the caller says)
    p.sync = a.state ? static_cast<uint32_t*>(a.state)
                     : reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(partial) + topk_partial_bytes(sc, cg, a.k));
    cudaError_t me = cudaMemsetAsync(p.sync, 0, topk_sync_bytes(sc), st);
    if (me != cudaSuccess) return static_cast<int>(me);
    if (launches) ++*launches;
  }
  // query blocks are re-read from L2 for every database tile: keep them; database tiles stream through
  const int hint = static_cast<int>(env_long("SEMGATE_L2_HINT", 2));
  p.policy_q = ((hint == 1 || hint == 2) && sc.a_resident) ? ptx::kL2EvictLast : ptx::kL2EvictNormal;
  // (run table: the units that share a database tile are not in lock-step, evict-first would throw a tile out
  //  before its last reader arrives -- measured 1.337 -> 1.290 ms at config 2 without it)
  p.policy_db = (hint >= 2 && sc.a_resident && sc.tab_runs == nullptr) ? ptx::kL2EvictFirst : ptx::kL2EvictNormal;

  // Two epilogue sets (plain list-building sweeps; each set keeps lists of its own, folded together when a run ends) wherever
  // the epilogue can be what bounds the kernel -- descriptors up to 4096 elements -- and the second set's lists still leave
  // a usable stage ring.  Measured (20 launches, same box, one set -> two): 5k x 512-d 52 -> 42 us, 20k x 512-d 404 -> 286 us,
  // on a sequence-like input with dense hits 5k x 512-d 111 -> 79 us, 20k x 2048-d 1.95 -> 1.67 ms; never slower.
  // SEMGATE_EPI_SETS=1 / 2 pins it.
  auto ring_depth = [&](int nsets) {
    int st = kMaxStages;
    while (st > 2 && smem_bytes_for(cg, st, p.kstride, sym, nsets) > kSmemLimit) --st;
    return smem_bytes_for(cg, st, p.kstride, sym, nsets) <= kSmemLimit ? st : 0;     // 0: does not fit at all
  };
  const bool sets_possible = cg != 4 && a.dense == nullptr;
  // (symmetric sweeps: run-table schedules only, i.e. up to ~57k keyframes -- 20k x 1024-d 815 -> 581 us, 20k x 4096-d with dense
  //  hits 2.07 -> 1.70 ms, the benchmark's config 2 unchanged (1.063 vs 1.060 ms); the long super-row sweeps are MMA-bound
  //  for seconds and keep the sixth stage: 1M x 4096-d 2.854 s with one set, 2.87 s with two)
  int sets = (sets_possible && (!sym || sc.tab_runs != nullptr) && p.kblocks <= 64 && ring_depth(2) >= (p.kblocks <= 16 ? 3 : 5)) ? 2 : 1;
  if (const char* e = getenv("SEMGATE_EPI_SETS")) { const int v = atoi(e); if (v == 1 || (v == 2 && sets_possible && ring_depth(2) >= 2)) sets = v; }
  // deepest ring that fits the 227 KB per-CTA limit
  int stages = std::min(ring_depth(sets), std::max(2, p.kblocks));
  p.stages = stages;
  const size_t smem = smem_bytes_for(cg, stages, p.kstride, sym, sets);

  const int units = topk_units(cg, a.sm_count);
  // only units that receive work in some super-row need to exist
  int used = 0;
  if (sc.tab_runs != nullptr) used = units;
  else if (sc.n_full > 0) used = std::max(used, sc.rm * std::min(sc.s_main, sc.ntiles));
  if (sc.r_last > 0) used = std::max(used, sc.r_last * std::min(sc.s_last, sym ? sc.r_last : sc.ntiles));
  used = std::min(std::max(used, 1), units);

  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(used * cg));
  cfg.blockDim = dim3(64 + 128 * sets);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cg); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = cg > 1 ? 1 : 0;
  cudaError_t e;
  auto launch = [&](auto kernel) -> cudaError_t {
    cudaError_t ee = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (ee != cudaSuccess) return ee;
    return cudaLaunchKernelEx(&cfg, kernel, tq, tdb, p);
  };
  if (sym) e = sets == 2 ? launch(gated_topk_kernel<2, 1, true, 2>) : launch(gated_topk_kernel<2, 1, true>);
  else if (cg == 4) e = launch(gated_topk_kernel<2, 2>);
  else if (cg == 2) e = sets == 2 ? launch(gated_topk_kernel<2, 1, false, 2>) : launch(gated_topk_kernel<2, 1>);
  else e = sets == 2 ? launch(gated_topk_kernel<1, 1, false, 2>) : launch(gated_topk_kernel<1, 1>);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (launches) ++*launches;
  return 0;
}

}  // namespace semgate
