// K2 — fused similarity GEMM + temporal-exclusion / floor gate / threshold /
// running top-k, sm_100a.
//
// Replaces, for a whole batch of query keyframes at once, the reference's
//   sim_matrix = X_n X_n^T                       (place_recognition.py:190)
//   per row: mask |t_j - t_i| < gap              (place_recognition.py:882-885)
//            argsort()[::-1][:k]                 (place_recognition.py:888)
//            drop sim < threshold                (place_recognition.py:891)
// and the single-query form  DB_n q_n + mask + top-k (place_recognition.py:140-154).
// The Q x N similarity matrix never leaves the SM: each 128 x 256 fp32 tile is
// accumulated in TMEM by tcgen05.mma from TMA-staged bf16 tiles, read back by the
// epilogue warps with tcgen05.ld, filtered and folded into a per-row top-k list
// held in shared memory.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA
// issuer, warps 2-5 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31 = tile rows).
// CG = 1: one CTA computes a 128 x 256 tile.  CG = 2: a CTA pair (cta_group::2)
// computes 256 x 256; each CTA stages its own 128 query rows and half of the
// database tile, and keeps the accumulator rows of its own queries.
// MC = 2 (with CG = 2): a cluster of two CTA pairs computes 512 x 256; the pairs own
// adjacent query blocks and share the database tile, every CTA fetches a quarter of it
// and TMA-multicasts it into the CTA of the other pair that needs it, which cuts the
// L2 -> SM traffic per FLOP by a quarter (the fused kernel sits at the L2 slice
// throughput at boost clocks).  The two pairs advance in lock-step through the stage ring.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace semgate {

constexpr int BM = 128;          // query rows per CTA (= TMEM lanes)
constexpr int BN = 256;          // database rows per tile (= TMEM columns per accumulator)
constexpr int BK = 64;           // bf16 per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;   // 16 KiB
constexpr uint64_t kPaceGiveUpNs = 2000000ull;    // 2 ms

struct TopkParams {
  int Q;                 // query rows
  int N;                 // database rows visible to this launch (local shard)
  int kblocks;           // padded descriptor length / BK
  int k;                 // list length (<= kMaxK)
  int kstride;           // smem list row stride in keys (odd -> conflict-free)
  int stages;            // smem ring depth
  float threshold;       // keep s >= threshold (compared in fp32, like numpy's weak-scalar rule)
  int use_time;          // 0: no temporal mask (query(timestamp=None), place_recognition.py:144)
  int window_skip;       // 1: chunks that lie wholly inside a row's exclusion window skip the slow path (A/B knob, default on)
  double gap;            // min_time_gap
  float gap_lo, gap_hi;  // fp32 neighbours of gap: gap_lo <= gap <= gap_hi (window pre-test, see window_excluded)
  int max_floor_diff;    // -1 off, 0 strict, 1 non-strict
  int gate_mode;         // 0 flag (reference order), 1 mask (exclude cross-floor before top-k)
  uint32_t db_index_offset;   // global index of local database row 0 (multi-GPU shards)
  const double* q_ts;
  const double* db_ts;
  const int32_t* q_floor;
  const int32_t* db_floor;
  uint64_t* partial;     // candidate keys, one k-entry list per (row, run): see sched_list_offset()
  float* dense;          // non-null: write the raw fp32 similarity tiles to dense[row * dense_ld + col] instead of
  int64_t dense_ld;      //           building lists (compute_all_pairwise_similarities, place_recognition.py:179-190)
  uint32_t* sync;        // pacing counters (zeroed per launch) or null: see Schedule::sync_window
  uint64_t policy_q;     // L2 eviction priority of the query-block loads (re-read once per database tile)
  uint64_t policy_db;    // ... of the streamed database tiles
  // symmetric sweep (SYM kernel): column-direction state, see the kernel's epilogue
  uint32_t* sym_bound;   // [N] admission bound of every keyframe as a query, ordered-score image; 0 = none published yet
  uint32_t* sym_cnt;     // [N] column-direction candidates appended so far
  uint32_t* sym_flag;    // != 0: some keyframe overflowed its candidate buffer -> the full sweep takes over
  uint64_t* sym_ovf;     // [N][sym_cap] appended candidate keys
  int sym_cap;
  const uint32_t* run_if;   // non-null: the whole launch is a no-op unless *run_if != 0 (the full sweep behind a symmetric one)
  // k > 64 runs as several sweeps: pass p only admits keys BELOW the last key pass p-1 kept for the row
  unsigned long long* clk;     // non-null: every CTA records {globaltimer, clock64} at entry and exit (4 words per CTA): the SM
                               // clock the kernel really ran at (a 2 ms NVML sample cannot resolve a 1 ms kernel)
  const uint64_t* ceil_keys;   // non-null: row r admits key < ceil_keys[r * ceil_stride] only (0: nothing is left for it)
  int64_t ceil_stride;
  Schedule sc;
};

// v[i] for a lane-varying i without a local-memory array: five levels of selects
__device__ __forceinline__ uint32_t pick32(const uint32_t (&v)[32], int i) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = (i & 1) ? v[2 * j + 1] : v[2 * j];
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = (i & 2) ? a[2 * j + 1] : a[2 * j];
#pragma unroll
  for (int j = 0; j < 4; ++j) c[j] = (i & 4) ? b[2 * j + 1] : b[2 * j];
#pragma unroll
  for (int j = 0; j < 2; ++j) d[j] = (i & 8) ? c[2 * j + 1] : c[2 * j];
  return (i & 16) ? d[1] : d[0];
}

// SYM (CTA pairs only): the queries are the database.  Only tiles on or above the block diagonal are
// computed; a tile right of the diagonal is read twice by the epilogue:
//   row direction     (as always) rows = queries, columns = candidates -> the thread's shared-memory list;
//   column direction  columns = queries, rows = candidates.  The owner of a query block publishes each
//     row's admission bound (k-th score once its list is full, monotone, atomicMax on the ordered image) in
//     sym_bound; whoever computes a tile stages the 256 column bounds (threshold if none yet) in shared
//     memory, compares every score of a column with that column's bound and appends the few that pass
//     to the keyframe's candidate buffer in HBM (sym_ovf, one atomic slot counter per keyframe).  A stale
//     bound only lets more candidates through: a bound is the k-th score of a subset of a keyframe's
//     admissible candidates, hence never above its final k-th score.  K3 merges buffer and lists.
//     A buffer that overflows (thresholds that admit most of the database) raises sym_flag, and the
//     full sweep launched right behind (run_if) redoes the job; nothing is lost, only time.
// Temporal exclusion without fp64 in the common case.  The reference predicate is fp64 `|t_db - t_q| < gap`
// (place_recognition.py:884); fp64 is slow on this part (43 % of the epilogue's stall samples at config 2 sat on its
// five instructions).  Both stamps are kept as fp32 offsets from a per-launch base, a = fl32(fl64(t_db - base)),
// b likewise: |(a - b) - (t_db - t_q)| <= 2^-24 (|a| + |b|) (1 + 2^-23), and the fp32 subtraction adds 2^-24 |a - b|.
// With the margin m = 2^-22 (|a| + |b| + d) (four times that bound, so its own fp32 rounding does not matter):
//   d > gap_hi + m  =>  |t_db - t_q| > gap by far more than an fp64 ulp  =>  not excluded,
//   d < gap_lo - m  =>  excluded,
// and only in between (pairs within ~1e-6 relative of the window edge, NaN / inf stamps) the exact fp64 test runs on the
// stamps themselves.  Decisions are bit-identical to the fp64 test (tests/test_window_pretest_model.py emulates this
// arithmetic in numpy float32 against the fp64 predicate on adversarial stamps).
__device__ __forceinline__ bool window_excluded(float a, float b, float gap_lo, float gap_hi, const double* t_db_ptr,
                                                double t_q, double gap) {
  const float d = fabsf(a - b);
  const float m = (fabsf(a) + fabsf(b) + d) * 0x1p-22f + 1e-30f;
  if (d > gap_hi + m) return false;
  if (d < gap_lo - m) return true;
  return time_excluded(__ldg(t_db_ptr), t_q, gap);
}

template <int CG, int MC, bool SYM = false, int SETS = 1>
__global__ void __launch_bounds__(64 + 128 * SETS, 1)
gated_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_db,
                  const TopkParams p) {
  static_assert(MC == 1 || (MC == 2 && CG == 2), "multicast needs CTA pairs");
  static_assert(!SYM || (CG == 2 && MC == 1), "symmetric sweep: a query block must be one database tile");
  static_assert(SETS == 1 || (SETS == 2 && MC == 1), "two epilogue sets: not with two-pair clusters");
  if (p.run_if != nullptr && ptx::ld_relaxed_gpu(p.run_if) == 0u) return;   // grid-uniform, before any barrier
  if (p.clk != nullptr && threadIdx.x == 0) {
    p.clk[4 * blockIdx.x + 0] = ptx::globaltimer_ns();
    p.clk[4 * blockIdx.x + 1] = static_cast<unsigned long long>(clock64());
  }
  constexpr int CSIZE = CG * MC;                 // CTAs per cluster = per schedule unit
  constexpr uint32_t B_ROWS = BN / CG;           // database rows held by one CTA
  constexpr uint32_t B_LOAD_ROWS = B_ROWS / MC;  // ... of which it fetches this many itself
  constexpr uint32_t B_STAGE_BYTES = B_ROWS * BK * 2;
  constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr uint32_t kTmemCols = 512;   // two 256-column fp32 accumulators

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is only guaranteed 16-byte aligned by the ABI; realign for SWIZZLE_128B
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int stages = p.stages;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + static_cast<size_t>(stages) * A_STAGE_BYTES;
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(stages) * STAGE_BYTES);
  // per-tile copies of the database timestamps / floor labels (one buffer per accumulator): a hit reads
  // them from shared memory instead of paying a dependent global load per hit column
  float* ts_s = reinterpret_cast<float*>(lists + static_cast<size_t>(BM) * p.kstride);   // fp32 offsets from ts_base
  int32_t* fl_s = reinterpret_cast<int32_t*>(ts_s + 2 * BN);
  // symmetric sweep: the tile's column bounds and their minimum per 32-column chunk (one set per accumulator)
  float* bd_s = reinterpret_cast<float*>(fl_s + 2 * BN);
  float* bmin_s = bd_s + (SYM ? 2 * BN : 0);
  // SETS == 2: the second epilogue set keeps lists of its own, [BM][kstride] behind everything else, and tells the first
  // set how many keys each holds when a run ends
  uint64_t* lists2 = reinterpret_cast<uint64_t*>(bmin_s + (SYM ? 16 : 0));
  int* cnt2_s = reinterpret_cast<int*>(lists2 + (SETS == 2 ? static_cast<size_t>(BM) * p.kstride : 0));
  // smallest / largest staged stamp of every 32-column chunk (one set per accumulator): [2][8][2]
  float* trange_s = reinterpret_cast<float*>(cnt2_s + (SETS == 2 ? BM : 0));
  uint64_t* bars = reinterpret_cast<uint64_t*>(trange_s + 32);
  // barrier slots: full[kMaxStages] empty[kMaxStages] tmem_full[2] tmem_empty[2]
  const uint32_t bar_full = ptx::smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * kMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kMaxStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CSIZE > 1) ? ptx::cluster_ctarank() : 0;
  const uint32_t half = cta_rank & 1u;           // position inside the CTA pair
  const uint32_t pair = cta_rank >> 1;           // which pair of the cluster
  const uint32_t pair_leader = cta_rank & ~1u;   // cluster rank of this pair's MMA-issuing CTA
  const bool leader = half == 0;
  const int unit = blockIdx.x / CSIZE;

  if constexpr (CSIZE > 1) ptx::cluster_sync();   // peers must be resident before a pair-wide TMEM allocation

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_db);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, CG);      // CG=2: leader's expect_tx arrive + peer's remote arrive
      ptx::mbar_init(bar_empty + 8 * s, MC);     // one tcgen05.commit per pair that reads what lands here
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(bar_tfull + 8 * a, 1);      // one tcgen05.commit
      ptx::mbar_init(bar_tempty + 8 * a, 4 * CG * SETS); // one arrive per epilogue warp (of both CTAs)
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) ptx::tmem_alloc<CG>(ptx::smem_u32(tmem_slot), kTmemCols);
  ptx::tc_fence_before();
  if constexpr (CSIZE > 1) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const Schedule sc = p.sc;

  // ring / accumulator state (each role advances its own copy identically)
  uint32_t stage = 0, phase = 0, it = 0;

  if (warp == 0) {
    // ===================================================== TMA producer (whole warp loops, one lane issues)
    const int window = (p.sync != nullptr) ? sc.sync_window : 0;
    const int pace_kb = sc.pace_kb, cpt = sc.cpt;
    for_each_run(sc, unit, [&](const Run& run) {
      const int m0 = (run.mb * CSIZE + static_cast<int>(cta_rank)) * BM;
      uint32_t* const pace = p.sync + run.sync_base;
      // Pacing is a hint, never a dependency: a unit that has waited kPaceGiveUpNs for the others
      // (e.g. one that idles through the full super-rows and reaches the tail super-row seconds
      // before anybody else) stops pacing for the rest of the run and just streams.
      bool paced = window > 0;
      if (window > 0 && ptx::elect_one()) {
        // tiles left of the diagonal (symmetric sweep) are not computed; the other units of the
        // super-row still count on this one's chunk arrivals
        for (int c = 0; c < (run.nt_first - run.nt0) * cpt; ++c) ptx::red_add_relaxed_gpu(pace + c, 1u);
      }
      __syncwarp();
      for (int nt = run.nt_first; nt < run.nt1; ++nt) {
        // rows of the database tile this CTA fetches: its half of the tile, and with MC = 2 the
        // quarter of that half that its pair is responsible for
        const int n0 = nt * BN + static_cast<int>(half * B_ROWS + (MC == 2 ? pair * B_LOAD_ROWS : 0u));
        int chunk = (nt - run.nt0) * cpt;            // position of this unit in its run, in chunks
        int kb_next_chunk = 0;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          if (window > 0 && kb == kb_next_chunk) {
            if (kb > 0) {                              // previous chunk fully issued
              if (ptx::elect_one()) ptx::red_add_relaxed_gpu(pace + chunk, 1u);
              ++chunk;
            }
            kb_next_chunk += pace_kb;
            const int back = chunk - window;
            if (paced && back >= 0) {
              // stay within `window` chunks of the slowest unit of this super-row
              const uint32_t need = static_cast<uint32_t>((back / cpt) < run.short_len ? run.units_all : run.units_long) * CSIZE;
              uint32_t spins = 0;
              uint64_t t0 = 0;
              while (ptx::ld_relaxed_gpu(pace + back) < need) {
                ptx::nanosleep(100);
                if ((++spins & 0x3Fu) == 0) {
                  const uint64_t now = ptx::globaltimer_ns();
                  if (t0 == 0) t0 = now;
                  else if (now - t0 > kPaceGiveUpNs) { paced = false; break; }
                }
              }
            }
          }
          ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t dst_a = ptx::smem_u32(smem_a) + stage * A_STAGE_BYTES;
          const uint32_t dst_b = ptx::smem_u32(smem_b) + stage * B_STAGE_BYTES;
          const uint32_t fb = bar_full + 8 * stage;
          if (ptx::elect_one()) {
            if constexpr (CG == 1) {
              ptx::mbar_arrive_expect_tx(fb, STAGE_BYTES);
              ptx::tma_load_2d(dst_a, &tmap_q, fb, kb * BK, m0, p.policy_q);
              ptx::tma_load_2d(dst_b, &tmap_db, fb, kb * BK, n0, p.policy_db);
            } else if constexpr (MC == 1) {
              // both CTAs' bytes are accounted on the leader's barrier (its MMA reads both smems)
              const uint32_t fb_leader = ptx::mapa(fb, 0);
              if (leader) ptx::mbar_arrive_expect_tx(fb, 2 * STAGE_BYTES);
              ptx::tma_load_2d_cg2(dst_a, &tmap_q, fb_leader, kb * BK, m0, p.policy_q);
              ptx::tma_load_2d_cg2(dst_b, &tmap_db, fb_leader, kb * BK, n0, p.policy_db);
              if (!leader) ptx::mbar_arrive_cluster(fb, 0);
            } else {
              // two pairs: the query rows are private, the database quarter goes to this CTA and to the
              // CTA at the same position of the other pair; every destination's bytes are accounted on
              // its own pair leader's barrier (2 * STAGE_BYTES per pair: 2 query blocks + 4 quarters)
              const uint32_t fb_leader = ptx::pair_leader_addr(fb);
              if (leader) ptx::mbar_arrive_expect_tx(fb, 2 * STAGE_BYTES);
              ptx::tma_load_2d_cg2(dst_a, &tmap_q, fb_leader, kb * BK, m0, p.policy_q);
              ptx::tma_load_2d_cg2_mc(dst_b + pair * (B_LOAD_ROWS * BK * 2), &tmap_db, fb_leader, kb * BK, n0,
                                      static_cast<uint16_t>((1u << half) | (1u << (half + 2))), p.policy_db);
              if (!leader) ptx::mbar_arrive_cluster(fb, pair_leader);
            }
          }
          __syncwarp();
          if (++stage == static_cast<uint32_t>(stages)) { stage = 0; phase ^= 1; }
        }
        if (window > 0 && ptx::elect_one()) ptx::red_add_relaxed_gpu(pace + chunk, 1u);   // last chunk of the tile
        __syncwarp();
      }
    });
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(BM * CG, BN);
      const uint64_t adesc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_a));
      const uint64_t bdesc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_b));
      for_each_run(sc, unit, [&](const Run& run) {
        for (int nt = run.nt_first; nt < run.nt1; ++nt, ++it) {
          const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
          ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            ptx::mbar_wait(bar_full + 8 * stage, phase);
            ptx::tc_fence_after();
            // descriptors differ from stage 0's only in the (address >> 4) field
            const uint64_t adesc = adesc0 + static_cast<uint64_t>((stage * A_STAGE_BYTES) >> 4);
            const uint64_t bdesc = bdesc0 + static_cast<uint64_t>((stage * B_STAGE_BYTES) >> 4);
            if (ptx::elect_one()) {
#pragma unroll
              for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                // advance 32 bytes (16 bf16) inside the 128-byte swizzle row: +2 in the >>4 address field
                ptx::umma_bf16<CG>(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb | kk) != 0);
              }
              if constexpr (CG == 1) {
                ptx::umma_commit_cg1(bar_empty + 8 * stage);
                if (kb == p.kblocks - 1) ptx::umma_commit_cg1(bar_tfull + 8 * acc);
              } else {
                // the stage is refilled by every CTA of the cluster that multicasts into this pair;
                // the accumulator is read by this pair's epilogues only
                ptx::umma_commit_cg2_mc(bar_empty + 8 * stage, static_cast<uint16_t>((1u << CSIZE) - 1u));
                if (kb == p.kblocks - 1) ptx::umma_commit_cg2_mc(bar_tfull + 8 * acc, static_cast<uint16_t>(0b11u << pair_leader));
              }
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(stages)) { stage = 0; phase ^= 1; }
          }
        }
      });
      if constexpr (CG == 2) {
        // the peer's epilogue arrives remotely on our barriers: drain before teardown
        if (it > 0) {
          const uint32_t last = it - 1;
          ptx::mbar_wait(bar_tempty + 8 * (last & 1), (last >> 1) & 1);
        }
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue: gate + threshold + running top-k
    // SETS == 2 (short descriptors: the epilogue, not the MMAs, bounds the kernel -- one warp per scheduler runs a chunk's
    // ~235 dependent instructions at 0.17 IPC): warps 2..5 and 6..9 are two sets, a warp reads the TMEM lanes of quad
    // warp % 4, set s takes the 32-column chunks c = s, s + 2, ....  Each of the two threads of a row keeps a list of its own
    // (its columns' top k); when a run ends the first set's thread folds the second's keys into its list and flushes that.
    // (A first version shared one list under a per-row lock: 3-11 % faster on sparse hits, 1.4-1.7x SLOWER on a sequence with
    // dense ones -- both threads insert all the time there.)
    const int set = SETS == 2 ? ((warp - 2) >> 2) : 0;
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int et = ((warp - 2) & 3) * 32 + lane;      // 0..127 among the epilogue threads of a set
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    auto epi_sync = [&]() { if constexpr (SETS == 2) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 1, 128;" ::: "memory"); };
    RowList L;
    L.keys = (set == 0 ? lists : lists2) + static_cast<size_t>(row_in_tile) * p.kstride;
    const int k = p.k;
    const bool mask_mode = p.gate_mode == 1 && p.max_floor_diff >= 0 && p.q_floor != nullptr && p.db_floor != nullptr;
    const bool use_time = p.use_time != 0;
    const float pos_inf = __int_as_float(0x7f800000);
    const double ts_base = use_time ? __ldg(p.db_ts) : 0.0;   // any finite stamp of the launch: offsets stay small

    for_each_run(sc, unit, [&](const Run& run) {
      const int grow = (run.mb * CSIZE + static_cast<int>(cta_rank)) * BM + row_in_tile;   // global query row
      const bool row_live = grow < p.Q;
      uint64_t* slot = p.partial + (row_live ? sched_run_list_offset(sc, run.mb, run.slot, grow, BM * CSIZE, k) : 0);
      double tq = 0.0;
      float tq32 = 0.f;
      int32_t qf = kFloorNone;
      if (row_live) {
        if (use_time) { tq = p.q_ts[grow]; tq32 = static_cast<float>(tq - ts_base); }
        if (mask_mode) qf = p.q_floor[grow];
      }
      const uint64_t ceil_key = (p.ceil_keys != nullptr && row_live) ? __ldg(p.ceil_keys + static_cast<int64_t>(grow) * p.ceil_stride) : ~0ull;
      L.reset(row_live ? p.threshold : pos_inf);
      if constexpr (SETS == 2) epi_sync();            // the flush of the run before has read the second set's lists
      float published = __int_as_float(0xff800000);   // SYM: last bound this thread published for its row
      // SYM: one column-direction append in flight per thread.  The slot number comes back from a global atomic
      // (about a microsecond); it is only looked at when the thread's next candidate arrives or the run ends, so
      // the round trip overlaps the chunks in between instead of stalling the warp at every hit.
      int pend_col = -1;
      uint32_t pend_at = 0;
      uint64_t pend_key = 0;
      auto complete_append = [&]() {
        if (pend_col >= 0) {
          if (pend_at < static_cast<uint32_t>(p.sym_cap)) p.sym_ovf[static_cast<int64_t>(pend_col) * p.sym_cap + pend_at] = pend_key;
          else *reinterpret_cast<volatile uint32_t*>(p.sym_flag) = 1u;
          pend_col = -1;
        }
      };

      // Per-tile copies of the database timestamps / labels (and, SYM, the columns' admission bounds with
      // their minimum per 32-column chunk) in shared memory, one buffer per accumulator: every thread fetches
      // two columns into registers (stage_load) and publishes them behind one named barrier (stage_store).
      // One barrier per tile is enough: whoever overwrites buffer `a` has passed the barrier of the tile before,
      // which every epilogue warp only reaches after it finished the tile before that (the buffer's last reader).
      const bool need_stage = use_time || mask_mode || SYM;
      double st_ts[2] = {0.0, 0.0};
      bool st_in[2] = {false, false};
      int32_t st_fl[2] = {kFloorNone, kFloorNone};
      float st_bd[2] = {pos_inf, pos_inf};
      auto stage_load = [&](int tile, bool cols) {
        if (SETS == 2 && set != 0) return;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int col = tile * BN + et + 128 * jj;
          if (use_time) { st_in[jj] = col < p.N; st_ts[jj] = st_in[jj] ? __ldg(p.db_ts + col) : 0.0; }
          if (mask_mode) st_fl[jj] = col < p.N ? __ldg(p.db_floor + col) : kFloorNone;
          if constexpr (SYM) {
            float b = pos_inf;                        // beyond N (TMA zero fill) and on the diagonal: nothing passes
            if (cols && col < p.N) {
              const uint32_t g = ptx::ld_relaxed_gpu(p.sym_bound + col);
              b = g != 0u ? ordered_to_score(g) : p.threshold;
            }
            st_bd[jj] = b;
          }
        }
      };
      auto stage_store = [&](uint32_t a) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          if (SETS == 2 && set != 0) break;
          const int j = et + 128 * jj;
          if (use_time) {
            const float off = static_cast<float>(st_ts[jj] - ts_base);
            ts_s[a * BN + j] = off;
            // the chunk's stamp range (this warp's columns in this pass are exactly chunk j / 32): columns beyond N do not
            // count, a NaN / inf stamp makes the range infinite (such a chunk is never skipped)
            const bool fin = off - off == 0.0f;
            float lo = !st_in[jj] ? pos_inf : (fin ? off : -pos_inf);
            float hi = !st_in[jj] ? -pos_inf : (fin ? off : pos_inf);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
              hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            if (lane == 0) { trange_s[(a * 8 + (j >> 5)) * 2] = lo; trange_s[(a * 8 + (j >> 5)) * 2 + 1] = hi; }
          }
          if (mask_mode) fl_s[a * BN + j] = st_fl[jj];
          if constexpr (SYM) {
            bd_s[a * BN + j] = st_bd[jj];
            // the columns of this warp in this pass are exactly chunk j / 32
            const uint32_t mn = __reduce_min_sync(0xffffffffu, score_to_ordered(st_bd[jj]));
            if (lane == 0) bmin_s[a * 8 + (j >> 5)] = ordered_to_score(mn);
          }
        }
        epi_sync();
      };
      bool staged = false;

      for (int nt = run.nt_first; nt < run.nt1; ++nt, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        const int col_base = nt * BN;
        const bool do_cols = SYM && nt != run.mb;   // the diagonal tile holds both orientations of its pairs
        if (need_stage && !staged) {    // first tile of a run: nothing was fetched ahead
          stage_load(nt, do_cols);
          stage_store(acc);
        }
        ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
        ptx::tc_fence_after();
        // the next tile's stamps / labels / bounds: loads in flight while this tile's chunks are processed
        const bool ahead = need_stage && nt + 1 < run.nt1;
        if (ahead) stage_load(nt + 1, SYM && nt + 1 != run.mb);
        const uint32_t t_acc = t_lane + acc * BN;
#pragma unroll 1
        for (int c = set; c < BN / 32; c += SETS) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_acc + c * 32, v);
          ptx::tmem_wait_ld();
          if (p.dense != nullptr) {
            // dense output: this thread owns 32 consecutive columns of its row
            const int col0 = col_base + c * 32;
            if (row_live && col0 < p.N) {
              float* dst = p.dense + static_cast<int64_t>(grow) * p.dense_ld + col0;
              if (col0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                  *reinterpret_cast<uint4*>(dst + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.N) dst[i] = __uint_as_float(v[i]);
              }
            }
            continue;
          }
          float mx = __uint_as_float(v[0]);
#pragma unroll
          for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
          const float bound = L.f;
          // Temporal neighbours score high and are all thrown out by the window: in an all-pairs sweep every row meets a
          // few chunks that lie INSIDE its exclusion window, where each of its 32 columns would walk the slow path only to
          // fail the window test (config 1: half of all slow-path iterations, all on the units that own diagonal tiles).
          // With the chunk's stamp range [lo, hi]: every column's d = |a - b| <= dmax = max(hi - b, b - lo) (fp32 subtraction
          // is monotone) and its margin m <= M, so dmax < gap_lo - M is window_excluded()'s own "excluded" verdict for all 32.
          bool in_window = false;
          if (use_time && p.window_skip) {
            const float lo = trange_s[(acc * 8 + c) * 2], hi = trange_s[(acc * 8 + c) * 2 + 1];
            const float dmax = fmaxf(hi - tq32, tq32 - lo);
            const float M = (fmaxf(fabsf(lo), fabsf(hi)) + fabsf(tq32) + dmax) * 0x1p-22f + 1e-30f;
            in_window = dmax < p.gap_lo - M;             // false for NaN / inf stamps and empty ranges (dmax = -inf handled below)
            if (lo > hi) in_window = false;
          }
          const bool hit = mx >= bound && !in_window;
          if (__any_sync(0xffffffffu, hit)) {
            uint32_t m = 0;
            if (hit) {
#pragma unroll
              for (int i = 0; i < 32; ++i) m |= (__uint_as_float(v[i]) >= bound ? 1u : 0u) << i;
            }
            // one hit column of this thread's row
            auto admit = [&](int i, float s) {
              if (s >= L.f) {                               // the bound may have risen since the mask was built
                const int col = col_base + c * 32 + i;      // local database row
                if (col < p.N) {                            // TMA zero-fill beyond N must not score
                  bool ok = true;
                  if (use_time) ok = !window_excluded(ts_s[acc * BN + c * 32 + i], tq32, p.gap_lo, p.gap_hi, p.db_ts + col, tq, p.gap);
                  if (ok && mask_mode) ok = floor_ok(qf, fl_s[acc * BN + c * 32 + i], p.max_floor_diff);
                  if (ok) {
                    const uint64_t key = pack_key(s, static_cast<uint32_t>(col) + p.db_index_offset);
                    if (key < ceil_key) L.insert(key, k);
                  }
                }
              }
            };
            // every lane walks its own hit columns; the score comes out of the registers already loaded (a 31-select tree; a
            // second trip to TMEM per hit column would serialise the warp: measured 1.7x slower at 512-d, 5 % at 4096-d).
            // (Tried for dense hits -- a revisit on a real sequence, where the warp's 32 rows hit the same run of columns:
            // walking the union of the hit columns with a warp-uniform index and a jump table instead of the select tree.
            // Slower on every input, sparse and dense: 367 -> 388 us and 605 -> 646 us at 20k x 512-d.)
            while (m) {
              const int i = __ffs(m) - 1;
              m &= m - 1;
              admit(i, __uint_as_float(pick32(v, i)));
            }
            __syncwarp();
          }
          if constexpr (SYM) {
            // column direction: this thread's row is a candidate of the 32 keyframes that head the columns
            const bool maybe = do_cols && row_live && mx >= bmin_s[acc * 8 + c];
            if (__any_sync(0xffffffffu, maybe)) {
              uint32_t cm = 0;
              if (maybe) {
                const float4* bd4 = reinterpret_cast<const float4*>(bd_s + acc * BN + c * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 b = bd4[i];
                  cm |= (__uint_as_float(v[4 * i + 0]) >= b.x ? 1u : 0u) << (4 * i + 0);
                  cm |= (__uint_as_float(v[4 * i + 1]) >= b.y ? 1u : 0u) << (4 * i + 1);
                  cm |= (__uint_as_float(v[4 * i + 2]) >= b.z ? 1u : 0u) << (4 * i + 2);
                  cm |= (__uint_as_float(v[4 * i + 3]) >= b.w ? 1u : 0u) << (4 * i + 3);
                }
              }
              while (cm) {
                const int i = __ffs(cm) - 1;
                cm &= cm - 1;
                const float s = __uint_as_float(pick32(v, i));
                const int j = acc * BN + c * 32 + i;
                bool ok = true;
                if (use_time) ok = !window_excluded(ts_s[j], tq32, p.gap_lo, p.gap_hi, p.db_ts + col_base + c * 32 + i, tq, p.gap);   // |a - b| is symmetric
                if (ok && mask_mode) ok = floor_ok(fl_s[j], qf, p.max_floor_diff);
                if (ok) {
                  complete_append();
                  pend_col = col_base + c * 32 + i;                        // < N: columns beyond carry +inf bounds
                  pend_key = pack_key(s, static_cast<uint32_t>(grow) + p.db_index_offset);
                  pend_at = atomicAdd(p.sym_cnt + pend_col, 1u);
                }
              }
              __syncwarp();
            }
          }
        }
        if constexpr (SYM) {
          // publish this row's admission bound once it is a real k-th score (monotone)
          if (row_live && L.cnt == k && L.f > published) {
            atomicMax(p.sym_bound + grow, score_to_ordered(L.f));
            published = L.f;
          }
        }
        // accumulator drained: hand it back to the MMA issuer
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 1) ptx::mbar_arrive(bar_tempty + 8 * acc);
          else ptx::mbar_arrive_cluster(bar_tempty + 8 * acc, pair_leader);
        }
        if (ahead) stage_store(acc ^ 1u);     // the next tile uses the other accumulator and the other buffer
        staged = ahead;
      }

      if constexpr (SYM) complete_append();
      // flush this run's list (unsorted, packed at the front; empty slots are key 0)
      if constexpr (SETS == 2) {
        if (set == 1) cnt2_s[row_in_tile] = L.cnt;
        epi_sync();                                   // both sets are done with the run's last tile
        if (set == 0) {
          const uint64_t* other = lists2 + static_cast<size_t>(row_in_tile) * p.kstride;
          const int c2 = cnt2_s[row_in_tile];
          for (int i = 0; i < c2; ++i) {              // keys are unique: the two sets saw different columns
            const uint64_t key = other[i];
            if (L.cnt < k || key > L.min_key) L.insert(key, k);
          }
          if (row_live && p.dense == nullptr)
            for (int i = 0; i < k; ++i) slot[i] = i < L.cnt ? L.keys[i] : 0ull;
        }
      } else if (row_live && p.dense == nullptr) {
        for (int i = 0; i < k; ++i) slot[i] = i < L.cnt ? L.keys[i] : 0ull;
      }
    });
  }

  // ----------------------------------------------------------------- teardown
  ptx::tc_fence_before();
  if constexpr (CSIZE > 1) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1) ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
  if (p.clk != nullptr && threadIdx.x == 0) {
    p.clk[4 * blockIdx.x + 2] = ptx::globaltimer_ns();
    p.clk[4 * blockIdx.x + 3] = static_cast<unsigned long long>(clock64());
  }
}

}  // namespace semgate
