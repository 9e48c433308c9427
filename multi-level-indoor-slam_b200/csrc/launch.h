// Internal host-side launch interface between the translation units of libsemgate.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <vector>

#include "common.cuh"

namespace semgate {

struct TopkLaunch {
  const void* q_bf16; int64_t Q;
  const void* db_bf16; int64_t N;
  int d_pad;
  const double* q_ts; const double* db_ts;      // both null -> no temporal mask
  const int32_t* q_floor; const int32_t* db_floor;
  float threshold; double gap; int k;
  int max_floor_diff; int gate_mode;
  uint32_t db_index_offset;
  int cta_group;                                  // CTAs per schedule unit: 1, 2 (cta_group::2 pair) or 4 (two pairs, multicast)
  int sm_count;
  float* dense; int64_t dense_ld;                 // non-null: dense fp32 similarity output instead of lists
  void* state;                                    // symmetric schedule: sym_state_bytes() of device memory (pacing counters,
                                                  // bounds, counts, flag, candidate buffers); else optional home of the
                                                  // pacing counters (null: behind the partial lists)
  const uint32_t* run_if;                         // non-null: device flag, the launch is a no-op unless it is non-zero
  unsigned long long* clk;                        // non-null: per-CTA {globaltimer, clock64} at entry / exit (4 words per CTA)
  const uint64_t* ceil_keys; int64_t ceil_stride; // non-null: row r only admits keys below ceil_keys[r * ceil_stride] (passes of k > 64)
};

// Tile schedule for a Q x N problem on `units` CTAs (1), CTA pairs (2) or two-pair clusters (4).
int topk_units(int cta_group, int sm_count);
Schedule make_schedule(int64_t Q, int64_t N, int d_pad, int cta_group, int sm_count, bool sym = false, int part_index = 0,
                       int part_count = 1);
// symmetric sweep (queries == database): state behind the partial lists, see gated_topk.cu
int sym_capacity(int k, int64_t N);
size_t topk_sync_bytes(const Schedule& sc);
size_t sym_zeroed_bytes(int64_t N, size_t sync_bytes);
size_t sym_state_bytes(int64_t N, int k, size_t sync_bytes);
int schedule_selfcheck(const Schedule& sc, int units, int64_t* tiles_computed, int64_t* makespan_tiles,
                       const std::vector<int>* owner_of_block = nullptr);   // run tables: which part owns each block
// Run table of a small symmetric sweep (see Schedule::tab_runs): built on the host; `sc` points into the vectors.
struct SymTable {
  Schedule sc;
  std::vector<RunEntry> runs;
  std::vector<int> unit_begin, block_first;
  std::vector<int> owner_of_block;   // part that computes each block's rows (the same function on every part)
  int64_t makespan;                  // tile-times of the busiest unit
  int64_t tiles;                     // tiles this part computes
};
bool sym_table_wanted(int64_t N);
void build_sym_table(int64_t N, int d_pad, int sm_count, int part_index, int part_count, SymTable* out);
size_t topk_partial_bytes(const Schedule& sc, int cta_group, int k);      // candidate-key lists (256-byte multiple)
size_t topk_workspace_bytes(const Schedule& sc, int cta_group, int k);    // lists + pacing counters

// K2: fills `partial` ([rows_padded][s_max][k] keys; the workspace also holds the pacing counters).
// Returns cudaError as int.
int launch_gated_topk(const TopkLaunch& a, const Schedule& sc, uint64_t* partial, cudaStream_t st,
                      int* launches);

// K6: streaming query (<= kStreamMaxQ query rows): bandwidth-built GEMV form of K2, one k-list per (query, block)
constexpr int kStreamMaxQ = 4;
bool stream_query_fits(int64_t Q, int d_pad, int k);
size_t stream_query_workspace_bytes(int64_t Q, int k, int sm_count);
int stream_query_lists(int sm_count);
int launch_stream_query(const TopkLaunch& a, uint64_t* partial /*[Q][lists][k]*/, cudaStream_t st);

// K1
int launch_normalize_cast(const float* x, int64_t n, int d, int64_t ld, void* out_bf16, int d_pad, cudaStream_t st);
int launch_normalize_cast_any(const void* x, int dtype /*SEMGATE_DTYPE_**/, int64_t n, int d, int64_t ld, void* out_bf16, int d_pad,
                              cudaStream_t st);

// K3: merge `n_lists` candidate lists per row into one sorted list.
struct MergeLaunch {
  const uint64_t* keys_in;
  int64_t Q; int k;
  int64_t row_stride, list_stride;   // in keys
  int n_lists;                       // >=0: constant; <0: per m-block from `sc` (offsets too: sched_list_offset)
  Schedule sc; int rows_per_mblock;
  const uint64_t* const* list_ptrs;  // optional device array of n_lists pointers: list g of row r = list_ptrs[g] + r*k
                                     // (per-GPU lists read in place over NVLink peer memory; keys_in unused)
  const uint64_t* seed_keys;         // optional [Q,k]: one more list per row (may alias keys_out)
  // symmetric sweep: while *sym_flag == 0 the lists follow `sc_sym` and row r also owns the first
  // min(sym_cnt[r], sym_cap) keys of sym_ovf + r * sym_cap; otherwise (overflow: the full sweep ran) `sc` holds
  const uint32_t* sym_flag; const uint32_t* sym_cnt; const uint64_t* sym_ovf; int sym_cap;
  uint32_t* sym_flag_copy;           // optional: *sym_flag is copied here (memory that outlives the sweep's workspace)
  int sym_force;                     // 1: no full sweep stands behind (one part of a multi-GPU sweep): `sc_sym` always holds
  Schedule sc_sym;
  // outputs (any may be null)
  uint64_t* keys_out;                // [Q,k] sorted descending, 0 padded
  float* scores; int32_t* idx; uint8_t* valid; int32_t* count;
  int64_t out_stride; int out_col;   // out_stride > 0: the [Q,k] outputs are columns out_col.. of rows out_stride wide
  int count_add;                     //   (one pass of a k > 64 sweep); count_add: count[row] += instead of =
  const int32_t* q_floor; const int32_t* db_floor; int64_t floor_index_offset;  // valid = gate(q_floor[row], db_floor[idx - off])
  int64_t floor_n;                   // > 0: labels in db_floor (an index outside is flagged invalid instead of read)
  int max_floor_diff;
  int64_t row_offset;                // merge rows [row_offset, row_offset + Q) of the input lists / q_floor; outputs are [Q, k]
  int64_t flag_offset;               // with list_ptrs and any_flag_out: word (u32) at list_ptrs[g] + flag_offset (in keys) is
  uint32_t* any_flag_out;            //   GPU g's overflow flag; their OR is written to *any_flag_out
  // one-call sweep (semgate_find_loop_closures_device): this launch also zeroes `zero_words` 64-bit words at `zero_ptr`
  // (the compaction's look-back state, so that no memset node sits between K3 and K4), and `pdl` launches the kernel as a
  // programmatic dependent of the kernel before it (its launch overlaps that kernel's tail; griddepcontrol.wait at entry)
  unsigned long long* zero_ptr; int zero_words; int pdl;
};
int launch_merge_topk(const MergeLaunch& a, cudaStream_t st);
void set_merge_dense(int on);   // 0: dense explicit lists take the general network kernel too (A/B, tests); process-wide

// K4: padded [Q,k] lists -> flat candidate arrays (query asc, score desc)
size_t compact_workspace_bytes(int64_t Q);
int compact_state_words(int64_t Q);   // 64-bit words of the workspace that must be zero when the kernel starts
int launch_compact(const float* scores, const int32_t* idx, const uint8_t* valid, const int32_t* count, int64_t Q, int k,
                   bool valid_only, int64_t q_offset /*added to the emitted query indices*/, int32_t* out_q, int32_t* out_m,
                   float* out_s, uint8_t* out_v, int64_t* out_total, void* workspace, cudaStream_t st, bool state_zeroed = false,
                   bool pdl = false);

// get_statistics (place_recognition.py:913-933) on the device: out[4] = total, valid, sum(sim), sum(valid sim), fp64
size_t stats_workspace_bytes();
int launch_candidate_stats(const float* sim, const uint8_t* valid, const int64_t* total_dev, int64_t M, void* workspace,
                           double* out, cudaStream_t st);

// floor gate over explicit candidate pairs (loop_closure_gate.py:105-126)
int launch_gate_candidates(const int32_t* floors, int64_t n_floors, const int32_t* q_idx, const int32_t* m_idx, int64_t M,
                           int max_floor_diff, uint8_t* out_valid, unsigned long long* counts /*[3]: accepted, rejected, bad index*/,
                           cudaStream_t st);

// spatial radius join (orb_slam3_integration.py:167-217): count+scan, then fill
size_t spatial_workspace_bytes(int64_t n);
int launch_spatial_count(const double* pos, int64_t n, double radius, int64_t gap, void* workspace, int64_t* total,
                         cudaStream_t st);
int launch_spatial_fill(const double* pos, int64_t n, double radius, int64_t gap, const void* workspace, int32_t* out_i,
                        int32_t* out_j, double* out_dist, int64_t capacity, cudaStream_t st);

// K5: CricaVPR cross-correlation score per candidate pair + per-query selection (place_recognition.py:669-757)
int launch_rerank(const void* feats_bf16, int n_feat, int P, int dl_pad, const int32_t* q_idx, const int32_t* m_idx,
                  const float* global_sim, int64_t M, float* out_cross, float* out_combined, int sm_count,
                  cudaStream_t st);
int launch_rerank_select(const int32_t* cand_idx, const float* combined, const int32_t* count, int64_t Q, int kc, int top_k,
                         int32_t* out_idx, float* out_score, int32_t* out_count, cudaStream_t st);

}  // namespace semgate
