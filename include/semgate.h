/* libsemgate — C ABI of the B200-native gated loop-closure retrieval path.
 *
 * The reference (wadewilliamsw1234/Multi-level-Indoor-SLAM) is pure Python and has
 * no FFI of its own; its seam for this path is the method surface of
 * scripts/semantic_gating/place_recognition.py and loop_closure_gate.py.  Each
 * entry point below names the reference interface it replaces (file:line relative
 * to scripts/semantic_gating/).  INTEGRATION.md shows the ctypes stub a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no allocation of caller-visible memory inside;
 *   - every function returns 0 on success, a negative SEMGATE_E* code or a positive
 *     cudaError_t otherwise; semgate_last_error() gives a thread-local message;
 *   - "device" pointers must live on the handle's GPU; *_host entry points take
 *     host pointers (pinned or pageable) and do the copies themselves;
 *   - sm_100a only: semgate_create fails on any other compute capability.  There
 *     is no CPU or other-architecture fallback.
 *   - floor labels are int32; SEMGATE_FLOOR_NONE encodes Python's floor_label=None
 *     (place_recognition.py:78,898), which passes every gate.
 */
#ifndef SEMGATE_H_
#define SEMGATE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEMGATE_VERSION 200
#define SEMGATE_MAX_K 64          /* candidates per query one sweep keeps (shared-memory list) */
#define SEMGATE_MAX_K_TOTAL 1024   /* largest k: k > 64 runs as ceil(k / 64) sweeps, see semgate_gated_topk */
#define SEMGATE_FLOOR_NONE INT32_MIN

#define SEMGATE_EINVAL (-1)       /* bad argument */
#define SEMGATE_EARCH (-2)        /* device is not compute capability 10.x */
#define SEMGATE_ENOMEM (-3)       /* workspace / capacity too small */
#define SEMGATE_EDRIVER (-4)      /* cuTensorMapEncodeTiled unavailable or failed */
#define SEMGATE_EINDEX (-5)       /* candidate index outside the floor-label array */

#define SEMGATE_GATE_FLAG 0       /* reference order: top-k, then flag cross-floor (place_recognition.py:888-899) */
#define SEMGATE_GATE_MASK 1       /* exclude cross-floor columns before top-k */

typedef struct semgate_ctx* semgate_handle_t;
typedef void* semgate_stream_t;   /* cudaStream_t */

/* Parameters of one gated top-k sweep.
 * Mirrors the knobs of SemanticPlaceRecognition(similarity_threshold, min_time_gap)
 * (place_recognition.py:814-818), find_loop_closures(enable_floor_gating, k) (:851-853),
 * query(k, min_time_gap) (:117-121) and SemanticLoopClosureGate(strict_mode)
 * (loop_closure_gate.py:42-44). */
typedef struct semgate_topk_params {
  float similarity_threshold;   /* keep s >= threshold, compared in fp32; -INFINITY disables (query()) */
  double min_time_gap;          /* exclude |t_db - t_q| < gap (strict, fp64) */
  int32_t k;                    /* 1..SEMGATE_MAX_K_TOTAL; above SEMGATE_MAX_K: several sweeps (no accumulate / parts) */
  int32_t max_floor_diff;       /* -1 gating off; 0 strict; 1 non-strict */
  int32_t gate_mode;            /* SEMGATE_GATE_FLAG | SEMGATE_GATE_MASK */
  uint32_t db_index_offset;     /* global index of database row 0 (row-sharded multi-GPU) */
  int32_t cta_group;            /* 0 = handle default (auto: <= 4 query rows -> streaming kernel, Q >= 4096 -> CTA
                                   pairs, else single CTAs); 1 = single-CTA tiles; 2 = CTA-pair tiles (cta_group::2);
                                   4 = clusters of two pairs sharing a multicast database tile */
  int32_t accumulate;           /* 1: out_keys already holds each query's list over OTHER database rows (earlier
                                   sweeps of disjoint slices); merge this sweep into it in place */
  int32_t symmetric;            /* all-pairs sweeps (queries == database, place_recognition.py:190 computes X X^T):
                                   0 = auto: when q_bf16 == db_bf16, q_ts == db_ts, q_floor == db_floor, Q == N,
                                   db_index_offset == 0, the tiles are CTA pairs and the sweep is long enough to be
                                   tensor-bound (Q >= 8192, d_pad >= 1024; handle option "symmetric" = 1 drops the size
                                   rule), every similarity is computed once and gated in both directions (half the
                                   tensor work, same lists);
                                   1 = require it (SEMGATE_EINVAL if the arguments do not allow it); -1 = never */
  int32_t part_index;           /* a symmetric sweep split over the GPUs of a box: with part_count = G > 1 (needs  */
  int32_t part_count;           /* symmetric = 1) this call computes part part_index of the tile triangle -- every G-th
                                   group of query blocks, dealt out boustrophedon so the parts are equal -- and its
                                   lists hold each query's best candidates among the pairs of THAT part; merging the G
                                   parts' out_keys (semgate_merge_topk / _peers) gives the sweep's lists.  No full
                                   sweep stands behind a part: if semgate_last_sweep_mode reports 2 on ANY part, the
                                   caller redoes the sweep row-sharded (semgate/dist.py does).  0 or 1: whole sweep */
} semgate_topk_params;

int semgate_version(void);
const char* semgate_last_error(void);

/* ---- lifetime ---------------------------------------------------------- */
int semgate_create(semgate_handle_t* out, int device);
int semgate_destroy(semgate_handle_t h);
int semgate_device_info(semgate_handle_t h, int* sm_count, int* cc_major, int* cc_minor);
/* options: "cta_group" (0 auto | 1 | 2 | 4); "symmetric" (0 auto by size | 1 whenever the arguments allow |
 * -1 never: handle default for semgate_topk_params.symmetric == 0); "profile" (0|1): bracket every fused-kernel launch with CUDA
 * events on its own stream; "clock_probe" (0|1): see semgate_clock_probe_read; "k3_dense" (1|0, process-wide): lists
 * handed to semgate_merge_topk* as arrays take the dense-list merge kernel (default) or the general one (A/B, tests) */
int semgate_set_option(semgate_handle_t h, const char* name, int64_t value);
/* sum of the fused kernel's (K2) device durations since the last read, and how many
 * launches that covers; synchronises on the recorded events and resets them. */
int semgate_profile_read(semgate_handle_t h, double* total_ms, int64_t* n_launches);
/* option "clock_probe" (0|1): every CTA of the fused kernel records {globaltimer, clock64} at entry and exit; this
 * reads the last sweep's first K2 launch back: the median and minimum over its CTAs of cycles / nanosecond (the SM
 * clock the kernel really ran at: NVML samples every few ms cannot resolve a 1 ms kernel), the time from the first
 * CTA's entry to the last one's exit, and the number of CTAs that reported.  Synchronises the sweep's stream. */
int semgate_clock_probe_read(semgate_handle_t h, double* sm_mhz_median, double* sm_mhz_min, double* span_us, int32_t* n_ctas);
/* kernels launched through this handle since creation (bench.py's gpu_launches) */
int64_t semgate_launch_count(semgate_handle_t h);

/* How the last semgate_gated_topk on this handle ran: *out_mode = 0 full sweep, 1 symmetric sweep,
 * 2 symmetric sweep whose candidate buffers overflowed, so that the full sweep behind it produced the
 * result (reads a device flag the sweep's merge kernel copied into handle-owned memory: synchronises the stream
 * of that call; the call's workspace may already be reused or freed).  *out_tiles (may be NULL) = 256-row x
 * 256-column (CTA pairs; 128 x 256 for single-CTA tiles) similarity tiles its schedule computes; mode 2 ran both. */
int semgate_last_sweep_mode(semgate_handle_t h, int32_t* out_mode, int64_t* out_tiles);

/* The same overflow flag without a host round trip: *out_flag_dev (device uint32) = non-zero iff the last
 * semgate_gated_topk on this handle was a symmetric sweep (or one part of one) whose candidate buffers
 * overflowed; stream-ordered copy.  A multi-GPU caller all-reduces it and reads it once, after the merge is queued. */
int semgate_last_sweep_overflow(semgate_handle_t h, uint32_t* out_flag_dev, semgate_stream_t stream);

/* Testing aid, needs no device: walks the fused kernel's tile schedule for a Q x N sweep on the host exactly as
 * the kernel's warp roles do and checks its invariants (every tile computed once -- in a symmetric sweep every
 * tile on or above the block diagonal and nothing else --, list slots, pacing counters).
 * symmetric: 0 full sweep, 1 symmetric sweep by the super-row formula, 2 symmetric sweep by the run table
 * (what sweeps of up to ~57k keyframes use).
 * out_shape[8] = blocks, tiles, rm, s_main (run table: most lists of any block), r_last, s_last, pacing window,
 * query blocks L2-resident;
 * out_tiles[2] = tiles computed, longest unit's tile count (the makespan in tile-times). */
int semgate_schedule_check(int64_t Q, int64_t N, int32_t d_pad, int32_t cta_group, int32_t sm_count, int32_t symmetric,
                           int32_t part_index, int32_t part_count, int32_t* out_shape, int64_t* out_tiles);

/* descriptor length padded to the kernel's K granule (64) */
int semgate_pad_dim(int d);

/* ---- K1: row normalisation + bf16 cast ----------------------------------
 * replaces `desc_matrix / (norms + 1e-8)` (place_recognition.py:186-187) and the
 * per-query re-normalisation in _compute_similarity (:169-170).
 * x: device fp32 [n, d], row stride ld elements.  out: device bf16 [n, d_pad]. */
int semgate_normalize_cast(semgate_handle_t h, const float* x, int64_t n, int32_t d, int64_t ld, void* out_bf16,
                           int32_t d_pad, semgate_stream_t stream);

/* The same for descriptors that are already half precision (the reference runs its extractors on the GPU,
 * place_recognition.py:291-297; under autocast their output is fp16 / bf16 and `.cpu().numpy()` keeps that).
 * dtype: SEMGATE_DTYPE_*.  Elements are widened exactly and every sum runs in the same order as the fp32
 * form, so a half-precision row gives bit for bit what its fp32 image gives. */
#define SEMGATE_DTYPE_F32 0
#define SEMGATE_DTYPE_F16 1
#define SEMGATE_DTYPE_BF16 2
int semgate_normalize_cast_dtype(semgate_handle_t h, const void* x, int32_t dtype, int64_t n, int32_t d, int64_t ld,
                                 void* out_bf16, int32_t d_pad, semgate_stream_t stream);

/* ---- dense similarity matrix (interface parity only) ---------------------
 * replaces `desc_matrix_norm @ desc_matrix_norm.T` (place_recognition.py:190) and
 * `np.dot(database_norm, query_norm)` (:171) for callers that really want the scores
 * themselves: out[i * ld_out + j] = <q_i, db_j>, fp32, same tcgen05 main loop as K2
 * with a store epilogue.  The retrieval path never calls this. */
int semgate_similarity_matrix(semgate_handle_t h, const void* q_bf16, int64_t Q, const void* db_bf16, int64_t N,
                              int32_t d_pad, float* out, int64_t ld_out, semgate_stream_t stream);

/* ---- K2+K3: fused similarity / exclusion window / gate / threshold / top-k ----
 * replaces compute_all_pairwise_similarities + the per-row loop of
 * find_loop_closures (place_recognition.py:868-899) and, with one query row,
 * _compute_similarity + the mask + argsort of query() (:140-154).
 *   q_bf16 [Q,d_pad], db_bf16 [N,d_pad]    normalised bf16 rows (K1 output)
 *   q_ts/db_ts   fp64 or both NULL (no temporal mask, query(timestamp=None))
 *   q_floor/db_floor  int32 or NULL (no gating)
 *   workspace    >= semgate_topk_workspace_bytes(...)
 * A handle is not thread-safe; calls on one handle are issued in order.  For small all-pairs sweeps the
 * symmetric schedule is a host-built table, cached in the handle per tile count (ceil(N/256)) and uploaded with
 * cudaMemcpyAsync on the calling stream at its first sweep (that call allocates a few hundred KB of device memory;
 * semgate_topk_workspace_bytes only builds the host copy).
 * accumulate = 1 cannot produce out_valid when gating is on (SEMGATE_EINVAL): the seeded lists hold indices of other
 * database slices, outside db_floor; flag the final lists with semgate_merge_topk and the whole label array.
 * outputs, each [Q,k], any may be NULL:
 *   out_keys    packed candidates (for semgate_merge_topk across GPUs)
 *   out_scores  fp32 descending, -inf padded;  out_idx int32 global index, -1 padded
 *   out_valid   uint8 floor flag;  out_count int32 [Q] */
size_t semgate_topk_workspace_bytes(semgate_handle_t h, int64_t Q, int64_t N, int32_t d_pad, const semgate_topk_params* p);
int semgate_gated_topk(semgate_handle_t h, const void* q_bf16, int64_t Q, const void* db_bf16, int64_t N, int32_t d_pad,
                       const double* q_ts, const double* db_ts, const int32_t* q_floor, const int32_t* db_floor,
                       const semgate_topk_params* p, void* workspace, size_t workspace_bytes, uint64_t* out_keys,
                       float* out_scores, int32_t* out_idx, uint8_t* out_valid, int32_t* out_count,
                       semgate_stream_t stream);

/* ---- K3 alone: merge G candidate lists per query (after an all-gather of out_keys) ----
 * keys_in [G,Q,k] device; db_floor_all: floor labels of the WHOLE database (index = global). */
int semgate_merge_topk(semgate_handle_t h, const uint64_t* keys_in, int32_t G, int64_t Q, int32_t k,
                       const int32_t* q_floor, const int32_t* db_floor_all, int32_t max_floor_diff, uint64_t* out_keys,
                       float* out_scores, int32_t* out_idx, uint8_t* out_valid, int32_t* out_count,
                       semgate_stream_t stream);

/* Same merge with the G per-GPU lists read IN PLACE: peer_keys is a DEVICE array of G device
 * pointers, list g of query row r = peer_keys[g] + r*k.  With the other ranks' buffers mapped into
 * this process (CUDA IPC / torch symmetric memory) the kernel pulls them over NVLink while it
 * merges, so the all-gather and its G*Q*k gathered copy disappear (semgate/dist.py, exchange="peer").
 * The caller orders the peers' writes before this launch (a cross-GPU barrier on the stream). */
int semgate_merge_topk_peers(semgate_handle_t h, const uint64_t* const* peer_keys, int32_t G, int64_t Q, int32_t k,
                             const int32_t* q_floor, const int32_t* db_floor_all, int32_t max_floor_diff,
                             uint64_t* out_keys, float* out_scores, int32_t* out_idx, uint8_t* out_valid,
                             int32_t* out_count, semgate_stream_t stream);

/* The same merge for a SLICE of the query rows, with the peers' overflow flags folded in.  Every rank of a
 * multi-GPU sweep merges only its own rows [row_begin, row_begin + row_count) of the G per-GPU lists (each
 * [Q_total, k]; list g of row r = peer_keys[g] + r*k): the NVLink reads drop to 1/G of the replicated merge's.
 * Outputs are [row_count, k] / [row_count]; q_floor is indexed by the global row.  With out_any_flag != NULL
 * the kernel also reads one uint32 at peer_keys[g] + flag_offset (offset in 8-byte keys, >= Q_total*k: a word
 * behind every rank's key buffer, written there by semgate_last_sweep_overflow) for every g and stores their OR
 * in *out_any_flag (device): whether any rank's part of a split symmetric sweep overflowed, learnt without a
 * collective of its own.  replaces nothing in the reference (single process); semgate/dist.py is the caller. */
int semgate_merge_topk_peers_rows(semgate_handle_t h, const uint64_t* const* peer_keys, int32_t G, int64_t Q_total,
                                  int32_t k, int64_t row_begin, int64_t row_count, int64_t flag_offset,
                                  const int32_t* q_floor, const int32_t* db_floor_all, int32_t max_floor_diff,
                                  uint64_t* out_keys, float* out_scores, int32_t* out_idx, uint8_t* out_valid,
                                  int32_t* out_count, uint32_t* out_any_flag, semgate_stream_t stream);

/* ---- K4: candidate compaction ----------------------------------------------
 * replaces the PlaceMatch append loop (place_recognition.py:890-909): flat arrays
 * ordered (query ascending, score descending).  Outputs need capacity Q*k.
 * out_total: device int64.  workspace >= semgate_compact_workspace_bytes(Q). */
size_t semgate_compact_workspace_bytes(int64_t Q);
int semgate_compact(semgate_handle_t h, const float* scores, const int32_t* idx, const uint8_t* valid, const int32_t* count,
                    int64_t Q, int32_t k, int32_t* out_query_idx, int32_t* out_match_idx, float* out_similarity,
                    uint8_t* out_is_valid, int64_t* out_total, void* workspace, semgate_stream_t stream);

/* Same, emitting only the floor-consistent candidates (is_valid != 0): the list handed on to geometric
 * verification, which skips cross-floor pairs (geometric_verification.py:709). */
int semgate_compact_valid(semgate_handle_t h, const float* scores, const int32_t* idx, const uint8_t* valid, const int32_t* count,
                          int64_t Q, int32_t k, int32_t* out_query_idx, int32_t* out_match_idx, float* out_similarity,
                          uint8_t* out_is_valid, int64_t* out_total, void* workspace, semgate_stream_t stream);

/* Compaction of a SLICE of the query rows (a rank's share of a multi-GPU sweep, semgate_merge_topk_peers_rows):
 * the lists are [Q, k] for the rows query_index_offset .. +Q, the emitted query indices are global.
 * valid_only != 0: as semgate_compact_valid. */
int semgate_compact_rows(semgate_handle_t h, const float* scores, const int32_t* idx, const uint8_t* valid, const int32_t* count,
                         int64_t Q, int32_t k, int64_t query_index_offset, int32_t valid_only, int32_t* out_query_idx,
                         int32_t* out_match_idx, float* out_similarity, uint8_t* out_is_valid, int64_t* out_total,
                         void* workspace, semgate_stream_t stream);

/* ---- match statistics -------------------------------------------------------
 * replaces SemanticPlaceRecognition.get_statistics (place_recognition.py:913-933) for a
 * device-resident candidate list: out_stats (device, 4 doubles) = { total, valid, sum(similarity),
 * sum(similarity of valid) }, accumulated in fp64 in a scheduling-independent order.  The number
 * of candidates is *total_dev (device int64, e.g. semgate_compact's out_total) when given, else M.
 * workspace >= semgate_stats_workspace_bytes(). */
size_t semgate_stats_workspace_bytes(void);
int semgate_candidate_stats(semgate_handle_t h, const float* similarity, const uint8_t* is_valid, const int64_t* total_dev,
                            int64_t M, void* workspace, double* out_stats, semgate_stream_t stream);

/* ---- floor gate over explicit candidate pairs --------------------------------
 * replaces SemanticLoopClosureGate.gate_candidates (loop_closure_gate.py:105-126).
 * max_floor_diff: 0 = strict_mode True, 1 = strict_mode False.
 * out_counts: device uint64[3] = accepted, rejected_cross_floor, out-of-range indices. */
int semgate_gate_candidates(semgate_handle_t h, const int32_t* floor_labels, int64_t n_labels, const int32_t* query_idx,
                            const int32_t* match_idx, int64_t M, int32_t max_floor_diff, uint8_t* out_is_valid,
                            uint64_t* out_counts, semgate_stream_t stream);

/* ---- spatial-proximity candidate generator (radius join over poses) ----------------
 * replaces detect_loop_closure_candidates (orb_slam3_integration.py:167-217 and the
 * droid_slam / lego_loam twins): pairs i < j with ||p_i - p_j|| <= radius (fp64) and
 * j - i >= min_index_gap, sorted by (i, j).  positions: device fp64 [n,3].
 * Two phases: _count fills the workspace and *out_total (device int64); _fill writes the
 * first `capacity` pairs (out_dist may be NULL). */
size_t semgate_spatial_workspace_bytes(int64_t n);
int semgate_spatial_count(semgate_handle_t h, const double* positions, int64_t n, double radius, int64_t min_index_gap,
                          void* workspace, int64_t* out_total, semgate_stream_t stream);
int semgate_spatial_fill(semgate_handle_t h, const double* positions, int64_t n, double radius, int64_t min_index_gap,
                         const void* workspace, int32_t* out_i, int32_t* out_j, double* out_dist, int64_t capacity,
                         semgate_stream_t stream);

/* ---- K5: CricaVPR cross-correlation re-rank ------------------------------------------
 * replaces compute_cross_correlation_score (place_recognition.py:669-710) over a batch of
 * (query, candidate) pairs and the score combination of rerank_candidates (:748).
 *   local_feats  device bf16 [n_feat, P, dl_pad]: patch features, rows L2-normalised
 *                (semgate_normalize_cast over the [n_feat*P, dl] matrix), dl_pad % 64 == 0
 *   query_idx / match_idx  int32 [M] keyframe numbers; a negative or out-of-range entry means
 *                "no cached local features" -> cross = NaN, combined = global (:749)
 *   out_cross[M] = sqrt(mean(row maxima) * mean(column maxima)) of q m^T
 *   out_combined[M] = 0.5 * global_sim + 0.5 * cross */
int semgate_rerank_scores(semgate_handle_t h, const void* local_feats, int64_t n_feat, int32_t P, int32_t dl_pad,
                          const int32_t* query_idx, const int32_t* match_idx, const float* global_sim, int64_t M,
                          float* out_cross, float* out_combined, semgate_stream_t stream);
/* per-query stable sort by combined score, descending, keep top_k (place_recognition.py:753-757).
 * cand_idx / combined: [Q, kc] padded lists (kc <= 64) with count[Q] live entries each;
 * outputs [Q, top_k] (-1 / -inf padded) and out_count[Q]. */
int semgate_rerank_select(semgate_handle_t h, const int32_t* cand_idx, const float* combined, const int32_t* count,
                          int64_t Q, int32_t kc, int32_t top_k, int32_t* out_idx, float* out_score, int32_t* out_count,
                          semgate_stream_t stream);

/* ---- find_loop_closures over a device-resident database: one call, caller-owned outputs ---------------
 * replaces SemanticPlaceRecognition.find_loop_closures (place_recognition.py:851-911) for a database that already
 * lives on the GPU as normalised bf16 rows (K1 output): K2 + K3 + K4 in one call, flat candidates (query
 * ascending, similarity descending) into the caller's arrays (capacity n*k each), their number in *out_total_dev
 * (device int64).  Nothing inside waits for the host.  use_graph != 0: the first call with a given argument set
 * runs eagerly, the second captures the same launch sequence into a CUDA graph, later calls replay it with one
 * cudaGraphLaunch (small sweeps are bound by launch overhead: BASELINE config 1 is a 60 us kernel).  The graph
 * is keyed on every argument, handle option and the params struct; up to 8 are cached per handle.  `stream` must not
 * be the legacy default stream for that (a capture cannot start there: the call then stays eager). */
size_t semgate_find_loop_closures_device_workspace_bytes(semgate_handle_t h, int64_t n, int32_t d_pad,
                                                         const semgate_topk_params* p);
int semgate_find_loop_closures_device(semgate_handle_t h, const void* x_bf16, int64_t n, int32_t d_pad, const double* ts,
                                      const int32_t* floor_labels, const semgate_topk_params* p, void* workspace,
                                      size_t workspace_bytes, int32_t* out_query_idx, int32_t* out_match_idx,
                                      float* out_similarity, uint8_t* out_is_valid, int64_t* out_total_dev,
                                      int32_t use_graph, semgate_stream_t stream);

/* ---- host-buffer entry points (the reference-facing calls) ----------------------
 * find_loop_closures over a whole database held in host memory
 * (SemanticPlaceRecognition.find_loop_closures, place_recognition.py:851-911):
 * H2D, K1, K2, K3, K4, D2H inside.  Output arrays need capacity >= min(n*k, capacity).
 * Returns SEMGATE_ENOMEM (and the required size in *out_total) if capacity is short. */
int semgate_find_loop_closures_host(semgate_handle_t h, const float* descriptors, int64_t n, int32_t d,
                                    const double* timestamps, const int32_t* floor_labels,
                                    const semgate_topk_params* p, int32_t* out_query_idx, int32_t* out_match_idx,
                                    float* out_similarity, uint8_t* out_is_valid, int64_t capacity, int64_t* out_total);

/* The same call for a host database of fp16 / bf16 descriptors (dtype: SEMGATE_DTYPE_*; _F32 is the call
 * above): half the bytes cross PCIe, which is what bounds the fp32 form.  Candidates
 * are bit-identical to the fp32 call on the widened descriptors. */
int semgate_find_loop_closures_host_dtype(semgate_handle_t h, const void* descriptors, int32_t dtype, int64_t n, int32_t d,
                                          const double* timestamps, const int32_t* floor_labels,
                                          const semgate_topk_params* p, int32_t* out_query_idx, int32_t* out_match_idx,
                                          float* out_similarity, uint8_t* out_is_valid, int64_t capacity,
                                          int64_t* out_total);

/* batched query() against a database in host memory (place_recognition.py:117-163):
 * padded outputs [nq,k] + count[nq]. */
int semgate_query_host(semgate_handle_t h, const float* queries, int64_t nq, const float* database, int64_t n, int32_t d,
                       const double* q_ts, const double* db_ts, const semgate_topk_params* p, float* out_scores,
                       int32_t* out_idx, int32_t* out_count);

/* gate_candidates over host arrays; out_counts: host uint64[3]. */
int semgate_gate_candidates_host(semgate_handle_t h, const int32_t* floor_labels, int64_t n_labels,
                                 const int32_t* query_idx, const int32_t* match_idx, int64_t M, int32_t max_floor_diff,
                                 uint8_t* out_is_valid, uint64_t* out_counts);

/* detect_loop_closure_candidates over host arrays.  *out_total always receives the number
 * of pairs; if it exceeds `capacity` (or the outputs are NULL) nothing is written and
 * SEMGATE_ENOMEM is returned, so a caller can size its buffers with a first call. */
int semgate_spatial_candidates_host(semgate_handle_t h, const double* positions, int64_t n, double radius,
                                    int64_t min_index_gap, int32_t* out_i, int32_t* out_j, double* out_dist,
                                    int64_t capacity, int64_t* out_total);

#ifdef __cplusplus
}
#endif
#endif /* SEMGATE_H_ */
