// K5 — CricaVPR cross-correlation re-rank score, batched over candidate pairs, sm_100a.
//
// Replaces compute_cross_correlation_score (place_recognition.py:669-710), called once per
// (query, candidate) pair by rerank_candidates (:712-757):
//   q, m  : L2-normalised patch features [P, D]          (:695-699, done once per keyframe by K1)
//   corr  = q m^T                                       (:702)
//   score = sqrt( mean_rows(max_cols corr) * mean_cols(max_rows corr) )   (:706-710)
//   combined = 0.5 * global + 0.5 * score               (:748)
// The P x P correlation matrix lives only in TMEM: 128 x <=256 fp32 tiles from tcgen05.mma
// (bf16 operands staged by 3-D TMA straight out of the per-keyframe feature store), row maxima
// kept per thread, column maxima by warp REDUX + shared-memory atomicMax on order-preserving
// integer images of the floats.  One CTA per pair at a time (persistent, pairs round-robin).
#include "common.cuh"
#include "launch.h"
#include "ptx.cuh"

#include <cudaTypedefs.h>
#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace semgate {

namespace {

constexpr int RBM = 128, RBN = 256, RBK = 64, RUK = 16;
constexpr int kRThreads = 192;
constexpr int kRMaxStages = 6;
constexpr uint32_t RA_BYTES = RBM * RBK * 2;      // 16 KiB
// The n-tile width `bn` (<= RBN) is chosen per launch so that the P columns split evenly (P = 529 -> three
// tiles of 192 instead of 256 + 256 + 17): fewer zero-filled rows are staged and multiplied.

struct RerankParams {
  int P;              // patches per keyframe
  int kblocks;        // padded feature length / 64
  int stages;
  int n_feat;         // keyframes in the feature store
  int colcap;         // P rounded up to 32 (shared-memory column-max slots)
  int bn;             // n-tile width (multiple of 32, <= RBN)
  uint32_t b_bytes;   // bn * RBK * 2: one stage of the B operand
  uint32_t stage_bytes;
  long long M;        // pairs
  const int32_t* q_idx;
  const int32_t* m_idx;
  const float* global_sim;
  float* out_cross;
  float* out_combined;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// the box lands at the same smem offset in every CTA of `mask`; each destination's barrier (same offset) gets the bytes
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}

// Column maxima of a 32 x 32 block held one row per lane: lane l returns max over the warp's rows of column l.
// Five exchange steps (xor 16 ... 1): at each step a lane keeps the half of its columns whose index bit equals its own
// lane bit, sends the other half to its partner and folds what it receives -- 31 SHFL + 31 FMNMX + 62 SEL, all
// independent within a step.  (One REDUX per column, the round-1 form, measured ~57 clocks per column: the
// reductions go through the uniform datapath one at a time and the epilogue, not operand delivery, bounded K5.)
__device__ __forceinline__ float colmax32(const float (&v)[32], int lane) {
  float a[16], b[8], c[4], d[2];
  bool hi = (lane & 16) != 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float send = hi ? v[j] : v[j + 16], keep = hi ? v[j + 16] : v[j];
    a[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 16));
  }
  hi = (lane & 8) != 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float send = hi ? a[j] : a[j + 8], keep = hi ? a[j + 8] : a[j];
    b[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8));
  }
  hi = (lane & 4) != 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = hi ? b[j] : b[j + 4], keep = hi ? b[j + 4] : b[j];
    c[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 4));
  }
  hi = (lane & 2) != 0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = hi ? c[j] : c[j + 2], keep = hi ? c[j + 2] : c[j];
    d[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 2));
  }
  hi = (lane & 1) != 0;
  const float send = hi ? d[0] : d[1], keep = hi ? d[1] : d[0];
  return fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void epi_bar2() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ bool pair_ok(const RerankParams& p, long long pr, int& q, int& m) {
  q = m = -1;
  if (pr >= p.M) return false;
  q = p.q_idx[pr];
  m = p.m_idx[pr];
  return q >= 0 && m >= 0 && q < p.n_feat && m < p.n_feat;
}

// C = 1: one CTA scores one pair at a time.  C = 2: a cluster of two CTAs scores two consecutive
// pairs in lock-step; when both have the same query keyframe (the normal case: a query's candidates
// are consecutive) each CTA fetches half of every query tile and TMA-multicasts it to both, which
// takes the most re-staged operand off the L2 -> SM path once per pair instead of twice.
// Operand roles: A (128-row tiles, tensor map `tmap_a`) = the CANDIDATE's patches, B (256-row tiles,
// `tmap_b`, box of 256 / C rows) = the QUERY's patches; the score is symmetric in the two.
template <int C>
__global__ void __launch_bounds__(kRThreads, 1)
rerank_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const RerankParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.stages;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + static_cast<size_t>(stages) * RA_BYTES;
  uint32_t* colmax = reinterpret_cast<uint32_t*>(smem + static_cast<size_t>(stages) * p.stage_bytes);
  float* scratch = reinterpret_cast<float*>(colmax + p.colcap);            // 8 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + 8);
  const uint32_t bar_full = ptx::smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * kRMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kRMaxStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kRMaxStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (C > 1) ? ptx::cluster_ctarank() : 0;
  const long long cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;
  constexpr uint16_t kAll = static_cast<uint16_t>((1u << C) - 1u);
  const int BN = p.bn;
  const uint32_t B_LOAD_ROWS = static_cast<uint32_t>(BN) / C;
  const uint32_t B_LOAD_BYTES = B_LOAD_ROWS * RBK * 2;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    for (int s = 0; s < stages; ++s) { ptx::mbar_init(bar_full + 8 * s, 1); ptx::mbar_init(bar_empty + 8 * s, C); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(bar_tfull + 8 * a, 1); ptx::mbar_init(bar_tempty + 8 * a, 4); }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  for (int c = threadIdx.x; c < p.colcap; c += kRThreads) colmax[c] = 0u;
  if (warp == 1) ptx::tmem_alloc<1>(ptx::smem_u32(tmem_slot), 512);
  ptx::tc_fence_before();
  if constexpr (C > 1) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int MT = (p.P + RBM - 1) / RBM;
  const int NT = (p.P + BN - 1) / BN;
  uint32_t stage = 0, phase = 0, it = 0;

  // One group = C consecutive pairs, one per CTA of the cluster.  Every role of both CTAs walks the
  // same (group, m-tile, n-tile, k-block) sequence; a CTA whose pair is missing keeps the stage ring
  // turning (it still arrives on the barriers) so that its peer can run.
  struct Group { int q, m; bool mine, any, shared; };
  auto load_group = [&](long long g) {
    Group G;
    G.mine = pair_ok(p, g * C + rank, G.q, G.m);
    G.any = G.mine; G.shared = false;
    if constexpr (C > 1) {
      int q2, m2;
      const bool peer = pair_ok(p, g * C + (rank ^ 1u), q2, m2);
      G.any = G.mine || peer;
      G.shared = G.mine && peer && q2 == G.q;
    }
    return G;
  };

  if (warp == 0) {
    // ----------------------------------------------------------- TMA producer
    for (long long g = cluster_id; g * C < p.M; g += n_clusters) {
      const Group G = load_group(g);
      if (!G.any) continue;
      for (int mt = 0; mt < MT; ++mt)
        for (int nt = 0; nt < NT; ++nt)
          for (int kb = 0; kb < p.kblocks; ++kb) {
            ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t fb = bar_full + 8 * stage;
            if (ptx::elect_one()) {
              if (G.mine) {
                const uint32_t dst_b = ptx::smem_u32(smem_b) + stage * p.b_bytes;
                ptx::mbar_arrive_expect_tx(fb, p.stage_bytes);
                tma_load_3d(ptx::smem_u32(smem_a) + stage * RA_BYTES, &tmap_a, fb, kb * RBK, mt * RBM, G.m);
                if constexpr (C == 1) {
                  tma_load_3d(dst_b, &tmap_b, fb, kb * RBK, nt * BN, G.q);
                } else if (G.shared) {
                  tma_load_3d_mc(dst_b + rank * B_LOAD_BYTES, &tmap_b, fb, kb * RBK, nt * BN + static_cast<int>(rank * B_LOAD_ROWS),
                                 G.q, kAll);
                } else {
#pragma unroll
                  for (int hsel = 0; hsel < C; ++hsel)
                    tma_load_3d(dst_b + hsel * B_LOAD_BYTES, &tmap_b, fb, kb * RBK, nt * BN + hsel * static_cast<int>(B_LOAD_ROWS), G.q);
                }
              } else {
                ptx::mbar_arrive(fb);                           // nothing to load: hand the stage straight on
              }
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(stages)) { stage = 0; phase ^= 1; }
          }
    }
  } else if (warp == 1) {
    // ----------------------------------------------------------- MMA issuer
    const uint64_t adesc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_a));
    const uint64_t bdesc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_b));
    for (long long g = cluster_id; g * C < p.M; g += n_clusters) {
      const Group G = load_group(g);
      if (!G.any) continue;
      for (int mt = 0; mt < MT; ++mt)
        for (int nt = 0; nt < NT; ++nt) {
          const int ncols = min(BN, p.P - nt * BN);
          const uint32_t nw = static_cast<uint32_t>((ncols + 15) & ~15);         // MMA N: multiple of 16
          const uint32_t idesc = ptx::make_idesc_bf16_f32(RBM, nw);
          const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
          if (G.mine) {
            ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
            ptx::tc_fence_after();
          }
          const uint32_t d_tmem = tmem_base + acc * RBN;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            ptx::mbar_wait(bar_full + 8 * stage, phase);
            ptx::tc_fence_after();
            const uint64_t adesc = adesc0 + static_cast<uint64_t>((stage * RA_BYTES) >> 4);
            const uint64_t bdesc = bdesc0 + static_cast<uint64_t>((stage * p.b_bytes) >> 4);
            if (ptx::elect_one()) {
              if (G.mine) {
#pragma unroll
                for (int kk = 0; kk < RBK / RUK; ++kk)
                  ptx::umma_bf16<1>(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb | kk) != 0);
              }
              // the stage is free when every CTA that reads what lands in it has consumed it
              if constexpr (C == 1) ptx::umma_commit_cg1(bar_empty + 8 * stage);
              else ptx::umma_commit_cg1_mc(bar_empty + 8 * stage, kAll);
              if (G.mine && kb == p.kblocks - 1) ptx::umma_commit_cg1(bar_tfull + 8 * acc);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(stages)) { stage = 0; phase ^= 1; }
          }
          if (G.mine) ++it;
        }
    }
  } else {
    // ----------------------------------------------------------- epilogue: row / column maxima
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int et = (warp - 2) * 32 + lane;                       // 0..127 among the epilogue threads
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float neg_inf = __int_as_float(0xff800000);
    for (long long g = cluster_id; g * C < p.M; g += n_clusters) {
      const Group G = load_group(g);
      const long long pr = g * C + rank;
      if (!G.mine) {
        if (et == 0 && pr < p.M) {
          p.out_cross[pr] = __int_as_float(0x7fc00000);          // no cached features: global score only (:749)
          p.out_combined[pr] = p.global_sim[pr];
        }
        continue;
      }
      float rsum = 0.f;
      for (int mt = 0; mt < MT; ++mt) {
        const bool row_valid = mt * RBM + row_in_tile < p.P;
        float rmax = neg_inf;
        for (int nt = 0; nt < NT; ++nt, ++it) {
          const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
          ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
          ptx::tc_fence_after();
          const uint32_t t_acc = t_lane + acc * RBN;
          const int ncols = min(BN, p.P - nt * BN);
          for (int c = 0; c * 32 < ncols; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(t_acc + c * 32, v);
            ptx::tmem_wait_ld();
            const int nv = min(32, ncols - c * 32);
            float x[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float xi = __uint_as_float(v[i]);
              if (i < nv) rmax = fmaxf(rmax, xi);                // warp-uniform
              x[i] = row_valid ? xi : neg_inf;
            }
            const float cm = colmax32(x, lane);
            if (lane < nv) atomicMax(&colmax[nt * BN + c * 32 + lane], score_to_ordered(cm));
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * acc);
        }
        if (row_valid) rsum += rmax;
      }
      // ---- finish the pair: means of the maxima
      epi_bar();                                                  // every warp's atomics have landed
      float csum = 0.f;
      for (int c = et; c < p.P; c += 128) {
        csum += ordered_to_score(colmax[c]);
        colmax[c] = 0u;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
        csum += __shfl_xor_sync(0xffffffffu, csum, o);
      }
      if (lane == 0) { scratch[warp - 2] = rsum; scratch[4 + warp - 2] = csum; }
      epi_bar();
      if (et == 0) {
        const float rs = scratch[0] + scratch[1] + scratch[2] + scratch[3];
        const float cs = scratch[4] + scratch[5] + scratch[6] + scratch[7];
        const float inv = 1.0f / static_cast<float>(p.P);
        const float cross = sqrtf((rs * inv) * (cs * inv));
        p.out_cross[pr] = cross;
        p.out_combined[pr] = 0.5f * p.global_sim[pr] + 0.5f * cross;
      }
      epi_bar();                                                  // scratch / colmax free for the next pair
    }
  }

  ptx::tc_fence_before();
  if constexpr (C > 1) ptx::cluster_sync(); else __syncthreads();   // a peer may still multicast into / arrive on this CTA
  ptx::tc_fence_after();
  if (warp == 1) ptx::tmem_dealloc<1>(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// Pair form (round 2): one CTA PAIR per (query, candidate) pair, `tcgen05.mma.cta_group::2` tiles of 256 query patches
// (M, 128 per CTA) x `bn` candidate patches (N, half staged by each CTA).  Measured history at 529 x 768, 100k pairs:
//   round 1, single CTAs, 128 x 192 tiles, one REDUX per column          73.1 ms   588 TFLOP/s useful
//   ... column maxima by a shuffle butterfly (colmax32)                  54.0 ms   796   (the epilogue had bounded it)
//   pair tiles 256 x 192, left-over patches as strip MMAs                42.7 ms  1007   (L2 -> SM delivery: 11.7 TB/s,
//                                                                                 tensor pipe 78 %, profiles/r02_k5_pair.md)
//   candidate tile kept in a shared-memory ring across the m-tiles       see DESIGN.md
//
// Left-over patches.  Patch counts are not multiples of 256 (DINOv2 at 322 x 322: P = 529 = 2 * 256 + 17).  Padding the
// M side to three pair tiles would spend a third of the MMAs on zero rows, so when at most 32 query patches are left
// over they become a STRIP: in the pass of the first m-tile the staged candidate rows (the B operand of the main MMA,
// K-major and 128-byte swizzled like any A operand) are multiplied once more, as the M operand of a 256 x 32 MMA
// against the left-over query patches.  Those small accumulators (one per n-tile) sit behind the two main ones in TMEM
// and are read with the tile they belong to.  P = 529: 512 x 544 + 544 x 32 multiplied for 529 x 529 wanted (94 %), where
// the single-CTA form multiplies 640 x 576 (76 %).
//
// RING.  With k-block stages that hold both operands, every (m-tile, n-tile) stages 128 + bn/2 rows per CTA and k-block.
// The ring form loops n-tiles outside, m-tiles inside and keeps the CTA's half of the candidate tile -- all its
// k-blocks -- in shared memory (one slot per k-block, 144 KB at 768-d x 96 rows): a slot is filled once, read by every
// m-tile's pass (and by the strip MMA), and released by the last one, so the next n-tile's (or next pair's) k-block is
// fetched a whole pass ahead, which also hides the DRAM latency of the candidate's features (the query's are L2-hot: its
// 25 candidates are consecutive pairs).  Only the query k-blocks stream through a short stage ring.  Rows staged per CTA
// and pair at P = 529: 6 x 128 + 3 x 96 = 1056 against 6 x (128 + 96) = 1344.
//
// Maxima.  Row maxima (a query patch's best candidate patch) live in a per-thread shared-memory slot per m-tile; column
// maxima (a candidate patch's best query patch) and the strip's go through shared-memory atomicMax on order-preserving
// integer images.  The odd CTA pushes its arrays into the even CTA's over distributed shared memory at the end of the
// pair (red.max + a release-arrive on a barrier of the even CTA), which finishes the score.  Arrays, partial sums and the
// merge barrier are double-buffered by pair parity; the MMA pipeline keeps the CTAs of a pair within two tiles of each
// other (and a pair has at least two tiles), so no other hand-shake is needed.
struct PairParams {
  int P, kblocks, n_feat;
  int stages;          // stage ring: query k-blocks (RING) or query + candidate k-blocks
  int mt2;             // 256-row pair tiles over the query's patches
  int nt;              // n-tiles over the candidate's patches
  int bn;              // n-tile width (multiple of 32)
  int strip_rows;      // query patches [mt2 * 256, P) handled by the strip MMAs (0: none)
  int strip_resident;  // the strip operand stays in shared memory for the whole pair (1) or rides in the first m-tile's stages (0)
  int pp;              // P rounded up to 32
  uint32_t b_bytes;       // (bn / 2) rows of one k-block: this CTA's half of the candidate operand
  uint32_t stage_bytes;   // one stage: [candidate half (not RING)] [query rows, 16 KiB] [strip rows, 2 KiB (streamed strip)]
  long long M;
  const int32_t* q_idx;
  const int32_t* m_idx;
  const float* global_sim;
  float* out_cross;
  float* out_combined;
};
constexpr int kPMaxStages = 8;
constexpr int kPThreads = 384;                       // warps: 0 stage producer, 1 MMA issuer, 2..5 and 8..11 epilogue, 6 ring producer, 7 idle
constexpr int kPMaxKb = 16;                        // ring slots (k-blocks of the candidate tile): features up to 1024-d
constexpr uint32_t kStripKbBytes = 16 * RBK * 2;   // one k-block of this CTA's half of the strip operand: 16 rows

__device__ __forceinline__ bool pair_ok2(const PairParams& p, long long pr, int& q, int& m) {
  q = p.q_idx[pr];
  m = p.m_idx[pr];
  return q >= 0 && m >= 0 && q < p.n_feat && m < p.n_feat;
}

template <bool RING, int ST>
__global__ void __launch_bounds__(kPThreads, 1)
rerank_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ CUtensorMap tmap_s, const PairParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.stages, KB = p.kblocks;
  const bool strip = p.strip_rows > 0;
  const bool s_res = strip && p.strip_resident != 0, s_str = strip && p.strip_resident == 0;
  // [candidate ring (RING)] [stages] [resident strip operand] [maxima] [row maxima] [partial sums] [barriers]
  uint8_t* smem_ring = smem;
  uint8_t* smem_st = smem + (RING ? static_cast<size_t>(KB) * p.b_bytes : 0);
  uint8_t* smem_s = smem_st + static_cast<size_t>(stages) * p.stage_bytes;
  const int arr_len = p.pp + 32;                                                        // column maxima, then the strip's 32
  uint32_t* arrs = reinterpret_cast<uint32_t*>(smem_s + (s_res ? static_cast<size_t>(KB) * kStripKbBytes : 0));
  float* rm_s = reinterpret_cast<float*>(arrs + 2 * arr_len);                           // [2 parities][2 sets][mt2][128]
  float* scratch = rm_s + 4 * p.mt2 * RBM;                                              // [2][24]
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + 48);
  const uint32_t bar_full = ptx::smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * kPMaxStages;
  const uint32_t bar_bfull = bar_empty + 8 * kPMaxStages;
  const uint32_t bar_bempty = bar_bfull + 8 * kPMaxKb;
  const uint32_t bar_tfull = bar_bempty + 8 * kPMaxKb;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t bar_sfull = bar_tempty + 16;
  const uint32_t bar_sfree = bar_sfull + 8;
  const uint32_t bar_merge = bar_sfree + 8;                                             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPMaxStages + 2 * kPMaxKb + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const long long cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int BN = p.bn, NT = p.nt, MT2 = p.mt2;
  // offsets inside a stage
  const uint32_t st_a = RING ? 0u : p.b_bytes;       // query rows (the candidate half comes first: the strip MMA reads 128 rows
  const uint32_t st_s = st_a + RA_BYTES;             //   from its base, which must stay inside the stage); strip rows last

  ptx::cluster_sync();                               // the peer must be resident before a pair-wide TMEM allocation
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    ptx::prefetch_tensormap(&tmap_s);
    for (int s = 0; s < stages; ++s) { ptx::mbar_init(bar_full + 8 * s, 2); ptx::mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < kPMaxKb; ++s) { ptx::mbar_init(bar_bfull + 8 * s, 2); ptx::mbar_init(bar_bempty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(bar_tfull + 8 * a, 1);
      ptx::mbar_init(bar_tempty + 8 * a, 16);        // eight epilogue warps of each CTA
      ptx::mbar_init(bar_merge + 8 * a, 256);        // every epilogue thread of the odd CTA
    }
    ptx::mbar_init(bar_sfull, 2);
    ptx::mbar_init(bar_sfree, 1);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  for (int c = threadIdx.x; c < 2 * arr_len; c += kPThreads) arrs[c] = 0u;
  if (warp == 1) ptx::tmem_alloc<2>(ptx::smem_u32(tmem_slot), 512);
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  // columns of n-tile `t` that are multiplied (multiple of 16) -- each CTA stages half of them
  auto tile_nw = [&](int t) { return (min(BN, p.P - t * BN) + 15) & ~15; };

  uint32_t stage = 0, phase = 0, it = 0, pc = 0;     // pc: valid pairs so far (every role counts the same)

  if (warp == 0) {
    // ----------------------------------------------------------- TMA producer (both CTAs, each its halves)
    for (long long g = cluster_id; g < p.M; g += n_clusters) {
      int q, m;
      if (!pair_ok2(p, g, q, m)) continue;
      if (s_res) {
        ptx::mbar_wait(bar_sfree, (pc & 1u) ^ 1u);   // the previous pair's strip MMAs have read it
        if (ptx::elect_one()) {
          const uint32_t fb = ptx::mapa(bar_sfull, 0);
          if (leader) ptx::mbar_arrive_expect_tx(bar_sfull, 2u * static_cast<uint32_t>(KB) * kStripKbBytes);
          for (int kb = 0; kb < KB; ++kb)
            ptx::tma_load_3d_cg2(ptx::smem_u32(smem_s) + kb * kStripKbBytes, &tmap_s, fb, kb * RBK, MT2 * 256 + static_cast<int>(rank) * 16, q);
          if (!leader) ptx::mbar_arrive_cluster(bar_sfull, 0);
        }
        __syncwarp();
      }
      for (int nt = 0; nt < NT; ++nt) {
        const int b_row = nt * BN + static_cast<int>(rank) * (tile_nw(nt) >> 1);
        for (int mt = 0; mt < MT2; ++mt) {
          const bool with_s = s_str && mt == 0;
          const int a_row = mt * 256 + static_cast<int>(rank) * RBM;
          const uint32_t tx = 2u * ((RING ? 0u : p.b_bytes) + RA_BYTES + (with_s ? kStripKbBytes : 0u));
          for (int kb = 0; kb < KB; ++kb) {
            ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            if (ptx::elect_one()) {
              const uint32_t fb_local = bar_full + 8 * stage;
              const uint32_t fb = ptx::mapa(fb_local, 0);
              const uint32_t dst = ptx::smem_u32(smem_st) + stage * p.stage_bytes;
              if (leader) ptx::mbar_arrive_expect_tx(fb_local, tx);
              if constexpr (!RING) ptx::tma_load_3d_cg2(dst, &tmap_b, fb, kb * RBK, b_row, m);
              ptx::tma_load_3d_cg2(dst + st_a, &tmap_a, fb, kb * RBK, a_row, q);
              if (with_s) ptx::tma_load_3d_cg2(dst + st_s, &tmap_s, fb, kb * RBK, MT2 * 256 + static_cast<int>(rank) * 16, q);
              if (!leader) ptx::mbar_arrive_cluster(fb_local, 0);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(stages)) { stage = 0; phase ^= 1; }
          }
        }
      }
      ++pc;
    }
  } else if (warp == 6) {
    // ----------------------------------------------------------- RING: the candidate ring's own producer (both CTAs)
    // One slot per k-block; a slot is refilled as soon as the last m-tile's pass of the tile before has read it, so the
    // ring runs a whole pass ahead of the MMAs by itself -- no look-ahead bookkeeping, and the stage producer's loop
    // stays as short as the plain form's (one warp issuing both streams took ~900 clocks per k-block: the MMA issuer
    // spent half its time waiting for query stages that were not even requested yet).
    if constexpr (RING) {
      uint32_t bt = 0;                               // candidate tiles so far = use count of every slot
      for (long long g = cluster_id; g < p.M; g += n_clusters) {
        int q, m;
        if (!pair_ok2(p, g, q, m)) continue;
        for (int nt = 0; nt < NT; ++nt, ++bt) {
          const int b_row = nt * BN + static_cast<int>(rank) * (tile_nw(nt) >> 1);
          for (int kb = 0; kb < KB; ++kb) {
            ptx::mbar_wait(bar_bempty + 8 * kb, (bt & 1u) ^ 1u);
            if (ptx::elect_one()) {
              const uint32_t fb_local = bar_bfull + 8 * kb;
              if (leader) ptx::mbar_arrive_expect_tx(fb_local, 2u * p.b_bytes);
              ptx::tma_load_3d_cg2(ptx::smem_u32(smem_ring) + kb * p.b_bytes, &tmap_b, ptx::mapa(fb_local, 0), kb * RBK, b_row, m);
              if (!leader) ptx::mbar_arrive_cluster(fb_local, 0);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ----------------------------------------------------------- MMA issuer (leader CTA)
    if (leader) {
      const uint32_t lo_st0 = ptx::smem_desc_lo_sw128(ptx::smem_u32(smem_st));      // low descriptor words (ptx::umma4_bf16_cg2)
      const uint32_t lo_ring0 = ptx::smem_desc_lo_sw128(ptx::smem_u32(smem_ring));
      const uint32_t lo_s0 = ptx::smem_desc_lo_sw128(ptx::smem_u32(smem_s));
      constexpr uint32_t idesc_strip = ptx::make_idesc_bf16_f32(256, 32);
      uint32_t bt = 0;                               // candidate tiles so far: use count of the ring slots
      for (long long g = cluster_id; g < p.M; g += n_clusters) {
        int q, m;
        if (!pair_ok2(p, g, q, m)) continue;
        for (int nt = 0; nt < NT; ++nt, ++bt) {
          const uint32_t idesc = ptx::make_idesc_bf16_f32(256, static_cast<uint32_t>(tile_nw(nt)));
          for (int mt = 0; mt < MT2; ++mt, ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const bool do_strip = strip && mt == 0;
            ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
            if (s_res && nt == 0 && mt == 0) ptx::mbar_wait(bar_sfull, pc & 1u);
            ptx::tc_fence_after();
            const uint32_t d_main = tmem_base + acc * BN;
            const uint32_t d_strip = tmem_base + 2 * BN + 32 * nt;
            // one k-block: wait for its stage (and, first pass of a RING tile, its ring slot), issue, release
            auto step = [&](int kb, uint32_t sg, uint32_t ph) {
              ptx::mbar_wait(bar_full + 8 * sg, ph);
              if (RING && mt == 0) ptx::mbar_wait(bar_bfull + 8 * kb, bt & 1u);
              ptx::tc_fence_after();
              const uint32_t st_lo = lo_st0 + ((sg * p.stage_bytes) >> 4);
              const uint32_t a_lo = st_lo + (st_a >> 4);
              const uint32_t b_lo = RING ? lo_ring0 + ((static_cast<uint32_t>(kb) * p.b_bytes) >> 4) : st_lo;
              if (ptx::elect_one()) {
                ptx::umma4_bf16_cg2(d_main, a_lo, b_lo, idesc, static_cast<uint32_t>(kb));
                if (do_strip) {
                  // the staged candidate rows once more, as the M operand against the left-over query patches
                  const uint32_t s_lo = s_res ? lo_s0 + ((static_cast<uint32_t>(kb) * kStripKbBytes) >> 4) : st_lo + (st_s >> 4);
                  ptx::umma4_bf16_cg2(d_strip, b_lo, s_lo, idesc_strip, static_cast<uint32_t>(kb));
                }
                ptx::umma_commit_cg2_mc(bar_empty + 8 * sg, 0b11);
                if (RING && mt == MT2 - 1) ptx::umma_commit_cg2_mc(bar_bempty + 8 * kb, 0b11);
                if (kb == KB - 1) {
                  if (s_res && mt == 0 && nt == NT - 1) ptx::umma_commit_cg2_mc(bar_sfree, 0b11);
                  ptx::umma_commit_cg2_mc(bar_tfull + 8 * acc, 0b11);
                }
              }
              __syncwarp();
            };
            if constexpr (ST > 0) {
              // the stage count divides the k-blocks of a tile: stage indices are compile-time inside the unrolled group, so
              // barrier addresses and descriptors are loop-invariant (the run-time form spent ~110 instructions per k-block
              // on them, 560 clocks on a scheduler it shares with an epilogue warp, against 384 clocks of MMAs)
              for (int kb0 = 0; kb0 < KB; kb0 += ST) {
#pragma unroll
                for (int sg = 0; sg < ST; ++sg) step(kb0 + sg, static_cast<uint32_t>(sg), phase);
                phase ^= 1;
              }
            } else {
              for (int kb = 0; kb < KB; ++kb) {
                step(kb, stage, phase);
                if (++stage == static_cast<uint32_t>(stages)) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
        ++pc;
      }
      if (it > 0) {                                  // the peer's epilogue arrives remotely on our barriers: drain
        const uint32_t last = it - 1;
        ptx::mbar_wait(bar_tempty + 8 * (last & 1), (last >> 1) & 1);
      }
    }
    __syncwarp();
  } else if (warp != 7) {
    // ----------------------------------------------------------- epilogue: row / column maxima (both CTAs)
    // Two sets of four warps (2..5 and 8..11; a warp reads the TMEM lanes of quad warp % 4): set s takes the 32-column
    // chunks c = s, s + 2, ... of every tile.  One set was as slow as the MMAs (a chunk is ~250 instructions and a chain of
    // five dependent shuffle levels), so every form of this kernel -- whatever it staged -- ended at ~30 us per pair.
    const int set = warp >= 8 ? 1 : 0;
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int et = (set ? (warp - 8) : (warp - 2)) * 32 + lane + set * 128;     // 0..255 among the epilogue threads
    const int ew = et >> 5;                                                      // 0..7
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float neg_inf = __int_as_float(0xff800000);
    const int row_limit = strip ? MT2 * 256 : p.P;   // query patches covered by main tiles
    for (long long g = cluster_id; g < p.M; g += n_clusters) {
      int q, m;
      if (!pair_ok2(p, g, q, m)) {
        if (leader && et == 0) {
          p.out_cross[g] = __int_as_float(0x7fc00000);          // no cached features: global score only (:749)
          p.out_combined[g] = p.global_sim[g];
        }
        continue;
      }
      const uint32_t par = pc & 1u;
      uint32_t* arr = arrs + par * arr_len;
      float* scr = scratch + par * 24;
      float* rm_pair = rm_s + par * (2 * MT2 * RBM);               // partial row maxima, [set][mt][row in tile]; the other set reads
      float* rm_mine = rm_pair + set * (MT2 * RBM);                // them at the end of the pair, hence one copy per pair parity
      for (int mt = 0; mt < MT2; ++mt) rm_mine[mt * RBM + row_in_tile] = neg_inf;     // this thread's slots only
      for (int nt = 0; nt < NT; ++nt) {
        const int ncols = min(BN, p.P - nt * BN);
        const int strip_set = ((ncols + 31) >> 5) & 1;           // the set with fewer main chunks (or set 0) reads the strip
        for (int mt = 0; mt < MT2; ++mt, ++it) {
          const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
          const bool row_valid = mt * 256 + static_cast<int>(rank) * RBM + row_in_tile < row_limit;
          float rmax = rm_mine[mt * RBM + row_in_tile];
          ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
          ptx::tc_fence_after();
          const uint32_t t_acc = t_lane + acc * BN;
          for (int c = set; c * 32 < ncols; c += 2) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(t_acc + c * 32, v);
            ptx::tmem_wait_ld();
            const int nv = min(32, ncols - c * 32);
            float x[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float xi = __uint_as_float(v[i]);
              if (i < nv) rmax = fmaxf(rmax, xi);                // warp-uniform
              x[i] = row_valid ? xi : neg_inf;
            }
            const float cm = colmax32(x, lane);
            if (lane < nv) atomicMax(&arr[nt * BN + c * 32 + lane], score_to_ordered(cm));
          }
          rm_mine[mt * RBM + row_in_tile] = rmax;
          if (strip && mt == 0 && set == strip_set) {
            // rows: the candidate patches this CTA staged for the tile; columns: the left-over query patches
            const int half = tile_nw(nt) >> 1;
            const int crow = nt * BN + static_cast<int>(rank) * half + row_in_tile;
            const bool cvld = row_in_tile < half && crow < p.P;
            uint32_t v[32];
            ptx::tmem_ld_32x32(t_lane + 2 * BN + 32 * nt, v);
            ptx::tmem_wait_ld();
            float r2 = neg_inf;
            float x[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float xi = __uint_as_float(v[i]);
              if (i < p.strip_rows) r2 = fmaxf(r2, xi);          // warp-uniform
              x[i] = cvld ? xi : neg_inf;
            }
            const float cm = colmax32(x, lane);
            if (lane < p.strip_rows) atomicMax(&arr[p.pp + lane], score_to_ordered(cm));
            if (cvld) atomicMax(&arr[crow], score_to_ordered(r2));
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) ptx::mbar_arrive(bar_tempty + 8 * acc);
            else ptx::mbar_arrive_cluster(bar_tempty + 8 * acc, 0);
          }
        }
      }
      // ---- finish the pair
      epi_bar2();                                                 // every warp's atomics and partial row maxima have landed
      float rsum = 0.f;
      if (set == 0) {
        for (int mt = 0; mt < MT2; ++mt)
          if (mt * 256 + static_cast<int>(rank) * RBM + row_in_tile < row_limit)
            rsum += fmaxf(rm_pair[mt * RBM + row_in_tile], rm_pair[(MT2 + mt) * RBM + row_in_tile]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
      }
      if (!leader) {
        for (int c = et; c < arr_len; c += 256) {
          const uint32_t vv = arr[c];
          if (vv != 0u) {
            ptx::red_max_u32_cluster(ptx::mapa(ptx::smem_u32(arr + c), 0), vv);
            arr[c] = 0u;
          }
        }
        if (set == 0 && lane == 0) ptx::st_f32_cluster(ptx::mapa(ptx::smem_u32(scr + 4 + ew), 0), rsum);
        ptx::mbar_arrive_release_cluster(bar_merge + 8 * par, 0);
      } else {
        if (set == 0 && lane == 0) scr[ew] = rsum;
        ptx::mbar_wait_acquire_cluster(bar_merge + 8 * par, (pc >> 1) & 1u);
        float csum = 0.f, ssum = 0.f;
        for (int c = et; c < p.P; c += 256) {
          csum += ordered_to_score(arr[c]);
          arr[c] = 0u;
        }
        if (et < 32) {
          if (et < p.strip_rows) ssum = ordered_to_score(arr[p.pp + et]);
          arr[p.pp + et] = 0u;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          csum += __shfl_xor_sync(0xffffffffu, csum, o);
          ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        }
        if (lane == 0) { scr[8 + ew] = csum; if (ew == 0) scr[16] = ssum; }
        epi_bar2();
        if (et == 0) {
          const float rs = ((scr[0] + scr[1]) + (scr[2] + scr[3])) + ((scr[4] + scr[5]) + (scr[6] + scr[7])) + scr[16];
          const float cs = ((scr[8] + scr[9]) + (scr[10] + scr[11])) + ((scr[12] + scr[13]) + (scr[14] + scr[15]));
          const float inv = 1.0f / static_cast<float>(p.P);
          const float cross = sqrtf((rs * inv) * (cs * inv));
          p.out_cross[g] = cross;
          p.out_combined[g] = 0.5f * p.global_sim[g] + 0.5f * cross;
        }
      }
      ++pc;
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  if (warp == 1) ptx::tmem_dealloc<2>(tmem_base, 512);
}


PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

// [n_feat, P, dl_pad] bf16, box = [1, box_rows, 64], 128-byte swizzle; rows >= P read as zero
int make_tmap3(CUtensorMap* m, const void* base, int n_feat, int P, int dl_pad, uint32_t box_rows) {
  auto enc = encode_fn();
  if (!enc) return -1;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(dl_pad), static_cast<cuuint64_t>(P), static_cast<cuuint64_t>(n_feat)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(dl_pad) * 2, static_cast<cuuint64_t>(P) * dl_pad * 2};
  cuuint32_t box[3] = {RBK, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

}  // namespace

namespace {
constexpr int kPairFormNotApplicable = -77;

// the pair form's tiling for P patches x dl_pad features; false: use a round-1 form.  `ring`: in = allowed, out = chosen.
bool pair_tiling(int P, int dl_pad, PairParams& p, size_t& smem, bool& ring, int strip_mode) {
  if (P <= RBM) return false;                          // one 128-row tile: a CTA pair would multiply zeros
  const int kblocks = dl_pad / RBK;
  const int full = P / 256, rem = P - full * 256;
  const bool ring_allowed = ring && kblocks <= kPMaxKb;
  for (int want_strip = 1; want_strip >= 0; --want_strip) {
    const bool strip = want_strip && full >= 1 && rem > 0 && rem <= 32;
    if (want_strip && !strip) continue;
    const int bnmax = strip ? 192 : RBN;
    // at least two tiles per pair: the double-buffered merge arrays rely on the MMA pipeline keeping the CTAs of a pair
    // within two TILES of each other, which must be less than two PAIRS
    const int mt2 = strip ? full : (P + 255) / 256;
    int nt = (P + bnmax - 1) / bnmax;
    if (mt2 * nt < 2) nt = 2;
    const int bn = (((P + nt - 1) / nt) + 31) & ~31;
    if (bn > bnmax) continue;
    if (strip && 2 * bn + 32 * nt > 512) continue;     // TMEM: two main accumulators + one strip accumulator per n-tile
    p.P = P; p.kblocks = kblocks; p.mt2 = mt2; p.nt = nt; p.bn = bn;
    p.strip_rows = strip ? rem : 0;
    p.pp = (P + 31) & ~31;
    p.b_bytes = static_cast<uint32_t>(bn / 2) * RBK * 2;
    const size_t fixed = 1024 + static_cast<size_t>(2 * (p.pp + 32)) * 4 + static_cast<size_t>(4 * mt2) * RBM * 4 + 192 +
                         (2 * kPMaxStages + 2 * kPMaxKb + 8) * 8 + 16;
    // candidate ring first (fewest bytes staged), then plain stages; the strip operand resident or streamed, whichever
    // leaves more stages (streamed on a tie: it frees a slot's worth of shared memory for the ring)
    for (int r = ring_allowed ? 1 : 0; r >= 0; --r) {
      int best_stages = 0, best_res = 0;
      for (int res = 1; res >= 0; --res) {
        if (!strip && res == 0) continue;
        if (strip && strip_mode >= 0 && res != strip_mode) continue;
        const size_t stage_bytes = (r ? 0 : p.b_bytes) + RA_BYTES + ((strip && !res) ? kStripKbBytes : 0);
        const size_t other = fixed + (r ? static_cast<size_t>(kblocks) * p.b_bytes : 0) + ((strip && res) ? static_cast<size_t>(kblocks) * kStripKbBytes : 0);
        if (other + 3 * stage_bytes > 232448) continue;
        const int st = static_cast<int>(std::min<size_t>((232448 - other) / stage_bytes, kPMaxStages));
        if (st >= best_stages) { best_stages = st; best_res = res; }
      }
      if (best_stages < 3) continue;
      p.strip_resident = best_res;
      p.stage_bytes = static_cast<uint32_t>((r ? 0 : p.b_bytes) + RA_BYTES + ((strip && !best_res) ? kStripKbBytes : 0));
      p.stages = best_stages;
      smem = fixed + (r ? static_cast<size_t>(kblocks) * p.b_bytes : 0) + ((strip && best_res) ? static_cast<size_t>(kblocks) * kStripKbBytes : 0) +
             static_cast<size_t>(p.stages) * p.stage_bytes;
      ring = r != 0;
      return true;
    }
  }
  return false;
}

int launch_rerank_pair(const void* feats_bf16, int n_feat, int P, int dl_pad, const int32_t* q_idx, const int32_t* m_idx,
                       const float* global_sim, int64_t M, float* out_cross, float* out_combined, int sm_count, cudaStream_t st) {
  PairParams p{};
  size_t smem = 0;
  // A/B knobs (defaults are the measured best): SEMGATE_RERANK_RING=0 forbids the candidate ring,
  // SEMGATE_RERANK_STRIP=1 / 0 pins the strip operand resident / streamed
  bool ring = true;
  int strip_mode = -1;
  if (const char* e = getenv("SEMGATE_RERANK_RING")) ring = atoi(e) != 0;
  if (const char* e = getenv("SEMGATE_RERANK_STRIP")) strip_mode = atoi(e) != 0 ? 1 : 0;
  if (!pair_tiling(P, dl_pad, p, smem, ring, strip_mode)) return kPairFormNotApplicable;
  CUtensorMap ta, tb, ts;
  int rc = make_tmap3(&ta, feats_bf16, n_feat, P, dl_pad, RBM);
  if (rc) return rc;
  rc = make_tmap3(&tb, feats_bf16, n_feat, P, dl_pad, static_cast<uint32_t>(p.bn / 2));
  if (rc) return rc;
  rc = make_tmap3(&ts, feats_bf16, n_feat, P, dl_pad, 16);
  if (rc) return rc;
  p.n_feat = n_feat;
  p.M = M;
  p.q_idx = q_idx; p.m_idx = m_idx; p.global_sim = global_sim;
  p.out_cross = out_cross; p.out_combined = out_combined;
  const unsigned clusters = static_cast<unsigned>(std::min<int64_t>(M, std::max(1, sm_count / 2)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * 2);
  cfg.blockDim = dim3(kPThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  auto launch = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, kernel, ta, tb, ts, p);
  };
  // compile-time stage counts that divide the k-blocks of a tile (see the MMA issuer); otherwise the run-time ring
  int ust = 0;
  if (const char* e = getenv("SEMGATE_RERANK_UNROLL")) ust = atoi(e) != 0 ? 0 : -1;
  if (ust == 0) {
    for (int cand : {6, 4, 3}) if (p.stages >= cand && p.kblocks % cand == 0) { ust = cand; break; }
    if (ust > 0) {
      smem -= static_cast<size_t>(p.stages - ust) * p.stage_bytes;
      p.stages = ust;
    }
  }
  cfg.dynamicSmemBytes = smem;
  cudaError_t e;
  if (ring) e = ust == 6 ? launch(rerank_pair_kernel<true, 6>) : ust == 4 ? launch(rerank_pair_kernel<true, 4>) : ust == 3 ? launch(rerank_pair_kernel<true, 3>) : launch(rerank_pair_kernel<true, 0>);
  else e = ust == 6 ? launch(rerank_pair_kernel<false, 6>) : ust == 4 ? launch(rerank_pair_kernel<false, 4>) : ust == 3 ? launch(rerank_pair_kernel<false, 3>) : launch(rerank_pair_kernel<false, 0>);
  return static_cast<int>(e);
}
}  // namespace


int launch_rerank(const void* feats_bf16, int n_feat, int P, int dl_pad, const int32_t* q_idx, const int32_t* m_idx,
                  const float* global_sim, int64_t M, float* out_cross, float* out_combined, int sm_count,
                  cudaStream_t st) {
  if (M <= 0) return 0;
  // The pair form (one CTA pair per candidate pair) wherever its tiling applies; SEMGATE_RERANK_CLUSTER=1 / 2 asks for
  // the round-1 forms (single CTAs / two pairs per 2-CTA cluster in lock-step), which also serve P <= 128.
  const char* form_env = getenv("SEMGATE_RERANK_CLUSTER");
  if (form_env == nullptr || atoi(form_env) == 0) {
    const int rc = launch_rerank_pair(feats_bf16, n_feat, P, dl_pad, q_idx, m_idx, global_sim, M, out_cross, out_combined, sm_count, st);
    if (rc != kPairFormNotApplicable) return rc;
  }
  int csize = 2;
  if (form_env) csize = atoi(form_env) == 1 ? 1 : 2;
  if (M < 2) csize = 1;
  // n-tile width: split the P columns evenly over ceil(P / 256) tiles, in multiples of 32
  const int nt_count = (P + RBN - 1) / RBN;
  const int bn = std::min(RBN, (((P + nt_count - 1) / nt_count) + 31) & ~31);
  CUtensorMap ta, tb;
  int rc = make_tmap3(&ta, feats_bf16, n_feat, P, dl_pad, RBM);
  if (rc) return rc;
  rc = make_tmap3(&tb, feats_bf16, n_feat, P, dl_pad, static_cast<uint32_t>(bn / csize));
  if (rc) return rc;
  RerankParams p{};
  p.P = P;
  p.kblocks = dl_pad / RBK;
  p.n_feat = n_feat;
  p.colcap = (P + 31) & ~31;
  p.bn = bn;
  p.b_bytes = static_cast<uint32_t>(bn) * RBK * 2;
  p.stage_bytes = RA_BYTES + p.b_bytes;
  p.M = M;
  p.q_idx = q_idx; p.m_idx = m_idx; p.global_sim = global_sim;
  p.out_cross = out_cross; p.out_combined = out_combined;
  const size_t fixed = 1024 + static_cast<size_t>(p.colcap) * 4 + 32 + 256;
  int stages = kRMaxStages;
  while (stages > 2 && fixed + static_cast<size_t>(stages) * p.stage_bytes > 232448) --stages;
  p.stages = std::min(stages, std::max(2, p.kblocks));
  const size_t smem = fixed + static_cast<size_t>(p.stages) * p.stage_bytes;
  const int64_t groups = (M + csize - 1) / csize;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(groups, sm_count / csize) * csize);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kRThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(csize); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = csize > 1 ? 1 : 0;
  auto launch = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, kernel, ta, tb, p);
  };
  return static_cast<int>(csize == 2 ? launch(rerank_kernel<2>) : launch(rerank_kernel<1>));
}

// ------------------------------------------------------------------------------------------
// Per-query selection after re-ranking (place_recognition.py:753-757): stable sort of each
// query's candidates by combined score, descending; keep top_k.  One warp per query, <= 64
// candidates per query; rank by counting.
__global__ void __launch_bounds__(256)
rerank_select_kernel(const int32_t* __restrict__ cand_idx, const float* __restrict__ combined, const int32_t* __restrict__ count,
                     int64_t Q, int kc, int top_k, int32_t* __restrict__ out_idx, float* __restrict__ out_score,
                     int32_t* __restrict__ out_count) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= Q) return;
  const int n = min(count[row], kc);
  const float neg_inf = __int_as_float(0xff800000);
  float s[2];
  int id[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int e = h * 32 + lane;
    const bool live = e < n;
    const float x = live ? combined[row * kc + e] : neg_inf;
    s[h] = (x == x) ? x : neg_inf;                               // NaN sorts last
    id[h] = live ? cand_idx[row * kc + e] : -1;
  }
  int rank[2] = {0, 0};
  for (int h2 = 0; h2 < 2; ++h2)
    for (int l2 = 0; l2 < 32; ++l2) {
      const int e2 = h2 * 32 + l2;
      if (e2 >= n) break;                                        // warp-uniform
      const float o = __shfl_sync(0xffffffffu, s[h2], l2);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = h * 32 + lane;
        rank[h] += (o > s[h] || (o == s[h] && e2 < e)) ? 1 : 0;  // stable: earlier entry wins ties
      }
    }
  const int keep = min(n, top_k);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int e = h * 32 + lane;
    if (e < n && rank[h] < keep) {
      out_idx[row * top_k + rank[h]] = id[h];
      out_score[row * top_k + rank[h]] = combined[row * kc + e];
    }
  }
  for (int t = keep + lane; t < top_k; t += 32) {
    out_idx[row * top_k + t] = -1;
    out_score[row * top_k + t] = neg_inf;
  }
  if (lane == 0) out_count[row] = keep;
}

int launch_rerank_select(const int32_t* cand_idx, const float* combined, const int32_t* count, int64_t Q, int kc, int top_k,
                         int32_t* out_idx, float* out_score, int32_t* out_count, cudaStream_t st) {
  if (Q <= 0) return 0;
  const unsigned grid = static_cast<unsigned>((Q + 7) / 8);
  rerank_select_kernel<<<grid, 256, 0, st>>>(cand_idx, combined, count, Q, kc, top_k, out_idx, out_score, out_count);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace semgate
