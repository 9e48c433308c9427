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

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

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
            uint32_t mine = 0u;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float x = __uint_as_float(v[i]);
              const bool cv = i < nv;                            // warp-uniform
              if (cv) rmax = fmaxf(rmax, x);
              const uint32_t o = (row_valid && cv) ? score_to_ordered(x) : 0u;
              const uint32_t red = __reduce_max_sync(0xffffffffu, o);
              if (lane == i) mine = red;
            }
            if (lane < nv) atomicMax(&colmax[nt * BN + c * 32 + lane], mine);
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

int launch_rerank(const void* feats_bf16, int n_feat, int P, int dl_pad, const int32_t* q_idx, const int32_t* m_idx,
                  const float* global_sim, int64_t M, float* out_cross, float* out_combined, int sm_count,
                  cudaStream_t st) {
  if (M <= 0) return 0;
  // two-CTA clusters with a multicast query operand unless SEMGATE_RERANK_CLUSTER=1 asks for single CTAs
  int csize = 2;
  if (const char* e = getenv("SEMGATE_RERANK_CLUSTER")) csize = atoi(e) == 1 ? 1 : 2;
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
