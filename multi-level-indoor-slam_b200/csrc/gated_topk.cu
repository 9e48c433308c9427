// Host side of K2: tile schedule, TMA tensor maps, launch.
#include "gated_topk.cuh"
#include "launch.h"

#include <cudaTypedefs.h>
#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace semgate {

// ------------------------------------------------------------------ schedule
Schedule make_schedule(int64_t Q, int64_t N, int d_pad, int cta_group, int sm_count) {
  Schedule sc{};
  const int units = std::max(1, sm_count / cta_group);
  const int bm_unit = BM * cta_group;
  sc.mblocks = static_cast<int>((Q + bm_unit - 1) / bm_unit);
  sc.ntiles = static_cast<int>((N + BN - 1) / BN);
  if (sc.mblocks <= 0 || sc.ntiles <= 0) { sc.mblocks = std::max(sc.mblocks, 0); return sc; }
  // Query blocks resident per super-row are re-read from L2 once per database
  // tile; keep them comfortably inside the 126 MB L2 next to the database tiles.
  const int64_t a_bytes = static_cast<int64_t>(bm_unit) * d_pad * 2;
  int64_t cap_mb = 48;
  if (const char* e = getenv("SEMGATE_RM_CAP_MB")) { const long v = atol(e); if (v > 0) cap_mb = v; }
  const int rm_cap = static_cast<int>(std::max<int64_t>(1, (cap_mb << 20) / std::max<int64_t>(a_bytes, 1)));
  const int rm_hi = std::min({sc.mblocks, units, rm_cap});
  int64_t best_cost = -1;
  int best_rm = 1;
  for (int rm = 1; rm <= rm_hi; ++rm) {
    const int s_main = std::min(sc.ntiles, units / rm);
    const int n_full = sc.mblocks / rm;
    const int r_last = sc.mblocks % rm;
    int64_t cost = static_cast<int64_t>(n_full) * ((sc.ntiles + s_main - 1) / s_main);
    if (r_last > 0) {
      const int s_last = std::min(sc.ntiles, units / r_last);
      cost += (sc.ntiles + s_last - 1) / s_last;
    }
    if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best_rm = rm; }   // ties -> larger rm (fewer DRAM passes)
  }
  sc.rm = best_rm;
  sc.s_main = std::min(sc.ntiles, units / sc.rm);
  sc.n_full = sc.mblocks / sc.rm;
  sc.r_last = sc.mblocks % sc.rm;
  sc.s_last = sc.r_last > 0 ? std::min(sc.ntiles, units / sc.r_last) : 0;
  sc.s_max = std::max(sc.n_full > 0 ? sc.s_main : 0, sc.s_last);
  // Column panels (off unless SEMGATE_PANEL_MB is set): panel width = a multiple of the split count
  // (no ceiling loss) whose tiles fit the given number of megabytes.
  sc.n_panels = 1;
  if (const char* e = getenv("SEMGATE_PANEL_MB")) {
    const long panel_mb = atol(e);
    if (panel_mb > 0) {
      const int64_t tile_bytes = static_cast<int64_t>(BN) * d_pad * 2;
      const int smax = std::max(sc.s_max, 1);
      int w_cap = static_cast<int>(std::max<int64_t>(1, (static_cast<int64_t>(panel_mb) << 20) / tile_bytes));
      w_cap = std::max(smax, (w_cap / smax) * smax);
      if (sc.ntiles > 4 * w_cap) {
        sc.n_panels = (sc.ntiles + w_cap - 1) / w_cap;
        sc.n_panels = std::max(1, std::min(sc.n_panels, sc.ntiles / smax));   // every panel keeps >= s_max tiles
      }
    }
  }
  return sc;
}

size_t topk_partial_bytes(const Schedule& sc, int cta_group, int k) {
  return static_cast<size_t>(sc.mblocks) * BM * cta_group * std::max(sc.s_max, 1) * k * sizeof(uint64_t);
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

static size_t smem_bytes_for(int cg, int stages, int kstride) {
  const size_t stage = A_STAGE_BYTES + static_cast<size_t>(BN / cg) * BK * 2;
  return 1024 /*realign slack*/ + stages * stage + static_cast<size_t>(BM) * kstride * 8 + 256 /*barriers + tmem slot*/;
}

int launch_gated_topk(const TopkLaunch& a, const Schedule& sc, uint64_t* partial, cudaStream_t st, int* launches) {
  if (a.Q <= 0 || a.N <= 0) return 0;
  const int cg = a.cta_group;
  CUtensorMap tq, tdb;
  int rc = make_tmap(&tq, a.q_bf16, a.Q, a.d_pad, BM);
  if (rc) return rc;
  rc = make_tmap(&tdb, a.db_bf16, a.N, a.d_pad, BN / cg);
  if (rc) return rc;

  TopkParams p{};
  p.Q = static_cast<int>(a.Q);
  p.N = static_cast<int>(a.N);
  p.kblocks = a.d_pad / BK;
  p.k = a.k;
  p.kstride = a.k | 1;
  p.threshold = a.threshold;
  p.use_time = (a.q_ts != nullptr && a.db_ts != nullptr) ? 1 : 0;
  p.gap = a.gap;
  p.max_floor_diff = a.max_floor_diff;
  p.gate_mode = a.gate_mode;
  p.db_index_offset = a.db_index_offset;
  p.q_ts = a.q_ts; p.db_ts = a.db_ts; p.q_floor = a.q_floor; p.db_floor = a.db_floor;
  p.partial = partial;
  p.sc = sc;

  // deepest ring that fits the 227 KB per-CTA limit
  const size_t limit = 232448;
  int stages = kMaxStages;
  while (stages > 2 && smem_bytes_for(cg, stages, p.kstride) > limit) --stages;
  stages = std::min(stages, std::max(2, p.kblocks));
  p.stages = stages;
  const size_t smem = smem_bytes_for(cg, stages, p.kstride);

  const int units = std::max(1, a.sm_count / cg);
  // only units that receive work in some super-row need to exist
  int used = 0;
  if (sc.n_full > 0) used = std::max(used, sc.rm * sc.s_main);
  if (sc.r_last > 0) used = std::max(used, sc.r_last * sc.s_last);
  used = std::min(std::max(used, 1), units);

  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(used * cg));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cudaError_t e;
  if (cg == 2) {
    e = cudaFuncSetAttribute(gated_topk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, gated_topk_kernel<2>, tq, tdb, p);
  } else {
    e = cudaFuncSetAttribute(gated_topk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    cfg.attrs = nullptr; cfg.numAttrs = 0;
    e = cudaLaunchKernelEx(&cfg, gated_topk_kernel<1>, tq, tdb, p);
  }
  if (e != cudaSuccess) return static_cast<int>(e);
  if (launches) ++*launches;
  return 0;
}

}  // namespace semgate
