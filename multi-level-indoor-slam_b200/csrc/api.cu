// C ABI of libsemgate (see include/semgate.h).
#include "../../include/semgate.h"
#include "launch.h"

#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

using namespace semgate;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(static_cast<int>(e), "%s: %s", what, cudaGetErrorString(e));
}

#define CUDA_TRY(expr)                                        \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return cuda_fail(_e, #expr);       \
  } while (0)

#define RC_TRY(expr, what)                                                                  \
  do {                                                                                      \
    int _rc = (expr);                                                                       \
    if (_rc > 0) return cuda_fail(static_cast<cudaError_t>(_rc), what);                     \
    if (_rc < 0) return fail(SEMGATE_EDRIVER, "%s failed (driver rc %d)", what, _rc);       \
  } while (0)

constexpr int kNumBufs = 17;
constexpr int kClkCtas = 1024;     // CTAs the clock probe has room for
enum BufId { B_X = 0, B_BF16, B_TS, B_FL, B_WS, B_SC, B_IX, B_VA, B_CT, B_OQ, B_OM, B_OS, B_OV, B_TOT, B_CWS, B_QBF16, B_KEYS };

inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

}  // namespace

// a run table (small symmetric sweeps) and its device copy, cached per shape
struct SymTableDev {
  int64_t nb; int d_pad, part_index, part_count;   // keyed by the tile count: ceil(N / 256)
  SymTable host;
  void* dev = nullptr;
  cudaStream_t upload_stream = nullptr;
  cudaEvent_t uploaded = nullptr;
  Schedule sc_dev;
};

// One captured find_loop_closures_device sequence; replayed while every argument is the same.
struct SweepGraphKey {
  const void* x; int64_t n; int32_t d_pad; const double* ts; const int32_t* floor; semgate_topk_params p;
  void* ws; size_t ws_bytes; int32_t* oq; int32_t* om; float* os; uint8_t* ov; int64_t* tot; int cta_group, symmetric;
};
struct SweepGraph {
  SweepGraphKey key;
  int seen = 0;                    // calls with this key so far (the first runs eagerly, the second is captured)
  cudaGraphExec_t exec = nullptr;
  int launches = 0;                // kernels + memsets of one replay
  int mode = 0; int64_t tiles = 0;
};

struct semgate_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  int cta_group = 0;               // 0 = auto (by problem size), 1, 2
  int symmetric = 0;               // aliased queries/database: 0 = symmetric sweep when it pays (size rule), 1 = whenever possible, -1 = never
  int last_mode = 0;               // last fused sweep: 0 full, 1 symmetric
  cudaStream_t last_stream = nullptr;
  int64_t last_tiles = 0;          // tiles the last fused sweep computed (its schedule's count)
  int64_t launches = 0;
  cudaStream_t stream = nullptr;   // used by the *_host entry points (compute)
  cudaStream_t copy_stream = nullptr;   // H2D of the next chunks, back to back, overlapped with the sweeps
  std::vector<cudaEvent_t> chunk_events;
  bool profile = false;            // record CUDA events around every K2 launch
  std::vector<struct SweepGraph*> graphs;   // semgate_find_loop_closures_device: captured launch sequences
  bool graphs_broken = false;      // a capture failed on this system: stay eager
  // one-call sweep in flight: K3 zeroes the compaction state, K3 / K4 launch as programmatic dependents (MergeLaunch)
  unsigned long long* fl_zero = nullptr; int fl_zero_words = 0; bool fl_pdl = false;
  unsigned long long* clk_dev = nullptr;   // option "clock_probe": per-CTA {globaltimer, clock64} pairs of the last K2 launch
  int clk_ctas = 0;
  std::vector<SymTableDev*> sym_tables, sym_retired;
  uint32_t* flag_dev = nullptr;    // overflow flag of the last symmetric sweep (copied by its merge kernel: the
                                   // caller's workspace may be reused or freed before anybody asks)
  std::vector<cudaEvent_t> prof_events;   // pairs (begin, end), on the launching stream
  size_t prof_used = 0;
  void* buf[kNumBufs] = {};
  size_t cap[kNumBufs] = {};

  int reserve(int id, size_t bytes, void** out) {
    if (bytes == 0) bytes = 256;
    if (cap[id] < bytes) {
      if (buf[id]) cudaFree(buf[id]);
      buf[id] = nullptr;
      cap[id] = 0;
      size_t want = align256(bytes + bytes / 8);
      cudaError_t e = cudaMalloc(&buf[id], want);
      if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc (scratch)");
      cap[id] = want;
    }
    *out = buf[id];
    return 0;
  }
};

namespace {

bool env_flag_default(const char* name, bool dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) != 0 : dflt;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int check_params(const semgate_topk_params* p) {
  if (!p) return fail(SEMGATE_EINVAL, "params is NULL");
  if (p->k < 1 || p->k > SEMGATE_MAX_K_TOTAL) return fail(SEMGATE_EINVAL, "k=%d outside 1..%d", p->k, SEMGATE_MAX_K_TOTAL);
  if (p->k > SEMGATE_MAX_K && (p->accumulate || p->part_count > 1 || p->symmetric == 1))
    return fail(SEMGATE_EINVAL, "k=%d > %d runs as several sweeps: not with accumulate, part_count > 1 or symmetric = 1", p->k, SEMGATE_MAX_K);
  if (p->max_floor_diff < -1) return fail(SEMGATE_EINVAL, "max_floor_diff=%d", p->max_floor_diff);
  if (p->gate_mode != SEMGATE_GATE_FLAG && p->gate_mode != SEMGATE_GATE_MASK) return fail(SEMGATE_EINVAL, "gate_mode=%d", p->gate_mode);
  if (p->cta_group != 0 && p->cta_group != 1 && p->cta_group != 2 && p->cta_group != 4) return fail(SEMGATE_EINVAL, "cta_group=%d", p->cta_group);
  if (p->accumulate != 0 && p->accumulate != 1) return fail(SEMGATE_EINVAL, "accumulate=%d", p->accumulate);
  if (p->symmetric < -1 || p->symmetric > 1) return fail(SEMGATE_EINVAL, "symmetric=%d", p->symmetric);
  if (p->part_count < 0 || p->part_index < 0 || p->part_index >= std::max(p->part_count, 1))
    return fail(SEMGATE_EINVAL, "part %d of %d", p->part_index, p->part_count);
  if (p->part_count > 1 && p->symmetric != 1) return fail(SEMGATE_EINVAL, "part_count > 1 splits a symmetric sweep: set symmetric = 1");
  if (std::isnan(p->similarity_threshold)) return fail(SEMGATE_EINVAL, "similarity_threshold is NaN");
  return 0;
}

// CTA-pair tiles (cta_group::2) halve the database-tile traffic per SM and win a few percent on
// large sweeps; with few query rows half of every 256-row pair tile would be padding.
// A few query rows against a whole database is a GEMV: the bandwidth-built streaming kernel (K6) takes it
// unless the caller pinned a tensor-core tile shape.
bool use_stream_path(semgate_handle_t h, const semgate_topk_params* p, int64_t Q, int32_t d_pad) {
  return p->cta_group == 0 && h->cta_group == 0 && stream_query_fits(Q, d_pad, p->k);
}

int resolve_cg(semgate_handle_t h, const semgate_topk_params* p, int64_t Q) {
  const int want = p->cta_group ? p->cta_group : h->cta_group;
  if (want == 1 || want == 2 || want == 4) return want;
  return Q >= 4096 ? 2 : 1;
}

// Symmetric sweep: possible when the shapes allow it (square problem on CTA-pair tiles, one index space,
// nothing to accumulate into) and wanted unless the caller or the handle said never.
bool sym_shape_ok(semgate_handle_t h, const semgate_topk_params* p, int64_t Q, int64_t N, int32_t d_pad) {
  return Q == N && Q > 256 && resolve_cg(h, p, Q) == 2 && !use_stream_path(h, p, Q, d_pad) && p->accumulate == 0;
}
// Automatic choice: halving the tensor work pays once the sweep is tensor-bound and long enough to amortise the
// two extra launches and the coarser tail of the triangular schedule; short descriptors leave the kernel
// epilogue-bound, where the column direction costs more than the saved MMAs (measured: 5k x 512-d is 1.4x
// SLOWER symmetric, 20k x 4096-d 1.5x faster, 1M x 4096-d 2.0x faster).
bool sym_wanted(semgate_handle_t h, const semgate_topk_params* p, int64_t Q, int32_t d_pad) {
  if (p->symmetric != 0) return p->symmetric == 1;
  if (h->symmetric != 0) return h->symmetric == 1;
  return Q >= 8192 && d_pad >= 1024;
}

// tiles a symmetric schedule computes (this part's share of the triangle)
int64_t sym_tiles_owned(const Schedule& sc) {
  if (sc.tab_runs != nullptr) return sc.tab_tiles;
  int64_t t = 0;
  int sr = 0;
  for (int lo = 0; lo < sc.ntiles; lo += sc.rm, ++sr) {
    if (!sched_owned(sc, sr)) continue;
    const int64_t r = std::min(sc.rm, sc.ntiles - lo), len = sc.ntiles - lo;
    t += r * len - r * (r - 1) / 2;
  }
  return t;
}

// The symmetric schedule of an N x N sweep: a run table (built once per shape, cached with its device copy) for
// small triangles, the super-row formula otherwise.  A table depends on N only through the tile count, so a database
// that grows keyframe by keyframe builds one per 256 keyframes.  `st_upload` == nullptr: the host copy is enough (size
// queries); else the device copy is made on first use with cudaMemcpyAsync on that stream (the host vectors live as
// long as the cache entry) and nothing here synchronises the device.  Evicted tables are freed at semgate_destroy.
int sym_schedule(semgate_handle_t h, int64_t N, int32_t d_pad, const semgate_topk_params* p, Schedule* out, bool want_device,
                 cudaStream_t st_upload) {
  if (!sym_table_wanted(N)) {
    *out = make_schedule(N, N, d_pad, 2, h->sm_count, true, p->part_index, p->part_count);
    return 0;
  }
  const int pc = std::max(p->part_count, 1);
  const int64_t nb = (N + 255) / 256;
  SymTableDev* t = nullptr;
  for (SymTableDev* c : h->sym_tables)
    if (c->nb == nb && c->d_pad == d_pad && c->part_index == p->part_index && c->part_count == pc) { t = c; break; }
  if (!t) {
    t = new (std::nothrow) SymTableDev();
    if (!t) return fail(SEMGATE_ENOMEM, "out of host memory");
    t->nb = nb; t->d_pad = d_pad; t->part_index = p->part_index; t->part_count = pc;
    build_sym_table(N, d_pad, h->sm_count, p->part_index, pc, &t->host);
    if (h->sym_tables.size() >= 256) {              // keep the cache bounded: retire the oldest (freed at destroy: a
      h->sym_retired.push_back(h->sym_tables.front());   // launch in flight may still read its device copy)
      h->sym_tables.erase(h->sym_tables.begin());
    }
    h->sym_tables.push_back(t);
  }
  if (!want_device) { *out = t->host.sc; return 0; }
  if (!t->dev) {
    const size_t b_runs = align256(t->host.runs.size() * sizeof(RunEntry)), b_ub = align256(t->host.unit_begin.size() * sizeof(int)),
                 b_bf = align256(t->host.block_first.size() * sizeof(int));
    DeviceGuard g(h->device);
    cudaError_t e = cudaMalloc(&t->dev, b_runs + b_ub + b_bf);
    if (e != cudaSuccess) { t->dev = nullptr; return cuda_fail(e, "cudaMalloc (run table)"); }
    char* d = static_cast<char*>(t->dev);
    e = cudaMemcpyAsync(d, t->host.runs.data(), t->host.runs.size() * sizeof(RunEntry), cudaMemcpyHostToDevice, st_upload);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + b_runs, t->host.unit_begin.data(), t->host.unit_begin.size() * sizeof(int), cudaMemcpyHostToDevice, st_upload);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + b_runs + b_ub, t->host.block_first.data(), t->host.block_first.size() * sizeof(int), cudaMemcpyHostToDevice, st_upload);
    if (e != cudaSuccess) { cudaFree(t->dev); t->dev = nullptr; return cuda_fail(e, "cudaMemcpyAsync (run table)"); }
    t->upload_stream = st_upload;
    if (cudaEventCreateWithFlags(&t->uploaded, cudaEventDisableTiming) == cudaSuccess) cudaEventRecord(t->uploaded, st_upload);
    t->sc_dev = t->host.sc;
    t->sc_dev.tab_runs = reinterpret_cast<const RunEntry*>(d);
    t->sc_dev.tab_unit_begin = reinterpret_cast<const int*>(d + b_runs);
    t->sc_dev.tab_block_first = reinterpret_cast<const int*>(d + b_runs + b_ub);
  } else if (t->upload_stream != st_upload && t->uploaded) {
    // first used on another stream: order this stream behind the upload (a no-op once it has completed)
    DeviceGuard g(h->device);
    cudaStreamWaitEvent(st_upload, t->uploaded, 0);
  }
  *out = t->sc_dev;
  return 0;
}

// workspace of a symmetric sweep: [partial lists of either schedule | pacing counters of the full schedule |
// symmetric state: its pacing counters, bounds, counts, flag, candidate buffers]
struct SymLayout { size_t partial, sync_full, state, total; };
SymLayout sym_layout(const Schedule& sc_full, const Schedule& sc_sym, int64_t N, int k) {
  SymLayout l;
  l.partial = std::max(topk_partial_bytes(sc_full, 2, k), topk_partial_bytes(sc_sym, 2, k));
  l.sync_full = align256(topk_sync_bytes(sc_full));
  l.state = sym_state_bytes(N, k, topk_sync_bytes(sc_sym));
  l.total = l.partial + l.sync_full + l.state;
  return l;
}

}  // namespace

extern "C" {

int semgate_version(void) { return SEMGATE_VERSION; }
const char* semgate_last_error(void) { return g_err; }
int semgate_pad_dim(int d) { return d <= 0 ? 0 : ((d + 63) / 64) * 64; }

int semgate_create(semgate_handle_t* out, int device) {
  if (!out) return fail(SEMGATE_EINVAL, "out is NULL");
  *out = nullptr;
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(SEMGATE_EINVAL, "device %d out of range (%d visible)", device, ndev);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(SEMGATE_EARCH, "device %d (%s) is compute capability %d.%d; libsemgate is built for sm_100a only and has no fallback",
                device, prop.name, prop.major, prop.minor);
  semgate_ctx* h = new (std::nothrow) semgate_ctx();
  if (!h) return fail(SEMGATE_ENOMEM, "out of host memory");
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->cc_major = prop.major;
  h->cc_minor = prop.minor;
  const char* env = getenv("SEMGATE_CTA_GROUP");
  if (env && (env[0] == '1' || env[0] == '2' || env[0] == '4')) h->cta_group = env[0] - '0';
  env = getenv("SEMGATE_SYMMETRIC");
  if (env && env[0] == '0') h->symmetric = -1;
  if (env && env[0] == '1') h->symmetric = 1;
  DeviceGuard g(device);
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaStreamCreate"); }
  e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { cudaStreamDestroy(h->stream); delete h; return cuda_fail(e, "cudaStreamCreate"); }
  e = cudaMalloc(reinterpret_cast<void**>(&h->flag_dev), 256);
  if (e == cudaSuccess) e = cudaMemset(h->flag_dev, 0, 256);
  if (e != cudaSuccess) { cudaStreamDestroy(h->stream); cudaStreamDestroy(h->copy_stream); delete h; return cuda_fail(e, "cudaMalloc (flag)"); }
  *out = h;
  return 0;
}

int semgate_destroy(semgate_handle_t h) {
  if (!h) return 0;
  DeviceGuard g(h->device);
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  for (cudaEvent_t e : h->chunk_events) cudaEventDestroy(e);
  for (int i = 0; i < kNumBufs; ++i) if (h->buf[i]) cudaFree(h->buf[i]);
  for (auto* v : {&h->sym_tables, &h->sym_retired})
    for (SymTableDev* t : *v) { if (t->uploaded) cudaEventDestroy(t->uploaded); if (t->dev) cudaFree(t->dev); delete t; }
  if (h->flag_dev) cudaFree(h->flag_dev);
  if (h->clk_dev) cudaFree(h->clk_dev);
  for (SweepGraph* gph : h->graphs) { if (gph->exec) cudaGraphExecDestroy(gph->exec); delete gph; }
  for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
  delete h;
  return 0;
}

int semgate_device_info(semgate_handle_t h, int* sm_count, int* cc_major, int* cc_minor) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (sm_count) *sm_count = h->sm_count;
  if (cc_major) *cc_major = h->cc_major;
  if (cc_minor) *cc_minor = h->cc_minor;
  return 0;
}

int semgate_set_option(semgate_handle_t h, const char* name, int64_t value) {
  if (!h || !name) return fail(SEMGATE_EINVAL, "NULL argument");
  if (strcmp(name, "cta_group") == 0) {
    if (value != 0 && value != 1 && value != 2 && value != 4) return fail(SEMGATE_EINVAL, "cta_group must be 0 (auto), 1, 2 or 4");
    h->cta_group = static_cast<int>(value);
    return 0;
  }
  if (strcmp(name, "symmetric") == 0) {
    if (value != 0 && value != -1 && value != 1) return fail(SEMGATE_EINVAL, "symmetric must be 0 (auto), 1 (whenever possible) or -1 (never)");
    h->symmetric = static_cast<int>(value);
    return 0;
  }
  if (strcmp(name, "profile") == 0) {
    h->profile = value != 0;
    return 0;
  }
  if (strcmp(name, "clock_probe") == 0) {
    DeviceGuard g(h->device);
    if (value != 0 && !h->clk_dev) {
      cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&h->clk_dev), kClkCtas * 4 * sizeof(unsigned long long));
      if (e != cudaSuccess) { h->clk_dev = nullptr; return cuda_fail(e, "cudaMalloc (clock probe)"); }
    } else if (value == 0 && h->clk_dev) {
      cudaFree(h->clk_dev);
      h->clk_dev = nullptr;
    }
    h->clk_ctas = 0;
    return 0;
  }
  if (strcmp(name, "k3_dense") == 0) {       // process-wide: K3's dense-list kernel on (default) / off
    set_merge_dense(value != 0);
    return 0;
  }
  return fail(SEMGATE_EINVAL, "unknown option '%s'", name);
}

int64_t semgate_launch_count(semgate_handle_t h) { return h ? h->launches : 0; }

int semgate_last_sweep_mode(semgate_handle_t h, int32_t* out_mode, int64_t* out_tiles) {
  if (!h || !out_mode) return fail(SEMGATE_EINVAL, "NULL argument");
  *out_mode = h->last_mode;
  if (out_tiles) *out_tiles = h->last_tiles;
  if (h->last_mode == 1) {
    DeviceGuard g(h->device);
    uint32_t flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, h->flag_dev, sizeof(flag), cudaMemcpyDeviceToHost, h->last_stream));
    CUDA_TRY(cudaStreamSynchronize(h->last_stream));
    if (flag != 0) *out_mode = 2;
  }
  return 0;
}

int semgate_profile_read(semgate_handle_t h, double* total_ms, int64_t* n_launches) {
  if (!h || !total_ms || !n_launches) return fail(SEMGATE_EINVAL, "NULL argument");
  DeviceGuard g(h->device);
  double sum = 0.0;
  for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
    CUDA_TRY(cudaEventSynchronize(h->prof_events[i + 1]));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->prof_events[i], h->prof_events[i + 1]));
    sum += ms;
  }
  *total_ms = sum;
  *n_launches = static_cast<int64_t>(h->prof_used / 2);
  h->prof_used = 0;
  return 0;
}

int semgate_clock_probe_read(semgate_handle_t h, double* sm_mhz_median, double* sm_mhz_min, double* span_us, int32_t* n_ctas) {
  if (!h || !sm_mhz_median) return fail(SEMGATE_EINVAL, "NULL argument");
  *sm_mhz_median = 0.0;
  if (sm_mhz_min) *sm_mhz_min = 0.0;
  if (span_us) *span_us = 0.0;
  if (n_ctas) *n_ctas = 0;
  if (!h->clk_dev || h->clk_ctas <= 0) return 0;
  DeviceGuard g(h->device);
  std::vector<unsigned long long> v(static_cast<size_t>(kClkCtas) * 4);
  CUDA_TRY(cudaStreamSynchronize(h->last_stream));
  CUDA_TRY(cudaMemcpy(v.data(), h->clk_dev, v.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  std::vector<double> mhz;
  unsigned long long t_lo = ~0ull, t_hi = 0;
  for (int i = 0; i < kClkCtas; ++i) {
    const unsigned long long t0 = v[4 * i], c0 = v[4 * i + 1], t1 = v[4 * i + 2], c1 = v[4 * i + 3];
    if (t0 == 0 || t1 <= t0 || c1 <= c0) continue;       // CTA did not run (smaller grid) or returned at once
    mhz.push_back(static_cast<double>(c1 - c0) / static_cast<double>(t1 - t0) * 1e3);
    t_lo = std::min(t_lo, t0); t_hi = std::max(t_hi, t1);
  }
  if (mhz.empty()) return 0;
  std::sort(mhz.begin(), mhz.end());
  *sm_mhz_median = mhz[mhz.size() / 2];
  if (sm_mhz_min) *sm_mhz_min = mhz.front();
  if (span_us) *span_us = static_cast<double>(t_hi - t_lo) * 1e-3;
  if (n_ctas) *n_ctas = static_cast<int32_t>(mhz.size());
  return 0;
}

int semgate_last_sweep_overflow(semgate_handle_t h, uint32_t* out_flag_dev, semgate_stream_t stream) {
  if (!h || !out_flag_dev) return fail(SEMGATE_EINVAL, "NULL argument");
  DeviceGuard g(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h->last_mode == 1)
    CUDA_TRY(cudaMemcpyAsync(out_flag_dev, h->flag_dev, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  else
    CUDA_TRY(cudaMemsetAsync(out_flag_dev, 0, sizeof(uint32_t), st));
  return 0;
}

// ---------------------------------------------------------------- schedule self-check (no device needed)
int semgate_schedule_check(int64_t Q, int64_t N, int32_t d_pad, int32_t cta_group, int32_t sm_count, int32_t symmetric,
                           int32_t part_index, int32_t part_count, int32_t* out_shape, int64_t* out_tiles) {
  if (Q <= 0 || N <= 0 || d_pad <= 0 || d_pad % 64 != 0 || sm_count < 4 || (cta_group != 1 && cta_group != 2))
    return fail(SEMGATE_EINVAL, "schedule_check: bad arguments");
  if (symmetric && (cta_group != 2 || Q != N)) return fail(SEMGATE_EINVAL, "schedule_check: a symmetric sweep needs Q == N and CTA pairs");
  if (part_count < 0 || part_index < 0 || part_index >= std::max(part_count, 1) || (part_count > 1 && !symmetric))
    return fail(SEMGATE_EINVAL, "schedule_check: bad part %d of %d", part_index, part_count);
  SymTable table;
  Schedule sc;
  if (symmetric == 2) {          // the run table small symmetric sweeps use (host copy)
    build_sym_table(N, d_pad, sm_count, part_index, part_count, &table);
    sc = table.sc;
  } else {
    sc = make_schedule(Q, N, d_pad, cta_group, sm_count, symmetric != 0, part_index, part_count);
  }
  int64_t computed = 0, makespan = 0;
  const int err = schedule_selfcheck(sc, topk_units(cta_group, sm_count), &computed, &makespan, symmetric == 2 ? &table.owner_of_block : nullptr);
  if (out_shape) {
    out_shape[0] = sc.mblocks; out_shape[1] = sc.ntiles; out_shape[2] = sc.rm; out_shape[3] = sc.s_main;
    out_shape[4] = sc.r_last; out_shape[5] = sc.s_last; out_shape[6] = sc.sync_window; out_shape[7] = sc.a_resident;
  }
  if (out_tiles) { out_tiles[0] = computed; out_tiles[1] = makespan; }
  if (err) return fail(SEMGATE_EINVAL, "schedule_check: invariant %d broken (Q=%lld N=%lld d_pad=%d cta_group=%d sym=%d)", err,
                       (long long)Q, (long long)N, d_pad, cta_group, symmetric);
  return 0;
}

// ---------------------------------------------------------------- K1
int semgate_normalize_cast_dtype(semgate_handle_t h, const void* x, int32_t dtype, int64_t n, int32_t d, int64_t ld,
                                 void* out_bf16, int32_t d_pad, semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (dtype != SEMGATE_DTYPE_F32 && dtype != SEMGATE_DTYPE_F16 && dtype != SEMGATE_DTYPE_BF16) return fail(SEMGATE_EINVAL, "normalize_cast: unknown dtype %d", dtype);
  if (n < 0 || d <= 0 || ld < d || d_pad < d || d_pad % 64 != 0) return fail(SEMGATE_EINVAL, "normalize_cast: bad shape n=%lld d=%d ld=%lld d_pad=%d", (long long)n, d, (long long)ld, d_pad);
  if (n == 0) return 0;
  if (!x || !out_bf16) return fail(SEMGATE_EINVAL, "normalize_cast: NULL pointer");
  DeviceGuard g(h->device);
  RC_TRY(launch_normalize_cast_any(x, dtype, n, d, ld, out_bf16, d_pad, static_cast<cudaStream_t>(stream)), "normalize_cast launch");
  h->launches += 1;
  return 0;
}

int semgate_normalize_cast(semgate_handle_t h, const float* x, int64_t n, int32_t d, int64_t ld, void* out_bf16, int32_t d_pad,
                           semgate_stream_t stream) {
  return semgate_normalize_cast_dtype(h, x, SEMGATE_DTYPE_F32, n, d, ld, out_bf16, d_pad, stream);
}

// ---------------------------------------------------------------- K2 + K3
size_t semgate_topk_workspace_bytes(semgate_handle_t h, int64_t Q, int64_t N, int32_t d_pad, const semgate_topk_params* p) {
  if (!h || !p || Q <= 0 || N <= 0 || p->k < 1 || p->k > SEMGATE_MAX_K_TOTAL) return 256;
  const int cg = resolve_cg(h, p, Q);
  Schedule sc = make_schedule(Q, N, d_pad, cg, h->sm_count);
  if (p->k > SEMGATE_MAX_K)   // several passes of <= 64: lists of one pass + the key lists of all passes (the ceilings)
    return align256(align256(topk_workspace_bytes(sc, cg, SEMGATE_MAX_K)) + static_cast<size_t>(Q) * p->k * sizeof(uint64_t));
  size_t need = topk_workspace_bytes(sc, cg, p->k);
  if (use_stream_path(h, p, Q, d_pad)) need = std::max(need, stream_query_workspace_bytes(Q, p->k, h->sm_count));
  if (sym_wanted(h, p, Q, d_pad) && sym_shape_ok(h, p, Q, N, d_pad)) {
    Schedule sc_sym{};
    if (sym_schedule(h, N, d_pad, p, &sc_sym, false, nullptr) == 0) need = std::max(need, sym_layout(sc, sc_sym, N, p->k).total);
  }
  return align256(need);
}

}  // extern "C"

namespace {
// one pass of a k > 64 sweep: the pass's lists are columns out_col .. out_col + k of rows k_total wide; only keys
// below the last key of the pass before are admitted
struct PassInfo { const uint64_t* ceil_keys; int64_t k_total; int out_col; };

int gated_topk_impl(semgate_handle_t h, const void* q_bf16, int64_t Q, const void* db_bf16, int64_t N, int32_t d_pad,
                    const double* q_ts, const double* db_ts, const int32_t* q_floor, const int32_t* db_floor,
                    const semgate_topk_params* p, void* workspace, size_t workspace_bytes, uint64_t* out_keys,
                    float* out_scores, int32_t* out_idx, uint8_t* out_valid, int32_t* out_count,
                    semgate_stream_t stream, const PassInfo* pass) {
  int rc = 0;
  if (Q < 0 || N < 0 || Q > INT32_MAX || N > INT32_MAX) return fail(SEMGATE_EINVAL, "gated_topk: bad sizes Q=%lld N=%lld", (long long)Q, (long long)N);
  if (d_pad <= 0 || d_pad % 64 != 0) return fail(SEMGATE_EINVAL, "gated_topk: d_pad=%d must be a positive multiple of 64", d_pad);
  if ((q_ts == nullptr) != (db_ts == nullptr)) return fail(SEMGATE_EINVAL, "gated_topk: q_ts and db_ts must both be given or both be NULL");
  if (static_cast<uint64_t>(p->db_index_offset) + static_cast<uint64_t>(N) > 0xFFFFFFFFull) return fail(SEMGATE_EINVAL, "gated_topk: global index overflows 32 bits");
  if (Q == 0) return 0;
  if (p->accumulate && !out_keys) return fail(SEMGATE_EINVAL, "gated_topk: accumulate needs out_keys");
  // the seeded lists hold global indices of OTHER database slices; db_floor only covers this one, so their floor
  // flags cannot be computed here: they come from a final semgate_merge_topk with the whole label array
  if (p->accumulate && out_valid && q_floor && db_floor && p->max_floor_diff >= 0)
    return fail(SEMGATE_EINVAL, "gated_topk: accumulate cannot produce out_valid (db_floor covers this slice only); "
                "flag the final lists with semgate_merge_topk and the whole label array");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard g(h->device);
  const int cg = resolve_cg(h, p, Q);
  const int k = p->k;

  MergeLaunch m{};
  m.Q = Q; m.k = k;
  m.keys_out = out_keys; m.scores = out_scores; m.idx = out_idx; m.valid = out_valid; m.count = out_count;
  m.seed_keys = p->accumulate ? out_keys : nullptr;
  if (pass) { m.out_stride = pass->k_total; m.out_col = pass->out_col; m.count_add = pass->out_col > 0 ? 1 : 0; }
  m.q_floor = q_floor; m.db_floor = db_floor; m.floor_index_offset = p->db_index_offset; m.floor_n = N;
  m.max_floor_diff = (q_floor && db_floor) ? p->max_floor_diff : -1;

  if (N == 0) {   // empty database: every list is empty (place_recognition.py:134)
    m.keys_in = nullptr; m.n_lists = 0; m.row_stride = 0; m.list_stride = 0;
    RC_TRY(launch_merge_topk(m, st), "merge_topk launch");
    h->launches += 1;
    return 0;
  }
  if (!q_bf16 || !db_bf16) return fail(SEMGATE_EINVAL, "gated_topk: NULL descriptor pointer");
  if ((reinterpret_cast<uintptr_t>(q_bf16) & 15) || (reinterpret_cast<uintptr_t>(db_bf16) & 15))
    return fail(SEMGATE_EINVAL, "gated_topk: descriptor matrices must be 16-byte aligned");

  const bool gemv = !pass && use_stream_path(h, p, Q, d_pad);
  Schedule sc = make_schedule(Q, N, d_pad, cg, h->sm_count);
  const size_t need = gemv ? stream_query_workspace_bytes(Q, k, h->sm_count) : topk_workspace_bytes(sc, cg, k);
  if (!workspace || workspace_bytes < need)
    return fail(SEMGATE_ENOMEM, "gated_topk: workspace %zu < required %zu bytes", workspace_bytes, need);

  // Symmetric sweep: the queries are the database itself (same rows, stamps and labels), so S = S^T and only
  // the tiles on or above the block diagonal are computed.  Automatic when the arguments alias.
  const bool aliased = q_bf16 == db_bf16 && q_ts == db_ts && q_floor == db_floor && p->db_index_offset == 0;
  bool sym = !pass && sym_wanted(h, p, Q, d_pad) && sym_shape_ok(h, p, Q, N, d_pad) && aliased;
  Schedule sc_sym{};
  SymLayout lay{};
  if (sym) {
    if ((rc = sym_schedule(h, N, d_pad, p, &sc_sym, true, st))) return rc;
    lay = sym_layout(sc, sc_sym, N, k);
    if (workspace_bytes < lay.total) sym = false;     // a caller that sized its workspace for the full sweep only
  }
  if (p->symmetric == 1 && !sym)
    return fail(SEMGATE_EINVAL, "gated_topk: symmetric=1 needs aliased query/database arguments, Q == N > 256, CTA-pair tiles, "
                "no accumulate, and a workspace of semgate_topk_workspace_bytes()");

  TopkLaunch a{};
  a.q_bf16 = q_bf16; a.Q = Q; a.db_bf16 = db_bf16; a.N = N; a.d_pad = d_pad;
  a.q_ts = q_ts; a.db_ts = db_ts;
  a.q_floor = q_floor; a.db_floor = db_floor;
  a.threshold = p->similarity_threshold; a.gap = p->min_time_gap; a.k = k;
  a.max_floor_diff = m.max_floor_diff; a.gate_mode = p->gate_mode;
  a.db_index_offset = p->db_index_offset;
  a.cta_group = cg; a.sm_count = h->sm_count;
  if (pass) { a.ceil_keys = pass->ceil_keys; a.ceil_stride = pass->k_total; }
  if (h->clk_dev && !gemv) {          // the probe covers the sweep's first K2 launch (the symmetric one, if any)
    a.clk = h->clk_dev;
    CUDA_TRY(cudaMemsetAsync(h->clk_dev, 0, kClkCtas * 4 * sizeof(unsigned long long), st));
    h->clk_ctas = std::min(kClkCtas, h->sm_count);
  }
  int launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  if (h->profile && cudaStreamIsCapturing(st, &capturing) != cudaSuccess) { cudaGetLastError(); capturing = cudaStreamCaptureStatusNone; }
  if (h->profile && capturing == cudaStreamCaptureStatusNone) {
    if (h->prof_used + 2 > h->prof_events.size()) {
      cudaEvent_t e0, e1;
      CUDA_TRY(cudaEventCreate(&e0));
      CUDA_TRY(cudaEventCreate(&e1));
      h->prof_events.push_back(e0);
      h->prof_events.push_back(e1);
    }
    ev0 = h->prof_events[h->prof_used];
    ev1 = h->prof_events[h->prof_used + 1];
    h->prof_used += 2;
    CUDA_TRY(cudaEventRecord(ev0, st));
  }
  uint32_t* sym_flag = nullptr;
  if (gemv) {
    RC_TRY(launch_stream_query(a, static_cast<uint64_t*>(workspace), st), "stream_query launch");
    launches = 1;
  } else if (sym) {
    char* ws = static_cast<char*>(workspace);
    TopkLaunch b = a;
    b.state = ws + lay.partial + lay.sync_full;
    RC_TRY(launch_gated_topk(b, sc_sym, static_cast<uint64_t*>(workspace), st, &launches), "gated_topk (symmetric) launch");
    sym_flag = reinterpret_cast<uint32_t*>(ws + lay.partial + lay.sync_full + topk_sync_bytes(sc_sym)) + 2 * N;
    if (p->part_count <= 1) {
      // the full sweep, armed by the overflow flag: a no-op unless some keyframe's candidate buffer ran over
      a.state = ws + lay.partial;
      a.run_if = sym_flag;
      a.clk = nullptr;
      RC_TRY(launch_gated_topk(a, sc, static_cast<uint64_t*>(workspace), st, &launches), "gated_topk launch");
    }   // one part of a multi-GPU sweep: the caller reads the flag (semgate_last_sweep_mode) and decides with its peers
  } else {
    RC_TRY(launch_gated_topk(a, sc, static_cast<uint64_t*>(workspace), st, &launches), "gated_topk launch");
  }
  if (ev1) CUDA_TRY(cudaEventRecord(ev1, st));
  h->launches += launches;

  m.keys_in = static_cast<const uint64_t*>(workspace);
  m.list_stride = k;
  if (gemv) {            // [Q][blocks][k]
    m.n_lists = stream_query_lists(h->sm_count);
    m.row_stride = static_cast<int64_t>(m.n_lists) * k;
  } else {               // per-row offsets follow the schedule
    m.row_stride = 0;
    m.n_lists = -1; m.sc = sc; m.rows_per_mblock = 128 * cg;   // cg = CTAs (128-row query blocks) per schedule unit
    if (sym) {
      m.sc_sym = sc_sym; m.sym_flag = sym_flag; m.sym_cnt = sym_flag - N; m.sym_flag_copy = h->flag_dev;
      m.sym_force = p->part_count > 1 ? 1 : 0;
      m.sym_cap = sym_capacity(k, N);
      m.sym_ovf = reinterpret_cast<const uint64_t*>(static_cast<char*>(workspace) + lay.partial + lay.sync_full +
                                                    sym_zeroed_bytes(N, topk_sync_bytes(sc_sym)));
    }
  }
  m.zero_ptr = h->fl_zero; m.zero_words = h->fl_zero_words; m.pdl = h->fl_pdl ? 1 : 0;
  RC_TRY(launch_merge_topk(m, st), "merge_topk launch");
  h->launches += 1;
  h->last_mode = sym ? 1 : 0;
  h->last_stream = st;
  h->last_tiles = gemv ? 0 : static_cast<int64_t>(sc.mblocks) * sc.ntiles;
  if (sym) h->last_tiles = sym_tiles_owned(sc_sym);
  return 0;
}
}  // namespace

extern "C" {

int semgate_gated_topk(semgate_handle_t h, const void* q_bf16, int64_t Q, const void* db_bf16, int64_t N, int32_t d_pad,
                       const double* q_ts, const double* db_ts, const int32_t* q_floor, const int32_t* db_floor,
                       const semgate_topk_params* p, void* workspace, size_t workspace_bytes, uint64_t* out_keys,
                       float* out_scores, int32_t* out_idx, uint8_t* out_valid, int32_t* out_count,
                       semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  int rc = check_params(p);
  if (rc) return rc;
  if (p->k <= SEMGATE_MAX_K)
    return gated_topk_impl(h, q_bf16, Q, db_bf16, N, d_pad, q_ts, db_ts, q_floor, db_floor, p, workspace, workspace_bytes, out_keys,
                           out_scores, out_idx, out_valid, out_count, stream, nullptr);
  // k > 64 (the reference's k is any integer, place_recognition.py:853,888): ceil(k / 64) sweeps.  A row's list in
  // shared memory holds 64 candidates; pass p keeps the best 64 among the candidates BELOW the last key pass p-1
  // kept (keys are unique and totally ordered, so the passes tile the sorted candidate list exactly), and the merge
  // kernel writes them into columns 64p.. of the k-wide outputs.  A row whose pass came back short has no more
  // candidates: its ceiling becomes 0 and later passes admit nothing for it.
  if (Q < 0 || N < 0 || Q > INT32_MAX || N > INT32_MAX) return fail(SEMGATE_EINVAL, "gated_topk: bad sizes Q=%lld N=%lld", (long long)Q, (long long)N);
  if (Q == 0) return 0;
  const int k_total = p->k;
  const size_t need = semgate_topk_workspace_bytes(h, Q, std::max<int64_t>(N, 1), d_pad, p);
  if (N > 0 && (!workspace || workspace_bytes < need)) return fail(SEMGATE_ENOMEM, "gated_topk: workspace %zu < required %zu bytes", workspace_bytes, need);
  // the ceilings live in the key lists of all passes: the caller's out_keys, else the tail of the workspace
  const int cg = resolve_cg(h, p, Q);
  const size_t pass_ws = align256(topk_workspace_bytes(make_schedule(Q, std::max<int64_t>(N, 1), d_pad, cg, h->sm_count), cg, SEMGATE_MAX_K));
  uint64_t* keys_all = out_keys ? out_keys : (N > 0 ? reinterpret_cast<uint64_t*>(static_cast<char*>(workspace) + pass_ws) : nullptr);
  semgate_topk_params pp = *p;
  pp.symmetric = -1;
  for (int col = 0; col < k_total; col += SEMGATE_MAX_K) {
    pp.k = std::min(SEMGATE_MAX_K, k_total - col);
    PassInfo pass{(col > 0 && keys_all) ? keys_all + (col - 1) : nullptr, k_total, col};
    rc = gated_topk_impl(h, q_bf16, Q, db_bf16, N, d_pad, q_ts, db_ts, q_floor, db_floor, &pp, workspace, pass_ws, keys_all, out_scores,
                         out_idx, out_valid, out_count, stream, &pass);
    if (rc) return rc;
  }
  return 0;
}

int semgate_merge_topk(semgate_handle_t h, const uint64_t* keys_in, int32_t G, int64_t Q, int32_t k, const int32_t* q_floor,
                       const int32_t* db_floor_all, int32_t max_floor_diff, uint64_t* out_keys, float* out_scores,
                       int32_t* out_idx, uint8_t* out_valid, int32_t* out_count, semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (G < 0 || Q < 0 || k < 1 || k > SEMGATE_MAX_K) return fail(SEMGATE_EINVAL, "merge_topk: bad sizes G=%d Q=%lld k=%d", G, (long long)Q, k);
  if (Q == 0) return 0;
  if (G > 0 && !keys_in) return fail(SEMGATE_EINVAL, "merge_topk: keys_in is NULL");
  DeviceGuard g(h->device);
  MergeLaunch m{};
  m.keys_in = keys_in; m.Q = Q; m.k = k;
  m.row_stride = k; m.list_stride = Q * k; m.n_lists = G;
  m.keys_out = out_keys; m.scores = out_scores; m.idx = out_idx; m.valid = out_valid; m.count = out_count;
  m.q_floor = q_floor; m.db_floor = db_floor_all; m.floor_index_offset = 0;
  m.max_floor_diff = (q_floor && db_floor_all) ? max_floor_diff : -1;
  RC_TRY(launch_merge_topk(m, static_cast<cudaStream_t>(stream)), "merge_topk launch");
  h->launches += 1;
  return 0;
}

// merge of per-GPU lists read in place from the peers' memory (NVLink P2P): no gathered copy
int semgate_merge_topk_peers(semgate_handle_t h, const uint64_t* const* peer_keys, int32_t G, int64_t Q, int32_t k,
                             const int32_t* q_floor, const int32_t* db_floor_all, int32_t max_floor_diff, uint64_t* out_keys,
                             float* out_scores, int32_t* out_idx, uint8_t* out_valid, int32_t* out_count,
                             semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (G < 1 || Q < 0 || k < 1 || k > SEMGATE_MAX_K) return fail(SEMGATE_EINVAL, "merge_topk_peers: bad sizes G=%d Q=%lld k=%d", G, (long long)Q, k);
  if (Q == 0) return 0;
  if (!peer_keys) return fail(SEMGATE_EINVAL, "merge_topk_peers: peer_keys is NULL");
  DeviceGuard g(h->device);
  MergeLaunch m{};
  m.keys_in = nullptr; m.list_ptrs = peer_keys; m.Q = Q; m.k = k;
  m.row_stride = 0; m.list_stride = 0; m.n_lists = G;
  m.keys_out = out_keys; m.scores = out_scores; m.idx = out_idx; m.valid = out_valid; m.count = out_count;
  m.q_floor = q_floor; m.db_floor = db_floor_all; m.floor_index_offset = 0;
  m.max_floor_diff = (q_floor && db_floor_all) ? max_floor_diff : -1;
  RC_TRY(launch_merge_topk(m, static_cast<cudaStream_t>(stream)), "merge_topk launch");
  h->launches += 1;
  return 0;
}

// The same for a slice of the rows, with the peers' overflow flags folded in: every rank merges only its own rows
// of the G per-GPU lists (1/G of the NVLink reads of the replicated merge) and learns in the same kernel whether any
// rank's symmetric part overflowed (no collective of its own for that).
int semgate_merge_topk_peers_rows(semgate_handle_t h, const uint64_t* const* peer_keys, int32_t G, int64_t Q_total, int32_t k,
                                  int64_t row_begin, int64_t row_count, int64_t flag_offset, const int32_t* q_floor,
                                  const int32_t* db_floor_all, int32_t max_floor_diff, uint64_t* out_keys, float* out_scores,
                                  int32_t* out_idx, uint8_t* out_valid, int32_t* out_count, uint32_t* out_any_flag,
                                  semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (G < 1 || G > 1024 || Q_total < 0 || k < 1 || k > SEMGATE_MAX_K || row_begin < 0 || row_count < 0 || row_begin + row_count > Q_total)
    return fail(SEMGATE_EINVAL, "merge_topk_peers_rows: bad sizes G=%d Q=%lld k=%d rows [%lld, +%lld)", G, (long long)Q_total, k,
                (long long)row_begin, (long long)row_count);
  if (!peer_keys) return fail(SEMGATE_EINVAL, "merge_topk_peers_rows: peer_keys is NULL");
  if (out_any_flag && flag_offset < Q_total * k) return fail(SEMGATE_EINVAL, "merge_topk_peers_rows: the flag word must lie behind the keys");
  DeviceGuard g(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (row_count == 0) {
    // nothing to merge here; the flags are still wanted: one block over zero rows
    if (!out_any_flag) return 0;
  }
  MergeLaunch m{};
  m.keys_in = nullptr; m.list_ptrs = peer_keys; m.Q = row_count; m.k = k;
  m.row_stride = 0; m.list_stride = 0; m.n_lists = G;
  m.row_offset = row_begin; m.flag_offset = flag_offset; m.any_flag_out = out_any_flag;
  m.keys_out = out_keys; m.scores = out_scores; m.idx = out_idx; m.valid = out_valid; m.count = out_count;
  m.q_floor = q_floor; m.db_floor = db_floor_all; m.floor_index_offset = 0;
  m.max_floor_diff = (q_floor && db_floor_all) ? max_floor_diff : -1;
  RC_TRY(launch_merge_topk(m, st), "merge_topk launch");
  h->launches += 1;
  return 0;
}

// ---------------------------------------------------------------- dense similarity (interface parity)
int semgate_similarity_matrix(semgate_handle_t h, const void* q_bf16, int64_t Q, const void* db_bf16, int64_t N, int32_t d_pad,
                              float* out, int64_t ld_out, semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (Q < 0 || N < 0 || Q > INT32_MAX || N > INT32_MAX) return fail(SEMGATE_EINVAL, "similarity_matrix: bad sizes Q=%lld N=%lld", (long long)Q, (long long)N);
  if (d_pad <= 0 || d_pad % 64 != 0) return fail(SEMGATE_EINVAL, "similarity_matrix: d_pad=%d must be a positive multiple of 64", d_pad);
  if (Q == 0 || N == 0) return 0;
  if (!q_bf16 || !db_bf16 || !out) return fail(SEMGATE_EINVAL, "similarity_matrix: NULL pointer");
  if (ld_out < N) return fail(SEMGATE_EINVAL, "similarity_matrix: ld_out=%lld < N=%lld", (long long)ld_out, (long long)N);
  if ((reinterpret_cast<uintptr_t>(q_bf16) & 15) || (reinterpret_cast<uintptr_t>(db_bf16) & 15))
    return fail(SEMGATE_EINVAL, "similarity_matrix: descriptor matrices must be 16-byte aligned");
  DeviceGuard g(h->device);
  const int cg = 1;
  Schedule sc = make_schedule(Q, N, d_pad, cg, h->sm_count);
  TopkLaunch a{};
  a.q_bf16 = q_bf16; a.Q = Q; a.db_bf16 = db_bf16; a.N = N; a.d_pad = d_pad;
  a.threshold = 0.f; a.gap = 0.0; a.k = 1; a.max_floor_diff = -1; a.gate_mode = 0; a.db_index_offset = 0;
  a.cta_group = cg; a.sm_count = h->sm_count;
  a.dense = out; a.dense_ld = ld_out;
  int launches = 0;
  RC_TRY(launch_gated_topk(a, sc, nullptr, static_cast<cudaStream_t>(stream), &launches), "similarity_matrix launch");
  h->launches += launches;
  return 0;
}

// ---------------------------------------------------------------- K4
size_t semgate_compact_workspace_bytes(int64_t Q) { return align256(compact_workspace_bytes(Q < 0 ? 0 : Q)); }

static int compact_impl(semgate_handle_t h, const float* scores, const int32_t* idx, const uint8_t* valid, const int32_t* count,
                        int64_t Q, int32_t k, bool valid_only, int64_t q_offset, int32_t* out_query_idx, int32_t* out_match_idx,
                        float* out_similarity, uint8_t* out_is_valid, int64_t* out_total, void* workspace,
                        semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (Q < 0 || k < 1 || !out_total) return fail(SEMGATE_EINVAL, "compact: bad arguments");
  if (Q > 0 && (!scores || !idx || !valid || !count || !out_query_idx || !out_match_idx || !out_similarity || !out_is_valid || !workspace))
    return fail(SEMGATE_EINVAL, "compact: NULL pointer");
  DeviceGuard g(h->device);
  if (q_offset < 0 || q_offset + Q > INT32_MAX) return fail(SEMGATE_EINVAL, "compact: query index offset out of range");
  const bool handoff = h->fl_zero != nullptr && h->fl_zero == workspace;     // the one-call sweep: K3 has zeroed the state
  RC_TRY(launch_compact(scores, idx, valid, count, Q, k, valid_only, q_offset, out_query_idx, out_match_idx, out_similarity, out_is_valid,
                        out_total, workspace, static_cast<cudaStream_t>(stream), handoff, handoff && h->fl_pdl), "compact launch");
  h->launches += Q > 0 ? (handoff ? 1 : 2) : 0;      // (state memset +) the one-pass kernel
  return 0;
}

int semgate_compact(semgate_handle_t h, const float* scores, const int32_t* idx, const uint8_t* valid, const int32_t* count,
                    int64_t Q, int32_t k, int32_t* out_query_idx, int32_t* out_match_idx, float* out_similarity,
                    uint8_t* out_is_valid, int64_t* out_total, void* workspace, semgate_stream_t stream) {
  return compact_impl(h, scores, idx, valid, count, Q, k, false, 0, out_query_idx, out_match_idx, out_similarity, out_is_valid,
                      out_total, workspace, stream);
}

int semgate_compact_valid(semgate_handle_t h, const float* scores, const int32_t* idx, const uint8_t* valid, const int32_t* count,
                          int64_t Q, int32_t k, int32_t* out_query_idx, int32_t* out_match_idx, float* out_similarity,
                          uint8_t* out_is_valid, int64_t* out_total, void* workspace, semgate_stream_t stream) {
  return compact_impl(h, scores, idx, valid, count, Q, k, true, 0, out_query_idx, out_match_idx, out_similarity, out_is_valid,
                      out_total, workspace, stream);
}

int semgate_compact_rows(semgate_handle_t h, const float* scores, const int32_t* idx, const uint8_t* valid, const int32_t* count,
                         int64_t Q, int32_t k, int64_t query_index_offset, int32_t valid_only, int32_t* out_query_idx,
                         int32_t* out_match_idx, float* out_similarity, uint8_t* out_is_valid, int64_t* out_total, void* workspace,
                         semgate_stream_t stream) {
  return compact_impl(h, scores, idx, valid, count, Q, k, valid_only != 0, query_index_offset, out_query_idx, out_match_idx,
                      out_similarity, out_is_valid, out_total, workspace, stream);
}

// ---------------------------------------------------------------- match statistics
size_t semgate_stats_workspace_bytes(void) { return align256(stats_workspace_bytes()); }

int semgate_candidate_stats(semgate_handle_t h, const float* similarity, const uint8_t* is_valid, const int64_t* total_dev,
                            int64_t M, void* workspace, double* out_stats, semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (M < 0 || !workspace || !out_stats) return fail(SEMGATE_EINVAL, "candidate_stats: bad arguments");
  if ((M > 0 || total_dev) && (!similarity || !is_valid)) return fail(SEMGATE_EINVAL, "candidate_stats: NULL pointer");
  DeviceGuard g(h->device);
  RC_TRY(launch_candidate_stats(similarity, is_valid, total_dev, M, workspace, out_stats, static_cast<cudaStream_t>(stream)),
         "candidate_stats launch");
  h->launches += 1;
  return 0;
}

// ---------------------------------------------------------------- gate over pairs
int semgate_gate_candidates(semgate_handle_t h, const int32_t* floor_labels, int64_t n_labels, const int32_t* query_idx,
                            const int32_t* match_idx, int64_t M, int32_t max_floor_diff, uint8_t* out_is_valid,
                            uint64_t* out_counts, semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (M < 0 || n_labels < 0 || max_floor_diff < 0 || !out_counts) return fail(SEMGATE_EINVAL, "gate_candidates: bad arguments");
  if (M > 0 && (!floor_labels || !query_idx || !match_idx || !out_is_valid)) return fail(SEMGATE_EINVAL, "gate_candidates: NULL pointer");
  DeviceGuard g(h->device);
  RC_TRY(launch_gate_candidates(floor_labels, n_labels, query_idx, match_idx, M, max_floor_diff, out_is_valid,
                                reinterpret_cast<unsigned long long*>(out_counts), static_cast<cudaStream_t>(stream)),
         "gate_candidates launch");
  h->launches += M > 0 ? 1 : 0;
  return 0;
}

// ---------------------------------------------------------------- spatial radius join
size_t semgate_spatial_workspace_bytes(int64_t n) { return align256(spatial_workspace_bytes(n < 0 ? 0 : n)); }

int semgate_spatial_count(semgate_handle_t h, const double* positions, int64_t n, double radius, int64_t min_index_gap,
                          void* workspace, int64_t* out_total, semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (n < 0 || n > INT32_MAX || !out_total || !(radius >= 0.0)) return fail(SEMGATE_EINVAL, "spatial_count: bad arguments");
  if (n > 0 && (!positions || !workspace)) return fail(SEMGATE_EINVAL, "spatial_count: NULL pointer");
  DeviceGuard g(h->device);
  RC_TRY(launch_spatial_count(positions, n, radius, min_index_gap, workspace, out_total, static_cast<cudaStream_t>(stream)),
         "spatial_count launch");
  h->launches += n > 0 ? 4 : 0;
  return 0;
}

int semgate_spatial_fill(semgate_handle_t h, const double* positions, int64_t n, double radius, int64_t min_index_gap,
                         const void* workspace, int32_t* out_i, int32_t* out_j, double* out_dist, int64_t capacity,
                         semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (n < 0 || capacity < 0) return fail(SEMGATE_EINVAL, "spatial_fill: bad arguments");
  if (n == 0 || capacity == 0) return 0;
  if (!positions || !workspace || !out_i || !out_j) return fail(SEMGATE_EINVAL, "spatial_fill: NULL pointer");
  DeviceGuard g(h->device);
  RC_TRY(launch_spatial_fill(positions, n, radius, min_index_gap, workspace, out_i, out_j, out_dist, capacity,
                             static_cast<cudaStream_t>(stream)), "spatial_fill launch");
  h->launches += 1;
  return 0;
}

// ---------------------------------------------------------------- K5 re-rank
int semgate_rerank_scores(semgate_handle_t h, const void* local_feats, int64_t n_feat, int32_t P, int32_t dl_pad,
                          const int32_t* query_idx, const int32_t* match_idx, const float* global_sim, int64_t M,
                          float* out_cross, float* out_combined, semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (M < 0 || n_feat < 0 || n_feat > INT32_MAX) return fail(SEMGATE_EINVAL, "rerank_scores: bad sizes");
  if (M == 0) return 0;
  if (P < 1 || P > 4096) return fail(SEMGATE_EINVAL, "rerank_scores: P=%d outside 1..4096", P);
  if (dl_pad <= 0 || dl_pad % 64 != 0) return fail(SEMGATE_EINVAL, "rerank_scores: dl_pad=%d must be a positive multiple of 64", dl_pad);
  if (!query_idx || !match_idx || !global_sim || !out_cross || !out_combined) return fail(SEMGATE_EINVAL, "rerank_scores: NULL pointer");
  if (n_feat > 0 && (!local_feats || (reinterpret_cast<uintptr_t>(local_feats) & 15)))
    return fail(SEMGATE_EINVAL, "rerank_scores: local_feats must be a 16-byte aligned device pointer");
  DeviceGuard g(h->device);
  if (n_feat == 0) {
    // no features at all: every pair falls back to its global score; the kernel handles it through
    // the out-of-range path, but a tensor map needs a valid base -> use the index array as a stand-in
    local_feats = query_idx;
  }
  RC_TRY(launch_rerank(local_feats, static_cast<int>(std::max<int64_t>(n_feat, 1)), P, dl_pad, query_idx, match_idx, global_sim, M,
                       out_cross, out_combined, h->sm_count, static_cast<cudaStream_t>(stream)), "rerank launch");
  h->launches += 1;
  return 0;
}

int semgate_rerank_select(semgate_handle_t h, const int32_t* cand_idx, const float* combined, const int32_t* count,
                          int64_t Q, int32_t kc, int32_t top_k, int32_t* out_idx, float* out_score, int32_t* out_count,
                          semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (Q < 0 || kc < 1 || kc > 64 || top_k < 1 || top_k > 64) return fail(SEMGATE_EINVAL, "rerank_select: bad sizes");
  if (Q == 0) return 0;
  if (!cand_idx || !combined || !count || !out_idx || !out_score || !out_count) return fail(SEMGATE_EINVAL, "rerank_select: NULL pointer");
  DeviceGuard g(h->device);
  RC_TRY(launch_rerank_select(cand_idx, combined, count, Q, kc, top_k, out_idx, out_score, out_count,
                              static_cast<cudaStream_t>(stream)), "rerank_select launch");
  h->launches += 1;
  return 0;
}

// ---------------------------------------------------------------- one call, device-resident: K2 + K3 + K4
namespace {
struct DeviceSweepLayout { size_t topk, sc, ix, va, ct, cws, total; };
DeviceSweepLayout device_sweep_layout(semgate_handle_t h, int64_t n, int32_t d_pad, const semgate_topk_params* p) {
  DeviceSweepLayout l{};
  const size_t nk = static_cast<size_t>(n) * p->k;
  l.topk = align256(semgate_topk_workspace_bytes(h, n, n, d_pad, p));
  l.sc = l.topk; l.ix = l.sc + align256(4 * nk); l.va = l.ix + align256(4 * nk); l.ct = l.va + align256(nk);
  l.cws = l.ct + align256(4 * static_cast<size_t>(n));
  l.total = l.cws + semgate_compact_workspace_bytes(n);
  return l;
}
}  // namespace

size_t semgate_find_loop_closures_device_workspace_bytes(semgate_handle_t h, int64_t n, int32_t d_pad, const semgate_topk_params* p) {
  if (!h || !p || n <= 0 || p->k < 1 || p->k > SEMGATE_MAX_K_TOTAL) return 256;
  return device_sweep_layout(h, n, d_pad, p).total;
}

int semgate_find_loop_closures_device(semgate_handle_t h, const void* x_bf16, int64_t n, int32_t d_pad, const double* ts,
                                      const int32_t* floor_labels, const semgate_topk_params* p, void* workspace,
                                      size_t workspace_bytes, int32_t* out_query_idx, int32_t* out_match_idx,
                                      float* out_similarity, uint8_t* out_is_valid, int64_t* out_total_dev, int32_t use_graph,
                                      semgate_stream_t stream) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  int rc = check_params(p);
  if (rc) return rc;
  if (!out_total_dev) return fail(SEMGATE_EINVAL, "find_loop_closures_device: out_total_dev is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard g(h->device);
  if (n < 2) {                                           // place_recognition.py:864
    CUDA_TRY(cudaMemsetAsync(out_total_dev, 0, sizeof(int64_t), st));
    return 0;
  }
  if (!x_bf16 || !out_query_idx || !out_match_idx || !out_similarity || !out_is_valid) return fail(SEMGATE_EINVAL, "find_loop_closures_device: NULL pointer");
  const DeviceSweepLayout l = device_sweep_layout(h, n, d_pad, p);
  if (!workspace || workspace_bytes < l.total) return fail(SEMGATE_ENOMEM, "find_loop_closures_device: workspace %zu < required %zu bytes", workspace_bytes, l.total);
  char* ws = static_cast<char*>(workspace);
  auto run = [&]() -> int {
    semgate_topk_params pa = *p;
    pa.db_index_offset = 0; pa.accumulate = 0; pa.part_index = 0; pa.part_count = 0;
    // K3 zeroes the compaction's look-back state (no memset node between K3 and K4) and both launch as programmatic
    // dependents of the kernel before them; single-pass sweeps only (k > 64 merges once per pass)
    struct Handoff {
      semgate_handle_t h;
      ~Handoff() { h->fl_zero = nullptr; h->fl_zero_words = 0; h->fl_pdl = false; }
    } handoff{h};
    if (p->k <= SEMGATE_MAX_K) {
      h->fl_zero = reinterpret_cast<unsigned long long*>(ws + l.cws);
      h->fl_zero_words = compact_state_words(n);
      h->fl_pdl = env_flag_default("SEMGATE_PDL", true);
    }
    int r = semgate_gated_topk(h, x_bf16, n, x_bf16, n, d_pad, ts, ts, floor_labels, floor_labels, &pa, ws, l.topk, nullptr,
                               reinterpret_cast<float*>(ws + l.sc), reinterpret_cast<int32_t*>(ws + l.ix),
                               reinterpret_cast<uint8_t*>(ws + l.va), reinterpret_cast<int32_t*>(ws + l.ct), st);
    if (r) return r;
    return semgate_compact(h, reinterpret_cast<float*>(ws + l.sc), reinterpret_cast<int32_t*>(ws + l.ix), reinterpret_cast<uint8_t*>(ws + l.va),
                           reinterpret_cast<int32_t*>(ws + l.ct), n, p->k, out_query_idx, out_match_idx, out_similarity, out_is_valid,
                           out_total_dev, ws + l.cws, st);
  };
  // (a capture cannot start on the legacy default stream: such callers stay eager)
  if (!use_graph || h->graphs_broken || st == nullptr || st == cudaStreamLegacy) return run();

  SweepGraphKey key;
  memset(&key, 0, sizeof(key));
  key.x = x_bf16; key.n = n; key.d_pad = d_pad; key.ts = ts; key.floor = floor_labels; key.ws = workspace;
  // field by field: the caller's struct may carry anything in its padding
  key.p.similarity_threshold = p->similarity_threshold; key.p.min_time_gap = p->min_time_gap; key.p.k = p->k;
  key.p.max_floor_diff = p->max_floor_diff; key.p.gate_mode = p->gate_mode; key.p.db_index_offset = p->db_index_offset;
  key.p.cta_group = p->cta_group; key.p.accumulate = p->accumulate; key.p.symmetric = p->symmetric;
  key.p.part_index = p->part_index; key.p.part_count = p->part_count;
  key.ws_bytes = workspace_bytes; key.oq = out_query_idx; key.om = out_match_idx; key.os = out_similarity; key.ov = out_is_valid;
  key.tot = out_total_dev; key.cta_group = h->cta_group; key.symmetric = h->symmetric;
  SweepGraph* gph = nullptr;
  for (SweepGraph* c : h->graphs)
    if (memcmp(&c->key, &key, sizeof(key)) == 0) { gph = c; break; }
  if (!gph) {
    gph = new (std::nothrow) SweepGraph();
    if (!gph) return fail(SEMGATE_ENOMEM, "out of host memory");
    gph->key = key;
    if (h->graphs.size() >= 8) {                       // small cache: drop the oldest
      if (h->graphs.front()->exec) cudaGraphExecDestroy(h->graphs.front()->exec);
      delete h->graphs.front();
      h->graphs.erase(h->graphs.begin());
    }
    h->graphs.push_back(gph);
  }
  if (gph->exec) {                                       // replay: one launch for the whole sweep
    CUDA_TRY(cudaGraphLaunch(gph->exec, st));
    h->launches += gph->launches;
    h->last_mode = gph->mode; h->last_tiles = gph->tiles; h->last_stream = st;
    return 0;
  }
  if (gph->seen++ == 0) return run();                    // first call: eager (sets kernel attributes, uploads schedules)
  // second call with the same arguments: capture the sequence (nothing in it waits for the host)
  const int64_t launches0 = h->launches;
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) { cudaGetLastError(); h->graphs_broken = true; return run(); }
  rc = run();
  e = cudaStreamEndCapture(st, &graph);
  if (rc != 0 || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    h->graphs_broken = true;
    h->launches = launches0;
    return run();
  }
  e = cudaGraphInstantiate(&gph->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { cudaGetLastError(); gph->exec = nullptr; h->graphs_broken = true; h->launches = launches0; return run(); }
  gph->launches = static_cast<int>(h->launches - launches0);
  gph->mode = h->last_mode; gph->tiles = h->last_tiles;
  CUDA_TRY(cudaGraphLaunch(gph->exec, st));
  return 0;
}

// ---------------------------------------------------------------- host-buffer entry points
int semgate_find_loop_closures_host(semgate_handle_t h, const float* descriptors, int64_t n, int32_t d,
                                    const double* timestamps, const int32_t* floor_labels, const semgate_topk_params* p,
                                    int32_t* out_query_idx, int32_t* out_match_idx, float* out_similarity,
                                    uint8_t* out_is_valid, int64_t capacity, int64_t* out_total) {
  return semgate_find_loop_closures_host_dtype(h, descriptors, SEMGATE_DTYPE_F32, n, d, timestamps, floor_labels, p, out_query_idx,
                                               out_match_idx, out_similarity, out_is_valid, capacity, out_total);
}

int semgate_find_loop_closures_host_dtype(semgate_handle_t h, const void* descriptors_any, int32_t dtype, int64_t n, int32_t d,
                                          const double* timestamps, const int32_t* floor_labels, const semgate_topk_params* p,
                                          int32_t* out_query_idx, int32_t* out_match_idx, float* out_similarity,
                                          uint8_t* out_is_valid, int64_t capacity, int64_t* out_total) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (dtype != SEMGATE_DTYPE_F32 && dtype != SEMGATE_DTYPE_F16 && dtype != SEMGATE_DTYPE_BF16) return fail(SEMGATE_EINVAL, "find_loop_closures: unknown dtype %d", dtype);
  const size_t esz = dtype == SEMGATE_DTYPE_F32 ? 4 : 2;           // bytes per descriptor element in host memory
  const char* descriptors = static_cast<const char*>(descriptors_any);
  int rc = check_params(p);
  if (rc) return rc;
  if (!out_total) return fail(SEMGATE_EINVAL, "out_total is NULL");
  *out_total = 0;
  if (n < 2) return 0;                                   // place_recognition.py:864
  if (d <= 0 || !descriptors) return fail(SEMGATE_EINVAL, "find_loop_closures: bad descriptors");
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  const int k = p->k;
  const int d_pad = semgate_pad_dim(d);
  void *dx = nullptr, *dbf = nullptr, *dts = nullptr, *dfl = nullptr, *ws = nullptr, *sc = nullptr, *ix = nullptr, *va = nullptr, *ct = nullptr, *oq = nullptr, *om = nullptr, *os = nullptr, *ov = nullptr, *tot = nullptr, *cws = nullptr;
  if ((rc = h->reserve(B_X, esz * n * d, &dx))) return rc;
  if ((rc = h->reserve(B_BF16, 2ull * n * d_pad, &dbf))) return rc;
  if (timestamps && (rc = h->reserve(B_TS, 8ull * n, &dts))) return rc;
  if (floor_labels && (rc = h->reserve(B_FL, 4ull * n, &dfl))) return rc;
  const size_t nk = static_cast<size_t>(n) * k;
  if ((rc = h->reserve(B_SC, 4 * nk, &sc)) || (rc = h->reserve(B_IX, 4 * nk, &ix)) || (rc = h->reserve(B_VA, nk, &va)) ||
      (rc = h->reserve(B_CT, 4ull * n, &ct)) || (rc = h->reserve(B_OQ, 4 * nk, &oq)) || (rc = h->reserve(B_OM, 4 * nk, &om)) ||
      (rc = h->reserve(B_OS, 4 * nk, &os)) || (rc = h->reserve(B_OV, nk, &ov)) || (rc = h->reserve(B_TOT, 8, &tot)) ||
      (rc = h->reserve(B_CWS, semgate_compact_workspace_bytes(n), &cws)))
    return rc;

  // Pipeline over row chunks: while chunk c+1 crosses PCIe (copy stream: H2D + K1), the compute
  // stream sweeps everything that chunk c completes: queries of chunk c against database rows
  // [0, end of c), and the earlier queries against chunk c's rows, accumulated into the per-query
  // key lists.  Top-k under a total order is an associative merge, so the result equals one sweep.
  // Chunk sizes shrink towards the end: what cannot overlap with PCIe is the work the LAST chunk
  // completes, and that is proportional to its size.
  const int64_t bytes_per_row = static_cast<int64_t>(d) * 4;   // chunks are cut by rows (launch cost per chunk), whatever the element size
  const int64_t approx_chunks = (n * bytes_per_row) / (24ll << 20);
  std::vector<int64_t> bounds;   // chunk c = rows [bounds[c], bounds[c+1])
  bounds.push_back(0);
  if (n < 4096 || approx_chunks < 2 || k > SEMGATE_MAX_K) {   // (k > 64 runs as several whole sweeps: no accumulation)
    bounds.push_back(n);
  } else {
    // Measured at config 2 (tools/e2e_chunks.py): a chunk's sweeps must finish under the next chunk's copy, and every chunk
    // costs the compute stream ~0.1 ms of launches whatever its size, so chunks below ~1 000 rows (16 MB) LENGTHEN the tail:
    // 4,4,4,4,3,3,2,2,1,1 (round 1) 6.76 ms, ...,2,2,2 6.54 ms, this one 6.51 ms; SEMGATE_E2E_CHUNKS overrides (A/B)
    int kWeights[12] = {0, 0, 6, 5, 5, 4, 4, 3, 3, 2, 2, 2};
    int nw = static_cast<int>(std::min<int64_t>(10, std::max<int64_t>(2, approx_chunks)));
    if (const char* e = getenv("SEMGATE_E2E_CHUNKS")) {     // comma-separated weights, first chunk first
      int vals[12], cnt = 0;
      for (const char* q = e; *q && cnt < 12;) { vals[cnt++] = std::max(1, atoi(q)); while (*q && *q != ',') ++q; if (*q == ',') ++q; }
      if (cnt >= 1) {
        for (int i = 0; i < 12; ++i) kWeights[i] = 0;
        for (int i = 0; i < cnt; ++i) kWeights[12 - cnt + i] = vals[i];
        nw = cnt;
      }
    }
    int wsum = 0;
    for (int i = 0; i < nw; ++i) wsum += kWeights[12 - nw + i];
    int acc = 0;
    for (int i = 0; i < nw; ++i) {
      acc += kWeights[12 - nw + i];
      int64_t b = (i == nw - 1) ? n : ((n * acc / wsum + 127) / 128) * 128;   // 128-row aligned cuts
      b = std::min(b, n);
      if (b > bounds.back()) bounds.push_back(b);
    }
    if (bounds.back() != n) bounds.push_back(n);
  }
  const int64_t nchunks = static_cast<int64_t>(bounds.size()) - 1;
  while (static_cast<int64_t>(h->chunk_events.size()) < nchunks) {
    cudaEvent_t ev;
    CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    h->chunk_events.push_back(ev);
  }
  void* keys;
  if ((rc = h->reserve(B_KEYS, 8 * nk, &keys))) return rc;
  // workspace: the largest any sub-sweep needs
  size_t ws_need = 0;
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t r0 = bounds[c], r1 = bounds[c + 1];
    ws_need = std::max(ws_need, semgate_topk_workspace_bytes(h, r1 - r0, r1, d_pad, p));
    if (r0 > 0) ws_need = std::max(ws_need, semgate_topk_workspace_bytes(h, r0, r1 - r0, d_pad, p));
  }
  if ((rc = h->reserve(B_WS, ws_need, &ws))) return rc;
  cudaStream_t cs = h->copy_stream;
  if (dts) CUDA_TRY(cudaMemcpyAsync(dts, timestamps, 8ull * n, cudaMemcpyHostToDevice, st));
  if (dfl) CUDA_TRY(cudaMemcpyAsync(dfl, floor_labels, 4ull * n, cudaMemcpyHostToDevice, st));
  double* d_ts = static_cast<double*>(dts);
  int32_t* d_fl = static_cast<int32_t*>(dfl);
  char* bf = static_cast<char*>(dbf);
  uint64_t* d_keys = static_cast<uint64_t*>(keys);
  for (int64_t ci = 0; ci < nchunks; ++ci) {
    const int64_t r0 = bounds[ci], r1 = bounds[ci + 1], rows = r1 - r0;
    // the copy stream carries nothing but copies, back to back; the chunk is normalised on the compute stream
    char* dxc = static_cast<char*>(dx) + esz * r0 * d;
    CUDA_TRY(cudaMemcpyAsync(dxc, descriptors + esz * r0 * d, esz * rows * d, cudaMemcpyHostToDevice, cs));
    CUDA_TRY(cudaEventRecord(h->chunk_events[ci], cs));
    CUDA_TRY(cudaStreamWaitEvent(st, h->chunk_events[ci], 0));
    if ((rc = semgate_normalize_cast_dtype(h, dxc, dtype, rows, d, d, bf + 2ull * r0 * d_pad, d_pad, st))) return rc;
    semgate_topk_params pa = *p;
    pa.db_index_offset = 0; pa.accumulate = 0;
    if (k > SEMGATE_MAX_K) {      // one chunk: the sweep's passes write the decoded lists themselves
      rc = semgate_gated_topk(h, bf, n, bf, n, d_pad, d_ts, d_ts, d_fl, d_fl, &pa, ws, h->cap[B_WS], d_keys, static_cast<float*>(sc),
                              static_cast<int32_t*>(ix), static_cast<uint8_t*>(va), static_cast<int32_t*>(ct), st);
      if (rc) return rc;
      break;
    }
    rc = semgate_gated_topk(h, bf + 2ull * r0 * d_pad, rows, bf, r1, d_pad, d_ts ? d_ts + r0 : nullptr, d_ts,
                            d_fl ? d_fl + r0 : nullptr, d_fl, &pa, ws, h->cap[B_WS], d_keys + r0 * k, nullptr, nullptr,
                            nullptr, nullptr, st);
    if (rc) return rc;
    if (r0 > 0) {
      semgate_topk_params pb = *p;
      pb.db_index_offset = static_cast<uint32_t>(r0); pb.accumulate = 1;
      rc = semgate_gated_topk(h, bf, r0, bf + 2ull * r0 * d_pad, rows, d_pad, d_ts, d_ts ? d_ts + r0 : nullptr, d_fl,
                              d_fl ? d_fl + r0 : nullptr, &pb, ws, h->cap[B_WS], d_keys, nullptr, nullptr, nullptr,
                              nullptr, st);
      if (rc) return rc;
    }
  }
  // decode the final key lists, apply the floor flag
  if (k <= SEMGATE_MAX_K) {
    rc = semgate_merge_topk(h, d_keys, 1, n, k, d_fl, d_fl, (d_fl != nullptr) ? p->max_floor_diff : -1, nullptr,
                            static_cast<float*>(sc), static_cast<int32_t*>(ix), static_cast<uint8_t*>(va),
                            static_cast<int32_t*>(ct), st);
    if (rc) return rc;
  }
  rc = semgate_compact(h, static_cast<float*>(sc), static_cast<int32_t*>(ix), static_cast<uint8_t*>(va), static_cast<int32_t*>(ct),
                       n, k, static_cast<int32_t*>(oq), static_cast<int32_t*>(om), static_cast<float*>(os),
                       static_cast<uint8_t*>(ov), static_cast<int64_t*>(tot), cws, st);
  if (rc) return rc;
  int64_t total = 0;
  CUDA_TRY(cudaMemcpyAsync(&total, tot, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *out_total = total;
  if (total > capacity) return fail(SEMGATE_ENOMEM, "find_loop_closures: %lld candidates exceed output capacity %lld", (long long)total, (long long)capacity);
  if (total > 0) {
    if (!out_query_idx || !out_match_idx || !out_similarity || !out_is_valid) return fail(SEMGATE_EINVAL, "find_loop_closures: NULL output");
    CUDA_TRY(cudaMemcpyAsync(out_query_idx, oq, 4ull * total, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_match_idx, om, 4ull * total, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_similarity, os, 4ull * total, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_is_valid, ov, 1ull * total, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return 0;
}

int semgate_query_host(semgate_handle_t h, const float* queries, int64_t nq, const float* database, int64_t n, int32_t d,
                       const double* q_ts, const double* db_ts, const semgate_topk_params* p, float* out_scores,
                       int32_t* out_idx, int32_t* out_count) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  int rc = check_params(p);
  if (rc) return rc;
  if (nq < 0 || n < 0 || d <= 0) return fail(SEMGATE_EINVAL, "query: bad sizes");
  if (nq == 0) return 0;
  if (!queries || !out_scores || !out_idx || !out_count) return fail(SEMGATE_EINVAL, "query: NULL pointer");
  const int k = p->k;
  if (n == 0) {                                          // place_recognition.py:134
    for (int64_t i = 0; i < nq * k; ++i) { out_scores[i] = -INFINITY; out_idx[i] = -1; }
    for (int64_t i = 0; i < nq; ++i) out_count[i] = 0;
    return 0;
  }
  if (!database) return fail(SEMGATE_EINVAL, "query: NULL database");
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  const int d_pad = semgate_pad_dim(d);
  void *dx = nullptr, *dbf = nullptr, *qbf = nullptr, *dts = nullptr, *ws = nullptr, *sc = nullptr, *ix = nullptr, *ct = nullptr;
  if ((rc = h->reserve(B_X, sizeof(float) * std::max(n, nq) * d, &dx))) return rc;
  if ((rc = h->reserve(B_BF16, 2ull * n * d_pad, &dbf))) return rc;
  if ((rc = h->reserve(B_QBF16, 2ull * nq * d_pad, &qbf))) return rc;
  const bool use_time = q_ts && db_ts;
  if (use_time && (rc = h->reserve(B_TS, 8ull * (n + nq), &dts))) return rc;
  const size_t wsb = semgate_topk_workspace_bytes(h, nq, n, d_pad, p);
  const size_t nk = static_cast<size_t>(nq) * k;
  if ((rc = h->reserve(B_WS, wsb, &ws)) || (rc = h->reserve(B_SC, 4 * nk, &sc)) || (rc = h->reserve(B_IX, 4 * nk, &ix)) ||
      (rc = h->reserve(B_CT, 4ull * nq, &ct)))
    return rc;
  CUDA_TRY(cudaMemcpyAsync(dx, database, sizeof(float) * n * d, cudaMemcpyHostToDevice, st));
  if ((rc = semgate_normalize_cast(h, static_cast<float*>(dx), n, d, d, dbf, d_pad, st))) return rc;
  CUDA_TRY(cudaMemcpyAsync(dx, queries, sizeof(float) * nq * d, cudaMemcpyHostToDevice, st));
  if ((rc = semgate_normalize_cast(h, static_cast<float*>(dx), nq, d, d, qbf, d_pad, st))) return rc;
  double* d_dbts = nullptr; double* d_qts = nullptr;
  if (use_time) {
    d_dbts = static_cast<double*>(dts); d_qts = d_dbts + n;
    CUDA_TRY(cudaMemcpyAsync(d_dbts, db_ts, 8ull * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_qts, q_ts, 8ull * nq, cudaMemcpyHostToDevice, st));
  }
  semgate_topk_params pp = *p;
  pp.max_floor_diff = -1;
  rc = semgate_gated_topk(h, qbf, nq, dbf, n, d_pad, d_qts, d_dbts, nullptr, nullptr, &pp, ws, h->cap[B_WS], nullptr,
                          static_cast<float*>(sc), static_cast<int32_t*>(ix), nullptr, static_cast<int32_t*>(ct), st);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(out_scores, sc, 4 * nk, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out_idx, ix, 4 * nk, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out_count, ct, 4ull * nq, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

int semgate_gate_candidates_host(semgate_handle_t h, const int32_t* floor_labels, int64_t n_labels, const int32_t* query_idx,
                                 const int32_t* match_idx, int64_t M, int32_t max_floor_diff, uint8_t* out_is_valid,
                                 uint64_t* out_counts) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (!out_counts || M < 0 || n_labels < 0 || max_floor_diff < 0) return fail(SEMGATE_EINVAL, "gate_candidates: bad arguments");
  out_counts[0] = out_counts[1] = out_counts[2] = 0;
  if (M == 0) return 0;
  if (!floor_labels || !query_idx || !match_idx || !out_is_valid) return fail(SEMGATE_EINVAL, "gate_candidates: NULL pointer");
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  int rc;
  void *dfl = nullptr, *dq = nullptr, *dm = nullptr, *dv = nullptr, *dc = nullptr;
  if ((rc = h->reserve(B_FL, 4ull * std::max<int64_t>(n_labels, 1), &dfl)) || (rc = h->reserve(B_OQ, 4ull * M, &dq)) ||
      (rc = h->reserve(B_OM, 4ull * M, &dm)) || (rc = h->reserve(B_OV, 1ull * M, &dv)) || (rc = h->reserve(B_TOT, 32, &dc)))
    return rc;
  CUDA_TRY(cudaMemcpyAsync(dfl, floor_labels, 4ull * n_labels, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dq, query_idx, 4ull * M, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dm, match_idx, 4ull * M, cudaMemcpyHostToDevice, st));
  rc = semgate_gate_candidates(h, static_cast<int32_t*>(dfl), n_labels, static_cast<int32_t*>(dq), static_cast<int32_t*>(dm), M,
                               max_floor_diff, static_cast<uint8_t*>(dv), static_cast<uint64_t*>(dc), st);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(out_is_valid, dv, 1ull * M, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out_counts, dc, 24, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (out_counts[2] != 0) return fail(SEMGATE_EINDEX, "gate_candidates: %llu candidate indices fall outside the %lld floor labels", (unsigned long long)out_counts[2], (long long)n_labels);
  return 0;
}

int semgate_spatial_candidates_host(semgate_handle_t h, const double* positions, int64_t n, double radius,
                                    int64_t min_index_gap, int32_t* out_i, int32_t* out_j, double* out_dist,
                                    int64_t capacity, int64_t* out_total) {
  if (!h) return fail(SEMGATE_EINVAL, "handle is NULL");
  if (!out_total || n < 0) return fail(SEMGATE_EINVAL, "spatial_candidates: bad arguments");
  *out_total = 0;
  if (n == 0) return 0;
  if (!positions) return fail(SEMGATE_EINVAL, "spatial_candidates: NULL positions");
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  int rc;
  void *dp = nullptr, *ws = nullptr, *tot = nullptr;
  if ((rc = h->reserve(B_X, 24ull * n, &dp)) || (rc = h->reserve(B_WS, semgate_spatial_workspace_bytes(n), &ws)) ||
      (rc = h->reserve(B_TOT, 32, &tot)))
    return rc;
  CUDA_TRY(cudaMemcpyAsync(dp, positions, 24ull * n, cudaMemcpyHostToDevice, st));
  if ((rc = semgate_spatial_count(h, static_cast<double*>(dp), n, radius, min_index_gap, ws, static_cast<int64_t*>(tot), st))) return rc;
  int64_t total = 0;
  CUDA_TRY(cudaMemcpyAsync(&total, tot, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *out_total = total;
  if (total == 0) return 0;
  if (total > capacity || !out_i || !out_j)
    return fail(SEMGATE_ENOMEM, "spatial_candidates: %lld pairs exceed output capacity %lld", (long long)total, (long long)capacity);
  void *di, *dj, *dd = nullptr;
  if ((rc = h->reserve(B_OQ, 4ull * total, &di)) || (rc = h->reserve(B_OM, 4ull * total, &dj))) return rc;
  if (out_dist && (rc = h->reserve(B_SC, 8ull * total, &dd))) return rc;
  if ((rc = semgate_spatial_fill(h, static_cast<double*>(dp), n, radius, min_index_gap, ws, static_cast<int32_t*>(di),
                                 static_cast<int32_t*>(dj), static_cast<double*>(dd), total, st)))
    return rc;
  CUDA_TRY(cudaMemcpyAsync(out_i, di, 4ull * total, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out_j, dj, 4ull * total, cudaMemcpyDeviceToHost, st));
  if (out_dist) CUDA_TRY(cudaMemcpyAsync(out_dist, dd, 8ull * total, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
