// torch.ops.semgate.* — the thin PyTorch extension over the C ABI of libsemgate (include/semgate.h).
//
// Nothing is computed here: every op checks its tensors, allocates the outputs and the workspace with
// torch's caching allocator, and calls the C entry point on torch's current CUDA stream.  The ops exist
// so that a maintainer of the reference can call the path from torch code without ctypes:
//   normalize_cast  <- desc / (norm + 1e-8)                       (place_recognition.py:186-187, :169-170)
//   gated_topk      <- X X^T + mask + argsort[:k] + threshold     (place_recognition.py:190, :882-899; :140-154)
//   merge_topk      <- the per-GPU list merge behind a row-sharded database
//   compact         <- the PlaceMatch append loop                 (place_recognition.py:890-909)
// There is no CPU or other-architecture path: semgate_create fails on anything but compute capability 10.x.
#include <torch/library.h>
#include <ATen/ATen.h>
#include <c10/cuda/CUDAStream.h>
#include <c10/cuda/CUDAGuard.h>

#include <cmath>
#include <mutex>
#include <tuple>
#include <vector>

#include "../../include/semgate.h"

namespace {

using at::Tensor;

semgate_handle_t handle_for(int device) {
  static std::mutex mu;
  static std::vector<semgate_handle_t> handles;
  std::lock_guard<std::mutex> lock(mu);
  if (static_cast<size_t>(device) >= handles.size()) handles.resize(device + 1, nullptr);
  if (!handles[device]) {
    const int rc = semgate_create(&handles[device], device);
    TORCH_CHECK(rc == 0, "semgate_create(device ", device, ") failed (", rc, "): ", semgate_last_error());
  }
  return handles[device];
}

void ok(int rc, const char* what) { TORCH_CHECK(rc == 0, what, " failed (", rc, "): ", semgate_last_error()); }

void* stream_of(const Tensor& t) { return c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

void want(const Tensor& t, at::ScalarType dtype, int64_t dim, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (there is no CPU path)");
  TORCH_CHECK(t.scalar_type() == dtype, name, " must be ", dtype, ", got ", t.scalar_type());
  TORCH_CHECK(t.dim() == dim, name, " must be ", dim, "-d");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}

template <typename T>
const T* opt_ptr(const c10::optional<Tensor>& t, at::ScalarType dtype, int64_t rows, const Tensor& like, const char* name) {
  if (!t.has_value() || !t->defined()) return nullptr;
  want(*t, dtype, 1, name);
  TORCH_CHECK(t->get_device() == like.get_device(), name, " lives on another device");
  TORCH_CHECK(t->size(0) >= rows, name, " is shorter than its matrix");
  return static_cast<const T*>(t->data_ptr());
}

// x fp32 / fp16 / bf16 [N, D] (row stride free) -> bf16 [N, pad64(D)], rows x / (||x|| + 1e-8).  Half-precision rows (an
// extractor under autocast, place_recognition.py:291-297) are widened exactly: same bits as the call on their fp32 image.
Tensor normalize_cast(const Tensor& x) {
  const auto st = x.scalar_type();
  TORCH_CHECK(x.is_cuda() && (st == at::kFloat || st == at::kHalf || st == at::kBFloat16) && x.dim() == 2 && x.stride(1) == 1,
              "normalize_cast: x must be a CUDA fp32 / fp16 / bf16 [N, D] tensor with unit column stride");
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t n = x.size(0);
  const int d = static_cast<int>(x.size(1));
  const int dp = semgate_pad_dim(d);
  const int32_t dtype = st == at::kFloat ? SEMGATE_DTYPE_F32 : st == at::kHalf ? SEMGATE_DTYPE_F16 : SEMGATE_DTYPE_BF16;
  Tensor out = at::empty({n, dp}, x.options().dtype(at::kBFloat16));
  if (n > 0)
    ok(semgate_normalize_cast_dtype(handle_for(x.get_device()), x.data_ptr(), dtype, n, d, x.stride(0), out.data_ptr(), dp, stream_of(x)),
       "semgate_normalize_cast_dtype");
  return out;
}

// -> (scores f32 [Q,k] descending / -inf padded, idx i32 [Q,k] global / -1 padded, valid u8 [Q,k], count i32 [Q], keys i64 [Q,k])
std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor> gated_topk(
    const Tensor& q_bf16, const Tensor& db_bf16, const c10::optional<Tensor>& q_floor, const c10::optional<Tensor>& db_floor,
    const c10::optional<Tensor>& q_ts, const c10::optional<Tensor>& db_ts, double min_time_gap, double threshold, int64_t k,
    int64_t max_floor_diff, int64_t gate_mode, int64_t db_index_offset) {
  want(q_bf16, at::kBFloat16, 2, "q_bf16");
  want(db_bf16, at::kBFloat16, 2, "db_bf16");
  TORCH_CHECK(q_bf16.get_device() == db_bf16.get_device(), "gated_topk: queries and database on different devices");
  const int64_t Q = q_bf16.size(0), N = db_bf16.size(0);
  const int dp = static_cast<int>(q_bf16.size(1));
  TORCH_CHECK(N == 0 || db_bf16.size(1) == dp, "gated_topk: query and database descriptor lengths differ");
  TORCH_CHECK(k >= 1 && k <= SEMGATE_MAX_K_TOTAL, "gated_topk: k outside 1..", SEMGATE_MAX_K_TOTAL);
  c10::cuda::CUDAGuard guard(q_bf16.device());
  semgate_topk_params p{};
  p.similarity_threshold = static_cast<float>(threshold);   // compared in fp32, like numpy's weak-scalar rule
  p.min_time_gap = min_time_gap;
  p.k = static_cast<int32_t>(k);
  p.max_floor_diff = static_cast<int32_t>(max_floor_diff);
  p.gate_mode = static_cast<int32_t>(gate_mode);
  p.db_index_offset = static_cast<uint32_t>(db_index_offset);
  const auto* qf = opt_ptr<int32_t>(q_floor, at::kInt, Q, q_bf16, "q_floor");
  const auto* df = opt_ptr<int32_t>(db_floor, at::kInt, N, q_bf16, "db_floor");
  const auto* qt = opt_ptr<double>(q_ts, at::kDouble, Q, q_bf16, "q_ts");
  const auto* dt = opt_ptr<double>(db_ts, at::kDouble, N, q_bf16, "db_ts");
  auto o = q_bf16.options();
  Tensor scores = at::empty({Q, k}, o.dtype(at::kFloat)), idx = at::empty({Q, k}, o.dtype(at::kInt)),
         valid = at::empty({Q, k}, o.dtype(at::kByte)), count = at::empty({Q}, o.dtype(at::kInt)),
         keys = at::empty({Q, k}, o.dtype(at::kLong));
  if (Q == 0) return {scores, idx, valid, count, keys};
  semgate_handle_t h = handle_for(q_bf16.get_device());
  const size_t wsb = semgate_topk_workspace_bytes(h, Q, N, dp, &p);
  Tensor ws = at::empty({static_cast<int64_t>(wsb)}, o.dtype(at::kByte));
  ok(semgate_gated_topk(h, q_bf16.data_ptr(), Q, N ? db_bf16.data_ptr() : nullptr, N, dp, qt, dt, qf, df, &p, ws.data_ptr(), wsb,
                        reinterpret_cast<uint64_t*>(keys.data_ptr<int64_t>()), scores.data_ptr<float>(), idx.data_ptr<int32_t>(),
                        valid.data_ptr<uint8_t>(), count.data_ptr<int32_t>(), stream_of(q_bf16)),
     "semgate_gated_topk");
  return {scores, idx, valid, count, keys};
}

// keys i64 [G,Q,k] (the gated_topk keys of G database shards, gathered) -> merged (scores, idx, valid, count, keys)
std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor> merge_topk(const Tensor& keys_gathered, const c10::optional<Tensor>& q_floor,
                                                              const c10::optional<Tensor>& db_floor_all, int64_t max_floor_diff) {
  want(keys_gathered, at::kLong, 3, "keys_gathered");
  c10::cuda::CUDAGuard guard(keys_gathered.device());
  const int64_t G = keys_gathered.size(0), Q = keys_gathered.size(1), k = keys_gathered.size(2);
  TORCH_CHECK(k >= 1 && k <= SEMGATE_MAX_K, "merge_topk: k outside 1..", SEMGATE_MAX_K);
  const auto* qf = opt_ptr<int32_t>(q_floor, at::kInt, Q, keys_gathered, "q_floor");
  const auto* df = opt_ptr<int32_t>(db_floor_all, at::kInt, 0, keys_gathered, "db_floor_all");
  auto o = keys_gathered.options();
  Tensor scores = at::empty({Q, k}, o.dtype(at::kFloat)), idx = at::empty({Q, k}, o.dtype(at::kInt)),
         valid = at::empty({Q, k}, o.dtype(at::kByte)), count = at::empty({Q}, o.dtype(at::kInt)),
         keys = at::empty({Q, k}, o.dtype(at::kLong));
  if (Q > 0)
    ok(semgate_merge_topk(handle_for(keys_gathered.get_device()), reinterpret_cast<const uint64_t*>(keys_gathered.data_ptr<int64_t>()),
                          static_cast<int32_t>(G), Q, static_cast<int32_t>(k), qf, df, static_cast<int32_t>(max_floor_diff),
                          reinterpret_cast<uint64_t*>(keys.data_ptr<int64_t>()), scores.data_ptr<float>(), idx.data_ptr<int32_t>(),
                          valid.data_ptr<uint8_t>(), count.data_ptr<int32_t>(), stream_of(keys_gathered)),
       "semgate_merge_topk");
  return {scores, idx, valid, count, keys};
}

// padded lists -> flat candidates in the reference's order (query ascending, similarity descending); the
// first total[0] entries of each output are live.  -> (query_idx i32, match_idx i32, similarity f32, is_valid u8, total i64 [1])
std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor> compact(const Tensor& scores, const Tensor& idx, const Tensor& valid,
                                                           const Tensor& count) {
  want(scores, at::kFloat, 2, "scores");
  want(idx, at::kInt, 2, "idx");
  want(valid, at::kByte, 2, "valid");
  want(count, at::kInt, 1, "count");
  const int64_t Q = scores.size(0), k = scores.size(1);
  TORCH_CHECK(idx.sizes() == scores.sizes() && valid.sizes() == scores.sizes() && count.size(0) == Q, "compact: shapes differ");
  c10::cuda::CUDAGuard guard(scores.device());
  auto o = scores.options();
  const int64_t cap = std::max<int64_t>(Q * k, 1);
  Tensor oq = at::empty({cap}, o.dtype(at::kInt)), om = at::empty({cap}, o.dtype(at::kInt)), os = at::empty({cap}, o.dtype(at::kFloat)),
         ov = at::empty({cap}, o.dtype(at::kByte)), total = at::zeros({1}, o.dtype(at::kLong));
  if (Q > 0) {
    Tensor ws = at::empty({static_cast<int64_t>(semgate_compact_workspace_bytes(Q))}, o.dtype(at::kByte));
    ok(semgate_compact(handle_for(scores.get_device()), scores.data_ptr<float>(), idx.data_ptr<int32_t>(), valid.data_ptr<uint8_t>(),
                       count.data_ptr<int32_t>(), Q, static_cast<int32_t>(k), oq.data_ptr<int32_t>(), om.data_ptr<int32_t>(),
                       os.data_ptr<float>(), ov.data_ptr<uint8_t>(), total.data_ptr<int64_t>(), ws.data_ptr(), stream_of(scores)),
       "semgate_compact");
  }
  return {oq, om, os, ov, total};
}

}  // namespace

TORCH_LIBRARY(semgate, m) {
  m.def("normalize_cast(Tensor x) -> Tensor");
  m.def("gated_topk(Tensor q_bf16, Tensor db_bf16, Tensor? q_floor, Tensor? db_floor, Tensor? q_ts, Tensor? db_ts, "
        "float min_time_gap, float threshold, int k, int max_floor_diff, int gate_mode, int db_index_offset) "
        "-> (Tensor, Tensor, Tensor, Tensor, Tensor)");
  m.def("merge_topk(Tensor keys_gathered, Tensor? q_floor, Tensor? db_floor_all, int max_floor_diff) "
        "-> (Tensor, Tensor, Tensor, Tensor, Tensor)");
  m.def("compact(Tensor scores, Tensor idx, Tensor valid, Tensor count) -> (Tensor, Tensor, Tensor, Tensor, Tensor)");
}

TORCH_LIBRARY_IMPL(semgate, CUDA, m) {
  m.impl("normalize_cast", normalize_cast);
  m.impl("gated_topk", gated_topk);
  m.impl("merge_topk", merge_topk);
  m.impl("compact", compact);
}
