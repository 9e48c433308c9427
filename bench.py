"""Benchmark of the gated loop-closure retrieval hot path (BASELINE.json metric:
gated similarity pairs/s + queries/s (top-25) at 1/2/4/8 B200; % of bf16 tensor peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE config 2, MixVPR shape): the all-pairs loop-closure sweep with floor gate over
n keyframes x 4096-d, 20 000^2 gated pairs PER GPU: n = 20 000 on one GPU (the configuration the metric is
quoted on), n ~ 20 000 * sqrt(N) on N GPUs (28 160 / 39 936 / 56 320 at N = 2 / 4 / 8), where the GPUs split the
triangle of similarity tiles -> per-GPU work is fixed, "weak".  A step = one pass of the hot path over resident,
already normalised bf16 descriptors: fused tcgen05 sweep (K2) + list merge (K3) [+ N>1: one cross-GPU barrier and
the merge of every rank's own rows over NVLink peer memory] + candidate compaction (K4).  `value` = query-database
pairs gated per second, whole job.  `e2e` = the same metric through the host-buffer API (fp32 descriptors in pinned
host memory -> candidates back in host memory: H2D, normalise, sweep, compaction, D2H all timed).

Both arms (this one and `--impl reference`) read the SAME host arrays: `semgate.synthetic.make_case(n, 4096, 3,
seed=0)`.  In the same run, on every N (the numbers are only reported beside a green parity block):
  parity          sampled query rows of the GPU lists against the CPU oracle on those arrays (north-star rule:
                  scores within 2e-3, sets equal up to the threshold / k-th-score boundary), window and floor
                  decisions bit-exact on every returned pair, candidate count equal to a row-sharded full-matrix sweep;
  north_star_c5   BASELINE config 5, the 1M-keyframe x 4096-d all-pairs sweep (16 floors), strong scaling: every rank
                  holds the matrix, the ranks split the triangle; with its own parity block (64 sampled rows against
                  an fp32 torch reference of the same op, bit-exact decisions, count == row-sharded full-matrix sweep);
and on one GPU also
  sustained       >= 2 s of back-to-back headline steps with the clock record (the power-capped regime);
  companions      K1 / K3 / K4 / gate kernels against the measured copy bandwidth;
  c1              BASELINE config 1 (5k x 512-d, the reference's CPU-runnable case) through the one-call, CUDA-graph
                  replayed entry point: step latency and parity against the oracle on the whole problem;
  cpu_baseline    the oracle port on all host threads (bounded sample) and, where /root/reference exists, the
                  unmodified reference function itself (else the committed measurement is cited).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "multi-level-indoor-slam_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

N_Q = 20000          # query keyframes
N_DB_PER_GPU = 20000  # database keyframes per GPU
DIM = 4096           # MixVPR descriptor length (place_recognition.py:197)
TOPK = 25
THRESHOLD = 0.5      # SemanticPlaceRecognition default (place_recognition.py:817)
MIN_TIME_GAP = 10.0  # (place_recognition.py:818)
NUM_FLOORS = 3
METRIC = "gated_similarity_pairs_per_s_top25"
UNIT = "pairs/s"
STRONG = False       # --workload c3/c4/c5: fixed problem, database rows (or the triangle) split over the ranks
ALLPAIRS = False     # default workload at N>1: all-pairs sweep, the ranks split the triangle of tiles (weak scaling)
WORKLOAD_NAME = "BASELINE configs[1]: MixVPR-shape 4096-d, 20k-keyframe all-pairs loop-closure sweep with floor gate"
HOST_SEED = 0        # semgate.synthetic seed of the host arrays both arms read


def set_workload(name: str):
    """Default (c2) is the configuration the metric is quoted on; the others are for DESIGN.md numbers."""
    global N_Q, N_DB_PER_GPU, DIM, NUM_FLOORS, STRONG, WORKLOAD_NAME
    if name == "c2":
        return
    if name == "c5":
        N_Q, N_DB_PER_GPU, DIM, NUM_FLOORS, STRONG = 1_000_000, 1_000_000, 4096, 16, True
        WORKLOAD_NAME = "BASELINE configs[4]: 1M-keyframe multi-floor database, full gated top-k sweep, triangle of tiles split over the GPUs"
    elif name == "c4":
        N_Q, N_DB_PER_GPU, DIM, NUM_FLOORS, STRONG = 8_192, 250_000, 49_152, 4, True
        WORKLOAD_NAME = "BASELINE configs[3]: AnyLoc-shape 49152-d VLAD, 250k database (rows sharded across the GPUs) x 8192-query batch"
    elif name == "c3":
        N_Q, N_DB_PER_GPU, DIM, NUM_FLOORS, STRONG = 10_000, 100_000, 8448, 4, True
        WORKLOAD_NAME = "BASELINE configs[2]: SALAD-shape 8448-d, 100k database (rows sharded across the GPUs) x 10k query batch, exclusion window"
    elif name == "c1":
        N_Q, N_DB_PER_GPU, DIM, NUM_FLOORS = 5_000, 5_000, 512, 3
        WORKLOAD_NAME = "BASELINE configs[0]: 5k keyframes x 512-d, 3 floors (the reference's CPU-runnable case)"
    else:
        raise SystemExit(f"unknown workload {name}")


def allpairs_keyframes(n_gpus: int) -> int:
    """Keyframes of the all-pairs sweep whose pair count is 20 000^2 per GPU (equal row shards of whole
    128-row blocks, so the e2e leg's all-gather of normalised rows is regular)."""
    if n_gpus <= 1:
        return 20000
    g = 128 * n_gpus
    return int(round(20000.0 * (n_gpus ** 0.5) / g)) * g


def enable_allpairs(n_gpus: int):
    global N_Q, N_DB_PER_GPU, STRONG, ALLPAIRS, WORKLOAD_NAME
    N_Q = N_DB_PER_GPU = allpairs_keyframes(n_gpus)
    STRONG = ALLPAIRS = True        # every rank holds the whole matrix; the work (not the rows) is split
    WORKLOAD_NAME += f" -- {n_gpus} GPUs: all-pairs sweep over {N_Q} keyframes (20000^2 pairs per GPU)"


def db_total(n_gpus: int) -> int:
    return N_DB_PER_GPU if STRONG else N_DB_PER_GPU * n_gpus


def workload_config(n_gpus: int):
    return {
        "workload": WORKLOAD_NAME,
        "queries": N_Q, "database_per_gpu": db_total(n_gpus) // n_gpus, "database_total": db_total(n_gpus), "dim": DIM,
        "top_k": TOPK, "similarity_threshold": THRESHOLD, "min_time_gap_s": MIN_TIME_GAP, "floors": NUM_FLOORS,
        "gate": "strict floor gate, flag mode (reference order)",
        "inputs": f"semgate.synthetic.make_case(n, {DIM}, {NUM_FLOORS}, seed={HOST_SEED}): the same host arrays in both arms",
        "sharding": ("none" if n_gpus == 1 else
                     f"triangle of similarity tiles x{n_gpus} (every rank holds all rows; each similarity is computed once, on one GPU)"
                     if (ALLPAIRS or (STRONG and N_Q == db_total(n_gpus))) else f"db-rows x{n_gpus}"),
        "pairs_per_gpu": float(N_Q) * db_total(n_gpus) / n_gpus,
        "l2": f"inputs larger than L2 ({2 * db_total(n_gpus) * DIM / 1e6:.0f} MB of bf16 rows per GPU vs 126 MB L2); no explicit flush",
    }


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            j = json.load(open(p))
            return (float(j["bf16_tflops"]), float(j.get("hbm_gbs", 0.0)), float(j.get("bf16_tflops_sustained", 0.0)),
                    "measured (MEASURED_PEAKS.json, burst)")
        except Exception:
            pass
    return 1590.0, 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / power / throttle reasons sampled through NVML every ~2 ms DURING the timed region
    (the default timed region is tens of milliseconds; `nvidia-smi -lms` cannot sample faster than 100 ms)."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.running = False
        self.thread = None
        self.nv = None
        self.h = None

    def start(self):
        try:
            if os.environ.get("SEMGATE_BENCH_NO_NVML"):      # A/B aid: does the sampler itself disturb the run?
                raise RuntimeError("disabled")
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[self.index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return self
        self.running = True
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()
        return self

    def _pump(self):
        nv = self.nv
        while self.running:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                watts = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(mhz), float(watts), float(util), time.perf_counter(), int(mask)))
            except Exception:
                pass
            time.sleep(0.002)

    def window(self, t0: float, t1: float):
        """Summary of the samples taken between two time.perf_counter() stamps (the sampler itself is started before
        the warm-up, so that NVML's first-call costs and its driver locks stay out of the timed region)."""
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        rows = [r for r in list(self.samples) if t0 <= r[3] <= t1]
        note = None
        if not rows:      # an NVML query can take longer than a 20 ms timed region: take the samples right around it
            rows = [r for r in list(self.samples) if t0 - 0.05 <= r[3] <= t1 + 0.05]
            note = "no sample fell inside the timed region; these are the samples within 50 ms of it"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        a = np.array([r[:3] for r in rows])
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
        reasons = sorted({nm for r in rows for nm, bit in names if r[4] & bit})
        out = {"sm_mhz": float(np.median(a[:, 0])), "sm_max_mhz": self.max_mhz, "reasons": reasons,
               "samples": int(a.shape[0]), "power_w": float(np.median(a[:, 1]))}
        if note:
            out["note"] = note
        return out

    def stop(self, t0: float = None, t1: float = None):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.running = False
        self.thread.join(timeout=2)
        return self.window(-1e30 if t0 is None else t0, 1e30 if t1 is None else t1)


# --------------------------------------------------------------------------- inputs (both arms)
def host_case(n: int, dim: int = None, floors: int = None):
    """The workload's host arrays: fp32 descriptors [n, dim], fp64 timestamps, int32 floor labels."""
    from semgate import synthetic
    dim = DIM if dim is None else dim
    floors = NUM_FLOORS if floors is None else floors
    desc, ts, fl = synthetic.make_case(n, dim, floors, seed=HOST_SEED)
    return desc, ts, fl.astype(np.int32)


# --------------------------------------------------------------------------- CPU arm
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its ranks; the CPU arm is meant to use every host core."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)          # process-wide, stays in effect
    except Exception:
        pass
    return n


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 0) for i in threadpool_info() if i.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return os.cpu_count() or 1


def cpu_sweep(desc, ts, fl, rows: int, want_result: bool = False):
    """The reference algorithm (oracle port, numpy/OpenBLAS on all host threads) on the first
    `rows` query keyframes against the whole database.  Returns seconds (and the candidates)."""
    from oracle import semgate_oracle as O
    t0 = time.perf_counter()
    dbn = O.l2_normalize(desc)
    res = O.gated_topk(dbn[:rows], dbn, ts[:rows], ts, fl[:rows], fl, k=TOPK, threshold=THRESHOLD,
                       min_time_gap=MIN_TIME_GAP, max_floor_diff=0, normalize=False, block=1024)
    c = O.compact(res)
    dt = time.perf_counter() - t0
    return (dt, c) if want_result else dt


def calibrate_rows(desc, ts, fl, target_s: float, lo: int = 256):
    """Query rows whose sweep takes about `target_s` seconds here.  Normalising the whole database is a
    fixed cost inside every sweep, so the time is affine in the rows: two probes give slope and offset."""
    cpu_sweep(desc, ts, fl, lo)                       # first call pays for page faults / thread start-up
    t1 = cpu_sweep(desc, ts, fl, lo)
    t2 = cpu_sweep(desc, ts, fl, 3 * lo)
    slope = max((t2 - t1) / (2 * lo), 1e-7)
    fixed = max(t1 - slope * lo, 0.0)
    rows = int((max(target_s, t2) - fixed) / slope)
    return int(min(max(rows, lo), desc.shape[0], N_Q)), t1


def reference_verbatim_block():
    """The unmodified reference function on this box's host cores where /root/reference exists (never on the GPU
    boxes: the reference does not travel); else the committed measurement from the build container."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import reference_verbatim as RV
        r = RV.measure()
        if r is not None:
            r["measured"] = "in this run"
            return r
    except Exception as e:      # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}
    p = os.path.join(ROOT, "profiles", "r02_reference_verbatim.json")
    if os.path.isfile(p):
        try:
            r = json.load(open(p))
            r["measured"] = ("NOT in this run: /root/reference does not exist on this box; figures measured in the build "
                             "container (profiles/r02_reference_verbatim.json, tools/reference_verbatim.py)")
            return r
        except Exception:
            pass
    return {"unavailable": "/root/reference is not present on this machine"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    desc, ts, fl = host_case(N_DB_PER_GPU)
    budget = 150.0 / max(args.steps + args.warmup, 1)
    rows, _ = calibrate_rows(desc, ts, fl, min(8.0, budget))
    for _ in range(args.warmup):
        cpu_sweep(desc, ts, fl, rows)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, cand = cpu_sweep(desc, ts, fl, rows, want_result=True)
    dt = time.perf_counter() - t0
    value = rows * float(N_DB_PER_GPU) * args.steps / dt
    sample = f"first {rows} of {N_Q} query keyframes against the full {N_DB_PER_GPU}-keyframe database per step"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus if ALLPAIRS else 1),
        "queries_per_s": rows * args.steps / dt,
        "candidates_in_sample": int(len(cand["query_idx"])), "valid_in_sample": int(np.asarray(cand["is_valid"]).sum()),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port", "sample": sample,
                         "reference_verbatim": reference_verbatim_block()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference algorithm (oracle/semgate_oracle.py: numpy + OpenBLAS, all host threads) on the same host arrays as "
                "the GPU arm; the reference's own function (an O(N^2) Python loop over an N x N matrix, "
                "place_recognition.py:882-885) is under cpu_baseline.reference_verbatim",
    }
    print(json.dumps(out))


# --------------------------------------------------------------------------- GPU arm: parity helpers
def decisions_exact_on_device(torch, res, row_lo, q_ts, db_ts, q_fl, db_fl, gap, thr, n_db):
    """Bit-exact checks on EVERY returned pair of `res` (rows row_lo.. of the queries): nothing inside the exclusion
    window (fp64), floor flag == (floor_q == floor_m), scores >= threshold and descending, indices in range."""
    k = res.scores.shape[1]
    rows = res.scores.shape[0]
    dev = res.scores.device
    pos = torch.arange(k, device=dev)[None, :]
    filled = pos < res.count[:, None]
    ok = bool(((res.idx >= 0) == filled).all()) and bool((res.scores[filled] >= thr).all())
    ok = ok and bool((res.scores[:, 1:][filled[:, 1:]] <= res.scores[:, :-1][filled[:, 1:]]).all())
    ok = ok and bool((res.idx[filled] < n_db).all()) and bool((res.valid[~filled] == 0).all())
    qi = (torch.arange(rows, device=dev) + row_lo)[:, None].expand(rows, k)[filled]
    mi = res.idx[filled].long()
    window_ok = not bool(((db_ts[mi] - q_ts[qi]).abs() < gap).any())
    floor_ok = bool(((q_fl[qi] == db_fl[mi]) == (res.valid[filled] != 0)).all())
    return ok, window_ok, floor_ok, int(res.count.sum().item())


def torch_reference_rows(torch, q_bf16, db_bf16, rows, q_ts, db_ts, q_fl, db_fl, k, thr, gap):
    """fp32 torch reference of the same op on sampled query rows (the full Q x N matrix of the large configs cannot
    exist): similarities of the bf16 operands accumulated in fp32, fp64 window, top-k, threshold, floor flag -- the
    order of place_recognition.py:882-899.  Returns candidate arrays in the oracle's compact form."""
    dev = q_bf16.device
    r = torch.from_numpy(rows).to(dev)
    q = q_bf16[r].float()
    n = db_bf16.shape[0]
    sims = torch.empty((len(rows), n), dtype=torch.float32, device=dev)
    step = max(1024, min(65536, (1 << 28) // max(db_bf16.shape[1], 1)))
    for s0 in range(0, n, step):
        sims[:, s0:s0 + step] = q @ db_bf16[s0:s0 + step].float().T
    sims[(db_ts[None, :] - q_ts[r][:, None]).abs() < gap] = -float("inf")
    top_s, top_i = torch.topk(sims, k, dim=1)
    keep = torch.isfinite(top_s) & (top_s >= thr)
    valid = q_fl[r][:, None] == db_fl[top_i]
    qq = r[:, None].expand_as(top_i)
    return {"query_idx": qq[keep].cpu().numpy().astype(np.int64), "match_idx": top_i[keep].cpu().numpy().astype(np.int64),
            "similarity": top_s[keep].cpu().numpy(), "is_valid": valid[keep].cpu().numpy()}


def lists_to_candidates(res, local_rows, global_rows):
    from oracle import semgate_oracle as O
    import torch
    rr = torch.from_numpy(local_rows).to(res.scores.device)
    sub = dict(scores=res.scores[rr].cpu().numpy(), idx=res.idx[rr].cpu().numpy().astype(np.int64),
               valid=res.valid[rr].cpu().numpy().astype(bool), count=res.count[rr].cpu().numpy())
    got = O.compact(sub)
    got["query_idx"] = global_rows[got["query_idx"]]
    return got


def parity_block(torch, dist, world, dev, part, ref_fn, q_ts, db_ts, q_fl, db_fl, n_db, count_expected, tol, what, n_rows=64):
    """Parity of this rank's rows (`part`: RowsResult) inside the bench run; all ranks must agree.  `ref_fn(rows)` gives
    the reference candidates of the sampled global rows."""
    import parity
    lo, hi = part.lo, part.hi
    res = part.result
    out = {"reference": what, "rows_sampled_per_rank": 0}
    ok_struct, ok_window, ok_floor, cnt = decisions_exact_on_device(torch, res, lo, q_ts, db_ts, q_fl, db_fl, MIN_TIME_GAP,
                                                                    THRESHOLD, n_db)
    status, err, bdiff = 1, 0.0, 0
    if hi > lo:
        n_rows = min(n_rows, hi - lo)
        rows = np.sort(np.random.default_rng(17 + lo).choice(hi - lo, n_rows, replace=False))
        try:
            rep = parity.compare_candidates(ref_fn(rows + lo), lists_to_candidates(res, rows, rows + lo), TOPK, THRESHOLD, tol=tol)
            err, bdiff = float(rep["max_score_err"]), int(rep["boundary_diffs"])
        except AssertionError as e:
            status = 0
            out["failure"] = str(e)[:300]
        out["rows_sampled_per_rank"] = int(n_rows)
    if not (ok_struct and ok_window and ok_floor):
        status = 0
    t = torch.tensor([status, cnt, bdiff], dtype=torch.int64, device=dev)
    e = torch.tensor([err], dtype=torch.float64, device=dev)
    if world > 1:
        m = t[:1].clone()
        dist.all_reduce(m, op=dist.ReduceOp.MIN)
        s = t[1:].clone()
        dist.all_reduce(s)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        status, cnt, bdiff = int(m[0]), int(s[0]), int(s[1])
    out.update({"max_score_err": float(e[0]), "score_tolerance": tol, "boundary_diffs": bdiff,
                "window_decisions_bit_exact": bool(ok_window), "floor_decisions_bit_exact": bool(ok_floor),
                "lists_well_formed": bool(ok_struct), "candidates": cnt,
                "candidates_row_sharded_full_matrix_sweep": count_expected,
                "count_matches": count_expected is None or cnt == count_expected})
    if count_expected is not None and cnt != count_expected:
        status = 0
    out["parity"] = "ok" if status == 1 else "FAILED"
    return out


# --------------------------------------------------------------------------- GPU arm: legs
def device_rows_bf16(torch, eng, n, dim, seed, dev):
    """normalised bf16 rows generated chunk-wise on the device (the fp32 form of 1M x 4096 rows would be 16 GB);
    the same seed gives the same matrix on every rank."""
    from semgate import _native
    dp = _native.pad_dim(dim)
    g = torch.Generator(device=dev); g.manual_seed(seed)
    places = max(8, n // 20)
    anchors = torch.randn((places, dim), generator=g, device=dev, dtype=torch.float32)
    out = torch.empty((n, dp), dtype=torch.bfloat16, device=dev)
    step = max(1024, min(65536, (1 << 28) // dim))
    for s0 in range(0, n, step):
        e0 = min(n, s0 + step)
        pid = torch.randint(0, places, (e0 - s0,), generator=g, device=dev)
        x = anchors[pid]
        x += 0.6 * torch.randn((e0 - s0, dim), generator=g, device=dev, dtype=torch.float32)
        eng.normalize_cast(x, out=out[s0:e0])
    return out


def k2_roofline(eng, k2_ms, k2_n, Q, n_db_local, n_db_total, dp, peak_tf, peak_src, step_ms):
    """Roofline of the dominant kernel from the library's own CUDA events around K2: FLOPs the tensor cores EXECUTED
    (a symmetric sweep computes only the tiles on or above the block diagonal) over the kernel's duration."""
    k2_avg = k2_ms / max(k2_n, 1)
    sweep_mode, sweep_tiles = eng.last_sweep_mode()   # 0 full, 1 symmetric (every similarity computed once), 2 overflowed
    flops_full = 2.0 * Q * n_db_local * dp
    flops = flops_full
    if sweep_mode == 1:
        nb = (Q + 255) // 256            # sweep_tiles: this rank's share of the nb*(nb+1)/2 tiles
        flops = 2.0 * Q * n_db_total * dp * sweep_tiles / float(nb * nb)
    tf = flops / (k2_avg * 1e-3) / 1e12 if k2_avg > 0 else 0.0
    return {"bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf, "traffic": None,
            "kernel": "gated_topk_kernel (K2)", "kernel_ms": k2_avg,
            "kernel_share_of_step": k2_avg / step_ms if step_ms > 0 else None, "peak_source": peak_src,
            "flops_per_launch": flops, "flops_counted": "executed by the tensor cores",
            "sweep": {0: "full matrix", 1: "symmetric: S_ij = S_ji computed once, gated in both directions",
                      2: "symmetric attempt overflowed, full sweep redone"}[sweep_mode],
            "tiles_per_launch": int(sweep_tiles), "full_matrix_flops": flops_full,
            "full_matrix_equivalent_tflops": flops_full / (k2_avg * 1e-3) / 1e12 if k2_avg > 0 else 0.0}, k2_avg


def in_kernel_clock(torch, eng, step_fn):
    """One extra, untimed step with the clock probe on: clock64 / globaltimer pairs from every CTA of K2."""
    try:
        eng.set_option("clock_probe", 1)
        step_fn()
        torch.cuda.synchronize()
        med, mn, span, ctas = eng.clock_probe_read()
        eng.set_option("clock_probe", 0)
        return {"sm_mhz_median": med, "sm_mhz_min": mn, "span_us": span, "ctas": ctas,
                "how": "clock64 / globaltimer at entry and exit of every CTA of K2, one untimed step"}
    except Exception as e:      # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}


def c5_leg(torch, dist, eng, sr, dev, rank, world, peak_tf, peak_src, steps=3, warmup=2):
    """BASELINE config 5 inside the default run: the 1M-keyframe all-pairs sweep, triangle split over the ranks."""
    from semgate import _native, synthetic
    n, dim, floors = 1_000_000, 4096, 16
    x = device_rows_bf16(torch, eng, n, dim, 1000, dev)
    ts_h = synthetic.make_timestamps(n)
    fl_h = synthetic.make_floors(n, floors).astype(np.int32)
    ts, fl = torch.from_numpy(ts_h).to(dev), torch.from_numpy(fl_h).to(dev)

    def mk(off):
        return _native.make_params(k=TOPK, similarity_threshold=THRESHOLD, min_time_gap=MIN_TIME_GAP, max_floor_diff=0,
                                   gate_mode=_native.GATE_FLAG, db_index_offset=off)

    def step(defer=True):
        return sr.sweep_all_pairs(x, mk, ts=ts, floor=fl, max_floor_diff=0, compact=True, gather=False, defer=defer)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step().result()
    barrier()
    eng.profile_read()
    sampler = ClockSampler(dev.index).start() if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    pend, cur = [], None
    for _ in range(steps):
        if cur is not None:
            pend.append(cur.forget_value())
        cur = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    k2_ms, k2_n = eng.profile_read()
    clocks = sampler.stop() if sampler else None
    out4 = cur.result()
    overflowed = sum(1 for p in pend if p.check()) + (1 if cur.overflowed else 0)
    total = out4[4].clone()
    roof, k2_avg = k2_roofline(eng, k2_ms, k2_n, n, n, n, _native.pad_dim(dim), peak_tf, peak_src, ms)
    how = sr.last_all_pairs if world > 1 else "one GPU"
    if world > 1:
        t = torch.tensor([ms, k2_avg], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, k2_max = float(t[0]), float(t[1])
        dist.all_reduce(total)
    else:
        k2_max = k2_avg
    # ---- parity, untimed: this rank's rows of the merged lists against an fp32 torch reference; the candidate count
    # against a row-sharded FULL-MATRIX sweep (every rank sweeps all queries against its slice of the rows)
    part = sr.sweep_all_pairs(x, mk, ts=ts, floor=fl, max_floor_diff=0, gather=False)
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world

    def mk_full(off):
        p = mk(off)
        p.symmetric = -1
        return p
    rows_part = sr.sweep(x, x[lo:hi], mk_full, lo, q_ts=ts, db_ts_shard=ts[lo:hi], q_floor=fl, db_floor_shard=fl[lo:hi],
                         db_floor_all=fl, max_floor_diff=0, gather=False)
    cnt = rows_part.result.count.sum().to(torch.int64).reshape(1)
    same_lists = torch.tensor([1 if (torch.equal(rows_part.result.idx, part.result.idx)
                                     and torch.equal(rows_part.result.scores, part.result.scores)) else 0], device=dev)
    if world > 1:
        dist.all_reduce(cnt)
        dist.all_reduce(same_lists, op=dist.ReduceOp.MIN)
    ref_fn = lambda rows: torch_reference_rows(torch, x, x, rows, ts, ts, fl, fl, TOPK, THRESHOLD, MIN_TIME_GAP)
    par = parity_block(torch, dist, world, dev, part, ref_fn, ts, ts, fl, fl, n, int(cnt.item()), 3e-4,
                       "fp32 torch reference of the same op on the same bf16 rows (sampled query rows), all ranks")
    par["lists_equal_row_sharded_full_matrix_sweep"] = bool(int(same_lists.item()) == 1)
    if not par["lists_equal_row_sharded_full_matrix_sweep"] or int(total.item()) != int(cnt.item()):
        par["parity"] = "FAILED"
    del x
    torch.cuda.empty_cache()
    return {
        "workload": "BASELINE configs[4]: 1M keyframes x 4096-d, 16 floors, all-pairs gated top-25 sweep",
        "scaling": "strong", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
        "pairs_per_s": float(n) * n / (ms * 1e-3), "queries_per_s": n / (ms * 1e-3), "candidates_per_step": int(total.item()),
        "k2_ms": k2_max, "k2_tflops_executed_per_gpu": roof["achieved"], "frac_of_peak": roof["frac"],
        "full_matrix_equivalent_tflops_per_gpu": roof["full_matrix_equivalent_tflops"], "sweep": roof["sweep"],
        "split": how, "steps_overflowed": overflowed, "clocks": clocks,
        "exchange": "none (one GPU)" if world == 1 else exchange_info(sr, world).get("exchange"),
        "data": "synthetic, generated on the device (same seed on every rank)", "parity": par,
    }


def sustained_leg(torch, eng, step_fn, seconds, flops_per_step, sus_peak_tf, dev_index):
    """>= `seconds` of back-to-back headline steps: the power-capped regime, with its clock record."""
    for _ in range(3):
        step_fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_probe0 = time.perf_counter()
    step_fn(); torch.cuda.synchronize()
    one = max(time.perf_counter() - t_probe0, 1e-4)
    n = int(max(50, min(20000, seconds / one * 1.3)))
    eng.profile_read()
    sampler = ClockSampler(dev_index).start()
    e0.record()
    for _ in range(n):
        step_fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    k2_ms, k2_n = eng.profile_read()
    k2_avg = k2_ms / max(k2_n, 1)
    tf = flops_per_step / (k2_avg * 1e-3) / 1e12 if k2_avg > 0 else 0.0
    return {"seconds": ms * 1e-3, "steps": n, "ms_per_step": ms / n, "k2_ms": k2_avg, "k2_tflops_executed": tf,
            "peak_sustained": sus_peak_tf, "frac_of_sustained_peak": tf / sus_peak_tf if sus_peak_tf else None,
            "pairs_per_s": float(N_Q) * N_Q * n / (ms * 1e-3), "clocks": clocks}


def companions_leg(torch, eng, peak_hbm):
    """The HBM-bound companion kernels against the measured copy bandwidth (CUDA events, median of 10, inputs
    larger than L2)."""
    from semgate import _native

    def timeit(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in evs:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b) for a, b in evs)
        return t[len(t) // 2]

    out = []

    def add(name, shape, ms, byts, **kw):
        d = {"kernel": name, "shape": shape, "ms": ms, "algorithmic_bytes": byts, "gbs": byts / ms / 1e6,
             "frac_of_hbm_peak": byts / ms / 1e6 / peak_hbm}
        d.update(kw)
        out.append(d)
    n, d = 65536, 4096
    x = torch.randn((n, d), device="cuda")
    o = torch.empty((n, d), dtype=torch.bfloat16, device="cuda")
    add("K1 normalize_cast", [n, d], timeit(lambda: eng.normalize_cast(x, out=o)), n * 6 * d)
    del x, o
    Q, k = 1_000_000, TOPK
    keys = torch.randint(1, 2 ** 62, (4, Q, k), device="cuda", dtype=torch.int64)
    add("K3 merge_topk (4 full lists per row, unsorted: a synthetic worst case)", [4, Q, k], timeit(lambda: eng.merge_topk(keys, k)),
        Q * k * (8 * 4 + 9) + 4 * Q)
    res = eng.merge_topk(keys, k)
    keys_sorted = torch.sort(keys, dim=2, descending=True).values.contiguous()
    add("K3 merge_topk (4 full sorted lists per row: the per-GPU lists of a sharded sweep)", [4, Q, k],
        timeit(lambda: eng.merge_topk(keys_sorted, k)), Q * k * (8 * 4 + 9) + 4 * Q)
    same = bool(torch.equal(eng.merge_topk(keys_sorted, k).idx, res.idx))
    out[-1]["equals_unsorted_merge"] = same
    del keys_sorted
    del keys
    add("K4 compact", [Q, k], timeit(lambda: eng.compact(res)), Q * k * (9 + 13) + 4 * Q)
    oq, om, os_, ov, tot = eng.compact(res)
    M = int(tot.item())
    fl = torch.randint(1, 6, (1 << 31 >> 10,), device="cuda", dtype=torch.int32)
    om = om % fl.shape[0]
    oq = oq % fl.shape[0]
    ms = timeit(lambda: eng.gate_candidates(fl, oq[:M], om[:M], 0))
    add("gate_candidates (pairs in the order K4 emits; 2M labels: gathers from global memory)", [M], ms, 9 * M, candidates_per_s=M / ms * 1e3)
    fl20 = fl[:20000].contiguous()                    # a trajectory-sized label table (the reference's 19 163 poses): kept in
    oq20, om20 = oq[:M] % 20000, om[:M] % 20000       # shared memory by every block
    ms = timeit(lambda: eng.gate_candidates(fl20, oq20, om20, 0))
    add("gate_candidates (the same pairs, 20 000 labels: table in shared memory)", [M], ms, 9 * M, candidates_per_s=M / ms * 1e3)
    del fl20, oq20, om20
    del res, oq, om, os_, ov, fl
    # K5: CricaVPR cross-correlation re-rank (place_recognition.py:669-757), DINOv2 shape (529 patches x 768-d), 25 candidates
    # per query -- tensor-bound: reported against the measured bf16 peak, useful FLOPs = 2 P^2 D per pair
    nf, P, Dl, kc, nq = 1000, 529, 768, 25, 2000
    feats = torch.empty((nf, P, _native.pad_dim(Dl)), dtype=torch.bfloat16, device="cuda")
    for s0 in range(0, nf, 250):
        xx = torch.randn((250 * P, Dl), device="cuda")
        eng.normalize_cast(xx, out=feats[s0:s0 + 250].view(250 * P, -1))
    qi = torch.arange(nq, device="cuda", dtype=torch.int32).repeat_interleave(kc) % nf
    mi = torch.randint(0, nf, (nq * kc,), device="cuda", dtype=torch.int32)
    gs = torch.rand((nq * kc,), device="cuda")
    ms = timeit(lambda: eng.rerank_scores(feats, qi, mi, gs), iters=5, warm=2)
    peak_tf = measured_peaks()[0]
    tf = 2.0 * P * P * Dl * nq * kc / ms / 1e9
    out.append({"kernel": "K5 rerank_scores (cross-correlation re-rank, pair form)", "shape": [nq * kc, P, Dl], "ms": ms,
                "pairs_per_s": nq * kc / ms * 1e3, "tflops_useful": tf, "frac_of_tensor_peak": tf / peak_tf,
                "note": "useful FLOPs 2*P*P*D per pair (the tiling multiplies 512 x 544 + 544 x 32 of the 529 x 529 wanted)"})
    return {"hbm_peak_gbs": peak_hbm, "kernels": out}


def c1_leg(torch, eng, peak_tf):
    """BASELINE config 1 (the reference's CPU-runnable case) through the one-call entry point, replayed as a CUDA
    graph: step latency, and parity of the WHOLE problem against the oracle on the same host arrays."""
    from semgate import _native
    from oracle import semgate_oracle as O
    import parity
    n, d, floors = 5000, 512, 3
    desc, ts_h, fl_h = host_case(n, d, floors)
    dev = torch.device("cuda", eng.device)
    x = eng.normalize_cast(torch.from_numpy(desc).to(dev))
    ts, fl = torch.from_numpy(ts_h).to(dev), torch.from_numpy(fl_h).to(dev)
    p = _native.make_params(k=TOPK, similarity_threshold=THRESHOLD, min_time_gap=MIN_TIME_GAP, max_floor_diff=0)
    side = torch.cuda.Stream(device=dev)
    out = {"workload": "BASELINE configs[0]: 5k keyframes x 512-d, 3 floors, all-pairs, top-25"}
    with torch.cuda.stream(side):
        for _ in range(5):
            res = eng.find_loop_closures_device(x, p, ts=ts, floor=fl, use_graph=True)
        side.synchronize()
        reps = 100
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.profile_read()
        l0 = eng.launch_count
        # five batches of 100 replays, median batch (the legs before this one leave the power controller in whatever state
        # their last kernels put it: single batches of this 60 us step were seen between 57 and 71 us on the same code)
        time.sleep(0.5)
        batches = []
        for _ in range(5):
            e0.record(side)
            for _ in range(reps):
                res = eng.find_loop_closures_device(x, p, ts=ts, floor=fl, use_graph=True)
            e1.record(side)
            side.synchronize()
            batches.append(e0.elapsed_time(e1) / reps * 1e3)
        batches.sort()
        out["us_per_step_graph"] = batches[len(batches) // 2]
        out["us_per_step_graph_batches"] = [round(b, 2) for b in batches]
        out["launches_per_step"] = (eng.launch_count - l0) / (5 * reps)
        reps = 200
        e0.record(side)
        for _ in range(reps):
            res3 = eng.compact(eng.gated_topk(x, x, p, q_ts=ts, db_ts=ts, q_floor=fl, db_floor=fl))
        e1.record(side)
        side.synchronize()
        out["us_per_step_three_calls"] = e0.elapsed_time(e1) / reps * 1e3
        k2_ms, k2_n = eng.profile_read()
        out["k2_us"] = k2_ms / max(k2_n, 1) * 1e3
    oq, om, os_, ov, tot = res
    t = int(tot.item())
    got = dict(query_idx=oq[:t].cpu().numpy().astype(np.int64), match_idx=om[:t].cpu().numpy().astype(np.int64),
               similarity=os_[:t].cpu().numpy(), is_valid=ov[:t].cpu().numpy().astype(bool))
    same3 = t == int(res3[4].item()) and torch.equal(om[:t], res3[1][:t]) and torch.equal(os_[:t], res3[2][:t])
    ref = O.find_loop_closures(desc, ts_h, fl_h, similarity_threshold=THRESHOLD, min_time_gap=MIN_TIME_GAP, k=TOPK)
    par = {"reference": "CPU oracle (fp32) on the same host arrays, the whole 5k x 5k problem"}
    try:
        rep = parity.compare_candidates(ref, got, TOPK, THRESHOLD)
        parity.check_decisions_exact(got, ts_h, fl_h, MIN_TIME_GAP, 0)
        parity.check_order(got)
        par.update({"parity": "ok" if same3 else "FAILED", "max_score_err": float(rep["max_score_err"]),
                    "boundary_diffs": int(rep["boundary_diffs"]), "candidates": t, "candidates_oracle": int(len(ref["query_idx"])),
                    "valid": int(got["is_valid"].sum()), "valid_oracle": int(np.asarray(ref["is_valid"]).sum()),
                    "graph_equals_three_call_path": bool(same3)})
    except AssertionError as e:
        par.update({"parity": "FAILED", "failure": str(e)[:300]})
    out["pairs_per_s"] = float(n) * n / (out["us_per_step_graph"] * 1e-6)
    out["queries_per_s"] = n / (out["us_per_step_graph"] * 1e-6)
    out["k2_tflops"] = 2.0 * n * n * d / (out["k2_us"] * 1e-6) / 1e12 if out["k2_us"] > 0 else None
    out["k2_frac_of_peak"] = out["k2_tflops"] / peak_tf if out["k2_tflops"] else None
    out["api"] = "semgate_find_loop_closures_device (K2 + K3 + K4 in one call, CUDA graph replay)"
    out["parity"] = par
    return out


def exchange_info(sr, world):
    if world == 1:
        return {}
    peer = bool(sr._peer_ok)
    d = {"exchange": "peer memory: one barrier, then every rank merges its own rows of the per-GPU lists in place over NVLink "
                     "(overflow flags folded into the merge kernel)" if peer
         else "NCCL all-gather of the per-GPU lists, then a replicated merge"}
    if sr.peer_error:
        d["peer_exchange_unavailable"] = sr.peer_error
    return d


def run_ours(args):
    import torch
    import torch.distributed as dist
    from semgate import _native, synthetic
    from semgate.dist import ShardedRetrieval, RowsResult, shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = _native.get_engine(local)
    if args.cta_group:
        eng.set_option("cta_group", args.cta_group)
    eng.set_option("profile", 1)
    if args.symmetric == "off":
        eng.set_option("symmetric", -1)
    sr = ShardedRetrieval(eng, exchange=args.exchange)
    peak_tf, peak_hbm, sus_tf, peak_src = measured_peaks()
    default_run = args.workload == "c2"

    n_db_total = db_total(world)
    lo, hi = shard_bounds(n_db_total, world, rank)
    dp = _native.pad_dim(DIM)

    # ---- inputs, resident in HBM.  The headline workload reads the host arrays the CPU arm reads; the large
    # configurations are generated on the device (their fp32 form does not fit a host comfortably).
    host = None
    if default_run or args.workload == "c1":
        host = host_case(n_db_total)
        full = eng.normalize_cast(torch.from_numpy(host[0]).to(dev))
        ts_h, fl_h = host[1], host[2]
    else:
        full = device_rows_bf16(torch, eng, n_db_total, DIM, 1000, dev)
        ts_h = synthetic.make_timestamps(n_db_total)
        fl_h = synthetic.make_floors(n_db_total, NUM_FLOORS).astype(np.int32)
    q_bf16 = full[:N_Q]
    db_bf16 = full[lo:hi] if (world > 1 and not (N_Q == n_db_total)) else full
    ts_all = torch.from_numpy(ts_h).to(dev)
    fl_all = torch.from_numpy(fl_h).to(dev)
    q_ts, q_fl = ts_all[:N_Q].contiguous(), fl_all[:N_Q].contiguous()

    def mk(offset):
        return _native.make_params(k=TOPK, similarity_threshold=THRESHOLD, min_time_gap=MIN_TIME_GAP, max_floor_diff=0,
                                   gate_mode=_native.GATE_FLAG, db_index_offset=offset,
                                   symmetric=-1 if args.symmetric == "off" else 0)

    allpairs = N_Q == n_db_total                       # the queries are the database

    def step(defer=True, want_rows=False):
        """One pass of the hot path; returns the flat candidates (oq, om, os, ov, total) of this rank's rows
        (N = 1: all rows), or with want_rows the RowsResult before compaction."""
        if allpairs and world > 1:
            r = sr.sweep_all_pairs(full, mk, ts=ts_all, floor=fl_all, max_floor_diff=0, compact=not want_rows, gather=False,
                                   defer=defer and not want_rows)
            return r
        if world > 1:
            part = sr.sweep(q_bf16, db_bf16, mk, lo, q_ts=q_ts, db_ts_shard=ts_all[lo:hi], q_floor=q_fl,
                            db_floor_shard=fl_all[lo:hi], db_floor_all=fl_all, max_floor_diff=0, gather=False)
            return part if want_rows else eng.compact(part.result, query_offset=part.lo)
        if allpairs:
            res = eng.gated_topk(full, full, mk(0), q_ts=ts_all, db_ts=ts_all, q_floor=fl_all, db_floor=fl_all)
        else:
            res = eng.gated_topk(q_bf16, full, mk(0), q_ts=q_ts, db_ts=ts_all, q_floor=q_fl, db_floor=fl_all)
        return RowsResult(0, N_Q, res) if want_rows else eng.compact(res)

    def resolve(o):
        return o.result() if hasattr(o, "result") else o

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import gc
    warm = max(args.warmup, 3)
    sampler = ClockSampler(local).start() if rank == 0 else None     # before the warm-up: see ClockSampler.window
    for _ in range(warm):
        out = resolve(step())
    barrier()
    eng.profile_read()
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gc.collect()
    gc.disable()                                  # no collector pauses between the enqueues of the timed region
    barrier()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_host0 = time.perf_counter()
    e0.record()
    # (only the newest step's outputs stay alive: holding every step's arrays would send the caching allocator to
    #  cudaMalloc inside the timed region -- measured as random 10-100 ms host stalls)
    pend, out = [], None
    for i in range(args.steps):
        marks[i].record()
        if hasattr(out, "forget_value"):
            pend.append(out.forget_value())
        out = step()
    marks[args.steps].record()
    e1.record()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3
    barrier()
    t_host1 = time.perf_counter()
    gc.enable()
    ms = e0.elapsed_time(e1)
    per_step = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps))
    launches = eng.launch_count - launches0
    k2_ms, k2_n = eng.profile_read()
    clocks = sampler.stop(t_host0, t_host1) if sampler else None
    steps_overflowed = sum(1 for o in pend if o.check())
    last_pending = out
    out = resolve(out)
    steps_overflowed += 1 if getattr(last_pending, "overflowed", False) else 0
    total = out[4].clone()
    roofline, k2_avg = k2_roofline(eng, k2_ms, k2_n, N_Q, (hi - lo) if (world > 1 and not allpairs) else n_db_total, n_db_total, dp,
                                   peak_tf, peak_src, ms / args.steps)
    if world > 1:
        t = torch.tensor([ms, k2_avg], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, k2_avg = float(t[0]), float(t[1])
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
        dist.all_reduce(total)
        roofline["kernel_ms"] = k2_avg
        roofline["kernel_share_of_step"] = k2_avg / (ms / args.steps)
    total_candidates = int(total.item())
    pairs_per_step = float(N_Q) * float(n_db_total)
    value = pairs_per_step * args.steps / (ms * 1e-3)
    roofline["in_kernel_clock"] = in_kernel_clock(torch, eng, lambda: resolve(step()))
    prof = os.path.join(ROOT, "profiles", "k2_sym_traffic.json" if roofline["sweep"].startswith("symmetric") else "k2_traffic.json")
    if os.path.isfile(prof) and default_run and world == 1:   # the committed ncu capture is of exactly this workload
        try:
            j = json.load(open(prof))
            roofline["traffic"] = j.get("dram_bytes_per_launch")
            roofline["traffic_source"] = ("NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of this kernel on "
                                          "this workload from the committed `ncu --set full` capture " + str(j.get("source", prof)))
        except Exception:
            pass

    # ---- parity inside the run (untimed): this rank's rows against the reference, counts against a row-sharded full sweep
    part = step(want_rows=True)
    if hasattr(part, "result") and not isinstance(part, RowsResult):
        part = part.result()

    def mk_full(off):
        p = mk(off)
        p.symmetric = -1
        return p
    if allpairs:
        rows_part = sr.sweep(full, full[lo:hi] if world > 1 else full, mk_full, lo if world > 1 else 0, q_ts=ts_all,
                             db_ts_shard=ts_all[lo:hi] if world > 1 else ts_all, q_floor=fl_all,
                             db_floor_shard=fl_all[lo:hi] if world > 1 else fl_all, db_floor_all=fl_all, max_floor_diff=0, gather=False)
        cnt = rows_part.result.count.sum().to(torch.int64).reshape(1)
        if world > 1:
            dist.all_reduce(cnt)
        count_expected = int(cnt.item())
        del rows_part
    else:
        count_expected = None
    if host is not None:
        from oracle import semgate_oracle as O
        desc_h = host[0]

        def ref_fn(rows):
            r = O.gated_topk(desc_h[rows], desc_h, ts_h[rows], ts_h, fl_h[rows], fl_h, k=TOPK, threshold=THRESHOLD,
                             min_time_gap=MIN_TIME_GAP, max_floor_diff=0)
            c = O.compact(r)
            c["query_idx"] = rows[c["query_idx"]]
            return c
        par = parity_block(torch, dist, world, dev, part, ref_fn, q_ts, ts_all, q_fl, fl_all, n_db_total, count_expected, 2e-3,
                           "CPU oracle (oracle/semgate_oracle.py, fp32 like the reference) on the same host arrays, sampled query rows")
    else:
        ref_fn = lambda rows: torch_reference_rows(torch, q_bf16, full, rows, q_ts, ts_all, q_fl, fl_all, TOPK, THRESHOLD, MIN_TIME_GAP)
        par = parity_block(torch, dist, world, dev, part, ref_fn, q_ts, ts_all, q_fl, fl_all, n_db_total, count_expected, 3e-4,
                           "fp32 torch reference of the same op on the same bf16 rows (sampled query rows), all ranks")
    if count_expected is not None and total_candidates != count_expected:
        par["parity"] = "FAILED"
        par["timed_steps_candidates"] = total_candidates
    del part

    base = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if (STRONG and not ALLPAIRS) else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(world), "queries_per_s": N_Q * args.steps / (ms * 1e-3),
        "candidates_per_step": total_candidates, "steps_overflowed": steps_overflowed, "parity": par,
        "step_ms_distribution": {"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1],
                                 "host_enqueue_ms_per_step": host_enqueue_ms / args.steps,
                                 "note": "per-step CUDA-event intervals inside the timed region (diagnostic; `ms_per_step` is the "
                                         "whole region / steps); a host slower than the GPU shows as host_enqueue > median"},
        "roofline": roofline, "gpu_launches": int(launches), "clocks": clocks,
        "all_pairs_split": sr.last_all_pairs if world > 1 else None,
        "cta_group": args.cta_group or os.environ.get("SEMGATE_CTA_GROUP", "auto (2 for Q >= 4096)"),
    }
    base.update(exchange_info(sr, world))

    if not default_run or args.no_e2e:
        if rank == 0:
            base.update({"cpu_baseline": None, "e2e": None,
                         "note": "non-default workload: the e2e / cpu_baseline / extra legs are only run for the default configuration"})
            print(json.dumps(base))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the public host-buffer API (pinned host memory in, host memory out)
    e2e_steps = max(3, min(args.steps, 10))
    desc_h = host[0]
    cap = N_Q * TOPK
    if world == 1:
        q_host = torch.from_numpy(desc_h).pin_memory()
        outs = tuple(torch.empty((cap,), dtype=dt, pin_memory=True).numpy()
                     for dt in (torch.int32, torch.int32, torch.float32, torch.uint8))
        qh = q_host.numpy()
        ts_pin, fl_pin = torch.from_numpy(ts_h).pin_memory(), torch.from_numpy(fl_h).pin_memory()   # every input pinned
        ts_p, fl_p = ts_pin.numpy(), fl_pin.numpy()
        p = mk(0)

        def e2e_step():
            r = eng.find_loop_closures_host(qh, ts_p, fl_p, p, out=outs)   # semgate_find_loop_closures_host
            return len(r[0])
        h2d = qh.nbytes + ts_h.nbytes + fl_h.nbytes
    else:
        q_host = torch.from_numpy(np.ascontiguousarray(desc_h[lo:hi])).pin_memory()      # this rank's rows of the database
        tsh = torch.from_numpy(ts_h).pin_memory()
        flh = torch.from_numpy(fl_h).pin_memory()
        rows_cap = (hi - lo) * TOPK
        ho = [torch.empty((rows_cap,), dtype=dt, pin_memory=True) for dt in (torch.int32, torch.int32, torch.float32, torch.uint8)]

        def e2e_step():
            # every rank uploads and normalises its rows; the bf16 rows meet over NVLink (all-gather); triangle sweep;
            # every rank merges and compacts its own rows and copies THEM to its host memory
            oq, om, os_, ov, tot = sr.sweep_all_pairs_from_host(q_host, tsh, flh, mk, lo, hi, N_Q, max_floor_diff=0, compact=True,
                                                                gather=False)
            t = int(tot.item())
            for h, d in zip(ho, (oq, om, os_, ov)):
                h[:t].copy_(d[:t], non_blocking=True)
            torch.cuda.synchronize()
            return t
        h2d = N_Q * DIM * 4 + (ts_h.nbytes + fl_h.nbytes) * world   # one fp32 row shard per rank
    for _ in range(max(args.warmup, 10)):         # the first calls after the device-resident legs run 5-8 % slower (same-box A/B,
        n_e2e = e2e_step()                        # tools/e2e_pin_ab.py: 7.0 ms over calls 2-11, 6.5 ms from then on)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        n_e2e = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt[0])
        nn = torch.tensor([n_e2e], device=dev, dtype=torch.int64)
        dist.all_reduce(nn)
        n_e2e = int(nn.item())
    e2e = {"value": pairs_per_step * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(n_e2e * 13 + 8 * world), "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
           "candidates": int(n_e2e), "candidates_match_device_path": bool(n_e2e == total_candidates),
           "api": "semgate_find_loop_closures_host (C ABI, pinned host buffers)" if world == 1 else
                  "semgate python API (ShardedRetrieval.sweep_all_pairs_from_host): pinned host row shard -> device per rank, normalise, "
                  "NCCL all-gather of the bf16 rows over NVLink, triangle sweep, each rank merges its own rows over NVLink peer memory, "
                  "compaction, D2H of each rank's candidates"}
    base["e2e"] = e2e

    # ---- the same call for a caller whose extractor already produces half precision (NOT the headline: the reference hands
    #      over fp32 rows, place_recognition.py:297): semgate_find_loop_closures_host_dtype on the same rows rounded to fp16
    if world == 1:
        try:
            q16 = torch.from_numpy(desc_h.astype(np.float16)).pin_memory()
            q16n = q16.numpy()
            for _ in range(max(args.warmup, 10)):
                n16 = len(eng.find_loop_closures_host(q16n, ts_p, fl_p, p, out=outs)[0])
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                n16 = len(eng.find_loop_closures_host(q16n, ts_p, fl_p, p, out=outs)[0])
            s16 = time.perf_counter() - t0
            base["e2e_fp16_host_input"] = {
                "value": pairs_per_step * e2e_steps / s16, "unit": UNIT, "ms_per_step": s16 / e2e_steps * 1e3,
                "h2d_bytes_per_step": int(q16n.nbytes + ts_h.nbytes + fl_h.nbytes), "d2h_bytes_per_step": int(n16 * 13 + 8),
                "steps": e2e_steps, "candidates": int(n16),
                "api": "semgate_find_loop_closures_host_dtype(SEMGATE_DTYPE_F16): the same host rows rounded to fp16, pinned",
                "note": "not the headline e2e: input precision differs from the reference's fp32 rows (candidates are those of "
                        "the fp16-rounded rows; bit-identical to the fp32 call on their widened image, tests/test_gpu_parity.py)"}
            del q16, q16n
        except Exception as e:      # noqa: BLE001
            base["e2e_fp16_host_input"] = {"failed": f"{type(e).__name__}: {e}"[:400]}

    # ---- BASELINE config 5 in the same run, every N (skipped only on request)
    if not args.no_c5:
        try:
            base["north_star_c5"] = c5_leg(torch, dist, eng, sr, dev, rank, world, peak_tf, peak_src)
        except Exception as e:      # noqa: BLE001
            base["north_star_c5"] = {"failed": f"{type(e).__name__}: {e}"[:400]}
            if world > 1:
                raise

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    if world == 1 and not args.no_extra:
        flops_step = roofline["flops_per_launch"]
        for name, fn in (("companions", lambda: companions_leg(torch, eng, peak_hbm)),
                         ("c1", lambda: c1_leg(torch, eng, peak_tf)),
                         ("sustained", lambda: sustained_leg(torch, eng, lambda: step(), 2.2, flops_step, sus_tf, local))):
            try:
                base[name] = fn()
            except Exception as e:      # noqa: BLE001
                base[name] = {"failed": f"{type(e).__name__}: {e}"[:400]}
            torch.cuda.empty_cache()

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only; bounded sample of the same host arrays)
    cpu = None
    if world == 1 and not args.no_cpu:
        use_all_host_threads()
        rows, _ = calibrate_rows(desc_h, ts_h, fl_h, 10.0)
        dt, cand = cpu_sweep(desc_h, ts_h, fl_h, rows, want_result=True)
        # the same rows on the GPU: candidate counts of the two arms side by side
        gq = out[0][:total_candidates]
        in_rows = int((gq < rows).sum().item())
        cpu = {"value": rows * float(N_DB_PER_GPU) / dt, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
               "sample": f"first {rows} of {N_Q} query keyframes against the full {N_DB_PER_GPU}-keyframe database "
                         f"({dt:.1f} s, oracle/semgate_oracle.py: numpy + OpenBLAS), the same host arrays as the GPU arm",
               "host_cpus": os.cpu_count(), "candidates_in_sample": int(len(cand["query_idx"])),
               "gpu_candidates_same_rows": in_rows, "reference_verbatim": reference_verbatim_block()}
    base["cpu_baseline"] = cpu
    print(json.dumps(base))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cta-group", type=int, default=0, choices=[0, 1, 2, 4])
    ap.add_argument("--exchange", default="auto", choices=["auto", "allgather", "peer"],
                    help="N>1: how the per-GPU candidate lists meet (peer memory over NVLink, or NCCL all-gather)")
    ap.add_argument("--symmetric", default="auto", choices=["auto", "off"],
                    help="auto: all-pairs sweeps (queries == database) compute every similarity once; off: always the full matrix")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (and everything after it)")
    ap.add_argument("--no-c5", action="store_true", help="skip the 1M-keyframe leg of the default run")
    ap.add_argument("--no-extra", action="store_true", help="skip the sustained / companions / c1 legs of the default run")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="c2 (default) is the configuration the metric is quoted on")
    args = ap.parse_args()
    set_workload(args.workload)
    if args.workload == "c2" and args.gpus > 1:
        enable_allpairs(args.gpus)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and args.gpus != world:
        if args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
