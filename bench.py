"""Benchmark of the gated loop-closure retrieval hot path (BASELINE.json metric:
gated similarity pairs/s + queries/s at top-25; % of bf16 tensor peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE config 2, MixVPR shape): the all-pairs loop-closure sweep with floor gate over
n keyframes x 4096-d, 20 000^2 gated pairs PER GPU: n = 20 000 on one GPU (the configuration the
metric is quoted on), n ~ 20 000 * sqrt(N) on N GPUs (28 160 / 39 936 / 56 320 at N = 2 / 4 / 8),
where the GPUs split the triangle of similarity tiles -> per-GPU work is fixed, "weak".
(`--symmetric off`: the full-matrix form; at N>1 then 20 000 queries against 20 000 database rows
per GPU, sharded by rows.)  A step = one pass of the hot path over resident, already normalised bf16
descriptors: fused tcgen05 sweep (K2) + list merge (K3) [+ exchange of the per-GPU lists over NVLink
+ merge at N>1] + candidate compaction (K4).  `value` = query-database pairs gated per second, whole job.
`e2e` = the same metric through the host-buffer C-ABI call (fp32 descriptors in pinned host
memory -> candidates back in host memory: H2D, normalise, sweep, compaction, D2H all timed).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "multi-level-indoor-slam_b200")):
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
STRONG = False       # --workload c5: fixed 1M x 1M sweep, database rows split over the ranks
ALLPAIRS = False     # default workload at N>1: all-pairs sweep, the ranks split the triangle of tiles (weak scaling)
WORKLOAD_NAME = "BASELINE configs[1]: MixVPR-shape 4096-d, 20k-keyframe all-pairs loop-closure sweep with floor gate"


def set_workload(name: str):
    """Default (c2) is the configuration the metric is quoted on; the others are for DESIGN.md numbers."""
    global N_Q, N_DB_PER_GPU, DIM, NUM_FLOORS, STRONG, WORKLOAD_NAME
    if name == "c2":
        return
    if name == "c5":
        N_Q, N_DB_PER_GPU, DIM, NUM_FLOORS, STRONG = 1_000_000, 1_000_000, 4096, 16, True
        WORKLOAD_NAME = "BASELINE configs[4]: 1M-keyframe multi-floor database, full gated top-k sweep, db rows sharded + NCCL merge"
    elif name == "c4":
        N_Q, N_DB_PER_GPU, DIM, NUM_FLOORS, STRONG = 8_192, 250_000, 49_152, 4, True
        WORKLOAD_NAME = "BASELINE configs[3]: AnyLoc-shape 49152-d VLAD, 250k database (rows sharded across the GPUs) x 8192-query batch"
    elif name == "c3":
        N_Q, N_DB_PER_GPU, DIM, NUM_FLOORS = 10_000, 100_000, 8448, 4
        WORKLOAD_NAME = "BASELINE configs[2]: SALAD-shape 8448-d, 100k database x 10k query batch, exclusion window"
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
        "sharding": ("none" if n_gpus == 1 else
                     f"triangle of similarity tiles x{n_gpus} (every rank holds all rows; each similarity is computed once, on one GPU)"
                     if ALLPAIRS else f"db-rows x{n_gpus}"),
        "pairs_per_gpu": float(N_Q) * db_total(n_gpus) / n_gpus,
        "l2": f"inputs larger than L2 ({2 * db_total(n_gpus) // n_gpus * DIM / 1e6:.0f} MB bf16 database per GPU vs 126 MB L2); no explicit flush",
    }


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            j = json.load(open(p))
            return float(j["bf16_tflops"]), float(j.get("hbm_gbs", 0.0)), "measured (MEASURED_PEAKS.json, burst)"
        except Exception:
            pass
    return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


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
            return
        self.running = True
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
        while self.running:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                watts = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(mhz), float(watts), float(util)))
                for nm, bit in names:
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.running = False
        self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        a = np.array(self.samples)
        return {"sm_mhz": float(np.median(a[:, 0])), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": int(a.shape[0]), "power_w": float(np.median(a[:, 1]))}


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


def cpu_inputs(n_db: int, seed: int = 0):
    """Host copy of the workload: fp32 descriptors, timestamps, floors (same model as the GPU arm)."""
    from semgate import synthetic
    desc = synthetic.make_descriptors(n_db, DIM, seed=seed)
    return desc, synthetic.make_timestamps(n_db), synthetic.make_floors(n_db, NUM_FLOORS).astype(np.int32)


def cpu_sweep(desc, ts, fl, rows: int):
    """The reference algorithm (oracle port, numpy/OpenBLAS on all host threads) on the first
    `rows` query keyframes against the whole database.  Returns seconds."""
    from oracle import semgate_oracle as O
    t0 = time.perf_counter()
    dbn = O.l2_normalize(desc)
    res = O.gated_topk(dbn[:rows], dbn, ts[:rows], ts, fl[:rows], fl, k=TOPK, threshold=THRESHOLD,
                       min_time_gap=MIN_TIME_GAP, max_floor_diff=0, normalize=False, block=1024)
    O.compact(res)
    return time.perf_counter() - t0


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    desc, ts, fl = cpu_inputs(N_DB_PER_GPU)
    budget = 150.0 / max(args.steps + args.warmup, 1)
    rows, _ = calibrate_rows(desc, ts, fl, min(8.0, budget))
    for _ in range(args.warmup):
        cpu_sweep(desc, ts, fl, rows)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sweep(desc, ts, fl, rows)
    dt = time.perf_counter() - t0
    value = rows * float(N_DB_PER_GPU) * args.steps / dt
    sample = f"first {rows} of {N_Q} query keyframes against the full {N_DB_PER_GPU}-keyframe database per step"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus if ALLPAIRS else 1),
        "queries_per_s": rows * args.steps / dt,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference algorithm (oracle/semgate_oracle.py, numpy + OpenBLAS, all host threads); the reference's own "
                "Python loop (place_recognition.py:882-885) is ~25x slower than this port and cannot hold 20k x 20k",
    }
    print(json.dumps(out))


# --------------------------------------------------------------------------- GPU arm
def exchange_info(sr, world):
    if world == 1:
        return {}
    peer = sr._symm is not None
    d = {"exchange": "peer memory: merge kernel reads the per-GPU lists in place over NVLink" if peer
         else "NCCL all-gather of the per-GPU lists, then merge"}
    if sr.peer_error:
        d["peer_exchange_unavailable"] = sr.peer_error
    return d


def run_ours(args):
    import torch
    import torch.distributed as dist
    from semgate import _native, synthetic
    from semgate.dist import ShardedRetrieval, shard_bounds

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

    n_db_total = db_total(world)
    lo, hi = shard_bounds(n_db_total, world, rank)
    dp = _native.pad_dim(DIM)

    # ---- synthetic inputs, resident in HBM (anchors shared by all ranks; rank 0's shard = the queries)
    g = torch.Generator(device=dev); g.manual_seed(1234)
    places = max(8, n_db_total // 20)
    anchors = torch.randn((places, DIM), generator=g, device=dev, dtype=torch.float32)

    def make_rows(n, seed):
        gg = torch.Generator(device=dev); gg.manual_seed(seed)
        pid = torch.randint(0, places, (n,), generator=gg, device=dev)
        x = anchors[pid]
        x += 0.6 * torch.randn((n, DIM), generator=gg, device=dev, dtype=torch.float32)
        return x

    def make_bf16(n, seed):
        """normalised bf16 rows generated chunk-wise (the fp32 form of 1M rows would be 16 GB)"""
        outb = torch.empty((n, dp), dtype=torch.bfloat16, device=dev)
        step = max(1024, min(65536, (1 << 28) // DIM))
        for s0 in range(0, n, step):
            e0 = min(n, s0 + step)
            eng.normalize_cast(make_rows(e0 - s0, seed * 7919 + s0), out=outb[s0:e0])
        return outb

    if STRONG:
        q_f32 = db_f32 = None
        full = make_bf16(n_db_total, 1000)             # same seed on every rank: identical matrix
        q_bf16 = full[:N_Q]                            # the queries are the first N_Q keyframes
        db_bf16 = full[lo:hi]                          # this rank's slice of the database rows
    else:
        # the query keyframes are the first N_Q rows of the database (rank 0's shard starts with them)
        q_f32 = make_rows(N_Q, 1000)
        if rank == 0 and N_Q == hi - lo:
            db_f32 = q_f32
        else:
            db_f32 = make_rows(hi - lo, 2000 + rank)
            if rank == 0:
                db_f32[:min(N_Q, hi - lo)] = q_f32[:min(N_Q, hi - lo)]
        q_bf16 = eng.normalize_cast(q_f32)
        db_bf16 = q_bf16 if db_f32 is q_f32 else eng.normalize_cast(db_f32)
    ts_all = torch.from_numpy(synthetic.make_timestamps(n_db_total)).to(dev)
    fl_all = torch.from_numpy(synthetic.make_floors(n_db_total, NUM_FLOORS).astype(np.int32)).to(dev)
    q_ts, q_fl = ts_all[:N_Q].contiguous(), fl_all[:N_Q].contiguous()
    db_ts, db_fl = ts_all[lo:hi].contiguous(), fl_all[lo:hi].contiguous()

    def mk(offset):
        return _native.make_params(k=TOPK, similarity_threshold=THRESHOLD, min_time_gap=MIN_TIME_GAP, max_floor_diff=0,
                                   gate_mode=_native.GATE_FLAG, db_index_offset=offset)

    # an all-pairs sweep over a database every rank holds in full: the ranks split the triangle of tiles
    triangle = STRONG and world > 1 and N_Q == n_db_total and args.symmetric == "auto"

    def step():
        if triangle:
            return sr.sweep_all_pairs(full, mk, ts=ts_all, floor=fl_all, max_floor_diff=0, compact=True)
        else:
            res = sr.sweep(q_bf16, db_bf16, mk, lo, q_ts=q_ts, db_ts_shard=db_ts, q_floor=q_fl, db_floor_shard=db_fl,
                           db_floor_all=fl_all, max_floor_diff=0)
        return eng.compact(res)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    eng.profile_read()
    launches0 = eng.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    k2_ms, k2_n = eng.profile_read()
    clocks = sampler.stop() if rank == 0 else None
    total_candidates = int(out[4].item())
    sweep_mode, sweep_tiles = eng.last_sweep_mode()   # 0 full, 1 symmetric (every similarity computed once), 2 overflowed
    if world > 1:
        t = torch.tensor([ms, k2_ms / max(k2_n, 1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, k2_avg = float(t[0]), float(t[1])
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    else:
        k2_avg = k2_ms / max(k2_n, 1)
    pairs_per_step = float(N_Q) * float(n_db_total)
    value = pairs_per_step * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (K2): algorithmic FLOPs = 2*Q*N_local*Dpad per launch
    peak_tf, peak_hbm, peak_src = measured_peaks()
    # A symmetric sweep (queries == database) computes only the tiles on or above the block diagonal: the
    # roofline counts the FLOPs the tensor cores EXECUTED, the full-matrix figure is given beside it.
    flops_full = 2.0 * N_Q * (hi - lo) * dp
    flops_per_launch = flops_full
    if sweep_mode == 1:
        nb = (N_Q + 255) // 256            # sweep_tiles: this rank's share of the nb*(nb+1)/2 tiles
        flops_per_launch = 2.0 * N_Q * n_db_total * dp * sweep_tiles / float(nb * nb)
    achieved_tf = flops_per_launch / (k2_avg * 1e-3) / 1e12 if k2_avg > 0 else 0.0
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": None, "kernel": "gated_topk_kernel (K2)",
                "kernel_ms": k2_avg, "kernel_share_of_step": k2_avg / (ms / args.steps), "peak_source": peak_src,
                "flops_per_launch": flops_per_launch, "flops_counted": "executed by the tensor cores",
                "sweep": {0: "full matrix", 1: "symmetric: S_ij = S_ji computed once, gated in both directions",
                          2: "symmetric attempt overflowed, full sweep redone"}[sweep_mode],
                "full_matrix_flops": flops_full,
                "full_matrix_equivalent_tflops": flops_full / (k2_avg * 1e-3) / 1e12 if k2_avg > 0 else 0.0}
    prof = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if sweep_mode == 1:
        prof = os.path.join(ROOT, "profiles", "k2_sym_traffic.json")
    if os.path.isfile(prof) and args.workload == "c2" and world == 1:   # the ncu capture is of this workload
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass

    if (STRONG and not ALLPAIRS) or args.no_e2e or args.workload != "c2":
        if rank == 0:
            out_json = {
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if (STRONG and not ALLPAIRS) else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(world), "queries_per_s": N_Q * args.steps / (ms * 1e-3),
                "candidates_per_step": total_candidates, "roofline": roofline, "cpu_baseline": None, "e2e": None,
                "gpu_launches": int(launches), "clocks": clocks,
                "note": "non-default workload: e2e / cpu_baseline legs are only run for the default configuration",
            }
            out_json.update(exchange_info(sr, world))
            print(json.dumps(out_json))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the public host-buffer API (pinned host memory in, host memory out)
    e2e_steps = max(3, min(args.steps, 10))
    if ALLPAIRS:
        q_host = torch.empty((hi - lo, DIM), dtype=torch.float32, pin_memory=True)   # this rank's rows of the database
        q_host.copy_(make_rows(hi - lo, 5000 + rank))
    else:
        q_host = torch.empty((N_Q, DIM), dtype=torch.float32, pin_memory=True)
        q_host.copy_(q_f32)
    ts_host = synthetic.make_timestamps(n_db_total)
    fl_host = synthetic.make_floors(n_db_total, NUM_FLOORS).astype(np.int32)
    cap = N_Q * TOPK
    if world == 1:
        outs = tuple(torch.empty((cap,), dtype=dt, pin_memory=True).numpy()
                     for dt in (torch.int32, torch.int32, torch.float32, torch.uint8))
        qh = q_host.numpy()
        p = mk(0)

        def e2e_step():
            r = eng.find_loop_closures_host(qh, ts_host, fl_host, p, out=outs)   # semgate_find_loop_closures_host
            return len(r[0])
        h2d = qh.nbytes + ts_host.nbytes + fl_host.nbytes
    elif ALLPAIRS:
        tsh = torch.from_numpy(ts_host).pin_memory()
        flh = torch.from_numpy(fl_host).pin_memory()
        ho = [torch.empty((cap,), dtype=dt, pin_memory=True) for dt in (torch.int32, torch.int32, torch.float32, torch.uint8)]

        def e2e_step():
            # every rank uploads and normalises its rows; the bf16 rows meet over NVLink (all-gather); triangle sweep
            oq, om, os_, ov, tot = sr.sweep_all_pairs_from_host(q_host, tsh, flh, mk, lo, hi, N_Q, max_floor_diff=0, compact=True)
            t = int(tot.item())
            if rank == 0:
                for h, d in zip(ho, (oq, om, os_, ov)):
                    h[:t].copy_(d[:t], non_blocking=True)
                torch.cuda.synchronize()
            return t
        h2d = N_Q * DIM * 4 + (ts_host.nbytes + fl_host.nbytes) * world   # one fp32 row shard per rank
    else:
        db_host = q_host if rank == 0 else torch.empty((hi - lo, DIM), dtype=torch.float32, pin_memory=True)
        if rank != 0:
            db_host.copy_(db_f32)
        tsh = torch.from_numpy(ts_host).pin_memory()
        flh = torch.from_numpy(fl_host).pin_memory()
        ho = [torch.empty((cap,), dtype=dt, pin_memory=True) for dt in (torch.int32, torch.int32, torch.float32, torch.uint8)]

        def e2e_step():
            # queries cross PCIe once (rank 0, they are its shard) and travel on as bf16 over NVLink
            res = sr.sweep_from_host(q_host if rank == 0 else None, db_host, tsh, flh, mk, lo, hi, N_Q, max_floor_diff=0)
            oq, om, os_, ov, tot = eng.compact(res)
            t = int(tot.item())
            if rank == 0:
                for h, d in zip(ho, (oq, om, os_, ov)):
                    h[:t].copy_(d[:t], non_blocking=True)
                torch.cuda.synchronize()
            return t
        h2d = q_host.numel() * 4 * world + (ts_host.nbytes + fl_host.nbytes) * world   # one fp32 shard per rank
    del q_f32, db_f32
    n_e2e = e2e_step()
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
    e2e = {"value": pairs_per_step * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(n_e2e * 13 + 8), "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
           "api": "semgate_find_loop_closures_host (C ABI, pinned host buffers)" if world == 1 else
                  "semgate python API (ShardedRetrieval.sweep_all_pairs_from_host): pinned host row shard -> device per rank, normalise, NCCL all-gather of the bf16 rows over NVLink, triangle sweep, list merge over NVLink, compaction, D2H" if ALLPAIRS else
                  "semgate python API (ShardedRetrieval.sweep_from_host): pinned host shard -> device per rank, normalise, NVLink broadcast of the bf16 queries, sharded sweep, NCCL merge, compaction, D2H"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only; bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu:
        use_all_host_threads()
        desc = q_host.numpy()
        rows, _ = calibrate_rows(desc, ts_host, fl_host, 10.0)
        dt = cpu_sweep(desc, ts_host, fl_host, rows)
        cpu = {"value": rows * float(N_DB_PER_GPU) / dt, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
               "sample": f"first {rows} of {N_Q} query keyframes against the full {N_DB_PER_GPU}-keyframe database "
                         f"({dt:.1f} s, oracle/semgate_oracle.py: numpy + OpenBLAS)",
               "host_cpus": os.cpu_count()}

    out_json = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
        "all_pairs_split": sr.last_all_pairs,
        "queries_per_s": N_Q * args.steps / (ms * 1e-3), "candidates_per_step": total_candidates,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "cta_group": args.cta_group or os.environ.get("SEMGATE_CTA_GROUP", "auto (2 for Q >= 4096)"),
    }
    out_json.update(exchange_info(sr, world))
    print(json.dumps(out_json))
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
                    help="N>1: how the per-GPU candidate lists meet (NCCL all-gather, or read in place over NVLink by the merge kernel)")
    ap.add_argument("--symmetric", default="auto", choices=["auto", "off"],
                    help="auto: all-pairs sweeps (queries == database) compute every similarity once; off: always the full matrix")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="c2 (default) is the configuration the metric is quoted on")
    args = ap.parse_args()
    set_workload(args.workload)
    if args.workload == "c2" and args.gpus > 1 and args.symmetric == "auto":
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
