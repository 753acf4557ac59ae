#!/usr/bin/env python
"""bench.py -- the hot path of PyWGCNA network construction on B200 (contract in the task brief).

A "step" is one pass of the hot path over one synthetic expression matrix that is already resident
in HBM:  standardise -> fused soft-threshold sweep (all powers) -> per-power statistics, the 2 x P
ten-point regressions and the power selection -> adjacency at the selected power -> TOM.

Workload at every N: BASELINE.json configs[2] ("C3": 50 000 genes x 500 samples, unsigned network and TOM,
powers 1..20 -- the configuration `north_star` names for 1/2/4/8 GPUs; it fits one GPU).  With N > 1 the
rows of the adjacency / TOM and the sweep columns are sharded in contiguous row blocks (strong scaling:
the total work is fixed).  The same JSON line carries a `secondary` record for configs[1] ("C2": 20 000 x
200, signed hybrid), measured the same way with fewer steps.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
  python bench.py --impl reference ...                      CPU arm: the UNMODIFIED reference
                                                            (baseline/_ref or /root/reference through
                                                            oracle/ref_import.py; the oracle port only if
                                                            neither tree is present)

One JSON line on stdout (rank 0).  metric: algorithmic TFLOP/s of the step,
flops = 2mn^2 (sweep correlation) + 2Pn^2 (power chain) + 2mn^2 (adjacency correlation) + 2n^3 (A.A).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("B200WGCNA_QUIET", "1")
os.environ.setdefault("B200W_REF", os.path.join(ROOT, "baseline", "_ref"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (m, n, F, seed, networkType, TOMType, powers)
    "C2": (200, 20000, 40, 20200, "signed hybrid", "signed", list(range(1, 21))),
    "C3": (500, 50000, 60, 50500, "unsigned", "unsigned", list(range(1, 21))),
    "C4": (5000, 30000, 60, 305000, "signed", "signed", list(range(1, 21))),
    "C5": (300, 80000, 80, 80300, "unsigned", "unsigned", list(range(1, 21))),
    "tiny": (60, 1500, 6, 7, "signed hybrid", "signed", list(range(1, 21))),
}
PRIMARY, SECONDARY = "C3", "C2"
METRIC = "TOMsimilarity+softThreshold TFLOP/s (sweep + adjacency + TOM, algorithmic flops)"
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel (tom_i8_kernel<2>) from `ncu --set full`,
# per launch, keyed by (workload, precision, GPUs) -> (bytes, committed ncu summary it was read from)
TRAFFIC = {
    ("C2", "i8", 1): (75.50e9 + 11.32e9, "profiles/r01_ncu_full_final_step_kernels.txt"),
}
try:
    TRAFFIC.update({tuple(k.split("|")[:2]) + (int(k.split("|")[2]),): (v["bytes"], v["file"])
                    for k, v in json.load(open(os.path.join(ROOT, "profiles", "traffic_table.json"))).items()})
except Exception:
    pass
# what the reference arm samples of a workload whose full size its CPU path cannot finish in minutes
REF_SAMPLE_GENES = {"C3": 20000, "C4": 12000, "C5": 20000}


def step_flops(m, n, P):
    return 2.0 * m * n * n + 2.0 * P * n * n + 2.0 * m * n * n + 2.0 * n ** 3


def config_of(wl, world):
    """The `config` object -- identical in both arms for the same command line."""
    m, n, F, seed, net, tom_type, powers = WORKLOADS[wl]
    return {"workload": wl, "genes": n, "samples": m, "powers": len(powers), "networkType": net,
            "TOMType": tom_type, "TOMDenom": "min",
            "l2": "256 MiB flush each step; operands (3.2 GB and up) exceed the 126 MB L2",
            "parallelism": f"row-block x{world}"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=PRIMARY, choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="auto", choices=["auto", "f64", "i8"])
    ap.add_argument("--cpu-sample-genes", type=int, default=5000, help="our arm's bounded cpu_baseline sample")
    ap.add_argument("--ref-sample-genes", type=int, default=0, help="reference arm: genes per step (0 = table)")
    ap.add_argument("--ref-budget-s", type=float, default=420.0, help="reference arm: stop adding steps after this")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md clocks line)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  NVML is read in-process from a thread
    (tens of microseconds per sample); `nvidia-smi -lms` in a subprocess -- the recipe's clocks line -- was
    measured to stretch a 68 ms step to 150-420 ms on this box, so it is only the fallback."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int, uuid: str | None = None, period: float = 0.02):
        self.idx, self.uuid, self.period = gpu_index, uuid, period
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.how = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + self.uuid) if not self.uuid.startswith("GPU-") else self.uuid)
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.samples.append((time.time(), float(mhz), int(rs)))
                    except Exception:
                        pass
                    self.stop_flag.wait(self.period)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.how = "nvml"
        except Exception:
            self.how = None

    def stop(self, t0, t1):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if self.how is None or not self.samples:
            return self._smi_once()
        rows = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        reasons = set()
        for _, _, rs in rows:
            for bit, nm in self.REASONS.items():
                if rs & bit:
                    reasons.add(nm)
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(rows), "how": "nvml thread, 20 ms period"}

    def _smi_once(self):
        try:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                 capture_output=True, text=True, timeout=20).stdout.strip().split(",")
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = [n for n, v in zip(names, out[2:6]) if v.strip().lower().startswith("active")]
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": reasons, "samples": 1,
                    "how": "nvidia-smi once, right after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own functions on the host cores
# ------------------------------------------------------------------------------------------------
def reference_step_fn():
    """(kind, fn): fn(X, net, tom_type, powers) runs one step of the hot path on the CPU.  kind "reference":
    the UNMODIFIED WGCNA.pickSoftThreshold / adjacency / TOMsimilarity (wgcna.py:788, :1047, :1160) out of
    baseline/_ref (the pip --target install that travels to the GPU box) or /root/reference, imported through
    oracle/ref_import.py; kind "port": oracle/restated.py, only when no reference tree is present."""
    import pandas as pd

    try:
        from oracle import ref_import as R

        if R.reference_available():
            R.load_reference()

            def ref(X, net, tom_type, powers):
                df = pd.DataFrame(X)
                t0 = time.perf_counter()
                power, _ = R.ref_pickSoftThreshold(df, networkType=net, powerVector=list(powers))
                t1 = time.perf_counter()
                A = R.ref_adjacency(df, adjacencyType=net, power=power)
                t2 = time.perf_counter()
                W = R.load_reference()
                with R._quiet(), np.errstate(all="ignore"):
                    T = W.WGCNA.TOMsimilarity(A, TOMType=tom_type)  # in place on A, as findModules calls it
                t3 = time.perf_counter()
                return int(power), float(T.iat[0, 1]), (t1 - t0, t2 - t1, t3 - t2)

            return "reference", ref
    except Exception as e:  # pragma: no cover
        print(f"[bench] reference tree unusable ({e!r}); timing the oracle port", file=sys.stderr)
    from oracle import restated as O

    def port(X, net, tom_type, powers):
        df = pd.DataFrame(X)
        t0 = time.perf_counter()
        power, _ = O.pickSoftThreshold(df, networkType=net, powerVector=powers, faithful=True)
        t1 = time.perf_counter()
        A = O.adjacency(df, adjacencyType=net, power=power)
        t2 = time.perf_counter()
        T = O.TOMsimilarity(A, TOMType=tom_type)
        t3 = time.perf_counter()
        return int(power), float(T.iloc[0, 1]), (t1 - t0, t2 - t1, t3 - t2)

    return "port", port


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else os.cpu_count()
    except Exception:
        return os.cpu_count()


def cpu_time_workload(wl, genes, steps, budget_s):
    """Time `steps` (at least one, fewer if the budget runs out) CPU steps on the first `genes` genes of `wl`."""
    from pywgcna_b200.synth import make_expression

    m, n, F, seed, net, tom_type, powers = WORKLOADS[wl]
    ns = min(n, genes) if genes else n
    X = make_expression(m, ns, F=F, seed=seed)  # the workload's generator (same m, F, seed) at ns genes
    kind, fn = reference_step_fn()
    times, phases, power = [], None, None
    t_start = time.perf_counter()
    while len(times) < max(1, steps):
        t0 = time.perf_counter()
        power, _, ph = fn(X, net, tom_type, powers)
        times.append(time.perf_counter() - t0)
        phases = ph
        if time.perf_counter() - t_start + times[-1] > budget_s:
            break
    dt = float(np.mean(times))
    what = {"reference": "UNMODIFIED reference WGCNA.pickSoftThreshold + adjacency + TOMsimilarity (baseline/_ref via "
                         "oracle/ref_import.py)",
            "port": "oracle/restated.py port of the reference algorithm (no reference tree on this box)"}[kind]
    sample = (f"{wl} generator, {'all' if ns == n else str(ns) + ' of'} {n} genes x {m} samples, "
              f"{len(times)} full step(s) of {what}; numpy/BLAS with all host threads")
    return {"value": step_flops(m, ns, len(powers)) / dt / 1e12, "unit": "TFLOP/s", "cores": cpu_threads(),
            "kind": kind, "seconds": dt, "steps": len(times), "sample_genes": ns, "full_workload": ns == n,
            "selected_power": power, "sample": sample,
            "phases_s": {"pickSoftThreshold": phases[0], "adjacency": phases[1], "TOMsimilarity": phases[2]},
            "host": {"cpu_count": os.cpu_count()}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    genes = args.ref_sample_genes or REF_SAMPLE_GENES.get(wl, 0)
    cb = cpu_time_workload(wl, genes, args.steps, args.ref_budget_s)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": cb["steps"], "warmup": 0, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": cb["seconds"] * 1e3, "seconds": cb["seconds"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(wl, args.gpus), "selected_power": cb["selected_power"],
        "sample_genes": cb["sample_genes"],
        "note": "every step is minutes of CPU: the number of steps is capped by --ref-budget-s and no warm-up step is "
                "spent (BLAS pools warm within the first seconds of a step)",
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "seconds", "phases_s",
                                            "full_workload", "host")},
        "e2e": {"value": cb["value"], "unit": "TFLOP/s", "seconds": cb["seconds"], "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    if wl == PRIMARY and not args.no_secondary and args.gpus == 1:
        # the like-for-like record: configs[1] at FULL size through the same unmodified functions (once: the
        # reference is a one-process CPU program, its numbers do not depend on --gpus)
        sb = cpu_time_workload(SECONDARY, 0, 1, args.ref_budget_s)
        line["secondary"] = {"config": config_of(SECONDARY, args.gpus), "value": sb["value"], "unit": "TFLOP/s",
                             "ms_per_step": sb["seconds"] * 1e3, "seconds": sb["seconds"], "steps": sb["steps"],
                             "selected_power": sb["selected_power"], "cpu_baseline": sb,
                             "e2e": {"value": sb["value"], "unit": "TFLOP/s", "seconds": sb["seconds"],
                                     "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def tom_parity_probe(net_dev, T, tom_type, rows_per_rank=32, seed=1234):
    """Hardware parity of whatever produced `T` (this rank's TOM rows): `rows_per_rank` sampled rows of every
    rank's block against the float64 formula of wgcna.py:1141-1157 evaluated independently with torch
    (checker only): L[i, :] = sum_u a_iu a_u: is accumulated over the ranks' adjacency row blocks and summed
    with an all-reduce.  Returns {max_abs, rows, ranks} (max over ranks)."""
    import torch
    import torch.distributed as dist

    world, rank, n = net_dev.world, net_dev.rank, net_dev.n
    dev = net_dev.dev
    A = net_dev.A[:net_dev.nrows]
    r0, nr = net_dev.row0, net_dev.nrows
    rng = np.random.default_rng(seed + rank)
    S = rows_per_rank
    local = np.sort(rng.choice(nr, size=min(S, nr), replace=False)) if nr > 0 else np.zeros(0, dtype=np.int64)
    idx = torch.full((S,), -1, dtype=torch.int64, device=dev)
    idx[:len(local)] = torch.from_numpy(local + r0).to(dev)
    rows = torch.zeros((S, n), dtype=torch.float64, device=dev)
    if len(local):
        rows[:len(local)] = A[torch.from_numpy(local).to(dev)]
    if world > 1:
        all_idx = torch.empty((world * S,), dtype=torch.int64, device=dev)
        all_rows = torch.empty((world * S, n), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(all_idx, idx)
        dist.all_gather_into_tensor(all_rows, rows)
    else:
        all_idx, all_rows = idx, rows
    valid = all_idx >= 0
    ar = torch.arange(all_idx.numel(), device=dev)
    all_rows = torch.nan_to_num(all_rows, nan=0.0)
    all_rows[ar[valid], all_idx[valid]] = 0.0  # fill_diagonal(A, 0) on the sampled rows (:1143)
    # this rank's block with NaN -> 0; its diagonal entries are removed from the products below instead of
    # cloning the block: L[i, j] = sum_u a0[i, u] A[u, j] - a0[i, j] A[j, j]
    L = torch.zeros((all_idx.numel(), n), dtype=torch.float64, device=dev)
    k = torch.zeros(n, dtype=torch.float64, device=dev)
    if nr > 0:
        Ab = torch.nan_to_num(A, nan=0.0) if bool(torch.isnan(A).any()) else A
        L = all_rows[:, r0:r0 + nr] @ Ab
        dg = torch.diagonal(Ab, offset=r0).clone()  # A[u, u] for this block's rows
        k[r0:r0 + nr] = Ab.sum(dim=1) - dg
        cols = torch.arange(r0, r0 + nr, device=dev)
        L[:, cols] -= all_rows[:, cols] * dg[None, :]
    if world > 1:
        dist.all_reduce(L)
        dist.all_reduce(k)
    a = all_rows
    ki = torch.where(valid, k[all_idx.clamp(min=0)], torch.zeros_like(k[:1]))
    D = torch.minimum(ki[:, None], k[None, :])
    ref = (L + a) / (D + 1 - a) if tom_type == "unsigned" else (L + a).abs() / (D + 1 - a.abs())
    ref[ar[valid], all_idx[valid]] = 1.0
    mine = ref[rank * S:rank * S + len(local)]
    got = T[torch.from_numpy(local).to(dev)] if len(local) else T[:0]
    err = (got - mine).abs().max() if len(local) else torch.zeros((), dtype=torch.float64, device=dev)
    err = err.reshape(1).clone()
    nrows = torch.tensor([float(len(local))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        dist.all_reduce(nrows)
    return {"max_abs": float(err.item()), "rows": int(nrows.item()), "ranks": world,
            "against": "float64 formula of wgcna.py:1141-1157 (torch, all-reduced over the row blocks)"}


def measure_ours(args, wl, steps, warmup, world, rank, local, with_e2e, with_fast, e2e_steps=2):
    """Everything measured for one workload; returns the dict the JSON line (or its `secondary`) is built from."""
    import torch
    import torch.distributed as dist

    import pywgcna_b200 as bw
    from pywgcna_b200 import _lib as L
    from pywgcna_b200 import wgcna as W
    from pywgcna_b200.device import DeviceNetwork
    from pywgcna_b200.shard import symmetric_tile_count
    from pywgcna_b200.synth import make_expression

    lib = L.load()
    m, n, F, seed, net, tom_type, powers = WORKLOADS[wl]
    P = len(powers)
    X_np = make_expression(m, n, F=F, seed=seed)
    X_host = torch.from_numpy(X_np).pin_memory()
    net_dev = DeviceNetwork(X_host, device=local, networkType=net, precision=args.precision, world=world, rank=rank)
    # schedule of the soft-threshold sweep over the ranks (device.py: sym_sweep); None on one GPU
    sharded_sweep = None if world == 1 else ("symmetric" if net_dev.sym_sweep else "full")
    prec_name = {L.PREC_F64: "f64", L.PREC_I8: "i8"}[
        L.PREC_F64 if lib.b200w_tom_operand_bytes(128, net_dev.prec) == 128 * 128 * 8 else L.PREC_I8]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tom_events, phase_events = [], []
    state = {}

    def step():
        if not os.environ.get("B200W_BENCH_NOFLUSH"):
            flush.zero_()  # L2 flush between steps (256 MiB > 126 MB L2); the big operands exceed L2 anyway
        p0, p1, p2, p3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        p0.record()
        net_dev.standardize()
        # sweep -> per-power histograms / medians / exact sums -> the sft table of wgcna.py:924-932 ->
        # power selection :945-953
        power, _table = net_dev.sft(powers, nBreaks=10, RsquaredCut=0.9, MeanCut=100)
        p1.record()
        net_dev.adjacency(power)
        p2.record()
        T = net_dev.tom(TOMType=tom_type, TOMDenom="min")
        p3.record()
        phase_events.append((p0, p1, p2, p3))
        state["power"] = power
        state["T"] = T
        return power

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        step()
    barrier()
    phase_events.clear()
    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid)
    if rank == 0 and not os.environ.get("B200W_BENCH_NOSAMPLER"):
        sampler.start()
    launches0 = lib.b200w_launch_count()
    barrier()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = lib.b200w_launch_count() - launches0
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    phases = {"soft_threshold": float(np.mean([a.elapsed_time(b) for a, b, _, _ in phase_events])),
              "adjacency": float(np.mean([b.elapsed_time(c) for _, b, c, _ in phase_events])),
              "tom": float(np.mean([c.elapsed_time(d) for _, _, c, d in phase_events]))}
    power = int(state["power"].item()) if hasattr(state["power"], "item") else int(state["power"])

    # hardware parity of the path that was just timed (every rank, real NCCL / IPC path at N > 1)
    parity = tom_parity_probe(net_dev, state["T"], tom_type)

    # dominant kernel alone (the A.A contraction + TOM epilogue): CUDA events on the launching stream
    kern_ms = []
    for i in range(2 + min(steps, 5)):
        flush.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        net_dev.tom_compute(TOMType=tom_type, TOMDenom="min")
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            kern_ms.append(e0.elapsed_time(e1))
    barrier()
    kern = float(np.mean(kern_ms))
    # the opt-in faster arithmetic (B200W_PREC_I8_FAST: 10 instead of 15 digit-pair products, max-abs error on
    # TOM ~1e-9 against a north-star tolerance of 1e-6) timed the same way -- beside the headline, not in it
    fast_ms = None
    if with_fast and prec_name == "i8" and world == 1:
        saved, ts = net_dev.prec, []
        net_dev.prec = L.PREC_I8_FAST
        for i in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            net_dev.tom_compute(TOMType=tom_type, TOMDenom="min")
            e1.record()
            torch.cuda.synchronize()
            if i >= 1:
                ts.append(e0.elapsed_time(e1))
        net_dev.prec = saved
        fast_ms = float(np.mean(ts))
    eff_flops = 2.0 * net_dev.nrows * n * n  # this rank's share of the algorithmic 2 n^3
    if prec_name == "i8" and (world == 1 or net_dev.fused):
        # symmetric schedule: only tiles on/above the diagonal are executed (256 x 256 x ldn MACs each)
        tiles = symmetric_tile_count(n, world, rank, net_dev.per)
        kern_flops = 2.0 * tiles * 256 * 256 * net_dev.ldn
    else:
        kern_flops = eff_flops
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = peaks.get("bf16_tflops", 1590.0)
    peak_src = "measured bf16 burst (MEASURED_PEAKS.json)" if "bf16_tflops" in peaks else "fallback 1590"
    if prec_name == "i8":
        mmas = int(os.environ.get("B200W_I8_PAIRS", "15"))
        try:
            i8 = json.load(open(os.path.join(ROOT, "profiles", "r01_int8_peak.json")))
            i8_tops = float(i8["int8_tops_burst"])
            i8_src = ("int8 GEMM peak measured with MEASURED_PEAKS.json's method (cuBLASLt via torch._int_mm 8192^3, "
                      f"burst; sustained {i8['int8_tops_sustained']:.0f}; tcgen05 issue-rate ceiling "
                      f"{i8['tcgen05_issue_rate_tops']:.0f}) -- profiles/r01_int8_peak.json")
            issue = float(i8["tcgen05_issue_rate_tops"]) / mmas
        except Exception:
            i8_tops, i8_src, issue = 2.0 * bf16, f"2 x {peak_src} (int8 not separately measured)", None
        peak = i8_tops / mmas
        peak_note = f"{i8_tops:.0f} TOP/s [{i8_src}] / {mmas} digit-pair s8 MMAs per logical float64-grade product"
    else:
        peak, issue = 37.0, None
        peak_note = "FP64 DMMA nominal 37 TFLOP/s (HGX B200 spec; MEASURED_PEAKS.json has no f64 figure)"
    traffic = TRAFFIC.get((wl, prec_name, world))
    roofline = {"bound": "tensor", "achieved": kern_flops / kern / 1e9, "peak": peak, "unit": "TFLOP/s",
                "frac": kern_flops / kern / 1e9 / peak, "traffic": traffic[0] if traffic else None,
                "traffic_source": traffic[1] if traffic else None,
                "kernel": f"tom_compute[{prec_name}]", "kernel_ms": kern, "share_of_step": kern / ms_step,
                "executed_flops": kern_flops, "algorithmic_flops": eff_flops,
                "effective": eff_flops / kern / 1e9,
                "frac_of_issue_ceiling": (kern_flops / kern / 1e9 / issue) if issue else None,
                "note": "achieved = EXECUTED flops / time (symmetric schedule: tiles on and above the diagonal "
                        "only, mirrored by the epilogue); effective = algorithmic 2 n^3 share / time",
                "peak_note": peak_note}
    if fast_ms is not None:
        roofline["i8_fast_kernel_ms"] = fast_ms  # not part of `value`

    # end to end through the public API with HOST buffers (pinned expression matrix in, numpy / pandas out)
    e2e = None
    e2e_fused = None
    del state["T"]
    if with_e2e:
        import pandas as pd

        del net_dev
        torch.cuda.empty_cache()
        df = pd.DataFrame(X_np)
        E = lib.b200w_sft_exp_buckets()

        def three_calls():
            # the reference-shaped drop-in: three calls, n x n through the host twice (wgcna.py:278, :310, :320)
            p, _ = bw.pickSoftThreshold(df, networkType=net, powerVector=powers, device=local)
            A = bw.adjacency(df, adjacencyType=net, power=p, device=local)
            return bw.TOMsimilarity(A, TOMType=tom_type, device=local, precision=args.precision)

        def one_call():
            # the sharded host entry: expression matrix up, this rank's TOM rows down
            return bw.network_blocks(df, powerVector=powers, networkType=net, TOMType=tom_type, device=local,
                                     precision=args.precision)["TOM"]

        def timed(fn):
            out = fn()
            del out
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                out = fn()
                del out
            dt = (time.perf_counter() - t0) / e2e_steps
            tt_ = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
            return float(tt_.item())

        small = P * (4 + 20) * 8 + P * E * 16
        dt1 = timed(one_call)
        e2e_fused = {"value": step_flops(m, n, P) / dt1 / 1e12, "unit": "TFLOP/s", "seconds": dt1, "steps": e2e_steps,
                     "h2d_bytes_per_step": world * m * n * 8, "d2h_bytes_per_step": n * n * 8 + world * small,
                     "api": "pywgcna_b200.network_blocks: expression matrix in, TOM row block out (one call per rank)"}
        if world == 1:
            dt3 = timed(three_calls)
            e2e = {"value": step_flops(m, n, P) / dt3 / 1e12, "unit": "TFLOP/s", "seconds": dt3, "steps": e2e_steps,
                   "h2d_bytes_per_step": 2 * m * n * 8 + n * n * 8, "d2h_bytes_per_step": small + 2 * n * n * 8,
                   "api": "pickSoftThreshold + adjacency + TOMsimilarity (pandas/numpy in, numpy/pandas out): the "
                          "reference's three calls"}
        else:
            e2e = e2e_fused
        from pywgcna_b200.device import release_workspace

        release_workspace()
        lib.b200w_dev_trim()
        lib.b200w_host_trim()
    else:
        del net_dev
    torch.cuda.empty_cache()
    return {"ms_per_step": ms_step, "value": step_flops(m, n, P) / ms_step / 1e9, "phases_ms": phases,
            "roofline": roofline, "e2e": e2e, "e2e_fused": e2e_fused, "clocks": clocks, "gpu_launches": int(launches),
            "selected_power": power, "tom_precision": prec_name, "parity": parity, "steps": steps,
            "warmup": max(warmup, 3), "sharded_sweep": sharded_sweep,
            "dtype": "f64" if prec_name == "f64" else "f64 (DMMA) + s8->s32 digit slices (tcgen05)"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from pywgcna_b200 import wgcna as W

    W.VERBOSE = False
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N > 1 with torchrun (python -m torch.distributed.run --nproc-per-node N ...)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = args.workload
    r = measure_ours(args, wl, args.steps, args.warmup, world, rank, local, with_e2e=not args.no_e2e, with_fast=True)
    sec = None
    if wl == PRIMARY and not args.no_secondary:
        s = measure_ours(args, SECONDARY, min(args.steps, 10), 3, world, rank, local, with_e2e=not args.no_e2e,
                         with_fast=False, e2e_steps=3)
        sec = {"config": config_of(SECONDARY, world), "value": s["value"], "unit": "TFLOP/s",
               "ms_per_step": s["ms_per_step"], "steps": s["steps"], "warmup": s["warmup"],
               "selected_power": s["selected_power"], "phases_ms": s["phases_ms"], "roofline": s["roofline"],
               "e2e": s["e2e"], "e2e_fused": s["e2e_fused"], "parity": s["parity"], "clocks": s["clocks"]}
    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        cpu = cpu_time_workload(wl, args.cpu_sample_genes, 1, 60.0)
    if rank == 0:
        line = {
            "metric": METRIC, "value": r["value"], "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": r["warmup"], "ms_per_step": r["ms_per_step"],
            "seconds": r["ms_per_step"] / 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": r["dtype"], "data": "synthetic", "config": config_of(wl, world),
            "selected_power": r["selected_power"], "tom_precision": r["tom_precision"],
            "clocks": r["clocks"], "gpu_launches": r["gpu_launches"], "phases_ms": r["phases_ms"],
            "sharded_sweep": r["sharded_sweep"],
            "parity": r["parity"], "roofline": r["roofline"], "cpu_baseline": cpu, "e2e": r["e2e"],
            "e2e_fused": r["e2e_fused"], "secondary": sec,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
