#!/usr/bin/env python
"""bench.py -- the hot path of PyWGCNA network construction on B200 (contract in the task brief).

A "step" is one pass of the hot path over one synthetic expression matrix that is already resident
in HBM:  standardise -> fused soft-threshold sweep (all powers) -> k to the host, scale-free fit and
power selection (host, 10 points per power) -> adjacency at the selected power -> TOM.
Workload at every N: BASELINE.json configs[1] ("C2": 20 000 genes x 200 samples, signed hybrid,
powers 1..20, signed TOM, min denominator); with N > 1 the rows of the adjacency / TOM and the sweep
columns are sharded in contiguous row blocks (strong scaling), operand and k exchanged with NCCL.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
  python bench.py --impl reference ...                      CPU arm: the oracle port of the reference
                                                            (oracle/restated.py) on a bounded sample

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

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (m, n, F, seed, networkType, TOMType, powers)
    "C2": (200, 20000, 40, 20200, "signed hybrid", "signed", list(range(1, 21))),
    "C3": (500, 50000, 60, 50500, "unsigned", "unsigned", list(range(1, 21))),
    "C4": (5000, 30000, 60, 305000, "signed", "signed", list(range(1, 21))),
    "C5": (300, 80000, 80, 80300, "unsigned", "unsigned", list(range(1, 21))),
    "tiny": (60, 1500, 6, 7, "signed hybrid", "signed", list(range(1, 21))),
}
METRIC = "TOMsimilarity+softThreshold TFLOP/s (sweep + adjacency + TOM, algorithmic flops)"
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from `ncu --set full` (profiles/), per launch
TRAFFIC = {("C2", "i8", 1): 75.50e9 + 11.32e9}  # profiles/r01_ncu_full_final_step_kernels.txt (inside a bench step)


def step_flops(m, n, P):
    return 2.0 * m * n * n + 2.0 * P * n * n + 2.0 * m * n * n + 2.0 * n ** 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="auto", choices=["auto", "f64", "i8"])
    ap.add_argument("--cpu-sample-genes", type=int, default=3000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
# CPU arm: oracle port of the reference on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_step(X, net, tom_type, powers):
    """One pass of the reference's CPU algorithm (oracle/restated.py, which follows wgcna.py line by
    line incl. the blocked np.corrcoef of :882) -- sweep, adjacency at the selected power, TOM."""
    import pandas as pd

    from oracle import restated as O

    df = pd.DataFrame(X)
    power, _ = O.pickSoftThreshold(df, networkType=net, powerVector=powers, faithful=True)
    A = O.adjacency(df, adjacencyType=net, power=power)
    T = O.TOMsimilarity(A, TOMType=tom_type)
    return float(T.iloc[0, 1])


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else os.cpu_count()
    except Exception:
        return os.cpu_count()


def cpu_baseline(wl, genes, reps=1):
    from pywgcna_b200.synth import make_expression

    m, n, F, seed, net, tom_type, powers = WORKLOADS[wl]
    ns = min(n, genes)
    X = make_expression(m, ns, F=F, seed=seed)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_step(X, net, tom_type, powers)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": step_flops(m, ns, len(powers)) / best / 1e12, "unit": "TFLOP/s", "cores": cpu_threads(),
            "kind": "port", "seconds": best,
            "sample": f"{wl} generator, first {ns} of {n} genes x {m} samples, one full step "
                      f"(pickSoftThreshold {len(powers)} powers + adjacency + TOMsimilarity), oracle/restated.py on numpy/BLAS"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pywgcna_b200.synth import make_expression

    m, n, F, seed, net, tom_type, powers = WORKLOADS[args.workload]
    ns = min(n, args.cpu_sample_genes)
    X = make_expression(m, ns, F=F, seed=seed)
    for _ in range(min(args.warmup, 1)):  # BLAS thread pools warm after one pass; each costs seconds
        cpu_step(X, net, tom_type, powers)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(X, net, tom_type, powers)
    dt = (time.perf_counter() - t0) / args.steps
    val = step_flops(m, ns, len(powers)) / dt / 1e12
    sample = (f"{args.workload} generator, first {ns} of {n} genes x {m} samples per step, reference algorithm "
              f"(oracle/restated.py port; the Python reference tree does not travel to the GPU box)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "genes": n, "samples": m, "powers": len(powers),
                   "networkType": net, "TOMType": tom_type, "sample_genes": ns},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cpu_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import pywgcna_b200 as bw
    from pywgcna_b200 import _lib as L
    from pywgcna_b200 import wgcna as W
    from pywgcna_b200.device import DeviceNetwork
    from pywgcna_b200.synth import make_expression

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
    lib = L.load()
    m, n, F, seed, net, tom_type, powers = WORKLOADS[args.workload]
    P = len(powers)
    X_host = torch.from_numpy(make_expression(m, n, F=F, seed=seed)).pin_memory()
    net_dev = DeviceNetwork(X_host, device=local, networkType=net, precision=args.precision,
                            world=world, rank=rank)
    prec_name = {L.PREC_F64: "f64", L.PREC_I8: "i8"}[
        L.PREC_F64 if lib.b200w_tom_operand_bytes(128, net_dev.prec) == 128 * 128 * 8 else L.PREC_I8]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tom_events, phase_events = [], []
    state = {}

    def step():
        if not os.environ.get("B200W_BENCH_NOFLUSH"):
            flush.zero_()  # L2 flush between steps (256 MiB > 126 MB L2); the big operands exceed L2 anyway
        p0, p1, p2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        p0.record()
        net_dev.standardize()
        # sweep -> per-power histograms / medians / exact sums on the device -> the whole sft table of
        # wgcna.py:924-932 (two 10-point regressions per power on the host) -> power selection :945-953
        power, _table = net_dev.sft(powers, nBreaks=10, RsquaredCut=0.9, MeanCut=100)
        power = int(power)
        p1.record()
        net_dev.adjacency(power)
        p2.record()
        phase_events.append((p0, p1, p2))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # events around the contraction only: prepare + exchange happen inside tom() before compute,
        # so bracket the whole call and subtract nothing -- the contraction dominates; the kernel-only
        # time comes from the second pair recorded by the library hook below
        e0.record()
        T = net_dev.tom(TOMType=tom_type, TOMDenom="min")
        e1.record()
        tom_events.append((e0, e1))
        state["power"] = power
        state["probe"] = T
        return power

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    tom_events.clear()
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
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = lib.b200w_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    tom_ms = float(np.mean([a.elapsed_time(b) for a, b in tom_events]))
    # where a step's time goes (device timeline; the soft-threshold phase includes the host's 2 x P regressions,
    # during which the GPU waits for the selected power)
    phases = {"soft_threshold": float(np.mean([a.elapsed_time(b) for a, b, _ in phase_events])),
              "adjacency": float(np.mean([b.elapsed_time(c) for _, b, c in phase_events])),
              "tom": tom_ms}

    # dominant kernel alone (the A.A contraction + TOM epilogue): CUDA events on the launching stream
    from pywgcna_b200.shard import symmetric_tile_count

    kern_ms = []
    for i in range(2 + min(args.steps, 5)):
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
    # TOM ~1e-9 against a north-star tolerance of 1e-6) timed the same way -- reported beside the headline, which
    # stays on the default precision
    fast_ms = None
    if prec_name == "i8" and world == 1:
        from pywgcna_b200 import _lib as L

        saved, ts = net_dev.prec, []
        net_dev.prec = L.PREC_I8_FAST
        for i in range(4):
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
        except Exception:
            i8_tops, i8_src = 2.0 * bf16, f"2 x {peak_src} (int8 not separately measured)"
        peak = i8_tops / mmas
        peak_note = f"{i8_tops:.0f} TOP/s [{i8_src}] / {mmas} digit-pair s8 MMAs per logical float64-grade product"
    else:
        peak = 37.0
        peak_note = "FP64 DMMA nominal 37 TFLOP/s (HGX B200 spec; MEASURED_PEAKS.json has no f64 figure)"
    roofline = {"bound": "tensor", "achieved": kern_flops / kern / 1e9, "peak": peak, "unit": "TFLOP/s",
                "frac": kern_flops / kern / 1e9 / peak, "traffic": TRAFFIC.get((args.workload, prec_name, world)),
                "kernel": f"tom_compute[{prec_name}]", "kernel_ms": kern, "share_of_step": kern / ms_step,
                "executed_flops": kern_flops, "algorithmic_flops": eff_flops,
                "effective": eff_flops / kern / 1e9,
                "note": "achieved = EXECUTED flops / time (symmetric schedule: tiles on and above the diagonal "
                        "only, mirrored by the epilogue); effective = algorithmic 2 n^3 share / time",
                "peak_note": peak_note}
    if fast_ms is not None:
        roofline["i8_fast_kernel_ms"] = fast_ms  # not part of `value`

    # end to end through the public drop-in API (host buffers in, host buffers out)
    e2e = None
    if not args.no_e2e:
        import pandas as pd

        df = pd.DataFrame(X_host.numpy())
        ksteps = max(1, min(args.steps, 3))
        r0, nr = net_dev.row0, net_dev.nrows
        E = lib.b200w_sft_exp_buckets()

        def e2e_step():
            # the reference-shaped host API: host buffers in, host buffers out.  N > 1: every rank runs the
            # (replicated) threshold pick and adjacency on its own GPU over its own PCIe link and asks
            # TOMsimilarity for its row block only -- the API shards by output rows, not by input.
            p, _ = bw.pickSoftThreshold(df, networkType=net, powerVector=powers, device=local)
            A = bw.adjacency(df, adjacencyType=net, power=p, device=local)
            if world == 1:
                return bw.TOMsimilarity(A, TOMType=tom_type, device=local, precision=args.precision)
            return W._tom(A, L.TOM_TYPES.index(tom_type), 0, check=True, device=local, rows=(r0, nr),
                          precision=args.precision)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            T = e2e_step()
        dt = (time.perf_counter() - t0) / ksteps
        tt_ = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
        dt = float(tt_.item())
        e2e = {"value": step_flops(m, n, P) / dt / 1e12, "unit": "TFLOP/s", "seconds": dt, "steps": ksteps,
               "h2d_bytes_per_step": world * (2 * m * n * 8 + n * n * 8),
               "d2h_bytes_per_step": world * (P * (4 + 20) * 8 + P * E * 16 + n * n * 8) + n * n * 8,
               "api": "pickSoftThreshold + adjacency + TOMsimilarity (pandas/numpy in, numpy/pandas out)"
                      + ("" if world == 1 else "; per rank: replicated pick + adjacency, TOM row block")}
        del T
    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        cpu = cpu_baseline(args.workload, args.cpu_sample_genes)
    if rank == 0:
        line = {
            "metric": METRIC, "value": step_flops(m, n, P) / ms_step / 1e9, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "seconds": ms_step / 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if prec_name == "f64" else "f64 (DMMA) + s8->s32 digit slices (tcgen05)",
            "data": "synthetic",
            "config": {"workload": args.workload, "genes": n, "samples": m, "powers": P, "networkType": net,
                       "TOMType": tom_type, "TOMDenom": "min", "selected_power": state["power"],
                       "tom_precision": prec_name, "l2": "256 MiB flush each step; operands (3.2 GB+) exceed L2",
                       "parallelism": f"row-block x{world}"},
            "clocks": clocks, "gpu_launches": int(launches), "tom_ms_in_step": tom_ms, "phases_ms": phases,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
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
