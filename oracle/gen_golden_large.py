"""TEST INFRASTRUCTURE ONLY -- full-size fixtures from the UNMODIFIED reference (run in the build container,
where /root/reference exists; minutes of CPU and up to ~45 GB of RAM):

    python -m oracle.gen_golden_large c2         # tests/golden/c2_fullsize.npz
    python -m oracle.gen_golden_large c1 [net]   # tests/golden/c1scale_<net>.npz

c2   BASELINE.json configs[1] (20 000 genes x 200 samples, signed hybrid, powers 1..20): the reference's own
     ``pickSoftThreshold`` table and selected power (wgcna.py:788-955), ``adjacency`` at that power and
     ``TOMsimilarity`` -- stored as the 20 x 7 table plus a few sampled rows of the two n x n matrices.
c1   the 192 x 23 844 stand-in for the 5xFAD tutorial matrix (the expression CSV itself is a
     missing blob) through the reference's own ``WGCNA.findModules`` (wgcna.py:265-404,
     ``oracle/findmodules_harness.py``): soft-threshold table, power, the gene tree (``linkage``), the
     ``cutreeHybrid`` labels, the merged module labels and sampled rows of the TOM.
Every fixture also records the wall time each reference stage took here (``bench.py``'s reference arm is
calibrated against them).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import pandas as pd

from oracle import findmodules_harness as H
from oracle import ref_import as R
from pywgcna_b200.synth import CONFIGS, make_expression

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SAMPLE_ROWS_C2 = np.array([0, 1, 777, 4095, 4096, 10000, 15871, 19999])
SAMPLE_ROWS_C1 = np.array([0, 5000, 11922, 23843])


def gen_c2():
    m, n, F, seed, net, tom = CONFIGS["C2"]
    X = make_expression(m, n, F=F, seed=seed)
    df = pd.DataFrame(X)
    powers = list(range(1, 21))
    t0 = time.perf_counter()
    p, sft = R.ref_pickSoftThreshold(df, networkType=net, powerVector=powers)
    t1 = time.perf_counter()
    A = R.ref_adjacency(df, adjacencyType=net, power=p)
    t2 = time.perf_counter()
    adj_rows = A[SAMPLE_ROWS_C2].copy()
    W = R.load_reference()
    with R._quiet(), np.errstate(all="ignore"):
        T = W.WGCNA.TOMsimilarity(A, TOMType=tom)  # in place on A, as findModules calls it
    t3 = time.perf_counter()
    tom_rows = T.to_numpy()[SAMPLE_ROWS_C2].copy()
    times = np.array([t1 - t0, t2 - t1, t3 - t2])
    np.savez_compressed(os.path.join(OUT, "c2_fullsize.npz"), powers=np.asarray(powers), power=np.int64(p),
                        sft=sft.to_numpy(dtype=np.float64), rows=SAMPLE_ROWS_C2, adj_rows=adj_rows,
                        tom_rows=tom_rows, ref_seconds=times, ref_cores=np.int64(os.cpu_count()))
    print("c2: power", p, "seconds (pick, adjacency, TOM)", times, flush=True)


def gen_c1(net):
    m, n, F, seed, _, _ = CONFIGS["C1"]
    tom = "signed" if net == "signed hybrid" else "unsigned"
    X = make_expression(m, n, F=F, seed=seed)
    obj = H.make_wgcna(X, networkType=net, TOMType=tom)
    r = H.run_findModules(obj)
    Z = r["linkage"]
    T = obj.TOM.to_numpy()
    tm = r["timings"]
    key = net.replace(" ", "_")
    np.savez_compressed(
        os.path.join(OUT, f"c1scale_{key}.npz"), power=r["power"], sft=r["sft"], powers=np.asarray(obj.powers),
        link_a=Z[:, 0].astype(np.int32), link_b=Z[:, 1].astype(np.int32), link_h=Z[:, 2],
        link_n=Z[:, 3].astype(np.int32), dynamicMods=r["dynamicMods"].astype(np.int16),
        moduleLabels=r["moduleLabels"].astype(np.int16), rows=SAMPLE_ROWS_C1, tom_rows=T[SAMPLE_ROWS_C1].copy(),
        ref_seconds=np.array([tm["pickSoftThreshold"], tm["adjacency"], tm["TOMsimilarity"], tm["cutreeHybrid"],
                              tm["findModules"]]), ref_cores=np.int64(os.cpu_count()))
    print(f"c1 {net}: power", r["power"], "modules", np.unique(r["dynamicMods"], return_counts=True),
          "merged", len(np.unique(r["moduleLabels"])), "timings", tm, flush=True)


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "c2":
        gen_c2()
    else:
        for net in (sys.argv[2:] or ["signed hybrid", "unsigned"]):
            gen_c1(net)
