"""Average linkage on the GPU against scipy on the box's host cores, on the rounded dissTOM of a synthetic network
(development aid).   python tools/linkagebench.py [n ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200WGCNA_QUIET", "1")
import numpy as np
import pandas as pd
from scipy.cluster.hierarchy import linkage as scipy_linkage

import pywgcna_b200 as bw
from pywgcna_b200.synth import make_expression

for n in [int(a) for a in sys.argv[1:]] or [6000, 23844]:
    X = make_expression(192, n, F=40, seed=5192)
    A = bw.adjacency(pd.DataFrame(X), adjacencyType="signed hybrid", power=10)
    y = bw.dissTOM(A, TOMType="signed", condensed=True)
    del A
    bw.linkage(y[: 45], "average")  # warm
    t0 = time.perf_counter(); Z = bw.linkage(y, "average"); t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter(); Z = bw.linkage(y, "average"); t_gpu2 = time.perf_counter() - t0
    os.environ["B200W_TRACE"] = "1"
    t0 = time.perf_counter(); Zs = scipy_linkage(y, "average"); t_cpu = time.perf_counter() - t0
    print(f"n = {n}: GPU linkage {t_gpu:.3f} s / {t_gpu2:.3f} s (host call incl. {y.nbytes / 1e9:.2f} GB H2D), scipy {t_cpu:.1f} s, "
          f"identical {np.array_equal(Z, Zs)}", flush=True)
