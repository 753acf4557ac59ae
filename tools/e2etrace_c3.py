"""Phase trace (B200W_TRACE=1) of the reference-shaped three calls and of network_blocks at C3 size (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200W_TRACE"] = "1"
os.environ.setdefault("B200WGCNA_QUIET", "1")
import numpy as np, pandas as pd
import pywgcna_b200 as bw
from pywgcna_b200.synth import make_expression
wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
m, n, F, seed, net, tom = {"C3": (500, 50000, 60, 50500, "unsigned", "unsigned"), "C2": (200, 20000, 40, 20200, "signed hybrid", "signed")}[wl]
df = pd.DataFrame(make_expression(m, n, F=F, seed=seed))
powers = list(range(1, 21))
for rep in range(3):
    t0 = time.perf_counter(); p, _ = bw.pickSoftThreshold(df, networkType=net, powerVector=powers)
    t1 = time.perf_counter(); A = bw.adjacency(df, adjacencyType=net, power=p)
    t2 = time.perf_counter(); T = bw.TOMsimilarity(A, TOMType=tom)
    t3 = time.perf_counter()
    print(f"rep {rep}: pick {t1 - t0:.3f}  adjacency {t2 - t1:.3f}  TOM {t3 - t2:.3f}  total {t3 - t0:.3f} s", file=sys.stderr, flush=True)
    del A, T
for rep in range(3):
    t0 = time.perf_counter(); r = bw.network_blocks(df, powerVector=powers, networkType=net, TOMType=tom); t1 = time.perf_counter()
    print(f"network_blocks rep {rep}: {t1 - t0:.3f} s", file=sys.stderr, flush=True)
    del r
