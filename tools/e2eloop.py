import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200WGCNA_QUIET", "1")
import numpy as np, pandas as pd
import pywgcna_b200 as bw
from pywgcna_b200.synth import make_expression
X = make_expression(200, 20000, F=40, seed=20200); df = pd.DataFrame(X)
powers = list(range(1, 21))
def step():
    t0 = time.perf_counter(); p, _ = bw.pickSoftThreshold(df, networkType="signed hybrid", powerVector=powers)
    t1 = time.perf_counter(); A = bw.adjacency(df, adjacencyType="signed hybrid", power=p)
    t2 = time.perf_counter(); T = bw.TOMsimilarity(A, TOMType="signed")
    t3 = time.perf_counter(); return T, (t1 - t0, t2 - t1, t3 - t2)
for i in range(8):
    T, ts = step(); print(i, [round(1e3 * t, 1) for t in ts], round(1e3 * sum(ts), 1), flush=True)
