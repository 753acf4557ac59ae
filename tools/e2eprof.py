import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200WGCNA_QUIET", "1")
import numpy as np, pandas as pd
import pywgcna_b200 as bw
from pywgcna_b200.synth import make_expression
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
X = make_expression(200, n, F=40, seed=20200); df = pd.DataFrame(X)
powers = list(range(1, 21))
def T(name, fn):
    t = time.perf_counter(); r = fn(); print(f"{name:20s} {1e3*(time.perf_counter()-t):9.1f} ms", flush=True); return r
for rep in range(4):
    print("--- rep", rep)
    p, sft = T("pickSoftThreshold", lambda: bw.pickSoftThreshold(df, networkType="signed hybrid", powerVector=powers))
    A = T("adjacency", lambda: bw.adjacency(df, adjacencyType="signed hybrid", power=p))
    Tm = T("TOMsimilarity", lambda: bw.TOMsimilarity(A, TOMType="signed"))
    D = T("dissTOM fused", lambda: bw.dissTOM(A, TOMType="signed"))
    del D
