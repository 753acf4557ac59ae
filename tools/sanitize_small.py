"""Small cases of every kernel for compute-sanitizer memcheck (one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200WGCNA_QUIET"] = "1"
import numpy as np, pandas as pd
import pywgcna_b200 as bw
rng = np.random.default_rng(0)
for n, m in ((37, 5), (300, 20), (513, 9)):
    X = rng.standard_normal((m, n)); X[:, 1] = 2.0
    df = pd.DataFrame(X)
    p, sft = bw.pickSoftThreshold(df, networkType="signed hybrid", powerVector=[1, 2, 4, 7])
    A = bw.adjacency(df, adjacencyType="signed hybrid", power=3)
    for prec in ("f64", "i8"):
        T = bw.TOMsimilarity(A.copy(), TOMType="signed", precision=prec)
        D = bw.dissTOM(A.copy(), precision=prec, condensed=True)
    bw.checkAdjMat(np.nan_to_num(A))
    bw.intramodularConnectivity(np.nan_to_num(A), rng.integers(0, 3, n))
    k = bw.softConnectivity(pd.DataFrame(rng.standard_normal((20, max(n, 12)))))
    print("ok", n, m, int(p), float(T.iloc[0, 2]), flush=True)
