import numpy as np, pandas as pd, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); os.environ["B200WGCNA_QUIET"]="1"
import pywgcna_b200 as bw
from oracle import restated as O
rng = np.random.default_rng(0)
for n in (3, 5, 17, 127, 129, 255, 257):
    for m in (4, 9):
        X = rng.standard_normal((m, n))
        df = pd.DataFrame(X)
        p, sft = bw.pickSoftThreshold(df, networkType="unsigned", powerVector=[1, 2, 3, 5])
        p0, sft0 = O.pickSoftThreshold(df, networkType="unsigned", powerVector=[1, 2, 3, 5], faithful=False)
        A = bw.adjacency(df, power=2)
        A0 = O.adjacency(df, power=2)
        e = []
        for prec in ("f64", "i8"):
            T = bw.TOMsimilarity(A.copy(), TOMType="unsigned", precision=prec).to_numpy()
            e.append(np.abs(T - O.TOMsimilarity(A0.copy(), TOMType="unsigned").to_numpy()).max())
        with np.errstate(all="ignore"):
            d = np.nanmax(np.abs(sft.to_numpy(float)[:, 4:] - sft0.to_numpy(float)[:, 4:]))
        print(n, m, "power", int(p), int(p0), "adj", np.abs(A - A0).max(), "tom", e, "sft k-cols", d, flush=True)
