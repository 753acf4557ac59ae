"""Per-phase wall/device times of one bench step (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200WGCNA_QUIET", "1")
import numpy as np, torch
from pywgcna_b200 import wgcna as W
from pywgcna_b200.device import DeviceNetwork
from pywgcna_b200.synth import make_expression
m, n = 200, int(sys.argv[1]) if len(sys.argv) > 1 else 20000
powers = list(range(1, 21))
X = torch.from_numpy(make_expression(m, n, F=40, seed=20200)).pin_memory()
net = DeviceNetwork(X, device=0, networkType="signed hybrid")
def T(name, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"{name:24s} {1e3*(time.perf_counter()-t):9.3f} ms", flush=True); return r
for rep in range(3):
    print("--- rep", rep)
    T("standardize", net.standardize)
    k = T("sweep", lambda: net.sweep(powers))
    kh = T("k.cpu", lambda: k.cpu().numpy())
    def fit():
        r2 = [W._scale_free_fit(kh[i], 10)[0] for i in range(len(powers))]
        return r2
    T("fit", fit)
    T("exact_mean", lambda: [W._exact_mean(kh[i]) for i in range(len(powers))])
    T("adjacency", lambda: net.adjacency(6))
    T("tom", lambda: net.tom(TOMType="signed"))
