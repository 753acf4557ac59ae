"""Device-side comparison of the int8 tensor-core TOM against the FP64 DMMA TOM (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pywgcna_b200.device import DeviceNetwork
from pywgcna_b200.synth import make_expression
n = int(sys.argv[1]); m = int(sys.argv[2]); net = sys.argv[3]; power = float(sys.argv[4]) if len(sys.argv) > 4 else 6
X = make_expression(m, n, F=40, seed=n + m)
d8 = DeviceNetwork(X, networkType=net, precision="i8")
A = d8.adjacency(power)
T8 = d8.tom(TOMType="signed").clone()
del d8.T
d64 = DeviceNetwork(X, networkType=net, precision="f64")
d64.A = A
T64 = d64.tom(TOMType="signed", A_rows=A)
D = (T8 - T64).abs()
mx = D.max().item(); idx = int(D.argmax().item())
print(f"n={n} m={m} net={net} power={power}: max|i8 - f64| = {mx:.3e} at ({idx // n}, {idx % n}); asym i8 = {(T8 - T8.T).abs().max().item():.3e}; asym f64 = {(T64 - T64.T).abs().max().item():.3e}")
bad = (D > 1e-8).nonzero()
if len(bad):
    r = bad[:, 0]; c = bad[:, 1]
    print("bad entries:", len(bad), "rows", r.min().item(), r.max().item(), "cols", c.min().item(), c.max().item())
    print("row tiles (256):", torch.unique(r // 256)[:20].tolist(), "col tiles:", torch.unique(c // 256)[:20].tolist())
