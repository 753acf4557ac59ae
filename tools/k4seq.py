import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pywgcna_b200.device import DeviceNetwork
from pywgcna_b200.synth import make_expression
n = int(sys.argv[1]); mode = sys.argv[2]
X = make_expression(200, n, F=40, seed=1)
net = DeviceNetwork(X, networkType="signed hybrid")
net.adjacency(6); net.tom_prepare(); net.tom_compute(TOMType="signed"); torch.cuda.synchronize()
evs = []
t0 = time.perf_counter()
for i in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); net.tom_compute(TOMType="signed"); b.record(); evs.append((a, b))
    if mode == "sync": torch.cuda.synchronize()
    if mode == "sleep": torch.cuda.synchronize(); time.sleep(0.05)
host = time.perf_counter() - t0
torch.cuda.synchronize()
print(mode, "host loop", round(host * 1e3, 1), "ms; per-launch:", [round(a.elapsed_time(b), 1) for a, b in evs], flush=True)
