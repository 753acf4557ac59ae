"""A/B of an env knob of the TOM contraction kernel: alternating batches, medians (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pywgcna_b200.device import DeviceNetwork
from pywgcna_b200.synth import make_expression
n = int(sys.argv[1]); knob = sys.argv[2]; vals = sys.argv[3].split(",")
X = make_expression(200, n, F=40, seed=1)
net = DeviceNetwork(X, networkType="signed hybrid")
net.adjacency(6); net.tom_prepare(); net.tom_compute(TOMType="signed"); torch.cuda.synchronize()
res = {v: [] for v in vals}
for rnd in range(6):
    for v in vals:
        os.environ[knob] = v
        for i in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); net.tom_compute(TOMType="signed"); b.record(); torch.cuda.synchronize()
            if i: res[v].append(a.elapsed_time(b))
for v in vals:
    r = np.array(res[v]); print(knob, v, "median", round(float(np.median(r)), 2), "min", round(float(r.min()), 2), "mean", round(float(r.mean()), 2), flush=True)
