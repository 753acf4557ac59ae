"""Sweep / adjacency kernels alone on both correlation engines (development aid; also the ncu target).
   python tools/sweepbench.py [--n 20000 --m 200 --P 20 --reps 3 --engines i8,f64 --nosim]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pywgcna_b200 as bw
from pywgcna_b200 import _lib as L
from pywgcna_b200.synth import make_expression

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--m", type=int, default=200)
ap.add_argument("--P", type=int, default=20)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--engines", default="i8,f64")
ap.add_argument("--nosim", action="store_true")
a = ap.parse_args()
lib = L.load()
torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
n, m, P = a.n, a.m, a.P
X = torch.from_numpy(make_expression(m, n, F=40, seed=1)).cuda()
ldk = lib.b200w_padded_k(m)
Z = torch.empty((n, ldk), dtype=torch.float64, device="cuda")
bad = torch.empty(n, dtype=torch.uint8, device="cuda")
K = torch.empty((P, n), dtype=torch.float64, device="cuda")
A = torch.empty((n, n), dtype=torch.float64, device="cuda")
steps = np.ones(P)
L.check(lib.b200w_dev_standardize(X.data_ptr(), m, n, Z.data_ptr(), bad.data_ptr(), 0, st))


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


res = {}
for eng in a.engines.split(","):
    bw.set_correlation_precision(eng)
    t_sweep = timeit(lambda: L.check(lib.b200w_dev_sweep(Z.data_ptr(), bad.data_ptr(), m, n, steps.ctypes.data, P, 2, 0,
                                                         n, K.data_ptr(), 0, st)))
    k = K.clone()
    t_sim = None if a.nosim else timeit(lambda: L.check(lib.b200w_dev_sweep_similarity(
        Z.data_ptr(), bad.data_ptr(), m, n, steps.ctypes.data, P, 2, 0, n, K.data_ptr(), A.data_ptr(), 0, st)))
    t_adj = timeit(lambda: L.check(lib.b200w_dev_adjacency(Z.data_ptr(), bad.data_ptr(), m, n, 2, 6.0, 0, n,
                                                           A.data_ptr(), 0, st)))
    res[eng] = (k, A[:64].clone())
    print(f"{eng}: sweep {t_sweep:.3f} ms  sweep+similarity {t_sim if t_sim is None else round(t_sim, 3)} ms  "
          f"adjacency {t_adj:.3f} ms   (n={n} m={m} P={P})", flush=True)
if len(res) == 2:
    (k0, a0), (k1, a1) = res.values()
    print("max |dk| rel", float(((k0 - k1).abs() / k1.abs().clamp(min=1e-300)).max()), "max |dA|", float((a0 - a1).abs().max()))
