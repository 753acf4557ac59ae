"""Quick device-resident timing of the individual kernels (development aid, not the bench contract)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pywgcna_b200 import _lib as L
from pywgcna_b200.synth import make_expression

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--m", type=int, default=200)
ap.add_argument("--P", type=int, default=20)
ap.add_argument("--prec", default="f64")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--skip-tom", action="store_true")
a = ap.parse_args()
lib = L.load()
dev = 0
torch.cuda.set_device(dev)
st = torch.cuda.current_stream().cuda_stream
n, m, P = a.n, a.m, a.P
X = torch.from_numpy(make_expression(m, n, F=40, seed=1)).cuda()
ldk = lib.b200w_padded_k(m)
Z = torch.empty((n, ldk), dtype=torch.float64, device="cuda")
bad = torch.empty(n, dtype=torch.uint8, device="cuda")
K = torch.empty((P, n), dtype=torch.float64, device="cuda")
A = torch.empty((n, n), dtype=torch.float64, device="cuda")
steps = np.ones(P)


def timeit(name, fn, flops=None, bytes_=None):
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
    t = min(ts)
    extra = ""
    if flops:
        extra += f"  {flops / t / 1e9:8.2f} TFLOP/s"
    if bytes_:
        extra += f"  {bytes_ / t / 1e6:8.1f} GB/s"
    print(f"{name:28s} {t:10.3f} ms{extra}", flush=True)


timeit("standardize", lambda: L.check(lib.b200w_dev_standardize(X.data_ptr(), m, n, Z.data_ptr(), bad.data_ptr(), dev, st)),
       bytes_=8 * m * n * 4)
timeit("sweep", lambda: L.check(lib.b200w_dev_sweep(Z.data_ptr(), bad.data_ptr(), m, n, steps.ctypes.data, P, 2, 0, n, K.data_ptr(), dev, st)),
       flops=2 * m * n * n + 2 * P * n * n)
timeit("adjacency", lambda: L.check(lib.b200w_dev_adjacency(Z.data_ptr(), bad.data_ptr(), m, n, 2, 6.0, 0, n, A.data_ptr(), dev, st)),
       flops=2 * m * n * n, bytes_=8 * n * n)
stats = torch.empty(6, dtype=torch.float64, device="cuda")
timeit("check_adjacency", lambda: L.check(lib.b200w_dev_check_adjacency(A.data_ptr(), n, stats.data_ptr(), dev, st)), bytes_=8 * n * n)
print("stats", stats.cpu().numpy())
if not a.skip_tom:
    prec = {"f64": L.PREC_F64, "i8": L.PREC_I8}[a.prec]
    op = torch.zeros(lib.b200w_tom_operand_bytes(n, prec), dtype=torch.uint8, device="cuda")
    oprows = (n + 127) // 128 * 128
    scale = torch.zeros(oprows, dtype=torch.float64, device="cuda")
    k = torch.zeros(oprows, dtype=torch.float64, device="cuda")
    out = torch.empty((n, n), dtype=torch.float64, device="cuda")
    timeit("tom_prepare", lambda: L.check(lib.b200w_dev_tom_prepare(A.data_ptr(), n, 0, n, prec, op.data_ptr(), oprows, scale.data_ptr(), k.data_ptr(), dev, st)),
           bytes_=8 * n * n)
    timeit("tom_compute", lambda: L.check(lib.b200w_dev_tom_compute(A.data_ptr(), op.data_ptr(), oprows, scale.data_ptr(), k.data_ptr(), n, 0, n, 0, 0, 0, prec, out.data_ptr(), dev, st)),
           flops=2 * n ** 3)
    print("TOM sample", out[1, :4].cpu().numpy(), "launches", lib.b200w_launch_count())
