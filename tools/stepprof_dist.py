"""Per-phase times of one bench step under torchrun (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200WGCNA_QUIET", "1")
import numpy as np, torch, torch.distributed as dist
from pywgcna_b200 import wgcna as W
from pywgcna_b200.device import DeviceNetwork
from pywgcna_b200.synth import make_expression
local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m, n = 200, 20000
powers = list(range(1, 21))
X = torch.from_numpy(make_expression(m, n, F=40, seed=20200)).pin_memory()
net = DeviceNetwork(X, device=local, networkType="signed hybrid", world=world, rank=rank)
def T(name, fn):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    if rank == 0: print(f"{name:24s} {1e3*(time.perf_counter()-t):9.3f} ms", flush=True)
    return r
for rep in range(3):
    if rank == 0: print("--- rep", rep)
    T("standardize", net.standardize)
    k = T("sweep(+allgather k)", lambda: net.sweep(powers))
    kh = T("k.cpu", lambda: k.cpu().numpy())
    T("fit_table", lambda: W.fit_table(kh, 10))
    T("adjacency", lambda: net.adjacency(11))
    T("tom_prepare", net.tom_prepare)
    T("tom_exchange", net.tom_exchange)
    T("tom_compute", lambda: net.tom_compute(TOMType="signed"))
    T("tom_finish", net.tom_finish)
if world > 1: dist.destroy_process_group()
