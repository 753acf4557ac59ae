"""Where the wall time of the sharded host entry goes (development aid): network_blocks under torchrun with a
synchronisation after every stage.   torchrun --nproc-per-node N tools/e2e_blocks_trace.py [C3|C2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200WGCNA_QUIET", "1")
import numpy as np, pandas as pd, torch, torch.distributed as dist
import pywgcna_b200 as bw
from pywgcna_b200 import _lib as L
from pywgcna_b200 import device as D
from pywgcna_b200.synth import make_expression

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
m, n, F, seed, net, tom = {"C3": (500, 50000, 60, 50500, "unsigned", "unsigned"), "C2": (200, 20000, 40, 20200, "signed hybrid", "signed")}[wl]
df = pd.DataFrame(make_expression(m, n, F=F, seed=seed))
powers = list(range(1, 21))
lib = L.load()
for rep in range(4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    X = L.as_f64_matrix(df); t.append(time.perf_counter())
    res = bw.network_blocks(df, powerVector=powers, networkType=net, TOMType=tom, device=local); t.append(time.perf_counter())
    del res; t.append(time.perf_counter())
    # the same stages by hand, synchronised
    netw = next(iter(D._NET_CACHE.values()))
    st = torch.cuda.current_stream(netw.dev).cuda_stream
    s = [time.perf_counter()]
    L.check(lib.b200w_dev_upload(netw.X.data_ptr(), L.hptr(X), X.nbytes, local, st)); s.append(time.perf_counter())
    netw._standardized = False; netw._sim_in_A = False
    p, tab = netw.sft(np.asarray(powers)); A = netw.adjacency(p); T = netw.tom(TOMType=tom); torch.cuda.synchronize(); s.append(time.perf_counter())
    _ = tab["SFT.R.sq"]; s.append(time.perf_counter())
    out = L.result_empty((netw.nrows, n)); s.append(time.perf_counter())
    L.check(lib.b200w_dev_download(L.hptr(out), T.data_ptr(), out.nbytes, local, st)); s.append(time.perf_counter())
    del out; s.append(time.perf_counter())
    if rank == 0:
        print(f"rep {rep}: as_f64 {t[1]-t[0]:.3f}  network_blocks {t[2]-t[1]:.3f}  del {t[3]-t[2]:.3f} | by hand: upload {s[1]-s[0]:.3f}  "
              f"chain {s[2]-s[1]:.3f}  table {s[3]-s[2]:.3f}  alloc {s[4]-s[3]:.3f}  download {s[5]-s[4]:.3f}  del {s[6]-s[5]:.3f}", flush=True)
if world > 1:
    dist.destroy_process_group()
