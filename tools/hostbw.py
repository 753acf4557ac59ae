import time, numpy as np, torch
n = 20000
nbytes = n * n * 8
def T(name, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    dt = time.perf_counter() - t; print(f"{name:34s} {dt*1e3:9.1f} ms  {nbytes/dt/1e9:7.2f} GB/s", flush=True); return r
d = torch.empty(n * n, dtype=torch.float64, device="cuda")
a = T("np.empty+touch (page faults)", lambda: np.ones(n * n))
T("H2D pageable", lambda: d.copy_(torch.from_numpy(a)))
T("H2D pageable again", lambda: d.copy_(torch.from_numpy(a)))
b = np.empty(n * n)
T("D2H into fresh np.empty", lambda: torch.from_numpy(b).copy_(d))
T("D2H into touched np", lambda: torch.from_numpy(b).copy_(d))
p = T("cudaHostAlloc 3.2GB (torch pin)", lambda: torch.empty(n * n, dtype=torch.float64).pin_memory())
T("H2D pinned", lambda: d.copy_(p, non_blocking=True))
T("D2H pinned", lambda: p.copy_(d, non_blocking=True))
T("cudaHostRegister 3.2GB", lambda: torch.cuda.cudart().cudaHostRegister(a.ctypes.data, nbytes, 0))
T("H2D registered", lambda: d.copy_(torch.from_numpy(a), non_blocking=True))
T("cudaHostUnregister", lambda: torch.cuda.cudart().cudaHostUnregister(a.ctypes.data))
c = T("np.copy 3.2GB (1 thread memcpy)", lambda: a.copy())
import pandas as pd
T("pd.DataFrame(ndarray)", lambda: pd.DataFrame(a.reshape(n, n)))
