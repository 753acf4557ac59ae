"""Copy-engine bandwidth between every ordered pair of visible GPUs (GB/s) and the driver's topology matrix."""
import json
import subprocess
import sys

import torch

n = torch.cuda.device_count()
size = 256 << 20
bufs = [torch.empty(size, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
mat = [[None] * n for _ in range(n)]
for i in range(n):
    for j in range(n):
        if i == j:
            continue
        torch.cuda.set_device(i)
        for _ in range(2):
            bufs[j].copy_(bufs[i], non_blocking=True)
        torch.cuda.synchronize(i)
        torch.cuda.synchronize(j)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4):
            bufs[j].copy_(bufs[i], non_blocking=True)
        b.record()
        torch.cuda.synchronize(i)
        torch.cuda.synchronize(j)
        mat[i][j] = round(4 * size / a.elapsed_time(b) / 1e6, 1)
access = [[bool(torch.cuda.can_device_access_peer(i, j)) if i != j else None for j in range(n)] for i in range(n)]
topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
print(json.dumps({"gpus": n, "copy_GBps": mat, "can_access_peer": access}))
print(topo)
