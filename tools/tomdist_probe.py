"""Does the sharded TOM kernel wait for its stores into the peers' row blocks?  Per rank (under torch.distributed.run):
tom() with the real peer mappings / with every mirrored store kept local / with only rank + 1 mapped."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200WGCNA_QUIET", "1")
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from pywgcna_b200.device import DeviceNetwork
    from pywgcna_b200.synth import make_expression

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    X = make_expression(500, n, F=10, seed=3)
    net = DeviceNetwork(X, device=local, networkType="unsigned")
    powers = list(range(1, 11)) + list(range(12, 31, 2))
    power, _ = net.sft(powers)
    net.adjacency(power)
    net.tom(TOMType="unsigned")
    torch.cuda.synchronize()
    real = net._peer_ptrs

    def table(peers):
        v = [int(net._T_raw)] * world
        for d in peers:
            q = (rank + d) % world
            v[q] = int(real[q])
        return (C.c_void_p * world)(*v)

    def timed(ptrs, reps=3):
        out = []
        net._peer_ptrs = ptrs
        for _ in range(reps):
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            net.tom(TOMType="unsigned")
            b.record()
            torch.cuda.synchronize()
            out.append(round(a.elapsed_time(b), 2))
        net._peer_ptrs = real
        return out

    res = {"rank": rank, "tom_real_peers_ms": timed(real), "tom_local_stores_ms": timed(table([])),
           "tom_one_peer_ms": timed(table([1])), "tom_real_again_ms": timed(real)}
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        print(json.dumps({"n": n, "world": world, "ranks": allres}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
