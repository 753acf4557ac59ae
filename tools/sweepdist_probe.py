"""Where the time of the sharded symmetric sweep goes, per rank (run under torch.distributed.run):
token all-reduce / kernel with peer stores / kernel with every store kept local / all-reduce of the partial k."""
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
    from pywgcna_b200 import _lib as L
    from pywgcna_b200.device import DeviceNetwork
    from pywgcna_b200.synth import make_expression

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    m = 500
    X = make_expression(m, n, F=10, seed=3)
    os.environ["B200WGCNA_SWEEP_DIST"] = "symmetric"
    net = DeviceNetwork(X, device=local, networkType="unsigned")
    assert net.sym_sweep
    powers = np.array(list(range(1, 11)) + list(range(12, 31, 2)), dtype=np.float64)
    steps = np.ascontiguousarray(np.diff(np.concatenate(([0.0], powers))))
    P = len(powers)
    net.sweep(powers)
    torch.cuda.synchronize()
    lib = net.lib
    own = (C.c_void_p * world)(*[int(net._A_raw)] * world)
    kpart = torch.empty((P, n), dtype=torch.float64, device=net.dev)
    tok = torch.zeros(1, device=net.dev)
    st = torch.cuda.current_stream().cuda_stream

    def timed(fn, reps=3):
        out = []
        for _ in range(reps):
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            out.append(round(a.elapsed_time(b), 3))
        return out

    def kern(ptrs):
        L.check(lib.b200w_dev_sweep_dist(net.Z.data_ptr(), net.bad.data_ptr(), net.m, net.n, steps.ctypes.data, P,
                                         net.net, world, rank, net.per, ptrs, kpart.data_ptr(), net.device, st))

    def mixed(peers):  # real pointers for the given peer offsets, every other store kept local
        v = [int(net._A_raw)] * world
        for d in peers:
            q = (rank + d) % world
            v[q] = int(net._peer_A_ptrs[q])
        return (C.c_void_p * world)(*v)

    res = {"rank": rank,
           "token_allreduce_ms": timed(lambda: dist.all_reduce(tok)),
           "kernel_peer_stores_ms": timed(lambda: kern(net._peer_A_ptrs)),
           "kernel_local_stores_ms": timed(lambda: kern(own)),
           "kernel_one_peer_ms": timed(lambda: kern(mixed([1]))),
           "kernel_two_peers_ms": timed(lambda: kern(mixed([1, 2]))),
           "kpart_allreduce_ms": timed(lambda: dist.all_reduce(kpart)),
           "sweep_call_ms": timed(lambda: net.sweep(powers))}
    os.environ["B200WGCNA_SWEEP_DIST"] = "full"
    net2 = DeviceNetwork(X, device=local, networkType="unsigned")
    net2.sweep(powers)
    res["sweep_call_full_ms"] = timed(lambda: net2.sweep(powers))
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        print(json.dumps({"n": n, "world": world, "ranks": allres}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
