"""Worker of tests/test_gpu_round2.py::test_ranks_real_nccl_and_ipc -- one process per GPU under
torch.distributed.run.  Runs the REAL sharded path (NCCL collectives, CUDA-IPC peer stores from inside the
tcgen05 kernel, copy-engine pulls of the digit planes gated per plane) and compares, on rank 0,

  * the union of the ranks' TOM row blocks with the single-GPU result of the same library: bit for bit
    (exact integer accumulation makes the sharded result independent of the number of GPUs),
  * with the float64 numpy restatement of the reference (oracle/restated.py): 1e-10,
  * k of the sweep, the selected power and the adjacency blocks likewise,

for both exchange modes ("pull", "gather"), both schedules of the sharded sweep ("symmetric": block pairs dealt out over
the ranks, similarity stored into the peers' blocks from inside the kernel; "full": column blocks) and for sizes that
leave the last rank short or empty.
Prints one JSON line on rank 0; exit code 0 only if every comparison held.
"""
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
    from oracle import restated as O
    from pywgcna_b200.device import DeviceNetwork
    from pywgcna_b200.synth import make_expression

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sizes = [int(s) for s in (sys.argv[1] if len(sys.argv) > 1 else "2304,2500,600,200").split(",")]
    powers = list(range(1, 11)) + list(range(11, 21, 2))
    report, ok = [], True
    for n in sizes:
        X = make_expression(48, n, F=5, seed=1000 + n)
        X[:, 7] = 3.25  # a constant gene: NaN row / column in the correlation, zeroed by the TOM
        # operand exchange of the TOM x schedule of the sharded sweep (device.py: sym_sweep)
        for mode, sweep_mode in (("pull", "symmetric"), ("gather", "symmetric"), ("pull", "full"), ("gather", "full")):
            os.environ["B200WGCNA_EXCHANGE"] = mode
            os.environ["B200WGCNA_SWEEP_DIST"] = sweep_mode
            net = DeviceNetwork(X, device=local, networkType="signed hybrid")
            for rep in range(2):  # twice: the second pass reuses buffers, epochs and peer mappings
                power, tab = net.sft(powers)
                A = net.adjacency(power)
                T = net.tom(TOMType="signed", TOMDenom="min")
            torch.cuda.synchronize()
            per, n_pad = net.per, net.per * world
            full_T = torch.zeros((n_pad, n), dtype=torch.float64, device=net.dev)
            full_A = torch.zeros((n_pad, n), dtype=torch.float64, device=net.dev)
            blk_T = torch.zeros((per, n), dtype=torch.float64, device=net.dev)
            blk_A = torch.zeros((per, n), dtype=torch.float64, device=net.dev)
            blk_T[:net.nrows] = T
            blk_A[:net.nrows] = A
            dist.all_gather_into_tensor(full_T, blk_T)
            dist.all_gather_into_tensor(full_A, blk_A)
            k_dist = net.sweep(powers).cpu().numpy()
            p_dist = float(power)
            if rank == 0:
                one = DeviceNetwork(X, device=local, networkType="signed hybrid", world=1, rank=0, use_dist=False)
                p1, _ = one.sft(powers)
                A1 = one.adjacency(p1).clone()  # (the sweep below reuses the adjacency buffer)
                T1 = one.tom(TOMType="signed", TOMDenom="min")
                k1 = one.sweep(powers).cpu().numpy()
                torch.cuda.synchronize()
                Td, Ad = full_T[:n].cpu().numpy(), full_A[:n].cpu().numpy()
                A_ref = O.adjacency(X, adjacencyType="signed hybrid", power=p_dist)
                T_ref = O.TOMsimilarity(A_ref.copy(), TOMType="signed").to_numpy()
                rec = {"n": n, "mode": mode, "sweep": sweep_mode, "world": world, "per": per, "power": p_dist,
                       "power_equal": p_dist == float(p1),
                       "k_bitwise": bool(np.array_equal(k_dist, k1)),
                       "k_max_abs": float(np.abs(k_dist - k1).max()),
                       "adj_bitwise": bool(np.array_equal(Ad, A1.cpu().numpy(), equal_nan=True)),
                       "tom_bitwise": bool(np.array_equal(Td, T1.cpu().numpy())),
                       "tom_vs_oracle": float(np.abs(Td - T_ref).max()),
                       "adj_vs_oracle": float(np.nanmax(np.abs(Ad - A_ref))),
                       "table_power": float(tab.host_power)}
                if not rec["adj_bitwise"]:
                    A1n = A1.cpu().numpy()
                    diff = ~((Ad == A1n) | (np.isnan(Ad) & np.isnan(A1n)))
                    ii, jj = np.nonzero(diff)
                    rec["adj_ndiff"] = int(diff.sum())
                    rec["adj_first_diffs"] = [(int(a), int(b), float(Ad[a, b]), float(A1n[a, b]))
                                              for a, b in list(zip(ii, jj))[:6]]
                good = (rec["power_equal"] and rec["k_max_abs"] <= 1e-9 and rec["adj_bitwise"] and rec["tom_bitwise"]
                        and rec["tom_vs_oracle"] <= 1e-10 and rec["adj_vs_oracle"] <= 1e-11
                        and rec["table_power"] == p_dist)
                rec["ok"] = bool(good)
                ok = ok and good
                report.append(rec)
                del one
            del net
            dist.barrier()
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.broadcast(flag, 0)
    if rank == 0:
        print(json.dumps({"ok": ok, "cases": report}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
