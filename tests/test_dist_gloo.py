"""N > 1 host logic on CPU: world_size-2 (and 3) gloo process groups exercise the row-block partition
and the all-gather exchange of pywgcna_b200/shard.py with a sharded numpy evaluation of the TOM formula
(each rank: its adjacency rows -> k of its rows -> exchange of operand rows and k -> its TOM rows), and
check that the union of the row blocks equals the oracle's single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pywgcna_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import restated as O
        from pywgcna_b200.synth import make_expression

        X = make_expression(40, n, F=4, seed=n)  # replicated input, as in DeviceNetwork
        A = O.adjacency(X, adjacencyType="unsigned", power=6)
        np.fill_diagonal(A, 0)
        per = shard.rows_per_rank(n, world, align=8)
        r0, nr = shard.row_block(n, world, rank, align=8)
        rows_pad = shard.padded_rows(n, world, align=8)
        # "prepare": this rank's rows of the operand and their k, written at their offset of the full buffers
        op = torch.zeros((rows_pad, n), dtype=torch.float64)
        k = torch.zeros((rows_pad, 1), dtype=torch.float64)
        op[r0:r0 + nr] = torch.from_numpy(A[r0:r0 + nr])
        k[r0:r0 + nr, 0] = torch.from_numpy(A[r0:r0 + nr].sum(axis=1))
        # exchange
        shard.all_gather_rows(op, per, rank, world)
        shard.all_gather_rows(k, per, rank, world)
        # "compute": own rows against the full operand
        opn, kn = op[:n].numpy(), k[:n, 0].numpy()
        L = opn[r0:r0 + nr] @ opn.T
        a = opn[r0:r0 + nr]
        T = (L + a) / (np.minimum(kn[r0:r0 + nr, None], kn[None, :]) + 1 - a)
        T[np.arange(nr), np.arange(r0, r0 + nr)] = 1.0
        np.save(os.path.join(out_dir, f"tom_{rank}.npy"), T)
        np.save(os.path.join(out_dir, f"blk_{rank}.npy"), np.array([r0, nr]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 97), (3, 130), (2, 8)])
def test_sharded_tom_equals_single_process(tmp_path, world, n):
    from oracle import restated as O
    from pywgcna_b200.synth import make_expression

    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    X = make_expression(40, n, F=4, seed=n)
    A = O.adjacency(X, adjacencyType="unsigned", power=6)
    T_ref = O.TOMsimilarity(A.copy(), TOMType="unsigned").to_numpy()
    T = np.full((n, n), np.nan)
    covered = 0
    for r in range(world):
        r0, nr = np.load(tmp_path / f"blk_{r}.npy")
        T[r0:r0 + nr] = np.load(tmp_path / f"tom_{r}.npy")
        covered += nr
    assert covered == n
    np.testing.assert_allclose(T, T_ref, rtol=0, atol=1e-12)


def test_row_block_partition_properties():
    for n in (1, 7, 128, 129, 20000, 23844, 50000, 80000):
        for world in (1, 2, 3, 4, 8):
            per = shard.rows_per_rank(n, world)
            assert per % 128 == 0 and per * world >= n
            blocks = [shard.row_block(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(b[1] for b in blocks) == n
            for (a0, an), (b0, bn) in zip(blocks, blocks[1:]):
                assert b0 == min(n, a0 + per) and an in (per, n - a0) or an == 0
            assert all(b[0] % 128 == 0 or b[1] == 0 for b in blocks)
