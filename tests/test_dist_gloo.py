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


# ------------------------------------------------------------------------------------------------
# the hardware parity probe of bench.py (every N > 1 line carries its result): the checker itself, on CPU
# ------------------------------------------------------------------------------------------------
class _FakeNet:
    """The members of DeviceNetwork that bench.tom_parity_probe reads, over CPU tensors."""

    def __init__(self, A_block, n, world, rank, per, row0):
        self.A, self.n, self.world, self.rank, self.per = A_block, n, world, rank, per
        self.row0, self.nrows = row0, A_block.shape[0]
        self.dev = torch.device("cpu")


def _probe_worker(rank, world, port, n, tom_type, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys

        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench
        from oracle import restated as O
        from pywgcna_b200.synth import make_expression

        X = make_expression(40, n, F=4, seed=n)
        X[:, 3] = 1.0  # a NaN gene: the probe has to clean it like the TOM does (wgcna.py:1185)
        A = O.adjacency(X, adjacencyType="unsigned", power=4)
        T = O.TOMsimilarity(A.copy(), TOMType=tom_type).to_numpy()
        per = shard.rows_per_rank(n, world, align=8)
        r0, nr = shard.row_block(n, world, rank, align=8)
        net = _FakeNet(torch.from_numpy(A[r0:r0 + nr].copy()), n, world, rank, per, r0)
        good = bench.tom_parity_probe(net, torch.from_numpy(T[r0:r0 + nr].copy()), tom_type, rows_per_rank=8)
        Tbad = T[r0:r0 + nr].copy()
        last_with_rows = max(r for r in range(world) if shard.row_block(n, world, r, align=8)[1] > 0)
        if rank == last_with_rows:
            Tbad[:, 5] += 1e-6  # one rank's block is off: the max over ranks must show it
        bad = bench.tom_parity_probe(net, torch.from_numpy(Tbad), tom_type, rows_per_rank=8)
        if rank == 0:
            np.save(os.path.join(out_dir, "probe.npy"), np.array([good["max_abs"], good["rows"], bad["max_abs"]]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,tom_type", [(2, 97, "unsigned"), (3, 61, "signed"), (2, 6, "unsigned")])
def test_bench_parity_probe_is_a_correct_checker(tmp_path, world, n, tom_type):
    """bench.py's probe (sampled TOM rows against the float64 formula accumulated over the ranks' adjacency
    blocks and all-reduced) agrees with the oracle to rounding when the blocks are right -- including a trailing
    rank without rows (n = 6, world = 2, align 8) -- and reports a block that is off by 1e-6."""
    mp.spawn(_probe_worker, args=(world, _free_port(), n, tom_type, str(tmp_path)), nprocs=world, join=True)
    good, rows, bad = np.load(tmp_path / "probe.npy")
    assert good < 1e-12 and rows >= 1
    assert 5e-7 < bad < 2e-6
