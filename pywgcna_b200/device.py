"""Device-resident chaining of the hot path and its row-block sharding over
several GPUs.  PyTorch is used for device memory, streams and ``torch.distributed``
only; every kernel launched here is one of libb200wgcna's, through the dev_* entry points of
``include/b200wgcna.h``.

Single GPU:   DeviceNetwork(X).sweep(powers) -> k ; .adjacency(power) ; .tom() -> TOM in HBM
Several GPUs: one process per GPU (torchrun); rank r owns the contiguous row block
              ``row_block(n, world, r)`` of the adjacency / TOM and the same block of sweep columns.
              Exchange steps (NCCL all-gather): k after the sweep; the TOM operand row blocks (int8
              digit planes or FP64 rows), their scales and k before the contraction.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .shard import all_gather_rows, row_block, rows_per_rank


class DeviceNetwork:
    """Expression matrix resident in HBM; sweep / adjacency / TOM without leaving the device."""

    def __init__(self, X, device: int = 0, networkType: str = "unsigned", group=None, world: int | None = None,
                 rank: int | None = None, precision: str = "auto"):
        self.lib = L.load()
        self.device = int(device)
        torch.cuda.set_device(self.device)
        self.dev = torch.device("cuda", self.device)
        self.net = L.NET_TYPES.index(networkType)
        self.prec = {"auto": L.PREC_AUTO, "f64": L.PREC_F64, "i8": L.PREC_I8}[precision]
        self.group = group
        if world is None:
            if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
                import torch.distributed as dist

                world, rank = dist.get_world_size(group), dist.get_rank(group)
            else:
                world, rank = 1, 0
        self.world, self.rank = int(world), int(rank)
        if isinstance(X, torch.Tensor):
            self.X = X.to(self.dev, torch.float64).contiguous()
        else:
            self.X = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(self.dev)
        self.m, self.n = self.X.shape
        self.ldk = self.lib.b200w_padded_k(self.m)
        self.ldn = self.lib.b200w_padded_n(self.n)
        self.Z = torch.empty((self.n, self.ldk), dtype=torch.float64, device=self.dev)
        self.bad = torch.empty(self.n, dtype=torch.uint8, device=self.dev)
        self.per = rows_per_rank(self.n, self.world)
        self.row0, self.nrows = row_block(self.n, self.world, self.rank)
        self.op_rows = self.per * self.world  # >= round_up(n, 128); equal blocks for the all-gather
        self.A = None
        self.T = None
        self._op = None
        self._standardized = False

    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def standardize(self):
        L.check(self.lib.b200w_dev_standardize(self.X.data_ptr(), self.m, self.n, self.Z.data_ptr(),
                                               self.bad.data_ptr(), self.device, self._stream()))
        self._standardized = True

    def _all_gather_rows(self, full: torch.Tensor):
        """In-place all-gather of the ``per``-row blocks of ``full`` ([op_rows, ...], contiguous)."""
        all_gather_rows(full, self.per, self.rank, self.world, self.group)

    # -- the three stages ------------------------------------------------------------------
    def sweep(self, powers) -> torch.Tensor:
        """k[P, n] (device, replicated on every rank) for the ascending power vector
        (wgcna.py:873-920)."""
        if not self._standardized:
            self.standardize()
        powers = np.sort(np.asarray(powers, dtype=np.float64))
        steps = np.ascontiguousarray(np.diff(np.concatenate(([0.0], powers))))
        P = len(powers)
        col0, ncols = (0, self.n) if self.world == 1 else (self.row0, self.nrows)
        kloc = torch.empty((P, max(ncols, 1)), dtype=torch.float64, device=self.dev)
        if ncols > 0:
            L.check(self.lib.b200w_dev_sweep(self.Z.data_ptr(), self.bad.data_ptr(), self.m, self.n,
                                             steps.ctypes.data, P, self.net, col0, ncols, kloc.data_ptr(),
                                             self.device, self._stream()))
        if self.world == 1:
            return kloc
        kT = torch.zeros((self.op_rows, P), dtype=torch.float64, device=self.dev)
        if ncols > 0:
            kT[self.row0:self.row0 + ncols] = kloc[:, :ncols].T
        self._all_gather_rows(kT)
        return kT[:self.n].T.contiguous()

    def adjacency(self, power: float) -> torch.Tensor:
        """This rank's row block of the adjacency (wgcna.py:1097-1111), kept in HBM."""
        if not self._standardized:
            self.standardize()
        if self.A is None:
            self.A = torch.empty((max(self.nrows, 1), self.n), dtype=torch.float64, device=self.dev)
        if self.nrows > 0:
            L.check(self.lib.b200w_dev_adjacency(self.Z.data_ptr(), self.bad.data_ptr(), self.m, self.n, self.net,
                                                 float(power), self.row0, self.nrows, self.A.data_ptr(),
                                                 self.device, self._stream()))
        return self.A[:self.nrows]

    def _operand(self):
        if self._op is None:
            elt = self.lib.b200w_tom_operand_bytes(128, self.prec) // (128 * 128)  # bytes / element / all planes
            self.planes, self.elt = (1, 8) if elt == 8 else (elt, 1)
            self._op = torch.zeros((self.planes, self.op_rows, self.ldn * self.elt), dtype=torch.uint8, device=self.dev)
            self._scale = torch.zeros(self.op_rows, dtype=torch.float64, device=self.dev)
            self._k = torch.zeros(self.op_rows, dtype=torch.float64, device=self.dev)
        return self._op

    def tom(self, TOMType: str = "signed", TOMDenom: str = "min", out_mode: int = L.OUT_TOM,
            A_rows: torch.Tensor | None = None) -> torch.Tensor:
        """This rank's row block of the TOM (wgcna.py:1141-1157) from its adjacency row block."""
        tt, td = L.TOM_TYPES.index(TOMType), L.TOM_DENOMS.index(TOMDenom)
        A_rows = self.A[:self.nrows] if A_rows is None else A_rows
        n = self.n
        op = self._operand()
        if self.T is None:
            self.T = torch.empty((max(self.nrows, 1), n), dtype=torch.float64, device=self.dev)
        if self.nrows > 0:
            L.check(self.lib.b200w_dev_tom_prepare(A_rows.data_ptr(), n, self.row0, self.nrows, self.prec,
                                                   op.data_ptr(), self.op_rows, self._scale.data_ptr(),
                                                   self._k.data_ptr(), self.device, self._stream()))
        if self.world > 1:
            for p in range(self.planes):
                self._all_gather_rows(op[p])
            self._all_gather_rows(self._k)
            self._all_gather_rows(self._scale)
        if self.nrows > 0:
            L.check(self.lib.b200w_dev_tom_compute(A_rows.data_ptr(), op.data_ptr(), self.op_rows,
                                                   self._scale.data_ptr(), self._k.data_ptr(), n, self.row0,
                                                   self.nrows, tt, td, out_mode, self.prec, self.T.data_ptr(),
                                                   self.device, self._stream()))
        return self.T[:self.nrows]
