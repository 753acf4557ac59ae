"""Device-resident chaining of the hot path and its sharding over several GPUs
(SURVEY 8e).  PyTorch is used for device memory, streams and ``torch.distributed`` only; every kernel
launched here is one of libb200wgcna's, through the dev_* entry points of ``include/b200wgcna.h``.

Single GPU:   DeviceNetwork(X).sweep(powers) -> k ; .adjacency(power) ; .tom() -> TOM in HBM
Several GPUs: one process per GPU (torchrun).  Rank r owns the contiguous row block
              ``row_block(n, world, r, 256)`` of the adjacency / TOM and the same block of sweep columns.
  sweep       own columns, then all-gather of k (NCCL)
  adjacency   own rows
  TOM         own rows -> int8 digit planes; all-gather of the planes, their scales and k (NCCL, 1 byte
              per element and plane instead of 8); then the SYMMETRIC contraction: the upper triangle of
              256 x 256 tiles is dealt out over the ranks and every tile's mirrored half is stored by the
              tcgen05 kernel itself into the owning peer's row block over NVLink (CUDA-IPC mapped
              buffers) -- compute and exchange in one kernel.  ``tom_finish()`` stream-orders a tiny
              all-reduce after it, which is the point where every rank's block is complete.
              ``B200WGCNA_DIST=rowblock`` selects the plain row-block contraction (no peer stores).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L
from .shard import all_gather_rows, row_block, rows_per_rank

ALIGN = 256  # row blocks start on 256-row tile boundaries (the CTA-pair tile of the int8 kernel)


class _DevView:
    """__cuda_array_interface__ over a raw device pointer, so torch can view library-owned memory."""

    def __init__(self, ptr: int, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class DeviceNetwork:
    """Expression matrix resident in HBM; sweep / adjacency / TOM without leaving the device."""

    def __init__(self, X, device: int = 0, networkType: str = "unsigned", group=None, world: int | None = None,
                 rank: int | None = None, precision: str = "auto", use_dist: bool | None = None,
                 keep_similarity: bool | None = None):
        self.lib = L.load()
        self.device = int(device)
        torch.cuda.set_device(self.device)
        self.dev = torch.device("cuda", self.device)
        self.net = L.NET_TYPES.index(networkType)
        self.prec = {"auto": L.PREC_AUTO, "f64": L.PREC_F64, "i8": L.PREC_I8, "i8-fast": L.PREC_I8_FAST}[precision]
        self.group = group
        self.use_dist = use_dist  # torch.distributed collectives (False: the caller moves the blocks, tests)
        if world is None:
            if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
                import torch.distributed as dist

                world, rank = dist.get_world_size(group), dist.get_rank(group)
                if self.use_dist is None:
                    self.use_dist = True
            else:
                world, rank = 1, 0
        elif self.use_dist is None:
            self.use_dist = torch.distributed.is_available() and torch.distributed.is_initialized()
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
        self.per = rows_per_rank(self.n, self.world, ALIGN)
        self.row0, self.nrows = row_block(self.n, self.world, self.rank, ALIGN)
        self.op_rows = self.per * self.world  # >= round_up(n, 128); equal blocks for the all-gather
        elt = self.lib.b200w_tom_operand_bytes(128, self.prec) // (128 * 128)  # bytes per element over all planes
        self.planes, self.elt = (1, 8) if elt == 8 else (elt, 1)
        self.fused = (self.world > 1 and self.elt == 1
                      and os.environ.get("B200WGCNA_DIST", "fused") != "rowblock")
        self.A = None
        self._sim_in_A = False  # self.A holds the similarity the last sweep left (see sweep / adjacency)
        # sweep() leaves this rank's rows of the similarity in the adjacency buffer (overwriting a block an
        # earlier adjacency() returned) and the next adjacency() finishes them elementwise
        if keep_similarity is None:
            keep_similarity = os.environ.get("B200WGCNA_KEEP_SIMILARITY", "1") != "0"
        self.keep_similarity = bool(keep_similarity)
        self.T = None
        self._T_raw = None
        self._peer_ptrs = None
        self._peer_open = []
        self._op = None
        self._standardized = False

    def __del__(self):
        try:
            for p in self._peer_open:
                self.lib.b200w_ipc_close(p, self.device)
            if self._T_raw:
                self.lib.b200w_dev_free(self._T_raw, self.device)
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def standardize(self):
        L.check(self.lib.b200w_dev_standardize(self.X.data_ptr(), self.m, self.n, self.Z.data_ptr(),
                                               self.bad.data_ptr(), self.device, self._stream()))
        self._standardized = True
        self._sim_in_A = False

    def _all_gather_rows(self, full: torch.Tensor):
        """In-place all-gather of the ``per``-row blocks of ``full`` ([op_rows, ...], contiguous)."""
        if self.use_dist:
            all_gather_rows(full, self.per, self.rank, self.world, self.group)

    # -- sweep and adjacency -------------------------------------------------------------------
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
        if ncols > 0 and self.keep_similarity:
            # the sweep also writes this rank's rows of the transformed similarity into the adjacency buffer,
            # and adjacency() of the same data finishes them with an elementwise S ** power instead of forming
            # the correlation a second time as the reference does (wgcna.py:882 and :1097)
            if self.A is None:
                self.A = torch.empty((max(self.nrows, 1), self.n), dtype=torch.float64, device=self.dev)
            L.check(self.lib.b200w_dev_sweep_similarity(self.Z.data_ptr(), self.bad.data_ptr(), self.m, self.n,
                                                        steps.ctypes.data, P, self.net, col0, ncols,
                                                        kloc.data_ptr(), self.A.data_ptr(), self.device,
                                                        self._stream()))
            self._sim_in_A = True
        elif ncols > 0:
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

    def sft(self, powers, nBreaks: int = 10, RsquaredCut: float = 0.9, MeanCut: float = 100):
        """Device-resident ``pickSoftThreshold`` (wgcna.py:873-953): sweep, then histograms, medians and
        exact mantissa sums on the device; 2 x P ten-point regressions on the host.  Returns
        (selected power, dict of the six table columns)."""
        from . import wgcna as W

        pv = np.sort(np.asarray(powers))
        k = self.sweep(pv)  # [P, n]
        P = len(pv)
        if not hasattr(self, "_sft_bufs") or self._sft_bufs[0].shape[0] != P or self._sft_bufs[3] != nBreaks:
            E = self.lib.b200w_sft_exp_buckets()
            self._sft_bufs = (torch.empty((P, 4 + 2 * nBreaks), dtype=torch.float64, device=self.dev),
                              torch.empty((P, E, 2), dtype=torch.int64, device=self.dev),
                              (torch.empty((P, 4 + 2 * nBreaks), dtype=torch.float64).pin_memory(),
                               torch.empty((P, E, 2), dtype=torch.int64).pin_memory()), nBreaks)
        d_stats, d_buck, (h_stats, h_buck), _ = self._sft_bufs
        L.check(self.lib.b200w_dev_sft_stats(k.data_ptr(), P, self.n, k.stride(0), nBreaks, d_stats.data_ptr(),
                                             d_buck.data_ptr(), self.device, self._stream()))
        h_stats.copy_(d_stats, non_blocking=True)
        h_buck.copy_(d_buck, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        r2, slope, r2adj, mean, med, mx = W.fit_table_from_stats(h_stats.numpy(), h_buck.numpy(), self.n, nBreaks)
        with np.errstate(invalid="ignore"):
            ind = (r2 > RsquaredCut) & (mean <= MeanCut)  # wgcna.py:945
        power = pv[ind].min() if ind.any() else pv[np.argsort(r2)[-1]]
        return power, {"SFT.R.sq": r2, "slope": slope, "truncated R.sq": r2adj, "mean(k)": mean,
                       "median(k)": med, "max(k)": mx}

    def adjacency(self, power: float) -> torch.Tensor:
        """This rank's row block of the adjacency (wgcna.py:1097-1111), kept in HBM."""
        if not self._standardized:
            self.standardize()
        if self.A is None:
            self.A = torch.empty((max(self.nrows, 1), self.n), dtype=torch.float64, device=self.dev)
        if self._sim_in_A:
            self._sim_in_A = False  # consumed: the buffer turns into the adjacency
            L.check(self.lib.b200w_dev_adjacency_from_similarity(self.A.data_ptr(), self.nrows, self.n,
                                                                 float(power), self.device, self._stream()))
            return self.A[:self.nrows]
        if self.nrows > 0:
            L.check(self.lib.b200w_dev_adjacency(self.Z.data_ptr(), self.bad.data_ptr(), self.m, self.n, self.net,
                                                 float(power), self.row0, self.nrows, self.A.data_ptr(),
                                                 self.device, self._stream()))
        return self.A[:self.nrows]

    # -- TOM -----------------------------------------------------------------------------------
    def _operand(self):
        if self._op is None:
            self._op = torch.zeros((self.planes, self.op_rows, self.ldn * self.elt), dtype=torch.uint8, device=self.dev)
            self._scale = torch.zeros(self.op_rows, dtype=torch.float64, device=self.dev)
            self._k = torch.zeros(self.op_rows, dtype=torch.float64, device=self.dev)
        return self._op

    def _out(self):
        """This rank's TOM row block.  In the fused multi-GPU mode it lives in a cudaMalloc buffer of
        ``per`` rows that the peers map through CUDA IPC."""
        if self.T is None:
            rows = self.per if self.fused else max(self.nrows, 1)
            if self.fused:
                p = C.c_void_p()
                L.check(self.lib.b200w_dev_malloc(rows * self.n * 8, self.device, C.byref(p)))
                self._T_raw = p.value
                self.T = torch.as_tensor(_DevView(p.value, (rows, self.n)), device=self.dev)
            else:
                self.T = torch.empty((rows, self.n), dtype=torch.float64, device=self.dev)
        return self.T

    def ipc_handle(self) -> bytes:
        self._out()
        h = C.create_string_buffer(64)
        L.check(self.lib.b200w_ipc_export(self._T_raw, h))
        return h.raw

    def connect_peers(self, ptrs=None):
        """Table of every rank's TOM block as seen from this GPU.  ``ptrs`` (tests: ranks emulated in one
        process) are used as they are; otherwise CUDA IPC handles are exchanged over the process group."""
        self._out()
        if ptrs is None:
            import torch.distributed as dist

            mine = torch.frombuffer(bytearray(self.ipc_handle()), dtype=torch.uint8).to(self.dev)
            allh = torch.zeros((self.world, 64), dtype=torch.uint8, device=self.dev)
            dist.all_gather_into_tensor(allh, mine, group=self.group)
            allh = allh.cpu().numpy()
            ptrs = []
            for r in range(self.world):
                if r == self.rank:
                    ptrs.append(self._T_raw)
                else:
                    p = C.c_void_p()
                    L.check(self.lib.b200w_ipc_open(allh[r].tobytes(), self.device, C.byref(p)))
                    self._peer_open.append(p.value)
                    ptrs.append(p.value)
        self._peer_ptrs = (C.c_void_p * self.world)(*[int(p) for p in ptrs])

    def tom_prepare(self, A_rows: torch.Tensor | None = None):
        if A_rows is None and self._sim_in_A:
            raise RuntimeError("DeviceNetwork: the adjacency buffer holds the similarity of the last sweep; "
                               "call adjacency(power) first")
        A_rows = self.A[:self.nrows] if A_rows is None else A_rows
        op = self._operand()
        self._A_rows = A_rows
        if self.nrows > 0:
            L.check(self.lib.b200w_dev_tom_prepare(A_rows.data_ptr(), self.n, self.row0, self.nrows, self.prec,
                                                   op.data_ptr(), self.op_rows, self._scale.data_ptr(),
                                                   self._k.data_ptr(), self.device, self._stream()))

    def tom_exchange(self):
        if self.world > 1:
            for p in range(self.planes):
                self._all_gather_rows(self._op[p])
            self._all_gather_rows(self._k)
            self._all_gather_rows(self._scale)

    def tom_compute(self, TOMType: str = "signed", TOMDenom: str = "min", out_mode: int = L.OUT_TOM):
        tt, td = L.TOM_TYPES.index(TOMType), L.TOM_DENOMS.index(TOMDenom)
        T = self._out()
        if self.fused and self._peer_ptrs is None:
            self.connect_peers()  # a collective: every rank takes part, also one without rows
        if self.nrows == 0:
            return T[:0]
        if self.fused:
            L.check(self.lib.b200w_dev_tom_compute_dist(self._A_rows.data_ptr(), self._op.data_ptr(), self.op_rows,
                                                        self._scale.data_ptr(), self._k.data_ptr(), self.n,
                                                        self.world, self.rank, self.per, tt, td, out_mode,
                                                        self._peer_ptrs, self.device, self._stream()))
        else:
            L.check(self.lib.b200w_dev_tom_compute(self._A_rows.data_ptr(), self._op.data_ptr(), self.op_rows,
                                                   self._scale.data_ptr(), self._k.data_ptr(), self.n, self.row0,
                                                   self.nrows, tt, td, out_mode, self.prec, T.data_ptr(),
                                                   self.device, self._stream()))
        return T[:self.nrows]

    def tom_finish(self):
        """Fused multi-GPU mode: peers store into this rank's block, so it is complete only when every
        rank's kernel has finished -- a one-element all-reduce on the stream orders exactly that."""
        if self.fused and self.use_dist:
            import torch.distributed as dist

            if not hasattr(self, "_flag"):
                self._flag = torch.zeros(1, dtype=torch.float32, device=self.dev)
            dist.all_reduce(self._flag, group=self.group)

    def tom(self, TOMType: str = "signed", TOMDenom: str = "min", out_mode: int = L.OUT_TOM,
            A_rows: torch.Tensor | None = None) -> torch.Tensor:
        """This rank's row block of the TOM (wgcna.py:1141-1157) from its adjacency row block."""
        self.tom_prepare(A_rows)
        self.tom_exchange()
        T = self.tom_compute(TOMType, TOMDenom, out_mode)
        self.tom_finish()
        return T


def network_blocks(datExpr, powerVector=None, networkType="unsigned", TOMType="signed", TOMDenom="min",
                   RsquaredCut=0.9, MeanCut=100, nBreaks=10, power=None, device=None, precision="auto",
                   out_mode: int = L.OUT_TOM, return_adjacency: bool = False):
    """Sharded host entry point: host expression matrix in, host results
    out, everything in between resident on this process's GPU.  The chain is the one ``findModules`` runs
    (wgcna.py:278 -> :310 -> :320): ``pickSoftThreshold`` (skipped when ``power`` is given), ``adjacency``
    at the selected power, ``TOMsimilarity``.

    Under torchrun (an initialised process group) every rank passes the same ``datExpr`` and gets back the
    rows ``[row0, row0 + nrows)`` it owns; per rank the PCIe traffic is the m x n expression matrix up and
    its n^2 / G block down, instead of the three n^2 transfers of the reference-shaped calls.

    Returns a dict: ``power``, ``sft`` (the P x 7 table, or None), ``rows`` = (row0, nrows), ``TOM``
    (nrows x n ndarray in a page-locked buffer; ``1 - TOM`` etc. per ``out_mode``) and, on request,
    ``adjacency`` (nrows x n)."""
    import pandas as pd

    X = L.as_f64_matrix(datExpr)
    lib = L.load()
    dev = L.default_device() if device is None else int(device)
    torch.cuda.set_device(dev)
    m, n = X.shape
    Xd = torch.empty((m, n), dtype=torch.float64, device=torch.device("cuda", dev))
    st = torch.cuda.current_stream(Xd.device).cuda_stream
    L.check(lib.b200w_dev_upload(Xd.data_ptr(), L.hptr(X), X.nbytes, dev, st))
    net = DeviceNetwork(Xd, device=dev, networkType=networkType, precision=precision)
    sft = None
    if power is None:
        pv = np.sort(np.asarray(list(range(1, 11)) + list(range(1, 21, 2)) if powerVector is None else powerVector))
        power, cols = net.sft(pv, nBreaks=nBreaks, RsquaredCut=RsquaredCut, MeanCut=MeanCut)
        names = ["Power", "SFT.R.sq", "slope", "truncated R.sq", "mean(k)", "median(k)", "max(k)"]
        sft = pd.DataFrame({"Power": pv, **{c: cols[c] for c in names[1:]}}, columns=names).astype(object)
    A = net.adjacency(float(power))
    res = {"power": power, "sft": sft, "rows": (net.row0, net.nrows)}
    if return_adjacency:
        a_host = L.result_empty((net.nrows, n))
        if net.nrows:
            L.check(lib.b200w_dev_download(L.hptr(a_host), A.data_ptr(), a_host.nbytes, dev, st))
        res["adjacency"] = a_host
    T = net.tom(TOMType=TOMType, TOMDenom=TOMDenom, out_mode=out_mode)
    t_host = L.result_empty((net.nrows, n))
    if net.nrows:
        L.check(lib.b200w_dev_download(L.hptr(t_host), T.data_ptr(), t_host.nbytes, dev, st))
    res["TOM"] = t_host
    return res
