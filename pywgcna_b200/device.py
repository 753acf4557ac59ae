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


class SftTable:
    """The six statistic columns of the reference's soft-threshold table (wgcna.py:924-932), evaluated on
    first access from the per-power statistics the device computed (``wgcna.fit_table_from_stats``).  Also
    carries what the device decided: ``device_fit`` ([P, 2]: R^2, exact mean) and ``device_power``; and
    ``host_power`` = the selection of wgcna.py:945-953 redone on the host table (the tests compare them)."""

    COLUMNS = ("SFT.R.sq", "slope", "truncated R.sq", "mean(k)", "median(k)", "max(k)")

    def __init__(self, event, host_bufs, powers, n, nBreaks, RsquaredCut, MeanCut, owner=None, generation=0):
        self._event, self._bufs = event, host_bufs
        self._owner, self._generation = owner, generation
        self.powers, self.n, self.nBreaks = powers, n, nBreaks
        self.RsquaredCut, self.MeanCut = RsquaredCut, MeanCut
        self._cols = None

    def _evaluate(self):
        if self._cols is None:
            from . import wgcna as W

            if self._owner is not None and getattr(self._owner, "_sft_generation", 0) != self._generation:
                raise RuntimeError("SftTable: a later sft() call of the same DeviceNetwork has reused the staging buffers; "
                                   "read a table before asking for the next one")
            self._event.synchronize()
            h_stats, h_buck, h_fit, h_power, _ = self._bufs
            self._device_fit = h_fit.numpy().copy()
            self._device_power = float(h_power[0])
            vals = W.fit_table_from_stats(h_stats.numpy(), h_buck.numpy(), self.n, self.nBreaks)
            self._cols = dict(zip(self.COLUMNS, vals))
            r2, mean = self._cols["SFT.R.sq"], self._cols["mean(k)"]
            with np.errstate(invalid="ignore"):
                ind = (r2 > self.RsquaredCut) & (mean <= self.MeanCut)  # wgcna.py:945
            self._host_power = self.powers[ind].min() if ind.any() else self.powers[np.argsort(r2)[-1]]
        return self._cols

    @property
    def device_fit(self):
        self._evaluate()
        return self._device_fit

    @property
    def device_power(self):
        self._evaluate()
        return self._device_power

    @property
    def host_power(self):
        self._evaluate()
        return self._host_power

    def __getitem__(self, name):
        return self._evaluate()[name]

    def keys(self):
        return self.COLUMNS

    def __iter__(self):
        return iter(self.COLUMNS)

    def __len__(self):
        return len(self.COLUMNS)


class DeviceNetwork:
    """Expression matrix resident in HBM; sweep / adjacency / TOM without leaving the device."""

    def __init__(self, X, device: int = 0, networkType: str = "unsigned", group=None, world: int | None = None,
                 rank: int | None = None, precision: str = "auto", use_dist: bool | None = None,
                 keep_similarity: bool | None = None, gene_major: bool = False):
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
        # gene_major: X holds the expression matrix genes x samples (n x m), see b200w_dev_standardize_gm
        self.gene_major = bool(gene_major)
        self.m, self.n = self.X.shape[::-1] if self.gene_major else self.X.shape
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
        # operand exchange of the fused mode: "pull" = copy engines fetch the peers' operand rows out of their
        # HBM, peer by peer, while the contraction already runs on the tiles whose operands are there (gated per
        # peer); "gather" = NCCL all-gathers before the kernel starts
        self.pull = (self.fused and bool(self.use_dist)
                     and os.environ.get("B200WGCNA_EXCHANGE", "pull") == "pull")
        self._A_raw = None
        self._peer_A_ptrs = None
        self._op_raw = None
        self._peer_op_ptrs = None
        self._epoch = 0
        self._gate = None
        self.A = None
        self._sim_in_A = False  # self.A holds the similarity the last sweep left (see sweep / adjacency)
        # sweep() leaves this rank's rows of the similarity in the adjacency buffer (overwriting a block an
        # earlier adjacency() returned) and the next adjacency() finishes them elementwise
        if keep_similarity is None:
            keep_similarity = os.environ.get("B200WGCNA_KEEP_SIMILARITY", "1") != "0"
        self.keep_similarity = bool(keep_similarity)
        # sweep of the sharded run: "symmetric" = block pairs dealt out over the ranks, every correlation tile formed
        # once and its similarity delivered to both owners' blocks (b200w_dev_sweep_dist: with one peer stored over
        # NVLink from inside the kernel; with more, staged in a local rectangle per peer that the copy engines push
        # while the kernel does the rank's own pairs -- plain stores to three peers made the kernel 4.5x slower,
        # profiles/r02_sweepdist_probe_n4*.json); "full" = every rank sweeps its whole column block.
        # C3: 2 GPUs 32.9 -> 21.1 ms, 4 GPUs 16.7 -> 12.7 ms.
        default_mode = "symmetric"
        self.sym_sweep = (self.world > 1 and bool(self.use_dist) and self.keep_similarity
                          and os.environ.get("B200WGCNA_SWEEP_DIST", default_mode) == "symmetric")
        self._sync_token = None
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
            if self._op_raw:
                self.lib.b200w_dev_free(self._op_raw, self.device)
            if self._A_raw:
                self.lib.b200w_dev_free(self._A_raw, self.device)
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def standardize(self):
        entry = self.lib.b200w_dev_standardize_gm if self.gene_major else self.lib.b200w_dev_standardize
        L.check(entry(self.X.data_ptr(), self.m, self.n, self.Z.data_ptr(), self.bad.data_ptr(), self.device,
                      self._stream()))
        self._standardized = True
        self._sim_in_A = False

    def _alloc_A(self):
        """This rank's rows of the similarity / adjacency.  With the symmetric sharded sweep the block lives in a
        cudaMalloc buffer of ``per`` rows that the peers map through CUDA IPC (they store into it)."""
        if self.A is not None:
            return
        if self.sym_sweep:
            p = C.c_void_p()
            L.check(self.lib.b200w_dev_malloc(self.per * self.n * 8, self.device, C.byref(p)))
            self._A_raw = p.value
            self.A = torch.as_tensor(_DevView(p.value, (self.per, self.n)), device=self.dev)
        else:
            self.A = torch.empty((max(self.nrows, 1), self.n), dtype=torch.float64, device=self.dev)

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
        if self.sym_sweep and P <= 20 and bool(np.all((steps == 1.0) | (steps == 2.0))):
            import torch.distributed as dist

            self._alloc_A()
            if self._peer_ptrs is None:
                self.connect_peers()
            kpart = torch.empty((P, self.n), dtype=torch.float64, device=self.dev)
            # the peers' kernels store into this rank's block: nobody starts before every rank is through with what
            # it still reads out of the block (the TOM epilogue of the previous pass) -- a stream-ordered collective
            if self._sync_token is None:
                self._sync_token = torch.zeros(1, dtype=torch.float32, device=self.dev)
            dist.all_reduce(self._sync_token, group=self.group)
            rc = self.lib.b200w_dev_sweep_dist(self.Z.data_ptr(), self.bad.data_ptr(), self.m, self.n,
                                               steps.ctypes.data, P, self.net, self.world, self.rank, self.per,
                                               self._peer_A_ptrs, kpart.data_ptr(), self.device, self._stream())
            if rc != L.E_UNSUPPORTED:  # (the FP64 correlation engine has no sharded schedule: the column-block path)
                L.check(rc)
                # the sum of the ranks' partial connectivities; once it has passed, every rank's kernel has finished,
                # so the similarity rows the peers stored into this rank's block are complete as well
                dist.all_reduce(kpart, group=self.group)
                self._sim_in_A = True
                return kpart
        col0, ncols = (0, self.n) if self.world == 1 else (self.row0, self.nrows)
        kloc = torch.empty((P, max(ncols, 1)), dtype=torch.float64, device=self.dev)
        if ncols > 0 and self.keep_similarity:
            # the sweep also writes this rank's rows of the transformed similarity into the adjacency buffer,
            # and adjacency() of the same data finishes them with an elementwise S ** power instead of forming
            # the correlation a second time as the reference does (wgcna.py:882 and :1097)
            self._alloc_A()
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
        """Device-resident ``pickSoftThreshold`` (wgcna.py:873-953): sweep, then histograms, medians, exact
        mantissa sums, the regression R^2 and the power selection (:945-953) on the device
        (``b200w_dev_sft_select``).  Nothing here waits for the GPU: returns ``(power, table)`` where ``power``
        is a 0-dim float64 DEVICE tensor that ``adjacency()`` reads in place, and ``table`` is a lazy mapping
        of the six columns of the reference's table -- its first access waits for the asynchronous copy of the
        per-power statistics and runs the 2 x P ten-point regressions on the host (float64, the reference's
        arithmetic), typically while the GPU is busy with the adjacency and the TOM."""
        pv = np.sort(np.asarray(powers))
        k = self.sweep(pv)  # [P, n]
        P = len(pv)
        key = (P, nBreaks)
        if getattr(self, "_sft_key", None) != key:
            E = self.lib.b200w_sft_exp_buckets()
            f64, i64 = torch.float64, torch.int64
            self._sft_dev = (torch.empty((P, 4 + 2 * nBreaks), dtype=f64, device=self.dev),
                             torch.empty((P, E, 2), dtype=i64, device=self.dev),
                             torch.empty((P, 2), dtype=f64, device=self.dev),
                             torch.empty(2, dtype=f64, device=self.dev),
                             torch.empty(P, dtype=f64, device=self.dev))
            self._sft_host = (torch.empty((P, 4 + 2 * nBreaks), dtype=f64).pin_memory(),
                              torch.empty((P, E, 2), dtype=i64).pin_memory(),
                              torch.empty((P, 2), dtype=f64).pin_memory(), torch.empty(2, dtype=f64).pin_memory(),
                              torch.empty(P, dtype=f64).pin_memory())
            self._sft_key = key
            self._sft_powers = None
        d_stats, d_buck, d_fit, d_power, d_powers = self._sft_dev
        h_stats, h_buck, h_fit, h_power, h_powers = self._sft_host
        if self._sft_powers is None or not np.array_equal(self._sft_powers, pv):
            h_powers.copy_(torch.from_numpy(pv.astype(np.float64)))
            d_powers.copy_(h_powers, non_blocking=True)
            self._sft_powers = pv.copy()
        L.check(self.lib.b200w_dev_sft_select(k.data_ptr(), P, self.n, k.stride(0), nBreaks, d_powers.data_ptr(),
                                              float(RsquaredCut), float(MeanCut), d_stats.data_ptr(),
                                              d_buck.data_ptr(), d_fit.data_ptr(), d_power.data_ptr(),
                                              self.device, self._stream()))
        h_stats.copy_(d_stats, non_blocking=True)
        h_buck.copy_(d_buck, non_blocking=True)
        h_fit.copy_(d_fit, non_blocking=True)
        h_power.copy_(d_power, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self._sft_generation = getattr(self, "_sft_generation", 0) + 1
        return d_power[0], SftTable(ev, self._sft_host, pv, self.n, nBreaks, RsquaredCut, MeanCut, owner=self,
                                    generation=self._sft_generation)

    def adjacency(self, power) -> torch.Tensor:
        """This rank's row block of the adjacency (wgcna.py:1097-1111), kept in HBM.  ``power`` is a number
        or the device scalar ``sft()`` returned (then the kernels read it from HBM: no host round trip)."""
        if not self._standardized:
            self.standardize()
        self._alloc_A()
        on_dev = isinstance(power, torch.Tensor) and power.is_cuda
        if on_dev:
            assert power.dtype == torch.float64 and power.device == self.dev
        if self._sim_in_A:
            self._sim_in_A = False  # consumed: the buffer turns into the adjacency
            if self.nrows == 0:
                return self.A[:0]
            if on_dev:
                L.check(self.lib.b200w_dev_adjacency_from_similarity_dp(self.A.data_ptr(), self.nrows, self.n,
                                                                        power.data_ptr(), self.device,
                                                                        self._stream()))
            else:
                L.check(self.lib.b200w_dev_adjacency_from_similarity(self.A.data_ptr(), self.nrows, self.n,
                                                                     float(power), self.device, self._stream()))
            return self.A[:self.nrows]
        if self.nrows > 0:
            if on_dev:
                L.check(self.lib.b200w_dev_adjacency_dp(self.Z.data_ptr(), self.bad.data_ptr(), self.m, self.n,
                                                        self.net, power.data_ptr(), self.row0, self.nrows,
                                                        self.A.data_ptr(), self.device, self._stream()))
            else:
                L.check(self.lib.b200w_dev_adjacency(self.Z.data_ptr(), self.bad.data_ptr(), self.m, self.n,
                                                     self.net, float(power), self.row0, self.nrows,
                                                     self.A.data_ptr(), self.device, self._stream()))
        return self.A[:self.nrows]

    # -- TOM -----------------------------------------------------------------------------------
    def _operand(self):
        if self._op is None and self.pull:
            # planes, scale and k in one cudaMalloc allocation that the peers map through CUDA IPC
            nbytes = self.lib.b200w_tom_shared_bytes(self.op_rows, self.n)
            off = self.lib.b200w_tom_shared_scale_offset(self.op_rows, self.n)
            p = C.c_void_p()
            L.check(self.lib.b200w_dev_malloc(nbytes, self.device, C.byref(p)))
            self._op_raw = p.value
            whole = torch.as_tensor(_DevView(p.value, (nbytes,), "|u1"), device=self.dev)
            whole.zero_()
            self._op = whole[:self.planes * self.op_rows * self.ldn].view(self.planes, self.op_rows, self.ldn)
            self._scale = whole[off:off + self.op_rows * 8].view(torch.float64)
            self._k = whole[off + self.op_rows * 8:off + self.op_rows * 16].view(torch.float64)
            self._plane_flags = torch.zeros(8, dtype=torch.int32, device=self.dev)
            self._copy_stream = torch.cuda.Stream(device=self.dev)
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

    def ipc_handle(self, raw=None) -> bytes:
        self._out()
        h = C.create_string_buffer(64)
        L.check(self.lib.b200w_ipc_export(self._T_raw if raw is None else raw, h))
        return h.raw

    def connect_peers(self, ptrs=None):
        """Table of every rank's TOM block (and, in pull mode, operand allocation) as seen from this GPU.
        ``ptrs`` (tests: ranks emulated in one process) are used as they are; otherwise CUDA IPC handles are
        exchanged over the process group -- a collective: every rank calls it, also one without rows."""
        self._out()
        if ptrs is None:
            import torch.distributed as dist

            self._operand()
            if self.sym_sweep:
                self._alloc_A()
            handles = (self.ipc_handle() + (self.ipc_handle(self._op_raw) if self.pull else bytes(64))
                       + (self.ipc_handle(self._A_raw) if self.sym_sweep else bytes(64)))
            mine = torch.frombuffer(bytearray(handles), dtype=torch.uint8).to(self.dev)
            allh = torch.zeros((self.world, 192), dtype=torch.uint8, device=self.dev)
            dist.all_gather_into_tensor(allh, mine, group=self.group)
            allh = allh.cpu().numpy()
            ptrs, op_ptrs, a_ptrs = [], [], []
            for r in range(self.world):
                if r == self.rank:
                    ptrs.append(self._T_raw)
                    op_ptrs.append(self._op_raw or 0)
                    a_ptrs.append(self._A_raw or 0)
                    continue
                p = C.c_void_p()
                L.check(self.lib.b200w_ipc_open(allh[r, :64].tobytes(), self.device, C.byref(p)))
                self._peer_open.append(p.value)
                ptrs.append(p.value)
                if self.pull:
                    q = C.c_void_p()
                    L.check(self.lib.b200w_ipc_open(allh[r, 64:128].tobytes(), self.device, C.byref(q)))
                    self._peer_open.append(q.value)
                    op_ptrs.append(q.value)
                if self.sym_sweep:
                    q = C.c_void_p()
                    L.check(self.lib.b200w_ipc_open(allh[r, 128:].tobytes(), self.device, C.byref(q)))
                    self._peer_open.append(q.value)
                    a_ptrs.append(q.value)
            if self.pull:
                self._peer_op_ptrs = (C.c_void_p * self.world)(*[int(p) for p in op_ptrs])
            if self.sym_sweep:
                self._peer_A_ptrs = (C.c_void_p * self.world)(*[int(p) for p in a_ptrs])
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

    def _sync_flag(self):
        if not hasattr(self, "_flag"):
            self._flag = torch.zeros(1, dtype=torch.float32, device=self.dev)
        return self._flag

    def tom_exchange(self):
        if self.world == 1:
            return
        if self.pull:
            import torch.distributed as dist

            if self._peer_ptrs is None:
                self.connect_peers()
            # every rank's rows of the operand are final once this tiny all-reduce has passed (stream order);
            # from then on the copy engines pull the peers' rows while the contraction below already runs
            dist.all_reduce(self._sync_flag(), group=self.group)
            main = torch.cuda.current_stream(self.dev)
            ev = torch.cuda.Event()
            ev.record(main)
            self._copy_stream.wait_event(ev)
            self._epoch = (self._epoch + 1) & 0x7fffffff
            L.check(self.lib.b200w_dev_tom_pull_planes(self._op_raw, self._peer_op_ptrs, self.world, self.rank,
                                                       self.per, self.op_rows, self.n,
                                                       self._plane_flags.data_ptr(), self._epoch, self.device,
                                                       self._copy_stream.cuda_stream))
            self._gate = self._epoch
            return
        for p in range(self.planes):
            self._all_gather_rows(self._op[p])
        self._all_gather_rows(self._k)
        self._all_gather_rows(self._scale)

    def tom_compute(self, TOMType: str = "signed", TOMDenom: str = "min", out_mode: int = L.OUT_TOM):
        tt, td = L.TOM_TYPES.index(TOMType), L.TOM_DENOMS.index(TOMDenom)
        T = self._out()
        if self.fused and self._peer_ptrs is None:
            self.connect_peers()  # a collective: every rank takes part, also one without rows
        gate, self._gate = self._gate, None
        if self.nrows == 0:
            if gate is not None:  # nothing to compute, but the pulls must not outlive this step
                torch.cuda.current_stream(self.dev).wait_stream(self._copy_stream)
            return T[:0]
        if self.fused and gate is not None:
            L.check(self.lib.b200w_dev_tom_compute_dist_gated(
                self._A_rows.data_ptr(), self._op.data_ptr(), self.op_rows, self._scale.data_ptr(),
                self._k.data_ptr(), self.n, self.world, self.rank, self.per, tt, td, out_mode, self._peer_ptrs,
                self._plane_flags.data_ptr(), gate, self.device, self._stream()))
        elif self.fused:
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

            dist.all_reduce(self._sync_flag(), group=self.group)

    def tom(self, TOMType: str = "signed", TOMDenom: str = "min", out_mode: int = L.OUT_TOM,
            A_rows: torch.Tensor | None = None) -> torch.Tensor:
        """This rank's row block of the TOM (wgcna.py:1141-1157) from its adjacency row block."""
        self.tom_prepare(A_rows)
        self.tom_exchange()
        T = self.tom_compute(TOMType, TOMDenom, out_mode)
        self.tom_finish()
        return T


def _gene_tree(self, TOMType: str = "signed", TOMDenom: str = "min") -> np.ndarray:
    """The gene tree of findModules (wgcna.py:320-328) without leaving the device: round(1 - TOM, 8) from the TOM
    epilogue in square form, then the average-linkage kernel on it (the matrix is consumed), then scipy's sort
    and labels on the n - 1 merges.  Single GPU (the walk is sequential); returns the scipy linkage matrix."""
    if self.world != 1:
        raise RuntimeError("gene_tree: gather the TOM row blocks on one GPU first")
    D = self.tom(TOMType=TOMType, TOMDenom=TOMDenom, out_mode=L.OUT_DISS_R8)
    Zraw = torch.empty((self.n - 1, 4), dtype=torch.float64, device=self.dev)
    L.check(self.lib.b200w_dev_linkage_average(D.data_ptr(), self.n, Zraw.data_ptr(), self.device, self._stream()))
    Z = np.ascontiguousarray(Zraw.cpu().numpy())
    L.check(self.lib.b200w_linkage_finish(L.hptr(Z), self.n))
    return Z


DeviceNetwork.gene_tree = _gene_tree

_NET_CACHE = {}


def release_workspace():
    """Drop the device workspace ``network_blocks`` keeps between calls."""
    _NET_CACHE.clear()
    torch.cuda.empty_cache()


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

    X, gene_major = L.as_f64_expression(datExpr)  # no host-side transposition for an F-ordered matrix
    lib = L.load()
    dev = L.default_device() if device is None else int(device)
    torch.cuda.set_device(dev)
    m, n = X.shape[::-1] if gene_major else X.shape
    # the device workspace (and, under torchrun, the peer mappings: IPC handle exchange + cudaIpcOpenMemHandle
    # cost hundreds of ms) is kept between calls of the same shape, like the DevCache of the C entry points
    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
    key = (dev, m, n, networkType, precision, torch.distributed.get_world_size() if dist_on else 1, gene_major)
    net = _NET_CACHE.get(key)
    if net is None:
        _NET_CACHE.clear()  # one workspace at a time: these are tens of GB
        net = DeviceNetwork(torch.empty(X.shape, dtype=torch.float64, device=torch.device("cuda", dev)), device=dev,
                            networkType=networkType, precision=precision, gene_major=gene_major)
        _NET_CACHE[key] = net
    st = torch.cuda.current_stream(net.dev).cuda_stream
    L.check(lib.b200w_dev_upload(net.X.data_ptr(), L.hptr(X), X.nbytes, dev, st))
    net._standardized = False
    net._sim_in_A = False
    sft = None
    picked = power is None
    if picked:
        pv = np.sort(np.asarray(list(range(1, 11)) + list(range(1, 21, 2)) if powerVector is None else powerVector))
        power, cols = net.sft(pv, nBreaks=nBreaks, RsquaredCut=RsquaredCut, MeanCut=MeanCut)
    A = net.adjacency(power)  # a device scalar after sft(): the chain below is enqueued without waiting
    res = {"rows": (net.row0, net.nrows)}
    if return_adjacency:
        a_host = L.result_empty((net.nrows, n))
        if net.nrows:
            L.check(lib.b200w_dev_download(L.hptr(a_host), A.data_ptr(), a_host.nbytes, dev, st))
        res["adjacency"] = a_host
    t_host = L.result_empty((net.nrows, n))
    overlapped = net.world == 1 and net.elt == 1 and out_mode in (L.OUT_TOM, L.OUT_DISS, L.OUT_DISS_R8)
    if overlapped:
        # one GPU: the result leaves the device row block by row block while the contraction is still running
        net.tom_prepare()
        tt, td = L.TOM_TYPES.index(TOMType), L.TOM_DENOMS.index(TOMDenom)
        T = net._out()
    else:
        T = net.tom(TOMType=TOMType, TOMDenom=TOMDenom, out_mode=out_mode)
    if picked:
        # the host half of the soft-threshold table (2 x P ten-point regressions) runs while the GPU works
        names = ["Power", "SFT.R.sq", "slope", "truncated R.sq", "mean(k)", "median(k)", "max(k)"]
        sft = pd.DataFrame({"Power": pv, **{c: cols[c] for c in names[1:]}}, columns=names).astype(object)
        power = type(pv[0])(cols.device_power)
    res["power"], res["sft"] = power, sft
    if overlapped:
        L.check(lib.b200w_dev_tom_compute_to_host(net._A_rows.data_ptr(), net._op.data_ptr(), net.op_rows,
                                                  net._scale.data_ptr(), net._k.data_ptr(), n, tt, td, out_mode,
                                                  net.prec, T.data_ptr(), L.hptr(t_host), dev, st))
    elif net.nrows:
        L.check(lib.b200w_dev_download(L.hptr(t_host), T.data_ptr(), t_host.nbytes, dev, st))
    res["TOM"] = t_host
    return res
