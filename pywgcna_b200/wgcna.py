"""Host-side mirror of the reference interface for the network-construction hot path.

Same names, argument meaning, return types and error behaviour as the static methods of
``PyWGCNA.wgcna.WGCNA`` (reference file ``PyWGCNA/wgcna.py``):

===========================  =======================  ==========================================
here                         reference                device work (include/b200wgcna.h)
===========================  =======================  ==========================================
``pickSoftThreshold``        ``wgcna.py:788-955``     ``b200w_connectivity_sweep`` (+ host fit)
``scaleFreeFitIndex``        ``wgcna.py:1008-1044``   host, 10-point OLS (float64)
``adjacency``                ``wgcna.py:1047-1111``   ``b200w_adjacency``
``checkAdjMat``              ``wgcna.py:1114-1138``   ``b200w_check_adjacency``
``TOMsimilarity``            ``wgcna.py:1160-1193``   ``b200w_tom``
``TomSimilarityFromAdj``     ``wgcna.py:1141-1157``   ``b200w_tom``
``softConnectivity``         ``wgcna.py:3342-3400``   ``b200w_connectivity_sweep``
``intramodularConnectivity`` ``wgcna.py:3403-3450``   ``b200w_intramodular``
===========================  =======================  ==========================================

``install()`` rebinds the three methods ``WGCNA.findModules`` looks up by class name
(``wgcna.py:278, :310, :320``) so the reference workflow runs on the GPU unchanged.

All arithmetic on the n x n objects happens in CUDA; this module never falls back to numpy for it.
"""
from __future__ import annotations

import math
import os
import sys
from fractions import Fraction

import numpy as np
import pandas as pd

from . import _lib as L

networkTypes = ["unsigned", "signed", "signed hybrid"]  # wgcna.py:53
adjacencyTypes = ["unsigned", "signed", "signed hybrid"]  # wgcna.py:54
TOMTypes = ["unsigned", "signed"]  # wgcna.py:55
TOMDenoms = ["min", "mean"]  # wgcna.py:56

# the reference's ANSI colours (wgcna.py:58-66) -- kept so logs look the same
OKCYAN = "\033[96m"
OKGREEN = "\033[92m"
WARNING = "\033[93m"
ENDC = "\033[0m"

#: set to False to silence the progress prints the reference emits
VERBOSE = os.environ.get("B200WGCNA_QUIET", "") == ""
#: arithmetic of the TOM contraction: "auto" | "f64" | "i8" | "i8-fast" (see B200W_PREC_* in the header)
PRECISION = os.environ.get("B200WGCNA_PRECISION", "auto")

_PREC = {"auto": L.PREC_AUTO, "f64": L.PREC_F64, "i8": L.PREC_I8, "i8-fast": L.PREC_I8_FAST}


def set_correlation_precision(precision: str = "auto"):
    """Arithmetic of the correlation product behind ``pickSoftThreshold`` / ``adjacency`` (process wide):
    ``"i8"`` (= ``"auto"``) the tcgen05 int8 digit-plane engine, ``"f64"`` the FP64 DMMA engine."""
    L.check(L.load().b200w_set_correlation_precision({"auto": L.PREC_AUTO, "f64": L.PREC_F64, "i8": L.PREC_I8}[precision]))


def _say(*a, **k):
    if VERBOSE:
        print(*a, **k)


def _device(device):
    return L.default_device() if device is None else int(device)


# ----------------------------------------------------------------------------------------------
# scale-free topology fit (host; 10 points per power)
# ----------------------------------------------------------------------------------------------
def _cut_edges(mn: float, mx: float, nBreaks: int) -> np.ndarray:
    """Bin edges of ``pd.cut(k, nBreaks)`` (pandas 2.2 ``_nbins_to_bins``, right-closed bins)."""
    if mn == mx:
        lo = mn - (0.001 * abs(mn) if mn != 0 else 0.001)
        hi = mx + (0.001 * abs(mx) if mx != 0 else 0.001)
        return np.linspace(lo, hi, nBreaks + 1, endpoint=True)
    edges = np.linspace(mn, mx, nBreaks + 1, endpoint=True)
    edges[0] -= (mx - mn) * 0.001
    return edges


def _ols(y: np.ndarray, X: np.ndarray):
    """Least squares as statsmodels ``OLS.fit()`` does it (pinv), returning
    (params, R^2, adjusted R^2) with the centred total sum of squares."""
    beta = np.linalg.pinv(X) @ y
    resid = y - X @ beta
    ssr = float(resid @ resid)
    yc = y - y.mean()
    tss = float(yc @ yc)
    with np.errstate(all="ignore"):
        rsq = 1.0 - ssr / tss
        rank = int(np.linalg.matrix_rank(X))
        nobs = len(y)
        rsq_adj = 1.0 - (nobs - 1) / (nobs - rank) * (1.0 - rsq)
    return beta, rsq, rsq_adj


def _scale_free_fit(k: np.ndarray, nBreaks: int):
    """(R^2, slope, truncated-exponential adjusted R^2) of wgcna.py:1017-1043, with the
    semantics of the pandas the reference pins (2.2.2: every bin kept, empty bins -> midpoint)."""
    k = np.asarray(k, dtype=np.float64)
    mn, mx = float(k.min()), float(k.max())
    edges = _cut_edges(mn, mx, nBreaks)
    ids = np.searchsorted(edges, k, side="left") - 1  # (e_i, e_i+1]
    ok = (ids >= 0) & (ids < nBreaks)
    cnt = np.bincount(ids[ok], minlength=nBreaks).astype(np.float64)
    tot = np.bincount(ids[ok], weights=k[ok], minlength=nBreaks)
    with np.errstate(all="ignore"):
        dk = tot / cnt  # NaN where a bin is empty            (:1019)
        p_dk = cnt / len(k)  # (:1022)
        mids = np.linspace(mn, mx, nBreaks + 1)  # (:1025)
        mids = 0.5 * (mids[1:] + mids[:-1])  # (:1027)
        empty = np.isnan(dk)
        dk[empty] = mids[empty]  # (:1029-1030)
        zero = dk == 0
        dk[zero] = mids[zero]  # (:1031-1032)
        log_dk = np.log10(dk)  # (:1035)
        log_p = np.log10(p_dk + 1e-09)  # (:1036)
        trunc = np.power(10, log_dk)  # (:1037)
        ones = np.ones(nBreaks)
        b1, r1, _ = _ols(log_p, np.column_stack([ones, log_dk]))  # (:1039)
        _, _, r2adj = _ols(log_p, np.column_stack([ones, log_dk, trunc]))  # (:1040)
    return r1, float(b1[1]), r2adj


def scaleFreeFitIndex(k, nBreaks=10):
    """``WGCNA.scaleFreeFitIndex`` (wgcna.py:1008-1044)."""
    r1, slope, r2adj = _scale_free_fit(k, nBreaks)
    return pd.DataFrame({"Rsquared.SFT": [r1], "slope.SFT": [slope],
                         "truncatedExponentialAdjRsquared": [r2adj]})


def _exact_mean(k: np.ndarray) -> np.float64:
    """``statistics.mean`` (wgcna.py:930): the exactly rounded mean.  fsum gives the correctly
    rounded sum, a second fsum its rounding error; their exact rational sum / n is rounded once."""
    vals = k.tolist()
    s = math.fsum(vals)
    if not math.isfinite(s):
        return np.float64(s / len(vals))
    r = math.fsum(vals + [-s])
    return np.float64(float((Fraction(s) + Fraction(r)) / len(vals)))


def _exact_mean_sorted(ks: np.ndarray) -> np.float64:
    """Same value as ``_exact_mean`` (= ``statistics.mean``), vectorised: every double is
    mantissa * 2^exponent with a 53-bit integer mantissa; mantissas are summed exactly per exponent
    (split in 27-bit halves so float64 bincount sums stay below 2^53), the few per-exponent totals are
    combined as Python integers and the exact rational sum / n is rounded once."""
    n = ks.shape[0]
    if n == 0 or not np.isfinite(ks[0]) or not np.isfinite(ks[-1]):
        return _exact_mean(ks)
    m, e = np.frexp(ks)
    M = (m * 9007199254740992.0).astype(np.int64)  # m * 2^53, exact
    nz = M != 0
    if not nz.any():
        return np.float64(0.0)
    emin = int(e[nz].min())
    sh = np.where(nz, e - emin, 0).astype(np.int64)
    hi = M >> 27
    lo = M - (hi << 27)
    W = int(sh.max()) + 1
    slo = np.bincount(sh, weights=lo.astype(np.float64), minlength=W)
    shi = np.bincount(sh, weights=hi.astype(np.float64), minlength=W)
    total = 0
    for s in np.nonzero((slo != 0) | (shi != 0))[0].tolist():
        total += ((int(shi[s]) << 27) + int(slo[s])) << s
    ex = emin - 53
    frac = Fraction(total * (1 << ex), n) if ex >= 0 else Fraction(total, n * (1 << -ex))
    return np.float64(float(frac))


def _ols_batch(y: np.ndarray, X: np.ndarray):
    """``_ols`` for a stack of problems: y [B, N], X [B, N, q] -> params [B, q], R^2 [B], adj R^2 [B]
    (pinv and rank from one batched SVD, like statsmodels' pinv solve + matrix_rank)."""
    with np.errstate(all="ignore"):
        pinv = np.linalg.pinv(X)
        beta = np.einsum("bqn,bn->bq", pinv, y)
        resid = y - np.einsum("bnq,bq->bn", X, beta)
        ssr = np.einsum("bn,bn->b", resid, resid)
        yc = y - y.mean(axis=1, keepdims=True)
        tss = np.einsum("bn,bn->b", yc, yc)
        rsq = 1.0 - ssr / tss
        rank = np.linalg.matrix_rank(X)
        nobs = y.shape[1]
        rsq_adj = 1.0 - (nobs - 1) / (nobs - rank) * (1.0 - rsq)
    return beta, rsq, rsq_adj


def fit_table(K: np.ndarray, nBreaks: int = 10):
    """All per-power statistics of wgcna.py:924-932 for k[P, n] at once: returns the arrays
    (SFT.R.sq, slope, truncated R.sq, mean(k), median(k), max(k)).  Same arithmetic as
    ``scaleFreeFitIndex`` / ``statistics.mean`` / ``statistics.median``; one sort per power serves the
    median, the histogram (bin boundaries by binary search of the 11 edges in the sorted k) and the
    exact mean, and the 2 x P least-squares problems are solved as two batched SVDs."""
    K = np.asarray(K, dtype=np.float64)
    P, n = K.shape
    dk = np.empty((P, nBreaks))
    p_dk = np.empty((P, nBreaks))
    mean = np.empty(P)
    med = np.empty(P)
    mx = np.empty(P)
    with np.errstate(all="ignore"):
        for i in range(P):
            ks = np.sort(K[i])
            lo_, hi_ = float(ks[0]), float(ks[-1])
            mx[i] = hi_
            med[i] = ks[n // 2] if n % 2 else (ks[n // 2 - 1] + ks[n // 2]) / 2  # statistics.median
            mean[i] = _exact_mean_sorted(ks)
            edges = _cut_edges(lo_, hi_, nBreaks)
            # bin b = (edges[b], edges[b+1]]  ->  boundaries = #elements <= edge
            bnd = np.searchsorted(ks, edges, side="right")  # bnd[0] = 0 and bnd[-1] = n by construction
            cnt = np.diff(bnd).astype(np.float64)
            tot = np.zeros(nBreaks)
            ne = np.nonzero(cnt > 0)[0]
            # the non-empty bins tile the sorted array contiguously, so reduceat over their starts sums
            # exactly the members of each bin
            tot[ne] = np.add.reduceat(ks, bnd[ne])
            d = tot / cnt  # NaN where a bin is empty (:1019)
            mids = np.linspace(lo_, hi_, nBreaks + 1)  # (:1025)
            mids = 0.5 * (mids[1:] + mids[:-1])  # (:1027)
            empty = np.isnan(d)
            d[empty] = mids[empty]  # (:1029-1030)
            zero = d == 0
            d[zero] = mids[zero]  # (:1031-1032)
            dk[i] = d
            p_dk[i] = cnt / n  # (:1022)
        log_dk = np.log10(dk)  # (:1035)
        log_p = np.log10(p_dk + 1e-09)  # (:1036)
        trunc = np.power(10, log_dk)  # (:1037)
        ones = np.ones_like(log_dk)
        b1, r1, _ = _ols_batch(log_p, np.stack([ones, log_dk], axis=2))  # (:1039)
        _, _, r2adj = _ols_batch(log_p, np.stack([ones, log_dk, trunc], axis=2))  # (:1040)
    return r1, b1[:, 1], r2adj, mean, med, mx


def fit_table_from_stats(stats: np.ndarray, buckets: np.ndarray, n: int, nBreaks: int = 10):
    """``fit_table`` from the per-power statistics the device computed (``b200w_dev_sft_stats``):
    stats [P, 4 + 2 nb] (min, max, median, sum, counts, bin sums), buckets [P, E, 2] exact mantissa
    sums.  Only the two batched 10-point regressions and the assembly of the exact mean run here."""
    stats = np.asarray(stats, dtype=np.float64)
    P = stats.shape[0]
    mn, mx, med = stats[:, 0], stats[:, 1], stats[:, 2]
    cnt = stats[:, 4:4 + nBreaks]
    tot = stats[:, 4 + nBreaks:4 + 2 * nBreaks]
    mean = np.empty(P)
    bk = np.asarray(buckets).reshape(P, -1, 2)
    for p in range(P):
        nzb = np.nonzero((bk[p, :, 0] != 0) | (bk[p, :, 1] != 0))[0]
        if nzb.size == 0:
            mean[p] = 0.0 if np.isfinite(stats[p, 3]) else stats[p, 3] / n
            continue
        e0 = int(nzb[0])
        total = 0
        for e in nzb.tolist():
            total += ((int(bk[p, e, 1]) << 27) + int(bk[p, e, 0])) << (e - e0)
        ex = e0 - 1100 - 53  # value = total * 2^ex
        frac = Fraction(total * (1 << ex), n) if ex >= 0 else Fraction(total, n * (1 << -ex))
        mean[p] = float(frac)
    with np.errstate(all="ignore"):
        dk = tot / cnt  # NaN where a bin is empty (:1019)
        p_dk = cnt / n  # (:1022)
        step = (mx - mn) / nBreaks  # np.linspace(min k, max k, nBreaks + 1) (:1025)
        grid = np.arange(nBreaks + 1)[None, :] * step[:, None] + mn[:, None]
        grid[:, -1] = mx
        mids = 0.5 * (grid[:, 1:] + grid[:, :-1])  # (:1027)
        empty = np.isnan(dk)
        dk = np.where(empty, mids, dk)  # (:1029-1030)
        dk = np.where(dk == 0, mids, dk)  # (:1031-1032)
        log_dk = np.log10(dk)  # (:1035)
        log_p = np.log10(p_dk + 1e-09)  # (:1036)
        trunc = np.power(10, log_dk)  # (:1037)
        ones = np.ones_like(log_dk)
        b1, r1, _ = _ols_batch(log_p, np.stack([ones, log_dk], axis=2))  # (:1039)
        _, _, r2adj = _ols_batch(log_p, np.stack([ones, log_dk, trunc], axis=2))  # (:1040)
    return r1, b1[:, 1], r2adj, mean, med.copy(), mx.copy()


def calBlockSize(matrixSize, rectangularBlocks=True, maxMemoryAllocation=None, overheadFactor=3):
    """``WGCNA.calBlockSize`` (wgcna.py:988-1004).  Only reported for log parity: the CUDA sweep
    tiles the n x n problem on chip and needs no host block size."""
    if maxMemoryAllocation is None:
        import resource

        maxAlloc = resource.getrlimit(resource.RLIMIT_AS)[1]
    else:
        maxAlloc = maxMemoryAllocation / 8
    maxAlloc = maxAlloc / overheadFactor
    if rectangularBlocks:
        blockSz = math.floor(maxAlloc / matrixSize)
    else:
        blockSz = math.floor(math.sqrt(maxAlloc))
    return min(matrixSize, blockSz)


# ----------------------------------------------------------------------------------------------
# pickSoftThreshold
# ----------------------------------------------------------------------------------------------
def connectivity(data, powerVector, networkType="unsigned", device=None, return_bad=False):
    """k[gene, power] of wgcna.py:856-920 (``datk``) for the ascending ``powerVector``."""
    X = L.as_f64_matrix(data)
    powers = np.sort(np.asarray(powerVector, dtype=np.float64))
    intType = networkTypes.index(networkType)
    m, n = X.shape
    steps = np.ascontiguousarray(np.diff(np.concatenate(([0.0], powers))))  # :901-903
    P = len(powers)
    kT = np.empty((P, n), dtype=np.float64)
    bad = np.empty(n, dtype=np.uint8)
    lib = L.load()
    L.check(lib.b200w_connectivity_sweep(L.hptr(X), m, n, L.hptr(steps), P, intType, L.hptr(kT),
                                         L.hptr(bad), _device(device)))
    if return_bad:
        return kT.T, bad.astype(bool)
    return kT.T


def pickSoftThreshold(data, dataIsExpr=True, weights=None, RsquaredCut=0.9, MeanCut=100, powerVector=None,
                      nBreaks=10, blockSize=None, corOptions=None, networkType="unsigned",
                      moreNetworkConcepts=False, gcInterval=None, device=None):
    """``WGCNA.pickSoftThreshold`` (wgcna.py:788-955): same arguments, same return
    ``(numpy.int64 powerEstimate, DataFrame[P x 7] of dtype object)``."""
    if powerVector is None:
        powerVector = list(range(1, 11)) + list(range(1, 21, 2))  # :823
    powerVector = np.sort(powerVector)  # :824
    networkTypes.index(networkType)  # :825 -> ValueError on an unknown string
    nGenes = data.shape[1]
    if nGenes < 3:  # :830-832
        sys.exit("The input data data contain fewer than 3 rows (nodes).\n"
                 "This would result in a trivial correlation network.")
    if not dataIsExpr:
        raise NotImplementedError("dataIsExpr=False: that branch of the reference cannot run (wgcna.py:839)")
    if weights is not None:
        raise NotImplementedError("observation weights are stored but never used by the reference "
                                  "(wgcna.py:866-871); refusing to ignore them silently")
    if moreNetworkConcepts:
        raise NotImplementedError("moreNetworkConcepts=True crashes in the reference (wgcna.py:851)")
    _say(f"{OKCYAN}pickSoftThreshold: calculating connectivity for given powers...{ENDC}")
    if blockSize is None:
        blockSize = calBlockSize(nGenes, rectangularBlocks=True, maxMemoryAllocation=2 ** 30)
        _say("will use block size ", blockSize, flush=True)

    colname1 = ["Power", "SFT.R.sq", "slope", "truncated R.sq", "mean(k)", "median(k)", "max(k)"]
    datout = pd.DataFrame(np.full((len(powerVector), len(colname1)), 666), columns=colname1, dtype=object)
    datout["Power"] = powerVector

    if 1 <= nBreaks <= 16:
        # sweep, histograms, medians and exact mantissa sums on the device (k never leaves HBM); only
        # the two 10-point regressions per power run on the host                       (:873-932)
        X, gene_major = L.as_f64_expression(data)
        m_, n_ = X.shape[::-1] if gene_major else X.shape
        pv = np.asarray(powerVector, dtype=np.float64)
        steps = np.ascontiguousarray(np.diff(np.concatenate(([0.0], pv))))  # :901-903
        lib = L.load()
        dstats = np.empty((len(pv), 4 + 2 * nBreaks), dtype=np.float64)
        dbuck = np.empty((len(pv), lib.b200w_sft_exp_buckets(), 2), dtype=np.int64)
        entry = lib.b200w_soft_threshold_stats_gm if gene_major else lib.b200w_soft_threshold_stats
        L.check(entry(L.hptr(X), m_, n_, L.hptr(steps), len(pv), networkTypes.index(networkType), nBreaks,
                      L.hptr(dstats), L.hptr(dbuck), _device(device)))
        stats = fit_table_from_stats(dstats, dbuck, n_, nBreaks)
    else:
        datk = connectivity(data, powerVector, networkType, device=device)
        stats = fit_table(np.ascontiguousarray(datk.T), nBreaks)  # :924-932, all powers at once
    for name, col in zip(colname1[1:], stats):
        datout[name] = pd.Series(list(col), dtype=object)
    _say(datout)

    # :945-953 -- raw R^2, not the slope-signed one
    r2 = datout["SFT.R.sq"].to_numpy(dtype=np.float64)
    mk = datout["mean(k)"].to_numpy(dtype=np.float64)
    with np.errstate(invalid="ignore"):
        ind = np.logical_and(r2 > RsquaredCut, mk <= MeanCut)
    if np.sum(ind) > 0:
        powerEstimate = np.min(powerVector[ind])
        _say(f"{OKGREEN}Selected power to have scale free network is {str(powerEstimate)}.{ENDC}")
    else:
        order = np.argsort(r2).tolist()
        powerEstimate = powerVector[order[-1]]
        _say(f"{OKGREEN}No power detected to have scale free network!\nFound the best given power which is "
             f"{str(powerEstimate)}.{ENDC}")
    return powerEstimate, datout


# ----------------------------------------------------------------------------------------------
# adjacency
# ----------------------------------------------------------------------------------------------
def adjacency(datExpr, selectCols=None, adjacencyType="unsigned", power=6, corOptions=None, weights=None,
              weightArgNames=None, device=None, rows=None):
    """``WGCNA.adjacency`` (wgcna.py:1047-1111).  Returns the n x n ``numpy.ndarray`` (float64,
    C-contiguous, writable) the reference returns.  ``rows=(row0, nrows)`` (extension) returns only
    that row block."""
    _say(f"{OKCYAN}calculating adjacency matrix ...{ENDC}")
    intType = adjacencyTypes.index(adjacencyType)  # :1073 -> ValueError
    if weights is not None:
        raise NotImplementedError("weights: not supported (np.corrcoef ignores them in the reference too)")
    if selectCols is not None:
        raise NotImplementedError("selectCols: that branch of the reference is invalid for DataFrames (wgcna.py:1100)")
    X, gene_major = L.as_f64_expression(datExpr)
    m, n = X.shape[::-1] if gene_major else X.shape
    row0, nrows = (0, n) if rows is None else (int(rows[0]), int(rows[1]))
    out = L.result_empty((nrows, n))
    lib = L.load()
    entry = lib.b200w_adjacency_gm if gene_major else lib.b200w_adjacency
    L.check(entry(L.hptr(X), m, n, intType, float(power), row0, nrows, L.hptr(out), _device(device)))
    _say("\tDone..\n")
    return out


# ----------------------------------------------------------------------------------------------
# TOM
# ----------------------------------------------------------------------------------------------
def _check_shape_dtype(adjMat):
    shape = adjMat.shape
    if shape is None or len(shape) != 2:  # :1128-1129
        sys.exit("adjacency is not two-dimensional")
    if not issubclass(adjMat.dtype.type, np.floating):  # :1131-1132
        sys.exit("adjacency is not numeric")
    if shape[0] != shape[1]:  # :1133-1134
        sys.exit("adjacency is not square")


def _raise_check(code, lo, hi):
    if code == 1:
        sys.exit("adjacency is not symmetric")  # :1136
    if code == 2:
        sys.exit(("some entries are not between", lo, "and", hi))  # :1138


def checkAdjMat(adjMat, min=0, max=1, device=None):
    """``WGCNA.checkAdjMat`` (wgcna.py:1114-1138); the three n^2 reductions run on the GPU."""
    if hasattr(adjMat, "to_numpy"):
        adjMat = adjMat.to_numpy()
    _check_shape_dtype(adjMat)
    A = np.ascontiguousarray(adjMat, dtype=np.float64)
    stats = np.empty(6, dtype=np.float64)
    L.check(L.load().b200w_check_adjacency(L.hptr(A), A.shape[0], L.hptr(stats), _device(device)))
    if stats[3] > 0:  # a NaN makes np.max / np.min NaN: both comparisons of the reference are False
        return
    if stats[0] > 1e-12:
        _raise_check(1, min, max)
    if stats[1] < min or stats[2] > max:
        _raise_check(2, min, max)


def checkSimilarity(adjMat, min=-1, max=1, device=None):
    """``WGCNA.checkSimilarity`` (wgcna.py:958-985): the similarity-matrix flavour of ``checkAdjMat``
    (only reached through ``pickSoftThreshold(dataIsExpr=False)``, a branch that cannot run in the
    reference, wgcna.py:839); same three n^2 reductions on the GPU, the reference's messages."""
    if hasattr(adjMat, "to_numpy"):
        adjMat = adjMat.to_numpy()
    dim = np.shape(adjMat)
    if dim is None or len(dim) != 2:
        sys.exit("adjacency is not two-dimensional")
    if not np.issubdtype(np.asarray(adjMat).dtype, np.float64):
        sys.exit("adjacency is not numeric")
    if dim[0] != dim[1]:
        sys.exit("adjacency is not square")
    A = np.ascontiguousarray(adjMat, dtype=np.float64)
    stats = np.empty(6, dtype=np.float64)
    L.check(L.load().b200w_check_adjacency(L.hptr(A), A.shape[0], L.hptr(stats), _device(device)))
    if stats[3] > 0:
        return
    if stats[0] > 1e-12:
        sys.exit("adjacency is not symmetric")
    if stats[1] < min or stats[2] > max:
        sys.exit(f"some entries are not between {min} and {max}.")


def _tom(adjMat, TOMTypeC, TOMDenomC, check, out_mode=L.OUT_TOM, device=None, rows=None, precision=None):
    if hasattr(adjMat, "to_numpy"):
        adjMat = adjMat.to_numpy()
    if check:
        _check_shape_dtype(adjMat)
    A = np.ascontiguousarray(adjMat, dtype=np.float64)
    n = A.shape[0]
    row0, nrows = (0, n) if rows is None else (int(rows[0]), int(rows[1]))
    condensed = out_mode in (L.OUT_DISS_COND, L.OUT_DISS_R8_COND)
    out = L.result_empty((n * (n - 1) // 2,) if condensed else (nrows, n))
    stats = np.full(4, np.nan)
    prec = _PREC[(precision or PRECISION).lower()]
    rc = L.load().b200w_tom(L.hptr(A), n, TOMTypeC, TOMDenomC, out_mode, prec, row0, nrows, L.hptr(out),
                            1 if check else 0, L.hptr(stats), _device(device))
    if rc > 0:
        _raise_check(rc, -1 if TOMTypeC == 1 else 0, 1)
    L.check(rc)
    # the reference works in place on its argument (NaN -> 0 :1185, diagonal -> 0 :1143); mirror the
    # visible side effect when the caller's buffer is writable (pandas < 3 hands findModules a view)
    if isinstance(adjMat, np.ndarray) and adjMat.flags.writeable and adjMat.dtype == np.float64:
        if check and stats[3] > 0:
            np.nan_to_num(adjMat, copy=False, nan=0)  # TOMsimilarity only (:1185)
        np.fill_diagonal(adjMat, 0)
    return out


def TomSimilarityFromAdj(adjMat, TOMDenom, TOMType, device=None):
    """``WGCNA.TomSimilarityFromAdj`` (wgcna.py:1141-1157): integer codes in, ndarray out, no checks and
    no ``nan_to_num``.  As in the reference, a NaN makes every entry of its row and of its column NaN
    (numpy's ``A @ A`` and plain sums; the diagonal is still set to 1), and a matrix that is not symmetric
    is contracted as ``A @ A`` with the column sums ``k_j`` (the library then takes its column operand
    from ``A^T`` instead of mirroring tiles)."""
    return _tom(adjMat, int(TOMType), int(TOMDenom), check=False, device=device)


def TOMsimilarity(adjMat, TOMType="signed", TOMDenom="min", device=None, precision=None):
    """``WGCNA.TOMsimilarity`` (wgcna.py:1160-1193): validates, then returns ``pd.DataFrame(tom)``
    with a default RangeIndex."""
    TOMTypeC = TOMTypes.index(TOMType)  # :1174 -> ValueError
    TOMDenomC = TOMDenoms.index(TOMDenom)  # :1177 -> ValueError
    if hasattr(adjMat, "to_numpy"):
        adjMat = adjMat.to_numpy()
    _check_shape_dtype(adjMat)
    _say(f"{OKCYAN}calculating TOM similarity matrix ...{ENDC}")
    tom = _tom(adjMat, TOMTypeC, TOMDenomC, check=True, device=device, precision=precision)
    _say("\tDone..\n")
    # copy=False: the reference's pandas (2.2) wraps the array; pandas >= 3 would copy n^2 doubles
    return pd.DataFrame(tom, copy=False)


def dissTOM(adjMat, TOMType="signed", TOMDenom="min", decimals=8, condensed=False, device=None, precision=None):
    """Fused hand-off of findModules (wgcna.py:320-326): ``(1 - TOM).round(8)`` straight from the TOM
    epilogue, as an ndarray; ``condensed=True`` returns the 1-D vector
    ``squareform((1 - TOM).round(8), checks=False)`` that ``linkage`` consumes (wgcna.py:326-328) --
    half the device-to-host bytes and no host pass over n^2."""
    TOMTypeC = TOMTypes.index(TOMType)
    TOMDenomC = TOMDenoms.index(TOMDenom)
    if decimals not in (None, 8):
        raise ValueError("decimals must be 8 (findModules) or None (no rounding)")
    if condensed:
        mode = L.OUT_DISS_COND if decimals is None else L.OUT_DISS_R8_COND
    else:
        mode = L.OUT_DISS if decimals is None else L.OUT_DISS_R8
    return _tom(adjMat, TOMTypeC, TOMDenomC, check=True, out_mode=mode, device=device, precision=precision)


def network(datExpr, power=6, networkType="unsigned", TOMType="signed", TOMDenom="min", return_adjacency=True,
            device=None, precision=None):
    """adjacency -> TOM in one call with the adjacency kept in HBM (one H2D of the expression
    matrix instead of an n x n round trip).  Returns ``(adjacency ndarray | None, TOM DataFrame)``."""
    intType = adjacencyTypes.index(networkType)
    TOMTypeC = TOMTypes.index(TOMType)
    TOMDenomC = TOMDenoms.index(TOMDenom)
    X = L.as_f64_matrix(datExpr)
    m, n = X.shape
    A = L.result_empty((n, n)) if return_adjacency else None
    T = L.result_empty((n, n))
    prec = _PREC[(precision or PRECISION).lower()]
    L.check(L.load().b200w_network(L.hptr(X), m, n, intType, float(power), TOMTypeC, TOMDenomC, L.OUT_TOM, prec,
                                   L.hptr(A) if A is not None else None, L.hptr(T), _device(device)))
    return A, pd.DataFrame(T, copy=False)


def linkage(y, method="average", metric="euclidean", optimal_ordering=False, device=None):
    """``scipy.cluster.hierarchy.linkage(y, method="average")`` as ``findModules`` calls it on the condensed
    ``round(1 - TOM, 8)`` (wgcna.py:326-328), computed on the GPU with scipy's exact merge order, heights and
    labels (``b200w_linkage_average``).  ``y`` is the condensed distance vector; anything else scipy's
    ``linkage`` accepts (other methods, observation matrices, optimal ordering) is handed to scipy."""
    y = np.asarray(y)
    if method != "average" or y.ndim != 1 or optimal_ordering or len(y) < 1:
        from scipy.cluster.hierarchy import linkage as scipy_linkage

        return scipy_linkage(y, method=method, metric=metric, optimal_ordering=optimal_ordering)
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = int(round((1 + math.sqrt(1 + 8 * len(y))) / 2))
    if n * (n - 1) // 2 != len(y):
        raise ValueError("Incompatible vector size. It must be a binomial coefficient n choose 2 for some integer n >= 2.")
    Z = np.empty((n - 1, 4), dtype=np.float64)
    L.check(L.load().b200w_linkage_average(L.hptr(y), n, L.hptr(Z), _device(device)))
    return Z


def network_blocks(datExpr, *args, **kwargs):
    """Sharded host entry point: expression matrix in, this rank's row block of the TOM out (one process per
    GPU under torchrun); see ``pywgcna_b200.device.network_blocks``.  (Imported lazily: it needs torch.)"""
    from .device import network_blocks as impl

    return impl(datExpr, *args, **kwargs)


# ----------------------------------------------------------------------------------------------
# connectivity helpers
# ----------------------------------------------------------------------------------------------
def softConnectivity(datExpr, corOptions=None, weights=None, type="unsigned", power=6, blockSize=1500,
                     minNSamples=None, device=None):
    """``WGCNA.softConnectivity`` (wgcna.py:3342-3400) as it was meant to work:
    ``k_i = sum_j adjacency_ij - 1``.  The reference function cannot run (1-based block loop :3388,
    RLIMIT_AS == -1 :3371, invalid ``np.corrcoef`` call :1100 -- SURVEY App. A.6); the argument
    checks, the forced ``power = 15`` for ``type == "signed"`` (:3365-3366) and numpy's plain
    (NaN-propagating) row sum (:3393) are kept."""
    if type == "signed":
        power = 15
    adjacencyTypes.index(type)
    nGenes = datExpr.shape[1]
    nSamples = datExpr.shape[0]
    if minNSamples is None:
        minNSamples = max(4, nSamples / 3)
    if nGenes < 10 or nSamples < minNSamples:
        sys.exit(f"Error: Something seems to be wrong. \n Make sure that the input data frame has genes as rows and "
                 f"array samples as columns.\n Alternatively, there seem to be fewer than 10 genes or fewer than "
                 f"{minNSamples} samples.")
    if nGenes < nSamples:
        _say(f"{WARNING}Warning: There are fewer genes than samples in the function softConnectivity. Maybe you "
             f"should transpose the data?{ENDC}")
    if weights is not None:
        raise NotImplementedError("weights: not supported")
    k, bad = connectivity(datExpr, [power], type, device=device, return_bad=True)
    k = np.ascontiguousarray(k[:, 0])
    if bad.any():  # a NaN gene puts a NaN in every row of the adjacency: ad1.sum(axis=1) is NaN
        k[:] = np.nan
    return k


def intramodularConnectivity(mat, colors, scaleByMax=False, index=None, device=None):
    """``WGCNA.intramodularConnectivity`` (wgcna.py:3403-3450)."""
    if hasattr(mat, "to_numpy"):
        mat = mat.to_numpy()
    colors = np.asarray(colors)
    if mat.shape[0] != mat.shape[1]:
        sys.exit("input matrix is not a square matrix.")
    if mat.shape[0] != len(colors):
        sys.exit("Dimensions of matrix (number of genes) and length of 'colors' differ.")
    nNodes = len(colors)
    cat = pd.Categorical(colors)
    labels = np.ascontiguousarray(cat.codes, dtype=np.int32)
    nmod = len(cat.categories)
    sizes = np.bincount(labels, minlength=nmod)
    small = np.ascontiguousarray(sizes < 3, dtype=np.uint8)  # :3436
    A = np.ascontiguousarray(mat, dtype=np.float64)
    kTotal = np.empty(nNodes)
    kWithin = np.empty(nNodes)
    L.check(L.load().b200w_intramodular(L.hptr(A), nNodes, L.hptr(labels), nmod, L.hptr(small), L.hptr(kTotal),
                                        L.hptr(kWithin), _device(device)))
    if scaleByMax:  # :3440-3441
        for c in range(nmod):
            if not small[c]:
                sel = labels == c
                kWithin[sel] = kWithin[sel] / max(kWithin[sel])
    kOut = np.subtract(kTotal, kWithin)
    if scaleByMax:
        kOut = np.zeros((nNodes,))
    kDiff = np.subtract(kWithin, kOut)
    res = pd.DataFrame({"kTotal": kTotal, "kWithin": kWithin, "kOut": kOut, "kDiff": kDiff})
    if index is not None:
        res.index = index
    return res


def signedKME(datExpr, datME, device=None):
    """Eigengene-based connectivity (module membership): ``cor(gene, eigengene)`` for every gene and
    module eigengene, the n x M block ``WGCNA.CalculateSignedKME`` slices out of
    ``np.corrcoef(datExpr.T, datME.T)`` (wgcna.py:3486-3489) -- computed directly, without the
    (n + M)^2 matrix.  Returns a DataFrame indexed by gene with columns ``"k" + eigengene name``."""
    if datME.shape[0] != datExpr.shape[0]:
        sys.exit("Number of samples (rows) in 'datExpr' and 'datME' must be the same.")  # :3462
    X = L.as_f64_matrix(datExpr)
    Y = L.as_f64_matrix(datME)
    m, n = X.shape
    M = Y.shape[1]
    out = np.empty((n, M), dtype=np.float64)
    L.check(L.load().b200w_cross_correlation(L.hptr(X), m, n, L.hptr(Y), M, L.hptr(out), _device(device)))
    idx = datExpr.columns if hasattr(datExpr, "columns") else None
    cols = ["k" + str(c) for c in (datME.columns if hasattr(datME, "columns") else range(M))]
    return pd.DataFrame(out, index=idx, columns=cols)


def CalculateSignedKME(self, exprWeights=None, MEWeights=None):
    """Instance-method replacement of ``WGCNA.CalculateSignedKME`` (wgcna.py:3452-3490): fills
    ``self.signedKME`` from ``self.datExpr`` / ``self.datME`` with the GPU cross-correlation."""
    if exprWeights is not None or MEWeights is not None:
        raise NotImplementedError("weights: np.corrcoef ignores them in the reference too")
    self.signedKME = signedKME(self.datExpr.to_df(), self.datME)


# ----------------------------------------------------------------------------------------------
# drop-in installation
# ----------------------------------------------------------------------------------------------
_PATCHED = ("pickSoftThreshold", "scaleFreeFitIndex", "calBlockSize", "checkSimilarity", "adjacency", "checkAdjMat",
            "TOMsimilarity", "TomSimilarityFromAdj", "softConnectivity", "intramodularConnectivity")
_saved = {}


def install(target=None, accelerate_linkage=True):
    """Rebind the hot-path static methods on the reference class so ``WGCNA.findModules()`` (which
    calls ``WGCNA.pickSoftThreshold / adjacency / TOMsimilarity`` by class name, wgcna.py:278, :310,
    :320) runs them on the GPU.  ``target`` defaults to ``PyWGCNA.wgcna.WGCNA``.  With ``accelerate_linkage``
    the module-level ``linkage`` that findModules calls next (wgcna.py:328) is rebound to the GPU average
    linkage as well (same result as scipy's, bit for bit)."""
    if target is None:
        from PyWGCNA.wgcna import WGCNA as target  # noqa: N811
    L.load()  # fail now, not in the middle of findModules
    g = globals()
    for name in _PATCHED:
        if (target, name) not in _saved:
            _saved[(target, name)] = target.__dict__.get(name)
        setattr(target, name, staticmethod(g[name]))
    if (target, "CalculateSignedKME") not in _saved:
        _saved[(target, "CalculateSignedKME")] = target.__dict__.get("CalculateSignedKME")
    target.CalculateSignedKME = CalculateSignedKME  # an instance method in the reference
    # findModules calls the module-level name `linkage` (from scipy.cluster.hierarchy, wgcna.py:10, :328)
    mod = sys.modules.get(getattr(target, "__module__", ""), None)
    if accelerate_linkage and mod is not None and hasattr(mod, "linkage") and (mod, "linkage") not in _saved:
        _saved[(mod, "linkage")] = mod.linkage
        mod.linkage = linkage
    return target


def uninstall(target=None):
    if target is None:
        from PyWGCNA.wgcna import WGCNA as target  # noqa: N811
    for name in _PATCHED + ("CalculateSignedKME",):
        if (target, name) not in _saved:
            continue
        orig = _saved.pop((target, name))
        if orig is not None:
            setattr(target, name, orig)
        elif name in target.__dict__:
            delattr(target, name)  # the class had no such attribute before install()
    mod = sys.modules.get(getattr(target, "__module__", ""), None)
    if mod is not None and (mod, "linkage") in _saved:
        mod.linkage = _saved.pop((mod, "linkage"))
    return target
