"""TEST INFRASTRUCTURE ONLY -- import harness for the *unmodified* PyWGCNA reference.

This module is only usable where the reference tree exists (this container:
``/root/reference``).  It is used to (a) validate ``oracle/restated.py`` and
(b) generate the committed golden fixtures under ``tests/golden/``
(``oracle/gen_golden.py``).  Nothing under ``pywgcna_b200/`` may import it.

The reference cannot be imported as-is: ``PyWGCNA/__init__.py:1-4`` star-imports
every module and ``PyWGCNA/wgcna.py:15-27`` pulls in statsmodels, matplotlib,
seaborn, gseapy, pyvis, reactome2py, anndata, biomart -- none installed here.
We pre-register permissive stub modules for those, provide a tiny numpy
stand-in for ``statsmodels.formula.api.ols`` (the only third-party arithmetic
on the hot path that is missing), then import ``PyWGCNA.wgcna``.

Dependency drift neutralised here (authoritative semantics = the versions the
reference pins in ``requirements.txt:6,7,13``):

* pandas>=3 ``groupby`` on a categorical defaults to ``observed=True`` and would
  silently drop empty connectivity bins in ``scaleFreeFitIndex``
  (``wgcna.py:1019``); ``reference_sft_semantics()`` restores pandas-2.2
  ``observed=False`` while the reference function runs.
* pandas>=3 copy-on-write makes ``DataFrame.to_numpy()`` read-only; the
  reference's in-place ``np.nan_to_num``/``fill_diagonal`` (``wgcna.py:1185,1143``)
  then raise.  ``ref_TOMsimilarity`` therefore always hands it a writable copy.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np
import pandas as pd

# $B200W_REF, the source tree of the build container, or the `pip install --target baseline/_ref` copy that
# travels to the GPU box (git-ignored; __graft_entry__.build() creates it where /root/reference exists)
_REF_CANDIDATES = [os.environ.get("B200W_REF", ""), "/root/reference",
                   os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")]


class _Stub(types.ModuleType):
    """A module whose every attribute is another stub / a no-op callable."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _Callable(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child


class _Callable(_Stub):
    def __call__(self, *a, **k):
        return _Callable(self.__name__ + "()")

    def __getitem__(self, k):
        return _Callable(self.__name__ + "[]")

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter(())


class _OLSResult:
    def __init__(self, rsq, rsq_adj, params):
        self.rsquared = rsq
        self.rsquared_adj = rsq_adj
        self.params = pd.Series(params)


class _OLSModel:
    """numpy least squares stand-in for statsmodels ``ols('y ~ a + b', df)``.

    statsmodels 0.14: ``rsquared = 1 - ssr/centered_tss``;
    ``rsquared_adj = 1 - (nobs-1)/df_resid * (1-rsquared)`` with
    ``df_resid = nobs - rank(X)`` (intercept included in X).
    """

    def __init__(self, formula, data):
        lhs, rhs = formula.split("~")
        self.y = np.asarray(data[lhs.strip()], dtype=np.float64)
        cols = [c.strip() for c in rhs.split("+")]
        self.X = np.column_stack([np.ones(len(self.y))] + [np.asarray(data[c], dtype=np.float64) for c in cols])

    def fit(self):
        beta, _, rank, _ = np.linalg.lstsq(self.X, self.y, rcond=None)
        resid = self.y - self.X @ beta
        ssr = float(resid @ resid)
        yc = self.y - self.y.mean()
        tss = float(yc @ yc)
        rsq = 1.0 - ssr / tss
        n = len(self.y)
        rsq_adj = 1.0 - (n - 1) / (n - rank) * (1.0 - rsq)
        return _OLSResult(rsq, rsq_adj, beta)


def _install_stubs():
    names = [
        "statsmodels", "statsmodels.formula", "statsmodels.formula.api",
        "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.gridspec",
        "matplotlib.colors", "matplotlib.lines", "matplotlib.cm",
        "seaborn", "gseapy", "gseapy.plot", "pyvis", "pyvis.network",
        "reactome2py", "anndata", "biomart",
    ]
    for n in names:
        if n in sys.modules:
            continue
        try:
            __import__(n)
            continue
        except Exception:
            pass
        sys.modules[n] = _Stub(n)
    for n in names:  # parent.child attribute links
        if "." in n:
            parent, child = n.rsplit(".", 1)
            if isinstance(sys.modules.get(parent), _Stub):
                setattr(sys.modules[parent], child, sys.modules[n])
    api = sys.modules["statsmodels.formula.api"]
    if isinstance(api, _Stub):
        api.ols = lambda formula, data: _OLSModel(formula, data)


_W = None


def reference_available() -> bool:
    return any(p and os.path.isdir(os.path.join(p, "PyWGCNA")) for p in _REF_CANDIDATES)


def load_reference():
    """Return the reference's ``PyWGCNA.wgcna`` module (unmodified source)."""
    global _W
    if _W is not None:
        return _W
    root = next((p for p in _REF_CANDIDATES if p and os.path.isdir(os.path.join(p, "PyWGCNA"))), None)
    if root is None:
        raise RuntimeError("PyWGCNA reference tree not found (set B200W_REF)")
    _install_stubs()
    # avoid PyWGCNA/__init__.py (it star-imports comparison/utils with more deps)
    pkg = types.ModuleType("PyWGCNA")
    pkg.__path__ = [os.path.join(root, "PyWGCNA")]
    sys.modules.setdefault("PyWGCNA", pkg)
    import importlib

    with open(os.devnull, "w") as devnull, contextlib.redirect_stdout(devnull):
        W = importlib.import_module("PyWGCNA.wgcna")
    _W = W
    return W


@contextlib.contextmanager
def reference_sft_semantics():
    """pandas-2.2 ``groupby(observed=False)`` while reference code runs."""
    orig = pd.DataFrame.groupby

    def groupby(self, *a, **k):
        k.setdefault("observed", False)
        return orig(self, *a, **k)

    pd.DataFrame.groupby = groupby
    try:
        yield
    finally:
        pd.DataFrame.groupby = orig


@contextlib.contextmanager
def _quiet():
    with open(os.devnull, "w") as devnull, contextlib.redirect_stdout(devnull):
        yield


def ref_pickSoftThreshold(data: pd.DataFrame, **kw):
    W = load_reference()
    with reference_sft_semantics(), _quiet(), np.errstate(all="ignore"):
        return W.WGCNA.pickSoftThreshold(data, **kw)


def ref_scaleFreeFitIndex(k, nBreaks=10):
    W = load_reference()
    with reference_sft_semantics(), np.errstate(all="ignore"):
        return W.WGCNA.scaleFreeFitIndex(k, nBreaks=nBreaks)


def ref_adjacency(data: pd.DataFrame, **kw):
    W = load_reference()
    with _quiet(), np.errstate(all="ignore"):
        return W.WGCNA.adjacency(data, **kw)


def ref_TOMsimilarity(adj: np.ndarray, **kw):
    W = load_reference()
    adj = np.array(adj, dtype=np.float64, copy=True)  # the reference mutates its input
    with _quiet(), np.errstate(all="ignore"):
        return W.WGCNA.TOMsimilarity(adj, **kw)


def ref_intramodularConnectivity(adj: np.ndarray, colors, **kw):
    W = load_reference()
    adj = np.array(adj, dtype=np.float64, copy=True)
    with _quiet(), np.errstate(all="ignore"):
        return W.WGCNA.intramodularConnectivity(adj, np.asarray(colors), **kw)


# --------------------------------------------------------------------------------------------
# end-to-end check: the reference's own clustering on top of a TOM (findModules, wgcna.py:323-334)
# --------------------------------------------------------------------------------------------
_CUTREE = None


def _reference_cutreeHybrid():
    """``WGCNA.cutreeHybrid`` with the two dependency-drift shims of SURVEY 8c: pandas >= 3
    copy-on-write turns the chained assignment ``OrdNumLabs.Value[i] = ...`` (wgcna.py:1855, :1859)
    into a no-op, and scipy >= 1.15 ``rankdata`` returns floats that the reference uses as indices
    (wgcna.py:1452-1453).  The function source is otherwise executed unmodified."""
    global _CUTREE
    if _CUTREE is not None:
        return _CUTREE
    import inspect
    import textwrap

    from scipy import stats as _stats

    W = load_reference()
    src = textwrap.dedent(inspect.getsource(W.WGCNA.cutreeHybrid))
    src = src.replace("@staticmethod\n", "")
    assert src.count("OrdNumLabs.Value[i] = ") == 2
    src = src.replace("OrdNumLabs.Value[i] = ", "OrdNumLabs.loc[i, 'Value'] = ")

    def rankdata_int(*a, **k):
        return _stats.rankdata(*a, **k).astype(np.int64)

    class _StatsShim:
        def __getattr__(self, name):
            return rankdata_int if name == "rankdata" else getattr(_stats, name)

    ns = dict(W.__dict__)
    ns["rankdata"] = rankdata_int
    ns["stats"] = _StatsShim()
    exec(compile(src, "<reference cutreeHybrid, CoW-safe>", "exec"), ns)
    _CUTREE = ns["cutreeHybrid"]
    return _CUTREE


def ref_modules_from_tom(tom: np.ndarray, minClusterSize=50, deepSplit=2, pamRespectsDendro=False):
    """dissTOM -> linkage(average) -> cutreeHybrid exactly as findModules does (wgcna.py:323-334,
    default kwargs of :265).  Returns (labels, linkage matrix)."""
    from scipy.cluster.hierarchy import linkage
    from scipy.spatial.distance import squareform

    prev = pd.options.future.infer_string if hasattr(pd.options.future, "infer_string") else None
    diss = pd.DataFrame(1 - np.asarray(tom, dtype=np.float64)).round(decimals=8)
    a = squareform(diss.values, checks=False)
    Z = linkage(a, method="average")
    cut = _reference_cutreeHybrid()
    with _quiet(), np.errstate(all="ignore"):
        labels = cut(dendro=Z, distM=diss, minClusterSize=minClusterSize, deepSplit=deepSplit,
                     pamRespectsDendro=pamRespectsDendro)
    del prev
    # cutreeHybrid returns the frame OrdNumLabs (wgcna.py:1848, :1864); "Value" holds the module labels
    return np.asarray(labels["Value"], dtype=np.int64), Z
