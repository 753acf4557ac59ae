"""TEST INFRASTRUCTURE ONLY -- runs the reference's own ``WGCNA.findModules`` (PyWGCNA/wgcna.py:265-404),
unmodified, on an in-memory expression matrix.  Used to generate the end-to-end fixtures
(``oracle/gen_golden_large.py``) and by the GPU end-to-end test, where ``pywgcna_b200.install()`` has
rebound the three hot-path static methods on the same class (wgcna.py:278, :310, :320) and everything
else -- linkage, cutreeHybrid, labels2colors, moduleEigengenes, mergeCloseModules -- stays the
reference's host code.  Nothing under ``pywgcna_b200/`` may import this module.

What is shimmed, and why (all of it dependency absence / drift; no arithmetic is replaced):

* ``anndata`` is not installed: ``AnnDataShim`` offers the handful of members findModules touches
  (``to_df()``, ``var``, ``obs``, ``shape``, ``X``, ``copy()``).
* ``matplotlib`` is not installed: the stub ``plt`` of ``oracle/ref_import.py`` gets a ``subplots`` that
  can be unpacked (wgcna.py:282), scipy's ``dendrogram`` (wgcna.py:354, :2387) runs with ``no_plot=True``
  (drawing imports matplotlib; the leaf order it returns is unchanged), and ``labels2colors`` is given an explicit ``colorSeq`` through findModules' own
  ``kwargs_function['labels2colors']`` (wgcna.py:339-341) because the default needs ``mcolors.CSS4_COLORS``.
* pandas >= 3 hands out read-only ``to_numpy()`` views (the reference mutates that array in place,
  wgcna.py:1185, :1143, as pandas 2.2 allowed): ``DataFrame.to_numpy`` returns the view writable again.
* pandas >= 3 copy-on-write / scipy >= 1.15 ``rankdata``: ``WGCNA.cutreeHybrid`` is rebound to the
  CoW-safe re-execution of its own source (``ref_import._reference_cutreeHybrid``), and
  ``pd.options.future.infer_string`` is switched off while the reference runs.
"""
from __future__ import annotations

import contextlib
import os
import time

import numpy as np
import pandas as pd

from oracle import ref_import as R

# explicit colour sequence for labels2colors (cosmetic: only names, the partition is what is compared)
COLOR_SEQ = [f"module{i:03d}" for i in range(400)]


class AnnDataShim:
    """The members of ``anndata.AnnData`` that WGCNA.findModules uses (wgcna.py:278-402)."""

    def __init__(self, X: np.ndarray, obs: pd.DataFrame, var: pd.DataFrame):
        self.X = X
        self.obs = obs
        self.var = var

    @property
    def shape(self):
        return self.X.shape

    def to_df(self) -> pd.DataFrame:
        return pd.DataFrame(self.X, index=self.obs.index, columns=self.var.index)

    def copy(self):
        return AnnDataShim(self.X.copy(), self.obs.copy(), self.var.copy())

    def __getitem__(self, key):
        """``adata[:, mask]`` as ``top_n_hub_genes`` slices it (wgcna.py:3746)."""
        rows, cols = key
        r = np.arange(self.X.shape[0])[rows] if not isinstance(rows, slice) or rows != slice(None) else slice(None)
        c = np.asarray(cols)
        return AnnDataShim(self.X[r][:, c], self.obs.iloc[r] if not isinstance(r, slice) else self.obs, self.var.loc[c]
                           if c.dtype == bool else self.var.iloc[c])


@contextlib.contextmanager
def _reference_environment(W):
    """Dependency-drift shims around a findModules run (see the module docstring)."""
    saved_cut = W.WGCNA.__dict__.get("cutreeHybrid")
    saved_dendro = W.dendrogram
    saved_subplots = getattr(W.plt, "subplots", None)
    saved_infer = pd.options.future.infer_string if hasattr(pd.options.future, "infer_string") else None
    W.WGCNA.cutreeHybrid = staticmethod(R._reference_cutreeHybrid())
    W.dendrogram = lambda *a, **k: saved_dendro(*a, **{**k, "no_plot": True})  # same leaves order, no drawing
    if isinstance(W.plt, R._Stub):
        W.plt.subplots = lambda *a, **k: (R._Callable("fig"), [R._Callable("ax0"), R._Callable("ax1")])
    if saved_infer is not None:
        pd.options.future.infer_string = False
    saved_to_numpy = pd.DataFrame.to_numpy

    def to_numpy_writable(self, *a, **k):
        # pandas 2.2 (the pinned version) hands findModules a writable view at wgcna.py:320, which
        # TOMsimilarity then mutates in place (:1185, :1143); pandas >= 3 marks the view read-only
        arr = saved_to_numpy(self, *a, **k)
        if not arr.flags.writeable:
            try:
                arr.flags.writeable = True
            except ValueError:
                arr = arr.copy()
        return arr

    pd.DataFrame.to_numpy = to_numpy_writable
    try:
        with R.reference_sft_semantics(), np.errstate(all="ignore"):
            yield
    finally:
        pd.DataFrame.to_numpy = saved_to_numpy
        W.WGCNA.cutreeHybrid = saved_cut
        W.dendrogram = saved_dendro
        if saved_subplots is not None and isinstance(W.plt, R._Stub):
            W.plt.subplots = saved_subplots
        if saved_infer is not None:
            pd.options.future.infer_string = saved_infer


def make_wgcna(X: np.ndarray, networkType="signed hybrid", TOMType="signed", powers=None, minModuleSize=50,
               RsquaredCut=0.9, MeanCut=100, MEDissThres=0.2, name="e2e"):
    """A reference ``WGCNA`` object over ``X`` (samples x genes) without going through ``GeneExp.__init__``
    (which needs the real anndata): the attributes are the ones ``WGCNA.__init__`` sets (wgcna.py:146-205)."""
    W = R.load_reference()
    m, n = X.shape
    genes = pd.Index([f"g{j:05d}" for j in range(n)], dtype=object)
    samples = pd.Index([f"s{i:03d}" for i in range(m)], dtype=object)
    ad = AnnDataShim(np.ascontiguousarray(X, dtype=np.float64), pd.DataFrame(index=samples), pd.DataFrame(index=genes))
    obj = object.__new__(W.WGCNA)
    obj.species, obj.level = None, "gene"
    obj.geneExpr = ad
    obj.name = name
    obj.save = False
    obj.outputPath = os.getcwd() + "/"
    obj.TPMcutoff, obj.cut = 1, float("inf")
    obj.datExpr = ad.copy()
    obj.metadataColors = {}
    obj.networkType = networkType
    obj.RsquaredCut, obj.MeanCut = RsquaredCut, MeanCut
    obj.powers = list(range(1, 11)) + list(range(11, 21, 2)) if powers is None else list(powers)
    obj.power, obj.sft = 0, None
    obj.geneTree = obj.adjacency = obj.TOM = None
    obj.TOMType = TOMType
    obj.minModuleSize = minModuleSize
    obj.dynamicMods = [None]
    obj.naColor = "grey"
    obj.MEs = None
    obj.MEDissThres = MEDissThres
    obj.datME = obj.signedKME = None
    obj.moduleTraitCor = obj.moduleTraitPvalue = None
    obj.figureType = "pdf"
    return obj


def run_findModules(obj, quiet=True, kwargs_function=None):
    """``obj.findModules(...)`` -- the reference's own method, as is -- inside the shim environment.
    Returns a dict with everything the end-to-end check compares, plus wall times of the stages the
    (possibly rebound) static methods took."""
    W = R.load_reference()
    kw = {"cutreeHybrid": {"deepSplit": 2, "pamRespectsDendro": False}, "labels2colors": {"colorSeq": COLOR_SEQ}}
    if kwargs_function:
        kw.update(kwargs_function)
    timings = {}
    raw = {}

    def timed(name, fn):
        def wrapper(*a, **k):
            t0 = time.perf_counter()
            out = fn(*a, **k)
            timings[name] = timings.get(name, 0.0) + time.perf_counter() - t0
            if name == "cutreeHybrid":
                raw["dynamicMods"] = np.asarray(out["Value"], dtype=np.int64).copy()
            return out
        return wrapper

    names = ("pickSoftThreshold", "adjacency", "TOMsimilarity", "cutreeHybrid")
    with _reference_environment(W):
        saved = {nm: W.WGCNA.__dict__[nm] for nm in names}
        for nm in names:
            fn = saved[nm].__func__ if isinstance(saved[nm], staticmethod) else saved[nm]
            setattr(W.WGCNA, nm, staticmethod(timed(nm, fn)))
        t0 = time.perf_counter()
        try:
            with (R._quiet() if quiet else contextlib.nullcontext()):
                W.WGCNA.findModules(obj, kwargs_function=kw)
        finally:
            for nm in names:
                setattr(W.WGCNA, nm, saved[nm])
        timings["findModules"] = time.perf_counter() - t0
    var = obj.datExpr.var
    return {
        "power": np.int64(obj.power),
        "sft": obj.sft.to_numpy(dtype=np.float64),
        "linkage": np.asarray(obj.geneTree, dtype=np.float64),
        "dynamicMods": raw["dynamicMods"],
        "dynamicColors": np.asarray(var["dynamicColors"]).astype("U16"),
        "moduleColors": np.asarray(var["moduleColors"]).astype("U16"),
        "moduleLabels": np.asarray(var["moduleLabels"], dtype=np.int64),
        "timings": timings,
    }
