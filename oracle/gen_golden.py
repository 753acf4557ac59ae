"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the
UNMODIFIED reference (``/root/reference/PyWGCNA/wgcna.py`` via
``oracle/ref_import.py``) on small seeded inputs.  Run in the build container:

    python -m oracle.gen_golden

The fixtures hold inputs AND reference outputs, so the tests that consume them
need neither the reference tree nor this generator at run time.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

from oracle import ref_import as R
from pywgcna_b200.synth import make_expression

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
NET = ["unsigned", "signed", "signed hybrid"]
DEFAULT_POWERS = list(range(1, 11)) + list(range(11, 21, 2))  # WGCNA.__init__, wgcna.py:159-160


def sft_to_array(sft: pd.DataFrame) -> np.ndarray:
    return sft.to_numpy(dtype=np.float64)


def case_pipeline(name, X, powers, tom_power=6):
    df = pd.DataFrame(X)
    out = {"X": X, "powers": np.asarray(powers)}
    for net in NET:
        key = net.replace(" ", "_")
        p, sft = R.ref_pickSoftThreshold(df, networkType=net, powerVector=list(powers))
        out[f"power_{key}"] = np.int64(p)
        out[f"sft_{key}"] = sft_to_array(sft)
        A = R.ref_adjacency(df, adjacencyType=net, power=tom_power)
        out[f"adj_{key}"] = A
        for tt in ["unsigned", "signed"]:
            for td in ["min", "mean"]:
                T = R.ref_TOMsimilarity(A, TOMType=tt, TOMDenom=td)
                out[f"tom_{key}_{tt}_{td}"] = T.to_numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: np.shape(v) for k, v in out.items() if k.startswith(("sft", "power"))})


def case_tom_from_adj():
    """TomSimilarityFromAdj called directly (wgcna.py:1141-1157: no checks, no nan_to_num) and TOMsimilarity
    on matrices that a NaN waves through checkAdjMat (:1135-1138 compare against NaN): asymmetric input
    (A @ A, column sums k_j) and NaN propagation."""
    W = R.load_reference()
    rng = np.random.default_rng(707)
    n = 131
    out = {}
    B = rng.uniform(0, 1, size=(n, n)) ** 3
    asym = B.copy()                                   # plainly asymmetric
    sym = (B + B.T) / 2
    sym_nan = sym.copy()
    sym_nan[7, 100] = sym_nan[100, 7] = np.nan        # symmetric NaN pattern
    sym_nan[90, 90] = np.nan                          # a NaN on the diagonal is zeroed first (:1143)
    asym_nan = sym.copy()
    asym_nan[3, 50] = np.nan                          # NaN in (3, 50) only, and a 0 opposite it
    asym_nan[50, 3] = 0.0
    asym_nan[110, 111] += 0.25                        # and a real asymmetry
    signed = (rng.uniform(-1, 1, size=(n, n)))        # asymmetric with negative entries (signed formula)
    for nm, A in (("asym", asym), ("sym_nan", sym_nan), ("asym_nan", asym_nan), ("signed_asym", signed)):
        out[f"A_{nm}"] = A
        for tt, td in ((0, 0), (1, 1)):
            with np.errstate(all="ignore"):
                out[f"fromadj_{nm}_{tt}_{td}"] = W.WGCNA.TomSimilarityFromAdj(A.copy(), td, tt)
    # the checked entry point: NaN -> checks pass vacuously -> nan_to_num -> asymmetric contraction
    for nm in ("sym_nan", "asym_nan"):
        for tt in ("unsigned", "signed"):
            out[f"tomsim_{nm}_{tt}"] = R.ref_TOMsimilarity(out[f"A_{nm}"], TOMType=tt).to_numpy()
    np.savez_compressed(os.path.join(OUT, "tom_from_adj.npz"), **out)
    print("tom_from_adj", len(out), "arrays")


def main():
    os.makedirs(OUT, exist_ok=True)
    # (1) planted modules, default power vector of WGCNA.__init__
    X = make_expression(48, 150, F=5, seed=101)
    case_pipeline("pipeline_small", X, DEFAULT_POWERS)
    # (2) ragged sizes (n not a multiple of any tile), pickSoftThreshold's own default
    #     power vector with duplicates (wgcna.py:823), expression-like values
    X = make_expression(37, 131, F=4, seed=202, expression_like=True)
    case_pipeline("pipeline_ragged", X, list(range(1, 11)) + list(range(1, 21, 2)))
    # (3) edge cases: zero-variance gene, a gene with a NaN, two identical genes,
    #     an exactly anti-correlated pair
    X = make_expression(30, 64, F=3, seed=303)
    X[:, 5] = 2.0            # zero variance -> NaN row/column in corrcoef
    X[3, 9] = np.nan         # NaN sample -> NaN row/column
    X[:, 20] = X[:, 21]      # duplicate genes (cor = 1)
    X[:, 30] = -X[:, 31]     # cor = -1
    case_pipeline("pipeline_edge", X, list(range(1, 21)))
    # (4) user-supplied signed adjacency with negative entries (only there do the
    #     signed and unsigned TOM formulas differ)
    rng = np.random.default_rng(404)
    n = 96
    B = rng.uniform(-1, 1, size=(n, n))
    A = (B + B.T) / 2
    np.fill_diagonal(A, 1.0)
    out = {"A": A}
    for td in ["min", "mean"]:
        out[f"tom_signed_{td}"] = R.ref_TOMsimilarity(A, TOMType="signed", TOMDenom=td).to_numpy()
    Apos = np.abs(A)
    for td in ["min", "mean"]:
        out[f"tom_unsigned_{td}"] = R.ref_TOMsimilarity(Apos, TOMType="unsigned", TOMDenom=td).to_numpy()
    out["Apos"] = Apos
    # intramodularConnectivity (wgcna.py:3403-3450)
    colors = np.array(["turquoise", "blue", "brown", "grey"])[rng.integers(0, 4, size=n)]
    colors[:2] = "tiny"  # a module with < 3 genes -> kWithin = 0
    imc = R.ref_intramodularConnectivity(Apos, colors)
    out["imc_colors"] = colors.astype("U16")
    out["imc"] = imc.to_numpy()
    np.savez_compressed(os.path.join(OUT, "tom_signed_user.npz"), **out)
    # (5) scaleFreeFitIndex alone, including a k vector that leaves bins empty
    ks = {"k_smooth": rng.gamma(2.0, 30.0, size=400),
          "k_gappy": np.concatenate([rng.uniform(0, 1, 300), [50.0, 51.0, 100.0]]),
          "k_with_zero": np.concatenate([[0.0, 0.0, 0.0], rng.uniform(0, 5, 200)])}
    out = {}
    for nm, k in ks.items():
        out[nm] = k
        out[nm + "_fit"] = R.ref_scaleFreeFitIndex(k).to_numpy(dtype=np.float64)[0]
    np.savez_compressed(os.path.join(OUT, "scale_free_fit.npz"), **out)
    # (6) end to end through the reference's own clustering (findModules, wgcna.py:278-334): power,
    #     adjacency, TOM, round(1 - TOM, 8), linkage(average), cutreeHybrid -> module labels
    X = make_expression(100, 600, F=5, seed=606, grey_frac=0.25, loading=(1.5, 3.0))  # clear modules
    df = pd.DataFrame(X)
    p, sft = R.ref_pickSoftThreshold(df, networkType="signed hybrid", powerVector=DEFAULT_POWERS)
    A = R.ref_adjacency(df, adjacencyType="signed hybrid", power=p)
    T = R.ref_TOMsimilarity(A, TOMType="signed").to_numpy()
    labels, Z = R.ref_modules_from_tom(T, minClusterSize=30)
    np.savez_compressed(os.path.join(OUT, "e2e_modules.npz"), X=X, powers=np.asarray(DEFAULT_POWERS), power=np.int64(p),
                        sft=sft_to_array(sft), labels=labels, linkage=Z, tom_sample=T[:8].copy())
    print("e2e_modules: power", p, "modules", np.unique(labels, return_counts=True))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    import sys

    if len(sys.argv) > 1 and sys.argv[1] == "tom_from_adj":
        case_tom_from_adj()
    else:
        main()
        case_tom_from_adj()
