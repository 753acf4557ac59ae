"""Frozen synthetic expression generators for the BASELINE.json configs
(SURVEY.md section 8d).  Planted-module data: F latent factors, 30% pure-noise
("grey") genes, loadings U(0.4, 1.2) with sign -1 w.p. 0.2 so that signed and
unsigned networks differ.  float64, samples x genes, C-contiguous."""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: (samples m, genes n, factors F, seed, networkType, TOMType)
    "C1": (192, 23844, 40, 5192, "signed hybrid", "signed"),
    "C2": (200, 20000, 40, 20200, "signed hybrid", "signed"),
    "C3": (500, 50000, 60, 50500, "unsigned", "unsigned"),
    "C4": (5000, 30000, 60, 305000, "signed", "signed"),
    "C5": (300, 80000, 80, 80300, "unsigned", "unsigned"),
}


def make_expression(m: int, n: int, F: int = 40, seed: int = 0, grey_frac: float = 0.3,
                    neg_frac: float = 0.2, expression_like: bool = False, return_modules: bool = False,
                    loading=(0.4, 1.2)):
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((m, F))
    module = rng.integers(0, F, size=n)
    grey = rng.random(n) < grey_frac
    lam = rng.uniform(loading[0], loading[1], size=n)
    lam[rng.random(n) < neg_frac] *= -1.0
    lam[grey] = 0.0
    X = rng.standard_normal((m, n))
    # chunked to bound the temporary Z[:, module] (m x n) at large n
    step = 8192
    for s in range(0, n, step):
        e = min(s + step, n)
        X[:, s:e] += Z[:, module[s:e]] * lam[s:e]
    if expression_like:
        X = np.log2(1.0 + np.exp(X))
    X = np.ascontiguousarray(X, dtype=np.float64)
    if return_modules:
        module = module.copy()
        module[grey] = -1
        return X, module
    return X


def make_config(name: str, scale: float = 1.0):
    """Expression matrix of a named config; ``scale`` < 1 shrinks the gene count
    (bounded CPU-baseline samples / parity cases), keeping m, F and the seed."""
    m, n, F, seed, net, tom = CONFIGS[name]
    n = max(16, int(round(n * scale)))
    return make_expression(m, n, F=F, seed=seed), net, tom
