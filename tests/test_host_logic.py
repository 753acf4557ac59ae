"""Host-side logic of pywgcna_b200/wgcna.py that needs no GPU: the scale-free fit, the exact mean,
argument validation and error conventions."""
import os
import statistics

import numpy as np
import pandas as pd
import pytest

import pywgcna_b200 as bw
from pywgcna_b200 import wgcna as W


@pytest.mark.parametrize("name", ["k_smooth", "k_gappy", "k_with_zero"])
def test_scale_free_fit_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "scale_free_fit.npz"))
    fit = bw.scaleFreeFitIndex(g[name])
    assert list(fit.columns) == ["Rsquared.SFT", "slope.SFT", "truncatedExponentialAdjRsquared"]
    np.testing.assert_allclose(fit.to_numpy(dtype=float)[0], g[name + "_fit"], rtol=1e-9, atol=1e-10)


def test_scale_free_fit_constant_k():
    fit = bw.scaleFreeFitIndex(np.full(50, 3.0))  # min == max branch of pd.cut
    assert fit.shape == (1, 3)


def test_exact_mean_equals_statistics_mean():
    rng = np.random.default_rng(3)
    for scale in (1.0, 1e-8, 1e6):
        k = rng.gamma(2.0, scale, size=2001)
        assert float(W._exact_mean(k)) == float(statistics.mean(k))


def test_fit_table_matches_scalar_path_and_statistics():
    rng = np.random.default_rng(4)
    K = rng.gamma(2.0, 30.0, size=(6, 3001))
    K[1] *= 1e-9
    K[2, :3] = 0.0
    K[3] = np.concatenate([rng.uniform(0, 1, 2998), [50.0, 51.0, 100.0]])  # leaves bins empty
    K[4] = 2.5  # min == max branch of pd.cut
    K[5, 0] = -1e-17  # nansum(...) - 1 can come out a hair below zero
    r2, slope, r2adj, mean, med, mx = W.fit_table(K)
    for i in range(len(K)):
        ref = W._scale_free_fit(K[i], 10)
        np.testing.assert_allclose([r2[i], slope[i], r2adj[i]], ref, rtol=1e-9, atol=1e-12)
        assert mean[i] == statistics.mean(K[i]) and med[i] == statistics.median(K[i]) and mx[i] == max(K[i])


def test_cal_block_size():
    assert bw.calBlockSize(20000, True, 2 ** 30) == 2236
    assert bw.calBlockSize(23844, True, 2 ** 30) == 1876  # C1
    assert bw.calBlockSize(100, True, 2 ** 30) == 100


def test_error_conventions_before_any_device_work():
    X = pd.DataFrame(np.random.default_rng(0).random((6, 8)))
    with pytest.raises(ValueError):  # list.index on an unknown string, wgcna.py:825
        bw.pickSoftThreshold(X, networkType="nope")
    with pytest.raises(ValueError):  # wgcna.py:1073
        bw.adjacency(X, adjacencyType="nope")
    with pytest.raises(ValueError):  # wgcna.py:1174 / :1177
        bw.TOMsimilarity(np.eye(4), TOMType="nope")
    with pytest.raises(ValueError):
        bw.TOMsimilarity(np.eye(4), TOMDenom="nope")
    with pytest.raises(SystemExit):  # wgcna.py:830-832
        bw.pickSoftThreshold(pd.DataFrame(np.zeros((5, 2))))
    with pytest.raises(SystemExit) as e:  # wgcna.py:1129
        bw.TOMsimilarity(np.zeros(5))
    assert "two-dimensional" in str(e.value)
    with pytest.raises(SystemExit) as e:  # wgcna.py:1132
        bw.TOMsimilarity(np.eye(4, dtype=np.int64))
    assert "not numeric" in str(e.value)
    with pytest.raises(SystemExit) as e:  # wgcna.py:1134
        bw.TOMsimilarity(np.zeros((3, 4)))
    assert "not square" in str(e.value)
    with pytest.raises(NotImplementedError):
        bw.adjacency(X, selectCols=[1, 2])
    with pytest.raises(NotImplementedError):
        bw.pickSoftThreshold(X, weights=X)
    with pytest.raises(SystemExit):
        bw.intramodularConnectivity(np.zeros((3, 4)), ["a"] * 3)
    with pytest.raises(SystemExit):
        bw.softConnectivity(pd.DataFrame(np.zeros((20, 5))))


def test_install_rebinds_static_methods():
    class Dummy:
        @staticmethod
        def pickSoftThreshold(*a, **k):
            return "ref"

        @staticmethod
        def adjacency(*a, **k):
            return "ref"

        @staticmethod
        def TOMsimilarity(*a, **k):
            return "ref"

    bw.install(Dummy)
    assert Dummy.TOMsimilarity is bw.TOMsimilarity and Dummy.adjacency is bw.adjacency
    assert Dummy.pickSoftThreshold is bw.pickSoftThreshold
    bw.uninstall(Dummy)
    assert Dummy.TOMsimilarity() == "ref" and Dummy.adjacency() == "ref"
