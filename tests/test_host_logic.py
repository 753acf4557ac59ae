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


def _emulate_device_stats(K, nBreaks):
    """numpy stand-in for b200w_dev_sft_stats: per power min / max / median / plain sum / pd.cut counts and bin sums,
    and the exact mantissa sums per frexp exponent as the (lo, hi) pairs the kernel hands over (value lo + hi * 2^27,
    however the kernel splits the 53-bit mantissas internally)."""
    P, n = K.shape
    E, bias = 2304, 1100
    stats = np.zeros((P, 4 + 2 * nBreaks))
    buck = np.zeros((P, E, 2), dtype=np.int64)
    for p in range(P):
        k = K[p]
        mn, mx = k.min(), k.max()
        edges = W._cut_edges(float(mn), float(mx), nBreaks)
        b = np.searchsorted(edges, k, side="left") - 1  # right-closed bins
        for i in range(nBreaks):
            stats[p, 4 + i] = (b == i).sum()
            stats[p, 4 + nBreaks + i] = k[b == i].sum()
        stats[p, :4] = mn, mx, statistics.median(k.tolist()), k.sum()
        nz = k[(k != 0) & np.isfinite(k)]
        m, e = np.frexp(nz)
        M = (m * 2.0 ** 53).astype(np.int64)  # exact
        # a different split than two 27-bit halves, to pin the (lo, hi) convention rather than one split
        pieces = [np.sign(M) * ((np.abs(M) >> (11 * q)) & 0x7FF) for q in range(5)]
        sums = [np.zeros(E, dtype=np.int64) for _ in range(5)]
        for q in range(5):
            np.add.at(sums[q], e + bias, pieces[q])
        buck[p, :, 0] = sums[0] + (sums[1] << 11) + (sums[2] << 22)
        buck[p, :, 1] = (sums[3] << 6) + (sums[4] << 17)
    return stats, buck


def test_fit_table_from_device_statistics_layout():
    """Host half of the device-resident pickSoftThreshold (DeviceNetwork.sft / b200w_soft_threshold_stats): the
    table assembled from per-power statistics equals the one computed from k itself -- mean, median, max bit for
    bit (statistics.mean is the correctly rounded exact mean), regressions to rounding."""
    rng = np.random.default_rng(11)
    n, nb = 3001, 10
    K = np.abs(rng.standard_normal((6, n))) * np.logspace(1.5, -7, 6)[:, None]
    K[2, 5] = 0.0  # zeros are skipped by the mantissa sums and sit in the lowest bin
    K[4] = np.round(K[4], 3) + 1e-3  # ties and an even spread over few distinct values
    stats, buck = _emulate_device_stats(K, nb)
    r2, slope, r2adj, mean, med, mx = W.fit_table_from_stats(stats, buck, n, nb)
    r2_0, slope_0, r2adj_0, mean_0, med_0, mx_0 = W.fit_table(K, nb)
    assert np.array_equal(mean, mean_0) and np.array_equal(med, med_0) and np.array_equal(mx, mx_0)
    for i in range(K.shape[0]):
        assert mean[i] == statistics.mean(K[i].tolist())
    np.testing.assert_allclose(r2, r2_0, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(slope, slope_0, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(r2adj, r2adj_0, rtol=1e-9, atol=1e-11)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) on the smallest workload."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "TFLOP/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["full_workload"] is True and line["steps"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("tiny") or "tiny" in json.dumps(line["config"])


def test_install_rebinds_the_real_reference_class():
    """install() / uninstall() on the REAL ``PyWGCNA.wgcna.WGCNA`` (imported through oracle/ref_import.py where a
    reference tree exists): the names findModules looks up on the class (wgcna.py:278, :310, :320), the helpers, the
    instance method CalculateSignedKME and the module-level ``linkage`` of :328 are rebound and restored."""
    import pytest

    from oracle import ref_import as R

    if not R.reference_available():
        pytest.skip("no reference tree here")
    Wm = R.load_reference()
    names = ("pickSoftThreshold", "adjacency", "TOMsimilarity", "TomSimilarityFromAdj", "checkAdjMat",
             "softConnectivity", "intramodularConnectivity", "scaleFreeFitIndex", "calBlockSize", "checkSimilarity",
             "CalculateSignedKME")
    before = {k: Wm.WGCNA.__dict__[k] for k in names}
    link = Wm.linkage
    bw.install(Wm.WGCNA)
    try:
        for k in ("pickSoftThreshold", "adjacency", "TOMsimilarity"):
            assert Wm.WGCNA.__dict__[k].__func__ is getattr(bw, k)
            assert getattr(Wm.WGCNA, k) is getattr(bw, k)  # what `WGCNA.<name>(...)` resolves to
        assert Wm.WGCNA.CalculateSignedKME is W.CalculateSignedKME
        assert Wm.linkage is bw.linkage
        bw.install(Wm.WGCNA)  # idempotent: the originals are remembered once
    finally:
        bw.uninstall(Wm.WGCNA)
    for k, v in before.items():
        assert Wm.WGCNA.__dict__[k] is v
    assert Wm.linkage is link
    bw.install(Wm.WGCNA, accelerate_linkage=False)
    try:
        assert Wm.linkage is link
    finally:
        bw.uninstall(Wm.WGCNA)


@pytest.mark.parametrize("n,world", [(2304, 2), (2500, 4), (600, 2), (200, 4), (5000, 8), (3001, 3)])
def test_sharded_symmetric_sweep_schedule_covers_every_block_pair_once(n, world):
    """Tile lists of b200w_dev_sweep_dist (corr.cu): over the ranks, every unordered pair of 128-gene blocks is formed
    exactly once -- as (rows = one block, both 64-column tiles of the other block) on the rank that owns the column
    block -- and every diagonal block once on its owner; the tiles are balanced over the ranks that have genes."""
    from pywgcna_b200 import _lib as L
    from pywgcna_b200.device import ALIGN, rows_per_rank

    lib = L.load()
    per = rows_per_rank(n, world, ALIGN)
    B = -(-n // 128)
    seen = {}
    counts = []
    for rank in range(world):
        cnt = lib.b200w_sweep_dist_tiles(n, world, rank, per, None, 0)
        assert cnt >= 0
        xy = np.zeros((max(cnt, 1), 2), dtype=np.int32)
        assert lib.b200w_sweep_dist_tiles(n, world, rank, per, xy.ctypes.data, cnt) == cnt
        counts.append(cnt)
        col0 = rank * per
        last, restarts = (-1, -1), 0
        for a, jt in xy[:cnt]:
            if (jt, a) <= last:  # column-major, rows ascending -- one run per launch (a peer's rows each, then the own rows)
                restarts += 1
            last = (jt, a)
        assert restarts <= world - 1
        for a, jt in xy[:cnt]:
            j0 = col0 + int(jt) * 64
            assert j0 < min(n, col0 + per) and 0 <= a < B
            key = (int(a), j0 // 64)  # (row block, global 64-column tile)
            assert key not in seen
            seen[key] = rank
    for a in range(B):
        for b in range(a, B):
            tiles_b = [c for c in (2 * b, 2 * b + 1) if c * 64 < n]  # a as rows, b's column tiles
            tiles_a = [c for c in (2 * a, 2 * a + 1) if c * 64 < n]
            got_ab = [(a, c) in seen for c in tiles_b]
            got_ba = [(b, c) in seen for c in tiles_a]
            if a == b:
                assert all(got_ab)
            else:  # exactly one orientation, complete
                assert (all(got_ab) and not any(got_ba)) or (all(got_ba) and not any(got_ab)), (a, b)
    busy = [c for c, r in zip(counts, range(world)) if r * per + per <= n]  # ranks with a full block
    if len(busy) > 1:
        assert max(busy) - min(busy) <= 0.1 * max(busy) + 4
    assert lib.b200w_sweep_dist_tiles(n, 1, 0, per, None, 0) == -1


@pytest.mark.parametrize("n,world", [(2304, 2), (2500, 4), (600, 2), (200, 4), (5000, 8), (3001, 3), (50000, 4), (1000, 8)])
def test_sharded_symmetric_sweep_delivers_every_cell_of_every_block_once(n, world):
    """Data flow of b200w_dev_sweep_dist on the host: who writes which 64 x 64 cell of which rank's similarity block.
    A tile (rows = block a, columns = 64-column tile c of the computing rank) is stored transposed into the computing
    rank's own rows and in the direct orientation to the owner of a -- into its block (own rows, or a peer's over NVLink
    with two ranks) or, with more ranks, into the staging rectangle that the copy engines push.  Every cell of every
    rank's block must be written exactly once, and a rectangle must hold exactly the tiles computed for that peer."""
    from pywgcna_b200 import _lib as L
    from pywgcna_b200.device import ALIGN, rows_per_rank

    lib = L.load()
    per = rows_per_rank(n, world, ALIGN)
    n64 = -(-n // 64)
    written = np.zeros((n64, n64), dtype=np.int32)  # global (row tile, column tile) of the n x n similarity
    staged = None
    for rank in range(world):
        cnt = lib.b200w_sweep_dist_tiles(n, world, rank, per, None, 0)
        xy = np.zeros((max(cnt, 1), 2), dtype=np.int32)
        lib.b200w_sweep_dist_tiles(n, world, rank, per, xy.ctypes.data, cnt)
        reg = np.zeros((world, 4), dtype=np.int64)
        mode = lib.b200w_sweep_dist_regions(n, world, rank, per, reg.ctypes.data)
        assert mode in (0, 1) and (staged is None or staged == mode)
        staged = mode
        assert mode == (1 if world > 2 else 0)
        owed = {o: set() for o in range(world)}
        for a, jt in xy[:cnt]:
            a, c = int(a), (rank * per) // 64 + int(jt)
            rows_a = [t for t in (2 * a, 2 * a + 1) if t * 64 < n]
            for t in rows_a:  # transposed: rows = the column tile's genes, columns = block a
                written[c, t] += 1
            if a != c // 2:  # off the diagonal block: the direct orientation goes to the owner of a
                owner = (a * 128) // per
                for t in rows_a:
                    written[t, c] += 1
                    if owner != rank:
                        owed[owner].add((t, c))
        for o in range(world):
            r0, r1, c0, c1 = (int(v) for v in reg[o])
            if not mode or o == rank:
                assert (r0, r1, c0, c1) == (0, 0, 0, 0)
                continue
            cells = {(t, c) for t in range(r0 // 64, -(-r1 // 64)) for c in range(c0 // 64, -(-c1 // 64))}
            assert cells == owed[o], (rank, o, len(cells), len(owed[o]))
            if cells:  # inside o's rows and this rank's columns, starting on tile boundaries
                assert r0 == o * per and r1 <= min(n, (o + 1) * per) and r0 % 128 == 0 and c0 % 128 == 0
                assert rank * per <= c0 and c1 == min(n, (rank + 1) * per)
    assert written.min() == 1 and written.max() == 1
