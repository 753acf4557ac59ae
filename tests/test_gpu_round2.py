"""GPU parity tests added in round 2 (all through the C ABI, against goldens of the unmodified reference or
the restated oracle): TomSimilarityFromAdj on unchecked input, full-size C2 soft-threshold table, several
devices driven from one process."""
import os
import threading

import numpy as np
import pandas as pd
import pytest

import pywgcna_b200 as bw
from oracle import restated as O
from pywgcna_b200 import _lib
from pywgcna_b200.synth import CONFIGS, make_expression

pytestmark = pytest.mark.gpu


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


# ------------------------------------------------------------------------------------------------
# TomSimilarityFromAdj, wgcna.py:1141-1157: no checks and no nan_to_num
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["f64", "i8"])
@pytest.mark.parametrize("nm", ["asym", "sym_nan", "asym_nan", "signed_asym"])
def test_tom_from_adj_unchecked(gpu, golden_dir, nm, prec, monkeypatch):
    """Public name, reference goldens: an asymmetric matrix is contracted as A @ A with column sums k_j,
    a NaN makes its whole row and column NaN (diagonal 1); tolerance 1e-10 (north-star: 1e-6)."""
    monkeypatch.setattr(bw.wgcna, "PRECISION", prec)
    g = load(golden_dir, "tom_from_adj")
    for tt, td in ((0, 0), (1, 1)):
        A = g[f"A_{nm}"].copy()
        T = bw.TomSimilarityFromAdj(A, td, tt)
        ref = g[f"fromadj_{nm}_{tt}_{td}"]
        assert np.array_equal(np.isnan(T), np.isnan(ref))
        # signed_asym: signed row sums put denominators min(k_i, k_j) + 1 - |a| near zero, so entries reach
        # 1e3 and more and the comparison has to be relative there
        np.testing.assert_allclose(T, ref, rtol=1e-7 if nm == "signed_asym" else 0, atol=1e-10, equal_nan=True)
        # the in-place side effect of the reference: diagonal zeroed (:1143), NaNs left alone
        assert np.all(np.diag(A) == 0)
        assert np.isnan(A).sum() == np.isnan(g[f"A_{nm}"]).sum() - np.isnan(np.diag(g[f"A_{nm}"])).sum()


@pytest.mark.parametrize("prec", ["f64", "i8"])
@pytest.mark.parametrize("nm", ["sym_nan", "asym_nan"])
def test_tomsimilarity_on_matrix_waved_through_by_nan(gpu, golden_dir, nm, prec):
    """TOMsimilarity: a NaN makes both comparisons of checkAdjMat False (wgcna.py:1135-1138), so even an
    asymmetric matrix reaches nan_to_num (:1185) and the contraction."""
    g = load(golden_dir, "tom_from_adj")
    for tt in ("unsigned", "signed"):
        A = g[f"A_{nm}"].copy()
        T = bw.TOMsimilarity(A, TOMType=tt, precision=prec).to_numpy()
        np.testing.assert_allclose(T, g[f"tomsim_{nm}_{tt}"], rtol=0, atol=1e-10)
        assert not np.isnan(A).any() and np.all(np.diag(A) == 0)  # :1185 and :1143, in place


def test_tom_from_adj_symmetric_matches_tomsimilarity(gpu):
    """For a clean symmetric adjacency the unchecked and the checked entry give the same bits."""
    X = make_expression(40, 517, F=4, seed=11)
    A = bw.adjacency(pd.DataFrame(X), adjacencyType="signed hybrid", power=6)
    T0 = bw.TOMsimilarity(A.copy(), TOMType="signed").to_numpy()
    T1 = bw.TomSimilarityFromAdj(A.copy(), 0, 1)
    assert np.array_equal(T0, T1)


# ------------------------------------------------------------------------------------------------
# the tcgen05 correlation engine (corr_i8.cu) against np.corrcoef itself
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n", [(200, 1500), (37, 131), (500, 2049), (5000, 700), (3, 300)])
def test_int8_correlation_matches_corrcoef(gpu, m, n):
    """unsigned adjacency with power 1 is |np.corrcoef| (wgcna.py:1097-1111): the int8 digit-plane product is
    within 1e-12 of numpy float64 (north-star tolerance 1e-6), bitwise symmetric, and both engines agree."""
    X = make_expression(m, n, F=5, seed=m * 1000 + n)
    X[:, 1] = X[:, 0]          # duplicate genes: cor = 1 exactly (clipped)
    X[:, 2] = -3.0 * X[:, 0]   # cor = -1
    X[:, 3] *= 1e-6            # tiny scale: the per-gene exponent carries it
    X[0, 4] += 1e3             # one dominant sample: the largest digit range in one entry
    ref = np.abs(np.corrcoef(X, rowvar=False))
    out = {}
    for eng in ("i8", "f64"):
        bw.set_correlation_precision(eng)
        try:
            out[eng] = bw.adjacency(pd.DataFrame(X), adjacencyType="unsigned", power=1)
        finally:
            bw.set_correlation_precision("auto")
        assert np.array_equal(out[eng], out[eng].T)
        np.testing.assert_allclose(out[eng], ref, rtol=0, atol=1e-12)
    np.testing.assert_allclose(out["i8"], out["f64"], rtol=0, atol=1e-12)
    # signed network, non-integer power: the pow() branch of the fused epilogue
    for eng in ("i8", "f64"):
        bw.set_correlation_precision(eng)
        try:
            A = bw.adjacency(pd.DataFrame(X), adjacencyType="signed", power=2.5)
        finally:
            bw.set_correlation_precision("auto")
        np.testing.assert_allclose(A, ((1 + np.corrcoef(X, rowvar=False)) / 2) ** 2.5, rtol=0, atol=1e-11)


# ------------------------------------------------------------------------------------------------
# full-size C2: the reference's own soft-threshold table and selected power (wgcna.py:924-953)
# ------------------------------------------------------------------------------------------------
def test_c2_fullsize_soft_threshold_table(gpu, golden_dir):
    """20 000 x 200, signed hybrid, powers 1..20: identical selected power, and every column of the 20 x 7
    table of the unmodified reference (tests/golden/c2_fullsize.npz, oracle/gen_golden_large.py)."""
    g = load(golden_dir, "c2_fullsize")
    m, n, F, seed, net, tom = CONFIGS["C2"]
    X = make_expression(m, n, F=F, seed=seed)
    df = pd.DataFrame(X)
    p, sft = bw.pickSoftThreshold(df, networkType=net, powerVector=list(g["powers"]))
    assert int(p) == int(g["power"])
    got, ref = sft.to_numpy(dtype=float), g["sft"]
    assert np.array_equal(got[:, 0], ref[:, 0])
    np.testing.assert_allclose(got[:, 1:4], ref[:, 1:4], rtol=0, atol=1e-9)   # R^2, slope, truncated R^2
    # mean, median, max of k (k = sum - 1: the absolute error of a sum of 20 000 terms stays when k is tiny)
    np.testing.assert_allclose(got[:, 4:], ref[:, 4:], rtol=1e-11, atol=1e-12)
    # sampled rows of the reference's adjacency and TOM at that power
    rows = g["rows"]
    A = bw.adjacency(df, adjacencyType=net, power=p)
    np.testing.assert_allclose(A[rows], g["adj_rows"], rtol=0, atol=1e-12)
    T = bw.TOMsimilarity(A, TOMType=tom).to_numpy()
    np.testing.assert_allclose(T[rows], g["tom_rows"], rtol=0, atol=1e-10)


# ------------------------------------------------------------------------------------------------
# one host process, several devices / several host threads (SURVEY 8b threading contract)
# ------------------------------------------------------------------------------------------------
def test_two_host_threads_one_device(gpu):
    """Two Python threads inside b200w_tom at once (ctypes releases the GIL): per-call parameters and
    per-thread progress buffers, same bits as the serial calls."""
    X = make_expression(50, 2304, F=5, seed=21)
    A = bw.adjacency(pd.DataFrame(X), adjacencyType="unsigned", power=4)
    ref = bw.TOMsimilarity(A.copy(), TOMType="unsigned").to_numpy().copy()
    out, err = [None, None], []

    def work(i):
        try:
            for _ in range(3):
                out[i] = bw.TOMsimilarity(A.copy(), TOMType="unsigned").to_numpy().copy()
        except Exception as e:  # pragma: no cover
            err.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not err, err
    assert np.array_equal(out[0], ref) and np.array_equal(out[1], ref)


def test_second_device_same_process(gpu):
    """b200w_tom on device 0 and then device 1 from one process (progress flags, tile lists and caches are
    keyed by device).  Skipped on a one-GPU box."""
    if _lib.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    X = make_expression(50, 2304, F=5, seed=22)
    A = bw.adjacency(pd.DataFrame(X), adjacencyType="unsigned", power=4, device=0)
    T0 = bw.TOMsimilarity(A.copy(), TOMType="unsigned", device=0).to_numpy().copy()
    T1 = bw.TOMsimilarity(A.copy(), TOMType="unsigned", device=1).to_numpy().copy()
    T0b = bw.TOMsimilarity(A.copy(), TOMType="unsigned", device=0).to_numpy().copy()
    assert np.array_equal(T0, T1) and np.array_equal(T0, T0b)
    ref = O.TOMsimilarity(A.copy(), TOMType="unsigned").to_numpy()
    np.testing.assert_allclose(T1, ref, rtol=0, atol=1e-10)


# ------------------------------------------------------------------------------------------------
# the real multi-rank path on hardware: NCCL + CUDA IPC peer stores + gated copy-engine plane pulls
# ------------------------------------------------------------------------------------------------
def _free_port():
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks_real_nccl_and_ipc(gpu, world):
    """`world` processes, one per GPU (tests/dist_worker.py): the union of the ranks' TOM blocks is bit-identical
    to the single-GPU result and within 1e-10 of the float64 oracle, for both exchange modes of the TOM operand, both
    schedules of the sharded sweep (symmetric with peer stores / column blocks) and for sizes that leave the last rank
    short (600) or empty (200).  Skipped when fewer GPUs are visible."""
    import json
    import subprocess
    import sys

    if _lib.device_count() < world:
        pytest.skip(f"needs {world} visible GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(root, "tests", "dist_worker.py"), "2304,2500,600,200"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    rep = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert rep["ok"] and len(rep["cases"]) == 16  # 4 sizes x (pull, gather) x (symmetric, full)
    assert {(c["mode"], c["sweep"]) for c in rep["cases"]} == {(e, s) for e in ("pull", "gather")
                                                               for s in ("symmetric", "full")}
    for c in rep["cases"]:
        assert c["tom_bitwise"] and c["adj_bitwise"] and c["power_equal"], c


# ------------------------------------------------------------------------------------------------
# average linkage on the GPU (findModules, wgcna.py:326-328) with scipy's exact result
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,dec", [(2, 8), (3, 8), (300, 2), (1000, 1), (1537, 8), (4000, 3)])
def test_linkage_average_matches_scipy(gpu, n, dec):
    """The whole linkage matrix -- merge order, heights, sizes -- bit for bit, also when rounding makes most
    distances tie (dec = 1: ten distinct values), against scipy itself and the numpy restatement."""
    from scipy.cluster.hierarchy import linkage as scipy_linkage
    from scipy.spatial.distance import squareform

    rng = np.random.default_rng(n * 10 + dec)
    B = rng.random((n, n))
    D = np.round((B + B.T) / 2, dec)
    np.fill_diagonal(D, 0)
    y = squareform(D, checks=False)
    Z = bw.linkage(y, method="average")
    Zs = scipy_linkage(y, method="average")
    assert np.array_equal(Z, Zs)
    if n <= 1000:
        assert np.array_equal(Z, O.linkage_average(y))


def test_linkage_on_gpu_dissimilarity_matches_reference_tree(gpu, golden_dir):
    """End of the accelerated chain on the e2e fixture: GPU power -> adjacency -> round(1 - TOM, 8) condensed ->
    GPU linkage gives the gene tree the unmodified reference built (merge order exact, heights to 2e-8)."""
    g = load(golden_dir, "e2e_modules")
    df = pd.DataFrame(g["X"])
    p, _ = bw.pickSoftThreshold(df, networkType="signed hybrid", powerVector=list(g["powers"]))
    A = bw.adjacency(df, adjacencyType="signed hybrid", power=p)
    y = bw.dissTOM(A, TOMType="signed", condensed=True)
    Z = bw.linkage(y, method="average")
    from scipy.cluster.hierarchy import linkage as scipy_linkage

    assert np.array_equal(Z, scipy_linkage(y, method="average"))
    assert np.array_equal(Z[:, [0, 1, 3]], g["linkage"][:, [0, 1, 3]])
    np.testing.assert_allclose(Z[:, 2], g["linkage"][:, 2], rtol=0, atol=2e-8)


# ------------------------------------------------------------------------------------------------
# C1 scale (192 x 23 844, the stand-in for the 5xFAD tutorial matrix): the reference's own results
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("net", ["signed hybrid", "unsigned"])
def test_c1scale_power_tom_and_gene_tree(gpu, golden_dir, net):
    """The accelerated part of findModules (wgcna.py:278-328) at the tutorial's size against what the UNMODIFIED
    reference produced (tests/golden/c1scale_*.npz): selected power, soft-threshold table, sampled TOM rows, and
    the gene tree -- merge order of all 23 843 merges identical, heights to 2e-8.  (cutreeHybrid on that tree takes
    minutes of pure Python; the full chain through the real WGCNA.findModules runs in
    test_c1scale_findmodules_labels / tools/e2e_c1scale.py.)"""
    g = load(golden_dir, "c1scale_" + net.replace(" ", "_"))
    m, n, F, seed, _, _ = CONFIGS["C1"]
    tom = "signed" if net == "signed hybrid" else "unsigned"
    X = make_expression(m, n, F=F, seed=seed)
    df = pd.DataFrame(X)
    p, sft = bw.pickSoftThreshold(df, networkType=net, powerVector=list(g["powers"]))
    assert int(p) == int(g["power"])
    got = sft.to_numpy(dtype=float)
    np.testing.assert_allclose(got[:, 1:4], g["sft"][:, 1:4], rtol=0, atol=1e-9)
    np.testing.assert_allclose(got[:, 4:], g["sft"][:, 4:], rtol=1e-11, atol=1e-12)
    A = bw.adjacency(df, adjacencyType=net, power=p)
    T = bw.TOMsimilarity(A, TOMType=tom).to_numpy()
    np.testing.assert_allclose(T[g["rows"]], g["tom_rows"], rtol=0, atol=1e-10)
    del T
    y = bw.dissTOM(A, TOMType=tom, condensed=True)  # round(1 - TOM, 8), squareform order (wgcna.py:323-326)
    del A
    Z = bw.linkage(y, method="average")              # wgcna.py:328 on the GPU
    assert np.array_equal(Z[:, 0].astype(np.int32), g["link_a"])
    assert np.array_equal(Z[:, 1].astype(np.int32), g["link_b"])
    assert np.array_equal(Z[:, 3].astype(np.int32), g["link_n"])
    np.testing.assert_allclose(Z[:, 2], g["link_h"], rtol=0, atol=2e-8)


@pytest.mark.skipif(not os.environ.get("B200W_LONG_TESTS"), reason="~10 min of reference host code; set B200W_LONG_TESTS=1")
def test_c1scale_findmodules_labels(gpu):
    """pywgcna_b200.install() on the REAL PyWGCNA.wgcna.WGCNA (baseline/_ref), then its own findModules on the
    C1-scale matrix for both network types and both TOM arithmetics: identical power, merge order, cutreeHybrid
    labels and merged module labels (tools/e2e_c1scale.py; the outcome of the last run is committed as
    profiles/r02_e2e_c1scale.json)."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out", "e2e_c1scale_test.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "e2e_c1scale.py"), out], capture_output=True,
                       text=True, timeout=3000, cwd=root)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    rep = json.load(open(out))
    assert rep["ok"] and len(rep["results"]) == 4


# ------------------------------------------------------------------------------------------------
# the one-call host entry and the device-resident gene tree
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("net,tom", [("signed hybrid", "signed"), ("unsigned", "unsigned")])
def test_network_blocks_equals_three_calls(gpu, net, tom):
    """pywgcna_b200.network_blocks (expression matrix in, TOM out, chain resident on the device, power selected on
    the device) gives the reference-shaped three calls' results: same power, same table, bitwise the same adjacency
    and TOM; a second call reuses the cached workspace."""
    X = make_expression(64, 1700, F=6, seed=31)
    X[:, 11] = 0.5  # a constant gene
    df = pd.DataFrame(X)
    powers = list(range(1, 11)) + list(range(11, 21, 2))
    p, sft = bw.pickSoftThreshold(df, networkType=net, powerVector=powers)
    A = bw.adjacency(df, adjacencyType=net, power=p)
    T = bw.TOMsimilarity(A.copy(), TOMType=tom).to_numpy()
    for _ in range(2):
        res = bw.network_blocks(df, powerVector=powers, networkType=net, TOMType=tom, return_adjacency=True)
        assert res["rows"] == (0, 1700) and int(res["power"]) == int(p)
        np.testing.assert_allclose(res["sft"].to_numpy(dtype=float), sft.to_numpy(dtype=float), rtol=1e-12, atol=1e-13)
        assert np.array_equal(res["adjacency"], A, equal_nan=True)
        assert np.array_equal(res["TOM"], T)
    from pywgcna_b200.device import release_workspace

    release_workspace()


def test_device_gene_tree_equals_scipy_on_host_dissimilarity(gpu):
    """DeviceNetwork.gene_tree(): round(1 - TOM, 8) and the average linkage without leaving the device give the
    linkage matrix scipy computes from the host-side dissTOM of the same network (wgcna.py:320-328)."""
    from scipy.cluster.hierarchy import linkage as scipy_linkage
    from scipy.spatial.distance import squareform

    from pywgcna_b200.device import DeviceNetwork

    X = make_expression(60, 1201, F=5, seed=41)
    net = DeviceNetwork(X, device=0, networkType="signed hybrid")
    net.adjacency(6.0)
    Z = net.gene_tree(TOMType="signed")
    A = bw.adjacency(pd.DataFrame(X), adjacencyType="signed hybrid", power=6)
    diss = bw.dissTOM(A, TOMType="signed")
    assert np.array_equal(Z, scipy_linkage(squareform(diss, checks=False), method="average"))


# ------------------------------------------------------------------------------------------------
# install() on the real reference class: the other callers of the rebound methods
# ------------------------------------------------------------------------------------------------
def test_install_on_real_class_hub_genes_and_signed_kme(gpu):
    """The reference's own ``top_n_hub_genes`` (it calls ``WGCNA.adjacency`` by class name, wgcna.py:3748) and
    ``CalculateSignedKME`` (:3452-3490), on a reference WGCNA object, before and after ``pywgcna_b200.install()`` on
    the REAL class out of baseline/_ref: same genes in the same order, connectivity and kME to 1e-10; uninstall()
    restores every attribute, also the module-level ``linkage``."""
    from oracle import findmodules_harness as H
    from oracle import ref_import as R

    if not R.reference_available():
        pytest.skip("no reference tree on this box (baseline/_ref missing)")
    W = R.load_reference()
    X, module = make_expression(80, 900, F=4, seed=77, return_modules=True)
    obj = H.make_wgcna(X, networkType="signed hybrid")
    obj.power = 6
    obj.datExpr.var["moduleColors"] = np.array(["grey", "c0", "c1", "c2", "c3"], dtype=object)[module + 1]
    rng = np.random.default_rng(3)
    obj.datME = pd.DataFrame(rng.standard_normal((80, 3)), index=obj.datExpr.obs.index, columns=["MEa", "MEb", "MEc"])
    before = {k: W.WGCNA.__dict__[k] for k in ("adjacency", "TOMsimilarity", "pickSoftThreshold", "CalculateSignedKME")}
    scipy_linkage = W.linkage
    with np.errstate(all="ignore"):
        hub_ref = W.WGCNA.top_n_hub_genes(obj, "c1", n=25)
        W.WGCNA.CalculateSignedKME(obj)
    kme_ref = obj.signedKME.copy()
    lib = _lib.load()
    bw.install(W.WGCNA)
    try:
        assert W.linkage is bw.linkage
        n0 = lib.b200w_launch_count()
        hub = W.WGCNA.top_n_hub_genes(obj, "c1", n=25)
        obj.CalculateSignedKME()
        assert lib.b200w_launch_count() > n0  # the rebound methods ran on the GPU
    finally:
        bw.uninstall(W.WGCNA)
    assert list(hub.index) == list(hub_ref.index)
    np.testing.assert_allclose(hub["connectivity"].to_numpy(float), hub_ref["connectivity"].to_numpy(float), rtol=0, atol=1e-10)
    assert list(obj.signedKME.columns) == list(kme_ref.columns)
    np.testing.assert_allclose(obj.signedKME.to_numpy(float), kme_ref.to_numpy(float), rtol=0, atol=1e-10)
    for k, v in before.items():
        assert W.WGCNA.__dict__[k] is v
    assert W.linkage is scipy_linkage


def test_gene_major_input_is_accepted_without_copy_and_gives_the_same_bits(gpu):
    """An F-ordered expression matrix (what DataFrame.to_numpy() returns under pandas >= 3) is passed to the *_gm
    entry points as it is; results equal the C-ordered calls bit for bit, including a constant and a NaN gene."""
    X = make_expression(57, 1111, F=5, seed=51)
    X[:, 5] = 2.0
    X[3, 9] = np.nan
    Xc, Xf = np.ascontiguousarray(X), np.asfortranarray(X)
    a, gm = _lib.as_f64_expression(Xf)
    assert gm and np.shares_memory(a, Xf) and not _lib.as_f64_expression(Xc)[1]
    assert _lib.as_f64_expression(pd.DataFrame(Xc))[0].shape in ((57, 1111), (1111, 57))
    powers = list(range(1, 21))
    p0, t0 = bw.pickSoftThreshold(Xc, networkType="signed", powerVector=powers)
    p1, t1 = bw.pickSoftThreshold(Xf, networkType="signed", powerVector=powers)
    assert p0 == p1 and np.array_equal(t0.to_numpy(dtype=float), t1.to_numpy(dtype=float), equal_nan=True)
    A0 = bw.adjacency(Xc, adjacencyType="signed", power=7)
    A1 = bw.adjacency(Xf, adjacencyType="signed", power=7)
    assert np.array_equal(A0, A1, equal_nan=True)
    r0 = bw.network_blocks(Xc, powerVector=powers, networkType="signed", TOMType="signed")
    T0 = r0["TOM"].copy()
    r1 = bw.network_blocks(Xf, powerVector=powers, networkType="signed", TOMType="signed")
    assert r0["power"] == r1["power"] and np.array_equal(T0, r1["TOM"])
    from pywgcna_b200.device import release_workspace

    release_workspace()
