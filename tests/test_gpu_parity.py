"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes, pywgcna_b200/wgcna.py),
against (a) golden outputs of the unmodified reference and (b) the numpy oracle on seeded inputs.

Tolerances (BASELINE.json north_star): max-abs <= 1e-6 on correlation / adjacency / TOM and the
identical selected power.  The FP64 kernels are held to a much tighter bound (1e-11); the int8-slice
TOM to 1e-9 (39-bit fixed point, exact accumulation)."""
import os

import numpy as np
import pandas as pd
import pytest

import pywgcna_b200 as bw
from oracle import restated as O
from pywgcna_b200 import _lib
from pywgcna_b200.synth import make_expression


@pytest.fixture(autouse=True, params=["i8", "f64"])
def corr_engine(request):
    """Every test of this module runs on both correlation engines: tcgen05 int8 digit planes (corr_i8.cu, the
    default) and FP64 DMMA (corr.cu)."""
    bw.set_correlation_precision(request.param)
    yield request.param
    bw.set_correlation_precision("auto")

pytestmark = pytest.mark.gpu

NETS = ["unsigned", "signed", "signed hybrid"]
CASES = ["pipeline_small", "pipeline_ragged", "pipeline_edge"]
TOL_NORTH_STAR = 1e-6
TOL_F64 = 1e-11
TOL_TOM = {"f64": 1e-11, "i8": 1e-9}
PRECISIONS = ["f64", "i8"]


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def maxabs(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    with np.errstate(invalid="ignore"):
        return float(np.nanmax(np.abs(a - b))) if a.size else 0.0


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("net", NETS)
def test_golden_soft_threshold(gpu, golden_dir, case, net):
    g = load(golden_dir, case)
    key = net.replace(" ", "_")
    p, sft = bw.pickSoftThreshold(pd.DataFrame(g["X"]), networkType=net, powerVector=list(g["powers"]))
    assert isinstance(p, np.int64) and int(p) == int(g[f"power_{key}"])
    assert list(sft.columns) == O.SFT_COLUMNS and sft.shape == g[f"sft_{key}"].shape
    got, ref = sft.to_numpy(dtype=float), g[f"sft_{key}"]
    assert np.array_equal(got[:, 0], ref[:, 0])
    # connectivity columns: relative 1e-10; fit columns: the 10-point OLS amplifies k's last bits
    np.testing.assert_allclose(got[:, 4:], ref[:, 4:], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(got[:, 1:4], ref[:, 1:4], rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("net", NETS)
def test_golden_adjacency(gpu, golden_dir, case, net):
    g = load(golden_dir, case)
    key = net.replace(" ", "_")
    A = bw.adjacency(pd.DataFrame(g["X"]), adjacencyType=net, power=6)
    assert isinstance(A, np.ndarray) and A.dtype == np.float64 and A.flags.c_contiguous and A.flags.writeable
    assert maxabs(A, g[f"adj_{key}"]) < TOL_F64
    finite = np.nan_to_num(A)
    assert np.array_equal(finite, finite.T), "adjacency must be exactly symmetric (checkAdjMat 1e-12)"


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("case", CASES)
def test_golden_tom(gpu, golden_dir, case, prec):
    g = load(golden_dir, case)
    for net in NETS:
        key = net.replace(" ", "_")
        for tt in ["unsigned", "signed"]:
            for td in ["min", "mean"]:
                A = g[f"adj_{key}"].copy()
                T = bw.TOMsimilarity(A, TOMType=tt, TOMDenom=td, precision=prec)
                assert isinstance(T, pd.DataFrame) and isinstance(T.index, pd.RangeIndex)
                assert maxabs(T.to_numpy(), g[f"tom_{key}_{tt}_{td}"]) < TOL_TOM[prec], (net, tt, td)
                # in-place side effects of the reference on a writable input (wgcna.py:1185, :1143)
                assert not np.isnan(A).any() and np.all(np.diag(A) == 0)


@pytest.mark.parametrize("prec", PRECISIONS)
def test_golden_signed_user_adjacency(gpu, golden_dir, prec):
    g = load(golden_dir, "tom_signed_user")
    for td in ["min", "mean"]:
        # entries of mixed sign: k = plain row sums nearly cancel, denominators get close to 0 and the
        # summation order of L and k shows (|TOM| reaches > 1 here) -> the north-star bound applies
        T = bw.TOMsimilarity(g["A"].copy(), TOMType="signed", TOMDenom=td, precision=prec).to_numpy()
        assert maxabs(T, g[f"tom_signed_{td}"]) < TOL_NORTH_STAR
        T = bw.TOMsimilarity(g["Apos"].copy(), TOMType="unsigned", TOMDenom=td, precision=prec).to_numpy()
        assert maxabs(T, g[f"tom_unsigned_{td}"]) < TOL_TOM[prec]
    # a negative entry is out of range for the unsigned TOM (wgcna.py:1137-1138)
    with pytest.raises(SystemExit):
        bw.TOMsimilarity(g["A"].copy(), TOMType="unsigned")


def test_golden_intramodular(gpu, golden_dir):
    g = load(golden_dir, "tom_signed_user")
    imc = bw.intramodularConnectivity(g["Apos"].copy(), g["imc_colors"])
    assert list(imc.columns) == ["kTotal", "kWithin", "kOut", "kDiff"]
    assert maxabs(imc.to_numpy(), g["imc"]) < 1e-11


def test_check_adjmat(gpu):
    rng = np.random.default_rng(5)
    B = rng.random((77, 77))
    A = (B + B.T) / 2
    bw.checkAdjMat(A)
    bad = A.copy()
    bad[3, 10] += 1e-9
    with pytest.raises(SystemExit) as e:
        bw.checkAdjMat(bad)
    assert "not symmetric" in str(e.value)
    with pytest.raises(SystemExit):
        bw.checkAdjMat(A * 3)
    with pytest.raises(SystemExit):
        bw.TOMsimilarity(bad)
    S = 2 * A - 1  # a similarity in [-1, 1]
    bw.checkSimilarity(S)
    with pytest.raises(SystemExit):
        bw.checkSimilarity(S * 1.5)
    with pytest.raises(SystemExit):
        bw.checkSimilarity(bad)
    nan = bad.copy()
    nan[0, 1] = np.nan  # numpy: max/min are NaN -> both comparisons False -> passes
    bw.checkAdjMat(nan)
    O.checkAdjMat(nan)


@pytest.mark.parametrize("net", NETS)
@pytest.mark.parametrize("m,n", [(50, 1000), (200, 3001), (33, 130)])
def test_oracle_sweep_adjacency(gpu, net, m, n):
    X = make_expression(m, n, F=8, seed=m + n)
    df = pd.DataFrame(X)
    powers = list(range(1, 21))
    pv, datk_ref = O.connectivity_sweep(X, powers, net, faithful=False)
    datk = bw.connectivity(df, powers, net)
    np.testing.assert_allclose(datk, datk_ref, rtol=1e-11, atol=1e-11)
    p, sft = bw.pickSoftThreshold(df, networkType=net, powerVector=powers)
    p_ref = O.select_power(pv, O.sft_table(pv, datk_ref))
    assert int(p) == int(p_ref)
    A = bw.adjacency(df, adjacencyType=net, power=6)
    assert maxabs(A, O.adjacency(df, adjacencyType=net, power=6)) < TOL_F64
    # row blocks (multi-GPU sharding) reproduce the full matrix bit for bit
    r0, nr = n // 3, n // 2
    blk = bw.adjacency(df, adjacencyType=net, power=6, rows=(r0, nr))
    assert np.array_equal(blk, A[r0:r0 + nr])


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("n", [64, 1000, 2500])
def test_oracle_tom(gpu, prec, n):
    X = make_expression(80, n, F=6, seed=n)
    for net, power in [("unsigned", 6), ("signed hybrid", 4), ("signed", 12)]:
        A = O.adjacency(X, adjacencyType=net, power=power)
        for tt, td in [("unsigned", "min"), ("signed", "mean")]:
            T_ref = O.TOMsimilarity(A.copy(), TOMType=tt, TOMDenom=td).to_numpy()
            T = bw.TOMsimilarity(A.copy(), TOMType=tt, TOMDenom=td, precision=prec).to_numpy()
            assert maxabs(T, T_ref) < TOL_TOM[prec], (net, tt, td)
            assert np.all(np.diag(T) == 1.0)
    # row block == slice of the full result; fused dissTOM == (1 - TOM).round(8).
    # np.corrcoef's output is asymmetric in the last bit; the whole-matrix int8 schedule computes the
    # upper triangle and mirrors it (TOM_ji := TOM_ij), a row block computes every entry from its own
    # a_ij -- so the two agree exactly only for an exactly symmetric input.
    As = np.maximum(A, A.T)
    T = bw.TOMsimilarity(As.copy(), precision=prec).to_numpy()
    blk = bw.wgcna._tom(As.copy(), 1, 0, check=False, rows=(n // 4, n // 2), precision=prec)
    assert np.array_equal(blk, T[n // 4:n // 4 + n // 2])
    assert np.array_equal(T, T.T)
    A = As
    D = bw.dissTOM(A.copy(), precision=prec)
    assert np.array_equal(D, (1 - T).round(decimals=8))
    # condensed hand-off == scipy squareform of the rounded dissimilarity (wgcna.py:326)
    from scipy.spatial.distance import squareform

    Dc = bw.dissTOM(A.copy(), precision=prec, condensed=True)
    assert Dc.shape == (n * (n - 1) // 2,) and np.array_equal(Dc, squareform(D, checks=False))


def test_fused_network_matches_two_calls(gpu):
    X = make_expression(60, 900, F=5, seed=9)
    A, T = bw.network(pd.DataFrame(X), power=6, networkType="signed hybrid", TOMType="signed")
    A2 = bw.adjacency(pd.DataFrame(X), adjacencyType="signed hybrid", power=6)
    T2 = bw.TOMsimilarity(A2.copy(), TOMType="signed")
    assert np.array_equal(A, A2) and np.array_equal(T.to_numpy(), T2.to_numpy())


def test_soft_connectivity(gpu):
    X = make_expression(40, 300, F=4, seed=21)
    for net in NETS:
        k = bw.softConnectivity(pd.DataFrame(X), type=net, power=6)
        np.testing.assert_allclose(k, O.softConnectivity(X, type=net, power=6), rtol=1e-11, atol=1e-11)
    Xb = X.copy()
    Xb[:, 7] = 1.0
    assert np.isnan(bw.softConnectivity(pd.DataFrame(Xb))).all()
    assert np.isnan(O.softConnectivity(Xb)).all()


@pytest.mark.parametrize("prec", PRECISIONS)
def test_properties_at_scale(gpu, prec):
    """Size-independent properties at a size the oracle cannot finish quickly."""
    n, m = 6000, 120
    X, mods = make_expression(m, n, F=20, seed=77, return_modules=True)
    A = bw.adjacency(pd.DataFrame(X), adjacencyType="unsigned", power=6)
    assert np.array_equal(A, A.T) and A.min() >= 0 and A.max() <= 1
    assert np.all(np.abs(np.diag(A) - 1) < 1e-12)
    # gene permutation equivariance
    perm = np.random.default_rng(1).permutation(n)
    Ap = bw.adjacency(pd.DataFrame(X[:, perm]), adjacencyType="unsigned", power=6)
    assert np.max(np.abs(Ap - A[np.ix_(perm, perm)])) < 1e-12
    # connectivity == row sums of the adjacency - 1
    k = bw.connectivity(pd.DataFrame(X), [6], "unsigned")[:, 0]
    np.testing.assert_allclose(k, A.sum(axis=1) - 1, rtol=1e-11)
    if True:
        T = bw.TOMsimilarity(A.copy(), TOMType="unsigned", precision=prec).to_numpy()
        assert np.max(np.abs(T - T.T)) < 1e-9 and T.min() >= 0 and T.max() <= 1 + 1e-12
        assert np.all(np.diag(T) == 1)
        # row-sampled check against float64 numpy (the oracle's formula on 40 rows)
        rows = np.random.default_rng(2).choice(n, 40, replace=False)
        A0 = A.copy()
        np.fill_diagonal(A0, 0)
        kk = A0.sum(axis=1)
        Lr = A0[rows] @ A0
        Tr = (Lr + A0[rows]) / (np.minimum(kk[rows][:, None], kk[None, :]) + 1 - A0[rows])
        Tr[np.arange(len(rows)), rows] = 1
        assert np.max(np.abs(T[rows] - Tr)) < TOL_TOM[prec]


@pytest.mark.parametrize("world,n", [(2, 1500), (3, 1500), (4, 2100), (8, 2304)])
def test_multi_gpu_path_emulated_on_one_gpu(gpu, world, n):
    """The N-GPU path with its ranks emulated sequentially on one device (B200_PROFILING.md: ranks that
    wait on one another must not run as separate launches on one GPU; these do not wait -- the exchange
    is done by the test).  Row blocks, the digit-plane exchange layout, the distributed symmetric tile
    deal and the mirrored stores into "peer" blocks must reproduce the single-GPU result bit for bit."""
    import torch

    from pywgcna_b200.device import DeviceNetwork

    X = make_expression(60, n, F=6, seed=n + world)
    powers = list(range(1, 13))
    ref = DeviceNetwork(X, device=0, networkType="signed hybrid")
    k_ref = ref.sweep(powers)
    A_ref = ref.adjacency(6).clone()
    T_ref = ref.tom(TOMType="signed").clone()

    nets = [DeviceNetwork(X, device=0, networkType="signed hybrid", world=world, rank=r, use_dist=False)
            for r in range(world)]
    assert sum(net.nrows for net in nets) == n and all(net.fused for net in nets)
    # sweep: every rank its own block of columns
    k = torch.zeros_like(k_ref)
    for net in nets:
        if net.nrows:
            steps = np.ones(len(powers))
            kloc = torch.empty((len(powers), net.nrows), dtype=torch.float64, device="cuda")
            net.standardize()
            _lib.check(net.lib.b200w_dev_sweep(net.Z.data_ptr(), net.bad.data_ptr(), net.m, net.n, steps.ctypes.data,
                                               len(powers), net.net, net.row0, net.nrows, kloc.data_ptr(), 0,
                                               torch.cuda.current_stream().cuda_stream))
            k[:, net.row0:net.row0 + net.nrows] = kloc
    torch.testing.assert_close(k, k_ref, rtol=1e-12, atol=1e-12)
    # adjacency row blocks
    for net in nets:
        net.adjacency(6)
    A = torch.cat([net.A[:net.nrows] for net in nets])
    assert torch.equal(A, A_ref)
    # TOM: prepare, emulate the all-gather of planes / scale / k, fused symmetric compute with peer stores
    for net in nets:
        net.tom_prepare()
    for dst in nets:
        for src in nets:
            if src is not dst and src.nrows:
                sl = slice(src.row0, src.row0 + src.nrows)
                dst._op[:, sl] = src._op[:, sl]
                dst._k[sl] = src._k[sl]
                dst._scale[sl] = src._scale[sl]
    for net in nets:
        net._out().fill_(float("nan"))
    ptrs = [net._T_raw for net in nets]
    for net in nets:
        net.connect_peers(ptrs)
    for net in nets:
        net.tom_compute(TOMType="signed")
    torch.cuda.synchronize()
    T = torch.cat([net.T[:net.nrows] for net in nets])
    assert not torch.isnan(T).any(), "a tile of the distributed schedule was never written"
    assert torch.equal(T, T_ref)


def test_end_to_end_dendrogram_matches_reference(gpu, golden_dir):
    """End-to-end check through the clustering findModules runs on top of the TOM (wgcna.py:323-328):
    GPU pickSoftThreshold -> adjacency -> TOMsimilarity -> round(1 - TOM, 8) -> scipy linkage(average)
    must give the dendrogram the unmodified reference produced (golden: same merges, same sizes, heights
    to 1e-8), which is the input cutreeHybrid turns into module labels.  (The label comparison itself
    needs the reference's cutreeHybrid, which does not travel to the GPU box: tools/e2e_labels.py.)"""
    from scipy.cluster.hierarchy import linkage
    from scipy.spatial.distance import squareform

    g = load(golden_dir, "e2e_modules")
    df = pd.DataFrame(g["X"])
    p, sft = bw.pickSoftThreshold(df, networkType="signed hybrid", powerVector=list(g["powers"]))
    assert int(p) == int(g["power"])
    np.testing.assert_allclose(sft.to_numpy(dtype=float)[:, 4:], g["sft"][:, 4:], rtol=1e-10, atol=1e-10)
    A = bw.adjacency(df, adjacencyType="signed hybrid", power=p)
    for prec in PRECISIONS:
        T = bw.TOMsimilarity(A.copy(), TOMType="signed", precision=prec)
        assert maxabs(T.to_numpy()[:8], g["tom_sample"]) < TOL_TOM[prec]
        diss = (1 - T).round(decimals=8)
        Z = linkage(squareform(diss.values, checks=False), method="average")
        Zref = g["linkage"]
        assert np.array_equal(Z[:, [0, 1, 3]], Zref[:, [0, 1, 3]]), f"merge order differs ({prec})"
        np.testing.assert_allclose(Z[:, 2], Zref[:, 2], rtol=0, atol=2e-8)
        # the fused epilogue hand-off gives the same rounded dissimilarity
        assert np.array_equal(bw.dissTOM(A.copy(), TOMType="signed", precision=prec), diss.values)


@pytest.mark.parametrize("n", [1001, 4000])
def test_device_sft_statistics(gpu, n):
    """Device histograms / medians / exact sums (b200w_dev_sft_stats) against the host fit on the same k:
    mean, median and max bit-identical to statistics.mean / statistics.median / max, regressions to 1e-9."""
    import statistics

    import torch

    from pywgcna_b200 import wgcna as W
    from pywgcna_b200.device import DeviceNetwork

    X = make_expression(70, n, F=6, seed=n)
    X[:, 3] = 1.5  # a NaN gene: k = 0 for it
    powers = list(range(1, 21))
    net = DeviceNetwork(X, device=0, networkType="signed hybrid")
    power, tab = net.sft(powers)
    k = net.sweep(powers).cpu().numpy()
    r2, slope, r2adj, mean, med, mx = W.fit_table(k, 10)
    assert np.array_equal(tab["mean(k)"], mean) and np.array_equal(tab["median(k)"], med)
    assert np.array_equal(tab["max(k)"], mx)
    for i in (0, 7, 19):
        assert tab["mean(k)"][i] == statistics.mean(k[i]) and tab["median(k)"][i] == statistics.median(k[i])
    np.testing.assert_allclose(tab["SFT.R.sq"], r2, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(tab["slope"], slope, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(tab["truncated R.sq"], r2adj, rtol=1e-9, atol=1e-11)
    p_ref, _ = O.pickSoftThreshold(pd.DataFrame(X), networkType="signed hybrid", powerVector=powers, faithful=False)
    assert int(power) == int(p_ref)
    # what the device decided without the host (b200w_dev_sft_select): the exactly rounded mean bit for bit, the
    # regression R^2 to rounding, and the same selection as wgcna.py:945-953 on the host table
    assert np.array_equal(tab.device_fit[:, 1], mean)
    np.testing.assert_allclose(tab.device_fit[:, 0], r2, rtol=1e-11, atol=1e-13)
    assert tab.device_power == float(tab.host_power) == float(p_ref)
    for cut, mcut in ((0.0, 1e9), (2.0, 100.0), (0.5, 30.0)):  # first branch / argmax branch / both limits active
        pw, tb = net.sft(powers, RsquaredCut=cut, MeanCut=mcut)
        assert float(pw) == float(tb.host_power), (cut, mcut)


@pytest.mark.parametrize("net", ["unsigned", "signed", "signed hybrid"])
@pytest.mark.parametrize("n,power", [(257, 6), (1001, 2.5), (130, 1)])
def test_adjacency_from_sweep_similarity(gpu, net, n, power):
    """Single-GPU chaining: the sweep leaves the transformed similarity in the adjacency buffer and adjacency()
    finishes it elementwise.  Bitwise the same matrix as the correlation-product route (NaN / constant genes,
    ragged sizes, integer and pow() powers), the same connectivities, and the buffer cannot be mistaken for
    an adjacency."""
    from pywgcna_b200.device import DeviceNetwork

    X = make_expression(40, n, F=5, seed=3 * n)
    X[:, 3] = 2.0       # zero variance -> NaN row / column
    X[7, n - 2] = np.nan  # a NaN sample -> NaN row / column
    powers = [1, 2, 3, 5]
    chained = DeviceNetwork(X, device=0, networkType=net, keep_similarity=True)
    plain = DeviceNetwork(X, device=0, networkType=net, keep_similarity=False)
    k1, k0 = chained.sweep(powers).cpu().numpy(), plain.sweep(powers).cpu().numpy()
    assert np.array_equal(k1, k0)
    with pytest.raises(RuntimeError):
        chained.tom_prepare()
    A1 = chained.adjacency(power).cpu().numpy()
    A0 = plain.adjacency(power).cpu().numpy()
    assert np.isnan(A0[3]).all() and np.isnan(A0[:, n - 2]).all()
    assert np.array_equal(A1, A0, equal_nan=True)
    assert maxabs(A0, O.adjacency(pd.DataFrame(X), adjacencyType=net, power=power)) < TOL_F64
    # a second adjacency() without a sweep in between takes the product route again
    assert np.array_equal(chained.adjacency(2).cpu().numpy(), plain.adjacency(2).cpu().numpy(), equal_nan=True)
    # sharded path: a rank's column-block sweep leaves its row block (transposed store)
    for rank in range(2):
        part = DeviceNetwork(X, device=0, networkType=net, world=2, rank=rank, use_dist=False, keep_similarity=True)
        kT = part.sweep(powers).cpu().numpy()
        rows = slice(part.row0, part.row0 + part.nrows)
        # (the whole-matrix sweep takes the symmetric tile schedule: same sums in another order)
        np.testing.assert_allclose(kT[:, rows], k0[:, rows], rtol=1e-12, atol=1e-12)
        if part.nrows:
            assert np.array_equal(part.adjacency(power).cpu().numpy(), A0[rows], equal_nan=True)


@pytest.mark.parametrize("n", [700, 2500])
def test_int8_fast_precision(gpu, n):
    """precision="i8-fast" (B200W_PREC_I8_FAST: the 10 digit-pair products of weight >= 5 instead of 15): well
    inside the north-star tolerance of 1e-6 on TOM, exactly symmetric, and reachable through every single-GPU
    entry point (host API, fused network call, device-resident chaining)."""
    from pywgcna_b200.device import DeviceNetwork

    X = make_expression(50, n, F=6, seed=17 * n)
    df = pd.DataFrame(X)
    A = bw.adjacency(df, adjacencyType="signed hybrid", power=6)
    T0 = O.TOMsimilarity(A.copy(), TOMType="signed").to_numpy()
    Tf = bw.TOMsimilarity(A.copy(), TOMType="signed", precision="i8-fast").to_numpy()
    err_fast, err_i8 = maxabs(Tf, T0), maxabs(bw.TOMsimilarity(A.copy(), TOMType="signed", precision="i8").to_numpy(), T0)
    print(f"n={n}: max-abs error i8 {err_i8:.2e}, i8-fast {err_fast:.2e}")
    assert err_fast < 1e-7 and err_i8 < TOL_TOM["i8"]
    assert np.array_equal(Tf, Tf.T)
    _, Tn = bw.network(df, power=6, networkType="signed hybrid", TOMType="signed", precision="i8-fast")
    assert np.array_equal(Tn.to_numpy(), Tf)
    net = DeviceNetwork(X, device=0, networkType="signed hybrid", precision="i8-fast")
    net.adjacency(6)
    assert np.array_equal(net.tom(TOMType="signed").cpu().numpy(), Tf)


@pytest.mark.parametrize("n,m", [(3, 4), (5, 9), (17, 4), (127, 9), (129, 4), (255, 9), (257, 4)])
def test_tiny_and_ragged_sizes(gpu, n, m):
    """Sizes below / straddling every tile (128-row DMMA tiles, 256 x 256 tcgen05 tiles, 128-byte K blocks)."""
    rng = np.random.default_rng(n * 31 + m)
    df = pd.DataFrame(rng.standard_normal((m, n)))
    p, sft = bw.pickSoftThreshold(df, networkType="unsigned", powerVector=[1, 2, 3, 5])
    p0, sft0 = O.pickSoftThreshold(df, networkType="unsigned", powerVector=[1, 2, 3, 5], faithful=False)
    assert int(p) == int(p0)
    np.testing.assert_allclose(sft.to_numpy(float)[:, 4:], sft0.to_numpy(float)[:, 4:], rtol=1e-10, atol=1e-10)
    A = bw.adjacency(df, power=2)
    A0 = O.adjacency(df, power=2)
    assert maxabs(A, A0) < TOL_F64
    T0 = O.TOMsimilarity(A0.copy(), TOMType="unsigned").to_numpy()
    for prec in PRECISIONS:
        assert maxabs(bw.TOMsimilarity(A.copy(), TOMType="unsigned", precision=prec).to_numpy(), T0) < TOL_TOM[prec]


def test_signed_kme(gpu):
    """cor(gene, eigengene): the n x M block of np.corrcoef(datExpr.T, datME.T) (wgcna.py:3486-3489)."""
    rng = np.random.default_rng(12)
    X = make_expression(45, 700, F=5, seed=12)
    X[:, 9] = 3.0  # constant gene -> NaN row
    ME = pd.DataFrame(rng.standard_normal((45, 7)), columns=[f"ME{c}" for c in "abcdefg"])
    df = pd.DataFrame(X, columns=[f"g{i}" for i in range(700)])
    kme = bw.signedKME(df, ME)
    with np.errstate(all="ignore"):
        ref = np.corrcoef(X.T, ME.to_numpy().T)[:700, 700:]
    assert list(kme.columns) == ["k" + c for c in ME.columns] and list(kme.index) == list(df.columns)
    assert maxabs(kme.to_numpy(), ref) < 1e-12


@pytest.mark.parametrize("prec", PRECISIONS)
def test_tom_degenerate_sizes(gpu, prec):
    for A in (np.array([[1.0]]), np.array([[1.0, 0.25], [0.25, 1.0]]), np.zeros((4, 4))):
        T = bw.TOMsimilarity(A.copy(), TOMType="unsigned", precision=prec).to_numpy()
        T0 = O.TOMsimilarity(A.copy(), TOMType="unsigned").to_numpy()
        assert maxabs(T, T0) < 1e-15


@pytest.mark.parametrize("env", [{"B200W_I8_CTAS": "1"}, {"B200W_I8_SYMMETRIC": "0"}, {"B200W_I8_FIN": "2"},
                                 {"B200W_I8_SYNC": "0"}, {"B200W_I8_NOCOOP": "1"}, {"B200W_I8_WMIN": "3"}])
def test_int8_kernel_variants(gpu, env, monkeypatch):
    """The schedule knobs of the tcgen05 kernel (1-CTA tiles, full instead of symmetric schedule, two
    finalizer warps, no producer barrier, non-cooperative launch as used under ncu, one more digit-pair
    class) change speed, never the result beyond its tolerance."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    n = 777
    X = make_expression(50, n, F=5, seed=n)
    A = O.adjacency(X, adjacencyType="unsigned", power=6)
    T_ref = O.TOMsimilarity(A.copy(), TOMType="unsigned").to_numpy()
    T = bw.TOMsimilarity(A.copy(), TOMType="unsigned", precision="i8").to_numpy()
    assert maxabs(T, T_ref) < TOL_TOM["i8"]
