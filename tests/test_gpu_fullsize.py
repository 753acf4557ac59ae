"""BASELINE.json configs at FULL size, device resident (pywgcna_b200/device.py -> dev_* C ABI), checked
through row-sampled float64 numpy evaluations of the oracle's formulas (the oracle itself cannot hold
n x n at these sizes: SURVEY 8d memory gates) and size-independent properties.

C2 20 000 x 200 signed hybrid | C3 50 000 x 500 unsigned | C4 30 000 x 5 000 signed | C5 80 000 x 300 unsigned
"""
import numpy as np
import pytest
import torch

from pywgcna_b200 import _lib as L
from pywgcna_b200.device import DeviceNetwork
from pywgcna_b200.synth import CONFIGS, make_expression

pytestmark = pytest.mark.gpu


def _transform(c, net):
    if net == "unsigned":
        return np.abs(c)
    if net == "signed":
        return (1 + c) / 2
    return np.where(c < 0, 0.0, c)


@pytest.mark.parametrize("name", ["C2", "C3", "C4", "C5"])
def test_config_full_size(gpu, name):
    m, n, F, seed, net, tom_type = CONFIGS[name]
    need = (2 * n * n * 8 + 5 * n * n + m * n * 16) * 1.05
    free, total = torch.cuda.mem_get_info()
    if need > free:
        pytest.skip(f"{name} needs {need / 1e9:.0f} GB of HBM, {free / 1e9:.0f} GB free")
    X = make_expression(m, n, F=F, seed=seed)
    net_dev = DeviceNetwork(X, device=0, networkType=net)
    rng = np.random.default_rng(seed)
    S = np.sort(rng.choice(n, 48, replace=False))
    # float64 reference rows: correlation of the sampled genes with all genes
    Xc = X - X.mean(axis=0)
    Xn = Xc / np.sqrt((Xc * Xc).sum(axis=0))
    cor_S = np.clip(Xn[:, S].T @ Xn, -1, 1)  # [48, n]
    s_S = _transform(cor_S, net)

    # --- sweep: k of the sampled genes for every power (wgcna.py:913-916)
    powers = list(range(1, 21))
    k = net_dev.sweep(powers)  # [P, n]
    k_S = k[:, torch.from_numpy(S).cuda()].cpu().numpy()
    for j, p in enumerate(powers):
        ref = (s_S ** p).sum(axis=1) - 1.0  # diagonal term s = 1 removed by the -1
        np.testing.assert_allclose(k_S[j], ref, rtol=1e-10, atol=1e-9)
    assert torch.isfinite(k).all()

    # --- adjacency rows
    power = 6
    A = net_dev.adjacency(power)  # [n, n] on device
    A_S = A[torch.from_numpy(S).cuda()].cpu().numpy()
    assert np.max(np.abs(A_S - s_S ** power)) < 1e-11
    # exactly symmetric on a sampled block, range [0, 1]
    blk = A[torch.from_numpy(S).cuda()][:, torch.from_numpy(S).cuda()].cpu().numpy()
    assert np.array_equal(blk, blk.T) and blk.min() >= 0 and blk.max() <= 1

    # --- TOM on the sampled block: L[S, S] = A0[S] A0[S]^T needs only the sampled rows
    T = net_dev.tom(TOMType=tom_type, TOMDenom="min")
    T_SS = T[torch.from_numpy(S).cuda()][:, torch.from_numpy(S).cuda()].cpu().numpy()
    A0 = A_S.copy()
    A0[np.arange(len(S)), S] = 0.0
    kk = A0.sum(axis=1)
    Lss = A0 @ A0.T
    a = A0[:, S]
    ref = (Lss + a) / (np.minimum(kk[:, None], kk[None, :]) + 1 - a)
    np.fill_diagonal(ref, 1.0)
    assert np.max(np.abs(T_SS - ref)) < 1e-9
    # symmetry / diagonal / range on whole sampled rows
    T_S = T[torch.from_numpy(S).cuda()].cpu().numpy()
    T_colS = T[:, torch.from_numpy(S).cuda()].cpu().numpy()
    assert np.array_equal(T_S, T_colS.T)
    assert np.all(T_S[np.arange(len(S)), S] == 1.0) and T_S.min() >= 0 and T_S.max() <= 1.0 + 1e-12
    del net_dev, A, T, k
    torch.cuda.empty_cache()
