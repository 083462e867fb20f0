"""the numpy statement of the twisted block-Thomas elimination / solve (oracle/twisted.py) against the dense LAPACK solve
of the assembled matrix (CPU), and the device kernels against that statement, block by block (GPU)"""
import numpy as np
import pytest

from cases import assemble_dense
from oracle import twisted


def _system(nb, bs, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal(bs * (nb * bs + 2 * (nb - 1)))
    for i in range(nb):
        blk = A[i * bs * bs:(i + 1) * bs * bs].reshape(bs, bs)
        blk[np.arange(bs), np.arange(bs)] += 2. * np.sqrt(bs) * np.sign(blk[np.arange(bs), np.arange(bs)])
    return A


@pytest.mark.parametrize('nb,bs', [(4, 3), (5, 7), (9, 11), (126, 5), (16, 53)])
def test_twisted_statement_solves_the_dense_system(nb, bs):
    A = _system(nb, bs, 100 * nb + bs)
    L, Di, m = twisted.twisted_invert(A, nb, bs)
    assert m == (nb - 1) // 2 and L[0] == m and L[1] == twisted.MAGIC
    b = np.random.default_rng(1).standard_normal(nb * bs)
    x = twisted.twisted_solve(A, L, Di, b, nb, bs)
    ref = np.linalg.solve(assemble_dense(A, nb, bs), b)
    assert np.max(np.abs(x - ref)) <= 1e-11 * np.max(np.abs(ref))


@pytest.mark.gpu
@pytest.mark.parametrize('nb,bs', [(4, 3), (9, 11), (12, 32), (126, 53)])
def test_device_twisted_factors_match_the_statement(nb, bs):
    """slot by slot: L_i / U_i, the inverses on both sides of the meeting block, the tag"""
    import torch
    from spitfire_b200 import griffon as gm
    A = _system(nb, bs, 7 * nb + bs)
    dA = torch.from_numpy(A[None, :].copy()).cuda()
    L = torch.zeros((1, nb * bs * bs), dtype=torch.float64, device='cuda')
    Di = torch.zeros_like(L)
    gm.btddod_full_invert(dA, nb, bs, L, Di, n_systems=1, twisted=True)
    Lr, Dr, m = twisted.twisted_invert(A, nb, bs)
    Lh, Dh = L.cpu().numpy()[0], Di.cpu().numpy()[0]
    assert Lh[0] == m and Lh[1] == twisted.MAGIC and not np.any(Lh[2:bs * bs])
    assert np.max(np.abs(Dh - Dr)) <= 1e-11 * np.max(np.abs(Dr))
    assert np.max(np.abs(Lh[bs * bs:] - Lr[bs * bs:])) <= 1e-11 * np.max(np.abs(Lr[bs * bs:]))
    b = np.random.default_rng(2).standard_normal(nb * bs)
    x = torch.zeros((1, nb * bs), dtype=torch.float64, device='cuda')
    gm.btddod_full_solve_inv(dA, L, Di, torch.from_numpy(b[None, :].copy()).cuda(), nb, bs, x, n_systems=1)
    xr = twisted.twisted_solve(A, Lr, Dr, b, nb, bs)
    assert np.max(np.abs(x.cpu().numpy()[0] - xr)) <= 1e-11 * np.max(np.abs(xr))
