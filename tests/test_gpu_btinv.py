"""k_btddod_invert (csrc/gb_btinv.cu): block-Thomas elimination by Gauss-Jordan inverses, against the LU-based
factorize_inv kernel (which reproduces the reference's dgetrf / dgetrs arithmetic) and against a dense LAPACK solve of
the assembled matrix (the reference's own check, tests/griffon/test_btddod.py: dense solve vs block solve)."""
import numpy as np
import pytest

from cases import assemble_dense


def _systems(n, nb, bs, seed, dominance):
    rng = np.random.default_rng(seed)
    nel = bs * (nb * bs + 2 * (nb - 1))
    A = rng.standard_normal((n, nel))
    for s in range(n):
        for i in range(nb):
            blk = A[s, i * bs * bs:(i + 1) * bs * bs].reshape(bs, bs)
            blk[np.arange(bs), np.arange(bs)] += dominance * np.sign(blk[np.arange(bs), np.arange(bs)])
    return A


@pytest.mark.gpu
@pytest.mark.parametrize('nb,bs', [(1, 1), (1, 53), (3, 5), (7, 11), (5, 32), (4, 33), (126, 53), (6, 56), (3, 64),
                                    (3, 65), (2, 97), (2, 128)])
def test_invert_matches_lu_inverse_and_dense_solve(nb, bs):
    import torch
    from spitfire_b200 import griffon as gm
    n = 3
    A = _systems(n, nb, bs, seed=nb * 1000 + bs, dominance=2. * np.sqrt(bs))
    dA = torch.from_numpy(A).cuda()
    keep = dA.clone()
    L = torch.zeros((n, nb * bs * bs), dtype=torch.float64, device='cuda')
    Di = torch.zeros_like(L)
    gm.btddod_full_invert(dA, nb, bs, L, Di, n_systems=n)
    assert torch.equal(dA, keep)  # the matrix is left intact
    if bs > 120:  # (the LU kernel stops at 120 x 120 blocks) first block against LAPACK's inverse
        inv0 = Di[0, :bs * bs].cpu().numpy().reshape(bs, bs).T
        blk0 = A[0, :bs * bs].reshape(bs, bs).T
        assert np.max(np.abs(inv0 @ blk0 - np.eye(bs))) <= 1e-11
        return
    # LU-based reference kernel
    J2 = keep.clone()
    L2, D2 = torch.zeros_like(L), torch.zeros_like(L)
    piv = torch.zeros((n, nb * bs), dtype=torch.int32, device='cuda')
    gm.btddod_full_factorize_inv(J2, nb, bs, L2, piv, D2, n_systems=n)
    sD = float(D2.abs().max())
    assert float((Di - D2).abs().max()) <= 1e-11 * sD
    if nb > 1:
        sL = float(L2[:, bs * bs:].abs().max())
        assert float((L[:, bs * bs:] - L2[:, bs * bs:]).abs().max()) <= 1e-11 * sL
    if bs > 80:
        return  # (k_btddod_solve_inv keeps a ring of four blocks in shared memory: blocks up to ~80 x 80)
    # solve through solve_inv with the intact matrix as d_factors, against the dense solve
    rng = np.random.default_rng(7)
    b = rng.standard_normal((n, nb * bs))
    db = torch.from_numpy(b).cuda()
    x = torch.zeros_like(db)
    gm.btddod_full_solve_inv(dA, L, Di, db, nb, bs, x, n_systems=n)
    xh = x.cpu().numpy()
    for s in range(n):
        ref = np.linalg.solve(assemble_dense(A[s], nb, bs), b[s])
        assert np.max(np.abs(xh[s] - ref)) <= 1e-10 * np.max(np.abs(ref))


@pytest.mark.gpu
def test_invert_pivots_on_a_permuted_block():
    """a block whose diagonal is zero needs the row interchanges"""
    import torch
    from spitfire_b200 import griffon as gm
    bs, nb = 53, 2
    rng = np.random.default_rng(3)
    P = np.eye(bs)[rng.permutation(bs)]
    blocks = [P * (1. + rng.random((bs, bs))) + 1e-3 * rng.standard_normal((bs, bs)) for _ in range(nb)]
    A = np.concatenate([b.T.ravel() for b in blocks] + [0.1 * rng.standard_normal(2 * (nb - 1) * bs)])[None, :]
    dA = torch.from_numpy(np.ascontiguousarray(A)).cuda()
    L = torch.zeros((1, nb * bs * bs), dtype=torch.float64, device='cuda')
    Di = torch.zeros_like(L)
    gm.btddod_full_invert(dA, nb, bs, L, Di, n_systems=1)
    inv0 = Di[0, :bs * bs].cpu().numpy().reshape(bs, bs).T
    assert np.max(np.abs(inv0 @ blocks[0] - np.eye(bs))) <= 1e-11
    b = rng.standard_normal((1, nb * bs))
    x = torch.zeros((1, nb * bs), dtype=torch.float64, device='cuda')
    gm.btddod_full_solve_inv(dA, L, Di, torch.from_numpy(b).cuda(), nb, bs, x, n_systems=1)
    ref = np.linalg.solve(assemble_dense(A[0], nb, bs), b[0])
    assert np.max(np.abs(x.cpu().numpy()[0] - ref)) <= 1e-9 * np.max(np.abs(ref))


@pytest.mark.gpu
@pytest.mark.parametrize('nb,bs', [(4, 2), (4, 5), (5, 11), (7, 32), (8, 33), (126, 53), (127, 53), (9, 64), (6, 65), (3, 7)])
def test_twisted_elimination_solves_like_the_dense_matrix(nb, bs):
    """gb_btddod_full_invert_twisted_batch + gb_btddod_full_solve_inv_batch (two CTAs of a cluster per system, meeting
    in the middle block) against the dense solve and against the one-sided elimination; (3, 7): below four blocks the
    twisted entry point falls back to the one-sided form"""
    import torch
    from spitfire_b200 import griffon as gm
    n = 5
    A = _systems(n, nb, bs, seed=nb * 77 + bs, dominance=2. * np.sqrt(bs))
    dA = torch.from_numpy(A).cuda()
    keep = dA.clone()
    L = torch.zeros((n, nb * bs * bs), dtype=torch.float64, device='cuda')
    Di = torch.zeros_like(L)
    gm.btddod_full_invert(dA, nb, bs, L, Di, n_systems=n, twisted=True)
    assert torch.equal(dA, keep)
    m = (nb - 1) // 2
    if nb >= 4 and bs <= 64:
        # the format tag, the untouched top half (same arithmetic as the one-sided elimination) and a different bottom
        assert float(L[0, 0]) == float(m) and float(L[0, 1]) == 2.718281828459045e-300
        L1, D1 = torch.zeros_like(L), torch.zeros_like(L)
        gm.btddod_full_invert(dA, nb, bs, L1, D1, n_systems=n)
        assert torch.equal(Di[:, :m * bs * bs], D1[:, :m * bs * bs])
        assert torch.equal(L[:, bs * bs:(m + 1) * bs * bs], L1[:, bs * bs:(m + 1) * bs * bs])
        assert not torch.equal(Di[:, (m + 1) * bs * bs:], D1[:, (m + 1) * bs * bs:])
    rng = np.random.default_rng(11)
    b = rng.standard_normal((n, nb * bs))
    db = torch.from_numpy(b).cuda()
    x = torch.zeros_like(db)
    gm.btddod_full_solve_inv(dA, L, Di, db, nb, bs, x, n_systems=n)
    xh = x.cpu().numpy()
    for s in range(n):
        ref = np.linalg.solve(assemble_dense(A[s], nb, bs), b[s])
        assert np.max(np.abs(xh[s] - ref)) <= 1e-10 * np.max(np.abs(ref))
    # a subset of the systems addressed in place, right-hand sides compact
    rows = torch.tensor([3, 0, 4], dtype=torch.int32, device='cuda')
    sub = torch.zeros((3, nb * bs), dtype=torch.float64, device='cuda')
    gm.btddod_full_solve_inv(dA, L, Di, db[rows.long()].contiguous(), nb, bs, sub, n_systems=3, system_rows=rows)
    assert torch.equal(sub, x[rows.long()])
    # repeatable bit for bit
    x2 = torch.zeros_like(db)
    gm.btddod_full_solve_inv(dA, L, Di, db, nb, bs, x2, n_systems=n)
    assert torch.equal(x, x2)


@pytest.mark.gpu
def test_twisted_elimination_large_batch():
    """more systems than cluster slots: the persistent loops of both kernels"""
    import torch
    from spitfire_b200 import griffon as gm
    n, nb, bs = 200, 6, 11
    A = _systems(n, nb, bs, seed=5, dominance=2. * np.sqrt(bs))
    dA = torch.from_numpy(A).cuda()
    L = torch.zeros((n, nb * bs * bs), dtype=torch.float64, device='cuda')
    Di = torch.zeros_like(L)
    gm.btddod_full_invert(dA, nb, bs, L, Di, n_systems=n, twisted=True)
    b = np.random.default_rng(2).standard_normal((n, nb * bs))
    db = torch.from_numpy(b).cuda()
    x = torch.zeros_like(db)
    gm.btddod_full_solve_inv(dA, L, Di, db, nb, bs, x, n_systems=n)
    xh = x.cpu().numpy()
    for s in (0, 73, 148, 199):
        ref = np.linalg.solve(assemble_dense(A[s], nb, bs), b[s])
        assert np.max(np.abs(xh[s] - ref)) <= 1e-10 * np.max(np.abs(ref))
