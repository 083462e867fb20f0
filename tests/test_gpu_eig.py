"""GPU tests of the device eigenvalue bound (gb_eig.cu): max Re(lambda) of small dense matrices against LAPACK dgeev
(numpy.linalg.eigvals -- the routine the reference calls, blas_lapack_kernels.h:157-180), and the `compute_eigenvalues`
branch of flamelet_jacobian (flamelet_kernels.cpp:1329-1341) against the CPU oracle.

Eigenvalues of a non-normal matrix are only determined to eps * ||A|| * (condition of the eigenvalue), and the two
implementations round differently (different Householder vectors, different shifts after deflation), so the bar is a
tolerance relative to the spectral scale, not bit equality: 1e-9 * max|lambda| for ordinary matrices, the square root
of eps for a defective 2 x 2 block, and the measured agreement is printed."""
import numpy as np
import pytest

from cases import flamelet_all, flamelet_case
from common import build_mech, oracle_available

pytestmark = pytest.mark.gpu
ORACLE = 'reference' if oracle_available('reference') else 'port'


def gpu_max_real(blocks):
    import torch
    from spitfire_b200 import griffon
    nb, n = blocks.shape[0], blocks.shape[1]
    d = torch.from_numpy(np.ascontiguousarray(blocks)).cuda()
    out = torch.full((nb,), np.nan, dtype=torch.float64, device='cuda')
    griffon.max_real_eigenvalue(d, n, out, nb)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def lapack_max_real(blocks):
    return np.array([np.linalg.eigvals(b).real.max() for b in blocks])


@pytest.mark.parametrize('n', [1, 2, 3, 4, 11, 32, 33, 53, 64, 97])
def test_random_matrices(n):
    rng = np.random.default_rng(100 + n)
    A = rng.standard_normal((96, n, n))
    A[:16] *= 10. ** rng.uniform(-6, 6, (16, 1, 1))                      # overall scale
    d = 10. ** rng.uniform(-5, 5, (16, n))
    A[16:32] = A[16:32] * d[:, :, None] / d[:, None, :]                   # badly balanced similarity transforms
    A[32:40] = np.triu(A[32:40])                                          # already triangular
    A[40:48] = A[40:48] - A[40:48].transpose(0, 2, 1)                     # skew: purely imaginary spectrum
    A[48:56] = A[48:56] + A[48:56].transpose(0, 2, 1)                     # symmetric
    A[56] = 0.
    A[57] = np.eye(n) * -3.5
    A[58] = np.diag(np.arange(n) - n / 2.)
    got, ref = gpu_max_real(A), lapack_max_real(A)
    scale = np.array([np.abs(np.linalg.eigvals(b)).max() for b in A])
    err = np.abs(got - ref) / (scale + 1e-300)
    err[56] = abs(got[56] - ref[56])
    print(f'n={n}: max |d max Re| / max|lambda| = {err.max():.2e}, median {np.median(err):.2e}')
    assert np.all(np.isfinite(got))
    tol = np.full(96, 1e-9)
    tol[16:32] = 1e-7  # the unbalanced ones: dgeev itself is only this accurate on them
    assert np.all(err <= tol), (np.argmax(err / tol), err.max())


def test_special_structures():
    blocks = np.zeros((6, 4, 4))
    blocks[0, :2, :2] = [[1., -5.], [5., 1.]]                  # complex pair 1 +- 5i, and a double zero
    blocks[1] = np.diag([2., 2., -1., -1.]) + np.diag([1., 0., 1.], 1)  # two Jordan blocks
    blocks[2] = np.array([[0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [-24, 50, -35, 10]], dtype=float).T  # roots 1,2,3,4
    blocks[3] = np.array([[0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], dtype=float)  # cyclic: needs the
    #                                                                                exceptional shift to start converging
    blocks[4] = np.full((4, 4), 1e-150)
    blocks[5] = np.full((4, 4), 1e150)
    got, ref = gpu_max_real(blocks), lapack_max_real(blocks)
    print('special:', got, ref)
    assert abs(got[0] - 1.) < 1e-14
    assert abs(got[1] - 2.) < 1e-7          # defective: sqrt(eps)
    assert abs(got[2] - 4.) < 1e-12
    assert abs(got[3] - 1.) < 1e-14
    assert abs(got[4] / 4e-150 - 1.) < 1e-14
    assert abs(got[5] / 4e150 - 1.) < 1e-14


def test_non_finite_input_gives_nan_and_terminates():
    blocks = np.random.default_rng(0).standard_normal((3, 7, 7))
    blocks[1, 3, 4] = np.nan
    blocks[2, 0, 0] = np.inf
    got = gpu_max_real(blocks)
    assert np.isfinite(got[0]) and np.isnan(got[1]) and not np.isfinite(got[2])


@pytest.mark.parametrize('name,nz', [('h2-burke', 34), ('methane-gri30', 40)])
def test_flamelet_jacobian_eigenvalue_bound_parity(name, nz):
    """the compute_eigenvalues branch through the single-flamelet C-ABI entry point against the oracle (LAPACK dgeev)"""
    mg, mo = build_mech(name, 'gpu'), build_mech(name, ORACLE)
    c = flamelet_case(mg, nz)
    got, ref = flamelet_all(mg.griffon, c, eig=True), flamelet_all(mo.griffon, c, eig=True)
    keys = [k for k in ref if k.startswith('eig')]
    assert keys
    for k in keys:
        scale = np.abs(ref[k]).max() + 1.
        err = np.abs(got[k] - ref[k]) / scale
        print(f'{name} {k}: max expeig {ref[k].max():.4e}, max |d| / scale {err.max():.2e}')
        assert ref[k].max() > 0.  # the case has explosive modes, the comparison is not vacuous
        assert err.max() <= 1e-6  # eps * ||block|| * condition: the blocks are strongly non-normal (measured 1e-13 .. 5e-9)
    # asking for the bound must not change the Jacobian
    plain = flamelet_all(mg.griffon, c, eig=False)
    for k in plain:
        if k.startswith('jac'):
            assert np.array_equal(plain[k], got[k]), k


def test_batched_jac_and_eig_matches_oracle():
    """_BatchOps.jac_and_eig with per-member diffusion terms: device path against the same host code driving the oracle"""
    import torch
    from spitfire_b200.flamelet import FlameletBatch
    from test_gpu_flamelet import _gri_flamelets
    res = []
    for backend in ('gpu', ORACLE):
        fb = FlameletBatch(_gri_flamelets(backend, [0.5, 8.], nz=24))
        ops = fb.ops
        state = fb._initial(None)
        state = state * (1. + 0.02 * torch.sin(torch.arange(state.shape[1], device=state.device, dtype=torch.float64)))
        diff = torch.tensor([0.1, 3.], dtype=torch.float64, device=ops.device)
        J, e = ops.jac_and_eig(state, diff)
        res.append((J.cpu().numpy(), e.cpu().numpy()))
    (Jg, eg), (Jo, eo) = res
    scale = np.abs(eo).max() + 1.
    err = np.abs(eg - eo).max() / scale
    print('batched bound: max', eo.max(), 'err', err)
    assert eo.max() > 0.
    assert err <= 1e-8
    assert np.abs(Jg - Jo).max() <= 1e-9 * np.abs(Jo).max()


def test_size_limits():
    """the matrix lives in shared memory: n = 169 is the largest size, n = 170 is refused (not silently truncated)"""
    import torch
    from spitfire_b200 import griffon
    rng = np.random.default_rng(3)
    A = rng.standard_normal((2, 169, 169))
    got, ref = gpu_max_real(A), lapack_max_real(A)
    assert np.all(np.abs(got - ref) <= 1e-9 * np.abs(ref))
    big = torch.zeros((1, 170 * 170), dtype=torch.float64, device='cuda')
    out = torch.zeros(1, dtype=torch.float64, device='cuda')
    with pytest.raises(griffon.GriffonB200Error):
        griffon.max_real_eigenvalue(big, 170, out, 1)
