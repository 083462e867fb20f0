"""GPU tests of the flamelet solvers and library builders: the same host code as tests/test_host_flamelet.py, with the
CUDA kernels (C-ABI library) behind it. Bars: the reference's gold libraries; for GRI-3.0, which has no gold file in the
reference, the CPU oracle driving the same solver on the same initial state (converged fields within 1e-8,
BASELINE.json)."""
import numpy as np
import pytest

from common import build_mech, oracle_available
from slfm_cases import build, compare_with_gold

pytestmark = pytest.mark.gpu
ORACLE = 'reference' if oracle_available('reference') else 'port'


def test_adiabatic_slfm_library_on_gpu_matches_reference_gold():
    lib = build('adiabatic_slfm', 'gpu')
    worst = compare_with_gold(lib, 'adiabatic_slfm', rtol=1e-8, atol=1e-10)
    print('adiabatic SLFM (GPU) vs gold: worst normalised error', worst)


def test_adiabatic_slfm_waves_on_gpu():
    lib = build('adiabatic_slfm', 'gpu', wave=8)
    compare_with_gold(lib, 'adiabatic_slfm', rtol=1e-6, atol=1e-6)


def test_nonadiabatic_steady_slfm_library_on_gpu_matches_reference_gold():
    lib = build('nonadiabatic_defect_steady_slfm', 'gpu')
    compare_with_gold(lib, 'nonadiabatic_defect_steady_slfm', rtol=1e-8, atol=1e-9)


def test_nonadiabatic_transient_slfm_library_on_gpu_matches_reference_gold():
    lib = build('nonadiabatic_defect_transient_slfm', 'gpu')
    compare_with_gold(lib, 'nonadiabatic_defect_transient_slfm', rtol=1e-6, atol=1e-6)


def _gri_flamelets(backend, chis, nz=48):
    from spitfire_b200.flamelet import Flamelet, FlameletSpec
    m = build_mech('methane-gri30', backend)
    air = m.stream(stp_air=True)
    fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
    fs = FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=nz, initial_condition='equilibrium',
                      stoich_dissipation_rate=float(chis[0]))
    first = Flamelet(fs)
    out = [first]
    for c in chis[1:]:
        fs.initial_condition = np.copy(first.initial_interior_state)
        fs.stoich_dissipation_rate = float(c)
        out.append(Flamelet(fs))
    return out


def test_gri30_steady_flamelets_gpu_vs_oracle():
    """GRI-3.0 methane/air flamelets: batched steady solve on the GPU against the same solver driven by the CPU
    oracle from the same (equilibrium) initial state; converged T and Y within 1e-8 relative (of the field's scale)"""
    from spitfire_b200.flamelet import FlameletBatch
    chis = [0.5, 2., 8.]
    sg, tg = FlameletBatch(_gri_flamelets('gpu', chis)).compute_steady_state(tolerance=1e-8)
    so, to = FlameletBatch(_gri_flamelets(ORACLE, chis)).compute_steady_state(tolerance=1e-8)
    assert list(tg) == list(to), (tg, to)
    ns = 53
    for k in range(len(chis)):
        a, b = sg[k].reshape(-1, ns), so[k].reshape(-1, ns)
        assert b[:, 0].max() > 1800.  # burning
        errT = np.max(np.abs(a[:, 0] - b[:, 0])) / np.max(np.abs(b[:, 0]))
        errY = np.max(np.abs(a[:, 1:] - b[:, 1:]) / (np.max(np.abs(b[:, 1:]), axis=0) + 1e-30 + 1e-12))
        print(f'chi_st {chis[k]}: solver {tg[k]}, rel err T {errT:.2e}, Y {errY:.2e}')
        assert errT <= 1e-8 and errY <= 1e-6


def test_homogeneous_reactor_on_gpu_matches_reference_gold():
    """BASELINE config 1: H2/air isobaric adiabatic ignition, single-state calls through the C-ABI host entry points"""
    from reactor_cases import compare_with_gold, run
    m, lib = run('gpu', 'adiabatic')
    print('steps', lib.time_values.size, 'max rel err T', compare_with_gold(m, lib, 'adiabatic'))


def test_inverse_based_block_thomas_solve_matches_lu_solve():
    """extension entry points gb_btddod_full_{factorize,solve}_inv_batch against the LU-based solve"""
    import torch
    from spitfire_b200 import griffon
    from cases import flamelet_all, flamelet_case
    mg, mo = build_mech('methane-gri30', 'gpu'), build_mech('methane-gri30', ORACLE)
    c = flamelet_case(mg, 24)
    ns, nzi = c['ns'], c['nzi']
    A0 = flamelet_all(mo.griffon, c, eig=False)['jac0True']
    F = 3
    rng = np.random.default_rng(5)
    rhs = rng.normal(size=(F, nzi * ns))
    dA = torch.from_numpy(np.array([A0] * F)).cuda()
    dA2 = dA.clone()
    dL, dL2 = torch.zeros((F, nzi * ns * ns), dtype=torch.float64, device='cuda'), torch.zeros((F, nzi * ns * ns), dtype=torch.float64, device='cuda')
    dP, dP2 = torch.zeros((F, nzi * ns), dtype=torch.int32, device='cuda'), torch.zeros((F, nzi * ns), dtype=torch.int32, device='cuda')
    dI = torch.zeros((F, nzi * ns * ns), dtype=torch.float64, device='cuda')
    dR = torch.from_numpy(rhs).cuda()
    x1, x2 = torch.zeros_like(dR), torch.zeros_like(dR)
    griffon.py_btddod_full_factorize(dA, nzi, ns, dL, dP, n_systems=F)
    griffon.py_btddod_full_solve(dA, dL, dP, dR, nzi, ns, x1, n_systems=F)
    griffon.btddod_full_factorize_inv(dA2, nzi, ns, dL2, dP2, dI, n_systems=F)
    griffon.btddod_full_solve_inv(dA2, dL2, dI, dR, nzi, ns, x2, n_systems=F)
    torch.cuda.synchronize()
    assert torch.equal(dA, dA2) and torch.equal(dL, dL2) and torch.equal(dP, dP2)
    a, b = x2.cpu().numpy(), x1.cpu().numpy()
    assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(b))


@pytest.mark.parametrize('bs,nb', [(11, 1), (11, 2), (11, 33), (53, 40), (64, 5), (70, 4), (5, 3), (11, 2600)])
def test_inverse_based_solve_shapes(bs, nb):
    """odd / even block sizes (8-byte vs 16-byte aligned blocks), one and two blocks, more than 64 rows per block, and a
    system whose right-hand side does not fit in shared memory: inverse-based solve against the LU-based one and
    against the residual of the original matrix"""
    import torch
    from spitfire_b200 import griffon
    rng = np.random.default_rng(bs * 1000 + nb)
    F = 3
    nelem = bs * (nb * bs + 2 * (nb - 1))
    A = rng.normal(size=(F, nelem))
    diag = A[:, :nb * bs * bs].reshape(F, nb, bs, bs)
    diag += 2. * bs * np.eye(bs)  # diagonally dominant blocks
    rhs = rng.normal(size=(F, nb * bs))
    dA0 = torch.from_numpy(A).cuda()
    dA, dA2 = dA0.clone(), dA0.clone()
    z = lambda *shape, dt=torch.float64: torch.zeros(shape, dtype=dt, device='cuda')
    dL, dL2, dI = z(F, nb * bs * bs), z(F, nb * bs * bs), z(F, nb * bs * bs)
    dP, dP2 = z(F, nb * bs, dt=torch.int32), z(F, nb * bs, dt=torch.int32)
    dR = torch.from_numpy(rhs).cuda()
    x1, x2, mv = torch.zeros_like(dR), torch.zeros_like(dR), torch.zeros_like(dR)
    griffon.py_btddod_full_factorize(dA, nb, bs, dL, dP, n_systems=F)
    griffon.py_btddod_full_solve(dA, dL, dP, dR, nb, bs, x1, n_systems=F)
    griffon.btddod_full_factorize_inv(dA2, nb, bs, dL2, dP2, dI, n_systems=F)
    griffon.btddod_full_solve_inv(dA2, dL2, dI, dR, nb, bs, x2, n_systems=F)
    griffon.py_btddod_full_matvec(dA0, x2, nb, bs, mv, n_systems=F)
    torch.cuda.synchronize()
    a, b = x2.cpu().numpy(), x1.cpu().numpy()
    assert np.all(np.isfinite(a))
    assert np.max(np.abs(a - b)) <= 1e-11 * np.max(np.abs(b))
    assert np.max(np.abs(mv.cpu().numpy() - rhs)) <= 1e-11 * np.max(np.abs(rhs))


def test_inverse_based_solve_addresses_a_subset_of_the_factors_in_place():
    """system_rows: right-hand sides k = 0..m-1 are solved with the factors of systems rows[k] of a larger batch"""
    import torch
    from spitfire_b200 import griffon
    rng = np.random.default_rng(7)
    F, bs, nb = 6, 11, 9
    nelem = bs * (nb * bs + 2 * (nb - 1))
    A = rng.normal(size=(F, nelem))
    A[:, :nb * bs * bs].reshape(F, nb, bs, bs)[...] += 2. * bs * np.eye(bs)
    rhs = rng.normal(size=(F, nb * bs))
    dA = torch.from_numpy(A).cuda()
    z = lambda *shape, dt=torch.float64: torch.zeros(shape, dtype=dt, device='cuda')
    dL, dI, dP = z(F, nb * bs * bs), z(F, nb * bs * bs), z(F, nb * bs, dt=torch.int32)
    griffon.btddod_full_factorize_inv(dA, nb, bs, dL, dP, dI, n_systems=F)
    dR = torch.from_numpy(rhs).cuda()
    full = torch.zeros_like(dR)
    griffon.btddod_full_solve_inv(dA, dL, dI, dR, nb, bs, full, n_systems=F)
    rows = torch.tensor([4, 1, 5], dtype=torch.int32, device='cuda')
    sub = torch.zeros((3, nb * bs), dtype=torch.float64, device='cuda')
    griffon.btddod_full_solve_inv(dA, dL, dI, dR[rows.long()].contiguous(), nb, bs, sub, n_systems=3, system_rows=rows)
    torch.cuda.synchronize()
    assert torch.equal(sub, full[rows.long()])
