"""Host-side flamelet / tabulation logic (spitfire_b200.flamelet, .tabulation, .equilibrium, .parallel) against the
reference's gold libraries, with the CPU oracle injected as the Griffon object -- no GPU needed. The same host code
drives the CUDA kernels in tests/test_gpu_flamelet.py.

The gold files were written by the reference with Cantera's equilibrium as the first initial guess; here that guess
comes from spitfire_b200.equilibrium. The converged fields nevertheless agree to ~1e-14 (adiabatic, steady) and ~1e-9
(transient), far inside the reference's own tolerances (1e-6; 2e-4 / 1e-4), which are the ones asserted where noted."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import ROOT, oracle_available
from slfm_cases import build, compare_with_gold, gold, h2_specs

ORACLE = 'reference' if oracle_available('reference') else 'port'


def test_equilibrium_initial_guess_is_sane():
    from spitfire_b200.flamelet import Flamelet, FlameletSpec
    fs = FlameletSpec(**h2_specs(ORACLE), initial_condition='equilibrium', stoich_dissipation_rate=1.)
    f = Flamelet(fs)
    T = f.initial_temperature
    assert 2500. < T.max() < 2700.  # H2 / 1200 K air adiabatic flame temperature
    y = f.initial_interior_state.reshape(-1, f.mechanism.n_species)[:, 1:]
    assert y.min() > -1e-12 and y.sum(axis=1).max() < 1. + 1e-12


def test_adiabatic_slfm_library_matches_reference_gold():
    """tests/tabulation/adiabatic_slfm/test.py: rtol = atol = 1e-6 in the reference; 1e-9 here"""
    lib = build('adiabatic_slfm', ORACLE)
    worst = compare_with_gold(lib, 'adiabatic_slfm', rtol=1e-9, atol=1e-12)
    print('adiabatic SLFM vs gold: worst normalised error', worst)


def test_adiabatic_slfm_waves_agree_with_chain():
    """solving several dissipation rates at once from the last converged member lands on the same table"""
    lib = build('adiabatic_slfm', ORACLE, wave=4)
    compare_with_gold(lib, 'adiabatic_slfm', rtol=1e-6, atol=1e-6)


def test_nonadiabatic_steady_slfm_library_matches_reference_gold():
    lib = build('nonadiabatic_defect_steady_slfm', ORACLE)
    compare_with_gold(lib, 'nonadiabatic_defect_steady_slfm', rtol=1e-8, atol=1e-9)


def test_nonadiabatic_transient_slfm_library_matches_reference_gold():
    """tests/tabulation/nonadiabatic_defect_transient_slfm/test.py: rtol 2e-4, atol 1e-4 in the reference"""
    lib = build('nonadiabatic_defect_transient_slfm', ORACLE)
    compare_with_gold(lib, 'nonadiabatic_defect_transient_slfm', rtol=1e-6, atol=1e-6)


def test_two_rank_gloo_sweep_matches_gold(tmp_path):
    """the reference runs its non-adiabatic builders with num_procs = 1 and 2; here: two torch.distributed ranks
    (gloo, CPU), chi_st dealt block-cyclically, one gather at the end"""
    out = tmp_path / 'lib.npz'
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'dist_worker.py'), ORACLE, str(out)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    d = np.load(out)
    g = gold('nonadiabatic_defect_transient_slfm')
    assert int(d['world']) == 2
    assert sorted(d['owned_0'].tolist() + d['owned_1'].tolist()) == [0, 1, 2, 3]
    # parallel.gather_profile_dicts: packed gather with uneven entry counts, and the fallback for ragged entries
    assert d['gather_profile_dicts_ok'].tolist() == [True, True]
    for p in ('temperature', 'mass fraction H2O', 'mass fraction OH', 'enthalpy_defect'):
        a, b = d[p], g['prop_' + p]
        assert np.max(np.abs(a - b) / (1e-6 * np.abs(b) + 1e-6)) <= 1., p


def test_library_slice_round_trip():
    """a Flamelet rebuilt from a one-dimensional library slice starts from that state (flamelet.py:135-180)"""
    from spitfire_b200.flamelet import Flamelet, FlameletSpec
    fs = FlameletSpec(**h2_specs(ORACLE), initial_condition='linear-TY', stoich_dissipation_rate=1.)
    f = Flamelet(fs)
    lib = f.make_library_from_interior_state(f.initial_interior_state)
    f2 = Flamelet(FlameletSpec(library_slice=lib, stoich_dissipation_rate=1.))
    assert np.allclose(f2.initial_interior_state, f.initial_interior_state, rtol=0, atol=1e-14)
    assert np.array_equal(f2.mixfrac_grid, f.mixfrac_grid)


@pytest.mark.parametrize('configuration', ['isobaric', 'isochoric'])
@pytest.mark.parametrize('heat_transfer', ['adiabatic', 'isothermal'])
def test_homogeneous_reactor_matches_reference_gold(heat_transfer, configuration):
    """BASELINE config 1 (H2/air ignition) through HomogeneousReactor + ESDIRK64 with the oracle kernels: the four
    configurations of the reference's closed_reactors regression test (tests/reactor/closed_reactors/test.py)"""
    from reactor_cases import compare_with_gold, run
    m, lib = run(ORACLE, heat_transfer, configuration)
    print(configuration, heat_transfer, 'steps', lib.time_values.size, 'max rel err T',
          compare_with_gold(m, lib, heat_transfer, configuration=configuration))


def test_structured_defect_interpolation_equals_numpy_interp_bit_for_bit():
    """tabulation.py:629-654 interpolates every property at every grid point with its own `interp` object; the
    column-wise restatement must return the same bits as np.interp"""
    from spitfire_b200.tabulation import _interp_columns
    rng = np.random.default_rng(1)
    for _ in range(300):
        n, nx, nc = rng.integers(1, 20), rng.integers(1, 40), rng.integers(1, 7)
        xp = np.sort(rng.standard_normal(n))
        if n > 1 and np.any(np.diff(xp) == 0):
            continue
        fp = rng.standard_normal((n, nc))
        x = np.concatenate([rng.uniform(xp[0] - 1, xp[-1] + 1, nx), xp[rng.integers(0, n, 3)]])
        ref = np.stack([np.interp(x, xp, fp[:, j]) for j in range(nc)], axis=1)
        assert np.array_equal(_interp_columns(x, xp, fp), ref)


def test_equilibrate_many_matches_the_one_at_a_time_solver():
    """the flamelet's 'equilibrium' initial condition solves all grid points in one array iteration; every stream must
    end where the single-stream Gibbs minimiser puts it (pure streams, trace mixtures and the interior)"""
    from common import build_mech
    from spitfire_b200.equilibrium import equilibrate_many
    for name, fu in (('methane-gri30', 'CH4:1'), ('h2-burke', 'H2:1')):
        m = build_mech(name, ORACLE)
        air = m.stream(stp_air=True)
        fuel = m.stream('TPX', (300., 101325., fu))
        zs = np.concatenate([[0., 1e-6], np.linspace(0.01, 0.99, 25), [1 - 1e-6, 1.]])
        mk = lambda z: m.mix_streams([(m.copy_stream(air), 1 - z), (m.copy_stream(fuel), z)], 'mass', 'HP')
        a = [mk(z) for z in zs]
        for q in a:
            q.equilibrate('HP')
        b = equilibrate_many([mk(z) for z in zs], 'HP')
        for x, y in zip(a, b):
            assert abs(x.T - y.T) <= 1e-11 * x.T and np.max(np.abs(x.Y - y.Y)) <= 1e-12
