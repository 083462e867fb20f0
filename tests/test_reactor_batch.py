"""HomogeneousReactorBatch (many isobaric reactors integrated together) against the serial
HomogeneousReactor and the reference's gold trajectory (tests/reactor/closed_reactors: H2/air, 1200 K, 1 atm).
CPU: the host logic with the oracle injected. GPU: the same through the batched C-ABI kernels."""
import os

import numpy as np
import pytest

from common import GOLDEN, build_mech, oracle_available

ORACLE = 'reference' if oracle_available('reference') else 'port'
T0S = [1100., 1200., 1300.]


def _template(backend, heat='adiabatic', T=1200.):
    from spitfire_b200.reactors import HomogeneousReactor
    m = build_mech('h2-burke', backend)
    air = m.stream(stp_air=True)
    fuel = m.stream('X', 'H2:1')
    mix = m.mix_for_equivalence_ratio(1.0, fuel, air)
    mix.TP = T, 101325.
    return m, mix, HomogeneousReactor(m, mix, 'isobaric', heat, 'closed')


def _batch(backend, heat='adiabatic'):
    from spitfire_b200.reactors import HomogeneousReactorBatch
    m, mix, r = _template(backend, heat)
    Y = np.tile(mix.Y, (len(T0S), 1))
    return m, mix, HomogeneousReactorBatch(r, T0S, Y)


def _check_against_gold_and_serial(backend):
    m, mix, b = _batch(backend)
    times, states, failed = b.integrate_to_steady(save_each_step=True)
    assert not np.any(failed)
    g = np.load(os.path.join(GOLDEN, 'gold_closed_reactors.npz'))
    t, T, Y = g['cp, adiabatic__t'], g['cp, adiabatic__T'], g['cp, adiabatic__Y']
    k = T0S.index(1200.)
    assert times[k].size == t.size, (times[k].size, t.size)
    assert np.allclose(times[k], t, rtol=1e-4, atol=1e-12)
    assert np.allclose(states[k][:, 0], T, rtol=1e-4)
    for i in range(m.n_species - 1):
        assert np.allclose(states[k][:, 1 + i], Y[i], rtol=1e-4, atol=1e-10)
    # every member follows the serial class's trajectory for its own initial state
    for k, T0 in enumerate(T0S):
        _, _, r = _template(backend, T=T0)
        lib = r.integrate_to_steady()
        assert lib.time_values.size == times[k].size, (T0, lib.time_values.size, times[k].size)
        # (same step count; the Newton iterations of the two drivers stop at slightly different iterates, which the
        # adaptive step sizes carry along: a few 1e-6 relative in time, more in T during the ignition front; the bar is the reference's own regression tolerance)
        assert np.allclose(lib.time_values, times[k], rtol=1e-4, atol=1e-14)
        assert np.allclose(lib['temperature'], states[k][:, 0], rtol=1e-4)
    return times


def test_reactor_batch_on_host_matches_gold_and_serial():
    _check_against_gold_and_serial(ORACLE)


def test_reactor_batch_ignition_delay_on_host():
    m, mix, b = _batch(ORACLE)
    tau = b.compute_ignition_delay()
    _, _, r = _template(ORACLE, T=1200.)
    assert abs(tau[1] - r.compute_ignition_delay()) <= 1e-5 * tau[1]
    assert tau[0] > tau[1] > tau[2] > 0.


def test_reactor_batch_refuses_time_dependent_parameters():
    from spitfire_b200.reactors import HomogeneousReactor, HomogeneousReactorBatch
    m, mix, _ = _template(ORACLE)
    r = HomogeneousReactor(m, mix, 'isobaric', 'adiabatic', 'open', mixing_tau=1e-3,
                           feed_temperature=lambda t: 1000. + t, feed_mass_fractions=mix.Y)
    with pytest.raises(ValueError):
        HomogeneousReactorBatch(r, [1200.], [mix.Y])


@pytest.mark.gpu
def test_reactor_batch_on_gpu_matches_gold_and_serial():
    _check_against_gold_and_serial('gpu')


@pytest.mark.gpu
def test_reactor_batch_gri_ignition_delays_on_gpu():
    """256 GRI-3.0 methane/air reactors, 1100-1900 K: delays fall monotonically with temperature and the batch agrees
    with the serial class on a member"""
    from spitfire_b200.reactors import HomogeneousReactor, HomogeneousReactorBatch
    m = build_mech('methane-gri30', 'gpu')
    air = m.stream(stp_air=True)
    fuel = m.stream('X', 'CH4:1')
    mix = m.mix_for_equivalence_ratio(1.0, fuel, air)
    mix.TP = 1500., 101325.
    T0 = np.linspace(1100., 1900., 256)
    b = HomogeneousReactorBatch(HomogeneousReactor(m, mix, 'isobaric', 'adiabatic', 'closed'), T0,
                                np.tile(mix.Y, (T0.size, 1)))
    tau = b.compute_ignition_delay()
    assert np.all(np.isfinite(tau)) and np.all(np.diff(tau) < 0.)
    k = 128
    mix.TP = float(T0[k]), 101325.
    serial = HomogeneousReactor(m, mix, 'isobaric', 'adiabatic', 'closed').compute_ignition_delay()
    print('GRI ignition delay at', T0[k], 'K:', tau[k], 'serial', serial)
    assert abs(tau[k] - serial) <= 1e-4 * serial


@pytest.mark.parametrize('heat,mass', [('isothermal', 'closed'), ('adiabatic', 'open'), ('diathermal', 'open')])
def test_reactor_batch_other_configurations_on_host(heat, mass):
    """open (constant feed), isothermal and diathermal reactors: the batch follows the serial class member by member"""
    from spitfire_b200.reactors import HomogeneousReactor, HomogeneousReactorBatch
    m, mix, _ = _template(ORACLE)
    feed = m.copy_stream(mix)
    feed.TP = 1400., 101325.
    kw = dict()
    if mass == 'open':
        kw.update(mixing_tau=1.e-4, feed_temperature=1400., feed_mass_fractions=feed.Y)
    if heat == 'diathermal':
        kw.update(convection_temperature=350., radiation_temperature=300., convection_coefficient=10.,
                  radiative_emissivity=0.5, shape_dimension_dict={'shape': 'sphere', 'char. length': 0.02})

    def make(T):
        mix.TP = T, 101325.
        return HomogeneousReactor(m, mix, 'isobaric', heat, mass, **kw)

    T0 = [1150., 1250.]
    b = HomogeneousReactorBatch(make(1200.), T0, np.tile(mix.Y, (2, 1)))
    times, states, failed = b.integrate_to_time(2.e-4, save_each_step=True)
    assert not np.any(failed)
    for k, T in enumerate(T0):
        lib = make(T).integrate_to_time(2.e-4)
        assert times[k][-1] == 2.e-4 and lib.time_values[-1] == 2.e-4  # both land on the final time exactly
        assert lib.time_values.size == times[k].size, (heat, mass, lib.time_values.size, times[k].size)
        assert np.allclose(lib.time_values, times[k], rtol=1e-4, atol=1e-14)
        assert np.allclose(lib['temperature'], states[k][:, 0], rtol=1e-4)


def test_two_rank_gloo_ignition_table(tmp_path):
    """the members of a batch are dealt to the ranks of torch.distributed (gloo on CPU here, NCCL on GPU boxes); the
    merged delays equal the single-process ones"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / 'tau.npz'
    port = 31500 + os.getpid() % 2000
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(port), os.path.join(root, 'tests', 'dist_ignition_worker.py'), ORACLE,
           str(out)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    d = np.load(out)
    assert int(d['world']) == 2
    from spitfire_b200.reactors import HomogeneousReactorBatch
    m, mix, r = _template(ORACLE)
    single = HomogeneousReactorBatch(r, d['T0'], np.tile(mix.Y, (d['T0'].size, 1))).compute_ignition_delay()
    assert np.array_equal(d['tau'], single)
    assert np.all(np.diff(d['tau']) < 0.)


@pytest.mark.gpu
@pytest.mark.parametrize('configuration', ['isobaric', 'isochoric'])
@pytest.mark.parametrize('heat_transfer', ['adiabatic', 'isothermal'])
def test_closed_reactor_gold_on_gpu(heat_transfer, configuration):
    """the four configurations of the reference's closed_reactors regression test (tests/reactor/closed_reactors),
    HomogeneousReactor on the CUDA kernels"""
    from reactor_cases import compare_with_gold, run
    m, lib = run('gpu', heat_transfer, configuration)
    compare_with_gold(m, lib, heat_transfer, configuration=configuration)


@pytest.mark.gpu
def test_isochoric_reactor_batch_matches_serial_on_gpu():
    from spitfire_b200.reactors import HomogeneousReactor, HomogeneousReactorBatch
    m = build_mech('h2-burke', 'gpu')
    air = m.stream(stp_air=True)
    fuel = m.stream('X', 'H2:1')
    mix = m.mix_for_equivalence_ratio(1.0, fuel, air)
    mix.TP = 1200., 101325.
    r = HomogeneousReactor(m, mix, 'isochoric', 'adiabatic', 'closed')
    b = HomogeneousReactorBatch(r, T0S, np.tile(mix.Y, (len(T0S), 1)))
    tau = b.compute_ignition_delay()
    for k, T0 in enumerate(T0S):
        mix.TP = T0, 101325.
        t1 = HomogeneousReactor(m, mix, 'isochoric', 'adiabatic', 'closed').compute_ignition_delay()
        assert abs(tau[k] - t1) <= 1e-4 * t1, (T0, tau[k], t1)
