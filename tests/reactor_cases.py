"""the reference's closed-reactor regression case (tests/reactor/closed_reactors/rebless.py): H2/air phi = 1, 1200 K,
one atmosphere, integrate_to_steady with defaults; gold arrays in tests/golden/gold_closed_reactors.npz"""
import os

import numpy as np

from common import GOLDEN, build_mech


def run(backend, heat_transfer, configuration='isobaric'):
    from spitfire_b200.reactors import HomogeneousReactor
    m = build_mech('h2-burke', backend)
    air = m.stream(stp_air=True)
    fuel = m.stream('X', 'H2:1')
    mix = m.mix_for_equivalence_ratio(1.0, fuel, air)
    mix.TP = 1200., 101325.
    r = HomogeneousReactor(m, mix, configuration, heat_transfer, 'closed')
    return m, r.integrate_to_steady()


def compare_with_gold(m, lib, heat_transfer, rtol=1e-4, configuration='isobaric'):
    """closed_reactors/test.py:26-28 uses rtol 1e-4 on times, temperatures and mass fractions"""
    g = np.load(os.path.join(GOLDEN, 'gold_closed_reactors.npz'))
    key = ('cp, ' if configuration == 'isobaric' else 'cv, ') + heat_transfer
    t, T, Y = g[key + '__t'], g[key + '__T'], g[key + '__Y']
    assert lib.time_values.size == t.size, (lib.time_values.size, t.size)
    assert np.allclose(lib.time_values, t, rtol=rtol, atol=1e-12)
    assert np.allclose(lib['temperature'], T, rtol=rtol)
    names = m.species_names
    for i, s in enumerate(names):
        assert np.allclose(lib['mass fraction ' + s], Y[i], rtol=rtol, atol=1e-10), s
    return float(np.max(np.abs(lib['temperature'] - T) / T))
