"""the reference's own SLFM regression cases (tests/tabulation/{adiabatic_slfm,nonadiabatic_defect_*_slfm}/rebless.py):
H2 (burke) 300 K against air 1200 K at one atmosphere, 34-point clustered grid; gold arrays in tests/golden/*.npz"""
import os

import numpy as np

from common import GOLDEN, build_mech


def h2_specs(backend):
    m = build_mech('h2-burke', backend)
    air = m.stream(stp_air=True)
    air.TP = 1200., 101325.
    fuel = m.stream('TPY', (300., 101325., 'H2:1'))
    return {'mech_spec': m, 'oxy_stream': air, 'fuel_stream': fuel, 'grid_points': 34}


def gold(name):
    return np.load(os.path.join(GOLDEN, f'gold_{name}.npz'))


def build(kind, backend, **kwargs):
    from spitfire_b200 import tabulation as tab
    fs = h2_specs(backend)
    if kind == 'adiabatic_slfm':
        return tab.build_adiabatic_slfm_library(fs, verbose=False, diss_rate_values=np.logspace(0, 1, 8), **kwargs)
    if kind == 'nonadiabatic_defect_transient_slfm':
        return tab.build_nonadiabatic_defect_transient_slfm_library(
            fs, verbose=False, diss_rate_values=np.logspace(0, 1, 4), integration_args={'transient_tolerance': 1e-10},
            **kwargs)
    if kind == 'nonadiabatic_defect_steady_slfm':
        return tab.build_nonadiabatic_defect_steady_slfm_library(
            fs, verbose=False, diss_rate_values=np.logspace(0, 1, 4), integration_args={'transient_tolerance': 1e-10},
            **kwargs)
    raise ValueError(kind)


def compare_with_gold(lib, kind, rtol, atol):
    """every dimension and every field the builder itself produces (temperature, pressure, mass fractions, enthalpy
    defect) against the reference's gold library. The gold files also hold fields that the reference's test adds
    afterwards with Cantera (enthalpy, cp, cv, density, viscosity); those are not outputs of this path."""
    g = gold(kind)
    worst = 0.
    for d in lib.dim_names:
        a, b = getattr(lib, d + '_values'), g['dim_' + d]
        assert a.shape == b.shape, (d, a.shape, b.shape)
        assert np.allclose(a, b, rtol=1e-8, atol=1e-12), d
    fields = ['temperature', 'pressure'] + [p for p in lib.props if p.startswith('mass fraction')]
    if 'enthalpy_defect' in lib.props and 'prop_enthalpy_defect' in g:
        fields.append('enthalpy_defect')
    for p in fields:
        a, b = lib[p], g['prop_' + p]
        assert a.shape == b.shape, (p, a.shape, b.shape)
        err = np.max(np.abs(a - b) / (rtol * np.abs(b) + atol))
        worst = max(worst, float(err))
        assert err <= 1., f'{kind} {p}: max |d|/(rtol |gold| + atol) = {err:.3e} (rtol {rtol}, atol {atol})'
    return worst
