"""BASELINE configs 4-5 at parity-test size: GRI-3.0 methane (300 K) / air (300 K) at one atmosphere on the reference's
default 128-point clustered grid.

  adiabatic : 17 of config 4's 64 dissipation rates (logspace(-3, 2, 64)): every fourth one over the burning range, the
              last burning member (index 55), the first extinguished one (56, which must end the table) and one beyond
  adiabatic_tight : the same table with the steady solver converged to 1e-9 instead of the reference's default 1e-6, so
              that the table no longer depends on the path the solver took (bar for the speculative waves)
  transient : config 5's transient heat-loss expansion for 4 dissipation rates x 16 stoichiometric enthalpy defects

The reference has no GRI-3.0 gold libraries, so the bar is the same host code driven by the unmodified
reference C++ kernels (oracle/_ref). tests/golden/make_gri_slfm.py runs that here and commits the tables as
tests/golden/ref_gri_slfm_*.npz; the GPU tests build the same libraries through the CUDA path and compare."""
import os

import numpy as np

from common import GOLDEN, build_mech

CHI64 = np.logspace(-3, 2, 64)
ADIABATIC_IDX = list(range(0, 56, 4)) + [55, 56, 60]
TRANSIENT_IDX = [8, 24, 40, 52]
N_DEFECT = 16


def gri_specs(backend, nz=128):
    from spitfire_b200.flamelet import FlameletSpec
    m = build_mech('methane-gri30', backend)
    air = m.stream(stp_air=True)
    fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
    return FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=nz)


TIGHT = 1.e-9  # steady-solver tolerance of the `adiabatic_tight` fixture (default: the reference's 1e-6; 1e-11 is below
              # the residual floor of the solver chain)


def build_adiabatic(backend, wave=1, tolerance=1.e-6):
    from spitfire_b200 import tabulation as tab
    return tab.build_adiabatic_slfm_library(gri_specs(backend), diss_rate_values=CHI64[ADIABATIC_IDX], verbose=False,
                                            wave=wave, tolerance=tolerance)


def build_transient(backend, wave=1):
    from spitfire_b200 import tabulation as tab
    return tab.build_nonadiabatic_defect_transient_slfm_library(gri_specs(backend),
                                                                diss_rate_values=CHI64[TRANSIENT_IDX], verbose=False,
                                                                n_defect_st=N_DEFECT, wave=wave)


def fixture_path(kind):
    return os.path.join(GOLDEN, f'ref_gri_slfm_{kind}.npz')


def save_fixture(lib, kind):
    data = {'dim_' + d: getattr(lib, d + '_values') for d in lib.dim_names}
    for p in lib.props:
        data['prop_' + p] = lib[p]
    np.savez_compressed(fixture_path(kind), **data)


def compare_with_fixture(lib, kind, tol_T, tol_Y):
    """max over the table of |a - b| / max|b| per field (the field's scale); mass fractions whose maximum is below
    1e-12 are compared on an absolute 1e-20 floor. Returns (err_T, err_Y)."""
    g = np.load(fixture_path(kind))
    for d in lib.dim_names:
        a, b = getattr(lib, d + '_values'), g['dim_' + d]
        assert a.shape == b.shape, (d, a.shape, b.shape)
        assert np.allclose(a, b, rtol=1e-8, atol=1e-8 * np.max(np.abs(b))), (d, a, b)
    T, Tr = lib['temperature'], g['prop_temperature']
    assert T.shape == Tr.shape
    err_T = float(np.max(np.abs(T - Tr)) / np.max(np.abs(Tr)))
    err_Y, worst = 0., None
    for p in lib.props:
        if not p.startswith('mass fraction'):
            continue
        a, b = lib[p], g['prop_' + p]
        e = float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-20))
        if np.max(np.abs(b)) > 1e-12 and e > err_Y:
            err_Y, worst = e, p
    print(f'{kind}: temperature differs by {err_T:.3e} of its scale, {worst} by {err_Y:.3e}')
    assert err_T <= tol_T, f'{kind}: temperature differs by {err_T:.3e} of its scale (bar {tol_T})'
    assert err_Y <= tol_Y, f'{kind}: {worst} differs by {err_Y:.3e} of its scale (bar {tol_Y})'
    return err_T, err_Y
