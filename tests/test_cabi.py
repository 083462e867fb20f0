"""C-ABI library: loads, exports every symbol include/griffon_b200.h declares, builds mechanisms on the host, and
refuses to compute without a GPU (there is no CPU fallback). No compute calls that need a device here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import ROOT, build_mech, load_mech_data
from spitfire_b200 import griffon
from spitfire_b200.griffon import GriffonB200Error, PyCombustionKernels


def declared_symbols():
    with open(os.path.join(ROOT, 'include', 'griffon_b200.h')) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(griffon.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f'missing exports: {missing}'


def test_library_has_no_torch_or_python_dependency():
    import subprocess
    out = subprocess.run(['ldd', griffon.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert 'torch' not in out and 'python' not in out


def test_mechanism_builds_on_host():
    m = build_mech('h2-burke', 'gpu')
    g = m.griffon
    assert g.n_species == 11 and g.n_reactions == 27
    np.testing.assert_allclose(g.molecular_weights[:3], [1.008, 2.016, 15.999], rtol=0, atol=1e-12)
    m2 = build_mech('methane-gri30', 'gpu')
    assert m2.griffon.n_species == 53 and m2.griffon.n_reactions == 325


def test_error_paths():
    g = PyCombustionKernels()
    g.mechanism_set_element_mw_map({'H': 1.008})
    g.mechanism_add_element('H')
    g.mechanism_add_species('H2', {'H': 2.0})
    with pytest.raises(GriffonB200Error):
        g.mechanism_add_species('H2', {'H': 2.0})  # twice, chemistry_setup.cpp:40-43
    with pytest.raises(GriffonB200Error):
        g.mechanism_add_species('O2', {'O': 2.0})  # unknown atom, chemistry_setup.cpp:45-51
    g.mechanism_add_species('H', {'H': 1.0})
    g.mechanism_resize_heat_capacity_data()
    with pytest.raises(GriffonB200Error):
        g.mechanism_add_reaction_simple({'H2': 1}, {'OH': 2}, True, 1., 0., 0.)  # unknown species
    with pytest.raises(GriffonB200Error):
        g.mechanism_add_reaction_simple({'H2': 1}, {'H2': 1}, True, 1., 0., 0.)  # < 2 net species, :499-527


@pytest.mark.skipif(griffon.load_library().gb_cuda_device_count() > 0, reason='checks the no-GPU behaviour')
def test_no_cpu_fallback_without_gpu():
    m = build_mech('h2-burke', 'gpu')
    ns = m.n_species
    y = np.ones(ns) / ns
    with pytest.raises(GriffonB200Error, match='no CUDA device|CUDA'):
        m.griffon.cp_mix(1000., y)
    state = np.hstack([1000., y[:-1]])
    with pytest.raises(GriffonB200Error):
        m.griffon.reactor_rhs_isobaric(state, 101325., 0., np.zeros(1), 0., 0., 0., 0., 0., 0., 0, False,
                                       np.zeros(ns))


def test_host_side_flamelet_helpers_match_oracle():
    """flamelet_stencils / flamelet_jac_indices are host helpers of the C-ABI: bit-equal to the oracle"""
    mg, mo = build_mech('h2-burke', 'gpu'), build_mech('h2-burke', 'port')
    ns, nz = mg.n_species, 20
    rng = np.random.default_rng(0)
    z = np.sort(np.hstack([0, rng.uniform(0, 1, nz - 2), 1]))
    dz, nzi = z[1:] - z[:-1], nz - 2
    chi = rng.uniform(0.1, 3., nz)
    outs = []
    for k in (mg.griffon, mo.griffon):
        arrs = [np.zeros(nzi * ns), np.zeros(nzi * ns), np.zeros(nzi * ns), np.zeros(nzi), np.zeros(nzi)]
        k.flamelet_stencils(dz, nzi, chi, np.ones(ns), *arrs)
        rows = np.zeros(ns * (nzi * ns + 2 * (nzi - 1)), dtype=np.int32)
        cols = np.zeros_like(rows)
        k.flamelet_jac_indices(nzi, rows, cols)
        outs.append(arrs + [rows, cols])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_nasa9_mechanism_is_accepted_and_malformed_coefficients_are_refused():
    """NASA9 species are packed for the device (parity: tests/test_gpu_parity.py); a coefficient list that is not
    {nregions, (Tlo, Thi, a0..a8) * nregions} is an argument error, never a silent path"""
    md = load_mech_data('old_xmls_nasa9_air_h2')
    from spitfire_b200.mechanism import ChemicalMechanismSpec
    m = ChemicalMechanismSpec(mech_data=md)
    g = m.griffon
    name = m.species_names[0]
    with pytest.raises(GriffonB200Error):
        g.mechanism_add_nasa9_cp(name, 200., 6000., [2., 200., 1000., 1., 2., 3.])


def test_reference_names_outside_the_path_are_callable_and_raise():
    """griffon.pyx:870-987, 1019-1105: the 2-D flamelet methods and the block-Jacobi helpers exist and raise
    GriffonB200Error ("keep them callable ... or raise clearly")"""
    import pytest
    from spitfire_b200 import griffon
    for name in ('flamelet2d_rhs', 'flamelet2d_factored_block_diag_jacobian', 'flamelet2d_offdiag_matvec',
                 'flamelet2d_matvec', 'flamelet2d_block_diag_solve'):
        f = getattr(griffon.PyCombustionKernels, name)
        with pytest.raises(griffon.GriffonB200Error):
            f(object.__new__(griffon.PyCombustionKernels))
    for name in ('py_btddod_blockdiag_matvec', 'py_btddod_blockdiag_factorize', 'py_btddod_blockdiag_solve',
                 'py_btddod_lowerfulltriangle_solve', 'py_btddod_upperfulltriangle_solve',
                 'py_btddod_scale_and_add_scaled_block_diagonal'):
        with pytest.raises(griffon.GriffonB200Error):
            getattr(griffon, name)()
    for name in ('reactor_rhs_isochoric', 'reactor_jac_isochoric', 'reactor_rhs_isochoric_batch',
                 'reactor_jac_isochoric_batch'):
        assert callable(getattr(griffon.PyCombustionKernels, name))
