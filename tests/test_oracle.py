"""The CPU oracle (oracle/griffon_oracle.c, the plain-C restatement) is pinned here:
 (1) bit-identical to the unmodified reference C++ (oracle/_ref, built from /root/reference) on every golden mechanism
     the reference's own tests use, for every function of the path;
 (2) the reference's own known-answer procedures that need no Cantera: analytic Jacobian vs central finite differences
     (tests/griffon/test_rhsjac_isobaric_closed_adiabatic.py:39-61), rate sensitivities vs finite differences
     (tests/griffon/test_reaction_rates.py:68-97), dense == sparse (tests/griffon/test_jac_dense_vs_sparse.py);
 (3) block-Thomas against a dense solve of the assembled matrix.
Gold-file pins that need the host solver loops live in test_reactor_host.py / test_flamelet_host.py."""
import numpy as np
import pytest

from common import (build_mech, golden_mech_names, has_nasa9, load_mech_data, oracle_available,
                    single_reaction_cases)
from cases import FLAGS, assemble_dense, block_thomas_all, call_all, flamelet_all, flamelet_case, random_case

MECHS = golden_mech_names() + single_reaction_cases()
need_ref = pytest.mark.skipif(not oracle_available('reference'), reason='oracle/_ref not built (no /root/reference)')


@need_ref
@pytest.mark.parametrize('name', MECHS)
def test_port_is_bit_identical_to_reference(name):
    a, b = build_mech(name, 'reference'), build_mech(name, 'port')
    ns = a.n_species
    rng = np.random.default_rng(1)
    n_trials = 6 if ns > 20 else 20
    for trial in range(n_trials):
        T, p, y = random_case(ns, rng)
        if trial % 5 == 4:
            y[rng.integers(ns)] = 0.
            y /= y.sum()
        yin = rng.dirichlet(np.ones(ns))
        oa, ob = call_all(a.griffon, ns, T, p, y, yin), call_all(b.griffon, ns, T, p, y, yin)
        for k in oa:
            assert np.array_equal(oa[k], ob[k]), f'{name}: {k} differs from the reference'
        assert np.array_equal(oa['s'], oa['s2'])  # dense == sparse (test_jac_dense_vs_sparse.py:81 allows 1e-10)


@need_ref
@pytest.mark.parametrize('name', ['h2-burke', 'methane-gri30', 'old_xmls_nasa9_air_h2',
                                  'old_xmls_rev_troe4_withN_withNTB', 'old_xmls_irr_elementary_noN_noT_generalstoich'])
def test_port_isochoric_reactor_bit_identical_to_reference(name):
    """isochoric_reactor_kernels.cpp:192-335: closed/open, adiabatic/isothermal/diathermal"""
    a, b = build_mech(name, 'reference'), build_mech(name, 'port')
    ns = a.n_species
    rng = np.random.default_rng(3)
    for trial in range(6):
        y = rng.dirichlet(np.ones(ns) * 0.5)
        state = np.concatenate([[rng.uniform(0.05, 5.), rng.uniform(300., 3000.)], y[:-1]])
        yin = rng.dirichlet(np.ones(ns))
        for heat in (0, 1, 2):
            for open_ in (False, True):
                args = (state, 1.1, 350., yin, 1e-3, 400., 500., 10., 0.3, 2.0, heat, open_)
                r = [np.zeros(ns + 1) for _ in range(4)]
                j = [np.zeros((ns + 1) ** 2) for _ in range(2)]
                a.griffon.reactor_rhs_isochoric(*args, r[0])
                b.griffon.reactor_rhs_isochoric(*args, r[1])
                a.griffon.reactor_jac_isochoric(*args, 0, r[2], j[0])
                b.griffon.reactor_jac_isochoric(*args, 0, r[3], j[1])
                assert np.array_equal(r[0], r[1]) and np.array_equal(r[2], r[3]) and np.array_equal(j[0], j[1])


@need_ref
@pytest.mark.parametrize('name,nz', [('h2-burke', 34), ('methane-gri30', 16), ('old_xmls_rev_troe4_withN_withNTB', 12)])
def test_port_flamelet_and_block_thomas_bit_identical_to_reference(name, nz):
    a, b = build_mech(name, 'reference'), build_mech(name, 'port')
    c = flamelet_case(a, nz)
    oa, ob = flamelet_all(a.griffon, c), flamelet_all(b.griffon, c)
    for k in oa:
        assert np.array_equal(oa[k], ob[k]), f'{name}: flamelet {k} differs from the reference'
    A0 = oa['jac0True']
    ta = block_thomas_all(a.griffon, A0, c['rhs'], c['nzi'], c['ns'])
    tb = block_thomas_all(b.griffon, A0, c['rhs'], c['nzi'], c['ns'])
    for k in ta:
        assert np.array_equal(ta[k], tb[k]), f'{name}: block-Thomas {k} differs from the reference'


@pytest.mark.parametrize('name,nz', [('h2-burke', 34), ('methane-gri30', 12)])
def test_port_block_thomas_matches_dense_solve(name, nz):
    m = build_mech(name, 'port')
    c = flamelet_case(m, nz)
    A0 = flamelet_all(m.griffon, c, eig=False)['jac0True']  # gamma*dt*J - I: well conditioned
    t = block_thomas_all(m.griffon, A0, c['rhs'], c['nzi'], c['ns'])
    M = assemble_dense(A0, c['nzi'], c['ns'])
    x = np.linalg.solve(M, c['rhs'])
    np.testing.assert_allclose(t['x'], x, rtol=1e-9, atol=1e-12 * np.max(np.abs(x)))
    np.testing.assert_allclose(t['mv'], M @ t['x'], rtol=1e-9, atol=1e-9 * np.max(np.abs(c['rhs'])))
    # jac indices describe the same matrix
    rows = np.zeros(A0.size, dtype=np.int32)
    cols = np.zeros_like(rows)
    m.griffon.flamelet_jac_indices(c['nzi'], rows, cols)
    M2 = np.zeros_like(M)
    M2[rows, cols] = A0
    assert np.array_equal(M, M2)


@pytest.mark.parametrize('name', ['h2-burke', 'methane-gri30', 'old_xmls_rev_troe4_withN_withNTB',
                                  'old_xmls_rev_lindemann_withN_withNTB', 'old_xmls_rev_3body_withN_withN_nonUnity',
                                  'old_xmls_irr_elementary_noN_noT_generalstoich'])
def test_port_jacobian_matches_finite_differences(name):
    """the reference's own procedure, tests/griffon/test_rhsjac_isobaric_closed_adiabatic.py:39-61 (tolerance 1e-2)"""
    m = build_mech(name, 'port')
    g, ns = m.griffon, m.n_species
    rng = np.random.default_rng(5)
    for T in (600., 1200.):
        for p in (101325., 2 * 101325.):
            y = rng.dirichlet(np.ones(ns))
            st = np.hstack([T, y[:-1]])
            args = (p, 0., np.zeros(1), 0., 0., 0., 0., 0., 0., 0, False)
            rhs, jac = np.zeros(ns), np.zeros(ns * ns)
            g.reactor_jac_isobaric(st, *args, 0, 0, rhs, jac)
            J = jac.reshape(ns, ns).T
            Jfd = np.zeros((ns, ns))
            for k in range(ns):
                h = 1e-6 * max(abs(st[k]), 1e-3)
                sp, sm = st.copy(), st.copy()
                sp[k] += h
                sm[k] -= h
                rp, rm = np.zeros(ns), np.zeros(ns)
                g.reactor_rhs_isobaric(sp, *args, rp)
                g.reactor_rhs_isobaric(sm, *args, rm)
                Jfd[:, k] = (rp - rm) / (2 * h)
            scale = np.max(np.abs(Jfd)) + 1.
            assert np.max(np.abs(J - Jfd)) / scale < 1e-2


def test_port_sensitivities_match_finite_differences():
    """tests/griffon/test_reaction_rates.py:68-97 (tolerance 1e-2) on the rho and T columns"""
    m = build_mech('h2-burke', 'port')
    g, ns = m.griffon, m.n_species
    rng = np.random.default_rng(7)
    y = rng.dirichlet(np.ones(ns))
    T, rho = 1300., 0.4
    s = np.zeros((ns + 1) ** 2)
    g.prod_rates_primitive_sensitivities(rho, T, y, 0, s)
    S = s.reshape(ns + 1, ns + 1).T

    def w(T_, rho_):
        o = np.zeros(ns)
        g.production_rates(T_, rho_, y, o)
        return o

    dT, dr = 1e-3, 1e-7
    np.testing.assert_allclose(S[:ns, 1], (w(T + dT, rho) - w(T - dT, rho)) / (2 * dT), rtol=1e-2,
                               atol=1e-6 * np.max(np.abs(S[:ns, 1])))
    np.testing.assert_allclose(S[:ns, 0], (w(T, rho + dr) - w(T, rho - dr)) / (2 * dr), rtol=1e-2,
                               atol=1e-6 * np.max(np.abs(S[:ns, 0])))
    assert np.all(S[ns, :] == 0.)
