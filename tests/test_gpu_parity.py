"""GPU parity tests (run on the B200 box: pytest -m gpu). Every call goes through the C-ABI library
(spitfire_b200/libgriffon_b200.so) and is compared with the CPU oracle on the same seeded inputs.

Tolerance (FP64). BASELINE.json asks for 1e-12 relative on RHS / Jacobian entries. Entries of w and J are sums of
signed reaction contributions, so per-entry relative accuracy is bounded by cancellation, not by the implementation:
recompiling the *reference algorithm itself* with FMA contraction moves GRI-3.0 Jacobian entries by up to 1.3e-9
relative (median 3e-16), entries that cancel exactly in the reference (e.g. the density
derivative of first-order decomposition rows) become 1e-17-sized, and entries at underflow scale (1e-280) carry
sub-normal rounding. The bars asserted here, per output array and with scale = max|ref| over the state's output:
    median  |d|/|ref|                                               <= 1e-15
    every well-conditioned entry (|ref| >= 1e-3 scale): |d|/|ref|    <= 1e-12   (the north-star bar)
    every entry:                      |d| <= 1e-12 * (|ref| + 1e-3 scale)
and NaN/Inf patterns must be identical. The library is built without FMA contraction for this reason.
"""
import numpy as np
import pytest

from cases import (FLAGS, assemble_dense, block_thomas_all, call_all, error_stats, flamelet_all, flamelet_case,
                   random_case)
from common import (build_mech, golden_mech_names, has_nasa9, load_mech_data, oracle_available,
                    single_reaction_cases)
from spitfire_b200 import griffon
from spitfire_b200.synthetic import edge_mixtures, synthetic_states

pytestmark = pytest.mark.gpu

MECHS = golden_mech_names() + single_reaction_cases()
ORACLE = 'reference' if oracle_available('reference') else 'port'

MEDIAN_TOL, BIG_TOL, SCALED_TOL = 1e-15, 1e-12, 1e-12


GROSS_TOL = 3e-14

# fixtures of the random-composition sweep that are held to a looser bar than 1e-12, with the cancelling term
SWEEP_TOL = {}


def assert_parity(a, ref, what, big_tol=None, scaled_tol=None, gross=None):
    """gross (optional, same shape): sum of the absolute values of the terms the reference adds up to form the entry.
    An entry whose difference is below GROSS_TOL * gross (about a hundred ulp of the summands) is rounding noise of a
    cancelling sum -- the reference moves by as much when its own summation order changes -- and is taken as equal."""
    big_tol = BIG_TOL if big_tol is None else big_tol
    scaled_tol = SCALED_TOL if scaled_tol is None else scaled_tol
    a, ref = np.asarray(a), np.asarray(ref)
    assert a.shape == ref.shape
    assert np.array_equal(np.isfinite(a), np.isfinite(ref)), f'{what}: NaN/Inf pattern differs'
    fin = np.isfinite(ref)
    if not fin.all():
        a, ref = np.where(fin, a, 0.), np.where(fin, ref, 0.)
    if gross is not None:
        with np.errstate(all='ignore'):
            noise = np.abs(a - ref) <= GROSS_TOL * np.where(np.isfinite(gross), gross, 0.)
        a = np.where(noise, ref, a)
    s = error_stats(a, ref)
    assert s['strict_median'] <= MEDIAN_TOL, f'{what}: {s}'
    assert s['big_max'] <= big_tol, f'{what}: {s}'
    assert s['scaled_max'] <= scaled_tol, f'{what}: {s}'
    return s


def oracle_batch(o, ns, state, y, p, rho):
    n = state.shape[0]
    out = dict(rhs=np.zeros((n, ns)), jrhs=np.zeros((n, ns)), jac=np.zeros((n, ns * ns)), w=np.zeros((n, ns)),
               sens=np.zeros((n, (ns + 1) ** 2)))
    dummy = np.zeros(1)
    for i in range(n):
        o.reactor_rhs_isobaric(state[i], p, 0., dummy, 0., 0., 0., 0., 0., 0., 0, False, out['rhs'][i])
        o.reactor_jac_isobaric(state[i], p, 0., dummy, 0., 0., 0., 0., 0., 0., 0, False, 0, 0, out['jrhs'][i],
                               out['jac'][i])
        o.production_rates(state[i, 0], rho[i], y[i], out['w'][i])
        o.prod_rates_primitive_sensitivities(rho[i], state[i, 0], y[i], 0, out['sens'][i])
    return out


def temperature_row_gross(o, ns, state, y, rho, sens):
    """sum_i |h_i dw_i/dx| / (rho cp) behind every entry of the Jacobian's temperature row (the inner products of
    isobaric_reactor_kernels.cpp:74-92 and the chain rule :319-343), from the oracle's own sensitivities; zero for the
    other rows. Shape [n, ns*ns] like the column-major Jacobian."""
    n = state.shape[0]
    out = np.zeros((n, ns * ns))
    h = np.zeros(ns)
    mw = np.array([o.mixture_molecular_weight(np.eye(ns)[i]) for i in range(ns)])
    for s in range(n):
        T = state[s, 0]
        o.species_enthalpies(T, h)
        cp = o.cp_mix(T, y[s])
        mmw = o.mixture_molecular_weight(y[s])
        ws = np.abs(sens[s].reshape(ns + 1, ns + 1).T[:ns, :])  # [i, col]: col 0 rho, 1 T, 2+k Y_k
        g = (np.abs(h)[:, None] * ws).sum(axis=0) / (rho[s] * cp)
        row = np.zeros(ns)
        row[0] = g[1] + rho[s] / T * g[0]
        for k in range(ns - 1):
            row[1 + k] = g[2 + k] + rho[s] * mmw * abs(1. / mw[k] - 1. / mw[ns - 1]) * g[0]
        out[s, 0::ns] = row
    return out


def temperature_rhs_gross(o, ns, state, y, rho, w):
    """sum_i |h_i w_i| / (rho cp) behind the temperature entry of the right-hand side (the inner product of
    isobaric_reactor_kernels.cpp:22); zero for the species entries. Shape [n, ns]. A one-ulp difference in a rate
    constant (libm vs CUDA exp) moves the entry by about 1e-16 of this sum, whatever is left of it after cancellation:
    the reference's single-reaction test mechanisms have heats of reaction that cancel to zero."""
    n = state.shape[0]
    out = np.zeros((n, ns))
    h = np.zeros(ns)
    for s in range(n):
        o.species_enthalpies(state[s, 0], h)
        out[s, 0] = np.sum(np.abs(h * w[s])) / (rho[s] * o.cp_mix(state[s, 0], y[s]))
    return out


def gpu_batch(g, ns, state, y, p, rho):
    n = state.shape[0]
    out = dict(rhs=np.zeros((n, ns)), jrhs=np.zeros((n, ns)), jac=np.zeros((n, ns * ns)), w=np.zeros((n, ns)),
               sens=np.zeros((n, (ns + 1) ** 2)))
    T = np.ascontiguousarray(state[:, 0])
    g.reactor_rhs_isobaric_batch(state, p, out['rhs'])
    g.reactor_jac_isobaric_batch(state, p, out['jrhs'], out['jac'])
    g.production_rates_batch(T, rho, y, out['w'])
    g.prod_rates_sens_batch(rho, T, y, 0, out['sens'])
    return out


def random_states(ns, n, rng, Tlo=250., Thi=3800.):
    y = rng.dirichlet(np.ones(ns) * 0.5, n)
    T = rng.uniform(Tlo, Thi, n)
    return np.ascontiguousarray(np.hstack([T[:, None], y[:, :-1]])), y


@pytest.mark.parametrize('name', MECHS)
def test_reactor_and_rates_parity_all_fixture_mechanisms(name):
    """the 34 old_xmls fixtures + h2-burke + GRI-3.0 + lu30 + heptane, T from below Tmin to above Tmax, 1/2/10 atm.
    The one- and two-reaction fixtures at random (non-physical) compositions have Jacobian entries in which the
    reactant term and the eliminated-last-species term of dq/dY cancel; the kernel sums these two parts in a different
    order than the reference ("dense part as two scalars"), so the bar for this sweep is 1e-11 on
    entries above 1e-3 of the scale; BASELINE's own mechanisms are held to 1e-12 in the tests below."""
    mg, mo = build_mech(name, 'gpu'), build_mech(name, ORACLE)
    ns = mg.n_species
    rng = np.random.default_rng(11)
    n = 96 if ns > 20 else 192
    for p in (101325., 2 * 101325., 10 * 101325.):
        state, y = random_states(ns, n, rng)
        rho = np.array([mo.griffon.ideal_gas_density(p, state[i, 0], y[i]) for i in range(n)])
        ref = oracle_batch(mo.griffon, ns, state, y, p, rho)
        got = gpu_batch(mg.griffon, ns, state, y, p, rho)
        gross = {'jac': temperature_row_gross(mo.griffon, ns, state, y, rho, ref['sens'])}
        gross['rhs'] = gross['jrhs'] = temperature_rhs_gross(mo.griffon, ns, state, y, rho, ref['w'])
        tol = SWEEP_TOL.get(name, (1e-12, ''))[0]
        for k in ref:
            assert_parity(got[k], ref[k], f'{name} p={p} {k}', big_tol=tol, scaled_tol=tol, gross=gross.get(k))


@pytest.mark.parametrize('name,fuel', [('h2-burke', 'H2'), ('methane-gri30', 'CH4')])
def test_synthetic_batch_parity_subset(name, fuel):
    """BASELINE configs 2-3: first 2048 states of the seeded synthetic batch through the CPU oracle"""
    mg, mo = build_mech(name, 'gpu'), build_mech(name, ORACLE)
    ns = mg.n_species
    n = 2048 if ns < 20 else 1024
    state, y = synthetic_states(mg.species_names, n, fuel)
    p = 101325.
    rho = np.array([mo.griffon.ideal_gas_density(p, state[i, 0], y[i]) for i in range(n)])
    ref = oracle_batch(mo.griffon, ns, state, y, p, rho)
    got = gpu_batch(mg.griffon, ns, state, y, p, rho)
    report = {}
    gross = temperature_row_gross(mo.griffon, ns, state, y, rho, ref['sens'])
    for k in ref:
        s = assert_parity(got[k], ref[k], f'{name} synthetic {k}', gross=gross if k == 'jac' else None)
        report[k] = s
        print(name, k, s)
    import json
    import os
    os.makedirs('gpurun_out', exist_ok=True)
    with open(os.path.join('gpurun_out', f'parity_{name}.json'), 'w') as f:
        json.dump(dict(mechanism=name, oracle=ORACLE, states=n, build=griffon.load_library().gb_build_info().decode(),
                       stats=report), f, indent=1)


@pytest.mark.parametrize('name', ['h2-burke', 'methane-gri30', 'old_xmls_rev_troe4_withN_withNTB'])
def test_edge_mixtures(name):
    """trace (1e-8, 1e-16) and exactly-zero species, single-species mixtures (test_reaction_rates.py:22-34)"""
    mg, mo = build_mech(name, 'gpu'), build_mech(name, ORACLE)
    ns = mg.n_species
    ys = edge_mixtures(ns)
    for T in (300., 1400., 2500.):
        n = ys.shape[0]
        state = np.ascontiguousarray(np.hstack([np.full((n, 1), T), ys[:, :-1]]))
        p = 101325.
        # the oracle sees y rebuilt from the state exactly like the kernels do
        rho = np.array([mo.griffon.ideal_gas_density(p, T, ys[i]) for i in range(n)])
        ref = oracle_batch(mo.griffon, ns, state, ys, p, rho)
        got = gpu_batch(mg.griffon, ns, state, ys, p, rho)
        for k in ref:
            assert_parity(got[k], ref[k], f'{name} edge T={T} {k}')


@pytest.mark.parametrize('name', ['h2-burke', 'methane-gri30', 'old_xmls_rev_lindemann_withN_withNTB'])
def test_open_isothermal_diathermal_reactor(name):
    mg, mo = build_mech(name, 'gpu'), build_mech(name, ORACLE)
    ns = mg.n_species
    rng = np.random.default_rng(2)
    n = 40
    state, y = random_states(ns, n, rng, 500., 2500.)
    yin = rng.dirichlet(np.ones(ns))
    p = 2 * 101325.
    for heat_option in (0, 1, 2):
        args = (p, 900., yin, 1e-3, 400., 500., 10., 0.3, 2.5, heat_option, True)
        ref_r, ref_jr, ref_j = np.zeros((n, ns)), np.zeros((n, ns)), np.zeros((n, ns * ns))
        for i in range(n):
            mo.griffon.reactor_rhs_isobaric(state[i], *args, ref_r[i])
            mo.griffon.reactor_jac_isobaric(state[i], *args, 2, 0, ref_jr[i], ref_j[i])
        r, jr, j = np.zeros((n, ns)), np.zeros((n, ns)), np.zeros((n, ns * ns))
        mg.griffon.reactor_rhs_isobaric_batch(state, p, r, *args[1:])
        mg.griffon.reactor_jac_isobaric_batch(state, p, jr, j, *args[1:], 2, 0)
        assert_parity(r, ref_r, f'{name} open rhs heat={heat_option}')
        assert_parity(jr, ref_jr, f'{name} open jac-rhs heat={heat_option}')
        assert_parity(j, ref_j, f'{name} open jac heat={heat_option}')


@pytest.mark.parametrize('name', ['h2-burke', 'methane-gri30', 'old_xmls_rev_lindemann_withN_withNTB',
                                  'old_xmls_nasa9_air_h2'])
def test_isochoric_reactor(name):
    """reactor_{rhs,jac}_isochoric (isochoric_reactor_kernels.cpp:192-335): state [rho, T, Y], (ns+1)^2 Jacobian;
    closed and open, the three heat-transfer options; batch entry points and the single-state method names"""
    mg, mo = build_mech(name, 'gpu'), build_mech(name, ORACLE)
    ns = mg.n_species
    rng = np.random.default_rng(6)
    n = 40
    y = rng.dirichlet(np.ones(ns) * 0.5, n)
    state = np.ascontiguousarray(np.hstack([rng.uniform(0.05, 4., (n, 1)), rng.uniform(500., 2500., (n, 1)), y[:, :-1]]))
    yin = rng.dirichlet(np.ones(ns))
    for open_ in (False, True):
        for heat_option in (0, 1, 2):
            args = (1.3, 900., yin, 1e-3, 400., 500., 10., 0.3, 2.5, heat_option, open_)
            ref_r, ref_jr, ref_j = np.zeros((n, ns + 1)), np.zeros((n, ns + 1)), np.zeros((n, (ns + 1) ** 2))
            for i in range(n):
                mo.griffon.reactor_rhs_isochoric(state[i], *args, ref_r[i])
                mo.griffon.reactor_jac_isochoric(state[i], *args, 0, ref_jr[i], ref_j[i])
            r, jr, j = np.zeros((n, ns + 1)), np.zeros((n, ns + 1)), np.zeros((n, (ns + 1) ** 2))
            mg.griffon.reactor_rhs_isochoric_batch(state, r, *args)
            mg.griffon.reactor_jac_isochoric_batch(state, jr, j, *args, 0)
            what = f'{name} isochoric open={open_} heat={heat_option}'
            assert_parity(r, ref_r, what + ' rhs')
            assert_parity(jr, ref_jr, what + ' jac-rhs')
            assert_parity(j, ref_j, what + ' jac')
            r1, j1 = np.zeros(ns + 1), np.zeros((ns + 1) ** 2)
            mg.griffon.reactor_jac_isochoric(state[3], *args, 0, r1, j1)
            assert np.array_equal(r1, jr[3]) and np.array_equal(j1, j[3])


def test_single_state_api_is_the_batch_path():
    """the reference's single-state method names are batch-of-one calls of the same kernels"""
    mg, mo = build_mech('h2-burke', 'gpu'), build_mech('h2-burke', ORACLE)
    ns = mg.n_species
    rng = np.random.default_rng(4)
    for _ in range(3):
        T, p, y = random_case(ns, rng)
        yin = rng.dirichlet(np.ones(ns))
        got, ref = call_all(mg.griffon, ns, T, p, y, yin), call_all(mo.griffon, ns, T, p, y, yin)
        for k in ref:
            d = np.abs(got[k] - ref[k])
            scale = np.max(np.abs(ref[k])) if ref[k].size else 1.
            assert np.all(d <= SCALED_TOL * (np.abs(ref[k]) + 1e-3 * scale)), k


@pytest.mark.parametrize('name', ['h2-burke', 'methane-gri30'])
def test_empty_ragged_and_tile_independence(name):
    """n = 0 is a no-op; results do not depend on batch size or on the position of a state inside a CTA tile"""
    mg = build_mech(name, 'gpu')
    g, ns = mg.griffon, mg.n_species
    state, y = synthetic_states(mg.species_names, 257, 'H2' if ns < 20 else 'CH4')
    p = 101325.
    g.reactor_rhs_isobaric_batch(np.zeros((0, ns)), p, np.zeros((0, ns)))
    g.reactor_jac_isobaric_batch(np.zeros((0, ns)), p, np.zeros((0, ns)), np.zeros((0, ns * ns)))
    full_r, full_j = np.zeros((257, ns)), np.zeros((257, ns * ns))
    g.reactor_jac_isobaric_batch(state, p, full_r, full_j)
    rr = np.zeros((257, ns))
    g.reactor_rhs_isobaric_batch(state, p, rr)
    for n in (1, 2, 7, 8, 15, 33, 100):
        off = 5
        r, j = np.zeros((n, ns)), np.zeros((n, ns * ns))
        g.reactor_jac_isobaric_batch(np.ascontiguousarray(state[off:off + n]), p, r, j)
        assert np.array_equal(r, full_r[off:off + n]) and np.array_equal(j, full_j[off:off + n])
        r2 = np.zeros((n, ns))
        g.reactor_rhs_isobaric_batch(np.ascontiguousarray(state[off:off + n]), p, r2)
        assert np.array_equal(r2, rr[off:off + n])


@pytest.mark.parametrize('name,fuel,n', [('h2-burke', 'H2', 1 << 20), ('methane-gri30', 'CH4', 1 << 18)])
def test_full_size_batch_properties(name, fuel, n):
    """BASELINE-size batches on the device: finite, identical to the small-batch evaluation (prefix), the Jacobian
    kernel's RHS agrees with the RHS kernel, and J v matches a directional finite difference of the RHS"""
    import torch
    mg = build_mech(name, 'gpu')
    g, ns = mg.griffon, mg.n_species
    state, _ = synthetic_states(mg.species_names, n, fuel)
    p = 101325.
    d_state = torch.from_numpy(state).cuda()
    d_rhs = torch.empty((n, ns), dtype=torch.float64, device='cuda')
    d_jrhs = torch.empty_like(d_rhs)
    d_jac = torch.empty((n, ns * ns), dtype=torch.float64, device='cuda')
    g.reactor_rhs_isobaric_batch(d_state, p, d_rhs)
    g.reactor_jac_isobaric_batch(d_state, p, d_jrhs, d_jac)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(d_rhs).all()) and bool(torch.isfinite(d_jac).all())
    k = 4096
    r, j = np.zeros((k, ns)), np.zeros((k, ns * ns))
    g.reactor_jac_isobaric_batch(np.ascontiguousarray(state[:k]), p, r, j)
    assert np.array_equal(j, d_jac[:k].cpu().numpy()) and np.array_equal(r, d_jrhs[:k].cpu().numpy())
    # tail of the batch too (last, partially filled tile)
    r, j = np.zeros((k, ns)), np.zeros((k, ns * ns))
    g.reactor_jac_isobaric_batch(np.ascontiguousarray(state[-k:]), p, r, j)
    assert np.array_equal(j, d_jac[-k:].cpu().numpy())
    # the two RHS evaluations use different but equivalent expressions for k_r and the Troe factor (SURVEY H2:
    # k_f/K_c vs k_f*exp(..), pow(F_cent, g) vs 10^..): they agree to rounding except where q = R_f - R_r cancels
    a, b = d_jrhs.cpu().numpy(), d_rhs.cpu().numpy()
    s = error_stats(a, b)
    assert s['strict_median'] <= MEDIAN_TOL and s['strict_p999'] <= 1e-11 and s['scaled_max'] <= 1e-9, s
    # directional derivative on a subset
    m = 2048
    rng = np.random.default_rng(0)
    v = rng.normal(size=(m, ns)) * np.abs(state[:m]) * 1e-7
    rp, rm = np.zeros((m, ns)), np.zeros((m, ns))
    g.reactor_rhs_isobaric_batch(np.ascontiguousarray(state[:m] + v), p, rp)
    g.reactor_rhs_isobaric_batch(np.ascontiguousarray(state[:m] - v), p, rm)
    J = d_jac[:m].cpu().numpy().reshape(m, ns, ns).transpose(0, 2, 1)
    jv = np.einsum('nij,nj->ni', J, v)
    fd = 0.5 * (rp - rm)
    scale = np.max(np.abs(jv), axis=1, keepdims=True) + 1e-300
    assert np.quantile(np.abs(jv - fd) / scale, 0.99) < 1e-4


def test_thermo_batch_helpers():
    mg, mo = build_mech('methane-gri30', 'gpu'), build_mech('methane-gri30', ORACLE)
    g, o, ns = mg.griffon, mo.griffon, mg.n_species
    rng = np.random.default_rng(9)
    n = 64
    y = rng.dirichlet(np.ones(ns), n)
    T = rng.uniform(150., 6500., n)  # below Tmin and above Tmax of several species
    out = np.zeros(n)
    g.thermo_batch(3, T, y, out)
    np.testing.assert_allclose(out, [o.cp_mix(T[i], y[i]) for i in range(n)], rtol=1e-14)
    g.thermo_batch(5, T, y, out)
    np.testing.assert_allclose(out, [o.enthalpy_mix(T[i], y[i]) for i in range(n)], rtol=1e-13, atol=1e-9)
    outs = np.zeros((n, ns))
    g.thermo_batch(9, T, y, outs)
    ref = np.zeros((n, ns))
    for i in range(n):
        o.species_enthalpies(T[i], ref[i])
    np.testing.assert_allclose(outs, ref, rtol=1e-14, atol=1e-7)


# ---- flamelet ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name,nz', [('h2-burke', 34), ('methane-gri30', 40), ('old_xmls_rev_troe4_withN_withNTB', 12)])
def test_flamelet_rhs_and_jacobian_parity(name, nz):
    mg, mo = build_mech(name, 'gpu'), build_mech(name, ORACLE)
    c = flamelet_case(mg, nz)
    got, ref = flamelet_all(mg.griffon, c, eig=False), flamelet_all(mo.griffon, c, eig=False)
    for k in ref:
        if k.startswith(('rhs', 'jac')):
            assert_parity(got[k], ref[k], f'{name} flamelet {k}')
        else:
            assert np.array_equal(got[k], ref[k]), k


def test_flamelet_batch_with_per_flamelet_dissipation_rates():
    """F flamelets in one launch, each with its own chi / stencil coefficients (the sweep's layout)"""
    import torch
    mg, mo = build_mech('h2-burke', 'gpu'), build_mech('h2-burke', ORACLE)
    g, o = mg.griffon, mo.griffon
    c = flamelet_case(mg, 24)
    ns, nzi = c['ns'], c['nzi']
    F = 5
    rng = np.random.default_rng(1)
    chis = np.array([c['chi'] * s for s in (0.5, 1., 2., 4., 8.)])
    states = np.array([c['state'] * (1 + 0.01 * rng.uniform(size=c['state'].size)) for _ in range(F)])
    Tc, Tr, hc, hr = c['heat']
    coeffs = [[np.zeros(nzi * ns) for _ in range(3)] + [np.zeros(nzi), np.zeros(nzi)] for _ in range(F)]
    ref_r, ref_j = [], []
    nj = ns * (nzi * ns + 2 * (nzi - 1))
    for f in range(F):
        o.flamelet_stencils(c['dz'], nzi, chis[f], np.ones(ns), *coeffs[f])
        cm, cs, cu, mc, nc = coeffs[f]
        r, J = np.zeros(nzi * ns), np.zeros(nj)
        o.flamelet_rhs(states[f], 101325., c['oxy'], c['fuel'], False, Tc, Tr, hc, hr, nzi, cm, cs, cu, mc, nc,
                       chis[f], True, True, True, r)
        o.flamelet_jacobian(states[f], 101325., c['oxy'], c['fuel'], False, Tc, Tr, hc, hr, nzi, cm, cs, cu, mc, nc,
                            chis[f], False, 0., True, 2e-8, 0, 0, True, True, True, np.zeros(nzi * ns), J)
        ref_r.append(r)
        ref_j.append(J)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d = dict(oxy=t(c['oxy']), fuel=t(c['fuel']), Tc=t(Tc), Tr=t(Tr), hc=t(hc), hr=t(hr),
             cm=t(np.array([x[0] for x in coeffs])), cs=t(np.array([x[1] for x in coeffs])),
             cu=t(np.array([x[2] for x in coeffs])), mc=t(coeffs[0][3]), nc=t(coeffs[0][4]), chi=t(chis))
    prm = g._flamelet_params(101325., d['oxy'], d['fuel'], False, d['Tc'], d['Tr'], d['hc'], d['hr'], nzi, d['cm'],
                             d['cs'], d['cu'], d['mc'], d['nc'], d['chi'], True, True, True,
                             strides=(0, nzi * ns, 0, nzi + 2))
    d_state = t(states)
    d_rhs = torch.zeros_like(d_state)
    d_jac = torch.zeros((F, nj), dtype=torch.float64, device='cuda')
    g.flamelet_rhs_batch(F, d_state, prm, d_rhs)
    g.flamelet_jacobian_batch(F, d_state, prm, d_jac, scale_and_offset=True, prefactor=2e-8)
    torch.cuda.synchronize()
    assert_parity(d_rhs.cpu().numpy(), np.array(ref_r), 'batched flamelet rhs')
    assert_parity(d_jac.cpu().numpy(), np.array(ref_j), 'batched flamelet jac')


# ---- block Thomas -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name,nz', [('h2-burke', 34), ('methane-gri30', 24)])
def test_block_thomas_parity(name, nz):
    """factor / solve / matvec / scale-add against the oracle (LAPACK dgetrf/dgetrs) and a dense solve.
    LU-level agreement is limited by the conditioning of the blocks, so the bars are on the solution and residual:
    identical pivot sequence, |x_gpu - x_ref| <= 1e-9 max|x|, residual no worse than the oracle's (x10)."""
    mg, mo = build_mech(name, 'gpu'), build_mech(name, ORACLE)
    c = flamelet_case(mg, nz)
    ns, nzi = c['ns'], c['nzi']
    A0 = flamelet_all(mo.griffon, c, eig=False)['jac0True']
    ref = block_thomas_all(mo.griffon, A0, c['rhs'], nzi, ns)
    got = block_thomas_all(griffon, A0, c['rhs'], nzi, ns)
    assert np.array_equal(got['piv'], ref['piv'])
    assert np.array_equal(got['B'], ref['B'])
    xs = np.max(np.abs(ref['x']))
    assert np.max(np.abs(got['x'] - ref['x'])) <= 1e-9 * xs
    M = assemble_dense(A0, nzi, ns)
    res_ref = np.max(np.abs(M @ ref['x'] - c['rhs']))
    res_got = np.max(np.abs(M @ got['x'] - c['rhs']))
    assert res_got <= 10 * res_ref + 1e-12 * np.max(np.abs(c['rhs']))
    np.testing.assert_allclose(got['mv'], ref['mv'], rtol=1e-9, atol=1e-9 * np.max(np.abs(c['rhs'])))
    s = error_stats(got['A'][:nzi * ns * ns], ref['A'][:nzi * ns * ns])
    assert s['scaled_max'] <= 1e-8, s


def test_block_thomas_batched_systems_match_single():
    import torch
    mg, mo = build_mech('h2-burke', 'gpu'), build_mech('h2-burke', ORACLE)
    c = flamelet_case(mg, 20)
    ns, nzi = c['ns'], c['nzi']
    base = flamelet_all(mo.griffon, c, eig=False)
    mats = np.array([base['jac0True'], base['jac1True'], base['jac2True'], base['jac3True']])
    rng = np.random.default_rng(0)
    rhs = rng.normal(size=(4, nzi * ns))
    singles = [block_thomas_all(griffon, mats[f], rhs[f], nzi, ns) for f in range(4)]
    dA = torch.from_numpy(mats.copy()).cuda()
    dL = torch.zeros((4, nzi * ns * ns), dtype=torch.float64, device='cuda')
    dP = torch.zeros((4, nzi * ns), dtype=torch.int32, device='cuda')
    dX = torch.zeros((4, nzi * ns), dtype=torch.float64, device='cuda')
    dR = torch.from_numpy(rhs).cuda()
    griffon.py_btddod_full_factorize(dA, nzi, ns, dL, dP, n_systems=4)
    griffon.py_btddod_full_solve(dA, dL, dP, dR, nzi, ns, dX, n_systems=4)
    torch.cuda.synchronize()
    for f in range(4):
        assert np.array_equal(dX[f].cpu().numpy(), singles[f]['x'])
        assert np.array_equal(dA[f].cpu().numpy(), singles[f]['A'])
        assert np.array_equal(dP[f].cpu().numpy(), singles[f]['piv'])
