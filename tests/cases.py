"""test-case builders shared by the oracle tests (CPU) and the parity tests (GPU): the same calls, made through the
reference's method names, on any backend object (`mech.griffon`)"""
import numpy as np


def random_case(ns, rng):
    y = rng.dirichlet(np.ones(ns) * 0.5)
    T = rng.uniform(250., 3800.)
    p = 101325. * rng.choice([1, 2, 10])
    return T, p, y


def call_all(g, ns, T, p, y, yin):
    """every single-state thermo / kinetics / reactor method of the path"""
    rho = g.ideal_gas_density(p, T, y)
    w = np.zeros(ns)
    g.production_rates(T, rho, y, w)
    s = np.zeros((ns + 1) ** 2)
    g.prod_rates_primitive_sensitivities(rho, T, y, 0, s)
    s2 = np.zeros((ns + 1) ** 2)
    g.prod_rates_primitive_sensitivities(rho, T, y, 2, s2)
    st = np.hstack([T, y[:-1]])
    r1, r2, j = np.zeros(ns), np.zeros(ns), np.zeros(ns * ns)
    g.reactor_rhs_isobaric(st, p, 0., np.zeros(1), 0., 0., 0., 0., 0., 0., 0, False, r1)
    g.reactor_jac_isobaric(st, p, 0., np.zeros(1), 0., 0., 0., 0., 0., 0., 0, False, 0, 0, r2, j)
    r3, j3, r4 = np.zeros(ns), np.zeros(ns * ns), np.zeros(ns)
    g.reactor_jac_isobaric(st, p, 900., yin, 1e-3, 400., 500., 10., 0.3, 2.5, 2, True, 2, 0, r3, j3)
    g.reactor_rhs_isobaric(st, p, 900., yin, 1e-3, 400., 500., 10., 0.3, 2.5, 2, True, r4)
    r5, j5 = np.zeros(ns), np.zeros(ns * ns)
    g.reactor_jac_isobaric(st, p, 900., yin, 1e-3, 400., 500., 10., 0.3, 2.5, 1, True, 0, 0, r5, j5)
    h, cp, d, e, cv, x = (np.zeros(ns) for _ in range(6))
    g.species_enthalpies(T, h)
    g.species_cp(T, cp)
    g.dcpdT_species(T, y, d)
    g.species_energies(T, e)
    g.species_cv(T, cv)
    g.mole_fractions(y, x)
    sc = np.array([g.cp_mix(T, y), g.cv_mix(T, y), g.enthalpy_mix(T, y), g.energy_mix(T, y),
                   g.mixture_molecular_weight(y), g.ideal_gas_pressure(rho, T, y), rho])
    return dict(w=w, s=s, s2=s2, r1=r1, r2=r2, j=j, r3=r3, j3=j3, r4=r4, r5=r5, j5=j5, h=h, cp=cp, d=d, e=e, cv=cv,
                x=x, sc=sc)


def flamelet_case(m, nz, seed=3):
    ns = m.n_species
    rng = np.random.default_rng(seed)
    z = np.sort(np.hstack([0, rng.uniform(0, 1, nz - 2), 1]))
    dz, nzi = z[1:] - z[:-1], nz - 2
    chi = 5.0 * np.exp(-2 * (z - 0.5) ** 2) * (z * (1 - z)) ** 0.5 + 0.01

    def rstate(T):
        y = rng.dirichlet(np.ones(ns))
        return np.hstack([T, y[:-1]])

    oxy, fuel = rstate(300.), rstate(350.)
    state = np.hstack([rstate(300 + 1800 * np.sin(np.pi * zz) ** 2) for zz in z[1:-1]])
    heat = (np.full(nzi, 350.), np.full(nzi, 320.), rng.uniform(1, 5, nzi) * 1e3, rng.uniform(0, 1, nzi))
    return dict(ns=ns, nz=nz, nzi=nzi, dz=dz, chi=chi, oxy=oxy, fuel=fuel, state=state, heat=heat,
                rhs=rng.normal(size=nzi * ns))


# (adiabatic, include_enthalpy_flux, include_variable_cp, use_scaled_heat_loss)
FLAGS = [(True, True, True, False), (False, True, True, True), (False, False, True, False),
         (True, False, False, False), (False, True, False, False)]


def flamelet_all(g, c, eig=True):
    ns, nzi = c['ns'], c['nzi']
    Tc, Tr, hc, hr = c['heat']
    arrs = [np.zeros(nzi * ns), np.zeros(nzi * ns), np.zeros(nzi * ns), np.zeros(nzi), np.zeros(nzi)]
    g.flamelet_stencils(c['dz'], nzi, c['chi'], np.ones(ns), *arrs)
    cmaj, csub, csup, mc, nc = arrs
    out = dict(cmaj=cmaj, csub=csub, csup=csup, mc=mc, nc=nc)
    nj = ns * (nzi * ns + 2 * (nzi - 1))
    for k, (ad, ef, vc, sh) in enumerate(FLAGS):
        r = np.zeros(nzi * ns)
        g.flamelet_rhs(c['state'], 101325., c['oxy'], c['fuel'], ad, Tc, Tr, hc, hr, nzi, cmaj, csub, csup, mc, nc,
                       c['chi'], ef, vc, sh, r)
        out[f'rhs{k}'] = r
        for ce, so in ((False, False), (eig, True)):
            J, ee = np.zeros(nj), np.zeros(nzi * ns)
            g.flamelet_jacobian(c['state'], 101325., c['oxy'], c['fuel'], ad, Tc, Tr, hc, hr, nzi, cmaj, csub, csup,
                                mc, nc, c['chi'], ce, 0.3, so, 1.7e-8, 0, 0, ef, vc, sh, ee, J)
            out[f'jac{k}{so}'] = J
            if ce:
                out[f'eig{k}'] = ee
    return out


def block_thomas_all(g, A0, rhs, nzi, ns):
    """g: an oracle kernels object (methods btddod_*) or the product module spitfire_b200.griffon (py_btddod_*)"""
    A, L = A0.copy(), np.zeros(nzi * ns * ns)
    piv, x, mv = np.zeros(nzi * ns, dtype=np.int32), np.zeros(nzi * ns), np.zeros(nzi * ns)
    fact = getattr(g, 'btddod_full_factorize', None) or g.py_btddod_full_factorize
    solve = getattr(g, 'btddod_full_solve', None) or g.py_btddod_full_solve
    matvec = getattr(g, 'btddod_full_matvec', None) or g.py_btddod_full_matvec
    sad = getattr(g, 'btddod_scale_and_add_diagonal', None) or g.py_btddod_scale_and_add_diagonal
    fact(A, nzi, ns, L, piv)
    solve(A, L, piv, rhs, nzi, ns, x)
    matvec(A0, x, nzi, ns, mv)
    B = A0.copy()
    sad(B, -2.0, rhs, 0.5, nzi, ns)
    return dict(A=A, L=L, piv=piv, x=x, mv=mv, B=B)


def assemble_dense(A, nzi, ns):
    n = nzi * ns
    M = np.zeros((n, n))
    for i in range(nzi):
        M[i * ns:(i + 1) * ns, i * ns:(i + 1) * ns] = A[i * ns * ns:(i + 1) * ns * ns].reshape(ns, ns).T
    sub = A[nzi * ns * ns:nzi * ns * ns + (nzi - 1) * ns]
    sup = A[nzi * ns * ns + (nzi - 1) * ns:]
    for k in range((nzi - 1) * ns):
        M[ns + k, k] = sub[k]
        M[k, ns + k] = sup[k]
    return M


def error_stats(a, ref):
    """error measures between a computed array and its reference; 2-D arrays are [state, entries] and every state is
    scaled by its own max|ref| (1-D arrays by the global max):
      strict_*   : |d|/|ref| over entries with ref != 0 (inf where ref == 0 != a)
      big_max    : max |d|/|ref| over the "well-conditioned" entries |ref| >= 1e-3 * scale
      scaled_max : max |d| / (|ref| + 1e-3 * scale) over all entries"""
    a, ref = np.asarray(a), np.asarray(ref)
    d = np.abs(a - ref)
    if a.ndim == 2:
        scale = np.max(np.abs(ref), axis=1, keepdims=True) * np.ones_like(ref)
    else:
        scale = (np.max(np.abs(ref)) if ref.size else 0.) * np.ones_like(ref)
    with np.errstate(all='ignore'):
        strict = np.where(np.abs(ref) > 0, d / np.abs(ref), np.where(d == 0, 0., np.inf))
        scaled = np.where(d == 0, 0., d / (np.abs(ref) + 1e-3 * scale))
    big = (np.abs(ref) >= 1e-3 * scale) & (np.abs(ref) > 0)
    finite = strict[np.isfinite(strict)]
    return dict(strict_max=float(np.max(strict)) if strict.size else 0.,
                strict_p999=float(np.quantile(finite, 0.999)) if finite.size else 0.,
                strict_median=float(np.median(strict)) if strict.size else 0.,
                big_max=float(np.max(strict[big])) if big.any() else 0.,
                scaled_max=float(np.max(scaled)) if scaled.size else 0.,
                nan=int(np.isnan(a).sum()))
