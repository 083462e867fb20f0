"""
Chemical equilibrium of an ideal-gas stream at constant (H,P) or (T,P), without Cantera.

The reference obtains its flamelet initial guess from Cantera's `Quantity.equilibrate('HP')` (flamelet.py:543-552);
Cantera is not a dependency here, so this module minimises the Gibbs energy directly with the element-potential
("reduced Gibbs iteration") method of Gordon & McBride (NASA RP-1311, 1994, sections 2.3-3.4): Newton's method on the
element multipliers pi_i, the total mole number and (for HP) the temperature, with their step-size control.
It runs on the host once per grid point of one flamelet -- a cold path, like the Cantera call it stands in for --
using the NASA7 / constant-cp data of the mechanism's `mech_data` dictionary.
"""
import numpy as np


class _HostThermo(object):
    """species cp/R, h/RT, s/R from mech_data (host-side, used only by the equilibrium solver)"""

    def __init__(self, mechanism):
        md = mechanism.mech_data
        self.R = mechanism.gas_constant
        self.p_ref = md['ref_pressure']
        self.names = list(mechanism.species_names)
        self.mw = np.asarray(mechanism.molecular_weights(), dtype=float) if callable(mechanism.molecular_weights) \
            else np.asarray(mechanism.molecular_weights, dtype=float)
        self.data = [md['species'][s]['cp'] for s in self.names]
        for d in self.data:
            if d[0] not in ('NASA7', 'constant'):
                raise NotImplementedError('equilibrium: only NASA7 and constant-cp species are supported')

    def evaluate(self, T):
        ns = len(self.data)
        cp, h, s = np.zeros(ns), np.zeros(ns), np.zeros(ns)
        lnT = np.log(T)
        for j, d in enumerate(self.data):
            if d[0] == 'NASA7':
                a = d[4] if T <= d[2] else d[5]
                cp[j] = a[0] + T * (a[1] + T * (a[2] + T * (a[3] + T * a[4])))
                h[j] = a[0] + T * (a[1] / 2 + T * (a[2] / 3 + T * (a[3] / 4 + T * a[4] / 5))) + a[5] / T
                s[j] = a[0] * lnT + T * (a[1] + T * (a[2] / 2 + T * (a[3] / 3 + T * a[4] / 4))) + a[6]
            else:  # ('constant', Tmin, Tmax, T0, h0, s0, cp) in molar units (J/kmol, J/kmol/K)
                T0, h0, s0, c = d[3], d[4], d[5], d[6]
                cp[j] = c / self.R
                h[j] = (h0 + c * (T - T0)) / (self.R * T)
                s[j] = (s0 + c * (lnT - np.log(T0))) / self.R
        return cp, h, s


def equilibrate(stream, XY='HP', max_iterations=400, tolerance=1.e-10):
    """bring `stream` (spitfire_b200.streams.Stream) to chemical equilibrium in place, holding (H,P) or (T,P)"""
    XY = XY.upper()
    if XY not in ('HP', 'TP'):
        raise ValueError('equilibrate supports "HP" and "TP"')
    mech = stream.mechanism
    th = _HostThermo(mech)
    ns = len(th.names)
    elements = list(mech.element_names)
    A = np.array([[mech.n_atoms(j, e) for j in range(ns)] for e in elements], dtype=float)  # [ne, ns]
    Y0 = np.asarray(stream.Y, dtype=float)
    n0 = Y0 / th.mw  # kmol of species per kg of mixture
    b0 = A @ n0
    keep_e = b0 > 1e-300
    # species that contain an absent element cannot exist
    possible = ~np.any((A[~keep_e] > 0), axis=0) if np.any(~keep_e) else np.ones(ns, dtype=bool)
    A = A[keep_e][:, possible]
    b0 = b0[keep_e]
    ne, nsp = A.shape
    mw = th.mw[possible]
    P = stream.P
    T = float(stream.T)
    h_target = None
    if XY == 'HP':
        cp0, h0, _ = th.evaluate(T)
        h_target = float(np.sum(n0 * h0 * th.R * T))  # J/kg
        T = max(T, 2000.) if T < 2000. else T  # CEA starts hot; the temperature correction is damped below
        T = min(T, 3800.)
    n = np.full(nsp, 0.1 / nsp / np.mean(mw) * 10.)  # rough positive start, kmol/kg
    n *= 1. / (np.sum(n * mw))  # one kilogram
    ntot = float(np.sum(n))
    lnpp = np.log(P / th.p_ref)
    size = ne + 1 + (1 if XY == 'HP' else 0)
    for it in range(max_iterations):
        cp_all, h_all, s_all = th.evaluate(T)
        cp, h, s = cp_all[possible], h_all[possible], s_all[possible]
        with np.errstate(divide='ignore'):
            mu = h - s + np.log(np.maximum(n, 1e-300) / ntot) + lnpp  # mu_j / RT
        G = np.zeros((size, size))
        r = np.zeros(size)
        An = A * n  # [ne, nsp]
        G[:ne, :ne] = An @ A.T
        G[:ne, ne] = An.sum(axis=1)
        G[ne, :ne] = G[:ne, ne]
        G[ne, ne] = np.sum(n) - ntot
        r[:ne] = b0 - An.sum(axis=1) + An @ mu
        r[ne] = ntot - np.sum(n) + np.sum(n * mu)
        if XY == 'HP':
            G[:ne, ne + 1] = An @ h
            G[ne, ne + 1] = np.sum(n * h)
            G[ne + 1, :ne] = G[:ne, ne + 1]
            G[ne + 1, ne] = G[ne, ne + 1]
            G[ne + 1, ne + 1] = np.sum(n * cp) + np.sum(n * h * h)
            r[ne + 1] = (h_target / (th.R * T) - np.sum(n * h)) + np.sum(n * h * mu)
        try:
            x = np.linalg.solve(G, r)
        except np.linalg.LinAlgError:
            x = np.linalg.lstsq(G, r, rcond=None)[0]
        pi, dlnn = x[:ne], x[ne]
        dlnT = x[ne + 1] if XY == 'HP' else 0.
        dlnnj = -mu + A.T @ pi + dlnn + h * dlnT
        # step-size control, RP-1311 eq. 3.1-3.3
        major = n / ntot > 1.e-8
        lam1 = max(5. * abs(dlnT), 5. * abs(dlnn), np.max(np.abs(dlnnj[major])) if np.any(major) else 0.)
        lam1 = 2. / lam1 if lam1 > 2. else 1.
        lam2 = 1.
        minor = (~major) & (dlnnj > 0)
        if np.any(minor):
            with np.errstate(divide='ignore', invalid='ignore'):
                cand = np.abs((-np.log(np.maximum(n[minor], 1e-300) / ntot) - 9.2103404) / (dlnnj[minor] - dlnn))
            cand = cand[np.isfinite(cand)]
            if cand.size:
                lam2 = min(1., float(np.min(cand)))
        lam = min(1., lam1, lam2)
        n = n * np.exp(np.clip(lam * dlnnj, -60., 60.))
        ntot = ntot * np.exp(lam * dlnn)
        T = T * np.exp(lam * dlnT)
        T = min(max(T, 150.), 6000.)
        conv = np.max(n * np.abs(dlnnj)) / np.sum(n) <= tolerance and abs(dlnn) * ntot / np.sum(n) <= tolerance and \
            abs(dlnT) <= tolerance and np.max(np.abs(b0 - (A * n).sum(axis=1))) <= tolerance * np.max(b0)
        if conv and lam == 1.:
            break
    else:
        # Cantera's equilibrate raises when it fails; a stalled iteration must not feed a flamelet initial condition or
        # an equilibrium library silently
        raise RuntimeError(
            f'equilibrate({XY!r}) did not converge in {max_iterations} iterations: T = {T:.2f} K, '
            f'max |n dln n| / sum n = {np.max(n * np.abs(dlnnj)) / np.sum(n):.3e}, |dln n| = {abs(dlnn):.3e}, '
            f'|dln T| = {abs(dlnT):.3e}, element balance error = '
            f'{np.max(np.abs(b0 - (A * n).sum(axis=1))) / np.max(b0):.3e} (tolerance {tolerance:.1e})')
    Y = np.zeros(ns)
    Y[possible] = n * mw
    Y /= np.sum(Y)
    if XY == 'HP':
        stream.TPY = T, P, Y
        stream.HPY = h_target, P, Y  # polish T against the enthalpy with the stream's own thermodynamics
    else:
        stream.TPY = stream.T, P, Y
    return stream
