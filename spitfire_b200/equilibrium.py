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

    def _tables(self):
        if getattr(self, '_tab', None) is None:
            ns = len(self.data)
            nasa = np.array([d[0] == 'NASA7' for d in self.data])
            tmid = np.array([d[2] if d[0] == 'NASA7' else 0. for d in self.data], dtype=float)
            lo = np.array([d[4] if d[0] == 'NASA7' else [0.] * 7 for d in self.data], dtype=float)
            hi = np.array([d[5] if d[0] == 'NASA7' else [0.] * 7 for d in self.data], dtype=float)
            const = np.array([[d[3], d[4], d[5], d[6]] if d[0] != 'NASA7' else [1., 0., 0., 0.] for d in self.data],
                             dtype=float)
            self._tab = (ns, nasa, tmid, lo, hi, const)
        return self._tab

    def evaluate(self, T):
        """(cp/R, h/RT, s/R) of every species at T: the species loop as array expressions, same operations in the same
        order per species"""
        ns, nasa, tmid, lo, hi, const = self._tables()
        lnT = np.log(T)
        a = np.where((T <= tmid)[:, None], lo, hi)
        cp = a[:, 0] + T * (a[:, 1] + T * (a[:, 2] + T * (a[:, 3] + T * a[:, 4])))
        h = a[:, 0] + T * (a[:, 1] / 2 + T * (a[:, 2] / 3 + T * (a[:, 3] / 4 + T * a[:, 4] / 5))) + a[:, 5] / T
        s = a[:, 0] * lnT + T * (a[:, 1] + T * (a[:, 2] / 2 + T * (a[:, 3] / 3 + T * a[:, 4] / 4))) + a[:, 6]
        if not np.all(nasa):  # ('constant', Tmin, Tmax, T0, h0, s0, cp) in molar units (J/kmol, J/kmol/K)
            T0, h0, s0, c = const[:, 0], const[:, 1], const[:, 2], const[:, 3]
            cp = np.where(nasa, cp, c / self.R)
            h = np.where(nasa, h, (h0 + c * (T - T0)) / (self.R * T))
            s = np.where(nasa, s, (s0 + c * (lnT - np.log(T0))) / self.R)
        return cp, h, s


def equilibrate(stream, XY='HP', max_iterations=400, tolerance=1.e-10):
    """bring `stream` (spitfire_b200.streams.Stream) to chemical equilibrium in place, holding (H,P) or (T,P)"""
    XY = XY.upper()
    if XY not in ('HP', 'TP'):
        raise ValueError('equilibrate supports "HP" and "TP"')
    mech = stream.mechanism
    th = getattr(mech, '_equilibrium_thermo', None)
    if th is None:
        th = _HostThermo(mech)
        try:
            mech._equilibrium_thermo = th
        except Exception:
            pass
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


def _evaluate_many(th, T):
    """(cp/R, h/RT, s/R) [P, ns] of every species at the temperatures T [P] (the array form of _HostThermo.evaluate)"""
    ns, nasa, tmid, lo, hi, const = th._tables()
    T = np.asarray(T, dtype=float)[:, None]
    lnT = np.log(T)
    a = np.where((T <= tmid[None, :])[:, :, None], lo[None], hi[None])
    cp = a[..., 0] + T * (a[..., 1] + T * (a[..., 2] + T * (a[..., 3] + T * a[..., 4])))
    h = a[..., 0] + T * (a[..., 1] / 2 + T * (a[..., 2] / 3 + T * (a[..., 3] / 4 + T * a[..., 4] / 5))) + a[..., 5] / T
    s = a[..., 0] * lnT + T * (a[..., 1] + T * (a[..., 2] / 2 + T * (a[..., 3] / 3 + T * a[..., 4] / 4))) + a[..., 6]
    if not np.all(nasa):
        T0, h0, s0, c = const[:, 0], const[:, 1], const[:, 2], const[:, 3]
        cp = np.where(nasa, cp, c / th.R)
        h = np.where(nasa, h, (h0 + c * (T - T0)) / (th.R * T))
        s = np.where(nasa, s, (s0 + c * (lnT - np.log(T0))) / th.R)
    return cp, h, s


def equilibrate_many(streams, XY='HP', max_iterations=400, tolerance=1.e-10):
    """`equilibrate` for a list of streams of one mechanism and one pressure at once (the grid points of a flamelet's
    'equilibrium' initial condition, flamelet.py:543-552): the same reduced Gibbs iteration with the same step-size
    control, every stream with its own iterates and its own convergence test, written as array expressions over the
    streams -- a converged stream is frozen. Streams whose set of possible species differs (a pure oxidiser without
    carbon, say) are grouped and solved group by group. The results agree with the one-at-a-time function to rounding
    (the small dense products run through batched BLAS calls); the streams are updated in place."""
    XY = XY.upper()
    if XY not in ('HP', 'TP'):
        raise ValueError('equilibrate supports "HP" and "TP"')
    if not streams:
        return streams
    mech = streams[0].mechanism
    th = getattr(mech, '_equilibrium_thermo', None)
    if th is None:
        th = _HostThermo(mech)
        try:
            mech._equilibrium_thermo = th
        except Exception:
            pass
    ns = len(th.names)
    elements = list(mech.element_names)
    A_full = np.array([[mech.n_atoms(j, e) for j in range(ns)] for e in elements], dtype=float)
    groups = dict()
    for k, st in enumerate(streams):
        n0 = np.asarray(st.Y, dtype=float) / th.mw
        keep_e = (A_full @ n0) > 1e-300
        possible = ~np.any((A_full[~keep_e] > 0), axis=0) if np.any(~keep_e) else np.ones(ns, dtype=bool)
        groups.setdefault((keep_e.tobytes(), possible.tobytes(), float(st.P)), []).append(k)
    for (ke, po, P), members in groups.items():
        keep_e, possible = np.frombuffer(ke, dtype=bool), np.frombuffer(po, dtype=bool)
        A = A_full[keep_e][:, possible]
        ne, nsp = A.shape
        mw = th.mw[possible]
        Pn = len(members)
        Y0 = np.array([np.asarray(streams[k].Y, dtype=float) for k in members])
        n0 = Y0 / th.mw
        b0 = (n0 @ A_full.T)[:, keep_e]
        T = np.array([float(streams[k].T) for k in members])
        h_target = None
        if XY == 'HP':
            _, h0, _ = _evaluate_many(th, T)
            h_target = np.sum(n0 * h0 * th.R * T[:, None], axis=1)
            T = np.minimum(np.where(T < 2000., np.maximum(T, 2000.), T), 3800.)
        n1 = np.full(nsp, 0.1 / nsp / np.mean(mw) * 10.)
        n1 *= 1. / (np.sum(n1 * mw))
        n = np.tile(n1, (Pn, 1))
        ntot = np.sum(n, axis=1)
        lnpp = np.log(P / th.p_ref)
        hp = XY == 'HP'
        size = ne + 1 + (1 if hp else 0)
        active = np.ones(Pn, dtype=bool)
        last = dict(dn=np.zeros(Pn), dlnn=np.zeros(Pn), dlnT=np.zeros(Pn), eb=np.zeros(Pn))
        for it in range(max_iterations):
            ia = np.nonzero(active)[0]
            if ia.size == 0:
                break
            na, nta, Ta, ba = n[ia], ntot[ia], T[ia], b0[ia]
            cp_all, h_all, s_all = _evaluate_many(th, Ta)
            cp, h, s_ = cp_all[:, possible], h_all[:, possible], s_all[:, possible]
            with np.errstate(divide='ignore'):
                mu = h - s_ + np.log(np.maximum(na, 1e-300) / nta[:, None]) + lnpp
            An = A[None, :, :] * na[:, None, :]
            G = np.zeros((ia.size, size, size))
            r = np.zeros((ia.size, size))
            G[:, :ne, :ne] = An @ A.T
            G[:, :ne, ne] = An.sum(axis=2)
            G[:, ne, :ne] = G[:, :ne, ne]
            G[:, ne, ne] = np.sum(na, axis=1) - nta
            r[:, :ne] = ba - An.sum(axis=2) + np.einsum('pes,ps->pe', An, mu)
            r[:, ne] = nta - np.sum(na, axis=1) + np.sum(na * mu, axis=1)
            if hp:
                G[:, :ne, ne + 1] = np.einsum('pes,ps->pe', An, h)
                G[:, ne, ne + 1] = np.sum(na * h, axis=1)
                G[:, ne + 1, :ne] = G[:, :ne, ne + 1]
                G[:, ne + 1, ne] = G[:, ne, ne + 1]
                G[:, ne + 1, ne + 1] = np.sum(na * cp, axis=1) + np.sum(na * h * h, axis=1)
                r[:, ne + 1] = (h_target[ia] / (th.R * Ta) - np.sum(na * h, axis=1)) + np.sum(na * h * mu, axis=1)
            try:
                x = np.linalg.solve(G, r[:, :, None])[:, :, 0]
            except np.linalg.LinAlgError:
                x = np.array([np.linalg.lstsq(G[q], r[q], rcond=None)[0] for q in range(ia.size)])
            pi, dlnn = x[:, :ne], x[:, ne]
            dlnT = x[:, ne + 1] if hp else np.zeros(ia.size)
            dlnnj = -mu + pi @ A + dlnn[:, None] + h * dlnT[:, None]
            # step-size control, RP-1311 eq. 3.1-3.3
            major = na / nta[:, None] > 1.e-8
            big = np.max(np.where(major, np.abs(dlnnj), 0.), axis=1)
            lam1 = np.maximum(np.maximum(5. * np.abs(dlnT), 5. * np.abs(dlnn)), big)
            lam1 = np.where(lam1 > 2., 2. / np.where(lam1 > 2., lam1, 1.), 1.)
            minor = (~major) & (dlnnj > 0)
            with np.errstate(divide='ignore', invalid='ignore'):
                cand = np.abs((-np.log(np.maximum(na, 1e-300) / nta[:, None]) - 9.2103404) / (dlnnj - dlnn[:, None]))
            cand = np.where(minor & np.isfinite(cand), cand, np.inf)
            lam2 = np.minimum(1., np.min(cand, axis=1))
            lam = np.minimum(1., np.minimum(lam1, lam2))
            na = na * np.exp(np.clip(lam[:, None] * dlnnj, -60., 60.))
            nta = nta * np.exp(lam * dlnn)
            Ta = np.minimum(np.maximum(Ta * np.exp(lam * dlnT), 150.), 6000.)
            sn = np.sum(na, axis=1)
            dn = np.max(na * np.abs(dlnnj), axis=1) / sn
            eb = np.max(np.abs(ba - np.einsum('es,ps->pe', A, na)), axis=1)
            conv = (dn <= tolerance) & (np.abs(dlnn) * nta / sn <= tolerance) & (np.abs(dlnT) <= tolerance) & \
                (eb <= tolerance * np.max(ba, axis=1))
            n[ia], ntot[ia], T[ia] = na, nta, Ta
            last['dn'][ia], last['dlnn'][ia], last['dlnT'][ia], last['eb'][ia] = dn, np.abs(dlnn), np.abs(dlnT), eb
            active[ia[conv & (lam == 1.)]] = False
        if np.any(active):
            k = int(np.nonzero(active)[0][0])
            raise RuntimeError(
                f'equilibrate_many({XY!r}): {int(active.sum())} of {Pn} streams did not converge in {max_iterations} '
                f'iterations; first: T = {T[k]:.2f} K, max |n dln n| / sum n = {last["dn"][k]:.3e}, |dln n| = '
                f'{last["dlnn"][k]:.3e}, |dln T| = {last["dlnT"][k]:.3e}, element balance error = {last["eb"][k]:.3e} '
                f'(tolerance {tolerance:.1e})')
        for q, k in enumerate(members):
            Y = np.zeros(ns)
            Y[possible] = n[q] * mw
            Y /= np.sum(Y)
            st = streams[k]
            if hp:
                st.TPY = T[q], P, Y
                st.HPY = h_target[q], P, Y  # polish T against the enthalpy with the stream's own thermodynamics
            else:
                st.TPY = st.T, P, Y
    return streams
