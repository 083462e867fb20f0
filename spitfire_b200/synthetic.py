"""
Synthetic thermochemical state batches for BASELINE.json configs 2-3:
N states, T ~ U(600, 2800) K, composition with probability 1/2 a Dirichlet(0.3) draw over all species, else a
fuel/air mixture at a random mixture fraction blended 50/50 with a Dirichlet draw (so radicals are non-zero).
State layout is the reference's reactor state [T, Y_0..Y_{ns-2}] (isobaric_reactor_kernels.cpp:182-184).
"""
import numpy as np


def synthetic_states(species_names, n, fuel='CH4', seed=20241017, Tmin=600., Tmax=2800., chunk=4096):
    """returns (state [n, ns] float64, y_full [n, ns]); the first k rows do not depend on n (fixed draw chunks)"""
    ns = len(species_names)
    rng = np.random.default_rng(seed)
    idx = {s.upper(): i for i, s in enumerate(species_names)}
    y_fuel = np.zeros(ns)
    if fuel.upper() in idx:
        y_fuel[idx[fuel.upper()]] = 1.
    else:
        y_fuel[0] = 1.
    y_air = np.zeros(ns)
    if 'O2' in idx and 'N2' in idx:
        x_o2, x_n2 = 1., 3.74
        m_o2, m_n2 = 31.998, 28.014
        y_air[idx['O2']] = x_o2 * m_o2 / (x_o2 * m_o2 + x_n2 * m_n2)
        y_air[idx['N2']] = 1. - y_air[idx['O2']]
    else:
        y_air[-1] = 1.
    state = np.empty((n, ns))
    yfull = np.empty((n, ns))
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        T = rng.uniform(Tmin, Tmax, chunk)[:m]
        d = rng.dirichlet(np.full(ns, 0.3), chunk)[:m]
        z = rng.uniform(0., 1., chunk)[:m, None]
        mix = 0.5 * (z * y_fuel[None, :] + (1. - z) * y_air[None, :]) + 0.5 * d
        pick = rng.uniform(size=chunk)[:m] < 0.5
        y = np.where(pick[:, None], d, mix)
        y /= y.sum(axis=1, keepdims=True)
        state[lo:lo + m, 0] = T
        state[lo:lo + m, 1:] = y[:, :-1]
        yfull[lo:lo + m] = y
    return state, yfull


def edge_mixtures(ns):
    """the reference's trace-species mixtures (tests/griffon/test_reaction_rates.py:22-34), generalised to ns species"""
    out = []
    base = np.ones(ns) / ns
    out.append(base.copy())
    for eps in (1.e-8, 1.e-16, 0.0):
        for k in (0, ns // 2, ns - 1):
            y = np.ones(ns)
            y[k] = eps
            out.append(y / y.sum())
    for k in (0, ns - 1):
        y = np.full(ns, 1.e-16)
        y[k] = 1.
        out.append(y / y.sum())
    return np.array(out)
