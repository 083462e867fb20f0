"""dev: k_jac4 against the oracle on the larger mechanisms; prints where the differences are"""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np
from common import build_mech, oracle_available
kind = 'reference' if oracle_available('reference') else 'port'
names = sys.argv[1:] or ['methane-gri30', 'heptane-liu']


def report(name, ns, rhs, jac, r_rhs, r_jac):
    for what, a, b in (('rhs', rhs, r_rhs), ('jac', jac, r_jac)):
        sc = np.max(np.abs(b), axis=1, keepdims=True)
        err = np.abs(a - b) / (np.abs(b) + 1e-3 * sc)
        err = np.where(np.isfinite(err), err, 1e300)
        bad = np.argwhere(err > 1e-10)
        print(f'  {what}: max scaled err {err.max():.3e}, bad entries {len(bad)} of {err.size}, nan {np.isnan(a).sum()} ref nan {np.isnan(b).sum()}')
        if what == 'jac' and len(bad):
            rows = sorted(set(int(e % ns) for _, e in bad)); cols = sorted(set(int(e // ns) for _, e in bad))
            print('   bad rows', rows[:60]); print('   bad cols', cols[:60]); print('   bad states', sorted(set(int(s) for s, _ in bad))[:40])
            for s_, e in bad[:12]:
                print(f'    state {s_} row {e % ns} col {e // ns}: got {a[s_, e]:.6e} ref {b[s_, e]:.6e}')
        if what == 'rhs' and len(bad):
            print('   bad', [(int(s_), int(e)) for s_, e in bad[:30]])
            for s_, e in bad[:8]:
                print(f'    state {s_} entry {e}: got {a[s_, e]:.6e} ref {b[s_, e]:.6e}')


for name in names:
    mg, mo = build_mech(name, 'gpu'), build_mech(name, kind)
    ns = mg.n_species
    rng = np.random.default_rng(11)
    n = 96
    for P in (101325., 202650., 1013250.):
        y = rng.dirichlet(np.ones(ns) * 0.5, n); T = rng.uniform(250., 3800., n)
        state = np.ascontiguousarray(np.hstack([T[:, None], y[:, :-1]]))
        rhs, jac = np.zeros((n, ns)), np.zeros((n, ns * ns))
        mg.griffon.reactor_jac_isobaric_batch(state, P, rhs, jac)
        rhs2, jac2 = np.zeros((n, ns)), np.zeros((n, ns * ns))
        mg.griffon.reactor_jac_isobaric_batch(state, P, rhs2, jac2)
        print('repeatable:', np.array_equal(rhs, rhs2, equal_nan=True), np.array_equal(jac, jac2, equal_nan=True))
        r_rhs, r_jac = np.zeros((n, ns)), np.zeros((n, ns * ns))
        mo.griffon.reactor_jac_isobaric_many(state, P, 0, r_rhs, r_jac)
        print(name, 'ns', ns, 'nr', mg.n_reactions, 'p', P)
        report(name, ns, rhs, jac, r_rhs, r_jac)
