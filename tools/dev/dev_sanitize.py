"""dev: one small call of every kernel, for compute-sanitizer (memcheck / racecheck / synccheck) runs:
    compute-sanitizer --tool racecheck python tools/dev/dev_sanitize.py [mechanism]"""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np, torch
from common import build_mech
from spitfire_b200 import griffon
from spitfire_b200.synthetic import synthetic_states
from spitfire_b200.flamelet import Flamelet, FlameletSpec, FlameletBatch

name = sys.argv[1] if len(sys.argv) > 1 else 'h2-burke'
m = build_mech(name, 'gpu'); g = m.griffon; ns = m.n_species
N = 37
st, _ = synthetic_states(m.species_names, N, 'H2' if ns < 20 else 'CH4')
d_state = torch.from_numpy(st).cuda(); d_rhs = torch.empty((N, ns), dtype=torch.float64, device='cuda')
d_jac = torch.empty((N, ns * ns), dtype=torch.float64, device='cuda')
g.reactor_rhs_isobaric_batch(d_state, 101325., d_rhs)
g.reactor_jac_isobaric_batch(d_state, 101325., d_rhs, d_jac)
torch.cuda.synchronize(); print('reactor ok', float(d_jac.abs().max()))
air = m.stream(stp_air=True); fuel = m.stream('TPX', (300., 101325., 'H2:1' if ns < 20 else 'CH4:1'))
f = Flamelet(FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=12, initial_condition='equilibrium', stoich_dissipation_rate=1.))
b = FlameletBatch([f, f, f]); ops = b.ops
state = b._initial(None)
r = ops.rhs(state); J, e = ops.jac_and_eig(state, torch.tensor([0.1, 0.2, 0.3], dtype=torch.float64, device='cuda'))
torch.cuda.synchronize(); print('flamelet ok', float(e.max()))
Jn = J.clone().neg_()
fact = ops.factorize(Jn.clone()); x1 = ops.solve(fact, r)
fact2 = ops.factorize(Jn.clone(), with_inverse=True); x2 = ops.solve(fact2, r)
x3 = ops.solve(fact2, r[1:], rows=torch.tensor([1, 2], device='cuda'))
mv = torch.zeros_like(r); griffon.py_btddod_full_matvec(Jn, x2, ops.nzi, ops.ns, mv, n_systems=3)
torch.cuda.synchronize(); print('block thomas ok', float((x1 - x2).abs().max()), float((mv - r).abs().max()), float((x3 - x2[1:]).abs().max()))
# ---- round-2 kernels: LU-based inverses against the Gauss-Jordan elimination, integrator vector kernels, non-finite scan,
# the C-level Newton stage loop, isochoric reactor
ops.gauss_jordan_inverses = False
fact_lu = ops.factorize(Jn.clone(), with_inverse=True); x_lu = ops.solve(fact_lu, r)
ops.gauss_jordan_inverses = True
torch.cuda.synchronize(); print('invert ok', float((x_lu - x2).abs().max()))
from spitfire_b200.time import batched
F3, ndof = r.shape
ks = [r.clone() * (1. + 0.1 * j) for j in range(6)]
dt = torch.full((F3,), 1e-6, dtype=torch.float64, device='cuda')
w = 1. / ops.scales
expl, res = torch.empty_like(r), torch.empty_like(r)
conv = torch.zeros(F3, dtype=torch.int32, device='cuda'); cnt = torch.zeros(1, dtype=torch.int32, device='cuda')
x, fq = state.clone(), r.clone()
griffon.esdirk_stage_begin(ks[:3], batched._A[3][:3], batched._G, dt, x, state, fq, expl, res, conv)
dx = ops.solve(fact2, res); xn = torch.empty_like(x)
griffon.newton_update(x, dx, conv, xn, cnt)
fn = ops.rhs(xn)
left = griffon.newton_tail(fn, xn, expl, state, dt, batched._G, w, 1e-12, x, fq, res, conv, cnt)
dq = torch.empty_like(r); stats = torch.empty((3, F3), dtype=torch.float64, device='cuda')
griffon.esdirk_finish(ks, batched._B, batched._BH, dt, w, dq, stats)
acc = torch.ones(F3, dtype=torch.int32, device='cuda'); griffon.accept_step(dq, acc, True, x)
nbad = griffon.count_nonfinite_members(x, dq)
work = torch.empty((3, F3, ndof), dtype=torch.float64, device='cuda')
conv.zero_()
left2, its = ops.newton_stage(fact2, None, ops._all(), None, expl, state, dt, batched._G, w, 1e-12, 3, x, fq, res, conv, work, cnt)
torch.cuda.synchronize(); print('integrator kernels ok', left, nbad, left2, its)
iso = np.concatenate([np.full((N, 1), 0.3), st], axis=1)
d_iso = torch.from_numpy(iso).cuda()
d_r2 = torch.empty((N, ns + 1), dtype=torch.float64, device='cuda'); d_j2 = torch.empty((N, (ns + 1) ** 2), dtype=torch.float64, device='cuda')
g.reactor_jac_isochoric_batch(d_iso, d_r2, d_j2)
torch.cuda.synchronize(); print('isochoric ok', float(d_j2.abs().max()))
