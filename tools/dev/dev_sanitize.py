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
