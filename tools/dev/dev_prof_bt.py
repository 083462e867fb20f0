"""dev: one GRI-3.0 128-point flamelet batch through jac_and_eig / factorize_inv / solve_inv, for ncu captures (not a test)"""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import torch
from common import build_mech
from spitfire_b200.flamelet import Flamelet, FlameletSpec, FlameletBatch
F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = build_mech('methane-gri30', 'gpu')
air = m.stream(stp_air=True); fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
f = Flamelet(FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128, initial_condition='linear-TY', stoich_dissipation_rate=1.))
b = FlameletBatch([f] * F); ops = b.ops
state = b._initial(None)
for _ in range(2):
    J, e = ops.jac_and_eig(state, torch.zeros(F, dtype=torch.float64, device='cuda'))
    r = ops.rhs(state)
    fact = ops.factorize(J.clone().neg_(), with_inverse=True)
    x = ops.solve(fact, r)
torch.cuda.synchronize()
print('ok', float(e.max()), float(x.abs().max()))
