"""dev: one-sided against twisted block elimination / solve on a GRI-3.0 128-point flamelet Jacobian (not a test)"""
import sys, os, time
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np, torch
from common import build_mech
from spitfire_b200.flamelet import Flamelet, FlameletSpec, FlameletBatch
m = build_mech('methane-gri30', 'gpu')
air = m.stream(stp_air=True); fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
f = Flamelet(FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128, initial_condition='linear-TY', stoich_dissipation_rate=1.))
def tm(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for F in (1, 7, 56):
    b = FlameletBatch([f] * F); ops = b.ops
    state = b._initial(None)
    J = ops.jac(state).neg_(); r = ops.rhs(state)
    out = {}
    for tw in (False, True):
        ops.twisted_elimination = tw
        fact = ops.factorize(J, with_inverse=True)
        out[tw] = ops.solve(fact, r)
        print(f'F={F:3d} twisted={tw!s:5s}: invert {tm(lambda: ops.factorize(J, with_inverse=True)):.3f} ms, solve_inv {tm(lambda: ops.solve(fact, r), 20):.4f} ms', flush=True)
    print(f'F={F:3d} solution twisted vs one-sided: max rel diff {float((out[True] - out[False]).abs().max() / out[False].abs().max()):.2e}', flush=True)
