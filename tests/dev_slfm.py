"""dev: time GRI-3.0 SLFM building blocks on the GPU (not a test)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
from common import build_mech
from spitfire_b200.flamelet import Flamelet, FlameletSpec, FlameletBatch
from spitfire_b200 import tabulation as tab
backend = sys.argv[1] if len(sys.argv) > 1 else 'gpu'
nchi = int(sys.argv[2]) if len(sys.argv) > 2 else 16
m = build_mech('methane-gri30', backend)
air = m.stream(stp_air=True); fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
fs = FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128)
chis = np.logspace(-1, 1.3, nchi)
for wave in (1, 8):
    t = time.time()
    lib = tab.build_adiabatic_slfm_library(FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128), diss_rate_values=chis, verbose=(wave == 1), wave=wave)
    print(f'adiabatic GRI-128 {nchi} chi wave={wave}: {time.time()-t:.2f} s, shape {lib.shape}, Tmax {lib["temperature"].max():.1f}', flush=True)
# kernel-level timings for one flamelet and a batch
f = Flamelet(FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128, initial_condition='linear-TY', stoich_dissipation_rate=1.))
for F in (1, 8, 64):
    b = FlameletBatch([f] * F); ops = b.ops
    st = torch.as_tensor(np.array([lib['temperature'][:, 0]] * F))  # dummy
    state = b._initial(None)
    def tm(fn, n=5):
        fn(); torch.cuda.synchronize(); t0 = time.time()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.time() - t0) / n * 1e3
    J = ops.jac(state).neg_()
    fact = ops.factorize(J.clone())
    r = ops.rhs(state)
    print(f'F={F}: rhs {tm(lambda: ops.rhs(state)):.3f} ms, jac {tm(lambda: ops.jac(state)):.3f} ms, factorize {tm(lambda: ops.factorize(J.clone())):.3f} ms, solve {tm(lambda: ops.solve(fact, r)):.3f} ms', flush=True)
