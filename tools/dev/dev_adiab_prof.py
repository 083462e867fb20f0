"""dev: where the adiabatic GRI-3.0 library build (config 4, wave 8) spends its time -- cProfile + synchronised op timings"""
import os, sys, time, cProfile, pstats
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np, torch
from common import build_mech
from spitfire_b200 import tabulation as tab
from spitfire_b200 import flamelet as fl
from spitfire_b200.flamelet import FlameletSpec
m = build_mech('methane-gri30', 'gpu')
air = m.stream(stp_air=True); fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
specs = lambda: FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128)
chis = np.logspace(-3, 2, 64)
for k in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tab.build_adiabatic_slfm_library(specs(), chis, verbose=False, wave=8)
    torch.cuda.synchronize(); print('build', k, time.perf_counter() - t0, flush=True)
pr = cProfile.Profile(); pr.enable(); t0 = time.perf_counter()
tab.build_adiabatic_slfm_library(specs(), chis, verbose=False, wave=8)
torch.cuda.synchronize(); print('profiled wall', time.perf_counter() - t0); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
acc = {}
def wrap(name):
    f = getattr(fl._BatchOps, name)
    def g(self, *a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = f(self, *a, **k)
        torch.cuda.synchronize()
        e = acc.setdefault(name, [0, 0.]); e[0] += 1; e[1] += time.perf_counter() - t
        return r
    setattr(fl._BatchOps, name, g)
for nm in ('rhs', 'jac', 'factorize', 'solve', 'jac_and_eig', 'add_to_block_diagonal'):
    wrap(nm)
t0 = time.perf_counter()
tab.build_adiabatic_slfm_library(specs(), chis, verbose=False, wave=8)
torch.cuda.synchronize(); print('op-synchronised wall', time.perf_counter() - t0)
print({k: (v[0], round(v[1], 3)) for k, v in acc.items()})
