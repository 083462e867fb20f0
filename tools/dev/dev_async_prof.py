"""dev: cProfile of the asynchronous batch integrator on the 56 heat-loss trajectories of config 5"""
import os, sys, time, cProfile, pstats
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np, torch
from common import build_mech
from spitfire_b200 import tabulation as tab
from spitfire_b200.flamelet import Flamelet, FlameletBatch, FlameletSpec
m = build_mech('methane-gri30', 'gpu')
air = m.stream(stp_air=True); fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
specs = FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128)
table, _, _ = tab.build_adiabatic_slfm_library(specs, np.logspace(-3, 2, 64), verbose=False, _return_intermediates=True, wave=8)
fls = [Flamelet(tab._transient_heat_loss_specs(specs, table, c)) for c in table.keys()]
args = tab._transient_integration_args(None, False)
b = FlameletBatch(fls)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); t0 = time.perf_counter()
b.integrate_for_heat_loss(**args)
torch.cuda.synchronize(); print('wall', time.perf_counter() - t0); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(28)
