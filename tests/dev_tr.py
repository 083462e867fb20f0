"""dev: profile the batched transient expansion (GRI-3.0, 128 points, 8 members)"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
from common import build_mech
from spitfire_b200.flamelet import FlameletSpec
from spitfire_b200 import tabulation as tab
m = build_mech('methane-gri30', 'gpu')
air = m.stream(stp_air=True); fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
fs = FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128)
chis = np.logspace(-1, 1, int(sys.argv[1]) if len(sys.argv) > 1 else 8)
td, z, x = tab.build_adiabatic_slfm_library(fs, diss_rate_values=chis, verbose=False, _return_intermediates=True, wave=8)
local = {}
pr = cProfile.Profile(); pr.enable(); t0 = time.time()
tab._expand_enthalpy_defect_dimension_transient_batch(list(td.keys()), local, fs, td, 1e4, True, None, False)
torch.cuda.synchronize(); print('batch expansion', time.time() - t0); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
