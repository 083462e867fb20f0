"""dev: the non-adiabatic GRI-3.0 library build (config 5) with its stage timers, contexts warm (not a test)"""
import os, sys, time
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np, torch
from common import build_mech
from spitfire_b200 import tabulation as tab
from spitfire_b200.flamelet import FlameletSpec
from spitfire_b200.time import batched
m = build_mech('methane-gri30', 'gpu')
air = m.stream(stp_air=True); fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
specs = lambda: FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128)
chis = np.logspace(-3, 2, 64)
tab.build_adiabatic_slfm_library(specs(), chis[::8], verbose=False, wave=8)  # warm-up
for fast in ([True, False] if len(sys.argv) > 1 else [True]):
    tab.FAST_TRANSIENT_STORE = fast
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lib = tab.build_nonadiabatic_defect_transient_slfm_library(specs(), diss_rate_values=chis, verbose=False, n_defect_st=16, wave=8)
    torch.cuda.synchronize()
    print(f'fast store {fast}: {time.perf_counter() - t0:.3f} s', list(lib.shape),
          {k: round(v, 3) for k, v in tab.LAST_BUILD_TIMES.items()},
          {k: (round(v, 3) if isinstance(v, float) else v) for k, v in batched.LAST_ASYNC_STATS.items()}, flush=True)
