"""dev: the asynchronous batch integrator against the lock-step one on GRI-3.0 heat-loss trajectories, bit for bit"""
import os, sys, time
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np, torch
from common import build_mech
from spitfire_b200 import tabulation as tab
from spitfire_b200.flamelet import Flamelet, FlameletBatch, FlameletSpec
from spitfire_b200.time import batched as tb
stride = int(sys.argv[1]) if len(sys.argv) > 1 else 4
m = build_mech('methane-gri30', 'gpu')
air = m.stream(stp_air=True); fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
specs = FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128)
table, _, _ = tab.build_adiabatic_slfm_library(specs, np.logspace(-3, 2, 64)[::stride], verbose=False, _return_intermediates=True)
out = []
for mode in (False, True, True):
    tb.ASYNC_MEMBERS = mode
    fls = [Flamelet(tab._transient_heat_loss_specs(specs, table, c)) for c in table.keys()]
    t0 = time.perf_counter()
    libs, failed = FlameletBatch(fls).integrate_for_heat_loss(**tab._transient_integration_args(None, False))
    print('async' if mode else 'lock-step', '%.2f s' % (time.perf_counter() - t0), 'steps', [l.shape[0] for l in libs], flush=True)
    out.append(libs)
for k in (1, 2):
    same = all(a.shape == b.shape and np.array_equal(a['temperature'], b['temperature']) and
               np.array_equal(a['mass fraction OH'], b['mass fraction OH']) for a, b in zip(out[0], out[k]))
    print('run', k, 'identical to lock-step:', same)
    if not same:
        for i, (a, b) in enumerate(zip(out[0], out[k])):
            if a.shape != b.shape:
                print('  member', i, 'shapes', a.shape, b.shape)
            else:
                print('  member', i, 'max |dT|', float(np.max(np.abs(a['temperature'] - b['temperature']))))
