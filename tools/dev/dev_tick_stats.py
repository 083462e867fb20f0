"""dev: host-clock accounting of the asynchronous integrator's ticks on the 56 heat-loss trajectories of config 5
(gb_debug_tick_stats); GB_TICK_PROFILE=1 adds a synchronisation after the solve and after the rhs of every round"""
import os, sys, time, ctypes as C
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np, torch
from common import build_mech
from spitfire_b200 import tabulation as tab, griffon
from spitfire_b200.flamelet import Flamelet, FlameletBatch, FlameletSpec
stride = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m = build_mech('methane-gri30', 'gpu')
air = m.stream(stp_air=True); fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
specs = FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128)
table, _, _ = tab.build_adiabatic_slfm_library(specs, np.logspace(-3, 2, 64)[::stride], verbose=False, _return_intermediates=True, wave=8)
lib = griffon.load_library()
buf = (C.c_double * 8)()
for rep in range(2):
    fls = [Flamelet(tab._transient_heat_loss_specs(specs, table, c)) for c in table.keys()]
    b = FlameletBatch(fls)
    lib.gb_debug_tick_stats(buf, 1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    libs, failed = b.integrate_for_heat_loss(**tab._transient_integration_args(None, False))
    torch.cuda.synchronize(); wall = time.perf_counter() - t0
    lib.gb_debug_tick_stats(buf, 1)
    s = list(buf)
    print(f'members {len(fls)} wall {wall:.3f} s  ticks {int(s[0])} rounds {int(s[1])}  in ticks {s[2]:.3f} s  in round loops {s[3]:.3f} s'
          f'  per round {s[3] / max(s[1], 1) * 1e3:.3f} ms  solve(sync) {s[4]:.3f} s  update+rhs(sync) {s[5]:.3f} s  steps {sum(l.shape[0] for l in libs)}', flush=True)
    from spitfire_b200.time import batched
    print('   ', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in batched.LAST_ASYNC_STATS.items()}, flush=True)
