"""dev: how much of the batched heat-loss integration is lock-step overhead? Newton iterations (rounds of kernels) taken
by the 56-member batch against each member integrated on its own (GRI-3.0, config 5)."""
import os, sys, time
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np, torch
from common import build_mech
from spitfire_b200 import tabulation as tab
from spitfire_b200 import flamelet as fl
from spitfire_b200.flamelet import Flamelet, FlameletBatch, FlameletSpec
from spitfire_b200.time import batched as _tb

m = build_mech('methane-gri30', 'gpu')
air = m.stream(stp_air=True); fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
specs = FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128)
chis = np.logspace(-3, 2, 64)
table, _, _ = tab.build_adiabatic_slfm_library(specs, chis, verbose=False, _return_intermediates=True, wave=8)
keys = list(table.keys())
count = dict(stages=0, its=0)
orig = fl._BatchOps.newton_stage
def counted(self, *a, **k):
    left, its = orig(self, *a, **k)
    count['stages'] += 1; count['its'] += its
    return left, its
fl._BatchOps.newton_stage = counted
orig2 = fl._BatchOps.esdirk_stages
def counted2(self, *a, **k):
    left, rounds = orig2(self, *a, **k)
    count['stages'] += 5; count['its'] += rounds
    return left, rounds
fl._BatchOps.esdirk_stages = counted2
def run(sel):
    count.update(stages=0, its=0)
    fls = [Flamelet(tab._transient_heat_loss_specs(specs, table, keys[i])) for i in sel]
    args = tab._transient_integration_args(None, False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    FlameletBatch(fls).integrate_for_heat_loss(**args)
    torch.cuda.synchronize()
    if _tb.LAST_ASYNC_STATS:
        print('   async stats:', dict(_tb.LAST_ASYNC_STATS)); _tb.LAST_ASYNC_STATS.clear()
    return time.perf_counter() - t0, count['stages'], count['its']
print('async members:', _tb.ASYNC_MEMBERS)
print('all %d members: %.2f s, %d stages, %d iterations' % ((len(keys),) + run(list(range(len(keys))))), flush=True)
for i in [0, len(keys) - 1]:
    print('member %2d (chi_st %.3g): %.2f s, %d stages, %d iterations' % ((i, keys[i]) + run([i])), flush=True)
print('members 0..6: %.2f s, %d stages, %d iterations' % run(list(range(7))), flush=True)
