import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from common import build_mech
from spitfire_b200.flamelet import FlameletSpec
from spitfire_b200 import tabulation as tab
m = build_mech('methane-gri30', 'gpu')
air = m.stream(stp_air=True); fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
t = time.time()
lib = tab.build_adiabatic_slfm_library(FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128), diss_rate_values=np.logspace(-3, 2, 64), verbose=True, wave=1)
print('total', time.time() - t)
