"""Per-warp phase timeline of k_jac4 (debug build).  On the GPU box:
    python -m spitfire_b200.build --timeline
    GRIFFON_B200_LIB=spitfire_b200/libgriffon_b200_tl.so python tools/timeline4.py [mechanism] [n_states]
Clocks of every warp of CTA 0 on its third tile, relative to the first consumer mark."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import torch
from common import build_mech
from spitfire_b200 import griffon
from spitfire_b200.synthetic import synthetic_states

name = sys.argv[1] if len(sys.argv) > 1 else 'methane-gri30'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 4 * 8
m = build_mech(name, 'gpu'); g = m.griffon; ns = m.n_species
st, _ = synthetic_states(m.species_names, N, 'H2' if ns < 20 else 'CH4')
d_state = torch.from_numpy(st).cuda(); d_rhs = torch.empty((N, ns), dtype=torch.float64, device='cuda')
d_jac = torch.empty((N, ns * ns), dtype=torch.float64, device='cuda')
for _ in range(2):
    g.reactor_jac_isobaric_batch(d_state, 101325., d_rhs, d_jac)
torch.cuda.synchronize()
lib = griffon.load_library()
buf = (C.c_longlong * (16 * 32))()
lib.gb_debug_jac4_timeline.argtypes = [C.c_void_p]
assert lib.gb_debug_jac4_timeline(buf) == 0
t = np.array(buf[:], dtype=np.int64).reshape(16, 32)
cons = np.nonzero(t[0])[0]; prod = np.nonzero(t[11])[0]
t0 = t[0, cons].min()
names = ['tile start', 'zero-filled', 'producer data ready', 'react done', 'react barrier', 'gather done', 'gather barrier',
         'consts done', 'consts barrier', 'transform done', 'store barrier']
print(f'consumer warps {len(cons)}, producer warps {len(prod)}; cycles relative to the first consumer mark')
prev = 0
for k in range(11):
    v = t[k, cons] - t0
    print(f'  {names[k]:20s} min {v.min():7d} max {v.max():7d} (+{v.max() - prev:6d})  ' + ' '.join(f'{int(x):6d}' for x in v))
    prev = v.max()
pn = ['loop top', 'buffer free', 'chains/thermo done', 'conc/cp done', 'factors done']
for k in range(5):
    v = t[11 + k, prod] - t0
    print(f'  P {pn[k]:18s} ' + ' '.join(f'{int(x):7d}' for x in v))
