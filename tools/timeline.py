"""Per-warp phase timeline of k_jac (debug build).  On the GPU box:
    python -m spitfire_b200.build --timeline
    GRIFFON_B200_LIB=spitfire_b200/libgriffon_b200_tl.so python tools/timeline.py [mechanism] [n_states]
Prints, for every warp of CTA 0 on its second tile, the clock at which it ARRIVED at each barrier (relative to the
earliest arrival at the tile's first barrier) and the barrier release times (= max over warps)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import torch
from common import build_mech
from spitfire_b200 import griffon
from spitfire_b200.synthetic import synthetic_states

name = sys.argv[1] if len(sys.argv) > 1 else 'methane-gri30'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 37888
m = build_mech(name, 'gpu'); g = m.griffon; ns = m.n_species
st, _ = synthetic_states(m.species_names, N, 'H2' if ns < 20 else 'CH4')
d_state = torch.from_numpy(st).cuda(); d_rhs = torch.empty((N, ns), dtype=torch.float64, device='cuda')
d_jac = torch.empty((N, ns * ns), dtype=torch.float64, device='cuda')
for _ in range(2):
    g.reactor_jac_isobaric_batch(d_state, 101325., d_rhs, d_jac)
torch.cuda.synchronize()
lib = griffon.load_library()
buf = (C.c_longlong * (20 * 32))()
lib.gb_debug_jac_timeline.argtypes = [C.c_void_p]
assert lib.gb_debug_jac_timeline(buf) == 0
t = np.array(buf[:], dtype=np.int64).reshape(20, 32)
nw = int((t[0] != 0).sum())
tk = t[11:17]
tw = t[17:19]
t = t[:11, :nw]
t0 = t[0].min()
names = ['top', 'load', 'thermo', 'conc', 'react', 'gather', 'write', 'fix', 'rows/cols', 'T-row', 'output']
print('barrier release (max over warps) and phase durations, cycles:')
rel = (t - t0).max(axis=1)
for k in range(11):
    print(f'  {names[k]:10s} release {rel[k]:8d}  phase {rel[k] - (rel[k-1] if k else 0):8d}   warp arrivals (rel. to previous release): '
          + ' '.join(f'{int(x - t0 - (rel[k-1] if k else 0)):6d}' for x in t[k]))


print('reaction phase, mean cycles per group by kind:')
for k, name in enumerate(['fast A+B<=>C+D', 'structured simple', 'third body', 'Lindemann', 'Troe', 'generic']):
    if tk[k, 1]:
        print(f'  {name:18s} {int(tk[k,1])//2:4d} groups  {tk[k,0]/tk[k,1]:8.0f} cycles')
print('reaction phase, cycles inside groups per warp:', ' '.join(str(int(x)//2) for x in tw[0, :nw]))
print('groups per warp:', ' '.join(str(int(x)//2) for x in tw[1, :nw]))
