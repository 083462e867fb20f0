"""dev: k_jac time against the start offset of the persistent CTAs (GB_JAC_STAGGER, cycles per slot) -- not a test"""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import torch
from common import build_mech
from spitfire_b200.synthetic import synthetic_states
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
m = build_mech('methane-gri30', 'gpu'); g = m.griffon; ns = m.n_species
st, _ = synthetic_states(m.species_names, N, 'CH4')
d_state = torch.from_numpy(st).cuda(); d_rhs = torch.empty((N, ns), dtype=torch.float64, device='cuda')
d_jac = torch.empty((N, ns * ns), dtype=torch.float64, device='cuda')
ref = None
for stg in [0] + [int(x) for x in sys.argv[2:]] + [0]:
    os.environ['GB_JAC_STAGGER'] = str(stg)
    g.reactor_jac_isobaric_batch(d_state, 101325., d_rhs, d_jac); torch.cuda.synchronize()
    if ref is None:
        ref = d_jac.clone()
    same = bool(torch.equal(ref, d_jac))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(5): g.reactor_jac_isobaric_batch(d_state, 101325., d_rhs, d_jac)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 5
    print(f'stagger {stg:6d}: N={N} {ms:.3f} ms {N/ms*1e3:.3e} states/s  {N*8*(ns*ns+2*ns)/ms/1e6:.1f} GB/s  identical={same}', flush=True)
