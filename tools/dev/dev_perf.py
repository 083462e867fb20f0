"""dev: time k_jac / k_rates on the synthetic batches (not a test)"""
import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import torch
from common import build_mech
from spitfire_b200.synthetic import synthetic_states
which = sys.argv[1] if len(sys.argv) > 1 else 'both'
for name, fuel, N in (('h2-burke', 'H2', 1 << 20), ('methane-gri30', 'CH4', 1 << 18)):
    m = build_mech(name, 'gpu'); g = m.griffon; ns = m.n_species
    st, _ = synthetic_states(m.species_names, N, fuel)
    d_state = torch.from_numpy(st).cuda(); d_rhs = torch.empty((N, ns), dtype=torch.float64, device='cuda')
    d_jac = torch.empty((N, ns * ns), dtype=torch.float64, device='cuda')
    for fn, tag in ((lambda: g.reactor_jac_isobaric_batch(d_state, 101325., d_rhs, d_jac), 'jac'),
                    (lambda: g.reactor_rhs_isobaric_batch(d_state, 101325., d_rhs), 'rhs')):
        if which not in ('both', tag):
            continue
        fn(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        for _ in range(3): fn()
        e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 3
        print(f'{name} {tag}: N={N} {ms:.3f} ms {N/ms*1e3:.3e} states/s  {N*8*(ns*ns+2*ns)/ms/1e6:.1f} GB/s', flush=True)
