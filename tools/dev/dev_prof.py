import sys, numpy as np, os
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import torch
from common import build_mech
from spitfire_b200.synthetic import synthetic_states
name = sys.argv[1] if len(sys.argv)>1 else 'methane-gri30'
N = int(sys.argv[2]) if len(sys.argv)>2 else 65536
which = sys.argv[3] if len(sys.argv)>3 else 'jac'
m = build_mech(name,'gpu'); g=m.griffon; ns=m.n_species
st,_ = synthetic_states(m.species_names, N, 'H2' if ns<20 else 'CH4')
d_state=torch.from_numpy(st).cuda(); d_rhs=torch.empty((N,ns),dtype=torch.float64,device='cuda'); d_jac=torch.empty((N,ns*ns),dtype=torch.float64,device='cuda')
for _ in range(3):
    if which=='jac': g.reactor_jac_isobaric_batch(d_state,101325.,d_rhs,d_jac)
    else: g.reactor_rhs_isobaric_batch(d_state,101325.,d_rhs)
torch.cuda.synchronize()
