import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from common import build_mech
from cases import error_stats
from test_gpu_parity import random_states, oracle_batch, gpu_batch
from spitfire_b200.synthetic import synthetic_states
name='heptane-liu'
mg, mo = build_mech(name,'gpu'), build_mech(name,'reference')
ns=mg.n_species; rng=np.random.default_rng(11); n=96; p=101325.
state,y = random_states(ns,n,rng)
rho = np.array([mo.griffon.ideal_gas_density(p, state[i,0], y[i]) for i in range(n)])
ref = oracle_batch(mo.griffon, ns, state, y, p, rho); got = gpu_batch(mg.griffon, ns, state, y, p, rho)
for k in ref:
    bad = np.argwhere((ref[k]==0)&(got[k]!=0))
    print(k, error_stats(got[k],ref[k]), 'zero-ref nonzero-got:', len(bad))
    for b in bad[:8]:
        e=b[1]
        if k=='jac': print('   state',b[0],'row',e%ns,'col',e//ns, mg.species_names[max(e%ns-1,0)], mg.species_names[max(e//ns-1,0)], got[k][tuple(b)], 'T',state[b[0],0])
        elif k=='sens': print('   state',b[0],'row',e%(ns+1),'col',e//(ns+1), got[k][tuple(b)])
        else: print('   state',b[0],'idx',e, got[k][tuple(b)])
print('last species', mg.species_names[-1], 'mw', mg.molecular_weights[-1])
# perf of the new kernel
for name,fuel,N in (('h2-burke','H2',1<<20),('methane-gri30','CH4',1<<18)):
    m = build_mech(name,'gpu'); g=m.griffon; ns=m.n_species
    st,_ = synthetic_states(m.species_names, N, fuel)
    d_state=torch.from_numpy(st).cuda(); d_rhs=torch.empty((N,ns),dtype=torch.float64,device='cuda'); d_jac=torch.empty((N,ns*ns),dtype=torch.float64,device='cuda')
    for fn,tag in ((lambda: g.reactor_jac_isobaric_batch(d_state,101325.,d_rhs,d_jac),'jac'),(lambda: g.reactor_rhs_isobaric_batch(d_state,101325.,d_rhs),'rhs')):
        fn(); torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
        for _ in range(3): fn()
        e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/3
        print(f'{name} {tag}: N={N} {ms:.3f} ms {N/ms*1e3:.3e} states/s')
