import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from common import build_mech
from spitfire_b200.synthetic import synthetic_states
print(torch.cuda.get_device_name(0))
for name, fuel in (('h2-burke','H2'),('methane-gri30','CH4')):
    mg = build_mech(name,'gpu'); mo = build_mech(name,'reference'); mp = build_mech(name, 'port')
    g, o = mg.griffon, mo.griffon
    ns = mg.n_species
    n = 512
    state, y = synthetic_states(mg.species_names, n, fuel)
    p = 101325.
    # oracle
    rhs_o = np.zeros((n,ns)); jac_o=np.zeros((n,ns*ns)); rhs2_o=np.zeros((n,ns)); w_o=np.zeros((n,ns)); s_o=np.zeros((n,(ns+1)**2))
    rho = np.zeros(n)
    t0=time.time()
    for i in range(n):
        o.reactor_rhs_isobaric(state[i],p,0.,np.zeros(1),0.,0.,0.,0.,0.,0.,0,False,rhs_o[i])
        o.reactor_jac_isobaric(state[i],p,0.,np.zeros(1),0.,0.,0.,0.,0.,0.,0,False,0,0,rhs2_o[i],jac_o[i])
        rho[i]=o.ideal_gas_density(p,state[i,0],y[i])
        o.production_rates(state[i,0],rho[i],y[i],w_o[i])
        o.prod_rates_primitive_sensitivities(rho[i],state[i,0],y[i],0,s_o[i])
    print(name,'oracle time per state (rhs+jac+w+sens)', (time.time()-t0)/n)
    rhs_g=np.zeros((n,ns)); jac_g=np.zeros((n,ns*ns)); rhs2_g=np.zeros((n,ns)); w_g=np.zeros((n,ns)); s_g=np.zeros((n,(ns+1)**2))
    g.reactor_rhs_isobaric_batch(state,p,rhs_g)
    g.reactor_jac_isobaric_batch(state,p,rhs2_g,jac_g)
    g.production_rates_batch(state[:,0].copy(),rho,y,w_g)
    g.prod_rates_sens_batch(rho,state[:,0].copy(),y,0,s_g)
    def stats(tag,a,b):
        scale = np.max(np.abs(b),axis=1,keepdims=True)
        d=np.abs(a-b)
        strict = np.where(np.abs(b)>0, d/np.abs(b), np.where(d==0,0,np.inf))
        rowrel = d/(np.abs(b)+1e-3*scale)
        print(f'  {tag:10s} strict rel: max {np.max(strict):.3e} p99.9 {np.quantile(strict,0.999):.3e} median {np.median(strict):.3e} | |d|/(|ref|+1e-3*max) max {np.max(rowrel):.3e}  nan={np.isnan(a).sum()}')
    stats('rhs',rhs_g,rhs_o); stats('rhs(jac)',rhs2_g,rhs2_o); stats('jac',jac_g,jac_o); stats('w',w_g,w_o); stats('sens',s_g,s_o)
    # single-state API
    r1=np.zeros(ns); j1=np.zeros(ns*ns)
    g.reactor_jac_isobaric(state[3],p,0.,np.zeros(1),0.,0.,0.,0.,0.,0.,0,False,0,0,r1,j1)
    print('  single-state == batch:', np.array_equal(r1,rhs2_g[3]), np.array_equal(j1,jac_g[3]))
    # open + diathermal
    yin = y[7].copy(); r3_o=np.zeros((16,ns)); j3_o=np.zeros((16,ns*ns)); r4_o=np.zeros((16,ns))
    for i in range(16):
        o.reactor_jac_isobaric(state[i],p,900.,yin,1e-3,400.,500.,10.,0.3,2.5,2,True,2,0,r3_o[i],j3_o[i])
        o.reactor_rhs_isobaric(state[i],p,900.,yin,1e-3,400.,500.,10.,0.3,2.5,2,True,r4_o[i])
    r3_g=np.zeros((16,ns)); j3_g=np.zeros((16,ns*ns)); r4_g=np.zeros((16,ns))
    g.reactor_jac_isobaric_batch(state[:16].copy(),p,r3_g,j3_g,900.,yin,1e-3,400.,500.,10.,0.3,2.5,2,True,2,0)
    g.reactor_rhs_isobaric_batch(state[:16].copy(),p,r4_g,900.,yin,1e-3,400.,500.,10.,0.3,2.5,2,True)
    stats('open rhs',r4_g,r4_o); stats('open jrhs',r3_g,r3_o); stats('open jac',j3_g,j3_o)
    # timing on device
    N = 1<<18 if ns>20 else 1<<20
    st_big, _ = synthetic_states(mg.species_names, N, fuel)
    d_state = torch.from_numpy(st_big).cuda(); d_rhs = torch.empty((N,ns),dtype=torch.float64,device='cuda'); d_jac=torch.empty((N,ns*ns),dtype=torch.float64,device='cuda')
    for fn,tag in ((lambda: g.reactor_jac_isobaric_batch(d_state,p,d_rhs,d_jac),'jac'),(lambda: g.reactor_rhs_isobaric_batch(d_state,p,d_rhs),'rhs')):
        fn(); torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); 
        for _ in range(3): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/3
        bytes_ = N*8*(ns*ns+2*ns) if tag=='jac' else N*8*2*ns
        print(f'  {tag}: N={N} {ms:.3f} ms  {N/ms*1e3:.3e} states/s  {bytes_/ms*1e-6:.1f} GB/s')
    print('  parity of big batch first rows vs small:', np.array_equal(d_jac[:n].cpu().numpy(), jac_g))
