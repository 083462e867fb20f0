import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from common import build_mech
from spitfire_b200 import griffon as G
def stats(tag,a,b):
    d=np.abs(a-b); scale=np.max(np.abs(b))
    with np.errstate(all='ignore'):
        strict = np.where(np.abs(b)>0, d/np.abs(b), np.where(d==0,0,np.inf))
    print(f'  {tag:14s} strict rel: max {np.max(strict):.3e} p99.9 {np.quantile(strict,0.999):.3e} median {np.median(strict):.3e} | |d|/(|ref|+1e-3*max) {np.max(d/(np.abs(b)+1e-3*scale)):.3e} nan={np.isnan(a).sum()}')
for name,nz in (('h2-burke',34),('methane-gri30',64)):
    mg = build_mech(name,'gpu'); mo = build_mech(name,'reference')
    g,o = mg.griffon, mo.griffon; ns=mg.n_species
    rng=np.random.default_rng(3)
    z = np.sort(np.hstack([0,rng.uniform(0,1,nz-2),1])); dz = z[1:]-z[:-1]; nzi=nz-2
    chi = 5.0*np.exp(-2*(z-0.5)**2)*(z*(1-z))**0.5+0.01
    def rstate(T):
        y=rng.dirichlet(np.ones(ns)); return np.hstack([T,y[:-1]])
    oxy=rstate(300.); fuel=rstate(350.)
    state=np.hstack([rstate(300+1800*np.sin(np.pi*zz)**2) for zz in z[1:-1]])
    Tc=np.full(nzi,350.);Tr=np.full(nzi,320.);hc=rng.uniform(1,5,nzi)*1e3;hr=rng.uniform(0,1,nzi)
    res={}
    for tag,k in (('o',o),('g',g)):
        cmaj=np.zeros(nzi*ns);csub=np.zeros(nzi*ns);csup=np.zeros(nzi*ns);mc=np.zeros(nzi);nc=np.zeros(nzi)
        k.flamelet_stencils(dz,nzi,chi,np.ones(ns),cmaj,csub,csup,mc,nc)
        rows=np.zeros(ns*(nzi*ns+2*(nzi-1)),dtype=np.int32); cols=np.zeros_like(rows); k.flamelet_jac_indices(nzi,rows,cols)
        out={'cmaj':cmaj,'csub':csub,'csup':csup,'mc':mc,'nc':nc,'rows':rows,'cols':cols}
        for kk,(ad,ef,vc,sh) in enumerate([(True,True,True,False),(False,True,True,True),(False,False,True,False),(True,False,False,False),(False,True,False,False)]):
            r=np.zeros(nzi*ns); k.flamelet_rhs(state,101325.,oxy,fuel,ad,Tc,Tr,hc,hr,nzi,cmaj,csub,csup,mc,nc,chi,ef,vc,sh,r)
            out[f'rhs{kk}']=r
            for so in (False,True):
                J=np.zeros(ns*(nzi*ns+2*(nzi-1))); ee=np.zeros(nzi*ns)
                k.flamelet_jacobian(state,101325.,oxy,fuel,ad,Tc,Tr,hc,hr,nzi,cmaj,csub,csup,mc,nc,chi,False,0.,so,1.7e-5,0,0,ef,vc,sh,ee,J)
                out[f'jac{kk}{so}']=J
        res[tag]=out
    print(name)
    for key in res['o']:
        if key in ('rows','cols'): print('  ',key,np.array_equal(res['o'][key],res['g'][key]))
        elif key.startswith(('rhs','jac')): stats(key,res['g'][key],res['o'][key])
        else: print('  ',key,np.array_equal(res['o'][key],res['g'][key]))
    # block Thomas: A = prefactor*J - I (ESDIRK form, well conditioned) and -J
    for key,label in (('jac0True','gdtJ-I'),('jac0False','-J')):
        A0 = res['o'][key].copy() if label=='gdtJ-I' else -res['o'][key]
        rhs=rng.normal(size=nzi*ns)
        sol={}
        for tag in ('o','g'):
            A=A0.copy(); L=np.zeros(nzi*ns*ns); piv=np.zeros(nzi*ns,dtype=np.int32); x=np.zeros(nzi*ns); mv=np.zeros(nzi*ns)
            if tag=='o':
                o.btddod_full_factorize(A,nzi,ns,L,piv); o.btddod_full_solve(A,L,piv,rhs,nzi,ns,x); o.btddod_full_matvec(A0,x,nzi,ns,mv)
                B=A0.copy(); o.btddod_scale_and_add_diagonal(B,-2.0,rhs,0.5,nzi,ns)
            else:
                t0=time.time(); G.py_btddod_full_factorize(A,nzi,ns,L,piv); t1=time.time(); G.py_btddod_full_solve(A,L,piv,rhs,nzi,ns,x); t2=time.time(); G.py_btddod_full_matvec(A0,x,nzi,ns,mv)
                B=A0.copy(); G.py_btddod_scale_and_add_diagonal(B,-2.0,rhs,0.5,nzi,ns)
                print(f'   gpu host-path factorize {1e3*(t1-t0):.2f} ms solve {1e3*(t2-t1):.2f} ms')
            sol[tag]=dict(A=A,L=L,piv=piv,x=x,mv=mv,B=B)
            print(f'  [{label}] {tag} residual |Ax-b|/|b| = {np.max(np.abs(mv-rhs))/np.max(np.abs(rhs)):.3e}')
        stats(label+' x',sol['g']['x'],sol['o']['x']); stats(label+' LU',sol['g']['A'],sol['o']['A']); stats(label+' L',sol['g']['L'][ns*ns:],sol['o']['L'][ns*ns:])
        print('   pivots equal:', np.array_equal(sol['g']['piv'],sol['o']['piv']), ' scale_add equal:', np.array_equal(sol['g']['B'],sol['o']['B']), ' matvec rel', np.max(np.abs(sol['g']['mv']-sol['o']['mv']))/np.max(np.abs(sol['o']['mv'])))
    # device timing, batch of F flamelets
    F=64
    J1 = res['o']['jac0True']; nj=J1.size
    dA = torch.from_numpy(np.tile(J1,(F,1))).cuda(); dL=torch.zeros((F,nzi*ns*ns),dtype=torch.float64,device='cuda'); dP=torch.zeros((F,nzi*ns),dtype=torch.int32,device='cuda')
    dR = torch.from_numpy(np.tile(rhs,(F,1))).cuda(); dX=torch.zeros_like(dR)
    def timeit(fn,reps=3):
        fn(); torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/reps
    dA0=dA.clone()
    t=timeit(lambda: (dA.copy_(dA0), G.py_btddod_full_factorize(dA,nzi,ns,dL,dP,n_systems=F))); print(f'  device factorize F={F}: {t:.3f} ms (incl copy)')
    dA.copy_(dA0); G.py_btddod_full_factorize(dA,nzi,ns,dL,dP,n_systems=F)
    t=timeit(lambda: G.py_btddod_full_solve(dA,dL,dP,dR,nzi,ns,dX,n_systems=F)); print(f'  device solve F={F}: {t:.3f} ms')
    t1=timeit(lambda: (dA[:1].copy_(dA0[:1]), G.py_btddod_full_factorize(dA[:1],nzi,ns,dL[:1],dP[:1],n_systems=1))); print(f'  device factorize F=1: {t1:.3f} ms')
    dA.copy_(dA0); G.py_btddod_full_factorize(dA,nzi,ns,dL,dP,n_systems=F)
    t1=timeit(lambda: G.py_btddod_full_solve(dA[:1],dL[:1],dP[:1],dR[:1],nzi,ns,dX[:1],n_systems=1)); print(f'  device solve F=1: {t1:.3f} ms')
    # flamelet rhs/jac device timing
    dS = torch.from_numpy(np.tile(state,(F,1))).cuda(); dRhs=torch.zeros_like(dS); dJ=torch.zeros((F,nj),dtype=torch.float64,device='cuda')
    tt = lambda a: torch.from_numpy(a).cuda()
    arrs = [tt(a) for a in (oxy,fuel,Tc,Tr,hc,hr,res['o']['cmaj'],res['o']['csub'],res['o']['csup'],res['o']['mc'],res['o']['nc'],chi)]
    prm = g._flamelet_params(101325.,arrs[0],arrs[1],False,arrs[2],arrs[3],arrs[4],arrs[5],nzi,arrs[6],arrs[7],arrs[8],arrs[9],arrs[10],arrs[11],True,True,True)
    t=timeit(lambda: g.flamelet_rhs_batch(F,dS,prm,dRhs)); print(f'  device flamelet_rhs F={F}: {t:.3f} ms')
    t=timeit(lambda: g.flamelet_jacobian_batch(F,dS,prm,dJ)); print(f'  device flamelet_jac F={F}: {t:.3f} ms')
    print('   batch rhs == single:', np.array_equal(dRhs[5].cpu().numpy(), res['g']['rhs1']), ' jac:', np.array_equal(dJ[7].cpu().numpy(), res['g']['jac1False']))
