"""dev: time / profile the block-Thomas kernels on a GRI-3.0 128-point flamelet Jacobian (not a test)"""
import sys, os, time
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np, torch
from common import build_mech
from spitfire_b200.flamelet import Flamelet, FlameletSpec, FlameletBatch
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m = build_mech('methane-gri30', 'gpu')
air = m.stream(stp_air=True); fuel = m.stream('TPX', (300., 101325., 'CH4:1'))
f = Flamelet(FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128, initial_condition='linear-TY', stoich_dissipation_rate=1.))
b = FlameletBatch([f] * F); ops = b.ops
state = b._initial(None)
J = ops.jac(state).neg_(); r = ops.rhs(state)
def tm(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.time() - t0) / n * 1e3
fact = ops.factorize(J.clone())
print(f'F={F}: factorize {tm(lambda: ops.factorize(J.clone())):.3f} ms, solve {tm(lambda: ops.solve(fact, r)):.3f} ms')
ops.gauss_jordan_inverses = False
fact_inv = ops.factorize(J.clone(), with_inverse=True)
print(f'F={F}: factorize_inv {tm(lambda: ops.factorize(J.clone(), with_inverse=True)):.3f} ms, solve_inv {tm(lambda: ops.solve(fact_inv, r)):.3f} ms, rhs {tm(lambda: ops.rhs(state)):.3f} ms, jac {tm(lambda: ops.jac(state)):.3f} ms')
ops.gauss_jordan_inverses = True
fact_gj = ops.factorize(J.clone(), with_inverse=True)
xa, xb = ops.solve(fact_inv, r), ops.solve(fact_gj, r)
print(f'F={F}: invert (Gauss-Jordan) {tm(lambda: ops.factorize(J, with_inverse=True)):.3f} ms; solution vs LU-based inverses: max rel diff {float((xa - xb).abs().max() / xa.abs().max()):.2e}')
x = state.clone(); e = r.clone(); dta = torch.ones(F, device='cuda', dtype=torch.float64)
def elem():
    rn = dta[:, None] * (0.25 * r + e) - (x - state)
    c = ((rn * ops.scales).abs().amax(dim=1) < 1e-12).cpu().numpy()
print(f'elementwise+sync {tm(elem, 20):.3f} ms')
if os.environ.get('GRIFFON_B200_LIB', '').endswith('_tl.so'):
    import ctypes as C
    from spitfire_b200 import griffon
    lib = griffon.load_library()
    buf = (C.c_longlong * 16)()
    lib.gb_debug_bt_timeline(buf)
    before = np.array(buf[:], dtype=np.int64)
    ops.factorize(J.clone(), with_inverse=True); torch.cuda.synchronize()
    lib.gb_debug_bt_timeline(buf)
    d = np.array(buf[:], dtype=np.int64) - before
    names = ['loop top (prev tail)', 'LU', 'store D/piv', 'inverse', 'store inv + wait + barrier', 'L/D update + pivot']
    for k in range(6):
        print(f'  {names[k]:28s} {d[k] / 126:10.0f} cycles per block')
    for k, nm in enumerate(['LU step: read piv + swap', 'update', 'pivot search', 'barrier wait']):
        print(f'  {nm:28s} {d[8 + k] / 126 / 52:10.0f} cycles per step (pivot column group)')
    lib.gb_debug_bt_timeline(buf)
    before = np.array(buf[:], dtype=np.int64)
    ops.solve(fact_inv, r); torch.cuda.synchronize()
    lib.gb_debug_bt_timeline(buf)
    d = np.array(buf[:], dtype=np.int64) - before
    for k, nm in enumerate(['solve_inv: wait_group+mbar', 'barrier', 'fetch issue', 'dot + epilogue']):
        print(f'  {nm:28s} {d[12 + k] / 251:10.0f} cycles per step (thread 0)')
    print(f'  of which dot + shuffles       {d[11] / 251:10.0f}')
