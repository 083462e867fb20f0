"""vector kernels of the batched implicit integrator (csrc/gb_newton.cu) against the tensor expressions they replace --
the numpy expressions of the reference's stage loop (time/methods.py:502-612), Newton solver (time/nonlinear.py:185-268)
and error estimate -- evaluated with eager torch operations in the same order: bit-identical."""
import numpy as np
import pytest

from spitfire_b200.time import batched


def _rand(torch, *shape, seed=0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize('n,ndof', [(1, 53), (7, 374), (56, 6678)])
def test_stage_kernels_equal_the_tensor_expressions(n, ndof):
    import torch
    from spitfire_b200 import griffon as gm
    G = batched._G
    ks = [_rand(torch, n, ndof, seed=10 + j) for j in range(6)]
    q, x = _rand(torch, n, ndof, seed=1), _rand(torch, n, ndof, seed=2)
    dt = torch.rand(n, dtype=torch.float64).cuda() * 1e-3 + 1e-6
    w = torch.rand(n, ndof, dtype=torch.float64).cuda() + 0.5
    for s in range(1, 6):
        explicit = batched._A[s][s - 1] * ks[s - 1]
        for j in range(s - 2, -1, -1):
            explicit = explicit + batched._A[s][j] * ks[j]
        f = ks[s - 1]
        res = dt[:, None] * (G * f + explicit) - (x - q)
        e2, r2 = torch.empty_like(q), torch.empty_like(q)
        conv = torch.ones(n, dtype=torch.int32, device='cuda')
        gm.esdirk_stage_begin(ks[:s], batched._A[s][:s], G, dt, x, q, f, e2, r2, conv)
        assert torch.equal(e2, explicit) and torch.equal(r2, res) and int(conv.sum()) == 0
    # Newton update + tail with a mix of converged and iterating members
    conv = torch.zeros(n, dtype=torch.int32, device='cuda')
    conv[::3] = 1
    dx = _rand(torch, n, ndof, seed=3) * 1e-3
    cnt = torch.full((1,), 99, dtype=torch.int32, device='cuda')
    xn = torch.empty_like(x)
    gm.newton_update(x, dx, conv, xn, cnt)
    keep = conv.bool()[:, None]
    assert torch.equal(xn, torch.where(keep, x, x - dx)) and int(cnt) == 0
    fn = _rand(torch, n, ndof, seed=4)
    rn = dt[:, None] * (G * fn + explicit) - (xn - q)
    norms = (rn * w).abs().amax(dim=1)
    tol = float(norms.median()) if n > 1 else float(norms[0]) * 2.
    x0, f0, r0 = x.clone(), f.clone(), res.clone()
    x1, f1, r1 = x.clone(), f.clone(), res.clone()
    conv1 = conv.clone()
    left = gm.newton_tail(fn, xn, explicit, q, dt, G, w, tol, x1, f1, r1, conv1, cnt)
    assert torch.equal(x1, torch.where(keep, x0, xn)) and torch.equal(f1, torch.where(keep, f0, fn))
    assert torch.equal(r1, torch.where(keep, r0, rn))
    want = conv.bool() | (norms < tol)
    assert torch.equal(conv1.bool(), want) and left == int((~want).sum())
    # a NaN in a member's residual leaves it unconverged
    fn2 = fn.clone()
    fn2[n - 1, ndof // 2] = float('nan')
    conv2 = torch.zeros(n, dtype=torch.int32, device='cuda')
    gm.newton_update(x, dx, conv2, xn, cnt)
    left = gm.newton_tail(fn2, xn, explicit, q, dt, G, w, 1e300, x1, f1, r1, conv2, cnt)
    assert left == 1 and int(conv2[n - 1]) == 0
    # error estimate
    dq = dt[:, None] * (batched._B[0] * ks[0] + batched._B[1] * ks[1] + batched._B[2] * ks[2] + batched._B[3] * ks[3] +
                        batched._B[4] * ks[4] + batched._B[5] * ks[5])
    dqh = dt[:, None] * (batched._BH[0] * ks[0] + batched._BH[1] * ks[1] + batched._BH[2] * ks[2] +
                         batched._BH[3] * ks[3] + batched._BH[4] * ks[4] + batched._BH[5] * ks[5])
    d2, stats = torch.empty_like(q), torch.empty((3, n), dtype=torch.float64, device='cuda')
    gm.esdirk_finish(ks, batched._B, batched._BH, dt, w, d2, stats)
    assert torch.equal(d2, dq)
    assert torch.equal(stats[0], ((dq - dqh) * w).abs().amax(dim=1)) and torch.equal(stats[1], (dq * w).abs().amax(dim=1))
    assert torch.equal(stats[2], torch.ones(n, dtype=torch.float64, device='cuda'))
    ks[2][0, 1] = float('inf')
    gm.esdirk_finish(ks, batched._B, batched._BH, dt, w, d2, stats)
    assert float(stats[2, 0]) == 0. and (n == 1 or float(stats[2, 1]) == 1.)
    # accepted step
    acc = torch.zeros(n, dtype=torch.int32, device='cuda')
    acc[::2] = 1
    qq = q.clone()
    gm.accept_step(dq, acc, True, qq)
    want = torch.where(acc.bool()[:, None], torch.clamp(q + dq, min=0.), q)
    assert torch.equal(qq, want)


@pytest.mark.gpu
def test_fused_stage_loop_reproduces_the_eager_one():
    """the transient heat-loss trajectories of the H2 gold case: the device-resident stage loop and the eager tensor
    loop take the same steps and give the same states, bit for bit"""
    import slfm_cases
    from spitfire_b200 import tabulation as tab
    from spitfire_b200.flamelet import Flamelet, FlameletBatch, FlameletSpec
    from spitfire_b200.time import batched as tb
    specs = FlameletSpec(**slfm_cases.h2_specs('gpu'))
    chis = np.logspace(0, 1, 4)
    table, _, _ = tab.build_adiabatic_slfm_library(specs, chis, verbose=False, _return_intermediates=True)
    out = []
    # eager tensor loop / device kernels driven from Python / one C-ABI call per stage / one call per step with the
    # members walking through the stages independently of each other
    # ... / the members advancing independently of each other across steps (integrate_batch_async)
    for fused, stage, asyn, members in ((False, False, False, False), (True, False, False, False),
                                        (True, True, False, False), (True, True, True, False),
                                        (True, True, False, True)):
        fls = [Flamelet(tab._transient_heat_loss_specs(specs, table, c)) for c in table.keys()]
        args = tab._transient_integration_args({'transient_tolerance': 1e-10}, False)
        saved = tb.FUSED_NEWTON, tb.STAGE_CALL, tb.ASYNC_STAGES, tb.ASYNC_MEMBERS
        tb.FUSED_NEWTON, tb.STAGE_CALL, tb.ASYNC_STAGES, tb.ASYNC_MEMBERS = fused, stage, asyn, members
        try:
            libs, failed = FlameletBatch(fls).integrate_for_heat_loss(**args)
        finally:
            tb.FUSED_NEWTON, tb.STAGE_CALL, tb.ASYNC_STAGES, tb.ASYNC_MEMBERS = saved
        assert not any(failed)
        out.append(libs)
    for other in out[1:]:
        for a, b in zip(out[0], other):
            assert a.shape == b.shape
            assert np.array_equal(a['temperature'], b['temperature'])
            assert np.array_equal(a['mass fraction H2O'], b['mass fraction H2O'])


@pytest.mark.gpu
def test_nonfinite_member_count_and_host_status():
    """'> 0 = number of members with non-finite output' -- as a call of its own on device arrays and as the
    status of the synchronous host entry points"""
    import torch
    from common import build_mech
    from spitfire_b200 import griffon as gm
    a = torch.zeros((9, 53), dtype=torch.float64, device='cuda')
    b = torch.ones((9, 2809), dtype=torch.float64, device='cuda')
    flags = torch.full((9,), 7, dtype=torch.int32, device='cuda')
    assert gm.count_nonfinite_members(a, b, flags) == 0 and int(flags.sum()) == 0
    a[2, 52] = float('nan')
    b[5, 0] = float('inf')
    b[2, 100] = -float('inf')
    assert gm.count_nonfinite_members(a, b, flags) == 2
    assert flags.tolist() == [0, 0, 1, 0, 0, 1, 0, 0, 0]
    assert gm.count_nonfinite_members(a) == 1
    # host entry points: three states, one of them with a NaN temperature
    m = build_mech('h2-burke', 'gpu')
    ns = m.n_species
    st = np.tile(np.concatenate([[1500.], np.full(ns - 1, 1. / ns)]), (3, 1))
    st[1, 0] = np.nan
    rhs, jac = np.zeros((3, ns)), np.zeros((3, ns * ns))
    assert m.griffon.reactor_rhs_isobaric_batch(st, 101325., rhs) == 1
    assert m.griffon.reactor_jac_isobaric_batch(st, 101325., rhs, jac) == 1
    assert np.isnan(rhs[1]).any() and np.isfinite(rhs[[0, 2]]).all() and np.isfinite(jac[[0, 2]]).all()
    st[1, 0] = 1400.
    assert m.griffon.reactor_jac_isobaric_batch(st, 101325., rhs, jac) == 0
