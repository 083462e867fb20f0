"""
ESDIRK64 + PI step control + lagged-Jacobian Newton for a BATCH of independent flamelets, device resident.

This is `odesolve` (integrator.py:229-694) + `KennedyCarpenterS6P4Q3.single_step` (methods.py:502-612) +
`SimpleNewtonSolver` (nonlinear.py:185-268) + `PIController` (stepcontrol.py:84-101) + the Jacobian-refresh policy
(integrator.py:100-130) restated so that every member of the batch follows, with its own time, step size, error
history, refresh flag and Newton iteration count, exactly the sequence of operations the serial code performs for it:
members are independent, so masking reproduces the serial arithmetic member by member. What is shared is the cost: one
right-hand-side launch, one block-Thomas solve launch etc. serve all active members at once, which is where the GPU
wins over the reference's one-flamelet-at-a-time loop (the kernels are latency-bound for a single flamelet).

Only what the flamelet drivers use is covered: implicit ESDIRK64, block-Thomas projector set up "in the governor",
`save_each_step=True`, a `stop(t, q, residual, nsteps) -> bool tensor` criterion and a minimum step count.
"""
import numpy as np

_G = 0.25
_A = [[0., 0., 0., 0., 0., 0.],
      [0.25, _G, 0., 0., 0., 0.],
      [8611. / 62500., -1743. / 31250., _G, 0., 0., 0.],
      [5012029. / 34652500., -654441. / 2922500., 174375. / 388108., _G, 0., 0.],
      [15267082809. / 155376265600., -71443401. / 120774400., 730878875. / 902184768., 2285395. / 8070912., _G, 0.],
      [82889. / 524892., 0., 15625. / 83664., 69875. / 102672., -2260. / 8211., _G]]
_C = [float(np.sum(np.array(row))) for row in _A]
_B = list(_A[5])
_BH = [4586570599. / 29645900160., 0., 178811875. / 945068544., 814220225. / 1159782912., -3700637. / 11593932.,
       61727. / 225920.]


# the device-resident stage loop (csrc/gb_newton.cu) is used for device batches of up to dense_limit members; False falls
# back to the eager tensor loop (kept for larger batches, where the iteration runs on the unconverged members only)
FUSED_NEWTON = True
# the Newton loop of a stage as ONE C-ABI call (gb_flamelet_newton_stage_batch) where the batch's operations offer it
STAGE_CALL = True
# ... and all implicit stages of a step as one call in which every member walks through the stages on its own
# (gb_flamelet_esdirk_stages_batch): the rounds of kernels of a step are then the largest per-member sum of Newton
# iterations instead of the sum over the stages of the largest per-member count
# Measured on config 5 (56 trajectories): 15,792 rounds against 15,815 -- the member that is slowest in a step is slowest
# at every stage of it, so nothing is gained within a step (the slowest member changes from step to step: the longest
# trajectory on its own takes 6,447). Off by default; bit-identical to the stage-by-stage loop (tests/test_gpu_newton.py).
ASYNC_STAGES = False
import os as _os
import time as _time
if _os.environ.get('GB_NEWTON_MODE') in ('eager', 'fused', 'stage', 'async'):  # (A/B switch for tools/bench_slfm.py)
    FUSED_NEWTON = _os.environ['GB_NEWTON_MODE'] != 'eager'
    STAGE_CALL = _os.environ['GB_NEWTON_MODE'] in ('stage', 'async')
    ASYNC_STAGES = _os.environ['GB_NEWTON_MODE'] == 'async'


def integrate_batch(ops, q0, stop, first_time_step=1.e-6, max_time_step=1.e-3, minimum_time_step_count=40,
                    transient_tolerance=1.e-10, maximum_steps_per_jacobian=10, nonlinear_solve_tolerance=1.e-12,
                    max_nonlinear_iter=20, max_ramp=1.1, ki=0.1333333333, maximum_steps=100000,
                    fail_factor=0.8, slow_factor=0.8, grow_limit=1.05, shrink_limit=0.9, clip_negative=True,
                    explicit_inverse_solves=True, save_each_step=True, stop_ignores_minimum=False, dense_limit=148,
                    fused_newton=None, stop_at_time=None):
    """advance every member of `ops` (flamelet._BatchOps) from q0 [F, ndof] until `stop(t, q, residual, nsteps)` (all
    [F]-shaped tensors; returns a bool tensor) holds for it and it has taken at least minimum_time_step_count steps.
    Returns per member the lists of saved times and states (numpy), initial state included, and a `failed` flag
    (non-finite update that the step-size reduction could not cure within maximum_steps). With save_each_step=False
    only the initial and the final time / state of every member are returned (large batches of 0-D reactors).
    stop_at_time: every member stops at this time exactly -- a step that would cross it is shortened, as odesolve does
    (integrator.py:590-593, 633) -- in addition to `stop`."""
    torch = ops.torch
    dev = ops.device
    F = q0.shape[0]
    w = 1. / ops.scales  # norm weighting
    q = q0.clone()
    # control state lives on the host (F is small; the step-size and refresh arithmetic is then literally the
    # reference's numpy arithmetic); q, the stage derivatives and the Newton iterates live on the device
    t = np.zeros(F)
    dt = np.full(F, float(first_time_step))
    nsteps = np.zeros(F, dtype=np.int64)
    setup_count = np.zeros(F, dtype=np.int64)
    refresh = np.ones(F, dtype=bool)
    going = np.ones(F, dtype=bool)
    attempts = np.zeros(F, dtype=np.int64)
    J = torch.zeros((F, ops.nelem), dtype=torch.float64, device=dev)
    L = torch.zeros((F, ops.nzi * ops.ns * ops.ns), dtype=torch.float64, device=dev)
    piv = torch.zeros((F, ops.ndof), dtype=torch.int32, device=dev)
    # on the device the Newton iterations use the explicit block inverses the factorisation forms anyway: the linear
    # solve only preconditions an iteration that converges on the true residual (to nonlinear_solve_tolerance)
    use_inv = bool(ops.on_device) and explicit_inverse_solves
    Dinv = torch.zeros_like(L) if use_inv else None
    ones = torch.ones((F, ops.ndof), dtype=torch.float64, device=dev)
    t_hist = [[0.] for _ in range(F)]
    q0_h = q0.cpu().numpy()
    q_hist = [[q0_h[f].copy()] for f in range(F)]
    residual_full = np.full(F, np.inf)

    def dev_idx(a):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.int64), device=dev)

    def dev_vec(a):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev)

    wcache = [None, None]

    def wnorm(x, widx):
        if wcache[0] is not widx:  # (the active set's weights are gathered once per step, not once per norm)
            wcache[0], wcache[1] = widx, w.index_select(0, widx)
        return (x * wcache[1]).abs().amax(dim=1)

    while True:
        idx_h = np.nonzero(going)[0]
        if idx_h.size == 0:
            break
        n = idx_h.size
        idx = dev_idx(idx_h)
        all_active = n == F
        qa = q if all_active else q.index_select(0, idx)
        if stop_at_time is not None:  # (the shortened step is the member's step size from here on, as in odesolve)
            dt[idx_h] = np.where(t[idx_h] + dt[idx_h] > stop_at_time, stop_at_time - t[idx_h], dt[idx_h])
        dta_h = dt[idx_h]
        dta = dev_vec(dta_h)
        # ---- projector: prefactor*J - I with prefactor = gamma*dt, for the members flagged for a refresh ----------------
        loc_h = np.nonzero(refresh[idx_h])[0]
        if loc_h.size:
            ij_h = idx_h[loc_h]
            ij = dev_idx(ij_h)
            Jn = ops.jac(qa if loc_h.size == n else qa.index_select(0, dev_idx(loc_h)), ij, key=tuple(ij_h.tolist()))
            Jn.mul_(dev_vec(dta_h[loc_h] * _G)[:, None])
            ops.add_to_block_diagonal(Jn, 1., ones[:ij_h.size], -1.)
            fact = ops.factorize(Jn, with_inverse=use_inv)
            J[ij], L[ij], piv[ij] = fact[:3]
            if use_inv:
                Dinv[ij] = fact[3]
        setup_count[idx_h] += 1
        factors = (J, L, piv, Dinv) if use_inv else (J, L, piv)
        key_all = tuple(idx_h.tolist())
        # ---- one ESDIRK64 step for every active member --------------------------------------------------------------------
        fused = bool(ops.on_device) and n <= dense_limit and (FUSED_NEWTON if fused_newton is None else fused_newton)
        if fused:
            # Device-resident stage loop (griffon_b200.h, "vector kernels of the batched implicit integrator"): one kernel
            # forms the explicit part and the first residual, and every Newton iteration is solve -> update -> rhs -> one
            # fused kernel (residual, select, weighted norm, convergence flags); the host reads one integer per iteration.
            gm = ops.gmod
            wa = w if all_active else w.index_select(0, idx)
            k = [ops.rhs(qa, idx, key=key_all)]
            qs = qa
            conv_i = torch.zeros(n, dtype=torch.int32, device=dev)
            left_d = torch.zeros(1, dtype=torch.int32, device=dev)
            expl, res = torch.empty_like(qa), torch.empty_like(qa)
            nl_ok = np.ones(n, dtype=bool)
            rows = None if all_active else idx
            stage_call = STAGE_CALL and use_inv and hasattr(ops, 'newton_stage')
            work = torch.empty((3,) + tuple(qa.shape), dtype=torch.float64, device=dev) if stage_call else None
            async_stages = stage_call and ASYNC_STAGES and hasattr(ops, 'esdirk_stages')
            if async_stages:
                K = torch.empty((6,) + tuple(qa.shape), dtype=torch.float64, device=dev)
                K[0].copy_(k[0])
                x, f = qa.clone(), k[0].clone()
                flags = torch.zeros((4, n), dtype=torch.int32, device=dev)  # stage, iterations, failed, done
                flags[0].fill_(1)
                gm.esdirk_stage_begin([K[0]], _A[1][:1], _G, dta, x, qa, f, expl, res, conv_i)
                left, _ = ops.esdirk_stages(factors, rows, idx, key_all, _A, qa, dta, _G, wa, nonlinear_solve_tolerance,
                                            max_nonlinear_iter, x, f, res, expl, K, flags[0], flags[1], flags[2],
                                            flags[3], work, left_d)
                if left:
                    raise RuntimeError('batched ESDIRK: the stage loop ended with members that are not done')
                nl_ok = ~(flags[2].cpu().numpy().astype(bool))
                k = [K[j] for j in range(6)]
            for s in range(1, 6 if not async_stages else 1):
                x, f = qs.clone(), k[-1].clone()
                gm.esdirk_stage_begin(k[:s], _A[s][:s], _G, dta, x, qa, f, expl, res, conv_i)
                left = n
                if stage_call:  # the whole Newton loop of the stage behind the C-ABI
                    left, _ = ops.newton_stage(factors, rows, idx, key_all, expl, qa, dta, _G, wa,
                                               nonlinear_solve_tolerance, max_nonlinear_iter, x, f, res, conv_i, work,
                                               left_d)
                for it in range(0 if stage_call else max_nonlinear_iter):
                    dx = ops.solve(factors, res, rows=rows)
                    xn = torch.empty_like(x)
                    gm.newton_update(x, dx, conv_i, xn, left_d)
                    fn = ops.rhs(xn, idx, key=key_all)
                    left = gm.newton_tail(fn, xn, expl, qa, dta, _G, wa, nonlinear_solve_tolerance, x, f, res, conv_i,
                                          left_d)
                    if left == 0:
                        break
                if left:
                    nl_ok &= conv_i.cpu().numpy().astype(bool)
                qs = x
                k.append(f)
            dq = torch.empty_like(qa)
            stats = torch.empty((3, n), dtype=torch.float64, device=dev)
            gm.esdirk_finish(k, _B, _BH, dta, wa, dq, stats)
            stats = stats.cpu().numpy()
        else:
            k = [ops.rhs(qa, idx, key=key_all)]
            qs = qa
            nl_ok = np.ones(n, dtype=bool)
            for s in range(1, 6):
                explicit = _A[s][s - 1] * k[s - 1]
                for j in range(s - 2, -1, -1):
                    explicit = explicit + _A[s][j] * k[j]
                # Newton with the lagged projector: x -= solve(res); res = dt*(gamma*f(x) + explicit) - (x - q_n)
                x = qs.clone()
                f = k[-1].clone()
                res = dta[:, None] * (_G * f + explicit) - (x - qa)
                conv = np.zeros(n, dtype=bool)
                conv_d = None
                for it in range(max_nonlinear_iter):
                    l_h = np.nonzero(~conv)[0]
                    if l_h.size == 0:
                        break
                    g_h = idx_h[l_h]
                    if l_h.size != n and n <= dense_limit and ops.on_device:
                        # Small batches are latency-bound: a kernel over all n members costs what a kernel over the
                        # unconverged ones costs, so the iteration runs on everybody and the converged members simply keep
                        # their values (no gathers, no index uploads; the unconverged members see the same arithmetic).
                        if conv_d is None:
                            conv_d = torch.as_tensor(conv, device=dev)
                        dx = ops.solve(factors, res, rows=None if all_active else idx)
                        xn = x - dx
                        fn = ops.rhs(xn, idx, key=key_all)
                        rn = dta[:, None] * (_G * fn + explicit) - (xn - qa)
                        keep = conv_d[:, None]
                        x, f, res = torch.where(keep, x, xn), torch.where(keep, f, fn), torch.where(keep, res, rn)
                        conv_d = conv_d | (wnorm(rn, idx) < nonlinear_solve_tolerance)
                        conv = conv_d.cpu().numpy()
                    elif l_h.size == n:
                        dx = ops.solve(factors, res, rows=None if all_active else idx)
                        xn = x - dx
                        fn = ops.rhs(xn, idx, key=key_all)
                        rn = dta[:, None] * (_G * fn + explicit) - (xn - qa)
                        x, f, res = xn, fn, rn
                        conv_d = wnorm(rn, idx) < nonlinear_solve_tolerance
                        conv = conv_d.cpu().numpy()
                    else:
                        loc, gl = dev_idx(l_h), dev_idx(g_h)
                        dx = ops.solve(factors, res.index_select(0, loc), rows=gl)
                        xn = x.index_select(0, loc) - dx
                        fn = ops.rhs(xn, gl, key=tuple(g_h.tolist()))
                        rn = dta.index_select(0, loc)[:, None] * (_G * fn + explicit.index_select(0, loc)) - \
                            (xn - qa.index_select(0, loc))
                        x[loc], f[loc], res[loc] = xn, fn, rn
                        conv[l_h] = (wnorm(rn, gl) < nonlinear_solve_tolerance).cpu().numpy()
                        conv_d = None
                nl_ok &= conv
                qs = x
                k.append(f)
            dq = dta[:, None] * (_B[0] * k[0] + _B[1] * k[1] + _B[2] * k[2] + _B[3] * k[3] + _B[4] * k[4] + _B[5] * k[5])
            dqh = dta[:, None] * (_BH[0] * k[0] + _BH[1] * k[1] + _BH[2] * k[2] + _BH[3] * k[3] + _BH[4] * k[4] +
                                  _BH[5] * k[5])
            stats = torch.stack([wnorm(dq - dqh, idx), wnorm(dq, idx), torch.isfinite(dq).all(dim=1).to(torch.float64)])
            stats = stats.cpu().numpy()
        err, ok = stats[0], stats[2] > 0.5
        with np.errstate(all='ignore'):
            residual = stats[1] / dta_h
        # ---- accepted members --------------------------------------------------------------------------------------------------
        a_h = np.nonzero(ok)[0]
        if a_h.size:
            ga_h = idx_h[a_h]
            sel = a_h.size != n
            qnew = (qa.index_select(0, dev_idx(a_h)) + dq.index_select(0, dev_idx(a_h))) if sel else qa + dq
            if clip_negative:
                qnew = torch.where(qnew < 0., torch.zeros_like(qnew), qnew)
            if sel or not all_active:
                q[dev_idx(ga_h)] = qnew
            else:
                q = qnew
            d = dta_h[a_h]
            t[ga_h] = t[ga_h] + d
            nsteps[ga_h] += 1
            # PI controller, stepcontrol.py:84-101 (its proportional factor divides the newest error by itself: one)
            e = err[a_h]
            with np.errstate(all='ignore'):
                ratio = (transient_tolerance / e) ** ki
            dnew = np.minimum(d * np.minimum(max_ramp, ratio), max_time_step)
            dnew = np.where(e < 1.e-16, np.minimum(d * max_ramp, max_time_step), dnew)
            # Jacobian-refresh policy, integrator.py:100-130
            cnt = setup_count[ga_h]
            okn = nl_ok[a_h]
            by_count = cnt == maximum_steps_per_jacobian
            # a Newton solve that ran out of iterations reports slow convergence (nonlinear.py:259-268), so it takes the
            # `nlslowness` branch and the slow-solve factor; fail_factor is for steps that are rejected outright (below)
            by_slow = ~by_count & ~okn
            by_size = ~by_count & okn & ((dnew > d * grow_limit) | (dnew < d * shrink_limit))
            dnew = np.where(by_slow, dnew * slow_factor, dnew)
            refresh[ga_h] = by_count | by_slow | by_size
            setup_count[ga_h] = np.where(by_count, 0, cnt)
            dt[ga_h] = dnew
            if save_each_step:
                qh = qnew.cpu().numpy()
                for m, f_ in enumerate(ga_h.tolist()):
                    t_hist[f_].append(float(t[f_]))
                    q_hist[f_].append(qh[m].copy())
        r_h = idx_h[np.nonzero(~ok)[0]]
        if r_h.size:
            dt[r_h] = dt[r_h] * fail_factor
            refresh[r_h] = True
        attempts[idx_h] += 1
        # ---- stopping --------------------------------------------------------------------------------------------------------------
        residual_full[idx_h] = np.where(np.isfinite(residual), residual, np.inf)
        done = stop(dev_vec(t), q, dev_vec(residual_full), dev_idx(nsteps)).cpu().numpy()
        if not stop_ignores_minimum:  # (odesolve's stop_at_steady test is not subject to the minimum, integrator.py:633-640)
            done = done & (nsteps >= minimum_time_step_count)
        if stop_at_time is not None:
            done = done | (t >= stop_at_time)
        done = done | (attempts > maximum_steps)
        going = going & ~done
    failed = attempts > maximum_steps
    if not save_each_step:
        qf = q.cpu().numpy()
        for f_ in range(F):
            t_hist[f_].append(float(t[f_]))
            q_hist[f_].append(qf[f_].copy())
    return [np.array(th) for th in t_hist], [np.array(qh) for qh in q_hist], failed


# ----------------------------------------------------------------------------------------------------------------------
# Members that advance independently of each other across steps.
#
# In integrate_batch every Newton iteration of a step is a round of latency-bound kernels that serves all members, and a
# step takes as many rounds as its SLOWEST member needs -- a different member at different times (measured on BASELINE
# config 5: 15.8 k rounds for 56 members against 6.4 k Newton iterations of the longest trajectory on its own). Here a
# round still serves all members, but every member is wherever its own trajectory is: its own step, its own stage, its
# own Newton iteration (device state machine, csrc/gb_newton.cu: k_async_update / k_async_tail). When a member
# completes the stages of a step the host does for it exactly what integrate_batch does for everybody at once -- error
# estimate, acceptance, PI controller, Jacobian-refresh policy, stopping test -- and sends it into its next step; a
# Jacobian refresh (Jacobian, scaling, Gauss-Jordan elimination: ~3.5 ms, seven rounds' worth) runs on a second stream
# with a second Griffon handle while the other members keep iterating. Every member's arithmetic is the sequence
# integrate_batch performs for it, so the trajectories are identical bit for bit (tests/test_gpu_newton.py).
ASYNC_MEMBERS = _os.environ.get('GB_ASYNC_MEMBERS', '1') != '0'
# Jacobian refreshes of the asynchronous integrator written and eliminated in place in the batch's factor arrays
# (GB_DIRECT_REFRESH=0: through temporaries and tensor copies, as the lock-step batch does; same numbers)
DIRECT_REFRESH = _os.environ.get('GB_DIRECT_REFRESH', '1') != '0'


LAST_ASYNC_STATS = dict()


def can_integrate_async(ops, F, explicit_inverse_solves=True):
    return bool(ASYNC_MEMBERS and getattr(ops, 'on_device', False) and explicit_inverse_solves and F > 1 and
                hasattr(ops, 'second_griffon') and hasattr(ops.g, 'flamelet_async_tick_batch'))


def integrate_batch_async(ops, q0, stop, first_time_step=1.e-6, max_time_step=1.e-3, minimum_time_step_count=40,
                          transient_tolerance=1.e-10, maximum_steps_per_jacobian=10, nonlinear_solve_tolerance=1.e-12,
                          max_nonlinear_iter=20, max_ramp=1.1, ki=0.1333333333, maximum_steps=100000,
                          fail_factor=0.8, slow_factor=0.8, grow_limit=1.05, shrink_limit=0.9, clip_negative=True,
                          save_each_step=True, stop_ignores_minimum=False, stop_at_time=None, stats_out=None):
    """integrate_batch with the members advancing independently of each other (device path, inverse-based solves).
    Same arguments, same results. stats_out (dict, optional) receives the number of rounds and refresh launches."""
    import ctypes as C
    torch, dev, gm = ops.torch, ops.device, ops.gmod
    F, ndof = q0.shape
    nst = 6
    w = (1. / ops.scales).contiguous()
    q = q0.clone()
    t = np.zeros(F)
    dt = np.full(F, float(first_time_step))
    nsteps = np.zeros(F, dtype=np.int64)
    setup_count = np.zeros(F, dtype=np.int64)
    refresh = np.ones(F, dtype=bool)
    attempts = np.zeros(F, dtype=np.int64)
    residual_full = np.full(F, np.inf)
    NEED, WAIT, RUN, FIN = 0, 1, 2, 3
    phase = np.full(F, NEED)
    f64 = dict(dtype=torch.float64, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    J = torch.zeros((F, ops.nelem), **f64)
    L = torch.zeros((F, ops.nzi * ops.ns * ops.ns), **f64)
    Dinv = torch.zeros_like(L)
    K = torch.zeros((nst, F, ndof), **f64)
    x, f, res, expl = q.clone(), torch.zeros_like(q), torch.zeros_like(q), torch.zeros_like(q)
    work = torch.zeros((3, F, ndof), **f64)
    work[1].copy_(q)  # (the kernels' x - dx buffer starts from a finite state for members that are not active yet)
    dt_d, dtin_d = torch.zeros(F, **f64), torch.zeros(F, **f64)
    state_d, stage_d, iters_d, nlfail_d, nits_d, start_d = (torch.zeros(F, **i32) for _ in range(6))
    dq = torch.zeros_like(q)
    stats = torch.zeros((3, F), **f64)
    ones = torch.ones((F, ndof), **f64)
    # host side of a tick (pinned): what the host sends (start flags, step sizes) and what comes back. Two sets: the
    # host processes the results of tick k (and prepares tick k+2) while tick k+1 runs
    pin = lambda shape, dtype: torch.zeros(shape, dtype=dtype).pin_memory()
    addr = lambda ten: C.c_void_p(ten.data_ptr())
    hps = [dict(start=pin((F,), torch.int32), dt=pin((F,), torch.float64), state=pin((F,), torch.int32),
                stage=pin((F,), torch.int32), stats=pin((3, F), torch.float64), nlfail=pin((F,), torch.int32),
                q=pin((F, ndof), torch.float64)) for _ in range(2)]
    hvs = [{k_: v.numpy() for k_, v in hp.items()} for hp in hps]
    hosts = [{k_: addr(v) for k_, v in hp.items()} for hp in hps]
    devp = dict(J=addr(J), L=addr(L), Dinv=addr(Dinv), q=addr(q), dt=addr(dt_d), w=addr(w), x=addr(x), f=addr(f),
                res=addr(res), expl=addr(expl), K=addr(K), state=addr(state_d), stage=addr(stage_d), iters=addr(iters_d),
                nlfail=addr(nlfail_d), nits=addr(nits_d), work=addr(work), dq=addr(dq), stats=addr(stats),
                start=addr(start_d), dtin=addr(dtin_d))
    tab = (C.c_double * (nst * nst))(*[float(_A[a][b]) for a in range(nst) for b in range(nst)])
    b_c, bh_c = (C.c_double * nst)(*_B), (C.c_double * nst)(*_BH)
    prm_all, keep_all = ops._params(ops._all(), tuple(range(F)))
    # Jacobian refreshes run on a few side streams, round robin, each with its own Griffon handle (a handle serves one
    # stream at a time): a refresh is ~3.5 ms on one SM per member and there are several thousand of them, so one side
    # stream would serialise more work than the whole integration takes
    n_side = 8
    sides = [(torch.cuda.Stream(), ops.second_griffon(k)) for k in range(n_side)]
    direct_refresh = DIRECT_REFRESH and getattr(ops, 'gauss_jordan_inverses', False)
    ctrl = torch.cuda.Stream()  # the stopping test runs here: it must not queue behind the rounds in flight
    device = torch.cuda.current_device()
    pending = []  # (event, member ids) of the Jacobian refreshes in flight
    t_hist = [[0.] for _ in range(F)]
    q0_h = q0.cpu().numpy()
    q_hist = [[q0_h[m].copy()] for m in range(F)]
    n_rounds = n_refresh = 0
    host_wait = host_ctrl = 0.  # seconds the host waited for a tick / spent on the step-end control
    _t_begin = _time.perf_counter()
    pack_d = torch.zeros((3, F), **f64)  # t, residual, step count of every member, for the stopping test
    torch.cuda.synchronize()

    # The ticks run on a worker thread (the C call releases the interpreter lock): while tick k+1 iterates on the device
    # the host does the step-end control of the members that completed in tick k and launches their Jacobian refreshes.
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(1)

    def run_tick(bi, with_start, max_rounds, members):
        torch.cuda.set_device(device)
        return ops.g.flamelet_async_tick_batch(F, prm_all, devp, hosts[bi], tab, b_c, bh_c, nst, _G,
                                               nonlinear_solve_tolerance, max_nonlinear_iter, clip_negative, max_rounds,
                                               with_start, members)

    def step_end(hv):
        """what integrate_batch does for everybody at once, for the members that completed a step in this tick"""
        fin = np.nonzero((phase == RUN) & (hv['state'] == 0) & (hv['stage'] == nst))[0]
        if fin.size == 0:
            return
        st_h = hv['stats'][:, fin]
        nl_ok = hv['nlfail'][fin] == 0
        d_all = dt[fin]
        err, ok = st_h[0], st_h[2] > 0.5
        with np.errstate(all='ignore'):
            residual = st_h[1] / d_all
        acc = fin[ok]
        if acc.size:
            d = d_all[ok]
            t[acc] = t[acc] + d
            nsteps[acc] += 1
            e = err[ok]
            with np.errstate(all='ignore'):
                ratio = (transient_tolerance / e) ** ki
            dnew = np.minimum(d * np.minimum(max_ramp, ratio), max_time_step)
            dnew = np.where(e < 1.e-16, np.minimum(d * max_ramp, max_time_step), dnew)
            cnt = setup_count[acc]
            okn = nl_ok[ok]
            by_count = cnt == maximum_steps_per_jacobian
            by_slow = ~by_count & ~okn
            by_size = ~by_count & okn & ((dnew > d * grow_limit) | (dnew < d * shrink_limit))
            dnew = np.where(by_slow, dnew * slow_factor, dnew)
            refresh[acc] = by_count | by_slow | by_size
            setup_count[acc] = np.where(by_count, 0, cnt)
            dt[acc] = dnew
            if save_each_step:
                for m in acc.tolist():
                    t_hist[m].append(float(t[m]))
                    q_hist[m].append(hv['q'][m].copy())
        rej = fin[~ok]
        if rej.size:
            dt[rej] = dt[rej] * fail_factor
            refresh[rej] = True
        attempts[fin] += 1
        residual_full[fin] = np.where(np.isfinite(residual), residual, np.inf)
        stop_host = getattr(stop, 'host', None)
        if stop_host is not None:
            # the stopping test on the host copies of the completed members' rows (they came back with the tick): no
            # device work, no synchronisation
            done = np.zeros(F, dtype=bool)
            done[fin] = stop_host(t[fin], hv['q'][fin], residual_full[fin], nsteps[fin])
        else:
            with torch.cuda.stream(ctrl):  # (the rows of these members are final: their tick has been synchronised)
                pack_d.copy_(torch.from_numpy(np.stack([t, residual_full, nsteps.astype(np.float64)])))
                done = stop(pack_d[0], q, pack_d[1], pack_d[2].to(torch.int64)).cpu().numpy()
        if not stop_ignores_minimum:
            done = done & (nsteps >= minimum_time_step_count)
        if stop_at_time is not None:
            done = done | (t >= stop_at_time)
        done = done | (attempts > maximum_steps)
        phase[fin] = np.where(done[fin], FIN, NEED)

    inflight, bi = None, 0
    try:
        while True:
            # ---- members at the start of a step: shorten it to the final time, refresh the projector if flagged --------
            need = np.nonzero(phase == NEED)[0]
            start = []
            if need.size:
                if stop_at_time is not None:
                    dt[need] = np.where(t[need] + dt[need] > stop_at_time, stop_at_time - t[need], dt[need])
                rf = need[refresh[need]]
                if rf.size and direct_refresh:
                    # dt gamma J - I straight out of the Jacobian kernel into the member's rows of the batch's arrays,
                    # eliminated in place there: two launches per member, no temporaries, no copies (the rounds in
                    # flight read these rows for the idle member and discard what they compute). One side stream per
                    # member, round robin: the members of a refresh do not wait for each other.
                    for m in rf.tolist():
                        side, g2 = sides[n_refresh % n_side]
                        with torch.cuda.stream(side):
                            ops.jac_scaled_row_on(g2, q, m, float(dt[m] * _G), J)
                            ops.invert_row(J, L, Dinv, m)
                            ev = torch.cuda.Event()
                            ev.record(side)
                        pending.append((ev, np.array([m])))
                        n_refresh += 1
                    phase[rf] = WAIT
                elif rf.size:
                    side, g2 = sides[n_refresh % n_side]
                    with torch.cuda.stream(side):
                        # (the members' states are final: the tick that accepted them has been synchronised. Nothing is
                        # copied from the host here: a synchronous copy would wait for the refresh queued before it)
                        if True:
                            Jn = torch.empty((rf.size, ops.nelem), **f64)
                            ops.jac_rows_on(g2, q, rf.tolist(), Jn)
                            for k_, m in enumerate(rf.tolist()):
                                Jn[k_].mul_(float(dt[m] * _G))
                            ops.add_to_block_diagonal(Jn, 1., ones[:rf.size], -1.)
                            fact = ops.factorize(Jn, with_inverse=True)
                            for k_, m in enumerate(rf.tolist()):
                                J[m].copy_(fact[0][k_]), L[m].copy_(fact[1][k_]), Dinv[m].copy_(fact[3][k_])
                        ev = torch.cuda.Event()
                        ev.record(side)
                    pending.append((ev, rf))
                    phase[rf] = WAIT
                    n_refresh += 1
                start.extend(need[~refresh[need]].tolist())
                setup_count[need] += 1
            # ---- refreshes that have completed -----------------------------------------------------------------------------
            if pending:
                idle = not start and inflight is None and not np.any(phase == RUN)
                still = []
                for ev, rf in pending:
                    if idle and not start:
                        ev.synchronize()
                    if ev.query():
                        start.extend(rf.tolist())
                    else:
                        still.append((ev, rf))
                pending = still
            # ---- next tick: start those members, rounds of kernels until a member completes its stages --------------------
            hv = hvs[bi]
            hv['start'][:] = 0
            if start:
                sid = np.array(start, dtype=np.int64)
                hv['start'][sid] = 1
                hv['dt'][sid] = dt[sid]
                phase[sid] = RUN
                refresh[sid] = False  # (set again by the policy at the end of the step)
            nxt = None
            if np.any(phase == RUN):
                # (the members inside a step: a superset of those the device still iterates on -- the ones that completed
                # in the tick in flight are idle there -- and all the right-hand side kernel of a round needs to visit)
                running = np.ascontiguousarray(np.nonzero(phase == RUN)[0], dtype=np.int32)
                nxt = (pool.submit(run_tick, bi, bool(start), 1 if (pending or inflight is not None) else 16, running), bi)
                bi ^= 1
            # ---- meanwhile: the step ends of the tick that has just finished ------------------------------------------------
            if inflight is not None:
                _t0 = _time.perf_counter()
                n_rounds += inflight[0].result()
                _t1 = _time.perf_counter()
                step_end(hvs[inflight[1]])
                host_wait += _t1 - _t0
                host_ctrl += _time.perf_counter() - _t1
            inflight = nxt
            if inflight is None and not pending and not np.any(phase == NEED):
                if np.all(phase == FIN):
                    break
    finally:
        pool.shutdown(wait=True)
    main = torch.cuda.current_stream()
    main.wait_stream(ctrl)
    del keep_all
    for side, _ in sides:
        main.wait_stream(side)
    failed = attempts > maximum_steps
    if not save_each_step:
        qf = q.cpu().numpy()
        for m in range(F):
            t_hist[m].append(float(t[m]))
            q_hist[m].append(qf[m].copy())
    LAST_ASYNC_STATS.update(rounds=n_rounds, refresh_launches=n_refresh, members=F, host_wait_s=host_wait,
                            host_step_end_s=host_ctrl, wall_s=_time.perf_counter() - _t_begin)
    if stats_out is not None:
        stats_out.update(rounds=n_rounds, refresh_launches=n_refresh, newton_iterations=nits_d.cpu().numpy().tolist())
    return [np.array(th) for th in t_hist], [np.array(qh) for qh in q_hist], failed
