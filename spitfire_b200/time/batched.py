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


def integrate_batch(ops, q0, stop, first_time_step=1.e-6, max_time_step=1.e-3, minimum_time_step_count=40,
                    transient_tolerance=1.e-10, maximum_steps_per_jacobian=10, nonlinear_solve_tolerance=1.e-12,
                    max_nonlinear_iter=20, max_ramp=1.1, ki=0.1333333333, maximum_steps=100000,
                    fail_factor=0.8, slow_factor=0.8, grow_limit=1.05, shrink_limit=0.9, clip_negative=True,
                    explicit_inverse_solves=True):
    """advance every member of `ops` (flamelet._BatchOps) from q0 [F, ndof] until `stop(t, q, residual, nsteps)` (all
    [F]-shaped tensors; returns a bool tensor) holds for it and it has taken at least minimum_time_step_count steps.
    Returns per member the lists of saved times and states (numpy), initial state included, and a `failed` flag
    (non-finite update that the step-size reduction could not cure within maximum_steps)."""
    torch = ops.torch
    dev = ops.device
    F = q0.shape[0]
    w = 1. / ops.scales  # norm weighting
    q = q0.clone()
    t = torch.zeros(F, dtype=torch.float64, device=dev)
    dt = torch.full((F,), float(first_time_step), dtype=torch.float64, device=dev)
    nsteps = torch.zeros(F, dtype=torch.int64, device=dev)
    setup_count = torch.zeros(F, dtype=torch.int64, device=dev)
    refresh = torch.ones(F, dtype=torch.bool, device=dev)
    going = torch.ones(F, dtype=torch.bool, device=dev)
    attempts = torch.zeros(F, dtype=torch.int64, device=dev)
    J = torch.zeros((F, ops.nelem), dtype=torch.float64, device=dev)
    L = torch.zeros((F, ops.nzi * ops.ns * ops.ns), dtype=torch.float64, device=dev)
    piv = torch.zeros((F, ops.ndof), dtype=torch.int32, device=dev)
    # on the device the Newton iterations use the explicit block inverses the factorisation forms anyway: the linear
    # solve only preconditions an iteration that converges on the true residual (to nonlinear_solve_tolerance)
    use_inv = bool(ops.on_device) and explicit_inverse_solves
    Dinv = torch.zeros_like(L) if use_inv else None
    ones = torch.ones((F, ops.ndof), dtype=torch.float64, device=dev)
    t_hist = [[0.] for _ in range(F)]
    q_hist = [[q0[f].cpu().numpy().copy()] for f in range(F)]

    def wnorm(x, idx):
        return (x * w.index_select(0, idx)).abs().amax(dim=1)

    while True:
        idx = torch.nonzero(going).flatten()
        if idx.numel() == 0:
            break
        n = idx.numel()
        qa, dta = q.index_select(0, idx), dt.index_select(0, idx)
        # ---- projector: prefactor*J - I with prefactor = gamma*dt, for the members flagged for a refresh ----------------
        ij_local = torch.nonzero(refresh.index_select(0, idx)).flatten()
        if ij_local.numel():
            ij = idx.index_select(0, ij_local)
            Jn = ops.jac(qa.index_select(0, ij_local), ij)
            Jn.mul_((dta.index_select(0, ij_local) * _G)[:, None])
            ops.add_to_block_diagonal(Jn, 1., ones[:ij.numel()], -1.)
            fact = ops.factorize(Jn, with_inverse=use_inv)
            J[ij], L[ij], piv[ij] = fact[:3]
            if use_inv:
                Dinv[ij] = fact[3]
        setup_count[idx] += 1
        # ---- one ESDIRK64 step for every active member --------------------------------------------------------------------
        k = [ops.rhs(qa, idx)]
        qs = qa
        nl_ok = torch.ones(n, dtype=torch.bool, device=dev)
        for s in range(1, 6):
            explicit = _A[s][s - 1] * k[s - 1]
            for j in range(s - 2, -1, -1):
                explicit = explicit + _A[s][j] * k[j]
            # Newton with the lagged projector: x -= solve(res); res = dt*(gamma*f(x) + explicit) - (x - q_n)
            x = qs.clone()
            f = k[-1].clone()
            res = dta[:, None] * (_G * f + explicit) - (x - qa)
            conv = torch.zeros(n, dtype=torch.bool, device=dev)
            for it in range(max_nonlinear_iter):
                loc = torch.nonzero(~conv).flatten()
                if loc.numel() == 0:
                    break
                gl = idx.index_select(0, loc)
                dx = ops.solve((J, L, piv, Dinv) if use_inv else (J, L, piv), res.index_select(0, loc), rows=gl)
                xn = x.index_select(0, loc) - dx
                fn = ops.rhs(xn, gl)
                rn = dta.index_select(0, loc)[:, None] * (_G * fn + explicit.index_select(0, loc)) - \
                    (xn - qa.index_select(0, loc))
                x[loc], f[loc], res[loc] = xn, fn, rn
                conv[loc] = wnorm(rn, gl) < nonlinear_solve_tolerance
            nl_ok &= conv
            qs = x
            k.append(f)
        dq = dta[:, None] * (_B[0] * k[0] + _B[1] * k[1] + _B[2] * k[2] + _B[3] * k[3] + _B[4] * k[4] + _B[5] * k[5])
        dqh = dta[:, None] * (_BH[0] * k[0] + _BH[1] * k[1] + _BH[2] * k[2] + _BH[3] * k[3] + _BH[4] * k[4] +
                              _BH[5] * k[5])
        err = wnorm(dq - dqh, idx)
        residual = wnorm(dq, idx) / dta
        ok = torch.isfinite(dq).all(dim=1)
        # ---- accepted members --------------------------------------------------------------------------------------------------
        acc = torch.nonzero(ok).flatten()
        ga = idx.index_select(0, acc)
        if acc.numel():
            qnew = qa.index_select(0, acc) + dq.index_select(0, acc)
            if clip_negative:
                qnew = torch.where(qnew < 0., torch.zeros_like(qnew), qnew)
            q[ga] = qnew
            t[ga] = t.index_select(0, ga) + dta.index_select(0, acc)
            nsteps[ga] += 1
            # PI controller (the proportional factor of the reference evaluates to one: it divides the error by itself)
            e = err.index_select(0, acc)
            d = dta.index_select(0, acc)
            ratio = (transient_tolerance / e) ** ki
            dnew = torch.minimum(d * torch.clamp(ratio, max=max_ramp), torch.full_like(d, max_time_step))
            dnew = torch.where(e < 1.e-16, torch.minimum(d * max_ramp, torch.full_like(d, max_time_step)), dnew)
            # Jacobian-refresh policy
            cnt = setup_count.index_select(0, ga)
            okn = nl_ok.index_select(0, acc)
            by_count = cnt == maximum_steps_per_jacobian
            by_fail = ~by_count & ~okn
            by_size = ~by_count & okn & ((dnew > d * grow_limit) | (dnew < d * shrink_limit))
            dnew = torch.where(by_fail, dnew * fail_factor, dnew)
            refresh[ga] = by_count | by_fail | by_size
            setup_count[ga] = torch.where(by_count, torch.zeros_like(cnt), cnt)
            dt[ga] = dnew
            th, qh = t.index_select(0, ga).cpu().numpy(), qnew.cpu().numpy()
            for m, f_ in enumerate(ga.tolist()):
                t_hist[f_].append(float(th[m]))
                q_hist[f_].append(qh[m].copy())
        rej = torch.nonzero(~ok).flatten()
        if rej.numel():
            gr = idx.index_select(0, rej)
            dt[gr] = dt.index_select(0, gr) * fail_factor
            refresh[gr] = True
        attempts[idx] += 1
        # ---- stopping --------------------------------------------------------------------------------------------------------------
        res_full = torch.zeros(F, dtype=torch.float64, device=dev)
        res_full[idx] = torch.where(torch.isfinite(residual), residual, torch.full_like(residual, float('inf')))
        done = stop(t, q, res_full, nsteps) & (nsteps >= minimum_time_step_count)
        done = done | (attempts > maximum_steps)
        going = going & ~done
        going[idx] = going.index_select(0, idx)
    failed = (attempts > maximum_steps).cpu().numpy()
    return [np.array(th) for th in t_hist], [np.array(qh) for qh in q_hist], failed
