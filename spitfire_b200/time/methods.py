"""Implicit Runge-Kutta steppers used by the chemistry path (mirror of src/spitfire/time/methods.py:35-125, 431-612).

`KennedyCarpenterS6P4Q3` is the six-stage, fourth-order ESDIRK with a third-order embedded error estimate of Kennedy &
Carpenter; the tableau coefficients are the published ones (methods.py:459-487)."""
import numpy as np
from numpy import inf
from scipy.linalg import norm


class StepOutput(object):
    """result of one time step (methods.py:85-125)"""
    __slots__ = ['solution_update', 'temporal_error', 'nonlinear_iter', 'linear_iter', 'nonlinear_converged',
                 'slow_nonlinear_convergence', 'projector_setups', 'extra_errors']

    def __init__(self, **kwargs):
        for slot in self.__slots__:
            setattr(self, slot, kwargs.get(slot, None))
        if self.temporal_error is None:
            self.temporal_error = -1.


class TimeStepperBase(object):
    """descriptor of a stepping method (methods.py:35-82)"""

    def __init__(self, name, order, n_stages=1, is_adaptive=False, norm_weighting=1., norm_order=inf,
                 is_implicit=False, nonlinear_solver=None, implicit_coefficient=None):
        self.name, self.order, self.n_stages = name, order, n_stages
        self.is_adaptive, self.norm_weighting, self.norm_order = is_adaptive, norm_weighting, norm_order
        self.is_implicit, self.nonlinear_solver = is_implicit, nonlinear_solver
        self.implicit_coefficient = implicit_coefficient

    def norm(self, x):
        return norm(x * self.norm_weighting, ord=self.norm_order)


class _DiagonallyImplicit(TimeStepperBase):
    """shared driver of (E)SDIRK steps: stage s solves  dt*(gamma*f(q_s) + sum_{j<s} a_sj k_j) - (q_s - q_n) = 0."""

    def _implicit_stage(self, rhs, lhs_setup, lhs_solve, t, dt, c, q_n, explicit_part, guess, guess_rhs, tally):
        gamma = self.gamma

        def residual(q, existing_rhs=None, evaluate_new_rhs=True):
            f = rhs(t + c * dt, q) if evaluate_new_rhs else existing_rhs
            return dt * (gamma * f + explicit_part) - (q - q_n), f

        out = self.nonlinear_solver(residual_method=residual,
                                    setup_method=lambda q: lhs_setup(t + c * dt, q),
                                    solve_method=lhs_solve, initial_guess=guess, initial_rhs=guess_rhs)
        tally['nonlinear_iter'] += out.iter
        tally['linear_iter'] += out.liter
        tally['converged'] = tally['converged'] and out.converged
        tally['slow'] = tally['slow'] or out.slow_convergence
        tally['setups'] += out.projector_setups
        return out.solution, out.rhs_at_converged


class BackwardEulerS1P1Q1(_DiagonallyImplicit):
    """backward Euler with the classical first-order error estimate (methods.py:405-428)"""

    def __init__(self, nonlinear_solver, norm_weighting=1., norm_order=inf):
        super().__init__(name='backward Euler', order=1, n_stages=1, is_implicit=True, implicit_coefficient=1.,
                         nonlinear_solver=nonlinear_solver, is_adaptive=True, norm_weighting=norm_weighting,
                         norm_order=norm_order)
        self.gamma = 1.

    def single_step(self, state, t, dt, rhs, lhs_setup, lhs_solve, *args, **kwargs):
        tally = dict(nonlinear_iter=0, linear_iter=0, converged=True, slow=False, setups=0)
        q_n = np.copy(state)
        f_n = rhs(t, state)
        q, f = self._implicit_stage(rhs, lhs_setup, lhs_solve, t, dt, 1., q_n, 0. * f_n, state, f_n, tally)
        dstate = dt * f
        return StepOutput(solution_update=dstate, temporal_error=self.norm(dstate - dt * f_n),
                          nonlinear_iter=tally['nonlinear_iter'], linear_iter=tally['linear_iter'],
                          nonlinear_converged=tally['converged'], slow_nonlinear_convergence=tally['slow'],
                          projector_setups=tally['setups'])


class KennedyCarpenterS6P4Q3(_DiagonallyImplicit):
    """Kennedy/Carpenter ESDIRK64: explicit first stage, five implicit stages with gamma = 1/4, stiffly accurate"""

    def __init__(self, nonlinear_solver, norm_weighting=1., norm_order=inf):
        super().__init__(name='Kennedy/Carpenter ESDIRK64', order=4, n_stages=6, is_implicit=True,
                         implicit_coefficient=0.25, nonlinear_solver=nonlinear_solver, is_adaptive=True,
                         norm_weighting=norm_weighting, norm_order=norm_order)
        g = 0.25
        self.gamma = g
        self.A = np.array([
            [0., 0., 0., 0., 0., 0.],
            [0.25, g, 0., 0., 0., 0.],
            [8611. / 62500., -1743. / 31250., g, 0., 0., 0.],
            [5012029. / 34652500., -654441. / 2922500., 174375. / 388108., g, 0., 0.],
            [15267082809. / 155376265600., -71443401. / 120774400., 730878875. / 902184768., 2285395. / 8070912., g,
             0.],
            [82889. / 524892., 0., 15625. / 83664., 69875. / 102672., -2260. / 8211., g]])
        self.c = np.sum(self.A, axis=1)
        self.b = np.copy(self.A[5])
        self.bh = np.array([4586570599. / 29645900160., 0., 178811875. / 945068544., 814220225. / 1159782912.,
                            -3700637. / 11593932., 61727. / 225920.])

    def single_step(self, state, t, dt, rhs, lhs_setup, lhs_solve, *args, **kwargs):
        A = self.A
        tally = dict(nonlinear_iter=0, linear_iter=0, converged=True, slow=False, setups=0)
        q_n = np.copy(state)
        k = [rhs(t, state)]
        q = state
        for s in range(1, 6):
            # explicit part summed from the newest stage to the oldest, as the reference writes it (methods.py:521-584)
            explicit_part = A[s, s - 1] * k[s - 1]
            for j in range(s - 2, -1, -1):
                explicit_part = explicit_part + A[s, j] * k[j]
            q, f = self._implicit_stage(rhs, lhs_setup, lhs_solve, t, dt, self.c[s], q_n, explicit_part, q, k[-1],
                                        tally)
            k.append(f)
        b, bh = self.b, self.bh
        dstate = dt * (b[0] * k[0] + b[1] * k[1] + b[2] * k[2] + b[3] * k[3] + b[4] * k[4] + b[5] * k[5])
        dstate_h = dt * (bh[0] * k[0] + bh[1] * k[1] + bh[2] * k[2] + bh[3] * k[3] + bh[4] * k[4] + bh[5] * k[5])
        return StepOutput(solution_update=dstate, temporal_error=self.norm(dstate - dstate_h),
                          nonlinear_iter=tally['nonlinear_iter'], linear_iter=tally['linear_iter'],
                          nonlinear_converged=tally['converged'], slow_nonlinear_convergence=tally['slow'],
                          projector_setups=tally['setups'])
