"""Newton iteration used inside the implicit stages (mirror of src/spitfire/time/nonlinear.py:20-268)."""
import numpy as np
from numpy import inf
from scipy.linalg import norm


def finite_difference_jacobian(residual_func, residual_value, state, offset_rel=1.e-5, offset_abs=1.e-8):
    """one-sided finite-difference Jacobian, column i perturbed by offset_rel*|q_i| + offset_abs (nonlinear.py:20-49)"""
    n = state.size
    jac = np.ndarray((n, n))
    for i in range(n):
        h = offset_rel * np.abs(state[i]) + offset_abs
        q = np.copy(state)
        q[i] += h
        jac[:, i] = (residual_func(q) - residual_value) / h
    return jac


class SolverOutput(object):
    """result of a nonlinear solve (nonlinear.py:52-84)"""
    __slots__ = ['solution', 'rhs_at_converged', 'iter', 'liter', 'converged', 'slow_convergence', 'projector_setups']

    def __init__(self, **kwargs):
        for slot in self.__slots__:
            setattr(self, slot, kwargs.get(slot, None))


class NonlinearSolver(object):
    """options shared by nonlinear solvers (nonlinear.py:87-140)"""
    defaults = {'max_nonlinear_iter': 20, 'slowness_detection_iter': inf, 'must_converge': False, 'tolerance': 1.e-12,
                'norm_weighting': 1., 'norm_order': inf, 'raise_naninf': False, 'custom_solution_check': None,
                'setup_projector_in_governor': True}

    def __init__(self, *args, **kwargs):
        for key, value in self.defaults.items():
            setattr(self, key, kwargs.get(key, value))

    def _guard(self, x, message):
        if self.raise_naninf and not np.all(np.isfinite(x)):
            raise ValueError(message)

    def _custom(self, solution, message):
        if self.custom_solution_check is not None:
            self.custom_solution_check(solution, message)


class SimpleNewtonSolver(NonlinearSolver):
    """Newton's method with a lagged (or per-iteration) linear projector (nonlinear.py:143-268).

    Each iteration solves with the projector handed in (`solve_method(residual) -> (dstate, n_linear, converged)`),
    SUBTRACTS the update (the projector is of the form gamma*dt*J - I) and re-evaluates the residual; convergence is
    ||residual * norm_weighting|| < tolerance."""

    def __init__(self, evaluate_jacobian_every_iter=False, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.evaluate_jacobian_every_iter = evaluate_jacobian_every_iter

    @property
    def evaluate_jacobian_every_iter(self):
        return self._evaluate_jacobian_every_iter

    @evaluate_jacobian_every_iter.setter
    def evaluate_jacobian_every_iter(self, value):
        self._evaluate_jacobian_every_iter = value
        self.setup_projector_in_governor = not value

    def __call__(self, residual_method, setup_method, solve_method, initial_guess, initial_rhs):
        x = np.copy(initial_guess)
        self._guard(x, 'NaN or Inf detected in Simple Newton Solve: in the initial solution!')
        self._custom(x, 'In the initial solution')
        measure = abs if x.size == 1 else (lambda v: norm(v, ord=self.norm_order))
        res, rhs = residual_method(x, existing_rhs=initial_rhs, evaluate_new_rhs=False)
        self._guard(res, 'NaN or Inf detected in Simple Newton Solve: in the initial residual!')
        setups = 0
        linear_iterations = 0
        for it in range(1, self.max_nonlinear_iter + 1):
            if self.evaluate_jacobian_every_iter:
                setup_method(x)
                setups += 1
            dx, n_lin, lin_ok = solve_method(res)
            note = f'On iteration {it}, linear solve convergence: {lin_ok} in {n_lin} iterations'
            self._guard(dx, 'NaN or Inf detected in Simple Newton Solve: solution update check! ' + note)
            linear_iterations += n_lin
            x -= dx
            self._guard(x, 'NaN or Inf detected in Simple Newton Solve: solution check! ' + note)
            self._custom(x, note)
            res, rhs = residual_method(x, evaluate_new_rhs=True)
            self._guard(res, 'NaN or Inf detected in Simple Newton Solve: solution check! ' + note)
            if measure(res * self.norm_weighting) < self.tolerance:
                return SolverOutput(solution=x, rhs_at_converged=rhs, iter=it, liter=linear_iterations,
                                    converged=True, slow_convergence=it > self.slowness_detection_iter,
                                    projector_setups=setups)
        if self.must_converge:
            raise ValueError('Simple Newton method did not converge and must_converge=True!')
        return SolverOutput(solution=x, rhs_at_converged=rhs, iter=self.max_nonlinear_iter, liter=linear_iterations,
                            converged=False, slow_convergence=True, projector_setups=setups)
