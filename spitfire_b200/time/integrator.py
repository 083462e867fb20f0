"""
`odesolve`: the time-integration governor (mirror of src/spitfire/time/integrator.py:229-713).

It owns the step loop around a stepper (`method.single_step`), a step-size controller and the linear-projector policy
(when to re-evaluate and re-factorize the Jacobian). The decision sequence of the reference is kept exactly -- order of
the dt coercions, of the accept/reject tests, of the controller call and of the projector-refresh tests
(integrator.py:520-640) -- because adaptive step sequences and Jacobian ages are part of the behaviour the gold
trajectories pin (e.g. 517 steps for the H2 ignition case).
"""
import datetime
import logging
import time as timer

import numpy as np
from numpy import inf
from scipy.linalg import norm
from scipy.linalg.lapack import dgetrf as lapack_lu_factor
from scipy.linalg.lapack import dgetrs as lapack_lu_solve

from spitfire_b200.time.methods import KennedyCarpenterS6P4Q3
from spitfire_b200.time.nonlinear import SimpleNewtonSolver, finite_difference_jacobian
from spitfire_b200.time.stepcontrol import ConstantTimeStep, PIController


class FailedODESolveException(ValueError):
    """raised when the integration dies; carries the partial history (integrator.py:214-226)"""

    def __init__(self, msg, times, states):
        super().__init__(msg)
        self._times, self._states = times, states

    times = property(lambda self: self._times)
    states = property(lambda self: self._states)


class _Logger(object):
    """the verbose table of integrator.py:133-211, one line every `log_rate` accepted steps"""

    def __init__(self, verbose, in_situ, rate, lines_per_header, title1, title2, extra):
        self.verbose, self.in_situ, self.rate, self.lines_per_header = verbose, in_situ, rate, lines_per_header
        self.title1, self.title2, self.extra = title1 or '', title2 or '', extra
        self.count = 0
        self.title_count = 0

    def __call__(self, state, t, dt, residual, nsteps, nnl, nlin, nsetup, cpu):
        self.count += 1
        if not self.verbose:
            return
        cols = [('number of', 'time steps', f' {nsteps:<9}'), ('simulation', 'time (s)', f'{float(t):<10.2e}'),
                ('time step', 'size (s)', f'{float(dt):<10.2e}')]
        if self.in_situ:
            if nnl == 'n/a':
                cols += [('nlin. iter', 'per step', f'{"n/a":<10}'), ('lin. iter', 'per nlin.', f'{"n/a":<10}'),
                         ('steps', 'per Jac.', f'{"n/a":<10}')]
            else:
                cols += [('nlin. iter', 'per step', f'{nnl / nsteps:<10.2f}'),
                         ('lin. iter', 'per nlin.', f'{nlin / nnl:<10.2f}'),
                         ('steps', 'per Jac.', f'{nsteps / nsetup:<10.2f}')]
        cols += [('diff. eqn.', '|residual|', f'{residual:<10.2e}'), ('total cpu', 'time (s)', f'{cpu:<10.2e}'),
                 ('cput per', 'step (ms)', f'{1.e3 * cpu / float(nsteps):<10.2e}')]
        line1 = '|' + ' | '.join(f'{a:<10}' for a, _, _ in cols)[:-1] + '|' + self.title1
        line2 = '|' + ' | '.join(f'{b:<10}' for _, b, _ in cols)[:-1] + '|' + self.title2
        body = '|' + ' | '.join(c for _, _, c in cols)[:-1] + '|'
        dashes = '-' * (len(line2) - 1)
        header = '\n' + line1 + '\n' + line2 + '\n' + dashes + '|'
        if self.extra is not None:
            body += self.extra(state, t, nsteps, nnl, nlin)
        if nsteps == 1:
            print(header)
        if self.count == self.rate:
            self.title_count += 1
            self.count = 0
            if self.title_count == self.lines_per_header:
                print(dashes)
                print(header)
                self.title_count = 0
            print(body, flush=True)


def _update_is_acceptable(state, dstate, time_error, target_error, nl_converged, strict, nl_must_converge, custom):
    """integrator.py:53-79: False means the step is rejected"""
    if time_error > target_error and strict:
        return False
    if not np.all(np.isfinite(dstate)):
        return False
    if nl_converged is not None and (not nl_converged and nl_must_converge):
        return False
    if custom is not None:
        return custom(state, dstate, time_error, nl_converged)
    return True


def _projector_policy(nl_ok, nl_slow, dt, dt_recent, count, rate, slow_factor, fail_factor, grow_limit, shrink_limit):
    """integrator.py:100-130 -> (re-evaluate?, dt, count)"""
    if count == rate:
        return True, dt, 0
    if nl_slow:
        return True, dt * slow_factor, count
    if not nl_ok:
        return True, dt * fail_factor, count
    if dt > dt_recent * grow_limit or dt < dt_recent * shrink_limit:
        return True, dt, count
    return False, dt, count


def odesolve(right_hand_side,
             initial_state,
             output_times=None,
             save_each_step=False,
             initial_time=0.,
             stop_criteria=None,
             stop_at_time=None,
             stop_at_steady=None,
             minimum_time_step_count=0,
             maximum_time_step_count=inf,
             pre_step_callback=None,
             post_step_callback=None,
             step_update_callback=None,
             method=None,
             step_size=None,
             linear_setup=None,
             linear_solve=None,
             linear_setup_rate=1,
             mass_setup=None,
             mass_matvec=None,
             verbose=False,
             debug_verbose=False,
             log_rate=1,
             log_lines_per_header=10,
             extra_logger_title_line1=None,
             extra_logger_title_line2=None,
             extra_logger_log=None,
             norm_weighting=1.,
             strict_temporal_error_control=False,
             nonlinear_solve_must_converge=False,
             warn_on_failed_step=False,
             return_on_failed_step=False,
             time_step_reduction_factor_on_failure=0.8,
             time_step_reduction_factor_on_slow_solve=0.8,
             time_step_increase_factor_to_force_jacobian=1.05,
             time_step_decrease_factor_to_force_jacobian=0.9,
             show_solver_stats_in_situ=False,
             return_info=False,
             throw_on_failure=True,
             print_exception_on_failure=True,
             maximum_residual=None):
    """Solve q' = f(t, q) -- same arguments, defaults and return conventions as the reference's `odesolve`
    (integrator.py:229-387): returns the states at `output_times`, or `(t, q)` histories with `save_each_step`, or
    `(q_final, t_final, dt_final)`; `return_info=True` appends the statistics dictionary."""
    if method is None:
        method = KennedyCarpenterS6P4Q3(SimpleNewtonSolver())
    if step_size is None:
        step_size = PIController()

    # ---- argument checks (integrator.py:389-436) -----------------------------------------------------------------
    if stop_at_steady is not None and not isinstance(stop_at_steady, (bool, float)):
        raise TypeError('Error in Spitfire odesolve, the stop_at_time argument must be provided as '
                        'a boolean (default tolerance if True) or a float.')
    if not isinstance(save_each_step, (bool, int)) or (isinstance(save_each_step, int) and save_each_step < 0):
        raise ValueError('Error in Spitfire odesolve, the save_each_step argument must be either True/False '
                         'or a positive integer (the step frequency at which data is saved).')
    if output_times is not None:
        for other, label in ((stop_at_time, 'stop_at_time'), (stop_at_steady, 'stop_at_steady'),
                             (stop_criteria, 'stop_criteria')):
            if other is not None:
                raise ValueError(f'Error in Spitfire odesolve, the {label} argument may not be provided if the '
                                 f'output_times argument is also in use.')
        if save_each_step:
            raise ValueError('Error in Spitfire odesolve, the save_each_step argument may not be provided if the '
                             'output_times argument is also in use.')
        if np.min(output_times) < initial_time:
            raise ValueError('Error in Spitfire odesolve, the provided output_times must be greater than or equal to'
                             ' the initial_time (defaults to 0.)')
        listed = output_times.tolist()
        if len(set(listed)) != len(listed):
            raise ValueError('Error in Spitfire odesolve, the provided output_times must be unique.')
        if listed != sorted(listed):
            raise ValueError('Error in Spitfire odesolve, the provided output_times must be increasing.')
    if output_times is None and stop_at_time is None and stop_at_steady is None and stop_criteria is None:
        raise ValueError('Error in Spitfire odesolve, you have not specified enough information to stop a simulation, '
                         'you must provide output_times, stop_at_time=tfinal, stop_at_steady=[True or tolerance], '
                         'or stop_criteria as a function(t, state, residual, nsteps)')
    if isinstance(step_size, PIController) and not method.is_adaptive:
        raise TypeError(f'The method provided {method.name} cannot be used with a PI controller'
                        ' (the default step_size argument), you must set step_size equal to a constant value.')

    log = logging.getLogger(__name__)
    t_hist, q_hist, out_states = None, None, None
    q, t = np.copy(initial_state), np.copy(initial_time)
    dt = None
    try:
        coerce_final = stop_at_time is not None
        out_idx = 0
        if output_times is not None:
            coerce_final = True
            stop_at_time = output_times[-1]
            out_states = np.zeros((output_times.size, initial_state.size))
            if output_times[0] < 1e-14:
                out_states[0, :] = np.copy(initial_state)
                out_idx = 1
        steady_tol = None
        if stop_at_steady is not None:
            steady_tol = 1.e-4 if isinstance(stop_at_steady, bool) else stop_at_steady

        fd_jacobian = method.is_implicit and linear_setup is None
        setup_in_governor = (method.is_implicit and method.nonlinear_solver.setup_projector_in_governor) or fd_jacobian
        refresh = True
        setup_count = 0
        if isinstance(step_size, float):
            step_size = ConstantTimeStep(step_size)
        if verbose:
            print('\n', datetime.datetime.now().strftime('%Y-%m-%d %H:%M'),
                  ': Spitfire running case with method:', method.name, flush=True)
        logger = _Logger(verbose, show_solver_stats_in_situ, log_rate, log_lines_per_header, extra_logger_title_line1,
                         extra_logger_title_line2, extra_logger_log)
        if save_each_step:
            t_hist, q_hist = [np.copy(t)], [np.copy(q)]

        nsteps = 0
        dt = step_size.first_step_size()
        dt_min, dt_max = min(1.e305, dt), max(0., dt)
        n_nonlinear, n_linear, n_setups = 0, 0, 0

        if fd_jacobian:
            # dense finite-difference projector factorized with LAPACK (integrator.py:490-511)
            diag = np.diag_indices(initial_state.size)
            holder = dict()

            def linear_setup(t_, q_, prefactor):
                f_of = lambda x: right_hand_side(t, x)
                jac = finite_difference_jacobian(f_of, f_of(q_), q_) * prefactor
                jac[diag] -= 1.
                holder['lu'] = lapack_lu_factor(jac)[:2]

            def linear_solve(res):
                return lapack_lu_solve(holder['lu'][0], holder['lu'][1], res)[0], 1, True

        gamma = method.implicit_coefficient if method.is_implicit else None
        if mass_matvec is None:
            mass_matvec = lambda x, *args: x
        keep_going = True
        residual = None
        clock0 = timer.perf_counter()
        while keep_going:
            if output_times is not None and t + dt > output_times[out_idx]:
                dt = output_times[out_idx] - t
            if coerce_final and t + dt > stop_at_time:
                dt = stop_at_time - t
            if pre_step_callback is not None:
                pre_step_callback(t, q, nsteps)
            if method.is_implicit and setup_in_governor and refresh:
                linear_setup(t, q, dt * gamma)
                n_setups += 1
            setup_count += 1

            dt_now = dt
            out = method.single_step(q, t, dt, right_hand_side,
                                     (lambda t_, x_: linear_setup(t_, x_, dt_now * gamma)) if method.is_implicit
                                     else None,
                                     linear_solve, mass_setup, mass_matvec)
            dq, err = out.solution_update, out.temporal_error
            nl_ok, nl_slow = out.nonlinear_converged, out.slow_nonlinear_convergence
            n_setups += out.projector_setups if out.projector_setups is not None else 0
            residual = norm(dq * norm_weighting, ord=np.inf) / dt
            if maximum_residual is not None and residual > maximum_residual:
                raise ValueError(f'residual of {residual:.2e} exceeded maximum_residual of {maximum_residual:.2e}.')

            dt_recent = dt
            if _update_is_acceptable(q, dq, err, step_size.target_error(), nl_ok, strict_temporal_error_control,
                                     nonlinear_solve_must_converge, step_update_callback):
                q += dq
                t += dt
                nsteps += 1
                n_nonlinear = n_nonlinear + out.nonlinear_iter if out.nonlinear_iter is not None else 'n/a'
                n_linear = n_linear + out.linear_iter if out.linear_iter is not None else 'n/a'
                if post_step_callback is not None:
                    modified = post_step_callback(t, q, residual, nsteps)
                    q = q if modified is None else np.copy(modified)
                if output_times is not None and t >= output_times[out_idx]:
                    out_states[out_idx, :] = np.copy(q)
                    out_idx += 1
                if save_each_step and (isinstance(save_each_step, bool) or not (nsteps % save_each_step)):
                    t_hist.append(np.copy(t))
                    q_hist.append(np.copy(q))
                logger(q, t, dt, residual, nsteps, n_nonlinear, n_linear, n_setups, timer.perf_counter() - clock0)
                dt = step_size(nsteps, dt, out)
                if method.is_implicit:
                    refresh, dt, setup_count = _projector_policy(
                        nl_ok, nl_slow, dt, dt_recent, setup_count, linear_setup_rate,
                        time_step_reduction_factor_on_slow_solve, time_step_reduction_factor_on_failure,
                        time_step_increase_factor_to_force_jacobian, time_step_decrease_factor_to_force_jacobian)
            else:
                if return_on_failed_step:
                    raise ValueError('Step failed and return_on_failed_step=True, stopping!')
                dt *= time_step_reduction_factor_on_failure
                refresh = True
                if warn_on_failed_step:
                    print('Warning! Step failed! return_on_failed_step=False so continuing on... retrying step...')

            dt_min, dt_max = min(dt_min, dt), max(dt_max, dt)
            if stop_criteria is not None:
                keep_going = not stop_criteria(t, q, residual, nsteps)
            if nsteps < minimum_time_step_count:
                keep_going = True
            if nsteps > maximum_time_step_count:
                keep_going = False
            elif coerce_final and t >= stop_at_time:
                keep_going = False
            elif steady_tol is not None and residual < steady_tol:
                keep_going = False

        runtime = timer.perf_counter() - clock0
        if verbose:
            print('\nIntegration successfully completed!\n\nStatistics:')
            print('- number of time steps :', nsteps)
            print('- final simulation time:', t)
            print('- smallest time step   :', dt_min)
            print('- average time step    :', t / nsteps)
            print('- largest time step    :', dt_max)
            print('\n  CPU time')
            print('- total    (s) : {:.6e}'.format(runtime))
            print('- per step (ms): {:.6e}'.format(1.e3 * runtime / nsteps))
            if method.is_implicit:
                print('\n  Nonlinear iterations')
                print('- total   : {:}'.format(n_nonlinear))
                print('- per step: {:.1f}'.format(n_nonlinear / nsteps))
                print('\n  Linear iterations')
                print('- total     : {:}'.format(n_linear))
                print('- per step  : {:.1f}'.format(n_linear / nsteps))
                print('- per nliter: {:.1f}'.format(n_linear / n_nonlinear))
                print('\n  Jacobian setups')
                print('- total     : {:}'.format(n_setups))
                print('- steps per : {:.1f}'.format(nsteps / n_setups))
                print('- nliter per: {:.1f}'.format(n_nonlinear / n_setups))
                print('- liter per : {:.1f}'.format(n_linear / n_setups))
            print('\n', datetime.datetime.now().strftime('%Y-%m-%d %H:%M'),
                  ': Spitfire finished in {:.8e} seconds!\n'.format(runtime), flush=True)
        stats = {'success': True, 'time steps': nsteps, 'simulation time': t, 'total cpu time (s)': runtime}
        if method.is_implicit:
            stats.update({'nonlinear iter': n_nonlinear, 'linear iter': n_linear, 'Jacobian setups': n_setups})

    except Exception as error:
        stats = {'success': False}
        if print_exception_on_failure:
            print('Spitfire odesolve caught the following Exception during time integration:\n')
            log.exception(error)
        if throw_on_failure:
            msg = 'odesolve failed to integrate the system due to an Exception being caught:\n' + str(error) + '\n'
            if output_times is not None:
                raise FailedODESolveException(msg=msg, times=output_times, states=out_states)
            if save_each_step and t_hist is not None:
                raise FailedODESolveException(msg=msg, times=np.array(t_hist), states=np.array(q_hist))
            raise FailedODESolveException(msg=msg, times=np.array([t]), states=np.array([q]))

    if output_times is not None:
        return (out_states, stats) if return_info else out_states
    if save_each_step:
        return (np.array(t_hist), np.array(q_hist), stats) if return_info else (np.array(t_hist), np.array(q_hist))
    return (q, t, dt, stats) if return_info else (q, t, dt)
