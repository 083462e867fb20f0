"""
Zero-dimensional homogeneous reactors on the B200 Griffon path: `HomogeneousReactor` with the reference's constructor
and `integrate*` signatures (reactors.py:27-794): isobaric and isochoric configurations, adiabatic / isothermal /
diathermal, closed / open.

The right-hand side and the analytical Jacobian come from `gb_reactor_{rhs,jac}_{isobaric,isochoric}_*` through
spitfire_b200.griffon; the dense ns x ns linear algebra of one reactor stays with SciPy's LAPACK on the host exactly
as in the reference (reactors.py:340-356), the time integration with spitfire_b200.time.
For many reactors at once use `PyCombustionKernels.reactor_{rhs,jac}_isobaric_batch` directly (bench.py).
"""
import numpy as np
from numpy import sqrt
from scipy.linalg.lapack import dgetrf as lapack_lu_factor
from scipy.linalg.lapack import dgetrs as lapack_lu_solve

from spitfire_b200.library import Dimension, Library
from spitfire_b200.time.integrator import odesolve
from spitfire_b200.time.methods import KennedyCarpenterS6P4Q3
from spitfire_b200.time.nonlinear import SimpleNewtonSolver
from spitfire_b200.time.stepcontrol import PIController


class HomogeneousReactor(object):
    """A zero-dimensional, well-mixed reactor (mirror of reactors.py:27-794, isobaric configurations)"""

    _configurations = ['constant pressure', 'constant volume', 'isobaric', 'isochoric']
    _configuration_dict = {'constant pressure': 'isobaric', 'isobaric': 'isobaric', 'constant volume': 'isochoric',
                           'isochoric': 'isochoric'}
    _heat_transfers = ['adiabatic', 'isothermal', 'diathermal']
    _mass_transfers = ['closed', 'open']
    _shape_dict = {'cube': {'l->sov': lambda a: 6. / a, 'v->sov': lambda v: 6. / (np.power(v, 1. / 3.))},
                   'sphere': {'l->sov': lambda a: 3. / a,
                              'v->sov': lambda v: 3. / (np.power(3. * v / (4. * np.pi), 1. / 3.))},
                   'capsule': {'l->sov': lambda a: 12. / (5. * a),
                               'v->sov': lambda v: 12. / (5. * np.power(3. * v / (10. * np.pi), 1. / 3.))},
                   'tetrahedron': {'l->sov': lambda a: 6. * sqrt(6.) / a,
                                   'v->sov': lambda v: 6. * sqrt(6.) / np.power(12. * v / np.sqrt(2.), 1. / 3.)},
                   'octahedron': {'l->sov': lambda a: 3. * sqrt(6.) / a,
                                  'v->sov': lambda v: 3. * sqrt(6.) / np.power(3. * v / np.sqrt(2.), 1. / 3.)},
                   'icosahedron': {'l->sov': lambda a: 12. * sqrt(3.) / ((3. + sqrt(5.)) * a),
                                   'v->sov': lambda v: 12. * sqrt(3.) / (
                                       (3. + sqrt(5.)) * np.power(12. * v / 5. / (3. + sqrt(5.)), 1. / 3.))}}
    _shapes = list(_shape_dict.keys())

    @classmethod
    def _check(cls, argument, description, acceptable):
        if argument.lower() not in acceptable:
            raise ValueError(f'Error in reactor construction:\n    Bad {description} argument detected.\n'
                             f'    Argument given: {argument}\n    Acceptable values: {acceptable}')

    @staticmethod
    def _need(value, name, why):
        if value is None:
            raise ValueError(f'Error in reactor construction:\n    {why} but the argument "{name}" was not given.')
        return value

    def __init__(self, mech_spec, initial_mixture, configuration, heat_transfer, mass_transfer,
                 convection_temperature=None, radiation_temperature=None, convection_coefficient=None,
                 radiative_emissivity=None, shape_dimension_dict=None, mixing_tau=None, feed_temperature=None,
                 feed_mass_fractions=None, feed_density=None, rates_sensitivity_type='dense',
                 sensitivity_transform_type='exact', initial_time=0.):
        self._check(configuration, 'configuration', self._configurations)
        self._check(heat_transfer, 'heat transfer', self._heat_transfers)
        self._check(mass_transfer, 'mass transfer', self._mass_transfers)
        self._configuration = self._configuration_dict[configuration.lower()]
        self._heat_transfer = heat_transfer.lower()
        self._mass_transfer = mass_transfer.lower()

        if self._heat_transfer == 'diathermal':
            why = 'heat transfer is set to diathermal'
            self._convection_temperature = self._need(convection_temperature, 'convection_temperature', why)
            self._radiation_temperature = self._need(radiation_temperature, 'radiation_temperature', why)
            self._convection_coefficient = self._need(convection_coefficient, 'convection_coefficient', why)
            self._radiative_emissivity = self._need(radiative_emissivity, 'radiative_emissivity', why)
            sdd = self._need(shape_dimension_dict, 'shape_dimension_dict', why)
            if 'shape' not in sdd:
                raise ValueError('Error in reactor construction:\n    The shape_dimension_dict argument did not have the '
                                 'required "shape" key')
            self._check(sdd['shape'], 'shape', self._shapes)
            has_l, has_v = 'char. length' in sdd, 'volume' in sdd
            if has_l == has_v:
                raise ValueError('Error in reactor construction:\n    The shape_dimension_dict argument needs exactly one '
                                 'of the "char. length" or "volume" keys')
            conv = self._shape_dict[sdd['shape']]
            self._surface_area_to_volume = conv['l->sov'](sdd['char. length']) if has_l else conv['v->sov'](sdd['volume'])
        else:
            self._convection_temperature = self._radiation_temperature = 0.
            self._convection_coefficient = self._radiative_emissivity = 0.
            self._surface_area_to_volume = 0.
        if self._mass_transfer == 'open':
            why = 'mass transfer is set to open'
            self._mixing_tau = np.inf if mixing_tau is None else mixing_tau
            self._feed_temperature = self._need(feed_temperature, 'feed_temperature', why)
            self._feed_mass_fractions = self._need(feed_mass_fractions, 'feed_mass_fractions', why)
        else:
            self._mixing_tau, self._feed_temperature = 0., 0.
            self._feed_mass_fractions = np.zeros(1)  # never dereferenced when closed (reactors.py:226)
        if self._mass_transfer == 'open' and self._configuration == 'isochoric':
            self._feed_density = self._need(feed_density, 'feed_density', 'mass transfer is set to open')
        else:
            self._feed_density = 0. if feed_density is None else feed_density

        # parameters may be constants or functions of time (reactors.py:246-271)
        self._timevar = {a: callable(getattr(self, a)) for a in
                         ('_convection_temperature', '_radiation_temperature', '_convection_coefficient',
                          '_radiative_emissivity', '_mixing_tau', '_feed_temperature', '_feed_mass_fractions',
                          '_feed_density')}
        at0 = lambda a: getattr(self, a)(0.) if self._timevar[a] else getattr(self, a)
        self._tc_value, self._tr_value = at0('_convection_temperature'), at0('_radiation_temperature')
        self._cc_value, self._re_value = at0('_convection_coefficient'), at0('_radiative_emissivity')
        self._tau_value, self._tf_value = at0('_mixing_tau'), at0('_feed_temperature')
        self._yf_value = at0('_feed_mass_fractions')
        self._rf_value = at0('_feed_density')

        self._rates_sensitivity_option = {'dense': 0, 'no-TBAF': 1, 'sparse': 2}[rates_sensitivity_type]
        self._sensitivity_transform_option = {'exact': 0}[sensitivity_transform_type]
        self._is_open = self._mass_transfer == 'open'
        self._heat_transfer_option = {'adiabatic': 0, 'isothermal': 1, 'diathermal': 2}[self._heat_transfer]

        self._mechanism = mech_spec
        self._griffon = mech_spec.griffon
        self._initial_pressure = float(initial_mixture.P)
        self._current_pressure = float(initial_mixture.P)
        self._initial_temperature = float(initial_mixture.T)
        self._current_temperature = float(initial_mixture.T)
        self._initial_mass_fractions = np.copy(initial_mixture.Y)
        self._current_mass_fractions = np.copy(initial_mixture.Y)
        self._initial_time = np.copy(initial_time)
        self._current_time = np.copy(initial_time)
        self._n_species = mech_spec.n_species
        self._n_reactions = mech_spec.n_reactions
        if self._configuration == 'isobaric':
            self._n_equations = self._n_species  # [T, Y_0..Y_{ns-2}]
            self._temperature_index = 0
            self._initial_state = np.hstack((self._initial_temperature, self._initial_mass_fractions[:-1]))
        else:
            self._n_equations = self._n_species + 1  # [rho, T, Y_0..Y_{ns-2}] (reactors.py:313-318)
            self._temperature_index = 1
            self._initial_state = np.hstack((float(initial_mixture.density), self._initial_temperature,
                                             self._initial_mass_fractions[:-1]))
        self._current_state = np.copy(self._initial_state)
        self._variable_scales = np.ones(self._n_equations)
        self._variable_scales[self._temperature_index] = 1.e3
        self._diag_indices = np.diag_indices(self._n_equations)
        self._left_hand_side_inverse_operator = None
        self._extra_logger_title_line1 = f'{"":<10} | {"":<10}|'
        self._extra_logger_title_line2 = f' {"T (K)":<8}  | {"T-T_0 (K)":<10}|'

    # -- read-only views ----------------------------------------------------------------------------------------------
    initial_state = property(lambda self: self._initial_state)
    current_state = property(lambda self: self._current_state)
    initial_temperature = property(lambda self: self._initial_temperature)
    current_temperature = property(lambda self: self._current_temperature)
    initial_pressure = property(lambda self: self._initial_pressure)
    current_pressure = property(lambda self: self._current_pressure)
    initial_mass_fractions = property(lambda self: self._initial_mass_fractions)
    current_mass_fractions = property(lambda self: self._current_mass_fractions)
    initial_time = property(lambda self: self._initial_time)
    current_time = property(lambda self: self._current_time)
    n_species = property(lambda self: self._n_species)
    n_reactions = property(lambda self: self._n_reactions)

    @classmethod
    def get_supported_reactor_shapes(cls):
        return HomogeneousReactor._shape_dict.keys()

    # -- callables handed to the time integrator ------------------------------------------------------------------------
    def _update_parameters(self, t):
        v = self._timevar
        if v['_convection_temperature']:
            self._tc_value = self._convection_temperature(t)
        if v['_radiation_temperature']:
            self._tr_value = self._radiation_temperature(t)
        if v['_convection_coefficient']:
            self._cc_value = self._convection_coefficient(t)
        if v['_radiative_emissivity']:
            self._re_value = self._radiative_emissivity(t)
        if v['_mixing_tau']:
            self._tau_value = self._mixing_tau(t)
        if v['_feed_temperature']:
            self._tf_value = self._feed_temperature(t)
        if v['_feed_mass_fractions']:
            self._yf_value = self._feed_mass_fractions(t)
        if v['_feed_density']:
            self._rf_value = self._feed_density(t)

    def _rhs(self, t, state):
        k = np.zeros(self._n_equations)
        self._update_parameters(t)
        if self._configuration == 'isochoric':
            self._griffon.reactor_rhs_isochoric(np.ascontiguousarray(state), self._rf_value, self._tf_value,
                                                np.ascontiguousarray(self._yf_value, dtype=np.float64),
                                                self._tau_value, self._tc_value, self._tr_value, self._cc_value,
                                                self._re_value, self._surface_area_to_volume,
                                                self._heat_transfer_option, self._is_open, k)
            return k
        self._griffon.reactor_rhs_isobaric(np.ascontiguousarray(state), self._initial_pressure, self._tf_value,
                                           np.ascontiguousarray(self._yf_value, dtype=np.float64), self._tau_value,
                                           self._tc_value, self._tr_value, self._cc_value, self._re_value,
                                           self._surface_area_to_volume, self._heat_transfer_option, self._is_open, k)
        return k

    def _jac(self, state):
        k = np.zeros(self._n_equations)
        j = np.zeros(self._n_equations * self._n_equations)
        if self._configuration == 'isochoric':
            self._griffon.reactor_jac_isochoric(np.ascontiguousarray(state), self._rf_value, self._tf_value,
                                                np.ascontiguousarray(self._yf_value, dtype=np.float64),
                                                self._tau_value, self._tc_value, self._tr_value, self._cc_value,
                                                self._re_value, self._surface_area_to_volume,
                                                self._heat_transfer_option, self._is_open,
                                                self._rates_sensitivity_option, k, j)
            return j.reshape((self._n_equations, self._n_equations), order='F')
        self._griffon.reactor_jac_isobaric(np.ascontiguousarray(state), self._initial_pressure, self._tf_value,
                                           np.ascontiguousarray(self._yf_value, dtype=np.float64), self._tau_value,
                                           self._tc_value, self._tr_value, self._cc_value, self._re_value,
                                           self._surface_area_to_volume, self._heat_transfer_option, self._is_open,
                                           self._rates_sensitivity_option, self._sensitivity_transform_option, k, j)
        return j.reshape((self._n_equations, self._n_equations), order='F')

    def _lapack_setup(self, t, state, prefactor):
        j = self._jac(state) * prefactor
        j[self._diag_indices] -= 1.
        self._left_hand_side_inverse_operator = lapack_lu_factor(j)[:2]

    def _lapack_solve(self, residual):
        lu, piv = self._left_hand_side_inverse_operator
        return lapack_lu_solve(lu, piv, residual)[0], 1, True

    def _extra_logger_log(self, state, *args, **kwargs):
        T = state[self._temperature_index]
        return f'{T:>10.2f} | {T - self._initial_temperature:>10.2f}|'

    # -- integration (reactors.py:429-794) ---------------------------------------------------------------------------------
    def integrate(self, stop_at_time=None, stop_at_steady=None, stop_criteria=None, first_time_step=1.e-6,
                  max_time_step=1.e6, minimum_time_step_count=40, transient_tolerance=1.e-10, write_log=False,
                  log_rate=100, maximum_steps_per_jacobian=1, nonlinear_solve_tolerance=1.e-12, linear_solver='lapack',
                  plot=None, stepper_type=KennedyCarpenterS6P4Q3, nlsolver_type=SimpleNewtonSolver,
                  stepcontrol_type=PIController, extra_integrator_args=dict(), extra_stepper_args=dict(),
                  extra_nlsolver_args=dict(), extra_stepcontrol_args=dict(), save_first_and_last_only=False):
        """Base method for reactor integration; same arguments as the reference (plotting is not provided).
        Returns a library over time with temperature, pressure and mass fractions."""
        if linear_solver != 'lapack':
            raise ValueError('only the "lapack" linear solver is provided')

        def post_step_callback(t, state, *args):
            state[state < 0.] = 0.
            return state

        iargs = {'stop_criteria': stop_criteria}
        if stop_at_time is not None:
            iargs['stop_at_time'] = stop_at_time
        if stop_at_steady is not None:
            iargs['stop_at_steady'] = stop_at_steady
        iargs.update(extra_integrator_args)
        sc = {'first_step': first_time_step, 'max_step': max_time_step, 'target_error': transient_tolerance}
        sc.update(extra_stepcontrol_args)
        nl = {'evaluate_jacobian_every_iter': False, 'norm_weighting': 1. / self._variable_scales,
              'tolerance': nonlinear_solve_tolerance}
        nl.update(extra_nlsolver_args)
        st = {'nonlinear_solver': nlsolver_type(**nl), 'norm_weighting': 1. / self._variable_scales}
        st.update(extra_stepper_args)
        output = odesolve(right_hand_side=self._rhs, initial_state=self._current_state, initial_time=self._current_time,
                          step_size=stepcontrol_type(**sc), method=stepper_type(**st), linear_setup=self._lapack_setup,
                          linear_solve=self._lapack_solve, minimum_time_step_count=minimum_time_step_count,
                          linear_setup_rate=maximum_steps_per_jacobian, verbose=write_log, log_rate=log_rate,
                          extra_logger_log=self._extra_logger_log,
                          extra_logger_title_line1=self._extra_logger_title_line1,
                          extra_logger_title_line2=self._extra_logger_title_line2,
                          norm_weighting=1. / self._variable_scales, post_step_callback=post_step_callback,
                          save_each_step=not save_first_and_last_only, **iargs)
        if save_first_and_last_only:
            state, time, _ = output
            states, t = np.array(state).reshape(1, -1), np.array([time], dtype=np.float64).ravel()
        else:
            t, states = output
        self._current_state = np.copy(states[-1, :])
        self._current_time = np.copy(t[-1])
        ti = self._temperature_index
        self._current_temperature = float(states[-1, ti])
        lib = Library(Dimension('time', t))
        if self._configuration == 'isobaric':
            lib['temperature'] = np.array(states[:, 0])
            lib['pressure'] = self._initial_pressure + np.zeros_like(lib['temperature'])
        else:  # reactors.py:637-645
            lib['density'] = np.array(states[:, 0])
            lib['temperature'] = np.array(states[:, 1])
        names = self._mechanism.species_names
        last = np.ones_like(lib['temperature'])
        lib['mass fraction ' + names[-1]] = last
        for i, s in enumerate(names[:-1]):
            lib['mass fraction ' + s] = np.array(states[:, ti + 1 + i])
            last = last - states[:, ti + 1 + i]
        lib['mass fraction ' + names[-1]] = last
        self._current_mass_fractions = np.array([lib['mass fraction ' + s][-1] for s in names])
        if self._configuration == 'isochoric':
            mw = np.asarray(self._mechanism.molecular_weights, dtype=np.float64)
            self._current_pressure = float(states[-1, 0] * states[-1, 1] / np.sum(self._current_mass_fractions / mw))
        lib.extra_attributes['mech_spec'] = self._mechanism
        return lib

    def integrate_to_steady(self, steady_tolerance=1.e-6, **kwargs):
        return self.integrate(stop_at_steady=steady_tolerance, **kwargs)

    def integrate_to_time(self, final_time, **kwargs):
        return self.integrate(stop_at_time=final_time, **kwargs)

    def _has_ignited(self, state, delta_temperature_ignition):
        return state[self._temperature_index] - self._initial_temperature > delta_temperature_ignition

    def integrate_to_steady_after_ignition(self, steady_tolerance=1.e-6, delta_temperature_ignition=400., **kwargs):
        def stop(t, state, residual, *args, **kw):
            return self._has_ignited(state, delta_temperature_ignition) and residual < steady_tolerance

        return self.integrate(stop_criteria=stop, **kwargs)

    def compute_ignition_delay(self, delta_temperature_ignition=400., minimum_allowable_residual=1.e-12,
                               return_solution=False, **kwargs):
        """time at which the temperature has risen by delta_temperature_ignition (reactors.py:724-779)"""

        def stop(t, state, residual, *args, **kw):
            if residual > minimum_allowable_residual:
                return self._has_ignited(state, delta_temperature_ignition)
            raise ValueError(f'From compute_ignition_delay(): residual < minimum allowable value '
                             f'({minimum_allowable_residual}), suggesting that the reactor will not ignite.')

        lib = self.integrate(stop_criteria=stop, save_first_and_last_only=not return_solution, **kwargs)
        tau = lib.time_values[-1]
        return (tau, lib) if return_solution else tau


# ---- many reactors at once ---------------------------------------------------------------------------------------------
class _ReactorBatchOps(object):
    """What time.batched.integrate_batch needs, for N independent isobaric reactors that share the pressure and the
    reactor parameters and differ in their states: right-hand side and Jacobian through the batched C-ABI entry points
    (`gb_reactor_{rhs,jac}_isobaric_batch` -- the headline kernels), the dense ns x ns Newton matrices handled as
    block-tridiagonal systems of ONE block (`gb_btddod_full_*_batch` with num_blocks = 1: dgetrf / dgetrs per reactor,
    which is what the reference does through SciPy, reactors.py:20-21, 571-587). With the CPU oracle injected
    (`griffon_factory`) the same interface loops over the members on the host."""

    def __init__(self, reactor, n):
        import torch
        from . import griffon as gmod
        self.torch, self.gmod = torch, gmod
        self.r = reactor
        self.g = reactor._griffon
        self.ns, self.nzi = reactor._n_equations, 1
        self.ndof = self.ns
        self.nelem = self.ns * self.ns
        self.F = n
        self.on_device = isinstance(self.g, gmod.PyCombustionKernels)
        if self.on_device and not torch.cuda.is_available():
            raise gmod.GriffonB200Error('the batched reactor integrator runs on the GPU: no CUDA device')
        self.device = torch.device('cuda') if self.on_device else torch.device('cpu')
        self.scales = torch.as_tensor(np.tile(reactor._variable_scales, (n, 1))).to(self.device)
        r = reactor
        self._args = (r._initial_pressure,)
        self._kw = dict(T_in=r._tf_value, y_in=np.ascontiguousarray(r._yf_value, dtype=np.float64), tau=r._tau_value,
                        T_inf=r._tc_value, T_surf=r._tr_value, h_conv=r._cc_value, eps_rad=r._re_value,
                        SoV=r._surface_area_to_volume, heat_option=r._heat_transfer_option, open_=r._is_open)
        if self.on_device and r._is_open:
            self._kw['y_in'] = torch.as_tensor(self._kw['y_in']).to(self.device)
        self._host = (r._initial_pressure, r._tf_value, np.ascontiguousarray(r._yf_value, dtype=np.float64), r._tau_value,
                      r._tc_value, r._tr_value, r._cc_value, r._re_value, r._surface_area_to_volume,
                      r._heat_transfer_option, r._is_open)

    def rhs(self, q, idx=None, key=None):
        torch = self.torch
        out = torch.empty_like(q)
        iso = self.r._configuration == 'isochoric'
        if self.on_device:
            if iso:
                self.g.reactor_rhs_isochoric_batch(q.contiguous(), out, rho_in=self.r._rf_value, **self._kw)
            else:
                self.g.reactor_rhs_isobaric_batch(q.contiguous(), self._args[0], out, **self._kw)
        else:
            qn, on = q.numpy(), out.numpy()
            for k in range(q.shape[0]):
                if iso:
                    self.g.reactor_rhs_isochoric(np.ascontiguousarray(qn[k]), self.r._rf_value, *self._host[1:], on[k])
                else:
                    self.g.reactor_rhs_isobaric(np.ascontiguousarray(qn[k]), *self._host, on[k])
        return out

    def jac(self, q, idx=None, key=None):
        torch = self.torch
        n = q.shape[0]
        rhs = torch.empty_like(q)
        iso = self.r._configuration == 'isochoric'
        if self.on_device:
            J = torch.empty((n, self.nelem), dtype=torch.float64, device=self.device)
            if iso:
                self.g.reactor_jac_isochoric_batch(q.contiguous(), rhs, J, rho_in=self.r._rf_value,
                                                   rates_sens_option=self.r._rates_sensitivity_option, **self._kw)
            else:
                self.g.reactor_jac_isobaric_batch(q.contiguous(), self._args[0], rhs, J,
                                                  rates_sens_option=self.r._rates_sensitivity_option,
                                                  sens_transform_option=self.r._sensitivity_transform_option,
                                                  **self._kw)
        else:
            J = torch.zeros((n, self.nelem), dtype=torch.float64)
            qn, rn, Jn = q.numpy(), rhs.numpy(), J.numpy()
            for k in range(n):
                if iso:
                    self.g.reactor_jac_isochoric(np.ascontiguousarray(qn[k]), self.r._rf_value, *self._host[1:],
                                                 self.r._rates_sensitivity_option, rn[k], Jn[k])
                else:
                    self.g.reactor_jac_isobaric(np.ascontiguousarray(qn[k]), *self._host,
                                                self.r._rates_sensitivity_option,
                                                self.r._sensitivity_transform_option, rn[k], Jn[k])
        return J  # column-major n_equations x n_equations per reactor = one BTDDOD block


def _borrow_linear_algebra():
    from .flamelet import _BatchOps
    for name in ('factorize', 'solve', 'add_to_block_diagonal', 'factor_store', 'factorize_into'):
        setattr(_ReactorBatchOps, name, getattr(_BatchOps, name))
    _ReactorBatchOps.explicit_inverse_solves = _BatchOps.explicit_inverse_solves
    _ReactorBatchOps._rows32 = (None, None)


class HomogeneousReactorBatch(object):
    """N reactors (isobaric or isochoric) integrated together on the GPU ("many ignition
    reactors at once"). Isochoric members start at the template's pressure: rho_i = p M_i / (R T_i).

    Same model and integrator as `HomogeneousReactor` (ESDIRK64, PI step control, Newton with the Jacobian refreshed
    every `maximum_steps_per_jacobian` steps, negative mass fractions clipped after each step); every member keeps its
    own time, step size and error history (time/batched.py), so a member's trajectory is the one the serial class
    produces for it. All members share the mechanism, the pressure and the (constant) reactor parameters of `template`
    and start from `temperatures[i]`, `mass_fractions[i]`.

        batch = HomogeneousReactorBatch(HomogeneousReactor(mech, mix, 'isobaric', 'adiabatic', 'closed'), T0s, Y0s)
        tau = batch.compute_ignition_delay()
    """

    def __init__(self, template, temperatures, mass_fractions):
        if any(template._timevar.values()):
            raise ValueError('HomogeneousReactorBatch needs constant reactor parameters')
        self._r = template
        T = np.atleast_1d(np.asarray(temperatures, dtype=np.float64))
        Y = np.atleast_2d(np.asarray(mass_fractions, dtype=np.float64))
        if Y.shape != (T.size, template._n_species):
            raise ValueError('mass_fractions must be [n_reactors, n_species]')
        self._ti = template._temperature_index
        if template._configuration == 'isochoric':
            m = template._mechanism
            mw = np.asarray(m.molecular_weights, dtype=np.float64)
            rho = template._initial_pressure / (m.gas_constant * T * np.sum(Y / mw[None, :], axis=1))
            self._initial_states = np.ascontiguousarray(np.hstack((rho[:, None], T[:, None], Y[:, :-1])))
        else:
            self._initial_states = np.ascontiguousarray(np.hstack((T[:, None], Y[:, :-1])))
        if not hasattr(_ReactorBatchOps, 'factorize'):
            _borrow_linear_algebra()
        self.ops = _ReactorBatchOps(template, T.size)

    n_reactors = property(lambda self: self._initial_states.shape[0])
    initial_states = property(lambda self: self._initial_states)

    def integrate(self, stop, first_time_step=1.e-6, max_time_step=1.e6, minimum_time_step_count=40,
                  transient_tolerance=1.e-10, maximum_steps_per_jacobian=1, nonlinear_solve_tolerance=1.e-12,
                  save_each_step=False, maximum_steps=100000, stop_ignores_minimum=False, stop_at_time=None):
        """stop(t, states, residual, nsteps) -> bool tensor, all arguments tensors over the members.
        Returns (times, states, failed): per member arrays of the saved times / states (first and last only unless
        save_each_step)."""
        from .time.batched import integrate_batch
        ops = self.ops
        q0 = ops.torch.as_tensor(self._initial_states).to(ops.device)
        return integrate_batch(ops, q0, stop, first_time_step=first_time_step, max_time_step=max_time_step,
                               minimum_time_step_count=minimum_time_step_count,
                               transient_tolerance=transient_tolerance,
                               maximum_steps_per_jacobian=maximum_steps_per_jacobian,
                               nonlinear_solve_tolerance=nonlinear_solve_tolerance, save_each_step=save_each_step,
                               maximum_steps=maximum_steps, stop_ignores_minimum=stop_ignores_minimum,
                               stop_at_time=stop_at_time)

    def integrate_to_steady(self, steady_tolerance=1.e-6, **kwargs):
        return self.integrate(lambda t, q, residual, nsteps: residual < steady_tolerance, stop_ignores_minimum=True,
                              **kwargs)

    def integrate_to_time(self, final_time, **kwargs):
        """every member lands on final_time exactly: the step that would cross it is shortened, as odesolve's
        stop_at_time does (integrator.py:590-593)"""
        return self.integrate(lambda t, q, residual, nsteps: t >= final_time, stop_at_time=float(final_time), **kwargs)

    def compute_ignition_delay(self, delta_temperature_ignition=400., minimum_allowable_residual=1.e-12, **kwargs):
        """time at which each member's temperature has risen by delta_temperature_ignition (reactors.py:724-779); NaN for
        a member whose residual falls below minimum_allowable_residual first (the serial class raises for it).
        Under torchrun (spitfire_b200.parallel initialised) the members are dealt block-cyclically to the ranks, every
        rank integrates its share on its own GPU and the delays are exchanged once at the end (no data-path
        collective: the reactors are independent)."""
        from . import parallel
        if parallel.world_size() > 1 and not getattr(self, '_is_share', False):
            mine = parallel.my_share(list(range(self.n_reactors)))
            local = dict()
            if mine:
                ti = self._ti
                Y = np.hstack((self._initial_states[mine, ti + 1:],
                               1. - self._initial_states[mine, ti + 1:].sum(axis=1, keepdims=True)))
                share = HomogeneousReactorBatch(self._r, self._initial_states[mine, ti], Y)
                share._initial_states = np.ascontiguousarray(self._initial_states[mine])  # (exactly the caller's states)
                share._is_share = True
                tau = share.compute_ignition_delay(delta_temperature_ignition, minimum_allowable_residual, **kwargs)
                local = {int(k): float(t) for k, t in zip(mine, tau)}
            merged = parallel.gather_dicts(local)
            return np.array([merged[k] for k in range(self.n_reactors)])
        ti = self._ti
        T0 = self.ops.torch.as_tensor(self._initial_states[:, ti]).to(self.ops.device)

        def stop(t, q, residual, nsteps):
            return ((q[:, ti] - T0) > delta_temperature_ignition) | (residual <= minimum_allowable_residual)

        times, states, failed = self.integrate(stop, **kwargs)
        tau = np.array([th[-1] for th in times])
        ignited = np.array([qs[-1][ti] for qs in states]) - self._initial_states[:, ti] > delta_temperature_ignition
        return np.where(ignited & ~np.asarray(failed), tau, np.nan)
