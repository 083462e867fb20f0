"""
Nonpremixed flamelets on the B200 Griffon path: `FlameletSpec`, `Flamelet` (the reference's API, flamelet.py:29-1778)
and `FlameletBatch` (new: F flamelets advanced together, device resident).

The reference's `Flamelet` calls Griffon once per flamelet and iteration (flamelet_rhs / flamelet_jacobian /
py_btddod_*). Here the solver loops are written once, for a batch of F flamelets with per-member control flow
(active masks, own line search, own Jacobian-refresh flag, own pseudo-time steps), on torch tensors that live where
the kernels run: on the GPU for the product path (`gb_flamelet_{rhs,jacobian}_batch`, `gb_btddod_*_batch` through
spitfire_b200.griffon), on the host when a test injects the CPU oracle as the mechanism's Griffon object. A
`Flamelet` is a batch of one, so its iterates follow the reference's arithmetic member by member; the library
builders (tabulation.py) use wider batches.

Deviations from the reference, all on cold or fallback paths:
  * initial conditions 'equilibrium' / 'unreacted' / 'Burke-Schumann' use spitfire_b200.streams / equilibrium instead of
    Cantera (flamelet.py:525-586);
  * the explosive eigenvalues of the pseudo-transient solver (LAPACK dgeev per grid point inside Griffon in the
    reference, flamelet_kernels.cpp:1329-1341) come from the device eigenvalue kernel (csrc/gb_eig.cu: balancing,
    Hessenberg reduction, Francis QR per block) through the compute_eigenvalues branch of the flamelet Jacobian;
  * the SuperLU linear solver option is not provided (block Thomas only).
"""
import numpy as np
from numpy import inf
from scipy.special import erfinv

from spitfire_b200.library import Dimension, Library
from spitfire_b200.time.integrator import odesolve
from spitfire_b200.time.methods import KennedyCarpenterS6P4Q3
from spitfire_b200.time.nonlinear import SimpleNewtonSolver
from spitfire_b200.time.stepcontrol import PIController

_pressure_depr_warning = 'Deprecation warning in building Spitfire Flamelet instance. ' \
                         'Specifying pressure in flamelet construction is no longer used, and will be removed in a ' \
                         'future version. The pressure is now obtained from the fuel and oxidizer streams.'


class FlameletSpec(object):
    """Boundary streams, mixture-fraction grid and model terms of a flamelet (flamelet.py:29-237); same keyword
    arguments, defaults and pickling behaviour as the reference."""

    def __init__(self, mech_spec=None, initial_condition=None, oxy_stream=None, fuel_stream=None,
                 grid=None, grid_points=None, grid_type='clustered', grid_cluster_intensity=4.,
                 grid_cluster_point='stoichiometric', library_slice=None,
                 max_dissipation_rate=None, stoich_dissipation_rate=None, dissipation_rate=None,
                 dissipation_rate_form='Peters',
                 heat_transfer='adiabatic', convection_temperature=None, radiation_temperature=None,
                 convection_coefficient=None, radiative_emissivity=None, scale_heat_loss_by_temp_range=False,
                 scale_convection_by_dissipation=False, use_linear_ref_temp_profile=False,
                 rates_sensitivity_type='dense', sensitivity_transform_type='exact',
                 include_enthalpy_flux=True, include_variable_cp=True, pressure=None):
        if pressure is not None:
            print(_pressure_depr_warning)
        if library_slice is not None:
            shape = library_slice.shape
            bad = ValueError(f'Error in Flamelet construction from library_slice. The library provided does not appear '
                             f'to have the right dimensions: shape = {shape}. The library must be one-dimensional, '
                             f'with mixture fraction as the first dimension. It can have trivial dimensions that can '
                             f'be squeezed to yield a one-dimensional form.')
            if len(shape) < 1 or any(s > 1 for s in shape[1:]):
                raise bad
            if len(shape) > 1:
                library_slice = Library.squeeze(library_slice)
            if 'mixture_fraction' not in library_slice.dims[0].name:
                raise ValueError(f'Error in Flamelet construction from library_slice. The library provided does not '
                                 f'appear to have mixture fraction as the first dimension. The library provided is:\n '
                                 f'{library_slice}')
            self.mech_spec = library_slice.extra_attributes['mech_spec']
            names = self.mech_spec.species_names
            p = library_slice['pressure'][0]
            self.oxy_stream = self.mech_spec.stream('TPY', (library_slice['temperature'][0], p,
                                                            [library_slice[f'mass fraction {s}'][0] for s in names]))
            self.fuel_stream = self.mech_spec.stream('TPY', (library_slice['temperature'][-1], p,
                                                             [library_slice[f'mass fraction {s}'][-1] for s in names]))
            self.grid = library_slice.mixture_fraction_values
            self.grid_points = self.grid_type = self.grid_cluster_intensity = self.grid_cluster_point = None
            ic = np.zeros((self.grid.size - 2, self.mech_spec.n_species))
            ic[:, 0] = library_slice['temperature'][1:-1]
            for i, s in enumerate(names[:-1]):
                ic[:, 1 + i] = library_slice['mass fraction ' + s][1:-1]
            self.initial_condition = ic.ravel()
        else:
            self.mech_spec = mech_spec
            self.initial_condition = initial_condition
            self.oxy_stream, self.fuel_stream = oxy_stream, fuel_stream
            self.grid, self.grid_points, self.grid_type = grid, grid_points, grid_type
            self.grid_cluster_intensity, self.grid_cluster_point = grid_cluster_intensity, grid_cluster_point
            if grid is not None:
                self.grid_points = self.grid_type = self.grid_cluster_intensity = self.grid_cluster_point = None
            if grid_points is not None:
                self.grid = None
        self.max_dissipation_rate = max_dissipation_rate
        self.stoich_dissipation_rate = stoich_dissipation_rate
        self.dissipation_rate = dissipation_rate
        self.dissipation_rate_form = dissipation_rate_form
        self.heat_transfer = heat_transfer
        self.convection_temperature = convection_temperature
        self.radiation_temperature = radiation_temperature
        self.convection_coefficient = convection_coefficient
        self.radiative_emissivity = radiative_emissivity
        self.scale_heat_loss_by_temp_range = scale_heat_loss_by_temp_range
        self.scale_convection_by_dissipation = scale_convection_by_dissipation
        self.use_linear_ref_temp_profile = use_linear_ref_temp_profile
        self.rates_sensitivity_type = rates_sensitivity_type
        self.sensitivity_transform_type = sensitivity_transform_type
        self.include_enthalpy_flux = include_enthalpy_flux
        self.include_variable_cp = include_variable_cp

    def __getstate__(self):
        d = dict(self.__dict__)
        oxy, fuel = d.pop('oxy_stream'), d.pop('fuel_stream')
        d.update(oxyY=np.copy(oxy.Y), oxyT=oxy.T, fuelY=np.copy(fuel.Y), fuelT=fuel.T, pressure=oxy.P)
        return d

    def __setstate__(self, state):
        state = dict(state)
        mech, p = state['mech_spec'], state.pop('pressure')
        state['oxy_stream'] = mech.stream('TPY', (state.pop('oxyT'), p, state.pop('oxyY')))
        state['fuel_stream'] = mech.stream('TPY', (state.pop('fuelT'), p, state.pop('fuelY')))
        self.__init__(**state)


def compute_dissipation_rate(mixture_fraction, max_dissipation_rate, form='Peters'):
    """chi(Z): the form of N. Peters, Turbulent Combustion (2000), or a constant (flamelet.py:240-266)"""
    if form in ('Peters', 'peters'):
        return max_dissipation_rate * np.exp(-2. * (erfinv(2. * mixture_fraction - 1.)) ** 2)
    return np.zeros_like(mixture_fraction) + max_dissipation_rate


# ----------------------------------------------------------------------------------------------------------------------
# kernels behind a batch of flamelets
# ----------------------------------------------------------------------------------------------------------------------
class _BatchOps(object):
    """Right-hand side, Jacobian and block-Thomas solves for F flamelets that share mechanism, grid size and model
    flags. All arrays are torch tensors on `self.device`; member f's data is row f."""

    def __init__(self, flamelets):
        import torch
        self.torch = torch
        f0 = flamelets[0]
        self.F = len(flamelets)
        self.g = f0._griffon
        self._mechanism_of_members = f0._mechanism
        self.ns, self.nzi = f0._n_equations, f0._nz_interior
        self.ndof = self.ns * self.nzi
        self.nelem = f0._jac_nelements_griffon
        self.pressure = float(f0._pressure)
        self.adiabatic = f0._heat_transfer == 'adiabatic'
        self.flags = (f0._include_enthalpy_flux, f0._include_variable_cp, f0._scale_heat_loss_by_temp_range)
        self.rsopt, self.stopt = f0._rsopt, f0._stopt
        for fl in flamelets[1:]:
            same = (fl._griffon is self.g and fl._nz_interior == self.nzi and abs(fl._pressure - self.pressure) < 1e-12
                    and (fl._heat_transfer == 'adiabatic') == self.adiabatic and
                    (fl._include_enthalpy_flux, fl._include_variable_cp, fl._scale_heat_loss_by_temp_range) == self.flags
                    and np.array_equal(fl._state_oxy, f0._state_oxy) and np.array_equal(fl._state_fuel, f0._state_fuel))
            if not same:
                raise ValueError('FlameletBatch: members must share mechanism, streams, pressure, grid size, '
                                 'heat-transfer type and model flags')
        from spitfire_b200 import griffon as gmod
        self.on_device = isinstance(self.g, gmod.PyCombustionKernels)
        if self.on_device and not torch.cuda.is_available():
            raise gmod.GriffonB200Error('the B200 Griffon path needs a CUDA device (there is no CPU fallback)')
        self.device = torch.device('cuda') if self.on_device else torch.device('cpu')
        self.gmod = gmod
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)
        self.oxy, self.fuel = t(f0._state_oxy), t(f0._state_fuel)
        self.cmaj = t(np.array([fl._maj_coeff_griffon for fl in flamelets]))
        self.csub = t(np.array([fl._sub_coeff_griffon for fl in flamelets]))
        self.csup = t(np.array([fl._sup_coeff_griffon for fl in flamelets]))
        self.mc = t(np.array([fl._mcoeff_griffon for fl in flamelets]))
        self.nc = t(np.array([fl._ncoeff_griffon for fl in flamelets]))
        self.chi = t(np.array([fl._x for fl in flamelets]))
        if self.adiabatic:
            self.heat = None
        else:
            self.heat = [t(np.array([getattr(fl, a) for fl in flamelets])) for a in
                         ('_T_conv', '_T_rad', '_h_conv', '_h_rad')]
        self.scales = t(np.array([fl._variable_scales for fl in flamelets]))
        self.max_chi = t(np.array([fl._max_dissipation_rate for fl in flamelets]))
        self._prm_cache = dict()

    # -- helpers -----------------------------------------------------------------------------------------------------
    def _params(self, idx, key=None):
        """C-ABI parameter block for the members `idx` (a LongTensor) -- compacts the per-member arrays. `key` (a
        hashable naming the member set, e.g. the tuple of indices) lets repeated selections reuse the compacted
        arrays and the parameter struct."""
        if key is not None and key in self._prm_cache:
            return self._prm_cache[key]
        full = idx.numel() == self.F and (key is None or key == tuple(range(self.F)))
        sel = (lambda a: a) if full else (lambda a: a.index_select(0, idx).contiguous())
        d = dict(cmaj=sel(self.cmaj), csub=sel(self.csub), csup=sel(self.csup), mc=sel(self.mc), nc=sel(self.nc),
                 chi=sel(self.chi))
        heat = [None] * 4 if self.adiabatic else [sel(h) for h in self.heat]
        nz = self.chi.shape[1]
        prm = self.g._flamelet_params(self.pressure, self.oxy, self.fuel, self.adiabatic, heat[0], heat[1], heat[2],
                                      heat[3], self.nzi, d['cmaj'], d['csub'], d['csup'], d['mc'], d['nc'], d['chi'],
                                      *self.flags, strides=(0 if self.adiabatic else self.nzi, self.ndof, self.nzi, nz))
        out = (prm, (d, heat))  # keep the tensors alive while the kernels run
        if key is not None:
            if len(self._prm_cache) > 256:
                self._prm_cache.clear()
            self._prm_cache[key] = out
        return out

    def _all(self):
        return self.torch.arange(self.F, device=self.device)

    # -- kernels -----------------------------------------------------------------------------------------------------
    def rhs(self, state, idx=None, key=None):
        """flamelet_rhs for the members idx (default all); state [len(idx), ndof]"""
        torch = self.torch
        idx = self._all() if idx is None else idx
        out = torch.empty_like(state)
        if self.on_device:
            prm, keep = self._params(idx, key)
            self.g.flamelet_rhs_batch(state.shape[0], state.contiguous(), prm, out)
            del keep
        else:
            sn, on = state.numpy(), out.numpy()
            for k, f in enumerate(idx.tolist()):
                self._host_rhs(f, sn[k], on[k])
        return out

    def _host_args(self, f):
        n1 = np.zeros(1)
        heat = [n1] * 4 if self.adiabatic else [h[f].numpy() for h in self.heat]
        return (self.pressure, self.oxy.numpy(), self.fuel.numpy(), self.adiabatic, heat[0], heat[1], heat[2], heat[3],
                self.nzi, self.cmaj[f].numpy(), self.csub[f].numpy(), self.csup[f].numpy(), self.mc[f].numpy(),
                self.nc[f].numpy(), self.chi[f].numpy())

    def _host_rhs(self, f, state, out):
        self.g.flamelet_rhs(np.ascontiguousarray(state), *self._host_args(f), *self.flags, out)

    def jac(self, state, idx=None, scale_and_offset=False, prefactor=1., key=None):
        """BTDDOD Jacobian (or prefactor*J - I) of the members idx; [len(idx), nelem]"""
        torch = self.torch
        idx = self._all() if idx is None else idx
        n = state.shape[0]
        if self.on_device:
            out = torch.empty((n, self.nelem), dtype=torch.float64, device=self.device)
            prm, keep = self._params(idx, key)
            self.g.flamelet_jacobian_batch(n, state.contiguous(), prm, out, scale_and_offset=scale_and_offset,
                                           prefactor=prefactor, rates_sens_option=self.rsopt,
                                           sens_transform_option=self.stopt)
            del keep
        else:
            out = torch.zeros((n, self.nelem), dtype=torch.float64)
            sn, on = state.numpy(), out.numpy()
            null = np.zeros(1)
            for k, f in enumerate(idx.tolist()):
                self.g.flamelet_jacobian(np.ascontiguousarray(sn[k]), *self._host_args(f), False, 0., scale_and_offset,
                                         prefactor, self.rsopt, self.stopt, *self.flags, null, on[k])
        return out

    def jac_and_eig(self, state, diffterm, idx=None, key=None):
        """(Jacobian, explosive-mode bound) of the members idx, the pair `Flamelet._*_jac_and_eig` returns
        (flamelet.py:715-741). diffterm: [n] (one value per member; the reference passes a scalar per call). On the
        device both come out of one C-ABI call: the eigenvalue bound is computed with diffterm = 0 and shifted per member
        here, max(max(Re) - d, 0) = max(max(max(Re), 0) - d, 0) for d >= 0."""
        torch = self.torch
        idx = self._all() if idx is None else idx
        n = state.shape[0]
        if not self.on_device:
            out = torch.zeros((n, self.nelem), dtype=torch.float64)
            e = torch.zeros((n, self.ndof), dtype=torch.float64)
            sn, on, en, dn = state.numpy(), out.numpy(), e.numpy(), diffterm.numpy()
            for k, f in enumerate(idx.tolist()):
                self.g.flamelet_jacobian(np.ascontiguousarray(sn[k]), *self._host_args(f), True, float(dn[k]), False, 1.,
                                         self.rsopt, self.stopt, *self.flags, en[k], on[k])
            return out, e
        out = torch.empty((n, self.nelem), dtype=torch.float64, device=self.device)
        e0 = torch.empty((n, self.ndof), dtype=torch.float64, device=self.device)
        prm, keep = self._params(idx, key)
        self.g.flamelet_jacobian_batch(n, state.contiguous(), prm, out, compute_eigenvalues=True, diffterm=0.,
                                       rates_sens_option=self.rsopt, sens_transform_option=self.stopt, out_expeig=e0)
        del keep
        return out, torch.clamp(e0 - torch.clamp(diffterm, min=0.)[:, None], min=0.)

    # On the device the steady solvers use the explicit block inverses the factorisation forms anyway: the linear solve
    # only proposes the update of an iteration that converges on the true residual, and the inverse-based back sweep is
    # 3-4x shorter than the pivoted triangular solves (griffon_b200.h, gb_btddod_full_*_inv_batch).
    explicit_inverse_solves = True
    gauss_jordan_inverses = True  # k_btddod_invert instead of LU + dgetrs-on-identity (gb_btinv.cu)
    twisted_elimination = True    # ... eliminating from both ends at once (two CTAs of a cluster per flamelet)
    _rows32 = (None, None)  # last (int64 rows, int32 copy) pair handed to the solve kernel

    def factor_store(self, F):
        """zeroed arrays for the factors of F systems: (J, L, pivots[, Dinv])"""
        torch = self.torch
        out = [torch.zeros((F, self.nelem), dtype=torch.float64, device=self.device),
               torch.zeros((F, self.nzi * self.ns * self.ns), dtype=torch.float64, device=self.device),
               torch.zeros((F, self.ndof), dtype=torch.int32, device=self.device)]
        if self.on_device and self.explicit_inverse_solves:
            out.append(torch.zeros((F, self.nzi * self.ns * self.ns), dtype=torch.float64, device=self.device))
        return tuple(out)

    def factorize_into(self, store, rows, J):
        """factorise the systems J and keep the factors at positions `rows` of `store`"""
        fact = self.factorize(J, with_inverse=len(store) == 4)
        for dst, src in zip(store, fact):
            dst[rows] = src

    def factorize(self, J, with_inverse=False):
        """block-Thomas factorisation in place of the systems J [n, nelem]; returns (J, L, pivots[, Dinv]).
        with_inverse (device path only) also keeps the explicit inverses of the factorised diagonal blocks so that
        `solve` can run its back sweep as matrix-vector products (griffon_b200.h, gb_btddod_full_*_inv_batch)."""
        torch = self.torch
        n = J.shape[0]
        L = torch.zeros((n, self.nzi * self.ns * self.ns), dtype=torch.float64, device=self.device)
        piv = torch.zeros((n, self.ndof), dtype=torch.int32, device=self.device)
        if self.on_device and with_inverse:
            Dinv = torch.empty((n, self.nzi * self.ns * self.ns), dtype=torch.float64, device=self.device)
            if getattr(self, 'gauss_jordan_inverses', True):
                # the solvers only ever apply these factors through solve_inv: inverses alone, by the short-chain kernel
                self.gmod.btddod_full_invert(J, self.nzi, self.ns, L, Dinv, n_systems=n,
                                             twisted=getattr(self, 'twisted_elimination', True))
            else:
                self.gmod.btddod_full_factorize_inv(J, self.nzi, self.ns, L, piv, Dinv, n_systems=n)
            return J, L, piv, Dinv
        if self.on_device:
            self.gmod.py_btddod_full_factorize(J, self.nzi, self.ns, L, piv, n_systems=n)
        else:
            Jn, Ln, pn = J.numpy(), L.numpy(), piv.numpy()
            for k in range(n):
                self.g.btddod_full_factorize(Jn[k], self.nzi, self.ns, Ln[k], pn[k])
        return J, L, piv

    def solve(self, fact, rhs, rows=None):
        """solve with the factors of the members `rows` (positions in the factor arrays; default all)"""
        torch = self.torch
        n = rhs.shape[0]
        if self.on_device and len(fact) == 4:
            # the kernel addresses the members' factors in place (no gather of ~8 MB per member)
            x = torch.empty_like(rhs)
            if rows is not None:
                if self._rows32[0] is not rows:
                    self._rows32 = (rows, rows.to(torch.int32))
                rows = self._rows32[1]
            self.gmod.btddod_full_solve_inv(fact[0], fact[1], fact[3], rhs.contiguous(), self.nzi, self.ns, x, n_systems=n,
                                            system_rows=rows)
            return x
        x = torch.zeros_like(rhs)
        if rows is not None:
            fact = tuple(a.index_select(0, rows) for a in fact)
        J, L, piv = fact[:3]
        if self.on_device:
            self.gmod.py_btddod_full_solve(J.contiguous(), L.contiguous(), piv.contiguous(), rhs.contiguous(), self.nzi,
                                           self.ns, x, n_systems=n)
        else:
            Jn, Ln, pn, rn, xn = J.numpy(), L.numpy(), piv.numpy(), rhs.numpy(), x.numpy()
            for k in range(n):
                self.g.btddod_full_solve(Jn[k], Ln[k], pn[k], np.ascontiguousarray(rn[k]), self.nzi, self.ns, xn[k])
        return x

    def newton_stage(self, fact, rows, idx, key, explicit, q, dt, gamma, weights, tolerance, max_iterations, x, f, res,
                     conv, work, counter):
        """device path only: the Newton loop of one implicit stage in one C-ABI call; (members left, iterations)"""
        prm, keep = self._params(idx, key)
        if rows is not None:
            if self._rows32[0] is not rows:
                self._rows32 = (rows, rows.to(self.torch.int32))
            rows = self._rows32[1]
        out = self.g.flamelet_newton_stage_batch(x.shape[0], prm, fact[0], fact[1], fact[3], rows, explicit, q, dt, gamma,
                                                 weights, tolerance, max_iterations, x, f, res, conv, work, counter)
        del keep
        return out

    def esdirk_stages(self, fact, rows, idx, key, tableau, q, dt, gamma, weights, tolerance, max_iterations, x, f, res,
                      explicit, K, stage, iters, nlfail, done, work, counter):
        """device path only: every implicit stage of one step in one C-ABI call, the members taking the stages
        independently of each other; (members not done, rounds of kernels)"""
        prm, keep = self._params(idx, key)
        if rows is not None:
            if self._rows32[0] is not rows:
                self._rows32 = (rows, rows.to(self.torch.int32))
            rows = self._rows32[1]
        out = self.g.flamelet_esdirk_stages_batch(x.shape[0], prm, fact[0], fact[1], fact[3], rows, tableau, q, dt, gamma,
                                                  weights, tolerance, max_iterations, x, f, res, explicit, K, stage, iters,
                                                  nlfail, done, work, counter)
        del keep
        return out

    def second_griffon(self, k=0):
        """the k-th extra Griffon handle of the same mechanism (its own work arrays), for kernels that run on another
        stream while this object's handle is busy -- a handle serves one stream at a time (griffon_b200.h)"""
        extra = self.__dict__.setdefault('_extra_mechs', dict())
        if k not in extra:
            from spitfire_b200.mechanism import ChemicalMechanismSpec
            m = self._mechanism_of_members
            extra[k] = ChemicalMechanismSpec(mech_data=m.mech_data, griffon_factory=m._griffon_factory)
        return extra[k].griffon

    def member_params(self, m):
        """C-ABI parameter block of the single member m, addressing its rows of the batch's arrays in place (cached)"""
        cache = self.__dict__.setdefault('_member_prm', dict())
        if m not in cache:
            heat = [None] * 4 if self.adiabatic else [h[m:m + 1] for h in self.heat]
            nz = self.chi.shape[1]
            tens = (self.cmaj[m:m + 1], self.csub[m:m + 1], self.csup[m:m + 1], self.mc[m:m + 1], self.nc[m:m + 1],
                    self.chi[m:m + 1])
            prm = self.g._flamelet_params(self.pressure, self.oxy, self.fuel, self.adiabatic, heat[0], heat[1], heat[2],
                                          heat[3], self.nzi, *tens, *self.flags,
                                          strides=(0 if self.adiabatic else self.nzi, self.ndof, self.nzi, nz))
            cache[m] = (prm, tens, heat)
        return cache[m][0]

    def jac_rows_on(self, g, q, members, out):
        """BTDDOD Jacobians of the listed members of q [F, ndof] into the rows of out [len(members), nelem], one launch
        per member through the Griffon handle g (see second_griffon) -- no gathers of states or coefficient arrays"""
        for k, m in enumerate(members):
            g.flamelet_jacobian_batch(1, q[m:m + 1], self.member_params(m), out[k:k + 1], scale_and_offset=False,
                                      prefactor=1., rates_sens_option=self.rsopt, sens_transform_option=self.stopt)
        return out

    def jac_scaled_row_on(self, g, q, m, prefactor, out):
        """row m of out [F, nelem] <- prefactor * J(q[m]) - I (the scale-and-offset option of the Jacobian kernel,
        flamelet_kernels.cpp:1395-1408), through the Griffon handle g"""
        g.flamelet_jacobian_batch(1, q[m:m + 1], self.member_params(m), out[m:m + 1], scale_and_offset=True,
                                  prefactor=prefactor, rates_sens_option=self.rsopt, sens_transform_option=self.stopt)

    def invert_row(self, J, L, Dinv, m):
        """block elimination by inverses of system m of J [F, nelem] into row m of L and Dinv (what `factorize` with
        with_inverse does for a batch, in place in the caller's arrays)"""
        self.gmod.btddod_full_invert(J[m:m + 1], self.nzi, self.ns, L[m:m + 1], Dinv[m:m + 1], n_systems=1,
                                     twisted=getattr(self, 'twisted_elimination', True))

    def nonfinite_rows(self, a, b):
        """bool [n]: the member's row of a or of b holds an Inf or NaN. On the device one kernel writes the flags
        (gb_count_nonfinite_members_batch), on the host the tensor expressions do."""
        torch = self.torch
        if self.on_device:
            flags = torch.empty(a.shape[0], dtype=torch.int32, device=self.device)
            self.gmod.count_nonfinite_members(a.contiguous(), b.contiguous(), flags)
            return flags.bool()
        return ~torch.isfinite(a).all(dim=1) | ~torch.isfinite(b).all(dim=1)

    def add_to_block_diagonal(self, J, matrix_scale, diagonal, diag_scale):
        """J <- matrix_scale*J + diag_scale*diag(diagonal) for every system (btddod_scale_and_add_diagonal)"""
        n = J.shape[0]
        if self.on_device:
            self.gmod.py_btddod_scale_and_add_diagonal(J, matrix_scale, diagonal.contiguous(), diag_scale, self.nzi,
                                                       self.ns, n_systems=n)
        else:
            Jn, dn = J.numpy(), diagonal.numpy()
            for k in range(n):
                self.g.btddod_scale_and_add_diagonal(Jn[k], matrix_scale, np.ascontiguousarray(dn[k]), diag_scale,
                                                     self.nzi, self.ns)
        return J


class FlameletBatch(object):
    """F flamelets (same mechanism, streams, grid size and model flags; their own dissipation rates, heat-transfer
    arrays and states) solved together. The steady solvers keep per-member control flow; a member that converges or
    fails simply drops out of the active set, so a diverging member does not stall the others."""

    def __init__(self, flamelets):
        self.flamelets = list(flamelets)
        self.ops = _BatchOps(self.flamelets)

    def __len__(self):
        return len(self.flamelets)

    def _initial(self, initial_guess):
        torch, ops = self.ops.torch, self.ops
        if initial_guess is None:
            g = np.array([fl._initial_state for fl in self.flamelets])
        else:
            g = np.array(initial_guess, dtype=np.float64).reshape(len(self.flamelets), ops.ndof)
        return torch.as_tensor(np.ascontiguousarray(g)).to(ops.device)

    @staticmethod
    def _norm(x, order):
        if order == inf or order == np.inf:
            return x.abs().amax(dim=1)
        return x.abs().pow(order).sum(dim=1).pow(1. / order)

    # ------------------------------------------------------------------------------------------------------------------
    def steady_solve_newton(self, initial_guess=None, tolerance=1.e-6, max_iterations=10, max_factor_line_search=1.5,
                            max_allowed_residual=1.e6, min_allowable_state_var=-1.e-6, norm_order=np.inf,
                            log_rate=100000, verbose=True):
        """Chord Newton with residual line search for every member (flamelet.py:1352-1485 member by member).
        Returns (states [F, ndof] ndarray, iteration counts [F], converged [F])."""
        torch, ops = self.ops.torch, self.ops
        F = len(self.flamelets)
        dev = ops.device
        state = self._initial(initial_guess)
        inv_scales = 1. / ops.scales
        rhs = ops.rhs(state)
        res = torch.full((F,), tolerance + 1., dtype=torch.float64, device=dev)
        iters = torch.zeros(F, dtype=torch.int64, device=dev)
        active = torch.ones(F, dtype=torch.bool, device=dev)
        failed = torch.zeros(F, dtype=torch.bool, device=dev)
        need_jac = torch.ones(F, dtype=torch.bool, device=dev)
        factors = ops.factor_store(F)
        while True:
            active = active & (res > tolerance) & (iters < max_iterations) & ~failed
            idx = torch.nonzero(active).flatten()
            if idx.numel() == 0:
                break
            iters[idx] += 1
            ij = torch.nonzero(active & need_jac).flatten()
            if ij.numel():
                Jn = ops.jac(state.index_select(0, ij), ij).neg_()
                ops.factorize_into(factors, ij, Jn)
                need_jac[ij] = False
            dstate = ops.solve(factors, rhs.index_select(0, idx), rows=idx)
            norm_old = self._norm(rhs.index_select(0, idx) * inv_scales.index_select(0, idx), norm_order)
            s_act = state.index_select(0, idx)
            rhs_new = ops.rhs(s_act + dstate, idx)
            bad = ops.nonfinite_rows(dstate, rhs_new)  # (the reference's NaN / Inf scans, flamelet.py:1435, 1443)
            alpha = torch.ones(idx.numel(), dtype=torch.float64, device=dev)
            while True:
                nrm = self._norm(rhs_new * inv_scales.index_select(0, idx), norm_order)
                cut = (nrm > max_factor_line_search * norm_old) & (alpha > 0.001) & ~bad
                ic = torch.nonzero(cut).flatten()
                if ic.numel() == 0:
                    break
                alpha[ic] *= 0.5
                dstate[ic] = dstate[ic] * alpha[ic, None]
                rhs_new[ic] = ops.rhs(s_act.index_select(0, ic) + dstate.index_select(0, ic), idx.index_select(0, ic))
                need_jac[idx.index_select(0, ic)] = True
                if verbose:
                    for a in alpha[ic].tolist():
                        print(f'  line search reducing step size to {a:.3f}')
            ok = ~bad
            new_state = s_act + dstate
            state[idx[ok]] = new_state[ok]
            rhs[idx[ok]] = rhs_new[ok]
            r = self._norm(rhs_new * inv_scales.index_select(0, idx), norm_order)
            res[idx[ok]] = r[ok]
            too_big = ok & (r > max_allowed_residual)
            negative = ok & (new_state.amin(dim=1) < min_allowable_state_var)
            failed[idx[bad | too_big | negative]] = True
            if verbose:
                for m, (b, t, n) in zip(idx.tolist(), zip(bad.tolist(), too_big.tolist(), negative.tolist())):
                    if b:
                        print('nan/inf detected in state update!')
                    elif t:
                        print('Convergence failure! Residual of {:.2e} detected, exceeds the maximum allowable value '
                              'of {:.2e}.'.format(float(res[m]), max_allowed_residual))
                    elif n:
                        print('Convergence failure! Mass fraction or temperature < min_allowable_state_var detected.')
        converged = (~failed) & (res <= tolerance)
        state = torch.where(state < 0, torch.zeros_like(state), state)
        if verbose:
            for c, f in zip(converged.tolist(), failed.tolist()):
                if not c and not f:
                    print('Convergence failure! Too many iterations required, more than allowable '
                          '{:}.'.format(max_iterations))
        return state.cpu().numpy(), iters.cpu().numpy(), converged.cpu().numpy()

    # ------------------------------------------------------------------------------------------------------------------
    def steady_solve_psitc(self, initial_guess=None, tolerance=1.e-6, max_iterations=400, min_allowable_state_var=-1.e-6,
                           ds_init=1., ds_init_decrease=4., adaptive_restart=True, diffusion_factor=4., global_ds=False,
                           ds_safety=0.1, ds_ramp=1.1, ds_max=1.e4, max_factor_line_search=1.5, max_allowed_residual=1.e6,
                           log_rate=100000, norm_order=np.inf, max_recursion_depth=20, verbose=True):
        """Adaptive pseudo-transient continuation for every member (flamelet.py:1487-1703): per-dof pseudo-time steps
        from the explosive-mode bound, Jacobian refreshed every 8 iterations or while the residual is above 1e-2,
        restart from the initial guess with a 4x smaller first step on failure (up to max_recursion_depth times).
        Returns (states, iteration counts, converged, min(ds)) per member."""
        torch, ops = self.ops.torch, self.ops
        F = len(self.flamelets)
        dev = ops.device
        guess = self._initial(initial_guess)
        state = guess.clone()
        inv_scales = 1. / ops.scales
        diffterm = diffusion_factor * ops.max_chi
        ds0 = torch.full((F,), float(ds_init), dtype=torch.float64, device=dev)
        ds = ds0[:, None].repeat(1, ops.ndof)
        depth = torch.zeros(F, dtype=torch.int64, device=dev)
        iters = torch.zeros(F, dtype=torch.int64, device=dev)
        jac_age = torch.zeros(F, dtype=torch.int64, device=dev)
        res = torch.full((F,), tolerance + 1., dtype=torch.float64, device=dev)
        rhs = ops.rhs(state)
        active = torch.ones(F, dtype=torch.bool, device=dev)
        failed = torch.zeros(F, dtype=torch.bool, device=dev)
        need_jac = torch.ones(F, dtype=torch.bool, device=dev)
        factors = ops.factor_store(F)
        jac_refresh_age = 8

        def restart(members):
            """members (LongTensor) failed an iteration: restart them or give up"""
            if members.numel() == 0:
                return
            give_up = (depth[members] > max_recursion_depth) | (not adaptive_restart)
            failed[members[give_up]] = True
            again = members[~give_up]
            if again.numel():
                if verbose:
                    print('Failure detected in steady_solve_psitc! Restarting...')
                depth[again] += 1
                ds0[again] = ds0[again] / ds_init_decrease
                ds[again] = ds0[again][:, None]
                state[again] = guess[again]
                rhs[again] = ops.rhs(state.index_select(0, again), again)
                iters[again] = 0
                jac_age[again] = 0
                res[again] = tolerance + 1.
                need_jac[again] = True

        while True:
            active = (res > tolerance) & (iters < max_iterations) & ~failed
            idx = torch.nonzero(active).flatten()
            if idx.numel() == 0:
                break
            iters[idx] += 1
            ij = torch.nonzero(active & need_jac).flatten()
            if ij.numel():
                Jp, expeig = ops.jac_and_eig(state.index_select(0, ij), diffterm.index_select(0, ij), ij)
                dsj = torch.minimum(torch.minimum(ds_safety / (expeig + 1.e-16), ds_ramp * ds.index_select(0, ij)),
                                    torch.full_like(expeig, ds_max))
                first = (iters.index_select(0, ij) == 1) | bool(global_ds)
                dsj = torch.where(first[:, None], dsj.amin(dim=1, keepdim=True).expand_as(dsj), dsj)
                ds[ij] = dsj
                ops.add_to_block_diagonal(Jp, -1., 1. / dsj, 1.)
                ops.factorize_into(factors, ij, Jp)
                jac_age[ij] = 0
            aged = torch.nonzero(active & ~need_jac).flatten()
            jac_age[aged] += 1
            need_jac[idx] = (jac_age.index_select(0, idx) == jac_refresh_age) | (res.index_select(0, idx) > 1.e-2)
            dstate = ops.solve(factors, rhs.index_select(0, idx), rows=idx)
            norm_old = self._norm(rhs.index_select(0, idx) * inv_scales.index_select(0, idx), norm_order)
            s_act = state.index_select(0, idx)
            rhs_new = ops.rhs(s_act + dstate, idx)
            bad = ops.nonfinite_rows(dstate, rhs_new)  # (the reference's NaN / Inf scans, flamelet.py:1435, 1443)
            alpha = torch.ones(idx.numel(), dtype=torch.float64, device=dev)
            while True:
                nrm = self._norm(rhs_new * inv_scales.index_select(0, idx), norm_order)
                cut = (nrm > max_factor_line_search * norm_old) & (alpha > 0.001) & ~bad
                ic = torch.nonzero(cut).flatten()
                if ic.numel() == 0:
                    break
                alpha[ic] *= 0.5
                dstate[ic] = dstate[ic] * alpha[ic, None]
                rhs_new[ic] = ops.rhs(s_act.index_select(0, ic) + dstate.index_select(0, ic), idx.index_select(0, ic))
                need_jac[idx.index_select(0, ic)] = True
            ok = ~bad
            new_state = s_act + dstate
            r = self._norm(rhs_new * inv_scales.index_select(0, idx), norm_order)
            trouble = bad | (r > max_allowed_residual) | (new_state.amin(dim=1) < min_allowable_state_var)
            keep = ~trouble
            state[idx[keep]] = new_state[keep]
            rhs[idx[keep]] = rhs_new[keep]
            res[idx[keep]] = r[keep]
            restart(idx[trouble])
        converged = (~failed) & (res <= tolerance)
        state = torch.where(state < 0, torch.zeros_like(state), state)
        return state.cpu().numpy(), iters.cpu().numpy(), converged.cpu().numpy(), ds.amin(dim=1).cpu().numpy()

    # ------------------------------------------------------------------------------------------------------------------
    def compute_steady_state(self, initial_guess=None, tolerance=1.e-6, verbose=False, use_psitc=True, newton_args=None,
                             psitc_args=None, transient_args=None):
        """Newton for every member, pseudo-transient continuation for those that fail, ESDIRK64 time integration for
        the rest (flamelet.py:1705-1778). Returns (states [F, ndof], solver tag per member)."""
        F = len(self.flamelets)
        nargs = dict(tolerance=tolerance, log_rate=1, verbose=verbose, max_iterations=38)
        nargs.update(newton_args or {})
        guess = np.array([fl._initial_state for fl in self.flamelets]) if initial_guess is None else \
            np.array(initial_guess, dtype=np.float64).reshape(F, -1)
        states, _, conv = self.steady_solve_newton(initial_guess=guess, **nargs)
        tags = np.array(['newton'] * F, dtype=object)
        out = np.array(states)
        todo = np.nonzero(~conv)[0]
        mds = np.full(F, 1.e-6)
        if todo.size and use_psitc:
            sub = FlameletBatch([self.flamelets[i] for i in todo])
            pargs = dict(tolerance=tolerance, log_rate=1, verbose=verbose, max_iterations=400)
            pargs.update(psitc_args or {})
            ps, _, pconv, pmds = sub.steady_solve_psitc(initial_guess=guess[todo], **pargs)
            for k, i in enumerate(todo):
                mds[i] = pmds[k]
                if pconv[k]:
                    out[i], tags[i], conv[i] = ps[k], 'psitc', True
        for i in np.nonzero(~conv)[0]:
            fl = self.flamelets[i]
            eargs = dict(steady_tolerance=tolerance, transient_tolerance=1.e-8, max_time_step=1e4, write_log=verbose,
                         log_rate=1, first_time_step=1e-2 * mds[i], maximum_steps_per_jacobian=10,
                         save_first_and_last_only=True)
            eargs.update(transient_args or {})
            fl._current_state = np.copy(guess[i])
            fl._current_time = 0.
            fl.integrate_to_steady(**eargs)
            out[i], tags[i] = fl._current_state, 'esdirk'
        for i, fl in enumerate(self.flamelets):
            fl._current_state = np.copy(out[i])
        return out, tags


    # ------------------------------------------------------------------------------------------------------------------
    def integrate_for_heat_loss(self, temperature_tolerance=0.05, steady_tolerance=1.e-4, first_time_step=1.e-6,
                                max_time_step=1.e-3, minimum_time_step_count=40, transient_tolerance=1.e-10,
                                maximum_steps_per_jacobian=10, nonlinear_solve_tolerance=1.e-12, **unused):
        """`Flamelet.integrate_for_heat_loss` (flamelet.py:1263-1286) for every member at once: ESDIRK64 with each
        member's own adaptive step until its temperature profile is nearly linear or it is steady
        (spitfire_b200.time.batched). Returns (libraries over (time, mixture fraction), failed flags)."""
        from spitfire_b200.time import batched
        ops = self.ops
        # (members advancing independently of each other where the device path offers it: time/batched.py)
        integrate_batch = batched.integrate_batch_async if batched.can_integrate_async(ops, len(self.flamelets)) else \
            batched.integrate_batch
        T_bc_max = max(self.flamelets[0]._oxy_stream.T, self.flamelets[0]._fuel_stream.T)

        def stop(t, q, residual, nsteps):
            return (q.amax(dim=1) < (1. + temperature_tolerance) * T_bc_max) | (residual < steady_tolerance)

        # the same test on host rows (numpy), for the integrator that has the completed members' states on the host
        stop.host = lambda t, q, residual, nsteps: \
            (q.max(axis=1) < (1. + temperature_tolerance) * T_bc_max) | (residual < steady_tolerance)

        q0 = ops.torch.as_tensor(np.array([fl._current_state for fl in self.flamelets])).to(ops.device)
        times, states, failed = integrate_batch(ops, q0, stop, first_time_step=first_time_step,
                                                max_time_step=max_time_step,
                                                minimum_time_step_count=minimum_time_step_count,
                                                transient_tolerance=transient_tolerance,
                                                maximum_steps_per_jacobian=maximum_steps_per_jacobian,
                                                nonlinear_solve_tolerance=nonlinear_solve_tolerance)
        libs = []
        for fl, t, st in zip(self.flamelets, times, states):
            fl._current_state, fl._current_time = np.copy(st[-1]), float(t[-1])
            lib = Library(Dimension('time', t), Dimension('mixture_fraction', fl._z))
            lib.extra_attributes['mech_spec'] = fl._mechanism
            fl._fill_library(lib, st, lead=True)
            libs.append(lib)
        return libs, failed


# ----------------------------------------------------------------------------------------------------------------------
class Flamelet(object):
    """Solve the nonpremixed flamelet equations (mirror of flamelet.py:269-1778 on the B200 Griffon path)"""

    _heat_transfers = ['adiabatic', 'nonadiabatic']
    _initializations = ['unreacted', 'equilibrium', 'Burke-Schumann', 'linear-TY']
    _grid_types = ['uniform', 'clustered']
    _rates_sensitivity_option_dict = {'dense': 0, 'no-TBAF': 1, 'sparse': 2}
    _sensitivity_transform_option_dict = {'exact': 0}
    _grid_cache = dict()

    @classmethod
    def _uniform_grid(cls, grid_points):
        z = np.linspace(0., 1., grid_points)
        return z, z[1:] - z[:-1]

    @classmethod
    def _clustered_grid(cls, grid_points, grid_cluster_point, grid_cluster_intensity=6.):
        """sinh clustering around a mixture fraction (J. D. Anderson, Computational Fluid Dynamics, 1995, pp. 585-588),
        flamelet.py:299-341"""
        if grid_cluster_intensity < 1.e-16:
            raise ValueError('cluster_coeff must be strictly positive! Given value: ' + str(grid_cluster_intensity))
        if grid_cluster_point < 0. or grid_cluster_point > 1.:
            raise ValueError('z_cluster must be between 0 and 1! Given value: ' + str(grid_cluster_point))
        b, zc = grid_cluster_intensity, grid_cluster_point
        key = (int(grid_points), float(zc), float(b))
        hit = cls._grid_cache.get(key)
        if hit is None:  # (a library build constructs a few hundred flamelets on the same grid)
            z = np.linspace(0., 1., grid_points)
            zo = 1.0 / (2.0 * b) * np.log((1. + (np.exp(b) - 1.) * zc) / (1. + (np.exp(-b) - 1.) * zc))
            a = np.sinh(b * zo)
            for i in range(grid_points):
                z[i] = zc / a * (np.sinh(b * (z[i] - zo)) + a)
            z[-1] = 1.
            if len(cls._grid_cache) > 64:
                cls._grid_cache.clear()
            hit = cls._grid_cache[key] = (z, z[1:] - z[:-1])
        return hit[0].copy(), hit[1].copy()

    @classmethod
    def make_clustered_grid(cls, grid_points, grid_cluster_point, grid_cluster_intensity=6.):
        return cls._clustered_grid(grid_points, grid_cluster_point, grid_cluster_intensity)[0]

    @classmethod
    def _compute_dissipation_rate(cls, mixture_fraction, max_dissipation_rate, form='Peters'):
        return compute_dissipation_rate(mixture_fraction, max_dissipation_rate, form)

    def _heat_array(self, value, name, attr):
        if value is None:
            raise ValueError('Flamelet specifications: Nonadiabatic heat transfer was selected but no ' + name +
                             ' argument was given.')
        if isinstance(value, float):
            setattr(self, attr, value + np.zeros(self._nz_interior))
        elif isinstance(value, np.ndarray):
            setattr(self, attr, np.array(value, dtype=np.float64))
        else:
            raise ValueError(name + ' was not given as a float (constant) or numpy array')

    def __init__(self, flamelet_specs=None, *args, **kwargs):
        if isinstance(flamelet_specs, dict):
            fs = FlameletSpec(**flamelet_specs)
        elif flamelet_specs is None:
            fs = FlameletSpec(*args, **kwargs)
        else:
            fs = flamelet_specs
        self._oxy_stream, self._fuel_stream = fs.oxy_stream, fs.fuel_stream
        perr = np.abs(self._fuel_stream.P - self._oxy_stream.P) / self._oxy_stream.P
        if perr > 1.e-6:
            raise ValueError(f'Error in building Spitfire Flamelet instance. The pressure of the fuel and oxidizer '
                             f'streams must be the same. Fuel pressure = {self._fuel_stream.P}, oxidizer pressure = '
                             f'{self._oxy_stream.P}. Error is |fuel.P - oxy.P|/oxy.P = {perr}')
        self._pressure = float(self._oxy_stream.P)
        self._mechanism = fs.mech_spec
        self._n_species = self._mechanism.n_species
        self._n_reactions = self._mechanism.n_reactions
        self._n_equations = self._n_species
        self._state_fuel = np.hstack([self._fuel_stream.T, self._fuel_stream.Y[:-1]])
        self._state_oxy = np.hstack([self._oxy_stream.T, self._oxy_stream.Y[:-1]])

        # grid (flamelet.py:387-419)
        if fs.grid is not None:
            self._z = np.array(fs.grid, dtype=np.float64)
            self._dz = self._z[1:] - self._z[:-1]
        else:
            if fs.grid_points is None:
                raise ValueError('Flamelet specifications: one of either grid or grid_points must be given.')
            if fs.grid_type == 'uniform':
                self._z, self._dz = self._uniform_grid(fs.grid_points)
            elif fs.grid_type == 'clustered':
                zc = fs.grid_cluster_point
                if zc == 'stoichiometric':
                    zc = self._mechanism.stoich_mixture_fraction(self._fuel_stream, self._oxy_stream)
                self._z, self._dz = self._clustered_grid(fs.grid_points, zc, fs.grid_cluster_intensity)
            else:
                raise ValueError('Flamelet specifications: Bad grid_type argument detected: ' + str(fs.grid_type) +
                                 '\n                         Acceptable values: ' + str(self._grid_types))
        self._nz_interior = self._z.size - 2
        self._n_dof = self._n_equations * self._nz_interior

        # dissipation rate (flamelet.py:424-460)
        if fs.dissipation_rate is not None:
            self._x = np.array(fs.dissipation_rate, dtype=np.float64)
            self._max_dissipation_rate = np.max(self._x)
            self._dissipation_rate_form = 'custom'
        elif fs.dissipation_rate_form not in ['peters', 'Peters', 'constant'] or \
                (fs.max_dissipation_rate is None and fs.stoich_dissipation_rate is None):
            self._x = np.zeros_like(self._z)
            self._max_dissipation_rate = 0.
            self._dissipation_rate_form = 'unspecified-set-to-0'
        else:
            self._dissipation_rate_form = fs.dissipation_rate_form
            if fs.max_dissipation_rate is not None:
                self._max_dissipation_rate = fs.max_dissipation_rate
            elif fs.dissipation_rate_form in ['peters', 'Peters']:
                z_st = self._mechanism.stoich_mixture_fraction(self._fuel_stream, self._oxy_stream)
                self._max_dissipation_rate = fs.stoich_dissipation_rate / np.exp(-2. * (erfinv(2. * z_st - 1.)) ** 2)
            else:
                self._max_dissipation_rate = fs.stoich_dissipation_rate
            self._x = compute_dissipation_rate(self._z, self._max_dissipation_rate, self._dissipation_rate_form)
        self._lewis_numbers = np.ones(self._n_species)

        # heat transfer (flamelet.py:463-519)
        if fs.heat_transfer not in self._heat_transfers:
            raise ValueError('Flamelet specifications: Bad heat_transfer argument detected: ' + str(fs.heat_transfer) +
                             '\n                         Acceptable values: ' + str(self._heat_transfers))
        self._heat_transfer = fs.heat_transfer
        self._T_conv = self._T_rad = self._h_conv = self._h_rad = None
        self._scale_heat_loss_by_temp_range = fs.scale_heat_loss_by_temp_range
        self._scale_convection_by_dissipation = fs.scale_convection_by_dissipation
        self._use_linear_ref_temp_profile = fs.use_linear_ref_temp_profile
        if self._heat_transfer != 'adiabatic':
            self._heat_array(fs.convection_coefficient, 'convection_coefficient', '_h_conv')
            self._heat_array(fs.radiative_emissivity, 'radiative_emissivity', '_h_rad')
            if self._use_linear_ref_temp_profile:
                self._T_conv = self._oxy_stream.T + self._z[1:-1] * (self._fuel_stream.T - self._oxy_stream.T)
                self._T_rad = self._T_conv.copy()
            else:
                self._heat_array(fs.radiation_temperature, 'radiation_temperature', '_T_rad')
                self._heat_array(fs.convection_temperature, 'convection_temperature', '_T_conv')
            if self._scale_convection_by_dissipation:
                zst = self._mechanism.stoich_mixture_fraction(self._fuel_stream, self._oxy_stream)
                factor = np.max(self._x) / (1. - zst) / zst
                self._h_conv = self._h_conv * factor
                self._h_rad = self._h_rad * factor

        # initial condition (flamelet.py:522-600)
        ic = fs.initial_condition
        if isinstance(ic, str):
            self._initial_state = self._initial_state_from_name(ic)
        elif isinstance(ic, np.ndarray):
            if ic.size != self._n_dof:
                raise ValueError('size of initial condition is incorrect!')
            self._initial_state = np.array(ic, dtype=np.float64).ravel()
        else:
            raise ValueError('Flamelet specifications: bad argument for initial_condition\n'
                             '                         must be either another Flamelet instance or a string\n'
                             '                         allowable strings: ' + str(self._initializations))
        self._current_state = np.copy(self._initial_state)
        self._initial_time = 0.
        self._current_time = 0.

        self._griffon = self._mechanism.griffon
        self._include_enthalpy_flux = fs.include_enthalpy_flux
        self._include_variable_cp = fs.include_variable_cp
        self._rsopt = self._rates_sensitivity_option_dict[fs.rates_sensitivity_type]
        self._stopt = self._sensitivity_transform_option_dict[fs.sensitivity_transform_type]
        self._variable_scales = np.ones(self._n_dof)
        self._variable_scales[::self._n_equations] = 1.e3
        self._solution_times = []
        self._maj_coeff_griffon = np.zeros(self._n_dof)
        self._sub_coeff_griffon = np.zeros(self._n_dof)
        self._sup_coeff_griffon = np.zeros(self._n_dof)
        self._mcoeff_griffon = np.zeros(self._nz_interior)
        self._ncoeff_griffon = np.zeros(self._nz_interior)
        self._griffon.flamelet_stencils(self._dz, self._nz_interior, self._x, 1. / self._lewis_numbers,
                                        self._maj_coeff_griffon, self._sub_coeff_griffon, self._sup_coeff_griffon,
                                        self._mcoeff_griffon, self._ncoeff_griffon)
        self._jac_nelements_griffon = int(self._n_equations * (self._nz_interior * self._n_equations +
                                                               2 * (self._nz_interior - 1)))
        self._iteration_count = None
        self._batch = None
        self._factors = None

    def _initial_state_from_name(self, name):
        m, oxy, fuel = self._mechanism, self._oxy_stream, self._fuel_stream
        nzi, neq = self._nz_interior, self._n_equations
        state = np.zeros((nzi, neq))
        mix = lambda z: m.mix_streams([(m.copy_stream(oxy), 1 - z), (m.copy_stream(fuel), z)], 'mass', 'HP')
        if name == 'unreacted':
            for i in range(nzi):
                q = mix(self._z[1 + i])
                state[i, :] = np.hstack((q.T, q.Y[:-1]))
        elif name == 'linear-TY':
            z = self._z[1:-1]
            for i in range(neq):
                state[:, i] = self._state_oxy[i] + (self._state_fuel[i] - self._state_oxy[i]) * z
        elif name == 'equilibrium':
            from spitfire_b200.equilibrium import equilibrate_many
            qs = equilibrate_many([mix(self._z[1 + i]) for i in range(nzi)], 'HP')  # all grid points in one iteration
            for i, q in enumerate(qs):
                state[i, :] = np.hstack((q.T, q.Y[:-1]))
        elif name == 'Burke-Schumann':
            zst = m.stoich_mixture_fraction(fuel, oxy)
            stmix = mix(zst)
            stat = m._get_atoms_in_stream(stmix, ['H', 'C', 'O', 'N'])
            comp = 'N2: ' + str(stat['N'] / 2) + ' H2O: ' + str(0.5 * stat['H'])
            if 'CO2' in m.species_names:
                comp += ' CO2: ' + str(stat['C'])
            sw = m.stream('HPX', (stmix.H, self._pressure, comp))
            Y_st, Y_o, Y_f = sw.Y, oxy.Y, fuel.Y
            for i in range(nzi):
                z = self._z[1 + i]
                h_mix = mix(z).H
                Y_mix = Y_o + z / zst * (Y_st - Y_o) if z <= zst else Y_st + (z - zst) / (1. - zst) * (Y_f - Y_st)
                q = m.stream('HPY', (h_mix, self._pressure, Y_mix))
                state[i, :] = np.hstack((q.T, q.Y[:-1]))
        else:
            raise ValueError('Flamelet specifications: bad string argument for initial_condition\n'
                             '                         given: ' + name + '\n'
                             '                         allowable: ' + str(self._initializations))
        return state.ravel()

    # -- read-only views (flamelet.py:918-1044) ----------------------------------------------------------------------
    mechanism = property(lambda self: self._mechanism)
    dissipation_rate = property(lambda self: self._x)
    mixfrac_grid = property(lambda self: self._z)
    oxy_stream = property(lambda self: self._oxy_stream)
    fuel_stream = property(lambda self: self._fuel_stream)
    pressure = property(lambda self: self._pressure)
    linear_temperature = property(lambda self: self._oxy_stream.T + (self._fuel_stream.T - self._oxy_stream.T) * self._z)
    iteration_count = property(lambda self: self._iteration_count)
    initial_interior_state = property(lambda self: self._initial_state)
    current_interior_state = property(lambda self: self._current_state)
    initial_state = property(lambda self: np.hstack((self._state_oxy, self._initial_state, self._state_fuel)))
    current_state = property(lambda self: np.hstack((self._state_oxy, self._current_state, self._state_fuel)))
    initial_temperature = property(lambda self: self.initial_state[::self._n_equations])
    current_temperature = property(lambda self: self.current_state[::self._n_equations])

    def _ops(self):
        if self._batch is None:
            self._batch = FlameletBatch([self])
        return self._batch.ops

    def _t(self, a):
        ops = self._ops()
        return ops.torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64).reshape(1, -1)).to(ops.device)

    # -- callables handed to the time integrator (flamelet.py:642-915) ---------------------------------------------------
    def _rhs(self, t, state_interior):
        return self._ops().rhs(self._t(state_interior)).cpu().numpy().ravel()

    _adiabatic_rhs = _rhs
    _nonadiabatic_rhs = _rhs

    def _jac(self, state_interior):
        return self._ops().jac(self._t(state_interior)).cpu().numpy().ravel()

    _adiabatic_jac = _jac
    _nonadiabatic_jac = _jac

    def _setup_block_thomas(self, t, state_interior, prefactor):
        ops = self._ops()
        J = ops.jac(self._t(state_interior), scale_and_offset=True, prefactor=float(prefactor))
        self._factors = ops.factorize(J)

    _adiabatic_setup_block_thomas = _setup_block_thomas
    _nonadiabatic_setup_block_thomas = _setup_block_thomas

    def _solve_block_thomas(self, residual):
        x = self._ops().solve(self._factors, self._t(residual))
        return x.cpu().numpy().ravel(), 1, True

    # -- time integration (flamelet.py:1046-1286) ---------------------------------------------------------------------------
    def integrate(self, stop_at_time=None, stop_at_steady=None, stop_criteria=None, first_time_step=1.e-6,
                  max_time_step=1.e-3, minimum_time_step_count=40, transient_tolerance=1.e-10, write_log=False,
                  log_rate=100, maximum_steps_per_jacobian=10, nonlinear_solve_tolerance=1.e-12,
                  linear_solver='block thomas', stepper_type=KennedyCarpenterS6P4Q3, nlsolver_type=SimpleNewtonSolver,
                  stepcontrol_type=PIController, extra_integrator_args=dict(), extra_stepper_args=dict(),
                  extra_nlsolver_args=dict(), extra_stepcontrol_args=dict(), save_first_and_last_only=False,
                  print_exception_on_failure=False):
        """Base method for flamelet integration; same arguments as the reference. Returns a (time, mixture_fraction)
        library of temperature, pressure and mass fractions."""
        if linear_solver.lower() != 'block thomas':
            raise ValueError('linear solver ' + linear_solver + ' is invalid here, only \'block thomas\' is provided')

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
                          step_size=stepcontrol_type(**sc), method=stepper_type(**st),
                          linear_setup=self._setup_block_thomas, linear_solve=self._solve_block_thomas,
                          minimum_time_step_count=minimum_time_step_count, linear_setup_rate=maximum_steps_per_jacobian,
                          verbose=write_log, log_rate=log_rate, norm_weighting=1. / self._variable_scales,
                          post_step_callback=post_step_callback, save_each_step=not save_first_and_last_only,
                          print_exception_on_failure=print_exception_on_failure, **iargs)
        if save_first_and_last_only:
            state, time, _ = output
            self._current_state, self._current_time = np.copy(state), np.copy(time)
            states, t = np.array(state).reshape(1, -1), np.array([time], dtype=np.float64).ravel()
        else:
            t, states = output
            self._current_state, self._current_time = np.copy(states[-1, :]), np.copy(t[-1])
        lib = Library(Dimension('time', t), Dimension('mixture_fraction', self._z))
        lib.extra_attributes['mech_spec'] = self._mechanism
        self._fill_library(lib, states, lead=True)
        return lib

    def integrate_to_steady(self, steady_tolerance=1.e-4, **kwargs):
        return self.integrate(stop_at_steady=steady_tolerance, **kwargs)

    def integrate_to_time(self, final_time, **kwargs):
        return self.integrate(stop_at_time=final_time, **kwargs)

    def _check_ignition_delay(self, state, delta_temperature_ignition):
        ne = self._n_equations
        return np.max(state[::ne] - self._initial_state[::ne]) > delta_temperature_ignition

    def integrate_to_steady_after_ignition(self, steady_tolerance=1.e-4, delta_temperature_ignition=400., **kwargs):
        def stop(t, state, residual, *args, **kw):
            return self._check_ignition_delay(state, delta_temperature_ignition) and residual < steady_tolerance

        return self.integrate(stop_criteria=stop, **kwargs)

    def integrate_for_heat_loss(self, temperature_tolerance=0.05, steady_tolerance=1.e-4, **kwargs):
        """Integrate until the temperature profile is nearly linear or steady: the heat-loss dimension of nonadiabatic
        libraries (flamelet.py:1263-1286)"""
        T_bc_max = max([self._oxy_stream.T, self._fuel_stream.T])

        def stop(t, state, residual, *args, **kw):
            return np.max(state) < (1. + temperature_tolerance) * T_bc_max or residual < steady_tolerance

        return self.integrate(stop_criteria=stop, **kwargs)

    def compute_ignition_delay(self, delta_temperature_ignition=400., minimum_allowable_residual=1.e-12,
                               return_solution=False, **kwargs):
        def stop(t, state, residual, *args, **kw):
            if residual > minimum_allowable_residual:
                return self._check_ignition_delay(state, delta_temperature_ignition)
            raise ValueError(f'From compute_ignition_delay(): residual < minimum allowable value '
                             f'({minimum_allowable_residual}), suggesting that the reactor will not ignite.')

        lib = self.integrate(stop_criteria=stop, save_first_and_last_only=not return_solution, **kwargs)
        tau = lib.time_values[-1]
        return (tau, lib) if return_solution else tau

    # -- libraries ------------------------------------------------------------------------------------------------------------
    def _fill_library(self, lib, states, lead):
        """temperature, pressure, mass fractions with the boundary streams attached; states [nt, ndof] if lead else
        [ndof]"""
        ne, names = self._n_equations, self._mechanism.species_names
        sl = (lambda a: (slice(None), a)) if lead else (lambda a: a)
        S = states if lead else states.reshape(-1)
        if lead and S.shape[0] > 8:
            # (one transposing pass, then every variable is a contiguous [nt, nzi] block: a saved trajectory is a few MB,
            # which 53 strided sweeps would each pull through the cache again)
            St = np.ascontiguousarray(S.reshape(S.shape[0], -1, ne).transpose(2, 0, 1))
            pick = lambda off: St[off]
        else:
            pick = (lambda off: S[:, off::ne]) if lead else (lambda off: S[off::ne])
        T = lib.get_empty_dataset()
        T[sl(0)], T[sl(slice(1, -1))], T[sl(-1)] = self._oxy_stream.T, pick(0), self._fuel_stream.T
        lib['temperature'] = T
        lib['pressure'] = np.zeros_like(T) + self._pressure
        last = np.ones_like(T)
        lib['mass fraction ' + names[-1]] = last
        for i, s in enumerate(names[:-1]):
            Y = lib.get_empty_dataset()
            Y[sl(0)], Y[sl(slice(1, -1))], Y[sl(-1)] = self._oxy_stream.Y[i], pick(1 + i), self._fuel_stream.Y[i]
            lib['mass fraction ' + s] = Y
            last = last - Y
        lib['mass fraction ' + names[-1]] = last

    def make_library_from_interior_state(self, state_in):
        lib = Library(Dimension('mixture_fraction', self._z))
        lib.extra_attributes['mech_spec'] = self._mechanism
        self._fill_library(lib, np.asarray(state_in, dtype=np.float64), lead=False)
        return lib

    # -- steady solvers: a batch of one -------------------------------------------------------------------------------------------
    def _one(self):
        if self._batch is None:
            self._batch = FlameletBatch([self])
        return self._batch

    def steady_solve_newton(self, initial_guess=None, **kwargs):
        """Newton's method with a chord Jacobian and residual line search (flamelet.py:1352-1485). Returns
        (library, iteration count, converged) or (None, None, False)."""
        guess = None if initial_guess is None else np.asarray(initial_guess).reshape(1, -1)
        states, iters, conv = self._one().steady_solve_newton(initial_guess=guess, **kwargs)
        if not conv[0]:
            return None, None, False
        self._current_state = np.copy(states[0])
        self._iteration_count = int(iters[0])
        return self.make_library_from_interior_state(states[0]), int(iters[0]), True

    def steady_solve_psitc(self, initial_guess=None, **kwargs):
        """Adaptive pseudo-transient continuation (flamelet.py:1487-1703). Returns (library, iteration count,
        converged, min(ds))."""
        kwargs.pop('_recursion_depth', None)
        guess = None if initial_guess is None else np.asarray(initial_guess).reshape(1, -1)
        states, iters, conv, mds = self._one().steady_solve_psitc(initial_guess=guess, **kwargs)
        if not conv[0]:
            return None, None, False, float(mds[0])
        self._current_state = np.copy(states[0])
        self._iteration_count = int(iters[0])
        return self.make_library_from_interior_state(states[0]), int(iters[0]), True, float(mds[0])

    def compute_steady_state(self, tolerance=1.e-6, verbose=False, use_psitc=True, newton_args=None, psitc_args=None,
                             transient_args=None):
        """Newton, then pseudo-transient continuation, then ESDIRK64 (flamelet.py:1705-1778)"""
        states, _ = self._one().compute_steady_state(tolerance=tolerance, verbose=verbose, use_psitc=use_psitc,
                                                     newton_args=newton_args, psitc_args=psitc_args,
                                                     transient_args=transient_args)
        self._current_state = np.copy(states[0])
        return self.make_library_from_interior_state(states[0])
