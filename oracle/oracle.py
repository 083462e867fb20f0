"""
oracle.py -- TEST INFRASTRUCTURE. ctypes binding of the two CPU checkers (see oracle/griffon_oracle.h).

    OracleKernels('port')       -> oracle/liboracle_port.so      (plain-C restatement)
    OracleKernels('reference')  -> oracle/_ref/libref_griffon.so (the unmodified reference C++, when built)

The class exposes the same method names / positional argument orders as the reference's
`PyCombustionKernels` (src/spitfire/griffon/griffon.pyx:220-987) so the parity tests read like the reference's
own tests. The product (spitfire_b200/) never imports this module.
"""
import ctypes as C
import os

import numpy as np

from spitfire_b200._cabi import MechanismSetters, declare_mech_abi, dptr, iptr, c_double_p, c_int_p

HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def lib_path(kind):
    return {'port': os.path.join(HERE, 'liboracle_port.so'),
            'reference': os.path.join(HERE, '_ref', 'libref_griffon.so')}[kind]


def available(kind):
    return os.path.exists(lib_path(kind))


def _load(kind):
    if kind in _LIBS:
        return _LIBS[kind]
    lib = C.CDLL(lib_path(kind))
    declare_mech_abi(lib, 'go_')
    P, D, I = C.c_void_p, C.c_double, C.c_int
    dp, ip = c_double_p, c_int_p
    lib.go_kind.restype = C.c_char_p

    def sig(name, restype, argtypes):
        f = getattr(lib, name)
        f.restype, f.argtypes = restype, argtypes

    sig('go_mixture_molecular_weight', D, [P, dp])
    sig('go_mole_fractions', None, [P, dp, dp])
    sig('go_ideal_gas_density', D, [P, D, D, dp])
    sig('go_ideal_gas_pressure', D, [P, D, D, dp])
    for n in ('cp_mix', 'cv_mix', 'enthalpy_mix', 'energy_mix'):
        sig('go_' + n, D, [P, D, dp])
    for n in ('species_cp', 'species_cv', 'species_enthalpies', 'species_energies'):
        sig('go_' + n, None, [P, D, dp])
    sig('go_cp_sens_T', None, [P, D, dp, dp, dp])
    sig('go_production_rates', None, [P, D, D, dp, dp])
    sig('go_prod_rates_primitive_sensitivities', None, [P, D, D, dp, I, dp])
    sig('go_reactor_rhs_isobaric', None, [P, dp, D, D, dp, D, D, D, D, D, D, I, I, dp])
    sig('go_reactor_jac_isobaric', None, [P, dp, D, D, dp, D, D, D, D, D, D, I, I, I, I, dp, dp])
    sig('go_reactor_jac_isobaric_many', None, [P, I, dp, D, I, dp, dp])
    sig('go_reactor_rhs_isochoric', None, [P, dp, D, D, dp, D, D, D, D, D, D, I, I, dp])
    sig('go_reactor_jac_isochoric', None, [P, dp, D, D, dp, D, D, D, D, D, D, I, I, I, dp, dp])
    sig('go_reactor_rhs_isobaric_many', None, [P, I, dp, D, dp])
    sig('go_flamelet_stencils', None, [P, dp, I, dp, dp, dp, dp, dp, dp, dp])
    sig('go_flamelet_jac_indices', None, [P, I, ip, ip])
    sig('go_flamelet_rhs', None, [P, dp, D, dp, dp, I, dp, dp, dp, dp, I, dp, dp, dp, dp, dp, dp, I, I, I, dp])
    sig('go_flamelet_jacobian', None,
        [P, dp, D, dp, dp, I, dp, dp, dp, dp, I, dp, dp, dp, dp, dp, dp, I, D, I, D, I, I, I, I, I, dp, dp])
    sig('go_btddod_full_factorize', None, [dp, I, I, dp, ip])
    sig('go_btddod_full_solve', None, [dp, dp, ip, dp, I, I, dp])
    sig('go_btddod_full_matvec', None, [dp, dp, I, I, dp])
    sig('go_btddod_scale_and_add_diagonal', None, [dp, D, dp, D, I, I])
    _LIBS[kind] = lib
    return lib


class OracleKernels(MechanismSetters):
    _prefix = 'go_'

    def __init__(self, kind='port'):
        self.kind = kind
        self._lib = _load(kind)
        self._h = C.c_void_p(self._lib.go_mech_create())

    def __del__(self):
        try:
            if self._h:
                self._lib.go_mech_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f'oracle({self.kind}) {what} failed with code {rc}')

    # thermodynamics -- griffon.pyx:684-758
    def mixture_molecular_weight(self, y):
        return self._lib.go_mixture_molecular_weight(self._h, dptr(y))

    def mole_fractions(self, y, x):
        self._lib.go_mole_fractions(self._h, dptr(y), dptr(x))

    def ideal_gas_density(self, p, T, y):
        return self._lib.go_ideal_gas_density(self._h, p, T, dptr(y))

    def ideal_gas_pressure(self, rho, T, y):
        return self._lib.go_ideal_gas_pressure(self._h, rho, T, dptr(y))

    def cp_mix(self, T, y):
        return self._lib.go_cp_mix(self._h, T, dptr(y))

    def cv_mix(self, T, y):
        return self._lib.go_cv_mix(self._h, T, dptr(y))

    def enthalpy_mix(self, T, y):
        return self._lib.go_enthalpy_mix(self._h, T, dptr(y))

    def energy_mix(self, T, y):
        return self._lib.go_energy_mix(self._h, T, dptr(y))

    def species_cp(self, T, out):
        self._lib.go_species_cp(self._h, T, dptr(out))

    def species_cv(self, T, out):
        self._lib.go_species_cv(self._h, T, dptr(out))

    def species_enthalpies(self, T, out):
        self._lib.go_species_enthalpies(self._h, T, dptr(out))

    def species_energies(self, T, out):
        self._lib.go_species_energies(self._h, T, dptr(out))

    def dcpdT_species(self, T, y, out):
        mix = C.c_double(0.)
        self._lib.go_cp_sens_T(self._h, T, dptr(y), C.cast(C.byref(mix), c_double_p), dptr(out))

    # kinetics -- griffon.pyx:763-783 (first two positional arguments are (T, rho) resp. (rho, T))
    def production_rates(self, T, rho, y, out_w):
        self._lib.go_production_rates(self._h, T, rho, dptr(y), dptr(out_w))

    def prod_rates_primitive_sensitivities(self, rho, T, y, option, out):
        self._lib.go_prod_rates_primitive_sensitivities(self._h, rho, T, dptr(y), option, dptr(out))

    # reactors -- griffon.pyx:788-824
    def reactor_rhs_isobaric(self, state, p, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open_,
                             out_rhs):
        self._lib.go_reactor_rhs_isobaric(self._h, dptr(state), p, T_in, dptr(y_in), tau, T_inf, T_surf, h_conv,
                                          eps_rad, SoV, int(heat_option), int(bool(open_)), dptr(out_rhs))

    def reactor_jac_isobaric(self, state, p, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open_,
                             rates_sens_option, sens_transform_option, out_rhs, out_jac):
        self._lib.go_reactor_jac_isobaric(self._h, dptr(state), p, T_in, dptr(y_in), tau, T_inf, T_surf, h_conv,
                                          eps_rad, SoV, int(heat_option), int(bool(open_)), int(rates_sens_option),
                                          int(sens_transform_option), dptr(out_rhs), dptr(out_jac))

    def reactor_jac_isobaric_many(self, state, p, rates_sens_option, out_rhs, out_jac):
        """plain C loop of the single-state call over state[n, ns] (ctypes releases the GIL: thread-parallel)"""
        self._lib.go_reactor_jac_isobaric_many(self._h, state.shape[0], dptr(state), p, int(rates_sens_option),
                                               dptr(out_rhs), dptr(out_jac))

    def reactor_rhs_isobaric_many(self, state, p, out_rhs):
        self._lib.go_reactor_rhs_isobaric_many(self._h, state.shape[0], dptr(state), p, dptr(out_rhs))

    # isochoric reactor -- griffon.pyx:831-866 (state [rho, T, Y_0..Y_{ns-2}])
    def reactor_rhs_isochoric(self, state, rho_in, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option,
                              open_, out_rhs):
        self._lib.go_reactor_rhs_isochoric(self._h, dptr(state), rho_in, T_in, dptr(y_in), tau, T_inf, T_surf, h_conv,
                                           eps_rad, SoV, int(heat_option), int(bool(open_)), dptr(out_rhs))

    def reactor_jac_isochoric(self, state, rho_in, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option,
                              open_, rates_sens_option, out_rhs, out_jac):
        self._lib.go_reactor_jac_isochoric(self._h, dptr(state), rho_in, T_in, dptr(y_in), tau, T_inf, T_surf, h_conv,
                                           eps_rad, SoV, int(heat_option), int(bool(open_)), int(rates_sens_option),
                                           dptr(out_rhs), dptr(out_jac))

    # flamelet -- griffon.pyx:556-679 (Python order: T_conv, T_rad, h_conv, h_rad)
    def flamelet_stencils(self, dz, nzi, chi, inv_lewis, out_cmajor, out_csub, out_csup, out_mcoeff, out_ncoeff):
        self._lib.go_flamelet_stencils(self._h, dptr(dz), nzi, dptr(chi), dptr(inv_lewis), dptr(out_cmajor),
                                       dptr(out_csub), dptr(out_csup), dptr(out_mcoeff), dptr(out_ncoeff))

    def flamelet_jac_indices(self, nzi, out_rows, out_cols):
        self._lib.go_flamelet_jac_indices(self._h, nzi, iptr(out_rows), iptr(out_cols))

    def flamelet_rhs(self, state, p, oxy, fuel, adiabatic, T_conv, T_rad, h_conv, h_rad, nzi, cmajor, csub, csup,
                     mcoeff, ncoeff, chi, include_enthalpy_flux, include_variable_cp, use_scaled_heat_loss, out_rhs):
        self._lib.go_flamelet_rhs(self._h, dptr(state), p, dptr(oxy), dptr(fuel), int(bool(adiabatic)), dptr(T_conv),
                                  dptr(h_conv), dptr(T_rad), dptr(h_rad), nzi, dptr(cmajor), dptr(csub), dptr(csup),
                                  dptr(mcoeff), dptr(ncoeff), dptr(chi), int(bool(include_enthalpy_flux)),
                                  int(bool(include_variable_cp)), int(bool(use_scaled_heat_loss)), dptr(out_rhs))

    def flamelet_jacobian(self, state, p, oxy, fuel, adiabatic, T_conv, T_rad, h_conv, h_rad, nzi, cmajor, csub, csup,
                          mcoeff, ncoeff, chi, compute_eigenvalues, diffterm, scale_and_offset, prefactor,
                          rates_sens_option, sens_transform_option, include_enthalpy_flux, include_variable_cp,
                          use_scaled_heat_loss, out_expeig, out_jac):
        self._lib.go_flamelet_jacobian(self._h, dptr(state), p, dptr(oxy), dptr(fuel), int(bool(adiabatic)),
                                       dptr(T_conv), dptr(h_conv), dptr(T_rad), dptr(h_rad), nzi, dptr(cmajor),
                                       dptr(csub), dptr(csup), dptr(mcoeff), dptr(ncoeff), dptr(chi),
                                       int(bool(compute_eigenvalues)), diffterm, int(bool(scale_and_offset)),
                                       prefactor, int(rates_sens_option), int(sens_transform_option),
                                       int(bool(include_enthalpy_flux)), int(bool(include_variable_cp)),
                                       int(bool(use_scaled_heat_loss)), dptr(out_expeig), dptr(out_jac))

    # BTDDOD -- griffon.pyx:1006-1113 (module-level functions in the reference)
    def btddod_full_factorize(self, d_factors, num_blocks, block_size, out_l_values, out_d_pivots):
        self._lib.go_btddod_full_factorize(dptr(d_factors), num_blocks, block_size, dptr(out_l_values),
                                           iptr(out_d_pivots))

    def btddod_full_solve(self, d_factors, l_values, d_pivots, rhs, num_blocks, block_size, out_solution):
        self._lib.go_btddod_full_solve(dptr(d_factors), dptr(l_values), iptr(d_pivots), dptr(rhs), num_blocks,
                                       block_size, dptr(out_solution))

    def btddod_full_matvec(self, matrix, vec, num_blocks, block_size, out):
        self._lib.go_btddod_full_matvec(dptr(matrix), dptr(vec), num_blocks, block_size, dptr(out))

    def btddod_scale_and_add_diagonal(self, matrix, a, diag, b, num_blocks, block_size):
        self._lib.go_btddod_scale_and_add_diagonal(dptr(matrix), a, dptr(diag), b, num_blocks, block_size)
