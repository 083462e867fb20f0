"""
ctypes plumbing shared by the product binding (spitfire_b200.griffon) and, in tests only, by the oracle binding.

The mechanism setters keep the exact names and positional argument orders of the reference's Cython class
(reference: src/spitfire/griffon/griffon.pyx:234-551) and marshal them onto the flattened C-ABI declared in
include/griffon_b200.h (`<prefix>mech_*`).
"""
import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_char_pp = C.POINTER(C.c_char_p)


def dptr(a):
    """pointer to a float64 C-contiguous ndarray (no copy; raises if the array is not usable in place)"""
    if a is None:
        return None
    if not isinstance(a, np.ndarray) or a.dtype != np.float64 or not a.flags['C_CONTIGUOUS']:
        raise TypeError('expected a C-contiguous float64 ndarray, as the reference binding does')
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    if a is None:
        return None
    if not isinstance(a, np.ndarray) or a.dtype != np.int32 or not a.flags['C_CONTIGUOUS']:
        raise TypeError('expected a C-contiguous int32 ndarray, as the reference binding does')
    return a.ctypes.data_as(c_int_p)


def _names(keys):
    arr = (C.c_char_p * max(len(keys), 1))()
    for i, k in enumerate(keys):
        arr[i] = str(k).encode()
    return arr


def _doubles(vals):
    arr = (C.c_double * max(len(vals), 1))()
    for i, v in enumerate(vals):
        arr[i] = float(v)
    return arr


def _ints(vals):
    arr = (C.c_int * max(len(vals), 1))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr


def declare_mech_abi(lib, prefix):
    """attach argtypes/restype for the `<prefix>mech_*` functions"""
    P = C.c_void_p
    f = getattr(lib, prefix + 'mech_create')
    f.restype, f.argtypes = P, []
    f = getattr(lib, prefix + 'mech_destroy')
    f.restype, f.argtypes = None, [P]
    for name in ('set_ref_pressure', 'set_ref_temperature', 'set_gas_constant'):
        f = getattr(lib, prefix + 'mech_' + name)
        f.restype, f.argtypes = C.c_int, [P, C.c_double]
    f = getattr(lib, prefix + 'mech_set_element_mw')
    f.restype, f.argtypes = C.c_int, [P, C.c_char_p, C.c_double]
    f = getattr(lib, prefix + 'mech_add_element')
    f.restype, f.argtypes = C.c_int, [P, C.c_char_p]
    f = getattr(lib, prefix + 'mech_add_species')
    f.restype, f.argtypes = C.c_int, [P, C.c_char_p, C.c_int, c_char_pp, c_double_p]
    f = getattr(lib, prefix + 'mech_resize_heat_capacity_data')
    f.restype, f.argtypes = C.c_int, [P]
    f = getattr(lib, prefix + 'mech_add_const_cp')
    f.restype, f.argtypes = C.c_int, [P, C.c_char_p] + [C.c_double] * 6
    f = getattr(lib, prefix + 'mech_add_nasa7_cp')
    f.restype, f.argtypes = C.c_int, [P, C.c_char_p, C.c_double, C.c_double, C.c_double, c_double_p, c_double_p]
    f = getattr(lib, prefix + 'mech_add_nasa9_cp')
    f.restype, f.argtypes = C.c_int, [P, C.c_char_p, C.c_double, C.c_double, C.c_int, c_double_p]
    f = getattr(lib, prefix + 'mech_add_reaction')
    f.restype = C.c_int
    f.argtypes = [P, C.c_int, C.c_int,
                  C.c_int, c_char_pp, c_int_p,
                  C.c_int, c_char_pp, c_int_p,
                  C.c_double, C.c_double, C.c_double,
                  C.c_int, c_char_pp, c_double_p, C.c_double,
                  C.c_double, C.c_double, C.c_double, c_double_p,
                  C.c_int, c_char_pp, c_double_p]
    for name in ('n_species', 'n_reactions'):
        f = getattr(lib, prefix + 'mech_' + name)
        f.restype, f.argtypes = C.c_int, [P]
    f = getattr(lib, prefix + 'mech_molecular_weights')
    f.restype, f.argtypes = C.c_int, [P, c_double_p]


class MechanismSetters(object):
    """The 18 `mechanism_*` setters of PyCombustionKernels (griffon.pyx:234-551), over `<prefix>mech_*`.

    Subclasses provide `self._lib`, `self._prefix`, `self._h` (the opaque handle) and `_check(rc, what)`.
    """

    def _mech_call(self, name, *args):
        rc = getattr(self._lib, self._prefix + 'mech_' + name)(self._h, *args)
        self._check(rc, name)

    def mechanism_set_element_mw_map(self, element_mw_map):
        for a in element_mw_map:
            self._mech_call('set_element_mw', str(a).encode(), float(element_mw_map[a]))

    def mechanism_add_element(self, element_name):
        self._mech_call('add_element', element_name.encode())

    def mechanism_add_species(self, species_name, atom_map):
        keys = list(atom_map.keys())
        self._mech_call('add_species', species_name.encode(), len(keys), _names(keys),
                        _doubles([atom_map[k] for k in keys]))

    def mechanism_set_ref_pressure(self, p):
        self._mech_call('set_ref_pressure', float(p))

    def mechanism_set_ref_temperature(self, T):
        self._mech_call('set_ref_temperature', float(T))

    def mechanism_set_gas_constant(self, Ru):
        self._mech_call('set_gas_constant', float(Ru))

    def mechanism_resize_heat_capacity_data(self):
        self._mech_call('resize_heat_capacity_data')

    def mechanism_add_const_cp(self, spec_name, Tmin, Tmax, T0, h0, s0, cp):
        self._mech_call('add_const_cp', spec_name.encode(), float(Tmin), float(Tmax), float(T0), float(h0),
                        float(s0), float(cp))

    def mechanism_add_nasa7_cp(self, spec_name, Tmin, Tmid, Tmax, low_coeffs, high_coeffs):
        if len(low_coeffs) != 7 or len(high_coeffs) != 7:
            raise ValueError('NASA7 polynomials need 7 low and 7 high coefficients')
        self._mech_call('add_nasa7_cp', spec_name.encode(), float(Tmin), float(Tmid), float(Tmax),
                        _doubles(low_coeffs), _doubles(high_coeffs))

    def mechanism_add_nasa9_cp(self, spec_name, Tmin, Tmax, coeffs):
        self._mech_call('add_nasa9_cp', spec_name.encode(), float(Tmin), float(Tmax), len(coeffs), _doubles(coeffs))

    def _add_reaction(self, rtype, reactants_stoich, products_stoich, reversible, A, b, Ea,
                      efficiencies=None, default_eff=1.0, flf=(0., 0., 0.), troe=None, orders=None):
        rk = list(reactants_stoich.keys())
        pk = list(products_stoich.keys())
        ek = list(efficiencies.keys()) if efficiencies else []
        ok = list(orders.keys()) if orders else []
        troe4 = [0., 0., 0., 0.]
        if troe is not None:
            for i, v in enumerate(troe):  # zero padded to [A, T3, T1, T2] as griffon.pyx:410-416
                troe4[i] = v
        self._mech_call('add_reaction', int(rtype), 1 if reversible else 0,
                        len(rk), _names(rk), _ints([reactants_stoich[k] for k in rk]),
                        len(pk), _names(pk), _ints([products_stoich[k] for k in pk]),
                        float(A), float(b), float(Ea),
                        len(ek), _names(ek), _doubles([efficiencies[k] for k in ek]), float(default_eff),
                        float(flf[0]), float(flf[1]), float(flf[2]), _doubles(troe4),
                        len(ok), _names(ok), _doubles([orders[k] for k in ok]))

    def mechanism_add_reaction_simple(self, reactants_stoich, products_stoich, reversible, fwd_pre_exp_value,
                                      fwd_temp_exponent, fwd_act_energy):
        self._add_reaction(1, reactants_stoich, products_stoich, reversible, fwd_pre_exp_value, fwd_temp_exponent,
                           fwd_act_energy)

    def mechanism_add_reaction_three_body(self, reactants_stoich, products_stoich, reversible, fwd_pre_exp_value,
                                          fwd_temp_exponent, fwd_act_energy, three_body_efficiencies,
                                          default_efficiency):
        self._add_reaction(2, reactants_stoich, products_stoich, reversible, fwd_pre_exp_value, fwd_temp_exponent,
                           fwd_act_energy, three_body_efficiencies, default_efficiency)

    def mechanism_add_reaction_Lindemann(self, reactants_stoich, products_stoich, reversible, fwd_pre_exp_value,
                                         fwd_temp_exponent, fwd_act_energy, three_body_efficiencies,
                                         default_efficiency, flf_pre_exp_value, flf_temp_exponent, flf_act_energy):
        self._add_reaction(3, reactants_stoich, products_stoich, reversible, fwd_pre_exp_value, fwd_temp_exponent,
                           fwd_act_energy, three_body_efficiencies, default_efficiency,
                           (flf_pre_exp_value, flf_temp_exponent, flf_act_energy))

    def mechanism_add_reaction_Troe(self, reactants_stoich, products_stoich, reversible, fwd_pre_exp_value,
                                    fwd_temp_exponent, fwd_act_energy, three_body_efficiencies, default_efficiency,
                                    flf_pre_exp_value, flf_temp_exponent, flf_act_energy, troe_parameters):
        self._add_reaction(4, reactants_stoich, products_stoich, reversible, fwd_pre_exp_value, fwd_temp_exponent,
                           fwd_act_energy, three_body_efficiencies, default_efficiency,
                           (flf_pre_exp_value, flf_temp_exponent, flf_act_energy), troe_parameters)

    def mechanism_add_reaction_simple_with_special_orders(self, reactants_stoich, products_stoich, reversible,
                                                          fwd_pre_exp_value, fwd_temp_exponent, fwd_act_energy,
                                                          special_orders):
        self._add_reaction(1, reactants_stoich, products_stoich, reversible, fwd_pre_exp_value, fwd_temp_exponent,
                           fwd_act_energy, orders=special_orders)

    def mechanism_add_reaction_three_body_with_special_orders(self, reactants_stoich, products_stoich, reversible,
                                                              fwd_pre_exp_value, fwd_temp_exponent, fwd_act_energy,
                                                              three_body_efficiencies, default_efficiency,
                                                              special_orders):
        self._add_reaction(2, reactants_stoich, products_stoich, reversible, fwd_pre_exp_value, fwd_temp_exponent,
                           fwd_act_energy, three_body_efficiencies, default_efficiency, orders=special_orders)

    def mechanism_add_reaction_Lindemann_with_special_orders(self, reactants_stoich, products_stoich, reversible,
                                                             fwd_pre_exp_value, fwd_temp_exponent, fwd_act_energy,
                                                             three_body_efficiencies, default_efficiency,
                                                             flf_pre_exp_value, flf_temp_exponent, flf_act_energy,
                                                             special_orders):
        self._add_reaction(3, reactants_stoich, products_stoich, reversible, fwd_pre_exp_value, fwd_temp_exponent,
                           fwd_act_energy, three_body_efficiencies, default_efficiency,
                           (flf_pre_exp_value, flf_temp_exponent, flf_act_energy), orders=special_orders)

    def mechanism_add_reaction_Troe_with_special_orders(self, reactants_stoich, products_stoich, reversible,
                                                        fwd_pre_exp_value, fwd_temp_exponent, fwd_act_energy,
                                                        three_body_efficiencies, default_efficiency,
                                                        flf_pre_exp_value, flf_temp_exponent, flf_act_energy,
                                                        troe_parameters, special_orders):
        self._add_reaction(4, reactants_stoich, products_stoich, reversible, fwd_pre_exp_value, fwd_temp_exponent,
                           fwd_act_energy, three_body_efficiencies, default_efficiency,
                           (flf_pre_exp_value, flf_temp_exponent, flf_act_energy), troe_parameters,
                           orders=special_orders)

    # introspection (not in the reference class; cheap and useful)
    @property
    def n_species(self):
        return getattr(self._lib, self._prefix + 'mech_n_species')(self._h)

    @property
    def n_reactions(self):
        return getattr(self._lib, self._prefix + 'mech_n_reactions')(self._h)

    @property
    def molecular_weights(self):
        out = np.zeros(self.n_species)
        getattr(self._lib, self._prefix + 'mech_molecular_weights')(self._h, dptr(out))
        return out
