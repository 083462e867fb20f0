"""
Python binding of the B200 Griffon library -- the drop-in for `spitfire.griffon.griffon`.

`PyCombustionKernels` keeps the method names and positional argument orders of the reference's Cython class
(reference: src/spitfire/griffon/griffon.pyx:220-987); the `py_btddod_*` module functions mirror griffon.pyx:1006-1113.
Every single-state method is a batch-of-one call of the CUDA path (`gb_*_host` in include/griffon_b200.h): outputs are
preallocated by the caller and filled in place, exactly as in the reference. New `*_batch` methods expose the batched
device entry points; they accept torch CUDA tensors (zero-copy, asynchronous on the current torch stream) or numpy
arrays (staged through the library's device scratch, synchronous).

There is no CPU fallback: if the shared library or a CUDA device is missing, calls raise `GriffonB200Error`.
"""
import ctypes as C
import os

import numpy as np

from spitfire_b200._cabi import MechanismSetters, declare_mech_abi, dptr, iptr, c_double_p, c_int_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GRIFFON_B200_LIB', os.path.join(HERE, 'libgriffon_b200.so'))
_lib = None


class GriffonB200Error(RuntimeError):
    pass


class ReactorParams(C.Structure):
    """gb_reactor_params of include/griffon_b200.h"""
    _fields_ = [('pressure', C.c_double), ('inflow_temperature', C.c_double), ('inflow_y', c_double_p),
                ('tau', C.c_double), ('fluid_temperature', C.c_double), ('surf_temperature', C.c_double),
                ('h_conv', C.c_double), ('eps_rad', C.c_double), ('surface_area_over_volume', C.c_double),
                ('heat_transfer_option', C.c_int), ('open', C.c_int)]


class FlameletParams(C.Structure):
    """gb_flamelet_params of include/griffon_b200.h"""
    _fields_ = [('nzi', C.c_int), ('pressure', C.c_double), ('oxy_state', c_double_p), ('fuel_state', c_double_p),
                ('adiabatic', C.c_int), ('T_convection', c_double_p), ('h_convection', c_double_p),
                ('T_radiation', c_double_p), ('h_radiation', c_double_p), ('cmajor', c_double_p),
                ('csub', c_double_p), ('csup', c_double_p), ('mcoeff', c_double_p), ('ncoeff', c_double_p),
                ('chi', c_double_p), ('include_enthalpy_flux', C.c_int), ('include_variable_cp', C.c_int),
                ('use_scaled_heat_loss', C.c_int), ('stride_heat', C.c_long), ('stride_coeff', C.c_long),
                ('stride_mn', C.c_long), ('stride_chi', C.c_long)]


def load_library():
    """dlopen the in-tree CUDA library and declare the C-ABI; raises GriffonB200Error if it is not built"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GriffonB200Error(f'{LIB_PATH} is not built; run `python -m spitfire_b200.build` (needs nvcc). '
                               'The B200 Griffon path has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    declare_mech_abi(lib, 'gb_')
    P, D, I, L = C.c_void_p, C.c_double, C.c_int, C.c_long
    dp, ip = c_double_p, c_int_p
    V = C.c_void_p  # raw device/host address

    def sig(name, restype, argtypes):
        f = getattr(lib, name)
        f.restype, f.argtypes = restype, argtypes

    sig('gb_last_error', C.c_char_p, [])
    sig('gb_cuda_device_count', I, [])
    sig('gb_mech_commit', I, [P])
    sig('gb_kernel_launch_count', L, [])
    sig('gb_build_info', C.c_char_p, [])
    sig('gb_measure_fp64_peak', I, [I, dp, dp])
    sig('gb_measure_fp64_latency', I, [I, dp])
    sig('gb_thermo_batch', I, [P, I, I, V, V, V, V, V])
    sig('gb_thermo_host', I, [P, I, I, V, V, V, V])
    sig('gb_production_rates_batch', I, [P, I, V, V, V, V, V])
    sig('gb_production_rates_host', I, [P, I, V, V, V, V])
    sig('gb_prod_rates_sens_batch', I, [P, I, V, V, V, I, V, V])
    sig('gb_prod_rates_sens_host', I, [P, I, V, V, V, I, V])
    RP = C.POINTER(ReactorParams)
    sig('gb_reactor_rhs_isobaric_batch', I, [P, I, V, RP, V, V])
    sig('gb_reactor_rhs_isobaric_host', I, [P, I, V, RP, V])
    sig('gb_reactor_jac_isobaric_batch', I, [P, I, V, RP, I, I, V, V, V])
    sig('gb_reactor_jac_isobaric_host', I, [P, I, V, RP, I, I, V, V])
    sig('gb_reactor_rhs_isochoric_batch', I, [P, I, V, RP, D, V, V])
    sig('gb_reactor_rhs_isochoric_host', I, [P, I, V, RP, D, V])
    sig('gb_reactor_jac_isochoric_batch', I, [P, I, V, RP, D, I, V, V, V])
    sig('gb_reactor_jac_isochoric_host', I, [P, I, V, RP, D, I, V, V])
    sig('gb_flamelet_stencils', I, [P, dp, I, dp, dp, dp, dp, dp, dp, dp])
    sig('gb_flamelet_jac_indices', I, [P, I, ip, ip])
    FP = C.POINTER(FlameletParams)
    sig('gb_flamelet_rhs_batch', I, [P, I, V, FP, V, V])
    sig('gb_flamelet_rhs_host', I, [P, I, V, FP, V])
    sig('gb_flamelet_jacobian_batch', I, [P, I, V, FP, I, D, I, D, I, I, V, V, V])
    sig('gb_flamelet_jacobian_host', I, [P, I, V, FP, I, D, I, D, I, I, V, V])
    sig('gb_btddod_full_factorize_batch', I, [I, V, I, I, V, V, V])
    sig('gb_btddod_full_solve_batch', I, [I, V, V, V, V, I, I, V, V])
    sig('gb_btddod_full_factorize_inv_batch', I, [I, V, I, I, V, V, V, V])
    sig('gb_btddod_full_solve_inv_batch', I, [I, V, V, V, V, I, I, V, V, V])
    sig('gb_btddod_full_invert_batch', I, [I, V, I, I, V, V, V])
    sig('gb_btddod_full_invert_twisted_batch', I, [I, V, I, I, V, V, V])
    sig('gb_max_real_eigenvalue_batch', I, [I, I, V, V, V])
    sig('gb_btddod_full_matvec_batch', I, [I, V, V, I, I, V, V])
    sig('gb_btddod_scale_and_add_diagonal_batch', I, [I, V, D, V, D, I, I, V])
    PP = C.POINTER(C.c_void_p)
    sig('gb_esdirk_stage_begin_batch', I, [I, I, I, PP, dp, D, V, V, V, V, V, V, V, V])
    sig('gb_newton_update_batch', I, [I, I, V, V, V, V, V, V])
    sig('gb_newton_tail_batch', I, [I, I, V, V, V, V, V, D, V, D, V, V, V, V, V, ip, V])
    sig('gb_esdirk_finish_batch', I, [I, I, I, PP, dp, dp, V, V, V, V, V])
    sig('gb_accept_step_batch', I, [I, I, V, V, I, V, V])
    sig('gb_count_nonfinite_members_batch', I, [I, L, V, L, V, V, V])
    sig('gb_newton_tail_staged_batch', I, [I, I, I, dp, I, V, V, V, V, D, V, D, V, V, V, V, V, V, V, V, V, V, ip, V])
    sig('gb_flamelet_esdirk_stages_batch', I, [P, I, FP, V, V, V, V, I, dp, V, V, D, V, D, I, V, V, V, V, V, V, V, V, V, V, V, ip, V])
    sig('gb_flamelet_async_tick_batch', I, [P, I, FP, V, V, V, I, dp, dp, dp, V, V, D, V, D, I, I] + [V] * 15 +
        [I, V, V, V, V, V, V, V, V, I, V])
    sig('gb_flamelet_newton_stage_batch', I, [P, I, FP, V, V, V, V, V, V, V, D, V, D, I, V, V, V, V, V, V, ip, V])
    sig('gb_btddod_full_factorize_host', I, [I, V, I, I, V, V])
    sig('gb_btddod_full_solve_host', I, [I, V, V, V, V, I, I, V])
    sig('gb_btddod_full_matvec_host', I, [I, V, V, I, I, V])
    sig('gb_btddod_scale_and_add_diagonal_host', I, [I, V, D, V, D, I, I])
    _lib = lib
    return lib


def check(rc, what):
    """raises on a negative status; a positive one is the number of members whose output holds an Inf or NaN
    (griffon_b200.h: the synchronous *_host entry points report it) and is returned"""
    if rc < 0:
        msg = load_library().gb_last_error()
        raise GriffonB200Error(f'{what} failed (code {rc}): {msg.decode() if msg else ""}')
    return rc


def kernel_launch_count():
    return load_library().gb_kernel_launch_count()


def measure_fp64_peak(kind=0):
    """(Tflop/s, FP64 thread instructions per clock and SM) of the device's FP64 pipe: kind 0 = DFMA, 1 = DMUL+DADD"""
    tf, ipc = C.c_double(0.), C.c_double(0.)
    check(load_library().gb_measure_fp64_peak(int(kind), C.byref(tf), C.byref(ipc)), 'measure_fp64_peak')
    return tf.value, ipc.value


def measure_fp64_latency(kind=2):
    """cycles per dependent FP64 instruction of a single warp: kind 2 = DFMA, 3 = DADD, 4 = DMUL"""
    c = C.c_double(0.)
    check(load_library().gb_measure_fp64_latency(int(kind), C.byref(c)), 'measure_fp64_latency')
    return c.value


def _is_torch(x):
    return type(x).__module__.startswith('torch')


def _addr(x, dtype=np.float64):
    """raw address of a numpy array (host) or torch tensor (host or device), with layout checks"""
    if x is None:
        return None
    if _is_torch(x):
        import torch
        want = {np.float64: torch.float64, np.int32: torch.int32}[dtype]
        if x.dtype != want or not x.is_contiguous():
            raise TypeError(f'expected a contiguous {want} tensor')
        return C.c_void_p(x.data_ptr())
    if not isinstance(x, np.ndarray) or x.dtype != dtype or not x.flags['C_CONTIGUOUS']:
        raise TypeError(f'expected a C-contiguous {np.dtype(dtype).name} ndarray')
    return C.c_void_p(x.ctypes.data)


def _on_device(*xs):
    flags = [(_is_torch(x) and x.is_cuda) for x in xs if x is not None]
    if any(flags) and not all(flags):
        raise TypeError('mixing device tensors and host arrays in one call')
    return bool(flags) and all(flags)


def _stream():
    """raw handle of torch's current CUDA stream (the direct binding: building a torch.cuda.Stream object per call costs
    ~20 us, more than the launch it parameterises)"""
    import torch
    raw = getattr(torch._C, '_cuda_getCurrentRawStream', None)
    if raw is not None:
        return C.c_void_p(raw(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class PyCombustionKernels(MechanismSetters):
    _prefix = 'gb_'

    def __init__(self):
        self._lib = load_library()
        self._h = C.c_void_p(self._lib.gb_mech_create())
        self._keep = []

    def __del__(self):
        try:
            if self._h:
                self._lib.gb_mech_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _check(self, rc, what):
        check(rc, what)

    def commit(self):
        """pack and upload the mechanism tables to the current CUDA device (also done lazily by every call)"""
        check(self._lib.gb_mech_commit(self._h), 'mech_commit')

    # ---- thermodynamics (griffon.pyx:684-758) -------------------------------------------------------------------
    def _thermo1(self, what, T=None, y=None, aux=None, per_species=False):
        ns = self.n_species
        Ta = None if T is None else np.array([T], dtype=np.float64)
        aa = None if aux is None else np.array([aux], dtype=np.float64)
        out = np.zeros(ns if per_species else 1)
        check(self._lib.gb_thermo_host(self._h, what, 1, _addr(aa), _addr(Ta), _addr(y), _addr(out)), 'thermo')
        return out

    def mixture_molecular_weight(self, y):
        return float(self._thermo1(0, None, y)[0])

    def ideal_gas_density(self, p, T, y):
        return float(self._thermo1(1, T, y, aux=p)[0])

    def ideal_gas_pressure(self, rho, T, y):
        return float(self._thermo1(2, T, y, aux=rho)[0])

    def cp_mix(self, T, y):
        return float(self._thermo1(3, T, y)[0])

    def cv_mix(self, T, y):
        return float(self._thermo1(4, T, y)[0])

    def enthalpy_mix(self, T, y):
        return float(self._thermo1(5, T, y)[0])

    def energy_mix(self, T, y):
        return float(self._thermo1(6, T, y)[0])

    def species_cp(self, T, out):
        out[:] = self._thermo1(7, T, None, per_species=True)

    def species_cv(self, T, out):
        out[:] = self._thermo1(8, T, None, per_species=True)

    def species_enthalpies(self, T, out):
        out[:] = self._thermo1(9, T, None, per_species=True)

    def species_energies(self, T, out):
        out[:] = self._thermo1(10, T, None, per_species=True)

    def dcpdT_species(self, T, y, out):
        out[:] = self._thermo1(11, T, y, per_species=True)

    def mole_fractions(self, y, x):
        x[:] = self._thermo1(12, None, y, per_species=True)

    def thermo_batch(self, what, T, y, out, aux=None):
        """batched thermodynamic helper; `what` is one of the GB_THERMO_* codes of include/griffon_b200.h"""
        n = out.shape[0]
        if _on_device(T, y, out, aux):
            check(self._lib.gb_thermo_batch(self._h, what, n, _addr(aux), _addr(T), _addr(y), _addr(out), _stream()),
                  'thermo_batch')
        else:
            check(self._lib.gb_thermo_host(self._h, what, n, _addr(aux), _addr(T), _addr(y), _addr(out)), 'thermo')

    # ---- kinetics (griffon.pyx:763-783) ---------------------------------------------------------------------------
    def production_rates(self, T, rho, y, out_w):
        """positional (T, rho, y, out): what every reference caller passes (griffon.pyx:763-768)"""
        Ta, ra = np.array([T], dtype=np.float64), np.array([rho], dtype=np.float64)
        check(self._lib.gb_production_rates_host(self._h, 1, _addr(Ta), _addr(ra), _addr(y), _addr(out_w)),
              'production_rates')

    def prod_rates_primitive_sensitivities(self, rho, T, y, option, out):
        Ta, ra = np.array([T], dtype=np.float64), np.array([rho], dtype=np.float64)
        check(self._lib.gb_prod_rates_sens_host(self._h, 1, _addr(ra), _addr(Ta), _addr(y), int(option), _addr(out)),
              'prod_rates_primitive_sensitivities')

    def production_rates_batch(self, T, rho, y, out_w):
        n = T.shape[0]
        if _on_device(T, rho, y, out_w):
            check(self._lib.gb_production_rates_batch(self._h, n, _addr(T), _addr(rho), _addr(y), _addr(out_w),
                                                      _stream()), 'production_rates_batch')
        else:
            check(self._lib.gb_production_rates_host(self._h, n, _addr(T), _addr(rho), _addr(y), _addr(out_w)),
                  'production_rates')

    def prod_rates_sens_batch(self, rho, T, y, option, out):
        n = T.shape[0]
        if _on_device(T, rho, y, out):
            check(self._lib.gb_prod_rates_sens_batch(self._h, n, _addr(rho), _addr(T), _addr(y), int(option),
                                                     _addr(out), _stream()), 'prod_rates_sens_batch')
        else:
            check(self._lib.gb_prod_rates_sens_host(self._h, n, _addr(rho), _addr(T), _addr(y), int(option),
                                                    _addr(out)), 'prod_rates_sens')

    # ---- isobaric reactor (griffon.pyx:788-824) -------------------------------------------------------------------
    @staticmethod
    def _reactor_params(p, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open_):
        prm = ReactorParams()
        prm.pressure, prm.inflow_temperature, prm.tau = float(p), float(T_in), float(tau)
        prm.fluid_temperature, prm.surf_temperature = float(T_inf), float(T_surf)
        prm.h_conv, prm.eps_rad, prm.surface_area_over_volume = float(h_conv), float(eps_rad), float(SoV)
        prm.heat_transfer_option, prm.open = int(heat_option), int(bool(open_))
        # closed reactors pass a length-1 dummy y_in (reactors.py:226); it is never dereferenced
        a = _addr(y_in) if open_ else None
        prm.inflow_y = C.cast(a, c_double_p) if a is not None else None
        return prm

    def reactor_rhs_isobaric(self, state, p, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open_,
                             out_rhs):
        prm = self._reactor_params(p, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open_)
        return check(self._lib.gb_reactor_rhs_isobaric_host(self._h, 1, _addr(state), C.byref(prm), _addr(out_rhs)),
              'reactor_rhs_isobaric')

    def reactor_jac_isobaric(self, state, p, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open_,
                             rates_sens_option, sens_transform_option, out_rhs, out_jac):
        prm = self._reactor_params(p, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open_)
        return check(self._lib.gb_reactor_jac_isobaric_host(self._h, 1, _addr(state), C.byref(prm), int(rates_sens_option),
                                                     int(sens_transform_option), _addr(out_rhs), _addr(out_jac)),
              'reactor_jac_isobaric')

    def reactor_rhs_isobaric_batch(self, state, p, out_rhs, T_in=0., y_in=None, tau=0., T_inf=0., T_surf=0.,
                                   h_conv=0., eps_rad=0., SoV=0., heat_option=0, open_=False):
        """state [n, ns] -> out_rhs [n, ns]; torch CUDA tensors (async) or numpy arrays (sync, host<->device inside)"""
        n = state.shape[0]
        prm = self._reactor_params(p, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open_)
        if _on_device(state, out_rhs):
            check(self._lib.gb_reactor_rhs_isobaric_batch(self._h, n, _addr(state), C.byref(prm), _addr(out_rhs),
                                                          _stream()), 'reactor_rhs_isobaric_batch')
        else:
            return check(self._lib.gb_reactor_rhs_isobaric_host(self._h, n, _addr(state), C.byref(prm), _addr(out_rhs)),
                  'reactor_rhs_isobaric')

    def reactor_jac_isobaric_batch(self, state, p, out_rhs, out_jac, T_in=0., y_in=None, tau=0., T_inf=0., T_surf=0.,
                                   h_conv=0., eps_rad=0., SoV=0., heat_option=0, open_=False, rates_sens_option=0,
                                   sens_transform_option=0):
        """state [n, ns] -> out_rhs [n, ns], out_jac [n, ns*ns] (column-major ns x ns per state)"""
        n = state.shape[0]
        prm = self._reactor_params(p, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open_)
        if _on_device(state, out_rhs, out_jac):
            check(self._lib.gb_reactor_jac_isobaric_batch(self._h, n, _addr(state), C.byref(prm),
                                                          int(rates_sens_option), int(sens_transform_option),
                                                          _addr(out_rhs), _addr(out_jac), _stream()),
                  'reactor_jac_isobaric_batch')
        else:
            return check(self._lib.gb_reactor_jac_isobaric_host(self._h, n, _addr(state), C.byref(prm),
                                                         int(rates_sens_option), int(sens_transform_option),
                                                         _addr(out_rhs), _addr(out_jac)), 'reactor_jac_isobaric')

    # ---- isochoric reactor (griffon.pyx:831-866): state [rho, T, Y_0..Y_{ns-2}], Jacobian (ns+1) x (ns+1) -------------
    def reactor_rhs_isochoric(self, state, rho_in, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option,
                              open_, out_rhs):
        prm = self._reactor_params(0., T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open_)
        return check(self._lib.gb_reactor_rhs_isochoric_host(self._h, 1, _addr(state), C.byref(prm), float(rho_in),
                                                      _addr(out_rhs)), 'reactor_rhs_isochoric')

    def reactor_jac_isochoric(self, state, rho_in, T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option,
                              open_, rates_sens_option, out_rhs, out_jac):
        prm = self._reactor_params(0., T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open_)
        return check(self._lib.gb_reactor_jac_isochoric_host(self._h, 1, _addr(state), C.byref(prm), float(rho_in),
                                                      int(rates_sens_option), _addr(out_rhs), _addr(out_jac)),
              'reactor_jac_isochoric')

    def reactor_rhs_isochoric_batch(self, state, out_rhs, rho_in=0., T_in=0., y_in=None, tau=0., T_inf=0., T_surf=0.,
                                    h_conv=0., eps_rad=0., SoV=0., heat_option=0, open_=False):
        """state [n, ns+1] -> out_rhs [n, ns+1]; torch CUDA tensors (async) or numpy arrays (sync)"""
        n = state.shape[0]
        prm = self._reactor_params(0., T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open_)
        if _on_device(state, out_rhs):
            check(self._lib.gb_reactor_rhs_isochoric_batch(self._h, n, _addr(state), C.byref(prm), float(rho_in),
                                                           _addr(out_rhs), _stream()), 'reactor_rhs_isochoric_batch')
        else:
            return check(self._lib.gb_reactor_rhs_isochoric_host(self._h, n, _addr(state), C.byref(prm), float(rho_in),
                                                          _addr(out_rhs)), 'reactor_rhs_isochoric')

    def reactor_jac_isochoric_batch(self, state, out_rhs, out_jac, rho_in=0., T_in=0., y_in=None, tau=0., T_inf=0.,
                                    T_surf=0., h_conv=0., eps_rad=0., SoV=0., heat_option=0, open_=False,
                                    rates_sens_option=0):
        """state [n, ns+1] -> out_rhs [n, ns+1], out_jac [n, (ns+1)^2] (column-major per state)"""
        n = state.shape[0]
        prm = self._reactor_params(0., T_in, y_in, tau, T_inf, T_surf, h_conv, eps_rad, SoV, heat_option, open_)
        if _on_device(state, out_rhs, out_jac):
            check(self._lib.gb_reactor_jac_isochoric_batch(self._h, n, _addr(state), C.byref(prm), float(rho_in),
                                                           int(rates_sens_option), _addr(out_rhs), _addr(out_jac),
                                                           _stream()), 'reactor_jac_isochoric_batch')
        else:
            return check(self._lib.gb_reactor_jac_isochoric_host(self._h, n, _addr(state), C.byref(prm), float(rho_in),
                                                          int(rates_sens_option), _addr(out_rhs), _addr(out_jac)),
                  'reactor_jac_isochoric')

    # ---- methods of the reference class that are not on the B200 path (griffon.pyx:870-987): callable, raise clearly ----
    def _not_on_path(self, name):
        raise GriffonB200Error(f'{name} is not part of the B200 Griffon hot path (SURVEY.md section 8: two-dimensional '
                               f'flamelets are out of scope); there is no CPU fallback')

    def flamelet2d_rhs(self, *args, **kwargs):
        self._not_on_path('flamelet2d_rhs')

    def flamelet2d_factored_block_diag_jacobian(self, *args, **kwargs):
        self._not_on_path('flamelet2d_factored_block_diag_jacobian')

    def flamelet2d_offdiag_matvec(self, *args, **kwargs):
        self._not_on_path('flamelet2d_offdiag_matvec')

    def flamelet2d_matvec(self, *args, **kwargs):
        self._not_on_path('flamelet2d_matvec')

    def flamelet2d_block_diag_solve(self, *args, **kwargs):
        self._not_on_path('flamelet2d_block_diag_solve')

    def flamelet_newton_stage_batch(self, n_flamelets, prm, d_factors, l_values, dinv, system_rows, explicit, q, dt, gamma,
                                    weights, tolerance, max_iterations, x, f, res, conv, work, n_unconverged):
        """the Newton loop of one implicit stage on the device (griffon_b200.h: gb_flamelet_newton_stage_batch); returns
        (members left unconverged, iterations taken)"""
        its = C.c_int(0)
        left = check(self._lib.gb_flamelet_newton_stage_batch(
            self._h, int(n_flamelets), C.byref(prm), _addr(d_factors), _addr(l_values), _addr(dinv),
            None if system_rows is None else _addr(system_rows, np.int32), _addr(explicit), _addr(q), _addr(dt),
            float(gamma), _addr(weights), float(tolerance), int(max_iterations), _addr(x), _addr(f), _addr(res),
            _addr(conv, np.int32), _addr(work), _addr(n_unconverged, np.int32), C.byref(its), _stream()),
            'flamelet_newton_stage_batch')
        return left, its.value

    def flamelet_esdirk_stages_batch(self, n_flamelets, prm, d_factors, l_values, dinv, system_rows, tableau, q, dt, gamma,
                                     weights, tolerance, max_iterations, x, f, res, explicit, K, stage, iters, nlfail, done,
                                     work, n_left):
        """all implicit stages of one ESDIRK step, members independent of each other (griffon_b200.h:
        gb_flamelet_esdirk_stages_batch); returns (members not done, rounds of kernels taken)"""
        ns_ = len(tableau)
        tab = (C.c_double * (ns_ * ns_))(*[float(tableau[a][b]) if b < len(tableau[a]) else 0. for a in range(ns_)
                                           for b in range(ns_)])
        rounds = C.c_int(0)
        i32 = lambda a: _addr(a, np.int32)
        left = check(self._lib.gb_flamelet_esdirk_stages_batch(
            self._h, int(n_flamelets), C.byref(prm), _addr(d_factors), _addr(l_values), _addr(dinv),
            None if system_rows is None else i32(system_rows), ns_, tab, _addr(q), _addr(dt), float(gamma),
            _addr(weights), float(tolerance), int(max_iterations), _addr(x), _addr(f), _addr(res), _addr(explicit),
            _addr(K), i32(stage), i32(iters), i32(nlfail), i32(done), _addr(work), i32(n_left), C.byref(rounds),
            _stream()), 'flamelet_esdirk_stages_batch')
        return left, rounds.value

    def flamelet_async_tick_batch(self, n_flamelets, prm, dev, host, tableau_c, b_c, bh_c, nstages, gamma, tolerance,
                                  max_iterations, clip_negative, max_rounds, with_start, members=None):
        """one tick of the asynchronous batch integrator (griffon_b200.h: gb_flamelet_async_tick_batch). dev: dict of
        raw device addresses (ints) keyed like the C arguments, host: dict of raw host addresses; returns the rounds."""
        d, h = dev, host
        return check(self._lib.gb_flamelet_async_tick_batch(
            self._h, int(n_flamelets), C.byref(prm), d['J'], d['L'], d['Dinv'], int(nstages), tableau_c, b_c, bh_c,
            d['q'], d['dt'], float(gamma), d['w'], float(tolerance), int(max_iterations), int(bool(clip_negative)),
            d['x'], d['f'], d['res'], d['expl'], d['K'], d['state'], d['stage'], d['iters'], d['nlfail'], d['nits'],
            d['work'], d['dq'], d['stats'], d['start'], d['dtin'], int(max_rounds),
            h['start'] if with_start else None, h['dt'], h['state'], h['stage'], h['stats'], h['nlfail'], h['q'],
            None if members is None else members.ctypes.data, 0 if members is None else int(members.size),
            _stream()), 'flamelet_async_tick_batch')

    # ---- flamelet (griffon.pyx:556-679) ---------------------------------------------------------------------------
    def flamelet_stencils(self, dz, nzi, chi, inv_lewis, out_cmajor, out_csub, out_csup, out_mcoeff, out_ncoeff):
        check(self._lib.gb_flamelet_stencils(self._h, dptr(dz), int(nzi), dptr(chi), dptr(inv_lewis),
                                             dptr(out_cmajor), dptr(out_csub), dptr(out_csup), dptr(out_mcoeff),
                                             dptr(out_ncoeff)), 'flamelet_stencils')

    def flamelet_jac_indices(self, nzi, out_rows, out_cols):
        check(self._lib.gb_flamelet_jac_indices(self._h, int(nzi), iptr(out_rows), iptr(out_cols)),
              'flamelet_jac_indices')

    @staticmethod
    def _flamelet_params(p, oxy, fuel, adiabatic, T_conv, T_rad, h_conv, h_rad, nzi, cmajor, csub, csup, mcoeff,
                         ncoeff, chi, include_enthalpy_flux, include_variable_cp, use_scaled_heat_loss,
                         strides=(0, 0, 0, 0)):
        prm = FlameletParams()
        prm.nzi, prm.pressure = int(nzi), float(p)

        def P(x):
            a = _addr(x)
            return C.cast(a, c_double_p) if a is not None else None

        prm.oxy_state, prm.fuel_state = P(oxy), P(fuel)
        prm.adiabatic = int(bool(adiabatic))
        if not adiabatic:
            prm.T_convection, prm.h_convection = P(T_conv), P(h_conv)
            prm.T_radiation, prm.h_radiation = P(T_rad), P(h_rad)
        prm.cmajor, prm.csub, prm.csup = P(cmajor), P(csub), P(csup)
        prm.mcoeff, prm.ncoeff, prm.chi = P(mcoeff), P(ncoeff), P(chi)
        prm.include_enthalpy_flux = int(bool(include_enthalpy_flux))
        prm.include_variable_cp = int(bool(include_variable_cp))
        prm.use_scaled_heat_loss = int(bool(use_scaled_heat_loss))
        prm.stride_heat, prm.stride_coeff, prm.stride_mn, prm.stride_chi = [int(s) for s in strides]
        return prm

    def flamelet_rhs(self, state, p, oxy, fuel, adiabatic, T_conv, T_rad, h_conv, h_rad, nzi, cmajor, csub, csup,
                     mcoeff, ncoeff, chi, include_enthalpy_flux, include_variable_cp, use_scaled_heat_loss, out_rhs):
        """Python argument order (T_conv, T_rad, h_conv, h_rad) as in griffon.pyx:580-620"""
        prm = self._flamelet_params(p, oxy, fuel, adiabatic, T_conv, T_rad, h_conv, h_rad, nzi, cmajor, csub, csup,
                                    mcoeff, ncoeff, chi, include_enthalpy_flux, include_variable_cp,
                                    use_scaled_heat_loss)
        return check(self._lib.gb_flamelet_rhs_host(self._h, 1, _addr(state), C.byref(prm), _addr(out_rhs)), 'flamelet_rhs')

    def flamelet_jacobian(self, state, p, oxy, fuel, adiabatic, T_conv, T_rad, h_conv, h_rad, nzi, cmajor, csub, csup,
                          mcoeff, ncoeff, chi, compute_eigenvalues, diffterm, scale_and_offset, prefactor,
                          rates_sens_option, sens_transform_option, include_enthalpy_flux, include_variable_cp,
                          use_scaled_heat_loss, out_expeig, out_jac):
        prm = self._flamelet_params(p, oxy, fuel, adiabatic, T_conv, T_rad, h_conv, h_rad, nzi, cmajor, csub, csup,
                                    mcoeff, ncoeff, chi, include_enthalpy_flux, include_variable_cp,
                                    use_scaled_heat_loss)
        return check(self._lib.gb_flamelet_jacobian_host(self._h, 1, _addr(state), C.byref(prm),
                                                  int(bool(compute_eigenvalues)), float(diffterm),
                                                  int(bool(scale_and_offset)), float(prefactor),
                                                  int(rates_sens_option), int(sens_transform_option),
                                                  _addr(out_expeig), _addr(out_jac)), 'flamelet_jacobian')

    def flamelet_rhs_batch(self, n_flamelets, state, prm, out_rhs):
        """state/out [F, nzi*ns]; prm from `_flamelet_params` with arrays living where `state` lives"""
        if _on_device(state, out_rhs):
            check(self._lib.gb_flamelet_rhs_batch(self._h, int(n_flamelets), _addr(state), C.byref(prm),
                                                  _addr(out_rhs), _stream()), 'flamelet_rhs_batch')
        else:
            return check(self._lib.gb_flamelet_rhs_host(self._h, int(n_flamelets), _addr(state), C.byref(prm),
                                                 _addr(out_rhs)), 'flamelet_rhs')

    def flamelet_jacobian_batch(self, n_flamelets, state, prm, out_jac, compute_eigenvalues=False, diffterm=0.,
                                scale_and_offset=False, prefactor=1., rates_sens_option=0, sens_transform_option=0,
                                out_expeig=None):
        if _on_device(state, out_jac):
            check(self._lib.gb_flamelet_jacobian_batch(self._h, int(n_flamelets), _addr(state), C.byref(prm),
                                                       int(bool(compute_eigenvalues)), float(diffterm),
                                                       int(bool(scale_and_offset)), float(prefactor),
                                                       int(rates_sens_option), int(sens_transform_option),
                                                       _addr(out_expeig), _addr(out_jac), _stream()),
                  'flamelet_jacobian_batch')
        else:
            return check(self._lib.gb_flamelet_jacobian_host(self._h, int(n_flamelets), _addr(state), C.byref(prm),
                                                      int(bool(compute_eigenvalues)), float(diffterm),
                                                      int(bool(scale_and_offset)), float(prefactor),
                                                      int(rates_sens_option), int(sens_transform_option),
                                                      _addr(out_expeig), _addr(out_jac)), 'flamelet_jacobian')


# ---- BTDDOD module functions (griffon.pyx:1006-1113) --------------------------------------------------------------
def _bt(name_batch, name_host, on_dev, *args):
    lib = load_library()
    if on_dev:
        check(getattr(lib, name_batch)(*args, _stream()), name_batch)
    else:
        check(getattr(lib, name_host)(*args), name_host)


def py_btddod_full_factorize(out_d_factors, num_blocks, block_size, out_l_values, out_d_pivots, n_systems=1):
    _bt('gb_btddod_full_factorize_batch', 'gb_btddod_full_factorize_host',
        _on_device(out_d_factors, out_l_values, out_d_pivots),
        int(n_systems), _addr(out_d_factors), int(num_blocks), int(block_size), _addr(out_l_values),
        _addr(out_d_pivots, np.int32))


def py_btddod_full_solve(d_factors, l_values, d_pivots, rhs, num_blocks, block_size, out_solution, n_systems=1):
    _bt('gb_btddod_full_solve_batch', 'gb_btddod_full_solve_host',
        _on_device(d_factors, l_values, d_pivots, rhs, out_solution),
        int(n_systems), _addr(d_factors), _addr(l_values), _addr(d_pivots, np.int32), _addr(rhs), int(num_blocks),
        int(block_size), _addr(out_solution))


def btddod_full_factorize_inv(out_d_factors, num_blocks, block_size, out_l_values, out_d_pivots, out_dinv, n_systems=1):
    """extension (device arrays only): factorisation that also returns the explicit inverses of the diagonal blocks"""
    check(load_library().gb_btddod_full_factorize_inv_batch(int(n_systems), _addr(out_d_factors), int(num_blocks),
                                                            int(block_size), _addr(out_l_values),
                                                            _addr(out_d_pivots, np.int32), _addr(out_dinv), _stream()),
          'gb_btddod_full_factorize_inv_batch')


def max_real_eigenvalue(blocks, n, out, n_blocks):
    """out[b] = max Re(lambda) of the b-th n x n block (device tensors) -- gb_max_real_eigenvalue_batch"""
    check(load_library().gb_max_real_eigenvalue_batch(int(n_blocks), int(n), _addr(blocks), _addr(out), _stream()),
          'gb_max_real_eigenvalue_batch')


def btddod_full_invert(matrix, num_blocks, block_size, out_l_values, out_dinv, n_systems=1, twisted=False):
    """extension (device arrays only): block-Thomas elimination by Gauss-Jordan inverses; `matrix` is left intact and
    serves as the d_factors argument of btddod_full_solve_inv. twisted: the two-sided elimination (two CTAs of a
    cluster per system meeting in the middle block; btddod_full_solve_inv recognises the format by its tag)"""
    lib = load_library()
    fn = lib.gb_btddod_full_invert_twisted_batch if twisted else lib.gb_btddod_full_invert_batch
    check(fn(int(n_systems), _addr(matrix), int(num_blocks), int(block_size), _addr(out_l_values), _addr(out_dinv),
             _stream()),
          'gb_btddod_full_invert_twisted_batch' if twisted else 'gb_btddod_full_invert_batch')


def btddod_full_solve_inv(d_factors, l_values, dinv, rhs, num_blocks, block_size, out_solution, n_systems=1,
                          system_rows=None):
    """extension (device arrays only): block-Thomas solve whose back sweep multiplies by the stored inverses.
    system_rows (int32 device tensor, optional): position in the factor arrays of each right-hand side's system"""
    check(load_library().gb_btddod_full_solve_inv_batch(int(n_systems), _addr(d_factors), _addr(l_values), _addr(dinv),
                                                        _addr(rhs), int(num_blocks), int(block_size),
                                                        _addr(out_solution),
                                                        None if system_rows is None else _addr(system_rows, np.int32), _stream()),
          'gb_btddod_full_solve_inv_batch')


# ---- vector kernels of the batched implicit integrator (device tensors; griffon_b200.h) ----------------------------
def _kptrs(ks):
    arr = (C.c_void_p * len(ks))(*[k.data_ptr() for k in ks])
    return arr


def _hvec(v):
    return (C.c_double * len(v))(*[float(x) for x in v])


def esdirk_stage_begin(ks, coefs, gamma, dt, x, q, f, explicit_out, res_out, conv):
    """explicit = sum_j coefs[j]*ks[j] (from the last term down), res = dt*(gamma*f + explicit) - (x - q), conv = 0"""
    n, ndof = x.shape
    check(load_library().gb_esdirk_stage_begin_batch(int(n), int(ndof), len(ks), _kptrs(ks), _hvec(coefs), float(gamma),
                                                     _addr(dt), _addr(x), _addr(q), _addr(f), _addr(explicit_out),
                                                     _addr(res_out), _addr(conv, np.int32), _stream()),
          'gb_esdirk_stage_begin_batch')


def newton_update(x, dx, conv, xn, n_unconverged):
    n, ndof = x.shape
    check(load_library().gb_newton_update_batch(int(n), int(ndof), _addr(x), _addr(dx), _addr(conv, np.int32), _addr(xn),
                                                _addr(n_unconverged, np.int32), _stream()), 'gb_newton_update_batch')


import threading as _threading

_tls = _threading.local()  # (batches integrated concurrently from several host threads each read their own count)


def _count_slot():
    c = getattr(_tls, 'count', None)
    if c is None:
        c = _tls.count = C.c_int(0)
    return c


def newton_tail(fn, xn, explicit, q, dt, gamma, weights, tolerance, x, f, res, conv, n_unconverged, read_count=True):
    """the fused residual / select / norm / convergence kernel; returns the number of members still iterating (after
    a stream synchronisation) when read_count, else None"""
    n, ndof = x.shape
    check(load_library().gb_newton_tail_batch(int(n), int(ndof), _addr(fn), _addr(xn), _addr(explicit), _addr(q),
                                              _addr(dt), float(gamma), _addr(weights), float(tolerance), _addr(x),
                                              _addr(f), _addr(res), _addr(conv, np.int32),
                                              _addr(n_unconverged, np.int32),
                                              C.byref(_count_slot()) if read_count else None, _stream()),
          'gb_newton_tail_batch')
    return _count_slot().value if read_count else None


def esdirk_finish(ks, b, bh, dt, weights, dq, stats):
    n, ndof = dq.shape
    check(load_library().gb_esdirk_finish_batch(int(n), int(ndof), len(ks), _kptrs(ks), _hvec(b), _hvec(bh), _addr(dt),
                                                _addr(weights), _addr(dq), _addr(stats), _stream()),
          'gb_esdirk_finish_batch')


def count_nonfinite_members(a, b=None, flags_out=None):
    """number of rows of the device tensor a [n, la] (or of b [n, lb]) that hold an Inf or NaN; flags_out (int32 [n],
    optional) receives the per-row flags. One kernel and one 4-byte read-back instead of isfinite / all / any passes."""
    n = a.shape[0]
    return check(load_library().gb_count_nonfinite_members_batch(
        int(n), int(a.shape[1]), _addr(a), 0 if b is None else int(b.shape[1]), _addr(b),
        None if flags_out is None else _addr(flags_out, np.int32), _stream()), 'gb_count_nonfinite_members_batch')


def accept_step(dq, accept, clip_negative, q):
    n, ndof = dq.shape
    check(load_library().gb_accept_step_batch(int(n), int(ndof), _addr(dq), _addr(accept, np.int32),
                                              int(bool(clip_negative)), _addr(q), _stream()), 'gb_accept_step_batch')


def py_btddod_full_matvec(matrix_values, vec, num_blocks, block_size, out_matvec, n_systems=1):
    _bt('gb_btddod_full_matvec_batch', 'gb_btddod_full_matvec_host', _on_device(matrix_values, vec, out_matvec),
        int(n_systems), _addr(matrix_values), _addr(vec), int(num_blocks), int(block_size), _addr(out_matvec))


def py_btddod_scale_and_add_diagonal(in_out_matrix_values, matrix_scale, diagonal, diag_scale, num_blocks, block_size,
                                     n_systems=1):
    _bt('gb_btddod_scale_and_add_diagonal_batch', 'gb_btddod_scale_and_add_diagonal_host',
        _on_device(in_out_matrix_values, diagonal),
        int(n_systems), _addr(in_out_matrix_values), float(matrix_scale), _addr(diagonal), float(diag_scale),
        int(num_blocks), int(block_size))


# ---- module functions of the reference that are not on the B200 path (griffon.pyx:1019-1105): the block-Jacobi /
# Gauss-Seidel helpers of the 2-D flamelet solver. Callable, raise clearly. ---------------------------------------------
def _btddod_not_on_path(name):
    def f(*args, **kwargs):
        raise GriffonB200Error(f'{name} is not part of the B200 Griffon hot path (SURVEY.md section 8: the flamelet '
                               f'Newton / ESDIRK loops use py_btddod_full_factorize / py_btddod_full_solve); there is no '
                               f'CPU fallback')
    f.__name__ = name
    return f


py_btddod_blockdiag_matvec = _btddod_not_on_path('py_btddod_blockdiag_matvec')
py_btddod_blockdiag_factorize = _btddod_not_on_path('py_btddod_blockdiag_factorize')
py_btddod_blockdiag_solve = _btddod_not_on_path('py_btddod_blockdiag_solve')
py_btddod_lowerfulltriangle_solve = _btddod_not_on_path('py_btddod_lowerfulltriangle_solve')
py_btddod_upperfulltriangle_solve = _btddod_not_on_path('py_btddod_upperfulltriangle_solve')
py_btddod_scale_and_add_scaled_block_diagonal = _btddod_not_on_path('py_btddod_scale_and_add_scaled_block_diagonal')
