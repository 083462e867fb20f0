"""
Tabulated-chemistry library builders on the B200 Griffon path (mirror of the sweep drivers of tabulation.py:55-842;
the presumed-PDF part of that file is out of scope).

Same public functions, arguments, defaults, library layout (dimension and property names) and extinction /
sub-sampling rules as the reference. What is different is how the flamelets are advanced:
  * every flamelet solve runs through spitfire_b200.flamelet on the GPU (batched kernels behind the C-ABI);
  * `build_adiabatic_slfm_library(..., wave=W)` may solve W consecutive dissipation rates at once, all started from
    the last converged member (the reference's chain is W = 1, the default, and is followed step for step then);
  * the per-chi_st heat-loss expansions of the non-adiabatic builders -- the reference's `Pool.starmap` units -- are
    dealt to the ranks of torch.distributed (one process per GPU) and gathered once at the end
    (spitfire_b200.parallel); `num_procs` > 1 forks a process pool over the dissipation rates as the reference does, but
    only for a host-side Griffon object (the injected CPU checker) -- on the device it is ignored;
  * specific enthalpies come from Griffon's `enthalpy_mix` instead of a Cantera SolutionArray (analysis.py:77-98);
  * the structured-defect interpolation (a Python triple loop over chi, property and grid point in the reference,
    tabulation.py:629-654) is done with array operations.
"""
import copy
from time import perf_counter

import numpy as np

from spitfire_b200 import parallel
from spitfire_b200.flamelet import Flamelet, FlameletBatch, FlameletSpec
from spitfire_b200.library import Dimension, Library

_mixture_fraction_name = 'mixture_fraction'
_dissipation_rate_name = 'dissipation_rate'
_enthalpy_defect_name = 'enthalpy_defect'
_stoich_suffix = '_stoich'


def _write_library_header(lib_type, mech, fuel, oxy, verbose):
    if verbose and parallel.rank() == 0:
        print('-' * 82)
        print(f'building {lib_type} library')
        print('-' * 82)
        print(f'- mechanism: {mech.mech_file_path}')
        print(f'- {mech.n_species} species, {mech.n_reactions} reactions')
        print(f'- stoichiometric mixture fraction: {mech.stoich_mixture_fraction(fuel, oxy):.3f}')
        print('-' * 82)
    return perf_counter()


def _write_library_footer(cput0, verbose):
    if verbose and parallel.rank() == 0:
        print('-' * 82)
        print(f'library built in {perf_counter() - cput0:6.2f} s')
        print('-' * 82, flush=True)


def compute_specific_enthalpy(mechanism, output_library):
    """add the mixture's specific enthalpy, 'enthalpy', to a library holding temperature and mass fractions
    (the role of analysis.compute_specific_enthalpy, analysis.py:77-98, with Griffon's enthalpy_mix)"""
    names = mechanism.species_names
    T = np.ascontiguousarray(output_library['temperature'], dtype=np.float64)
    Y = np.stack([np.asarray(output_library['mass fraction ' + s], dtype=np.float64) for s in names], axis=-1)
    Tf, Yf = T.reshape(-1), np.ascontiguousarray(Y.reshape(-1, len(names)))
    g = mechanism.griffon
    h = np.zeros(Tf.size)
    if hasattr(g, 'thermo_batch'):
        from spitfire_b200.griffon import load_library  # noqa: F401  (product path: one batched launch)
        g.thermo_batch(5, Tf, Yf, h)  # GB_THERMO_H_MIX
    else:
        for i in range(Tf.size):
            h[i] = g.enthalpy_mix(Tf[i], Yf[i])
    output_library['enthalpy'] = h.reshape(T.shape)
    return output_library


def _enthalpy_of_states(mechanism, T, Y):
    """specific enthalpy of the states (T [n], Y [n, ns]) -- the arithmetic of compute_specific_enthalpy for a selection"""
    Tf = np.ascontiguousarray(T, dtype=np.float64).reshape(-1)
    Yf = np.ascontiguousarray(Y, dtype=np.float64).reshape(Tf.size, -1)
    g = mechanism.griffon
    h = np.zeros(Tf.size)
    if hasattr(g, 'thermo_batch'):
        g.thermo_batch(5, Tf, Yf, h)  # GB_THERMO_H_MIX
    else:
        for i in range(Tf.size):
            h[i] = g.enthalpy_mix(Tf[i], Yf[i])
    return h


def _copy_specs(flamelet_specs):
    return FlameletSpec(**flamelet_specs) if isinstance(flamelet_specs, dict) else copy.copy(flamelet_specs)


def _initial_state_library(flamelet_specs, name):
    fs = _copy_specs(flamelet_specs)
    fs.initial_condition = name
    flamelet = Flamelet(fs)
    return flamelet.make_library_from_interior_state(flamelet.initial_interior_state)


def build_unreacted_library(flamelet_specs, verbose=True):
    """pure mixing of the streams, no reaction (tabulation.py:55-73)"""
    return _initial_state_library(flamelet_specs, 'unreacted')


def build_adiabatic_eq_library(flamelet_specs, verbose=True):
    """chemical equilibrium (infinitely fast chemistry) at every mixture fraction (tabulation.py:76-95)"""
    return _initial_state_library(flamelet_specs, 'equilibrium')


def build_adiabatic_bs_library(flamelet_specs, verbose=True):
    """Burke-Schumann (idealised complete combustion) chemistry (tabulation.py:98-116)"""
    return _initial_state_library(flamelet_specs, 'Burke-Schumann')


# ----------------------------------------------------------------------------------------------------------------------
def build_adiabatic_slfm_library(flamelet_specs, diss_rate_values=np.logspace(-3, 2, 16),
                                 diss_rate_ref='stoichiometric', verbose=True, solver_verbose=False,
                                 _return_intermediates=False, include_extinguished=False, diss_rate_log_scaled=True,
                                 wave=1, tolerance=1.e-6):
    """Adiabatic strained-laminar-flamelet library over (mixture fraction, stoichiometric dissipation rate)
    (tabulation.py:229-336): equilibrium start, steady solve per dissipation rate in the given order, each started from
    the previous solution, stopping at the first extinguished member (max(T - T_linear) < 10 K) unless
    include_extinguished.

    wave : int
        (extension) how many consecutive dissipation rates are attempted together on the device with Newton's method,
        all started from the last converged solution; the converged prefix is kept and the chain continues from its
        last member (a member Newton cannot reach is solved on its own with the full solver chain, as in the
        reference). 1 reproduces the reference's chain exactly."""
    if isinstance(flamelet_specs, dict):
        flamelet_specs = FlameletSpec(**flamelet_specs)
    m, fuel, oxy = flamelet_specs.mech_spec, flamelet_specs.fuel_stream, flamelet_specs.oxy_stream
    flamelet_specs.initial_condition = 'equilibrium'
    use_max = diss_rate_ref == 'maximum'

    def set_chi(value):
        if use_max:
            flamelet_specs.max_dissipation_rate = value
        else:
            flamelet_specs.stoich_dissipation_rate = value

    set_chi(0.)
    cput00 = _write_library_header('adiabatic SLFM', m, fuel, oxy, verbose)
    f = Flamelet(flamelet_specs)  # evaluates the equilibrium initial condition once
    flamelet_specs.initial_condition = np.copy(f.initial_interior_state)
    say = verbose and parallel.rank() == 0
    table_dict, x_values = dict(), list()
    nchi = diss_rate_values.size
    suffix = _stoich_suffix if not use_max else '_max'
    z_st = m.stoich_mixture_fraction(fuel, oxy)
    idx, stop = 0, False
    wave = max(1, int(wave))
    while idx < nchi and not stop:
        members, libs = [], []
        for chival in diss_rate_values[idx:idx + wave]:
            set_chi(chival)
            members.append(Flamelet(flamelet_specs))
        cput0 = perf_counter()
        if len(members) > 1:
            # speculative wave: Newton only, all members from the last converged state; keep the converged prefix
            states, _, conv = FlameletBatch(members).steady_solve_newton(tolerance=tolerance, max_iterations=38,
                                                                         verbose=solver_verbose)
            n_ok = 0
            while n_ok < len(members) and conv[n_ok]:
                members[n_ok]._current_state = np.copy(states[n_ok])
                n_ok += 1
            members = members[:max(n_ok, 1)]
            libs = [fl.make_library_from_interior_state(fl._current_state) for fl in members[:n_ok]]
        if len(members) == 1 and (wave == 1 or not libs):
            libs = [members[0].compute_steady_state(tolerance=tolerance, verbose=solver_verbose, use_psitc=True)]
        dcput = perf_counter() - cput0
        for k, (flamelet, x_library) in enumerate(zip(members, libs)):
            if say:
                print(f'{idx + k + 1:4}/{nchi:4} (chi{suffix} = {diss_rate_values[idx + k]:8.1e} 1/s) ', end='')
            if np.max(flamelet.current_temperature - flamelet.linear_temperature) < 10. and not include_extinguished:
                if say:
                    print(' extinction detected, stopping. The extinguished state will not be included in the table.')
                stop = True
                break
            if say:
                print(f' converged in {dcput / len(members):6.2f} s, T_max = {np.max(flamelet.current_temperature):6.1f}',
                      flush=True)
            chi_st = flamelet._compute_dissipation_rate(np.array([z_st]), flamelet._max_dissipation_rate,
                                                        flamelet._dissipation_rate_form)[0]
            x_values.append(chi_st)
            table_dict[chi_st] = {k2: x_library[k2].ravel() for k2 in x_library.props}
            flamelet_specs.initial_condition = np.copy(flamelet.current_interior_state)
            if _return_intermediates:
                table_dict[chi_st]['adiabatic_state'] = np.copy(flamelet.current_interior_state)
        idx += len(members)
    if _return_intermediates:
        _write_library_footer(cput00, verbose)
        return table_dict, f.mixfrac_grid, np.array(x_values)
    z_dim = Dimension(_mixture_fraction_name, f.mixfrac_grid)
    x_dim = Dimension(_dissipation_rate_name + _stoich_suffix, np.array(x_values), diss_rate_log_scaled)
    output_library = Library(z_dim, x_dim)
    output_library.extra_attributes['mech_spec'] = m
    for quantity in table_dict[x_values[-1]]:
        values = output_library.get_empty_dataset()
        for ix, x in enumerate(x_values):
            values[:, ix] = table_dict[x][quantity]
        output_library[quantity] = values
    _write_library_footer(cput00, verbose)
    return output_library


# ----------------------------------------------------------------------------------------------------------------------
def _subsample_by_stoich_enthalpy(z, z_st, h_profiles, h_stoich_spacing, include_last):
    """indices of the profiles (rows of h_profiles) kept: the first one and every one whose stoichiometric enthalpy has
    dropped by more than h_stoich_spacing since the last kept one (tabulation.py:381-390, 498-505)"""
    indices = [0]
    n = h_profiles.shape[0]
    last_hst = np.interp(z_st, z, h_profiles[0])
    for i in range(n if include_last else n - 1):
        this_hst = np.interp(z_st, z, h_profiles[i])
        if last_hst - this_hst > h_stoich_spacing:
            indices.append(i)
            last_hst = this_hst
    if include_last and n - 1 not in indices:
        indices.append(n - 1)
    return indices


def _store_defect_profiles(managed_dict, chi_st, z, z_st, lib_get, props, h_profiles, indices):
    h_ad = h_profiles[0]
    for i in indices:
        defect = h_profiles[i] - h_ad
        gst = float(np.interp(z_st, z, defect))
        data = dict()
        data['enthalpy_defect'] = np.copy(defect)
        data['enthalpy_cons'] = np.copy(h_ad)
        data['enthalpy'] = np.copy(h_profiles[i])
        data[_mixture_fraction_name] = z
        for q in props:
            data[q] = lib_get(q, i)
        managed_dict[(chi_st, gst)] = data


def _transient_heat_loss_specs(flamelet_specs, table_dict, chi_st):
    fs = copy.copy(flamelet_specs)
    fs.initial_condition = table_dict[chi_st]['adiabatic_state']
    fs.stoich_dissipation_rate = chi_st
    fs.heat_transfer = 'nonadiabatic'
    fs.scale_heat_loss_by_temp_range = True
    fs.scale_convection_by_dissipation = True
    fs.use_linear_ref_temp_profile = True
    fs.convection_coefficient = fs.convection_coefficient if fs.convection_coefficient is not None else 1.e7
    fs.radiative_emissivity = 0.
    return fs


def _transient_integration_args(input_integration_args, solver_verbose):
    integration_args = {'first_time_step': 1.e-9, 'max_time_step': 1.e-1, 'write_log': solver_verbose, 'log_rate': 100,
                        'print_exception_on_failure': False}
    if input_integration_args is not None:
        integration_args.update(input_integration_args)
    integration_args.setdefault('transient_tolerance', 1.e-8)
    return integration_args


def _store_transient_library(managed_dict, chi_st, fnonad, transient_lib, h_stoich_spacing):
    """sub-sample the trajectory in stoichiometric enthalpy and store the kept profiles (tabulation.py:375-400)"""
    z = fnonad.mixfrac_grid
    mech = fnonad.mechanism
    z_st = mech.stoich_mixture_fraction(fnonad.fuel_stream, fnonad.oxy_stream)
    names = mech.species_names
    nt, nz = transient_lib['temperature'].shape
    if not FAST_TRANSIENT_STORE or nz < 8:
        h_tz = compute_specific_enthalpy(mech, transient_lib)['enthalpy']
        indices = _subsample_by_stoich_enthalpy(z, z_st, h_tz, h_stoich_spacing, include_last=True)
        props = [q for q in transient_lib.props]
        _store_defect_profiles(managed_dict, chi_st, z, z_st, lambda q, i: transient_lib[q][i, :], props, h_tz, indices)
        return
    # The sub-sampling looks at the enthalpy interpolated to z_st only -- a value np.interp forms from the two grid
    # points that bracket z_st -- and full profiles are stored for the kept time levels alone. So the enthalpy is
    # evaluated at a four-point window around z_st for every time level and on the whole grid for the kept ones:
    # ~5 x fewer states to assemble and evaluate than the whole (time, z) plane, same values (the enthalpy of a state
    # does not depend on which other states are evaluated with it).
    j = int(np.clip(np.searchsorted(z, z_st, side='right') - 1, 0, nz - 2))
    cols = np.arange(max(j - 1, 0), min(j + 3, nz))
    Tw = transient_lib['temperature'][:, cols]
    Yw = np.stack([transient_lib['mass fraction ' + sp][:, cols] for sp in names], axis=-1)
    h_win = _enthalpy_of_states(mech, Tw, Yw).reshape(nt, cols.size)
    indices = _subsample_by_stoich_enthalpy(z[cols], z_st, h_win, h_stoich_spacing, include_last=True)
    Tk = transient_lib['temperature'][indices, :]
    Yk = np.stack([transient_lib['mass fraction ' + sp][indices, :] for sp in names], axis=-1)
    h_keep = _enthalpy_of_states(mech, Tk, Yk).reshape(len(indices), nz)
    h_rows = {i: h_keep[k] for k, i in enumerate(indices)}
    props = [q for q in transient_lib.props]
    _store_defect_profiles(managed_dict, chi_st, z, z_st, lambda q, i: transient_lib[q][i, :], props, h_rows, indices)


def _expand_enthalpy_defect_dimension_transient(chi_st, managed_dict, flamelet_specs, table_dict, h_stoich_spacing,
                                                verbose, input_integration_args, solver_verbose):
    """one rapid-extinction trajectory at chi_st: ESDIRK64 with strong scaled convective heat loss until the
    temperature profile is nearly linear, sub-sampled in stoichiometric enthalpy (tabulation.py:339-405)"""
    fs = _transient_heat_loss_specs(flamelet_specs, table_dict, chi_st)
    integration_args = _transient_integration_args(input_integration_args, solver_verbose)
    cput0 = perf_counter()
    transient_lib, fnonad = None, None
    while transient_lib is None and integration_args['transient_tolerance'] > 1.e-15:
        try:
            fnonad = Flamelet(fs)
            transient_lib = fnonad.integrate_for_heat_loss(**integration_args)
        except Exception:
            if solver_verbose:
                print(f'Transient heat loss calculation failed with tolerance of '
                      f'{integration_args["transient_tolerance"]:.1e}, retrying with 100x lower...')
            integration_args['transient_tolerance'] *= 1.e-2
    if transient_lib is None:
        raise RuntimeError(f'heat-loss expansion at chi_st = {chi_st} failed at every tolerance')
    _store_transient_library(managed_dict, chi_st, fnonad, transient_lib, h_stoich_spacing)
    if verbose:
        print('chi_st = {:8.1e} 1/s converged in {:6.2f} s'.format(chi_st, perf_counter() - cput0), flush=True)


# Concurrent sub-batches of a rank's heat-loss trajectories (see _integrate_heat_loss_in_groups). Measured on config 5
# (56 trajectories, one B200): 1 group 15.3 s, 4 groups 19.2 s, 8 groups 22.6 s -- the rounds of the groups do not
# overlap on the device, because a small flamelet batch is spread over all SMs on purpose (k_rates: one tile of a few
# grid points per SM), so eight small batches cost eight times the latency of one. Off by default; the trajectories
# are bit-identical either way (the GPU library tests were run with 8 groups).
TRAJECTORY_GROUPS = 1  # None: chosen from the number of members (at most 8); 1: one lock-step batch


def _trajectory_groups(flamelet_specs, n_members):
    """how many independently advancing groups the heat-loss trajectories of a rank are split into (device path only)"""
    import os
    if not _griffon_on_device(flamelet_specs):
        return 1
    g = os.environ.get('GB_TRAJECTORY_GROUPS', TRAJECTORY_GROUPS)
    g = int(g) if g is not None else min(8, n_members // 2)
    return max(1, min(g, n_members))


def _integrate_heat_loss_in_groups(chi_list, flamelet_specs, table_dict, integration_args, groups):
    """The members of a lock-step batch wait for each other: every Newton iteration of a step is a round of
    latency-bound kernels that serves all members, and a step takes as many rounds as its SLOWEST member needs -- a
    different member at different times (measured on config 5: 15.8 k rounds for 56 members against 6.4 k Newton
    iterations of the longest trajectory on its own). Here the members are dealt to `groups` smaller batches that
    advance independently of each other: one host thread, one CUDA stream and one Griffon handle (its own work arrays)
    per group; the kernels of different groups overlap on the device (a round occupies one SM per member), the C-ABI
    calls release the interpreter lock. Every member's own arithmetic is unchanged, so the trajectories are the ones
    the single batch computes, bit for bit."""
    import threading

    import torch
    from spitfire_b200.mechanism import ChemicalMechanismSpec
    m = flamelet_specs.mech_spec
    parts = [list(range(g, len(chi_list), groups)) for g in range(groups)]
    batches = []
    for part in parts:  # (construction on the calling thread: it goes through the shared stream / mechanism objects)
        fs = copy.copy(flamelet_specs)
        fs.mech_spec = ChemicalMechanismSpec(mech_data=m.mech_data, griffon_factory=m._griffon_factory)
        fls = [Flamelet(_transient_heat_loss_specs(fs, table_dict, chi_list[k])) for k in part]
        batches.append((fls, FlameletBatch(fls)))
    torch.cuda.synchronize()
    device = torch.cuda.current_device()
    results, errors = [None] * groups, [None] * groups

    def work(g):
        try:
            torch.cuda.set_device(device)
            with torch.cuda.stream(torch.cuda.Stream()):
                results[g] = batches[g][1].integrate_for_heat_loss(**integration_args)
                torch.cuda.current_stream().synchronize()
        except BaseException as e:  # re-raised on the calling thread
            errors[g] = e

    threads = [threading.Thread(target=work, args=(g,)) for g in range(groups)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    flamelets, libs, failed = [None] * len(chi_list), [None] * len(chi_list), [None] * len(chi_list)
    for g, part in enumerate(parts):
        for j, k in enumerate(part):
            flamelets[k], libs[k], failed[k] = batches[g][0][j], results[g][0][j], results[g][1][j]
    return flamelets, libs, failed


def _expand_enthalpy_defect_dimension_transient_batch(chi_list, managed_dict, flamelet_specs, table_dict,
                                                      h_stoich_spacing, verbose, input_integration_args,
                                                      solver_verbose):
    """all of this rank's rapid-extinction trajectories advanced together on the device (FlameletBatch); every member
    keeps its own adaptive step sequence, so its trajectory is the one the one-at-a-time path computes. A member whose
    batched run fails is redone on its own with the reference's retry ladder."""
    if not chi_list:
        return
    cput0 = perf_counter()
    integration_args = _transient_integration_args(input_integration_args, solver_verbose)
    groups = _trajectory_groups(flamelet_specs, len(chi_list))
    if groups > 1:
        flamelets, libs, failed = _integrate_heat_loss_in_groups(chi_list, flamelet_specs, table_dict, integration_args,
                                                                 groups)
    else:
        flamelets = [Flamelet(_transient_heat_loss_specs(flamelet_specs, table_dict, chi_st)) for chi_st in chi_list]
        libs, failed = FlameletBatch(flamelets).integrate_for_heat_loss(**integration_args)
    for chi_st, fl, lib, bad in zip(chi_list, flamelets, libs, failed):
        if bad:
            _expand_enthalpy_defect_dimension_transient(chi_st, managed_dict, flamelet_specs, table_dict,
                                                        h_stoich_spacing, verbose, input_integration_args,
                                                        solver_verbose)
        else:
            _store_transient_library(managed_dict, chi_st, fl, lib, h_stoich_spacing)
    if verbose:
        print('{:} heat-loss trajectories (chi_st {:8.1e} .. {:8.1e} 1/s) advanced together in {:6.2f} s'.format(
            len(chi_list), min(chi_list), max(chi_list), perf_counter() - cput0), flush=True)


def _expand_enthalpy_defect_dimension_steady(chi_st, managed_dict, flamelet_specs, table_dict, h_stoich_spacing,
                                             verbose, input_integration_args, solver_verbose):
    """quasi-steady heat loss at chi_st: continuation in the convection coefficient with an adaptive increment until
    the flamelet extinguishes (tabulation.py:408-519)"""
    _expand_enthalpy_defect_dimension_steady_batch([chi_st], managed_dict, flamelet_specs, table_dict,
                                                   h_stoich_spacing, verbose, input_integration_args, solver_verbose)


def _expand_enthalpy_defect_dimension_steady_batch(chi_list, managed_dict, flamelet_specs, table_dict,
                                                   h_stoich_spacing, verbose, input_integration_args, solver_verbose):
    """the quasi-steady heat-loss continuations (tabulation.py:408-519) of several dissipation rates advanced together:
    every member keeps its own convection coefficient, increment and extinction test -- the sequence of steady problems
    it solves is the one the one-at-a-time path solves -- and each round's steady solves share their kernel launches
    (FlameletBatch.compute_steady_state: Newton, then pseudo-transient continuation, then ESDIRK, member by member)."""
    if not chi_list:
        return
    cput0 = perf_counter()
    diff_target, hval_max = 1.e-1, 1.e10

    class _Member(object):
        pass

    members = []
    for chi_st in chi_list:
        mb = _Member()
        fs = copy.copy(flamelet_specs)
        fs.initial_condition = table_dict[chi_st]['adiabatic_state']
        fs.stoich_dissipation_rate = chi_st
        fs.heat_transfer = 'nonadiabatic'
        fs.scale_heat_loss_by_temp_range = False
        fs.scale_convection_by_dissipation = False
        fs.use_linear_ref_temp_profile = True
        fs.radiative_emissivity = 0.
        fs.convection_coefficient = 0.
        mb.chi_st, mb.fs = chi_st, fs
        mb.flamelet = Flamelet(fs)
        mb.state_old = np.copy(mb.flamelet.current_interior_state)
        mb.hval, mb.dh = 0., 1.e-1
        mb.solutions = [{p: table_dict[chi_st][p] for p in table_dict[chi_st] if p != 'adiabatic_state'}]
        mb.hvalues = [0.]
        mb.current_state = table_dict[chi_st]['adiabatic_state']
        mb.first, mb.extinguished = True, False
        members.append(mb)
    while True:
        active = [mb for mb in members if mb.first or (not mb.extinguished and mb.hval < hval_max)]
        if not active:
            break
        for mb in active:
            mb.hval += mb.dh
            mb.first = False
            mb.fs.convection_coefficient = mb.hval
            mb.fs.initial_condition = mb.current_state
            mb.flamelet = Flamelet(mb.fs)
        if len(active) == 1:
            libs = [active[0].flamelet.compute_steady_state(verbose=solver_verbose)]
        else:
            FlameletBatch([mb.flamelet for mb in active]).compute_steady_state(verbose=solver_verbose)
            libs = [mb.flamelet.make_library_from_interior_state(mb.flamelet.current_interior_state) for mb in active]
        for mb, g_library in zip(active, libs):
            flamelet = mb.flamelet
            mb.current_state = flamelet.current_interior_state
            maxT = np.max(mb.current_state)
            diff_norm = np.max(np.abs(mb.current_state - mb.state_old) / (np.abs(mb.current_state) + 1.e-4))
            mb.extinguished = maxT < (np.max([flamelet.oxy_stream.T, flamelet.fuel_stream.T]) + 10.)
            mb.state_old = np.copy(mb.current_state)
            mb.dh *= np.min([np.max([np.sqrt(diff_target / diff_norm), 0.1]), 2.])
            mb.hvalues.append(mb.hval)
            mb.solutions.append({p: g_library[p].ravel() for p in g_library.props})
    for mb in members:
        flamelet, fs, chi_st = mb.flamelet, mb.fs, mb.chi_st
        z = flamelet.mixfrac_grid
        steady_lib = Library(Dimension(_mixture_fraction_name, z),
                             Dimension(_enthalpy_defect_name + _stoich_suffix, np.array(mb.hvalues)))
        steady_lib.extra_attributes['mech_spec'] = fs.mech_spec
        props = [p for p in table_dict[chi_st] if p != 'adiabatic_state']
        for p in props:
            values = steady_lib.get_empty_dataset()
            for ig, sol in enumerate(mb.solutions):
                values[:, ig] = sol[p].ravel()
            steady_lib[p] = values
        z_st = flamelet.mechanism.stoich_mixture_fraction(flamelet.fuel_stream, flamelet.oxy_stream)
        h_zt = compute_specific_enthalpy(fs.mech_spec, steady_lib)['enthalpy']
        indices = _subsample_by_stoich_enthalpy(z, z_st, h_zt.T, h_stoich_spacing, include_last=False)
        _store_defect_profiles(managed_dict, chi_st, z, z_st, lambda q, i: steady_lib[q][:, i], steady_lib.props, h_zt.T,
                               indices)
    if verbose:
        print('{:} quasi-steady heat-loss continuations (chi_st {:8.1e} .. {:8.1e} 1/s) converged in {:6.2f} s'.format(
            len(chi_list), min(chi_list), max(chi_list), perf_counter() - cput0), flush=True)


# heat-loss trajectories stored through the reduced enthalpy evaluation of _store_transient_library (same values)
FAST_TRANSIENT_STORE = True
_pool_job = None
# where the last non-adiabatic build of this process spent its time (seconds per stage; bench.py reports it)
LAST_BUILD_TIMES = dict()


def _pool_expand(chi_st):
    expand, flamelet_specs, table_dict, h_stoich_spacing, integration_args, solver_verbose = _pool_job
    try:  # one thread per worker: the pool already uses every core
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass
    out = dict()
    expand(chi_st, out, copy.copy(flamelet_specs), table_dict, h_stoich_spacing, False, integration_args, solver_verbose)
    return out


def _griffon_on_device(flamelet_specs):
    from spitfire_b200 import griffon as gmod
    specs = flamelet_specs if not isinstance(flamelet_specs, dict) else FlameletSpec(**flamelet_specs)
    return isinstance(specs.mech_spec.griffon, gmod.PyCombustionKernels)


def _build_unstructured_nonadiabatic_defect_slfm_library(flamelet_specs, heat_loss_expansion='transient',
                                                         diss_rate_values=np.logspace(-3, 2, 16),
                                                         diss_rate_ref='stoichiometric', verbose=True,
                                                         solver_verbose=False, h_stoich_spacing=10.e3, num_procs=1,
                                                         integration_args=None, wave=1, batch_expansions=True):
    """adiabatic chain on every rank (it is short and every rank needs all of it), then this rank's share of the
    independent heat-loss expansions, then one gather (tabulation.py:522-591)"""
    _t0 = perf_counter()
    table_dict, z_values, x_values = build_adiabatic_slfm_library(flamelet_specs, diss_rate_values, diss_rate_ref,
                                                                  verbose, solver_verbose, _return_intermediates=True,
                                                                  wave=wave)
    LAST_BUILD_TIMES.clear()
    LAST_BUILD_TIMES['adiabatic_chain_s'] = perf_counter() - _t0
    expand = _expand_enthalpy_defect_dimension_transient if heat_loss_expansion == 'transient' else \
        _expand_enthalpy_defect_dimension_steady
    if verbose and parallel.rank() == 0:
        print(f'expanding ({heat_loss_expansion}) enthalpy defect dimension on {parallel.world_size()} rank(s) ...',
              flush=True)
    cput0 = perf_counter()
    local = dict()
    mine = parallel.my_share(list(table_dict.keys()))
    on_device = _griffon_on_device(flamelet_specs)
    if num_procs > 1 and not on_device and len(mine) > 1:
        # the reference's own parallel mode (tabulation.py:542-568): a process pool over the dissipation rates, results
        # collected in a shared dictionary. Only meaningful for a host Griffon object (the CPU checker injected through
        # ChemicalMechanismSpec(griffon_factory=...)); on the device the rank's trajectories advance as one batch.
        # The workers are forked, so they inherit the specification (with its Griffon object) instead of unpickling it.
        import multiprocessing as mp
        global _pool_job
        _pool_job = (expand, flamelet_specs, table_dict, h_stoich_spacing, integration_args, solver_verbose)
        ctx = mp.get_context('fork')
        try:
            with ctx.Pool(processes=min(int(num_procs), len(mine))) as pool:
                for part in pool.map(_pool_expand, mine, chunksize=1):
                    local.update(part)
        finally:
            _pool_job = None
    elif heat_loss_expansion == 'transient' and batch_expansions and len(mine) > 1:
        _expand_enthalpy_defect_dimension_transient_batch(mine, local, flamelet_specs, table_dict, h_stoich_spacing,
                                                          verbose, integration_args, solver_verbose)
    elif heat_loss_expansion == 'steady' and batch_expansions and len(mine) > 1:
        _expand_enthalpy_defect_dimension_steady_batch(mine, local, flamelet_specs, table_dict, h_stoich_spacing,
                                                       verbose, integration_args, solver_verbose)
    else:
        for chi_st in mine:
            expand(chi_st, local, flamelet_specs, table_dict, h_stoich_spacing, verbose, integration_args,
                   solver_verbose)
    LAST_BUILD_TIMES['own_expansions_s'] = perf_counter() - cput0
    _t0 = perf_counter()
    parallel.barrier()
    LAST_BUILD_TIMES['wait_for_other_ranks_s'] = perf_counter() - _t0
    _t0 = perf_counter()
    merged = parallel.gather_profile_dicts(local)
    LAST_BUILD_TIMES['gather_s'] = perf_counter() - _t0
    if verbose and parallel.rank() == 0:
        print('-' * 82)
        print('enthalpy defect dimension expanded in {:6.2f} s'.format(perf_counter() - cput0))
        print('-' * 82, flush=True)
    return merged


def _interp_columns(x, xp, fp):
    """np.interp(x, xp, fp[:, j]) for every column j at once (xp increasing, values held constant outside its range):
    the bracketing interval and the abscissa differences are the same for all columns, so they are found once. Same
    arithmetic as numpy's kernel -- slope = (fp[k+1] - fp[k]) / (xp[k+1] - xp[k]); slope * (x - xp[k]) + fp[k] -- hence
    the same bits (tests/test_host_flamelet.py)."""
    x, xp, fp = np.asarray(x, dtype=np.float64), np.asarray(xp, dtype=np.float64), np.asarray(fp, dtype=np.float64)
    n = xp.size
    if n == 1:
        return np.repeat(fp[:1], x.size, axis=0)
    k = np.clip(np.searchsorted(xp, x, side='right') - 1, 0, n - 2)
    slope = (fp[k + 1] - fp[k]) / (xp[k + 1] - xp[k])[:, None]
    out = slope * (x - xp[k])[:, None] + fp[k]
    out[x <= xp[0]] = fp[0]
    out[x >= xp[-1]] = fp[-1]
    exact = np.nonzero(x == xp[np.clip(k + 1, 0, n - 1)])[0]  # (a knot hit from the left interval returns the knot value)
    if exact.size:
        out[exact] = fp[np.clip(k[exact] + 1, 0, n - 1)]
    return out


def _interpolate_to_structured_defect_dimension(unstructured_table, n_defect_stoich, verbose=False, extend=False):
    """piecewise-linear interpolation of every property at every (chi_st, z) onto n_defect_stoich equispaced
    stoichiometric enthalpy defects, held constant outside each chi_st's range (tabulation.py:594-665)"""
    cput0 = perf_counter()
    by_chi = dict()
    for (chi_st, g_st) in unstructured_table.keys():
        by_chi.setdefault(chi_st, []).append(g_st)
    all_g = [g for gs in by_chi.values() for g in gs]
    min_g, max_g = np.min(all_g), np.max(all_g)
    defect_space = np.linspace(min_g, max_g, n_defect_stoich)
    if extend:
        spacing = np.abs(defect_space[1] - defect_space[0])
        defect_space = np.linspace(min_g - 2 * spacing, max_g, n_defect_stoich + 2)
    structured = dict()
    for chi_st, gs in by_chi.items():
        g_sorted = np.sort(np.array(gs))
        entries = [unstructured_table[(chi_st, g)] for g in g_sorted]
        names = list(entries[0].keys())
        # all properties of this chi_st in one interpolation: [defect, property, z] -> columns (property, z); the
        # columns are independent, so every value is the one the property-by-property loop computes
        data_all = np.array([[np.asarray(e[q], dtype=np.float64) for q in names] for e in entries])
        ng, nq, nz = data_all.shape
        out_all = _interp_columns(defect_space, g_sorted, data_all.reshape(ng, nq * nz)).reshape(-1, nq, nz)
        for j, q in enumerate(names):
            data, out = data_all[:, j], out_all[:, j]
            if extend and q in ('enthalpy', 'enthalpy_defect') and g_sorted.size > 1:
                lo = defect_space < g_sorted[0]
                slope = (data[1] - data[0]) / (g_sorted[1] - g_sorted[0])
                out[lo] = data[0] + (defect_space[lo, None] - g_sorted[0]) * slope
                hi = defect_space > g_sorted[-1]
                slope = (data[-1] - data[-2]) / (g_sorted[-1] - g_sorted[-2])
                out[hi] = data[-1] + (defect_space[hi, None] - g_sorted[-1]) * slope
            if q in ('density', 'temperature') and np.any(out < 1.e-14):
                raise ValueError(f'{q} < 1.e-14 detected!')
        for ig, g in enumerate(defect_space):
            structured[(chi_st, g)] = {q: out_all[ig, j] for j, q in enumerate(names)}
    if verbose and parallel.rank() == 0:
        print('Structured enthalpy defect dimension built in {:6.2f} s'.format(perf_counter() - cput0), flush=True)
    return structured, np.array(sorted(by_chi.keys())), defect_space[::-1]


def _build_nonadiabatic_defect_slfm_library(flamelet_specs, heat_loss_expansion='transient',
                                            diss_rate_values=np.logspace(-3, 2, 16), diss_rate_ref='stoichiometric',
                                            verbose=True, solver_verbose=False, h_stoich_spacing=10.e3, num_procs=1,
                                            integration_args=None, n_defect_st=32, extend_defect_dim=False,
                                            diss_rate_log_scaled=True, wave=1):
    if isinstance(flamelet_specs, dict):
        flamelet_specs = FlameletSpec(**flamelet_specs)
    m, fuel, oxy = flamelet_specs.mech_spec, flamelet_specs.fuel_stream, flamelet_specs.oxy_stream
    cput00 = _write_library_header('nonadiabatic (defect) SLFM', m, fuel, oxy, verbose)
    ugt = _build_unstructured_nonadiabatic_defect_slfm_library(flamelet_specs, heat_loss_expansion, diss_rate_values,
                                                               diss_rate_ref, verbose, solver_verbose, h_stoich_spacing,
                                                               num_procs, integration_args, wave=wave)
    _t0 = perf_counter()
    table, x_values, g_values = _interpolate_to_structured_defect_dimension(ugt, n_defect_st, verbose=verbose,
                                                                            extend=extend_defect_dim)
    LAST_BUILD_TIMES['structured_interpolation_s'] = perf_counter() - _t0
    _t0 = perf_counter()
    key0 = list(table.keys())[0]
    z_values = table[key0][_mixture_fraction_name]
    output_library = Library(Dimension(_mixture_fraction_name, z_values),
                             Dimension(_dissipation_rate_name + _stoich_suffix, x_values, diss_rate_log_scaled),
                             Dimension(_enthalpy_defect_name + _stoich_suffix, g_values))
    output_library.extra_attributes['mech_spec'] = m
    for quantity in table[key0]:
        values = output_library.get_empty_dataset()
        for ix, x in enumerate(x_values):
            for ig, g in enumerate(g_values):
                values[:, ix, ig] = table[(x, g)][quantity]
        output_library[quantity] = values
    LAST_BUILD_TIMES['assembly_s'] = perf_counter() - _t0
    _write_library_footer(cput00, verbose)
    return output_library


def build_nonadiabatic_defect_transient_slfm_library(flamelet_specs, diss_rate_values=np.logspace(-3, 2, 16),
                                                     diss_rate_ref='stoichiometric', verbose=True, solver_verbose=False,
                                                     h_stoich_spacing=10.e3, num_procs=1, integration_args=None,
                                                     n_defect_st=32, extend_defect_dim=False, diss_rate_log_scaled=True,
                                                     wave=1):
    """SLFM library with heat loss through the enthalpy defect, the heat-loss profiles generated by rapid transient
    extinction (tabulation.py:727-783)"""
    return _build_nonadiabatic_defect_slfm_library(flamelet_specs, 'transient', diss_rate_values, diss_rate_ref, verbose,
                                                   solver_verbose, h_stoich_spacing, num_procs, integration_args,
                                                   n_defect_st, extend_defect_dim, diss_rate_log_scaled, wave=wave)


def build_nonadiabatic_defect_steady_slfm_library(flamelet_specs, diss_rate_values=np.logspace(-3, 2, 16),
                                                  diss_rate_ref='stoichiometric', verbose=True, solver_verbose=False,
                                                  h_stoich_spacing=10.e3, num_procs=1, integration_args=None,
                                                  n_defect_st=32, extend_defect_dim=False, diss_rate_log_scaled=True,
                                                  wave=1):
    """SLFM library with heat loss through the enthalpy defect, the heat-loss profiles generated by quasi-steady
    extinction (tabulation.py:786-842)"""
    return _build_nonadiabatic_defect_slfm_library(flamelet_specs, 'steady', diss_rate_values, diss_rate_ref, verbose,
                                                   solver_verbose, h_stoich_spacing, num_procs, integration_args,
                                                   n_defect_st, extend_defect_dim, diss_rate_log_scaled, wave=wave)
