#!/usr/bin/env python
"""
bench.py -- headline benchmark of the B200 Griffon hot path (contract in the task statement).

Metric (BASELINE.json): batched isobaric-reactor RHS + analytical Jacobian, thermochemical states per second, on
BASELINE config 3: GRI-3.0 methane/air (53 species, 325 reactions), 1,048,576 synthetic states per GPU
(spitfire_b200.synthetic, seed 20241017), p = 1 atm, closed adiabatic reactor. One "step" = one pass of
`reactor_jac_isobaric_batch` (which returns both the RHS and the Jacobian, like the reference's
`reactor_jac_isobaric`) over the whole batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--states S] [--mech NAME]

  default arm      : `value` = device-resident throughput (inputs in HBM, CUDA-event timed, max over ranks);
                     `e2e`   = same metric through the host-buffer C-ABI entry point (pinned host memory in, results
                               copied back to the host every step);
                     `roofline` for k_jac against the measured HBM peak; `cpu_baseline` = the reference's CPU path on
                     this box's host cores (bounded sample).
  --impl reference : the reference's own CPU implementation (oracle/_ref when present, else the oracle port) on all
                     host cores, bounded sample per step.
Under torchrun (N > 1) every rank owns its own batch (weak scaling, no collective in the data path).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# one thread per process on the host side: the CPU arm runs one process per core (OPENBLAS_NUM_THREADS=1)
os.environ.setdefault('OMP_NUM_THREADS', '1')
os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
os.environ.setdefault('MKL_NUM_THREADS', '1')

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402

METRIC = 'reactor RHS+Jacobian states/sec (GRI-3.0)'
UNIT = 'states/s'
PRESSURE = 101325.


def load_mech_data(name):
    with open(os.path.join(ROOT, 'tests', 'golden', 'mech', name + '.json')) as f:
        return json.load(f)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---- CPU arm (the only place bench.py executes oracle/) ------------------------------------------------------------
def _cpu_worker(args):
    kind, mech_name, fuel, lo, n, reps, keep = args
    from oracle.oracle import OracleKernels
    from spitfire_b200.mechanism import ChemicalMechanismSpec
    from spitfire_b200.synthetic import synthetic_states
    m = ChemicalMechanismSpec(mech_data=load_mech_data(mech_name), griffon_factory=lambda: OracleKernels(kind))
    ns = m.n_species
    state, _ = synthetic_states(m.species_names, lo + n, fuel)
    state = np.ascontiguousarray(state[lo:lo + n])
    rhs, jac = np.zeros((n, ns)), np.zeros((n, ns * ns))
    m.griffon.reactor_jac_isobaric_many(state[:8], PRESSURE, 0, rhs[:8], jac[:8])  # touch
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        m.griffon.reactor_jac_isobaric_many(state, PRESSURE, 0, rhs, jac)
        times.append(time.perf_counter() - t0)
    # the first `keep` states' outputs go back to the parent: they check the GPU results of the same states (parity)
    return times, (rhs[:keep].copy(), jac[:keep].copy()) if keep else None


class CpuArm(object):
    """the reference's single-state `reactor_jac_isobaric` looped over a sample of the workload on every host core
    (one process per core, like the reference's own multiprocessing mode, tabulation.py:542-556)"""

    def __init__(self, mech_name, fuel, per_core_states):
        from oracle import oracle
        self.kind = 'reference' if oracle.available('reference') else 'port'
        self.mech_name, self.fuel = mech_name, fuel
        self.cores = host_cores()
        self.per_core = per_core_states
        # load the library in the parent first: the forked workers inherit the mapping, and the driver's record of the
        # shared objects this process loaded then shows which CPU implementation was timed
        oracle._load(self.kind)
        self.lib_path = oracle.lib_path(self.kind)
        import multiprocessing as mp
        self.pool = mp.get_context('fork').Pool(self.cores)
        self.reference_outputs = None

    def run(self, reps):
        """returns the list over `reps` of wall seconds for cores*per_core states (all cores busy concurrently)"""
        jobs = [(self.kind, self.mech_name, self.fuel, c * self.per_core, self.per_core, reps,
                 min(self.per_core, 1024) if c == 0 else 0) for c in range(self.cores)]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
        per_worker = [r[0] for r in res]
        self.reference_outputs = res[0][1]
        # per repetition the slowest worker bounds the throughput
        step_times = [max(w[k] for w in per_worker) for k in range(reps)]
        return step_times, wall

    def close(self):
        self.pool.close()
        self.pool.join()

    @property
    def sample(self):
        return (f'{self.cores * self.per_core} states of the workload per step ({self.per_core} per core), single-state '
                f'reactor_jac_isobaric looped in C, one process per core')


def _oracle_mech(name, kind):
    from oracle.oracle import OracleKernels
    from spitfire_b200.mechanism import ChemicalMechanismSpec
    return ChemicalMechanismSpec(mech_data=load_mech_data(name), griffon_factory=lambda: OracleKernels(kind))


def _gri_flamelet_specs(m):
    from spitfire_b200.flamelet import FlameletSpec
    air = m.stream(stp_air=True)
    fuel = m.stream('TPX', (300., PRESSURE, 'CH4:1'))
    return FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=128)


def library_fields(lib):
    names = ['temperature'] + [k for k in lib.props if k.startswith('mass fraction ')]
    return names


def library_parity(lib, ref):
    """largest difference of the temperature and mass-fraction fields of two libraries of equal shape, each field
    relative to its own largest reference value"""
    if list(lib.shape) != list(ref.shape):
        return {'error': f'shape {list(lib.shape)} vs {list(ref.shape)}'}
    worst, worst_name = 0., None
    for k in library_fields(ref):
        a, b = np.asarray(lib[k]), np.asarray(ref[k])
        scale = float(np.max(np.abs(b)))
        if scale <= 1e-12:  # (species that never form, e.g. AR in CH4/air: rounding noise around zero)
            continue
        e = float(np.max(np.abs(a - b))) / scale
        if e > worst:
            worst, worst_name = e, k
    eT = float(np.max(np.abs(np.asarray(lib['temperature']) - np.asarray(ref['temperature'])) /
                      np.asarray(ref['temperature'])))
    return {'T_max_rel': eT, 'fields_max_rel_to_field_scale': worst, 'worst_field': worst_name,
            'n_fields': len(library_fields(ref))}


def cpu_library_sample(args, kind, cores):
    """CPU baseline of the secondary metric on a bounded sample of BASELINE configs 4-5: every `stride`-th dissipation
    rate of the configuration, the SAME host code driving the reference's CPU kernels, the heat-loss expansions in a
    process pool over the dissipation rates (the reference's num_procs mode, tabulation.py:542-568)."""
    from spitfire_b200 import tabulation as tab
    m = _oracle_mech('methane-gri30', kind)
    chis = np.logspace(-3, 2, args.library_chi)[::args.library_cpu_stride]
    t0 = time.perf_counter()
    ad = tab.build_adiabatic_slfm_library(_gri_flamelet_specs(m), diss_rate_values=chis, verbose=False)
    t_ad = time.perf_counter() - t0
    t0 = time.perf_counter()
    na = tab.build_nonadiabatic_defect_transient_slfm_library(_gri_flamelet_specs(m), diss_rate_values=chis,
                                                              verbose=False, n_defect_st=16, num_procs=cores)
    t_na = time.perf_counter() - t0
    return {'chis': chis, 'adiabatic': ad, 'nonadiabatic': na,
            'report': {'kind': kind, 'cores': cores,
                       'sample': f'every {args.library_cpu_stride}th chi_st of the configuration '
                                 f'({chis.size} requested, {ad.shape[1]} burn), same host code on the reference CPU '
                                 f'kernels; reference-order chain, heat-loss expansions in a pool of {cores} processes',
                       'adiabatic_slfm_s': t_ad, 'nonadiabatic_defect_slfm_s': t_na,
                       'adiabatic_shape': list(ad.shape), 'nonadiabatic_shape': list(na.shape)}}


def config1_ignition(backend):
    """BASELINE config 1: H2/air phi = 1, 1200 K, 1 atm, closed adiabatic isobaric reactor, integrate_to_steady with
    defaults, against the reference's gold trajectory (tests/reactor/closed_reactors, rtol 1e-4)"""
    from reactor_cases import compare_with_gold
    from spitfire_b200.reactors import HomogeneousReactor
    from common import build_mech
    m = build_mech('h2-burke', backend)
    air = m.stream(stp_air=True)
    fuel = m.stream('X', 'H2:1')
    mix = m.mix_for_equivalence_ratio(1.0, fuel, air)
    mix.TP = 1200., 101325.
    r = HomogeneousReactor(m, mix, 'isobaric', 'adiabatic', 'closed')
    t0 = time.perf_counter()
    lib = r.integrate_to_steady()
    wall = time.perf_counter() - t0
    err = compare_with_gold(m, lib, 'adiabatic')
    return {'wall_s': wall, 'steps': int(lib.time_values.size), 'gold_T_max_rel': err, 'gold_rtol': 1e-4,
            'T_end': float(lib['temperature'][-1])}


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    arm = CpuArm(args.mech, args.fuel, args.cpu_states_per_core)
    step_times, _ = arm.run(args.warmup + args.steps)
    arm.close()
    timed = step_times[args.warmup:]
    n = arm.cores * arm.per_core
    ms = 1e3 * float(np.mean(timed))
    value = n / (ms * 1e-3)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, n_states=n),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': arm.cores, 'kind': arm.kind, 'sample': arm.sample,
                         'library': os.path.relpath(arm.lib_path, ROOT),
                         'note': 'a rate: the sample is a prefix of the same seeded workload, not the whole batch'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm -------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """Multi-rank runs: keep this rank's threads -- and with them, by first touch, its pinned host buffers -- on the NUMA
    node its GPU hangs off, so that the end-to-end leg (6 GB device->host per step and rank) does not cross the socket
    interconnect. Best effort: returns the node, or None if the topology is not visible or the cpuset forbids it."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), 'pci_domain_id', 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), 'pci_device_id', 0)
        path = '/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node' % (dom, bus, dev)
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            cpus = set()
            for part in f.read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def workload_config(args, n_states):
    return {'workload': f'BASELINE config 3: GRI-3.0 methane/air batched isobaric reactor RHS+analytical Jacobian '
                        f'({args.mech}, closed adiabatic, 1 atm)',
            'mechanism': args.mech, 'states_per_gpu': n_states, 'seed': 20241017,
            'l2': 'inputs_and_outputs_larger_than_L2',
            'parallelism': f'independent state shards x{args.gpus}, no data-path collective'}


class ClockSampler(object):
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nme)
        if not sm:
            return None
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(np.max(smax)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def measured_hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic_per_launch(mech, n_states):
    """dram bytes of one k_jac launch from the committed ncu capture (profiles/roofline_traffic.json), scaled to the
    launch size; None if no capture has been committed"""
    p = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f)
    e = d.get(mech)
    if not e:
        return None
    return float(e['dram_bytes']) / float(e['states']) * n_states


def time_library_builds(args, world, cpu_sample=None):
    """secondary metric of BASELINE.json, "SLFM library build wall time": BASELINE configs 4 and 5 (GRI-3.0, CH4 300 K /
    air 300 K, 1 atm, 128-point clustered grid, chi_st in logspace(-3, 2, 64), 16 enthalpy defects) through the
    repo's public builders, heat-loss trajectories dealt to the ranks. Every rank must call this (one gather).
    cpu_sample (N = 1): the CPU leg's libraries of a sub-sampled configuration; the same sample is rebuilt on the GPU
    and compared field by field (`parity`)."""
    from spitfire_b200 import tabulation as tab
    from spitfire_b200.mechanism import ChemicalMechanismSpec
    m = ChemicalMechanismSpec(mech_data=load_mech_data('methane-gri30'))
    chis = np.logspace(-3, 2, args.library_chi)
    out = {'config': f'GRI-3.0 CH4/air 300 K 1 atm, {args.library_chi} chi_st in logspace(-3,2), 128-point grid',
           'n_gpus': world, 'wave': 8}
    t0 = time.perf_counter()
    lib = tab.build_adiabatic_slfm_library(_gri_flamelet_specs(m), diss_rate_values=chis, verbose=False, wave=8)
    out['adiabatic_slfm_s'] = time.perf_counter() - t0
    out['adiabatic_shape'] = list(lib.shape)
    t0 = time.perf_counter()
    lib = tab.build_nonadiabatic_defect_transient_slfm_library(
        _gri_flamelet_specs(m), diss_rate_values=chis, verbose=False, n_defect_st=16, wave=8)
    out['nonadiabatic_defect_slfm_s'] = time.perf_counter() - t0
    out['nonadiabatic_shape'] = list(lib.shape)
    try:  # this rank's heat-loss trajectories: rounds of the asynchronous integrator, seconds inside it
        from spitfire_b200.time import batched as _batched
        out['trajectories'] = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in _batched.LAST_ASYNC_STATS.items()}
        out['stages_s'] = {k: round(v, 4) for k, v in tab.LAST_BUILD_TIMES.items()}
    except Exception:
        pass
    out['T_max'] = float(lib['temperature'].max())
    if cpu_sample is not None:
        sc = cpu_sample['chis']
        same = {}
        for wave in (1, 8):
            t0 = time.perf_counter()
            ad = tab.build_adiabatic_slfm_library(_gri_flamelet_specs(m), diss_rate_values=sc, verbose=False, wave=wave)
            t_ad = time.perf_counter() - t0
            t0 = time.perf_counter()
            na = tab.build_nonadiabatic_defect_transient_slfm_library(
                _gri_flamelet_specs(m), diss_rate_values=sc, verbose=False, n_defect_st=16, wave=wave)
            t_na = time.perf_counter() - t0
            same[f'wave{wave}'] = {'adiabatic_slfm_s': t_ad, 'nonadiabatic_defect_slfm_s': t_na,
                                   'parity_adiabatic': library_parity(ad, cpu_sample['adiabatic']),
                                   'parity_nonadiabatic': library_parity(na, cpu_sample['nonadiabatic'])}
        out['cpu_baseline'] = cpu_sample['report']
        out['gpu_same_sample'] = same
    return out


def run_gpu_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    # CPU legs first, before any CUDA context exists in this process (fork-safe); rank 0 at N=1 only
    cpu_baseline, cpu_ref_out, cpu_lib, config1 = None, None, None, {}
    if world == 1 and not args.no_cpu_baseline:
        try:
            arm = CpuArm(args.mech, args.fuel, args.cpu_states_per_core)
            st, _ = arm.run(2)
            arm.close()
            n_cpu = arm.cores * arm.per_core
            cpu_baseline = {'value': n_cpu / st[-1], 'unit': UNIT, 'cores': arm.cores, 'kind': arm.kind,
                            'sample': arm.sample, 'library': os.path.relpath(arm.lib_path, ROOT)}
            cpu_ref_out = arm.reference_outputs
            if args.library_chi > 0:
                cpu_lib = cpu_library_sample(args, arm.kind, arm.cores)
            config1['cpu_reference'] = dict(config1_ignition(arm.kind), kind=arm.kind, cores=1)
        except Exception as e:  # the checker is optional for the measurement itself
            if cpu_baseline is None:
                cpu_baseline = {'value': None, 'unit': UNIT, 'cores': host_cores(), 'kind': 'unavailable',
                                'sample': f'failed: {e!r}'[:300]}
            else:
                cpu_baseline['secondary_failed'] = repr(e)[:300]

    import torch
    import torch.distributed as dist
    from spitfire_b200 import griffon
    from spitfire_b200.mechanism import ChemicalMechanismSpec
    from spitfire_b200.synthetic import synthetic_states

    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    mech = ChemicalMechanismSpec(mech_data=load_mech_data(args.mech))
    g, ns = mech.griffon, mech.n_species
    n = args.states
    # every rank owns a different slice of the seeded sequence
    state_np, _ = synthetic_states(mech.species_names, n, args.fuel, seed=20241017 + rank)
    d_state = torch.from_numpy(state_np).cuda()
    d_rhs = torch.empty((n, ns), dtype=torch.float64, device='cuda')
    d_jac = torch.empty((n, ns * ns), dtype=torch.float64, device='cuda')

    def step():
        g.reactor_jac_isobaric_batch(d_state, PRESSURE, d_rhs, d_jac)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = griffon.kernel_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    barrier()
    launches = griffon.kernel_launch_count() - launches0
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    clocks = sampler.stop()

    # ---- end to end through the host-buffer entry point: pinned host state in, rhs + jac back to the host ----------
    e2e_states = min(n, args.e2e_states if (world > 1 or args.e2e_states_set) else n)
    h_state = torch.empty((e2e_states, ns), dtype=torch.float64, pin_memory=True)
    h_state.copy_(torch.from_numpy(state_np[:e2e_states]))
    h_rhs = torch.empty((e2e_states, ns), dtype=torch.float64, pin_memory=True)
    h_jac = torch.empty((e2e_states, ns * ns), dtype=torch.float64, pin_memory=True)
    hs, hr, hj = h_state.numpy(), h_rhs.numpy(), h_jac.numpy()
    g.reactor_jac_isobaric_batch(hs, PRESSURE, hr, hj)
    barrier()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        g.reactor_jac_isobaric_batch(hs, PRESSURE, hr, hj)  # synchronous: returns when the results are on the host
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    check = float(hr[0, 0])  # read of the step's result on the host

    # ---- parity of the timed kernel on the states the CPU leg evaluated (same seeded inputs) ------------------------
    parity = None
    if cpu_ref_out is not None:
        from cases import error_stats
        k = cpu_ref_out[0].shape[0]
        parity = {'states': int(k), 'against': cpu_baseline['kind'],
                  'rhs': error_stats(d_rhs[:k].cpu().numpy(), cpu_ref_out[0]),
                  'jac': error_stats(d_jac[:k].cpu().numpy(), cpu_ref_out[1]),
                  'e2e_equals_device': bool(np.array_equal(hj[:k], d_jac[:k].cpu().numpy()) and
                                            np.array_equal(hr[:k], d_rhs[:k].cpu().numpy()))}
    # ---- second roofline bound: FP64 pipe, measured on this device ---------------------------------------
    fp64_peak = None
    if rank == 0:
        try:
            fp64_peak = griffon.measure_fp64_peak(0)[0]
        except Exception:
            fp64_peak = None
    # ---- BASELINE config 2 (H2, 1M states) and config 1 (one ignition through the public reactor class) -------------
    config2 = None
    if rank == 0 and not args.no_extra_configs and args.mech != 'h2-burke':
        try:
            del h_jac, hj
            mh = ChemicalMechanismSpec(mech_data=load_mech_data('h2-burke'))
            gh, nh = mh.griffon, mh.n_species
            sh, _ = synthetic_states(mh.species_names, n, 'H2', seed=20241017)
            dsh = torch.from_numpy(sh).cuda()
            drh = torch.empty((n, nh), dtype=torch.float64, device='cuda')
            djh = torch.empty((n, nh * nh), dtype=torch.float64, device='cuda')
            for _ in range(3):
                gh.reactor_jac_isobaric_batch(dsh, PRESSURE, drh, djh)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                gh.reactor_jac_isobaric_batch(dsh, PRESSURE, drh, djh)
            e1.record()
            torch.cuda.synchronize()
            msh = e0.elapsed_time(e1) / args.steps
            bh = 8 * (nh * nh + 2 * nh)
            config2 = {'workload': 'BASELINE config 2: h2-burke, 1 atm, closed adiabatic', 'states': n,
                       'value': n / (msh * 1e-3), 'unit': UNIT, 'ms_per_step': msh,
                       'hbm_gbs': bh * n / (msh * 1e-3) / 1e9}
            del dsh, drh, djh
            config1['gpu'] = config1_ignition('gpu')
        except Exception as e:
            config2 = {'error': repr(e)[:300]}

    t_ms = torch.tensor([total_ms / args.steps, e2e_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step, e2e_ms = float(t_ms[0]), float(t_ms[1])
    library = None
    if args.library_chi > 0:
        try:
            if world > 1:
                from spitfire_b200 import parallel  # the builders find the process group bench.py initialised
                assert parallel.world_size() == world
            library = time_library_builds(args, world, cpu_lib)
        except Exception as e:  # the headline line must survive a failure of the secondary metric
            library = {'error': repr(e)[:300]}
    if rank == 0:
        value = world * n / (ms_per_step * 1e-3)
        peak, peak_src = measured_hbm_peak()
        bytes_per_state = 8 * (ns * ns + 2 * ns)  # ns in, ns + ns^2 out
        kernel_ms = float(np.mean(step_ms))
        achieved = bytes_per_state * n / (kernel_ms * 1e-3) / 1e9
        # the kernel is measured against both roofs, HBM (8*(ns^2+2ns) bytes per state) and the FP64 pipe
        # (flop model of the sparse-exact formulation, transcendental calls not counted); the larger fraction is the bound
        flops_per_state = {'methane-gri30': 0.093e6, 'h2-burke': 5.7e3}.get(args.mech)
        hbm_frac = achieved / peak
        roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': hbm_frac,
                    'traffic': ncu_traffic_per_launch(args.mech, n), 'kernel': 'k_jac',
                    'algorithmic_bytes_per_launch': bytes_per_state * n, 'kernel_ms': kernel_ms,
                    'peak_source': peak_src}
        if fp64_peak and flops_per_state:
            tf = flops_per_state * n / (kernel_ms * 1e-3) / 1e12
            roofline['fp64'] = {'achieved': tf, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': tf / fp64_peak,
                                'flops_per_state': flops_per_state,
                                'peak_source': 'measured now (gb_measure_fp64_peak: dependent-free DFMA, all SMs)'}
            roofline['frac_max_of_both'] = max(hbm_frac, tf / fp64_peak)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, n),
            'roofline': roofline,
            'cpu_baseline': cpu_baseline,
            'e2e': {'value': world * e2e_states / (e2e_ms * 1e-3), 'unit': UNIT,
                    'h2d_bytes_per_step': int(8 * e2e_states * ns),
                    'd2h_bytes_per_step': int(8 * e2e_states * (ns + ns * ns)),
                    'states_per_step': e2e_states, 'ms_per_step': e2e_ms, 'steps': e2e_steps,
                    'api': 'PyCombustionKernels.reactor_jac_isobaric_batch(numpy) -> gb_reactor_jac_isobaric_host',
                    'result_check': check},
            'gpu_launches': int(launches),
            'parity': parity,
            'config1_h2_ignition': config1 or None,
            'config2_h2_states': config2,
            'library_build': library,
            'numa_node': numa,
            'clocks': clocks,
            'build': load_build_info(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def load_build_info():
    from spitfire_b200 import griffon
    return griffon.load_library().gb_build_info().decode()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--mech', default='methane-gri30')
    ap.add_argument('--fuel', default=None)
    ap.add_argument('--states', type=int, default=1 << 20)
    ap.add_argument('--e2e-states', type=int, default=None,
                    help='states per end-to-end step (default: the whole batch at N = 1, 262144 per rank at N > 1)')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--cpu-states-per-core', type=int, default=2048)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra-configs', action='store_true', help='skip the BASELINE config 1 / 2 extras')
    ap.add_argument('--library-cpu-stride', type=int, default=4,
                    help='the CPU leg of the library metric builds every stride-th dissipation rate')
    ap.add_argument('--library-chi', type=int, default=64,
                    help='dissipation rates of the SLFM library builds timed as the secondary metric (0 = skip)')
    args = ap.parse_args()
    args.e2e_states_set = args.e2e_states is not None
    if args.e2e_states is None:
        args.e2e_states = 1 << 18
    if args.fuel is None:
        args.fuel = 'H2' if args.mech.startswith('h2') else 'CH4'
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
