"""Ignition-delay table: N isobaric GRI-3.0 methane/air reactors (phi = 1, 1 atm, T0 in [1100, 1900] K) integrated
together on one B200 (HomogeneousReactorBatch). Prints one JSON line.

    python tools/bench_ignition.py [--reactors 4096] [--mech methane-gri30] [--serial K]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_ignition.py   # ranks share the table

--serial K: also time K members with the serial HomogeneousReactor on the same backend (one state per C-ABI call)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402

from common import build_mech  # noqa: E402
from spitfire_b200 import parallel  # noqa: E402
from spitfire_b200.reactors import HomogeneousReactor, HomogeneousReactorBatch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reactors', dest='n', type=int, default=4096)
    ap.add_argument('--mech', default='methane-gri30')
    ap.add_argument('--backend', default='gpu')
    ap.add_argument('--serial', type=int, default=0)
    args = ap.parse_args()
    rank, world = parallel.init_from_env(None if args.backend == 'gpu' else 'gloo')
    m = build_mech(args.mech, args.backend)
    air = m.stream(stp_air=True)
    fuel = m.stream('X', 'CH4:1' if 'gri' in args.mech or 'methane' in args.mech else 'H2:1')
    mix = m.mix_for_equivalence_ratio(1.0, fuel, air)
    mix.TP = 1500., 101325.
    T0 = np.linspace(1100., 1900., args.n)
    b = HomogeneousReactorBatch(HomogeneousReactor(m, mix, 'isobaric', 'adiabatic', 'closed'), T0,
                                np.tile(mix.Y, (args.n, 1)))
    b.compute_ignition_delay(maximum_steps=3)  # warm-up (module load, allocations)
    t0 = time.perf_counter()
    tau = b.compute_ignition_delay()
    wall = time.perf_counter() - t0
    out = dict(metric='ignition-delay table wall time', unit='s', value=wall, n_reactors=args.n, mechanism=args.mech,
               backend=args.backend, n_ranks=world, reactors_per_s=args.n / wall, tau_min=float(np.nanmin(tau)),
               tau_max=float(np.nanmax(tau)), not_ignited=int(np.isnan(tau).sum()))
    if args.serial and rank == 0:
        ks = np.linspace(0, args.n - 1, args.serial).astype(int)
        t0 = time.perf_counter()
        err = 0.
        for k in ks:
            mix.TP = float(T0[k]), 101325.
            s = HomogeneousReactor(m, mix, 'isobaric', 'adiabatic', 'closed').compute_ignition_delay()
            err = max(err, abs(s - tau[k]) / s)
        out.update(serial_s_per_reactor=(time.perf_counter() - t0) / ks.size, max_rel_diff_vs_serial=err)
    if rank == 0:
        print(json.dumps(out), flush=True)
    parallel.finalize()


if __name__ == '__main__':
    main()
