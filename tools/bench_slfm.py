"""SLFM library build timings (BASELINE configs 4-5) -- secondary metric "SLFM library build wall time".

    python tools/bench_slfm.py [--backend gpu|reference|port] [--nchi 64] [--nz 128] [--wave 8] [--nonadiabatic K]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_slfm.py --nonadiabatic 8

Prints one JSON line per measurement (rank 0). GRI-3.0 methane (300 K) / air (300 K), 1 atm, clustered grid,
chi_st in logspace(-3, 2, nchi) (the reference's default range), defaults otherwise. --nonadiabatic K: transient
heat-loss expansion of the first K burning members (dealt to the ranks), n_defect_st = 16.
With --backend reference|port the same host code drives the CPU oracle (the reference's C++ built from
/root/reference, or its C restatement): that is the CPU baseline of this metric."""
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
from spitfire_b200 import tabulation as tab  # noqa: E402
from spitfire_b200.flamelet import FlameletSpec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--backend', default='gpu')
    ap.add_argument('--nchi', type=int, default=64)
    ap.add_argument('--nz', type=int, default=128)
    ap.add_argument('--wave', type=int, default=8)
    ap.add_argument('--nonadiabatic', type=int, default=0)
    ap.add_argument('--mech', default='methane-gri30')
    ap.add_argument('--profile-ops', action='store_true', help='synchronise around every batched operation and '
                    'print where the time goes (slows the build down)')
    args = ap.parse_args()
    rank, world = parallel.init_from_env()
    acc = dict()
    if args.profile_ops:
        import torch
        from spitfire_b200 import flamelet as fl

        def wrap(name):
            f = getattr(fl._BatchOps, name)

            def g(self, *a, **k):
                if self.on_device:
                    torch.cuda.synchronize()
                t = time.perf_counter()
                r = f(self, *a, **k)
                if self.on_device:
                    torch.cuda.synchronize()
                e = acc.setdefault(name, [0, 0.])
                e[0] += 1
                e[1] += time.perf_counter() - t
                return r
            setattr(fl._BatchOps, name, g)
        for nm in ('rhs', 'jac', 'factorize', 'solve', 'jac_and_eig', 'add_to_block_diagonal'):
            wrap(nm)
    m = build_mech(args.mech, args.backend)
    air = m.stream(stp_air=True)
    fuel = m.stream('TPX', (300., 101325., 'CH4:1' if 'methane' in args.mech or 'gri' in args.mech else 'H2:1'))
    chis = np.logspace(-3, 2, args.nchi)

    def specs():
        return FlameletSpec(mech_spec=m, oxy_stream=air, fuel_stream=fuel, grid_points=args.nz)

    def say(d):
        if rank == 0:
            d.update(backend=args.backend, mechanism=args.mech, grid_points=args.nz, n_gpus=world if args.backend == 'gpu' else 0)
            print(json.dumps(d), flush=True)

    if args.nonadiabatic == 0:
        for wave in sorted({1, args.wave}):
            t0 = time.perf_counter()
            lib = tab.build_adiabatic_slfm_library(specs(), diss_rate_values=chis, verbose=False, wave=wave)
            say(dict(metric='adiabatic SLFM library build wall time', unit='s', value=time.perf_counter() - t0,
                     higher_is_better=False, n_chi_requested=args.nchi, n_chi_burning=int(lib.shape[1]), wave=wave,
                     T_max=float(lib['temperature'].max())))
    else:
        t0 = time.perf_counter()
        lib = tab.build_nonadiabatic_defect_transient_slfm_library(specs(), diss_rate_values=chis[:args.nonadiabatic],
                                                                   verbose=False, n_defect_st=16, wave=args.wave)
        say(dict(metric='nonadiabatic (transient defect) SLFM library build wall time', unit='s',
                 value=time.perf_counter() - t0, higher_is_better=False, n_chi=args.nonadiabatic, n_defect_st=16,
                 shape=list(lib.shape), wave=args.wave))
    if args.profile_ops and rank == 0:
        print(json.dumps({k: dict(calls=v[0], seconds=round(v[1], 3)) for k, v in acc.items()}), flush=True)
    parallel.finalize()


if __name__ == '__main__':
    main()
