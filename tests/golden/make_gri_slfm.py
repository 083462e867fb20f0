"""
Generates tests/golden/ref_gri_slfm_{adiabatic,transient}.npz: the GRI-3.0 SLFM libraries of tests/gri_slfm_cases.py built
by this repository's host code on top of the UNMODIFIED reference C++ kernels (oracle/_ref, compiled from
/root/reference by oracle/build_oracle.py). Run HERE (CPU, a few minutes):

    python tests/golden/make_gri_slfm.py [adiabatic] [adiabatic_tight] [transient]

Prints the wall time of each build on one host core: the reference-order CPU cost of the same work.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')

import gri_slfm_cases as cases  # noqa: E402

which = sys.argv[1:] or ['adiabatic', 'adiabatic_tight', 'transient']
for kind in which:
    t0 = time.perf_counter()
    if kind == 'adiabatic':
        lib = cases.build_adiabatic('reference')
    elif kind == 'adiabatic_tight':
        lib = cases.build_adiabatic('reference', tolerance=cases.TIGHT)
    else:
        lib = cases.build_transient('reference')
    dt = time.perf_counter() - t0
    cases.save_fixture(lib, kind)
    print(f'{kind}: shape {lib.shape}, T_max {lib["temperature"].max():.3f} K, {dt:.1f} s on one core -> '
          f'{cases.fixture_path(kind)}', flush=True)
