"""worker of test_two_rank_gloo_ignition_table: run under torchrun; a batch of H2 ignition reactors dealt to the ranks"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from spitfire_b200 import parallel  # noqa: E402
from test_reactor_batch import _template  # noqa: E402
from spitfire_b200.reactors import HomogeneousReactorBatch  # noqa: E402

backend, out = sys.argv[1], sys.argv[2]
rank, world = parallel.init_from_env('gloo' if backend != 'gpu' else None)
m, mix, r = _template(backend)
T0 = np.array([1100., 1150., 1200., 1250., 1300.])
tau = HomogeneousReactorBatch(r, T0, np.tile(mix.Y, (T0.size, 1))).compute_ignition_delay()
if rank == 0:
    np.savez(out, world=world, tau=tau, T0=T0)
parallel.finalize()
