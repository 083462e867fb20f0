"""worker of test_two_rank_gloo_sweep_matches_gold: run under torchrun; builds the reference's transient non-adiabatic
SLFM case with the chi_st values dealt to the ranks and writes rank 0's merged library"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from spitfire_b200 import parallel  # noqa: E402
from slfm_cases import build  # noqa: E402

backend, out = sys.argv[1], sys.argv[2]
rank, world = parallel.init_from_env('gloo' if backend != 'gpu' else None)
owned = parallel.my_share(list(range(4)))
lib = build('nonadiabatic_defect_transient_slfm', backend)
all_owned = parallel.gather_dicts({rank: owned})
if rank == 0:
    np.savez(out, world=world, **{f'owned_{r}': np.array(all_owned[r]) for r in range(world)},
             **{p: lib[p] for p in lib.props})
parallel.barrier()
