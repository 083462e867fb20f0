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
# the packed gather of profile dictionaries: uneven entry counts (rank r holds r + 1 entries), and the fallback for
# entries that are not uniform
prof = {(float(rank), float(k)): {'a': np.full(5, rank + 0.5 * k), 'b': np.arange(5.) * (rank + 1)} for k in range(rank + 1)}
merged = parallel.gather_profile_dicts(prof)
ok_packed = len(merged) == sum(r + 1 for r in range(world)) and all(
    np.array_equal(merged[(float(r), float(k))]['a'], np.full(5, r + 0.5 * k)) and
    np.array_equal(merged[(float(r), float(k))]['b'], np.arange(5.) * (r + 1)) for r in range(world) for k in range(r + 1))
ragged = {(float(rank), 0.): {'a': np.zeros(3 + rank)}}
merged2 = parallel.gather_profile_dicts(ragged)
ok_ragged = all(merged2[(float(r), 0.)]['a'].shape == (3 + r,) for r in range(world))
if rank == 0:
    np.savez(out, world=world, gather_profile_dicts_ok=np.array([ok_packed, ok_ragged]),
             **{f'owned_{r}': np.array(all_owned[r]) for r in range(world)}, **{p: lib[p] for p in lib.props})
parallel.barrier()
