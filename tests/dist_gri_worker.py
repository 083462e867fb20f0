"""worker of test_gri_transient_library_two_ranks_equals_one_rank: run under torchrun; builds the GRI-3.0 transient
heat-loss library of gri_slfm_cases.py on the GPU path with the chi_st values dealt to the ranks (NCCL when every rank
has a GPU of its own, else gloo with the ranks sharing cuda:0) and writes rank 0's merged library"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from spitfire_b200 import parallel  # noqa: E402
import gri_slfm_cases as cases  # noqa: E402

out = sys.argv[1]
world_env = int(os.environ.get('WORLD_SIZE', '1'))
backend = 'nccl' if torch.cuda.device_count() >= world_env else 'gloo'
rank, world = parallel.init_from_env(backend)
lib = cases.build_transient('gpu', wave=int(sys.argv[2]) if len(sys.argv) > 2 else 1)
if rank == 0:
    np.savez(out, world=world, backend=backend, **{'dim_' + d: getattr(lib, d + '_values') for d in lib.dim_names},
             **{'prop_' + p: lib[p] for p in lib.props})
parallel.finalize()
