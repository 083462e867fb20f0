"""
Multi-GPU partition of the library sweeps: one process per GPU (torchrun), whole flamelets owned by one rank.

The reference's only parallel construct is `multiprocessing.Pool.starmap` over the stoichiometric dissipation rates of
the non-adiabatic builder with the results collected through a `Manager().dict()` (tabulation.py:542-568). Here the same
units of work (one heat-loss expansion per chi_st, one member of a wave of adiabatic flamelets) are dealt block-cyclically
to the ranks of `torch.distributed` -- high-chi members near extinction take more iterations, cyclic assignment
balances them -- and the per-rank result dictionaries are exchanged once at the end with an all-gather over NCCL
(NVLink) on GPU boxes or gloo in the CPU tests. There is no collective inside the data path.
"""
import os


def _dist():
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


def rank():
    d = _dist()
    return d.get_rank() if d else 0


def world_size():
    d = _dist()
    return d.get_world_size() if d else 1


def init_from_env(backend=None):
    """initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / MASTER_*) if it is present and
    not initialised yet; binds the rank to its GPU. Returns (rank, world_size)."""
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    if 'RANK' not in os.environ or int(os.environ.get('WORLD_SIZE', '1')) < 2:
        return 0, 1
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dist.init_process_group(backend=backend)
    return dist.get_rank(), dist.get_world_size()


def my_share(items):
    """the items (a sequence) this rank owns: block-cyclic by position"""
    r, w = rank(), world_size()
    return [x for k, x in enumerate(items) if k % w == r]


def gather_dicts(local):
    """merge the per-rank dictionaries on every rank (the `Manager().dict()` of the reference). Keys must be unique
    across ranks."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return dict(local)
    parts = [None] * d.get_world_size()
    d.all_gather_object(parts, dict(local))
    merged = dict()
    for p in parts:
        merged.update(p)
    return merged


def gather_profile_dicts(local):
    """gather_dicts for dictionaries {key: {name: 1-D float64 array}} whose entries all carry the same names and array
    length (the per-(chi_st, defect) profile sets of the non-adiabatic builders: ~1e5 arrays of one grid line each).
    Pickling that many small arrays costs seconds; here a rank's profiles travel as ONE array [entry, name, point] next
    to the key and name lists, and the merged dictionary holds views into the gathered arrays. Falls back to
    gather_dicts when the entries are not uniform."""
    import numpy as np
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return dict(local)
    keys = list(local.keys())
    names, npts, uniform = [], 0, True
    if keys:
        names = list(local[keys[0]].keys())
        npts = int(np.asarray(local[keys[0]][names[0]]).size) if names else 0
        for k in keys:
            e = local[k]
            if list(e.keys()) != names or any(np.asarray(e[q]).shape != (npts,) for q in names):
                uniform = False
                break
    # one small object per rank (keys, names, sizes), then the numbers as one tensor collective
    metas = [None] * d.get_world_size()
    d.all_gather_object(metas, (bool(uniform), keys, names, npts))
    if not all(m[0] for m in metas):
        return gather_dicts(local)
    import torch
    live = [m for m in metas if m[1]]
    if not live:
        return dict()
    nn, nz = len(live[0][2]), live[0][3]
    if any(len(m[2]) != nn or m[3] != nz for m in live):
        return gather_dicts(local)
    nmax = max(len(m[1]) for m in metas)
    pack = np.zeros((nmax, nn, nz))
    for i, k in enumerate(keys):
        e = local[k]
        for j, q in enumerate(names):
            pack[i, j] = e[q]
    dev = torch.device('cuda', torch.cuda.current_device()) if d.get_backend() == 'nccl' else torch.device('cpu')
    mine = torch.from_numpy(pack).to(dev)
    parts = [torch.empty_like(mine) for _ in metas]
    d.all_gather(parts, mine)
    merged = dict()
    for (_, pkeys, pnames, _), part in zip(metas, parts):
        ppack = part.cpu().numpy()
        for i, k in enumerate(pkeys):
            merged[k] = {q: ppack[i, j] for j, q in enumerate(pnames)}
    return merged


def barrier():
    d = _dist()
    if d is not None:
        d.barrier()


def finalize():
    """tear the process group down (torchrun scripts call this last)"""
    d = _dist()
    if d is not None:
        d.barrier()
        d.destroy_process_group()
