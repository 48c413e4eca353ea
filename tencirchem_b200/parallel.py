"""Multi-GPU plumbing for the parameter-set regime: independent theta sets (PES scans,
multi-start kUpCCGSD, ADAPT pools) shard across ranks with no data-path collective; only (E, grad E) are gathered.
One process per GPU, `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) for the plumbing.

The other regime -- ONE CI vector sharded by alpha rows over the GPUs, (18e,18o) / (20e,20o) -- lives in
tencirchem_b200/sharded.py.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of `n_items` work units for `rank`."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_parameter_sets(thetas: np.ndarray, rank: int, world: int) -> np.ndarray:
    lo, hi = shard_bounds(len(thetas), rank, world)
    return np.asarray(thetas)[lo:hi]


def _device_for_backend():
    import torch
    import torch.distributed as dist

    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def reduce_timing_max(value: float) -> float:
    """Device time of a step = MAX over ranks (never a wall clock of one rank)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(value)], dtype=torch.float64, device=_device_for_backend())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_results(energies: np.ndarray, grads: np.ndarray, n_total: int, rank: int, world: int):
    """All ranks receive the energies [n_total] and gradients [n_total, P] in the original order."""
    import torch
    import torch.distributed as dist

    dev = _device_for_backend()
    n_params = grads.shape[1] if grads.ndim == 2 else 0
    e_full = torch.zeros(n_total, dtype=torch.float64, device=dev)
    g_full = torch.zeros((n_total, n_params), dtype=torch.float64, device=dev)
    lo, hi = shard_bounds(n_total, rank, world)
    e_full[lo:hi] = torch.as_tensor(energies, dtype=torch.float64, device=dev)
    if n_params:
        g_full[lo:hi] = torch.as_tensor(grads, dtype=torch.float64, device=dev)
    # disjoint slices: a SUM all-reduce is a gather that works the same on NCCL and gloo
    dist.all_reduce(e_full, op=dist.ReduceOp.SUM)
    dist.all_reduce(g_full, op=dist.ReduceOp.SUM)
    return e_full.cpu().numpy(), g_full.cpu().numpy()


def energy_and_grad_batch(plan, thetas: np.ndarray, rank: int = 0, world: int = 1):
    """Evaluate many parameter sets (API gap (i) of SURVEY.md section 8b: the reference accepts one theta per
    call): each rank evaluates its slice on its own GPU with the device-resident path; results are gathered."""
    thetas = np.asarray(thetas, dtype=np.float64)
    mine = shard_parameter_sets(thetas, rank, world)
    # all sets of the rank's slice share every kernel launch (tcc_energy_and_grad_batch)
    energies, grads = plan.energy_and_grad_batch(mine) if len(mine) else (np.zeros(0), np.zeros((0, plan.n_params)))
    if world == 1:
        return energies, grads
    return gather_results(energies, grads, len(thetas), rank, world)
