"""
Multi-GPU plumbing: trials shard across ranks (one process per GPU), the only data-path
collective is the all-reduce of the trial-summed cross-spectral matrix.

The reference farms one task per trial to Dask workers and "reduces" through a
lock-serialised read-modify-write of one HDF5 dataset
(syncopy/shared/kwarg_decorators.py:722-735, computational_routine.py:938-942);
here every rank accumulates its partial sum in HBM and one NCCL all-reduce (gloo in the CPU
tests) combines them.  Time-frequency results (keeptrials=True) need no collective at all:
trial k's rows are disjoint (`trial_shard` gives the contiguous block of each rank).
"""
import torch
import torch.distributed as dist


def trial_shard(n_trials, rank, world):
    """Contiguous block [lo, hi) of the (selection-ordered) trial list owned by `rank`."""
    base, rem = divmod(n_trials, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def allreduce_csd(csd_sum, n_trials, group=None):
    """
    In-place SUM all-reduce of a complex64 [nFreq, C, C] partial CSD sum; also reduces the trial
    count.  Returns the global number of trials.
    """
    flat = torch.view_as_real(csd_sum)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    cnt = torch.tensor([float(n_trials)], dtype=torch.float64, device=csd_sum.device)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return int(round(cnt.item()))
