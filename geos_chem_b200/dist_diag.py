"""The only collective on the chemistry path: reduce per-rank integrator diagnostics.

The reference reduces an error count (OpenMP REDUCTION(+:errorCount), fullchem_mod.F90:543) and
raises a global Failed2x flag (:1245-1269, :1557); State_Diag%Kpp* counters stay per cell.
Here each rank owns a block of (I,J) columns and the same quantities are summed across ranks
with one small all-reduce (NCCL on GPUs, gloo in the CPU tests): cells, failed cells, sum of
Nstp/Nacc/Nrej, max Nstp, and 64-bin histograms of Nstp and Nrej.
"""
import numpy as np


def reduce_diagnostics(istatus, ierr, device=None, nbins=64):
    """istatus [8, ncell], ierr [ncell] (numpy or torch) -> dict of global diagnostics (valid on every rank)"""
    import torch
    import torch.distributed as dist
    ist = torch.as_tensor(istatus)
    ier = torch.as_tensor(ierr)
    dev = device or ist.device
    ist, ier = ist.to(dev), ier.to(dev)
    nstp, nacc, nrej = ist[2].to(torch.int64), ist[3].to(torch.int64), ist[4].to(torch.int64)
    sums = torch.stack([torch.tensor(ier.numel(), device=dev, dtype=torch.int64), (ier < 0).sum(),
                        nstp.sum(), nacc.sum(), nrej.sum()])
    hist = torch.cat([torch.bincount(nstp.clamp(max=nbins - 1), minlength=nbins),
                      torch.bincount(nrej.clamp(max=nbins - 1), minlength=nbins)])
    mx = nstp.max() if nstp.numel() else torch.tensor(0, device=dev)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums)
        dist.all_reduce(hist)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    s = sums.cpu().tolist()
    h = hist.cpu().numpy()
    return {"cells": s[0], "failed": s[1], "sum_nstp": s[2], "sum_nacc": s[3], "sum_nrej": s[4],
            "max_nstp": int(mx), "hist_nstp": h[:nbins].tolist(), "hist_nrej": h[nbins:].tolist()}
