"""Row-slab decomposition helpers: rank g owns grid rows [g*n/P, (g+1)*n/P) of each of
the fields [u_n | v_n | u_s | v_s | p]; a rank's vector is the concatenation of its field slabs."""
from __future__ import annotations

import numpy as np
import torch


def slab_rows(n: int, rank: int, nranks: int):
    """(first row, number of rows) owned by `rank`."""
    if n % nranks:
        raise ValueError(f"n={n} is not divisible by nranks={nranks}")
    rows = n // nranks
    return rank * rows, rows


def scatter_slab(x_global, n: int, nfields: int, rank: int, nranks: int):
    """The rank's slab vector cut out of a global [nfields * n * n] vector (numpy or torch)."""
    r0, rows = slab_rows(n, rank, nranks)
    v = x_global.reshape(nfields, n, n)[:, r0:r0 + rows, :]
    return v.reshape(-1).clone() if isinstance(v, torch.Tensor) else np.ascontiguousarray(v).reshape(-1)


def assemble_global(slabs, n: int, nfields: int):
    """Inverse of scatter_slab for a list of per-rank slab vectors (rank order)."""
    nranks = len(slabs)
    rows = n // nranks
    parts = [s.reshape(nfields, rows, n) for s in slabs]
    if isinstance(parts[0], torch.Tensor):
        return torch.cat(parts, dim=1).reshape(-1)
    return np.concatenate(parts, axis=1).reshape(-1)


def gather_slabs(x_local: torch.Tensor, n: int, nfields: int, nranks: int):
    """All-gather the ranks' slab vectors into the global vector (torch.distributed collective)."""
    import torch.distributed as dist
    outs = [torch.empty_like(x_local) for _ in range(nranks)]
    dist.all_gather(outs, x_local.contiguous())
    return assemble_global(outs, n, nfields)
