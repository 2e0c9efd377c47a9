"""Multi-GPU sharding of the embarrassingly parallel units of the hot path.

One process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).
The reference's only parallel hook is ``pool.map`` over independent chunks
(scintools/dynspec.py:1715-1719); here the independent units are
  * the curvatures of one eta sweep (each needs only the read-only CS, which
    every rank recomputes from the dynamic spectrum: cheaper than moving it),
  * phase-screen realisations / whole dynamic spectra of a batch.
Units are block-partitioned over ranks, there is NO data-path collective;
the only communication is one all-gather of the per-unit results
(8 bytes per eta) at the end.
"""
import numpy as np


def block_range(n, rank, world):
    """Contiguous block of ``range(n)`` owned by ``rank`` (sizes differ by <= 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def world_info(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _gather_device(group=None, device=None):
    """Device of the all-gather buffers: NCCL has no CPU backend, so under
    one-process-per-GPU the buffers must be CUDA tensors; gloo takes CPU."""
    import torch.distributed as dist
    if device is not None:
        return device
    backend = str(dist.get_backend(group)).lower()
    if "nccl" in backend:
        from . import _device as D
        return D.device()
    return None


def all_gather_blocks(local, n_total, group=None, device=None):
    """All-gather variable-length float64 blocks (block_range layout) into one
    array of length n_total on every rank.  One collective."""
    import torch
    import torch.distributed as dist
    rank, world = world_info(group)
    local = np.ascontiguousarray(local, dtype=np.float64)
    if world == 1:
        return local.copy()
    width = -(-n_total // world)
    buf = torch.full((width,), float("nan"), dtype=torch.float64)
    buf[:local.shape[0]] = torch.from_numpy(local)
    device = _gather_device(group, device)
    if device is not None:
        buf = buf.to(device)
    out = torch.empty((world * width,), dtype=torch.float64, device=buf.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy().reshape(world, width)
    parts = []
    for r in range(world):
        lo, hi = block_range(n_total, r, world)
        parts.append(out[r, :hi - lo])
    return np.concatenate(parts)


def sharded_eta_sweep(sweep_fn, etas, group=None, device=None):
    """Run ``sweep_fn(etas_block) -> eigs_block`` on this rank's block of the
    curvature grid and all-gather the eigenvalues.  ``sweep_fn`` is e.g.
    ``lambda e: ththmod.eta_sweep(cs, tau, fd, e, edges)``."""
    etas = np.asarray(etas, dtype=np.float64)
    rank, world = world_info(group)
    lo, hi = block_range(etas.shape[0], rank, world)
    local = sweep_fn(etas[lo:hi]) if hi > lo else np.zeros(0)
    return all_gather_blocks(local, etas.shape[0], group, device)


def sharded_items(items, group=None):
    """The items (seeds, dynspec indices ...) this rank owns."""
    rank, world = world_info(group)
    lo, hi = block_range(len(items), rank, world)
    return list(items[lo:hi])


def sharded_map(fn, items, width, group=None, device=None):
    """Apply ``fn(item) -> float64 array of length width`` to this rank's block
    of ``items`` and all-gather the rows: every rank gets the [len(items)][width]
    array.  Used for the independent retrieval chunks of
    Dynspec.thetatheta_chunks (the reference's pool.map over chunks,
    scintools/dynspec.py:1815-1828); one collective, rows padded to whole blocks."""
    import torch
    import torch.distributed as dist
    rank, world = world_info(group)
    n = len(items)
    lo, hi = block_range(n, rank, world)
    local = np.full((hi - lo, width), np.nan)
    for k, i in enumerate(range(lo, hi)):
        local[k] = np.asarray(fn(items[i]), dtype=np.float64).reshape(width)
    if world == 1:
        return local
    rows = -(-n // world)
    buf = torch.full((rows, width), float("nan"), dtype=torch.float64)
    buf[:hi - lo] = torch.from_numpy(local)
    device = _gather_device(group, device)
    if device is not None:
        buf = buf.to(device)
    out = torch.empty((world * rows, width), dtype=torch.float64, device=buf.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy().reshape(world, rows, width)
    parts = []
    for r in range(world):
        a, b = block_range(n, r, world)
        parts.append(out[r, :b - a])
    return np.concatenate(parts, axis=0)
