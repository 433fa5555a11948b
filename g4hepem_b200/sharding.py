"""Multi-GPU host logic: tracks shard across ranks, tables are replicated, one collective at the end.

Tracks are independent (SURVEY.md par. 8e): rank r of W owns a contiguous slice of the primaries / of the batch and
steps it with its own replica of the table arena; there is no data-path collective.  The one exchange of the path
is the final sum of the scored quantities (deposited energy per layer, step counters) over ranks -- the batch
equivalent of TestEm3's Run::Merge (apps/examples/TestEm3/src/Run.cc:146-190) -- done with one all_reduce
(NCCL over NVLink on the GPU box; the same code runs on gloo in the CPU tests)."""
import numpy as np


def shard_bounds(n_total, rank, world):
    """Contiguous slice [lo, hi) of n_total units owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_host_batch(batch, rank, world):
    """The slice of a host batch owned by `rank` (views of the same arrays, ids untouched so that the per-track
    uniform streams -- keyed by track id -- do not depend on the sharding)."""
    lo, hi = shard_bounds(batch.n, rank, world)
    out = type(batch)(hi - lo)
    for g in batch.groups() + ("meta", "winner"):
        getattr(out, g)[...] = getattr(batch, g)[lo:hi]
    return out


def layer_histogram(edep, layer_index, n_layers):
    """Per-layer sum of the deposited energy of one rank (float64, deterministic order: numpy bincount)."""
    return np.bincount(np.asarray(layer_index, dtype=np.int64), weights=np.asarray(edep, dtype=np.float64), minlength=n_layers)


def allreduce_scores(hist, counters=None, dist=None, device=None):
    """Sum the per-rank score arrays over all ranks.  `dist` is torch.distributed (already initialised) or None for
    a single process.  Returns numpy arrays."""
    import torch

    hist = np.asarray(hist, dtype=np.float64)
    counters = np.zeros(0, dtype=np.float64) if counters is None else np.asarray(counters, dtype=np.float64)
    flat = torch.from_numpy(np.concatenate([hist, counters]))
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None:
            flat = flat.to(device)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat = flat.cpu()
    out = flat.numpy()
    return out[: hist.size].copy(), out[hist.size:].copy()
