"""Sharding of a batch across the GPUs of one box.

Streams are independent (reference src/batch.rs:34,79: `par_iter` over
streams), so a batch splits into contiguous ranges of stream indices, one per
rank, balanced by input bytes; each rank works on its slice with offsets
rebased to 0 and no collective is needed on the data path.  Only the optional
gather of per-stream results (sizes / status) to every rank uses
torch.distributed, with whatever backend the process group has (NCCL on the
GPUs, gloo in the CPU tests).
"""
import numpy as np


def partition(in_off, world):
    """Contiguous stream ranges [lo, hi) per rank, balanced by bytes.

    in_off: uint64 offsets[n+1].  Returns a list of `world` (lo, hi) pairs that
    cover [0, n) in order; empty ranges are allowed when n < world."""
    in_off = np.asarray(in_off, dtype=np.uint64)
    n = len(in_off) - 1
    total = int(in_off[-1]) if n > 0 else 0
    cuts = [0]
    for r in range(1, world):
        if n == 0:
            cuts.append(0)
            continue
        target = total * r // world
        # first stream whose start offset is >= target; keep cuts monotone
        k = int(np.searchsorted(in_off[:-1], np.uint64(target), side="left"))
        if total == 0:
            k = n * r // world
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def take(flat, in_off, lo, hi):
    """Rank-local view of streams [lo, hi): (flat slice, offsets rebased to 0)."""
    in_off = np.asarray(in_off, dtype=np.uint64)
    base = in_off[lo]
    return flat[int(base):int(in_off[hi])], (in_off[lo:hi + 1] - base).astype(np.uint64)


def gather_results(local, n_total, ranges, group=None):
    """All-gather a per-stream result vector (e.g. out_size or status) so that
    every rank holds the full, order-preserving array of length n_total."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local)
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    width = max(hi - lo for lo, hi in ranges) if ranges else 0
    device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    pad = np.zeros(width, dtype=local.dtype)
    pad[:len(local)] = local
    t = torch.from_numpy(pad.view(np.uint8).copy()).to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    full = np.empty(n_total, dtype=local.dtype)
    for r, (lo, hi) in enumerate(ranges):
        full[lo:hi] = outs[r].cpu().numpy().view(local.dtype)[:hi - lo]
    return full
