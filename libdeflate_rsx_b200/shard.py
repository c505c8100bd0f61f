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


class ShardedBatch:
    """One process, several GPUs: the batch API of src/batch.rs with the streams split over the
    devices of the box (contiguous ranges balanced by bytes, `partition`), one `bdf_ctx` and one
    host thread per device, results concatenated in stream order.  No collective: the shards never
    talk to each other.  This is what a single-process caller (the Rust crate behind the C ABI) does;
    `bench.py --gpus N` uses one process per GPU instead."""

    def __init__(self, devices=None, context_factory=None):
        from . import _native as N
        from .batch import Context
        if devices is None:
            devices = list(range(N.lib().bdf_device_count()))
        if not devices:
            raise N.BdfError("no CUDA device: the engine has no CPU fallback")
        make = context_factory or Context
        self.contexts = [make(d) for d in devices]

    def _run(self, inputs, fn):
        """fn(ctx, lo, hi) -> list of per-stream results for streams [lo, hi)."""
        from concurrent.futures import ThreadPoolExecutor
        n = len(inputs)
        if n == 0:
            return []
        lens = np.array([len(b) for b in inputs], dtype=np.uint64)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        ranges = partition(off, len(self.contexts))
        with ThreadPoolExecutor(len(self.contexts)) as pool:
            futs = [pool.submit(fn, ctx, lo, hi) if hi > lo else None
                    for ctx, (lo, hi) in zip(self.contexts, ranges)]
            out = []
            for f in futs:
                out += f.result() if f is not None else []
        return out

    def compress_batch(self, inputs, level, format=0):
        from .batch import BatchCompressor
        return self._run(inputs, lambda ctx, lo, hi: BatchCompressor(level, format, ctx).compress_batch(inputs[lo:hi]))

    def compress_to_size_batch(self, inputs, level, final_block=True):
        from .batch import BatchCompressor
        return self._run(inputs, lambda ctx, lo, hi: BatchCompressor(level, 0, ctx).compress_to_size_batch(
            inputs[lo:hi], final_block))

    def decompress_batch(self, inputs, max_out_sizes, format=0):
        from .batch import BatchDecompressor
        n = min(len(inputs), len(max_out_sizes))
        inputs, sizes = list(inputs[:n]), list(max_out_sizes[:n])
        return self._run(inputs, lambda ctx, lo, hi: BatchDecompressor(format, ctx).decompress_batch(
            inputs[lo:hi], sizes[lo:hi]))

    def checksum_batch(self, inputs, kind):
        from .batch import checksum_batch
        return self._run(inputs, lambda ctx, lo, hi: checksum_batch(inputs[lo:hi], kind, ctx))
