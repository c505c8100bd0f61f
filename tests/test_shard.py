"""CPU tests of the multi-GPU sharding logic: partitioning and a world_size-2
gloo run in which each rank processes its shard (with the oracle standing in
for the GPU) and the gathered result equals the single-process result."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

import corpus
import oracle_lib as o
from libdeflate_rsx_b200 import shard


def test_partition_covers_and_balances():
    rng = np.random.default_rng(3)
    for n, world in ((0, 4), (1, 2), (3, 8), (1000, 2), (1000, 8), (65536, 8)):
        lens = rng.integers(0, 70000, n).astype(np.uint64)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        parts = shard.partition(off, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert b == c and a <= b
        if n >= 1000:
            loads = [int(off[hi] - off[lo]) for lo, hi in parts]
            assert max(loads) - min(loads) <= 2 * 70000
    # equal-size streams split evenly
    off = np.arange(65537, dtype=np.uint64) * np.uint64(65536)
    assert shard.partition(off, 8) == [(i * 8192, (i + 1) * 8192) for i in range(8)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bufs = [corpus.corpus_b_stream(k, 3000 + 100 * k) for k in range(11)] + [b"", b"x"]
    flat, off = o.flatten(bufs)
    ranges = shard.partition(off, world)
    lo, hi = ranges[rank]
    sub, sub_off = shard.take(flat, off, lo, hi)
    out, out_off, out_size, status = o.compress_batch(sub if len(sub) else np.zeros(1, np.uint8), sub_off, 6)
    sizes = shard.gather_results(out_size, len(bufs), ranges)
    stats = shard.gather_results(status, len(bufs), ranges)
    crcs = shard.gather_results(o.checksum_batch(sub if len(sub) else np.zeros(1, np.uint8), sub_off, 1),
                                len(bufs), ranges)
    dist.barrier()
    if rank == 0:
        q.put((sizes.tolist(), stats.tolist(), crcs.tolist()))
    dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_process():
    import zlib
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sizes, stats, crcs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bufs = [corpus.corpus_b_stream(k, 3000 + 100 * k) for k in range(11)] + [b"", b"x"]
    exp = [o.compress(b, 6) for b in bufs]
    assert sizes == [len(e) for e in exp]
    assert stats == [0] * len(bufs)
    assert crcs == [zlib.crc32(b) for b in bufs]


def test_sharded_batch_orders_results_without_a_gpu():
    """Host logic of the single-process multi-GPU mirror: partitioning, one worker per device,
    order-preserving concatenation — with a fake context (no GPU here) and the oracle as the codec."""
    import libdeflate_rsx_b200.batch as batch

    class FakeCtx:
        def __init__(self, device):
            self.device = device

    calls = []

    class FakeCompressor:
        def __init__(self, level, format, ctx):
            self.level, self.format, self.ctx = level, format, ctx

        def compress_batch(self, bufs):
            calls.append((self.ctx.device, len(bufs)))
            return [o.compress(b, self.level, self.format) or b"" for b in bufs]

        def compress_to_size_batch(self, bufs, final_block=True):
            return [o.compress_to_size(b, self.level, final_block) for b in bufs]

    real = batch.BatchCompressor
    batch.BatchCompressor = FakeCompressor
    try:
        sb = shard.ShardedBatch(devices=[0, 1, 2], context_factory=FakeCtx)
        bufs = [corpus.corpus_b_stream(k, 500 + 300 * k) for k in range(17)] + [b""]
        got = sb.compress_batch(bufs, 6, 1)
        assert got == [o.compress(b, 6, 1) for b in bufs]
        assert sorted(d for d, _ in calls) == [0, 1, 2] and sum(c for _, c in calls) == len(bufs)
        assert sb.compress_batch([], 6) == []
        assert sb.compress_to_size_batch(bufs, 6) == [o.compress_to_size(b, 6) for b in bufs]
    finally:
        batch.BatchCompressor = real
