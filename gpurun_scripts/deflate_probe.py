"""Device-resident compress throughput per level and corpus kind (CUDA events), output checked
against the oracle on the distinct streams.   usage: deflate_probe.py LEVELS N [kinds]"""
import ctypes as C
import os
import sys

sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import torch

import corpus
import oracle_lib as o
import libdeflate_rsx_b200 as b

levels = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [6]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
kinds = sys.argv[3].split(',') if len(sys.argv) > 3 else ["corpusA", "text", "binary", "lowent", "mixedB"]
GEN = {"text": corpus.text_stream, "binary": corpus.binary_stream, "lowent": corpus.lowentropy_stream,
       "mixedB": corpus.corpus_b_stream, "corpusA": lambda k: corpus.corpus_a_stream(k % 16), "periodic": corpus.periodic_stream}
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
ctx = b.Context(0)
D = 64
for kind in kinds:
    plain = [GEN[kind](k) for k in range(D)]
    tile = torch.from_numpy(np.frombuffer(b"".join(plain), dtype=np.uint8).copy()).to(dev)
    d_in = tile.repeat(n // D)
    d_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * 65536
    bound = b.compress_bound(0, 65536)
    d_out = torch.empty(n * bound, dtype=torch.uint8, device=dev)
    d_ooff = torch.arange(n, dtype=torch.int64, device=dev) * bound
    d_size = torch.zeros(n, dtype=torch.int64, device=dev)
    d_stat = torch.zeros(n, dtype=torch.int32, device=dev)
    for lvl in levels:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        best = None
        for it in range(3):
            d_stat.fill_(-1)
            ev[0].record(stream)
            ctx.check(ctx._lib.bdf_compress_batch_device(ctx.handle, lvl, 0, d_in.data_ptr(), d_off.data_ptr(), n, d_out.data_ptr(),
                                                         d_ooff.data_ptr(), d_size.data_ptr(), d_stat.data_ptr(), C.c_void_p(stream.cuda_stream)))
            ev[1].record(stream)
            torch.cuda.synchronize(dev)
            ms = ev[0].elapsed_time(ev[1])
            best = ms if best is None or ms < best else best
        sizes = d_size.cpu().numpy()
        stat = d_stat.cpu().numpy()
        bad = []
        for k in list(range(D)) + [n - 1]:
            got = d_out[k * bound:k * bound + int(sizes[k])].cpu().numpy().tobytes()
            exp = o.compress(plain[k % D], lvl)
            if stat[k] != 0 or got != exp:
                bad.append(k)
        print(f"{kind:8s} L{lvl}: {n * 65536 / best / 1e6:8.2f} GB/s  ms {best:8.2f} ratio {n * 65536 / max(int(sizes.sum()), 1):6.2f}  "
              f"{'ok' if not bad and (stat == 0).all() else 'BAD ' + str(bad[:8])}", flush=True)
