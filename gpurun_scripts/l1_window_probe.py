"""Level 1: whole-window rounds (BDF_L1_WINDOW=1) against one-match rounds, and resident CTAs per SM
(both read per launch), device-resident, output of the distinct streams compared with the oracle.
usage: [VARIANTS=window:ctas,...] [BDF_LIBRARY=...] l1_window_probe.py N [kinds]"""
import ctypes as C
import os
import sys

sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import torch

import corpus
import oracle_lib as o
import libdeflate_rsx_b200 as b

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
kinds = sys.argv[2].split(',') if len(sys.argv) > 2 else ["text", "mixedB", "binary", "lowent", "corpusA"]
GEN = {"text": corpus.text_stream, "binary": corpus.binary_stream, "lowent": corpus.lowentropy_stream,
       "mixedB": corpus.corpus_b_stream, "corpusA": lambda k: corpus.corpus_a_stream(k % 16)}
# window:ctas pairs; window bits: 1 = whole-window rounds, 2 / 4 = prefetch of the next window's buckets into L1 / L2
VARIANTS = [tuple(x.split(":")) for x in os.environ.get("VARIANTS", "0:8,1:8,1:6,1:4,1:3,1:2,0:4").split(",")]
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
ctx = b.Context(0)
D = 64
for kind in kinds:
    plain = [GEN[kind](k) for k in range(D)]
    exp = [o.compress(p, 1) for p in plain]
    tile = torch.from_numpy(np.frombuffer(b"".join(plain), dtype=np.uint8).copy()).to(dev)
    d_in = tile.repeat(n // D)
    d_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * 65536
    bound = b.compress_bound(0, 65536)
    d_out = torch.empty(n * bound, dtype=torch.uint8, device=dev)
    d_ooff = torch.arange(n, dtype=torch.int64, device=dev) * bound
    d_size = torch.zeros(n, dtype=torch.int64, device=dev)
    d_stat = torch.zeros(n, dtype=torch.int32, device=dev)
    res = []
    for win, ctas in VARIANTS:
        os.environ["BDF_L1_WINDOW"] = win
        os.environ["BDF_L1_CTAS_PER_SM"] = ctas
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        best = None
        for it in range(3):
            d_stat.fill_(-1)
            d_out.zero_()
            ev[0].record(stream)
            ctx.check(ctx._lib.bdf_compress_batch_device(ctx.handle, 1, 0, d_in.data_ptr(), d_off.data_ptr(), n, d_out.data_ptr(),
                                                         d_ooff.data_ptr(), d_size.data_ptr(), d_stat.data_ptr(), C.c_void_p(stream.cuda_stream)))
            ev[1].record(stream)
            torch.cuda.synchronize(dev)
            ms = ev[0].elapsed_time(ev[1])
            best = ms if best is None or ms < best else best
        sizes = d_size.cpu().numpy()
        stat = d_stat.cpu().numpy()
        bad = []
        for k in list(range(D)) + [n - 1, n // 2 + 7]:
            got = d_out[k * bound:k * bound + int(sizes[k])].cpu().numpy().tobytes()
            if stat[k] != 0 or got != exp[k % D]:
                bad.append(k)
        ok = not bad and bool((stat == 0).all()) and bool((sizes.reshape(-1, D) == sizes[:D]).all())
        res.append(f"w{win}c{ctas} {n * 65536 / best / 1e6:7.2f} {'ok' if ok else 'BAD ' + str(bad[:6])}")
    print(f"{kind:8s} L1 GB/s | " + " | ".join(res), flush=True)
