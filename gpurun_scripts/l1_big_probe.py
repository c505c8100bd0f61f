"""Level 1 on units above 64 KiB (block-split rounds) and in the size estimator: whole-window rounds (BDF_L1_WINDOW bit 3)
against one-match rounds; every output compared with the oracle.  usage: l1_big_probe.py [n_buffers]"""
import os
import sys
import time

sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np

import corpus
import oracle_lib as o
import libdeflate_rsx_b200 as bdf

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rng = np.random.default_rng(5)
gens = [corpus.text_stream, corpus.binary_stream, corpus.lowentropy_stream, corpus.corpus_b_stream]
base = [b"".join(g(4 * k + j) for j in range(4)) for k in range(4) for g in gens]       # 16 x 256 KiB, one kind each
bufs = []
for i in range(n):
    b = base[i % len(base)]
    r = i % 8
    ln = 262144 if r < 4 else int(rng.integers(65537, 262145)) if r < 7 else 262144 * 3 + int(rng.integers(0, 5000))
    bufs.append((b * 4)[:ln])
total = sum(map(len, bufs))
flat, off = o.flatten(bufs)
eout, eoff, esize, est = o.compress_batch(flat, off, 1, 0)
exp = [eout[int(eoff[i]):int(eoff[i]) + int(esize[i])].tobytes() if est[i] == 0 else b"" for i in range(n)]
small = [b[:262144] for b in bufs[:256]]
exp_sz = [o.compress_to_size(b, 1) for b in small]
ctx = bdf.default_context()
for win in os.environ.get("WINDOWS", "3,11,3,11").split(","):
    os.environ["BDF_L1_WINDOW"] = win
    c = bdf.BatchCompressor(1)
    t0 = time.time()
    got = c.compress_batch(bufs)
    t1 = time.time()
    ms = ctx.last_kernel_ms
    bad = [i for i in range(n) if got[i] != exp[i]]
    t2 = time.time()
    sz = c.compress_to_size_batch(small)
    t3 = time.time()
    ms_sz = ctx.last_kernel_ms
    bad_sz = [i for i in range(len(small)) if sz[i] != exp_sz[i]]
    print(f"BDF_L1_WINDOW={win:3s} compress {n} buffers ({total >> 20} MiB): kernels {ms:8.2f} ms = {total / ms / 1e6:6.2f} GB/s, call {t1 - t0:.2f} s, "
          f"{'ok' if not bad else 'BAD ' + str(bad[:6])} | size estimate of {len(small)} x 256 KiB: kernels {ms_sz:7.2f} ms = "
          f"{sum(map(len, small)) / ms_sz / 1e6:6.2f} GB/s {'ok' if not bad_sz else 'BAD ' + str(bad_sz[:6])}", flush=True)
