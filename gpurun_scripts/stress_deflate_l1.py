"""Large randomised differential batch for the level-1 compressor (whole-window rounds, deflate_l1.cuh): tens of thousands of
streams per call, lengths 0 .. 64 KiB (the window rounds) with a share of longer ones (the block-split rounds), contents chosen
to stress the window walk — text, records, small alphabets, short periods (same-hash lower lanes), periods around 258 and
around 32, runs with noise, incompressible bytes; every framing.  Sizes, status and bytes must equal the oracle's, stream by
stream.   usage: stress_deflate_l1.py [n_streams] [seed]"""
import os
import sys
import time

sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np

import corpus
import oracle_lib as o
import libdeflate_rsx_b200 as bdf

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)


def periodic(period, noise):
    a = np.tile(rng.integers(0, 256, period, dtype=np.uint8), 70000 // period + 1)[:70000].copy()
    if noise:
        idx = rng.integers(0, 70000, noise)
        a[idx] = rng.integers(0, 256, noise, dtype=np.uint8)
    return a.tobytes()


pool = [corpus.text_stream(k, 70000) for k in range(10)] + [corpus.binary_stream(k, 70000) for k in range(10)] + \
       [corpus.lowentropy_stream(k, 70000) for k in range(8)] + [corpus.periodic_stream(k, 70000) for k in range(6)] + \
       [corpus.corpus_a_stream(k, 70000) for k in range(6)] + [rng.integers(0, 256, 70000, dtype=np.uint8).tobytes() for _ in range(3)] + \
       [bytes(rng.integers(0, a, 70000, dtype=np.uint8)) for a in (2, 2, 3, 4, 4, 6, 16)] + \
       [periodic(p, z) for p in (1, 2, 3, 4, 5, 7, 16, 29, 31, 32, 33, 64, 100, 255, 256, 257, 258, 259, 260, 516, 1000) for z in (0, 60)]
bad_total = 0
for fmt in [int(x) for x in os.environ.get("STRESS_FORMATS", "0,1,2").split(",")]:
    t0 = time.time()
    bufs = []
    for i in range(n):
        base = pool[int(rng.integers(0, len(pool)))]
        r = int(rng.integers(0, 40))
        if r == 0:
            ln = int(rng.integers(65537, 200000))            # one 256 KiB unit: the block-split rounds
        elif r < 4:
            ln = int(rng.choice([0, 1, 2, 3, 4, 31, 32, 33, 34, 35, 63, 64, 65, 257, 258, 259, 65533, 65534, 65535, 65536]))
        elif r < 20:
            ln = int(rng.integers(0, 3000))
        else:
            ln = int(rng.integers(0, 65537))
        st = int(rng.integers(0, 4096))
        s = (base * (ln // len(base) + 2))[st:st + ln]
        bufs.append(s)
    t1 = time.time()
    got = bdf.BatchCompressor(1, format=fmt).compress_batch(bufs)
    t2 = time.time()
    flat, off = o.flatten(bufs)
    eout, eoff, esize, est = o.compress_batch(flat, off, 1, fmt)
    bad = 0
    for i, g in enumerate(got):
        exp = b"" if est[i] != 0 else eout[int(eoff[i]):int(eoff[i]) + int(esize[i])].tobytes()
        if g != exp:
            bad += 1
            if bad <= 5:
                print("MISMATCH", fmt, i, len(bufs[i]), len(g), len(exp), int(est[i]))
    nfail = sum(1 for e in est if e != 0)
    print(f"format {fmt}: {n} streams ({sum(map(len, bufs)) >> 20} MiB), {nfail} failed in-band as expected, mismatches {bad}  "
          f"(gen {t1 - t0:.1f} s, gpu call {t2 - t1:.2f} s)", flush=True)
    bad_total += bad
ctx = bdf.default_context()
chk = int(ctx._lib.bdf_debug_check_failures(ctx.handle))       # -1: not the -DBDF_CHECK build; 0: no assertion failed
print("device assertions:", "not a check build" if chk == -1 else ("none failed" if chk == 0 else f"FAILED at line {chk & 0x7FFFFFFF}"))
if chk > 0:
    bad_total += 1
print("STRESS", "OK" if bad_total == 0 else "FAILED")
sys.exit(1 if bad_total else 0)
