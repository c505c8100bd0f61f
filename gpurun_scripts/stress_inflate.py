"""Large randomised differential batch for the decompress path: tens of thousands of small streams in ONE
call (so that the large-batch configuration runs: header pre-pass, lane kernel with tables in global
memory and per-block litlen width, both engines), every framing, several zlib strategies, random output
capacities and corruptions; status and bytes must agree with the oracle stream by stream.
usage: stress_inflate.py [n_streams] [seed]"""
import os
import sys
import time
import zlib

sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np

import corpus
import oracle_lib as o
import libdeflate_rsx_b200 as bdf

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
WBITS = {0: -15, 1: 15, 2: 31}
pool = [corpus.text_stream(k, 65536) for k in range(12)] + [corpus.binary_stream(k, 65536) for k in range(12)] + \
       [corpus.lowentropy_stream(k, 65536) for k in range(8)] + [corpus.periodic_stream(k, 65536) for k in range(8)] + \
       [corpus.corpus_a_stream(k) for k in range(8)] + [rng.integers(0, 256, 65536, dtype=np.uint8).tobytes() for _ in range(4)] + \
       [bytes(rng.choice(np.frombuffer(b"ab\0\xff", dtype=np.uint8), 65536)) for _ in range(4)]
STRATS = [zlib.Z_DEFAULT_STRATEGY] * 5 + [zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]
bad_total = 0
for fmt in [int(x) for x in os.environ.get("STRESS_FORMATS", "1,0,2").split(",")]:
    t0 = time.time()
    streams, caps, plains = [], [], []
    for i in range(n):
        base = pool[int(rng.integers(0, len(pool)))]
        ln = int(rng.integers(0, 9000)) if rng.integers(0, 20) else int(rng.integers(0, 66000))
        st = int(rng.integers(0, 65536 - min(ln, 65535)))
        s = (base + base)[st:st + ln]
        z = zlib.compressobj(int(rng.integers(1, 10)), zlib.DEFLATED, WBITS[fmt], 9, STRATS[int(rng.integers(0, len(STRATS)))])
        c = z.compress(s)
        if rng.integers(0, 4) == 0 and len(s) > 100:
            c += z.flush(zlib.Z_FULL_FLUSH)          # more blocks: later headers are read inside the kernels
            c += z.compress(s[: len(s) // 3])
            s = s + s[: len(s) // 3]
        c += z.flush()
        mode = int(rng.integers(0, 25))
        cap = len(s)
        if mode == 0 and len(s) > 0:
            cap = len(s) - 1
        elif mode == 1:
            cap = len(s) + int(rng.integers(0, 3000))
        elif mode == 2 and len(c) > 8:
            b = bytearray(c)
            b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
            c = bytes(b)
        elif mode == 3 and len(c) > 4:
            c = c[:int(rng.integers(1, len(c)))]
        streams.append(c); caps.append(cap); plains.append(s)
    t1 = time.time()
    got = bdf.BatchDecompressor(format=fmt).decompress_batch(streams, caps)
    t2 = time.time()
    flat, off = o.flatten(streams)
    eout, eoff, esize, est = o.decompress_batch(flat, off, caps, fmt)
    bad = 0
    for i, g in enumerate(got):
        exp = None if est[i] != 0 else eout[int(eoff[i]):int(eoff[i]) + int(esize[i])].tobytes()
        if g != exp:
            bad += 1
            if bad <= 5:
                print("MISMATCH", fmt, i, len(streams[i]), caps[i], None if g is None else len(g), None if exp is None else len(exp), int(est[i]))
    nfail = sum(1 for g in got if g is None)
    print(f"format {fmt}: {n} streams, {nfail} failed in-band as expected, mismatches {bad}  (gen {t1 - t0:.1f} s, gpu call {t2 - t1:.2f} s)", flush=True)
    bad_total += bad
ctx = bdf.default_context()
chk = int(ctx._lib.bdf_debug_check_failures(ctx.handle))       # -1: not the -DBDF_CHECK build; 0: no assertion failed
print("device assertions:", "not a check build" if chk == -1 else ("none failed" if chk == 0 else f"FAILED at line {chk & 0x7FFFFFFF}"))
if chk > 0:
    bad_total += 1
print("STRESS", "OK" if bad_total == 0 else "FAILED")
sys.exit(1 if bad_total else 0)
