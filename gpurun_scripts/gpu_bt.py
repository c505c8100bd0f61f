import sys, time
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import corpus
import libdeflate_rsx_b200 as b
ctx = b.default_context()
n = int(sys.argv[1]); lvl = int(sys.argv[2])
base = [corpus.corpus_b_stream(k) for k in range(16)]
flat, off = b.flatten([base[k % 16] for k in range(n)])
c = b.BatchCompressor(lvl)
for it in range(2):
    out, out_off, out_size, status = c.compress_flat(flat, off)
    print(f"mixed L{lvl}: n={n} ok {int((status==0).sum())} kernel ms {ctx.last_kernel_ms:.1f} GB/s {n*65536/ctx.last_kernel_ms/1e6:.3f} ratio {n*65536/max(int(out_size.sum()),1):.2f}", flush=True)
