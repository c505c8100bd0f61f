import sys, zlib, time
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import corpus, oracle_lib as o
import libdeflate_rsx_b200 as b
ctx = b.default_context()
levels = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [1, 2, 6, 9]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
streams = corpus.small_cases() + [corpus.corpus_a_stream(0), corpus.text_stream(1), corpus.binary_stream(2), corpus.lowentropy_stream(3), corpus.offset_stream(3), corpus.offset_stream(1), corpus.offset_stream(32), np.random.default_rng(1).integers(0,256,3000,dtype=np.uint8).tobytes()]
for lvl in levels:
    got = b.BatchCompressor(lvl).compress_batch(streams)
    exp = [o.compress(s, lvl) or b"" for s in streams]
    bad = [i for i,(g,e) in enumerate(zip(got,exp)) if g != e]
    print("L%d parity:" % lvl, "OK" if not bad else ("BAD", bad, [(len(got[i]), len(exp[i])) for i in bad]))
    for i in bad[:3]:
        g, e = got[i], exp[i]
        k = next((j for j in range(min(len(g), len(e))) if g[j] != e[j]), min(len(g), len(e)))
        print("   stream", i, "len", len(streams[i]), "first diff at byte", k, g[max(0,k-4):k+8].hex(), e[max(0,k-4):k+8].hex())
        try:
            print("   zlib inflate of ours ok:", zlib.decompress(g, -15) == streams[i])
        except Exception as ex:
            print("   zlib inflate of ours failed:", ex)
for name, gen in (("corpusA", corpus.corpus_a_stream), ("text", corpus.text_stream), ("mixed", corpus.corpus_b_stream)):
    base = [gen(k) for k in range(16)]
    flat, off = b.flatten([base[k % 16] for k in range(n)])
    for lvl in levels:
        c = b.BatchCompressor(lvl)
        for it in range(2):
            out, out_off, out_size, status = c.compress_flat(flat, off)
        print(f"{name} L{lvl}: ok {int((status==0).sum())}/{n} kernel ms {ctx.last_kernel_ms:.2f} GB/s {n*65536/ctx.last_kernel_ms/1e6:.2f} ratio {n*65536/max(int(out_size.sum()),1):.2f}")
