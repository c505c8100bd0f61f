import sys, zlib, time
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import corpus, oracle_lib as o
import libdeflate_rsx_b200 as b
ctx = b.default_context()
# checksum
bufs = corpus.small_cases() + [corpus.corpus_a_stream(0), corpus.text_stream(1), b"\xff"*100000]
for kind, ref in ((b.ADLER32, lambda x: zlib.adler32(x)), (b.CRC32, lambda x: zlib.crc32(x))):
    got = b.checksum_batch(bufs, kind)
    exp = [ref(x) for x in bufs]
    print("checksum", kind, "OK" if got == exp else [(i,hex(g),hex(e)) for i,(g,e) in enumerate(zip(got,exp)) if g!=e])
# inflate
streams = bufs + [corpus.binary_stream(2), corpus.lowentropy_stream(3), corpus.offset_stream(3), corpus.offset_stream(1), corpus.offset_stream(32)]
for fmt, wb in ((b.RAW,-15),(b.ZLIB,15),(b.GZIP,31)):
    comp = []
    for s in streams:
        for lvl in (1,6,9):
            c = zlib.compressobj(lvl, zlib.DEFLATED, wb); comp.append((c.compress(s)+c.flush(), s))
        oc = o.compress(s, 6, fmt)
        if oc is not None: comp.append((oc, s))
        oc = o.compress(s, 1, fmt)
        if oc is not None: comp.append((oc, s))
        if len(s): comp.append((o.compress(s, 0, fmt), s))
    d = b.BatchDecompressor(format=fmt)
    res = d.decompress_batch([c for c,_ in comp], [len(s) for _,s in comp])
    bad = [i for i,(r,(c,s)) in enumerate(zip(res,comp)) if r != s]
    print("inflate fmt", fmt, len(comp), "streams", "OK" if not bad else ("BAD", bad[:10], [ (None if res[i] is None else len(res[i])) for i in bad[:10]]))
# errors
d = b.BatchDecompressor()
good = o.compress(b"hello hello hello hello", 6)
r = d.decompress_batch([bytes([0,1,2,3,4,5]), bytes([0,1,2,3]), good, good, good[:-2], b""], [100,100,100,5,100,10])
print("errors:", [None if x is None else len(x) for x in r])
# level 0 compress
c0 = b.BatchCompressor(0)
outs = c0.compress_batch(streams)
print("L0 compress", "OK" if all(x == (o.compress(s,0) or b"") for x,s in zip(outs,streams)) else "BAD")
for fmt in (b.ZLIB,b.GZIP):
    outs = b.BatchCompressor(0, format=fmt).compress_batch(streams)
    print("L0 fmt",fmt, "OK" if all(x == o.compress(s,0,fmt) for x,s in zip(outs,streams)) else "BAD")
# perf: corpus A zlib
n = 16384
base = [o.compress(corpus.corpus_a_stream(k), 6, b.ZLIB) for k in range(16)]
ins = [base[k%16] for k in range(n)]
flat, off = b.flatten(ins)
dz = b.BatchDecompressor(format=b.ZLIB)
for it in range(3):
    t=time.time(); out, out_off, out_size, status = dz.decompress_flat(flat, off, np.full(n, 65536, dtype=np.uint64)); dt=time.time()-t
    print("decompress", n, "streams status ok:", int((status==0).sum()), "kernel ms", ctx.last_kernel_ms, "GB/s kernel", n*65536/ctx.last_kernel_ms/1e6, "e2e s", dt)
exp = np.frombuffer(b"".join(corpus.corpus_a_stream(k) for k in range(16)), dtype=np.uint8)
print("content ok:", all(np.array_equal(out[i*65536:(i+1)*65536], exp[(i%16)*65536:(i%16+1)*65536]) for i in range(0,n,97)))
# text perf
base = [o.compress(corpus.text_stream(k), 6, b.ZLIB) for k in range(64)]
ins = [base[k%64] for k in range(n)]
flat, off = b.flatten(ins)
for it in range(2):
    out, out_off, out_size, status = dz.decompress_flat(flat, off, np.full(n, 65536, dtype=np.uint64))
    print("text decompress ok:", int((status==0).sum()), "kernel ms", ctx.last_kernel_ms, "GB/s kernel", n*65536/ctx.last_kernel_ms/1e6)
