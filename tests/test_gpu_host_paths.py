"""Host-call plumbing of the C ABI: the three compress entry points (bound-spaced slots, packed
result, scattered input) give the oracle's bytes; the host decompress call writes nothing outside
what the streams produced; launches on different caller streams do not share scratch."""
import ctypes as C
import zlib

import numpy as np
import pytest

import corpus
import oracle_lib as o
import tier

pytestmark = pytest.mark.gpu


def batch():
    rng = np.random.default_rng(17)
    bufs = corpus.small_cases() + [
        corpus.text_stream(1), corpus.binary_stream(2, 40000), corpus.lowentropy_stream(3, 65536),
        corpus.corpus_a_stream(4), b"", rng.integers(0, 256, 30000, dtype=np.uint8).tobytes(),    # incompressible: fails at level 1
        (corpus.text_stream(5) * 5)[:300001], corpus.periodic_stream(6, 65535), b"x",
    ]
    return bufs


@pytest.mark.parametrize("level,fmt", [(0, 0), (1, 0), (1, 2), (6, 1), (9, 0), (10, 0)])
def test_three_compress_entry_points(engine, level, fmt):
    bufs = batch()
    c = engine.BatchCompressor(level, format=fmt)
    exp = [o.compress(b, level, fmt) for b in bufs]
    first = c.compress_batch(bufs)                         # scattered input, packed result
    for g, b, e in zip(first, bufs, exp):
        tier.check_stream(g, b, e, level, fmt)
    assert c.compress_batch_slots(bufs) == first           # flat input, bound-spaced slots
    flat, off = engine.flatten(bufs)
    out, out_off, status = c.compress_dense(flat, off)      # flat input, packed result
    got = [out[int(out_off[i]):int(out_off[i + 1])].tobytes() for i in range(len(bufs))]
    assert got == first
    assert all((status[i] == 0) == (first[i] != b"" or (len(b) == 0 and level == 0)) for i, b in enumerate(bufs))


def test_scattered_input_larger_than_a_staging_buffer(engine):
    """One 70 MiB buffer between small ones: it crosses the 64 MiB pinned staging buffers."""
    big = (corpus.text_stream(2) * 1121)[:70 * 1024 * 1024 + 13]
    bufs = [b"head", big, corpus.binary_stream(1, 5000), b""]
    got = engine.BatchCompressor(1).compress_batch(bufs)
    assert [zlib.decompress(g, -15) for g in got] == bufs
    assert got[2] == o.compress(bufs[2], 1) and got[0] == o.compress(bufs[0], 1)
    assert len(got[1]) == len(o.compress(big, 1))


def test_dense_output_too_small_is_an_argument_error(engine):
    flat, off = engine.flatten([corpus.text_stream(1), corpus.text_stream(2)])
    ctx = engine.default_context()
    out = np.empty(1000, dtype=np.uint8)
    out_off = np.zeros(3, dtype=np.uint64)
    status = np.zeros(2, dtype=np.int32)
    rc = ctx._lib.bdf_compress_batch_host_dense(ctx.handle, 6, 0, flat.ctypes.data, off.ctypes.data, 2,
                                                out.ctypes.data, 1000, out_off.ctypes.data, status.ctypes.data)
    assert rc == engine.E_ARG and int(out_off[2]) > 1000        # the needed size is reported


def test_host_decompress_writes_only_what_was_produced(engine):
    """Slots with gaps between them, capacities larger than the output, one failing stream: every
    byte outside [out_off[i], out_off[i] + out_size[i]) keeps the caller's fill pattern."""
    plain = [corpus.text_stream(1, 5000), corpus.corpus_a_stream(2, 70000), b"", corpus.binary_stream(3, 333)]
    comp = [o.compress(p, 6, 1) for p in plain]
    bad = bytearray(comp[0]); bad[40] ^= 0x10
    comp.append(bytes(bad)); plain.append(None)
    caps = np.array([6000, 70000, 10, 400, 5000], dtype=np.uint64)
    out_off = np.array([7, 7000, 80000, 80011, 90003], dtype=np.uint64)
    out = np.full(100000, 0xAA, dtype=np.uint8)
    flat, off = engine.flatten(comp)
    size = np.zeros(5, dtype=np.uint64)
    status = np.zeros(5, dtype=np.int32)
    sums = np.zeros(5, dtype=np.uint32)
    ctx = engine.default_context()
    ctx.check(ctx._lib.bdf_decompress_batch_host(ctx.handle, 1, flat.ctypes.data, off.ctypes.data, 5, out.ctypes.data,
                                                 out_off.ctypes.data, caps.ctypes.data, size.ctypes.data,
                                                 sums.ctypes.data, status.ctypes.data))
    keep = np.ones(100000, dtype=bool)
    for i, p in enumerate(plain):
        if p is None:
            assert status[i] != 0 and size[i] == 0
            continue
        assert status[i] == 0 and int(size[i]) == len(p)
        a = int(out_off[i])
        assert out[a:a + len(p)].tobytes() == p
        keep[a:a + len(p)] = False
    assert (out[keep] == 0xAA).all()


def test_huge_capacities_are_rejected_not_wrapped(engine):
    c = o.compress(b"abc" * 10, 6)
    flat, off = engine.flatten([c, c])
    ctx = engine.default_context()
    caps = np.array([16, 2 ** 64 - 16], dtype=np.uint64)
    out_off = np.array([0, 16], dtype=np.uint64)
    out = np.zeros(64, dtype=np.uint8)
    size = np.zeros(2, dtype=np.uint64)
    status = np.zeros(2, dtype=np.int32)
    rc = ctx._lib.bdf_decompress_batch_host(ctx.handle, 0, flat.ctypes.data, off.ctypes.data, 2, out.ctypes.data,
                                            out_off.ctypes.data, caps.ctypes.data, size.ctypes.data, None, status.ctypes.data)
    assert rc == engine.E_ARG
    with pytest.raises(engine.BdfError):
        engine.BatchDecompressor().decompress_batch([c, c], [16, 2 ** 63])


@pytest.mark.parametrize("level", [1, 6, 10])
def test_two_caller_streams_do_not_share_scratch(engine, level):
    """Two compress calls and one size call enqueued back to back on different CUDA streams
    through one ctx (the kernels work in ctx-owned slabs): all three results are the oracle's."""
    import torch
    dev = torch.device("cuda", 0)
    ctx = engine.Context(0)
    lib = ctx._lib
    n = 96 if level >= 10 else 600
    sets = [[corpus.corpus_b_stream(k, 20000 + 37 * k) for k in range(n)],
            [corpus.corpus_b_stream(1000 + k, 30000 - 11 * k) for k in range(n)]]
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    keep, outs = [], []
    for bufs, st in zip(sets, streams):
        flat, off = engine.flatten(bufs)
        bound = np.array([engine.compress_bound(0, len(b)) for b in bufs], dtype=np.uint64)
        ooff = engine.exclusive_offsets(bound)
        d = {"in": torch.from_numpy(flat).to(dev), "off": torch.from_numpy(off.view(np.int64)).to(dev),
             "ooff": torch.from_numpy(ooff.view(np.int64)).to(dev),
             "out": torch.zeros(int(bound.sum()), dtype=torch.uint8, device=dev),
             "size": torch.zeros(n, dtype=torch.int64, device=dev), "stat": torch.full((n,), -1, dtype=torch.int32, device=dev)}
        keep.append(d); outs.append((bufs, ooff))
    d_est = torch.zeros(n, dtype=torch.int64, device=dev)
    d_est_stat = torch.full((n,), -1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    for d, st in zip(keep, streams):
        ctx.check(lib.bdf_compress_batch_device(ctx.handle, level, 0, d["in"].data_ptr(), d["off"].data_ptr(), n,
                                                d["out"].data_ptr(), d["ooff"].data_ptr(), d["size"].data_ptr(),
                                                d["stat"].data_ptr(), C.c_void_p(st.cuda_stream)))
    ctx.check(lib.bdf_compress_size_batch_device(ctx.handle, level, keep[0]["in"].data_ptr(), keep[0]["off"].data_ptr(), n, 1,
                                                 d_est.data_ptr(), d_est_stat.data_ptr(), C.c_void_p(streams[2].cuda_stream)))
    torch.cuda.synchronize(dev)
    for d, (bufs, ooff) in zip(keep, outs):
        size = d["size"].cpu().numpy(); stat = d["stat"].cpu().numpy(); out = d["out"].cpu().numpy()
        for i, b in enumerate(bufs):
            assert stat[i] == 0, i
            tier.check_stream(out[int(ooff[i]):int(ooff[i]) + int(size[i])].tobytes(), b, o.compress(b, level), level, 0, i)
    est = d_est.cpu().numpy()
    assert (d_est_stat.cpu().numpy() == 0).all()
    assert [int(v) for v in est[:20]] == [o.compress_to_size(b, level) for b in sets[0][:20]]
    ctx.close()
