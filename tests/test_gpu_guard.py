"""Out-of-bounds write detection (compute-sanitizer is not available on the GPU pool): outputs are
laid out with canary gaps in front of, between and behind the streams' slots, at odd alignments,
and every gap must be intact after the kernels ran.  Covers the zero-fill-ahead of the inflate
kernel, its 16-byte chunk stores and TMA bulk copies, the bit sinks of the compressors and the
pack kernel.  Device-pointer entry points, buffers owned by torch."""
import ctypes as C
import zlib

import numpy as np
import pytest

import corpus
import oracle_lib as o
import tier

pytestmark = pytest.mark.gpu
CANARY = 0xA5


def _layout(sizes, rng):
    """Slots of exactly `sizes[i]` bytes separated by gaps of 1..97 bytes (so that slot starts hit
    every alignment); returns (offsets, total)."""
    off, at = [], int(rng.integers(1, 50))
    for s in sizes:
        off.append(at)
        at += int(s) + int(rng.integers(1, 98))
    return np.array(off, dtype=np.uint64), at + 64


def _check_gaps(buf, off, sizes, what):
    mask = np.ones(len(buf), dtype=bool)
    for o_, s in zip(off, sizes):
        mask[int(o_):int(o_) + int(s)] = False
    assert (buf[mask] == CANARY).all(), f"{what}: bytes outside the streams' slots were written"


def test_no_writes_outside_output_slots(engine):
    import torch
    dev = torch.device("cuda", 0)
    ctx = engine.default_context()
    lib = ctx._lib
    rng = np.random.default_rng(9)
    t = corpus.text_stream(1)
    plain = [corpus.corpus_a_stream(2), corpus.corpus_a_stream(5)[:40001], t[:33333], bytes(70000), b"ab" * 30000,
             corpus.binary_stream(3, 50001), corpus.lowentropy_stream(1, 65536), b"", b"q", t[:15], (t * 3)[:150001],
             corpus.periodic_stream(7, 65536), bytes([7]) * 300 + t[:100] + bytes([9]) * 40000]
    for fmt in (0, 1, 2):
        # ---------------- decompress: slot = exactly the output size
        comp = [o.compress(p, 6, fmt) for p in plain]
        flat, in_off = o.flatten(comp)
        sizes = [len(p) for p in plain]
        out_off, total = _layout(sizes, rng)
        d_in = torch.from_numpy(flat.copy()).to(dev)
        d_in_off = torch.from_numpy(in_off.view(np.int64).copy()).to(dev)
        d_out = torch.full((total,), CANARY, dtype=torch.uint8, device=dev)
        d_out_off = torch.from_numpy(out_off.view(np.int64).copy()).to(dev)
        d_max = torch.tensor(sizes, dtype=torch.int64, device=dev)
        d_size = torch.zeros(len(plain), dtype=torch.int64, device=dev)
        d_status = torch.full((len(plain),), -1, dtype=torch.int32, device=dev)
        d_sum = torch.zeros(len(plain), dtype=torch.int32, device=dev)
        s = torch.cuda.current_stream(dev)
        ctx.check(lib.bdf_decompress_batch_device(ctx.handle, fmt, d_in.data_ptr(), d_in_off.data_ptr(), len(plain),
                                                  d_out.data_ptr(), d_out_off.data_ptr(), d_max.data_ptr(),
                                                  d_size.data_ptr(), d_sum.data_ptr(), d_status.data_ptr(),
                                                  C.c_void_p(s.cuda_stream) if s.cuda_stream else None))
        torch.cuda.synchronize(dev)
        out = d_out.cpu().numpy()
        assert (d_status.cpu().numpy() == 0).all()
        for i, p in enumerate(plain):
            assert out[int(out_off[i]):int(out_off[i]) + len(p)].tobytes() == p, (fmt, i)
        _check_gaps(out, out_off, sizes, f"inflate fmt {fmt}")
        # one byte too little room everywhere: failures, and still nothing outside the (smaller) slots
        small = [max(x - 1, 0) for x in sizes]
        d_out.fill_(CANARY)
        d_max = torch.tensor(small, dtype=torch.int64, device=dev)
        ctx.check(lib.bdf_decompress_batch_device(ctx.handle, fmt, d_in.data_ptr(), d_in_off.data_ptr(), len(plain),
                                                  d_out.data_ptr(), d_out_off.data_ptr(), d_max.data_ptr(),
                                                  d_size.data_ptr(), d_sum.data_ptr(), d_status.data_ptr(), None))
        torch.cuda.synchronize(dev)
        _check_gaps(d_out.cpu().numpy(), out_off, small, f"inflate fmt {fmt}, short room")
        # ---------------- compress (streams <= 64 KiB through the device entry point)
        ins = [p for p in plain if len(p) <= 65536]
        pflat, pin_off = o.flatten(ins)
        for level in (0, 1, 6, 10):
            bounds = [int(lib.bdf_compress_bound(fmt, len(p))) for p in ins]
            c_off, ctotal = _layout(bounds, rng)
            d_pin = torch.from_numpy(pflat.copy()).to(dev)
            d_pin_off = torch.from_numpy(pin_off.view(np.int64).copy()).to(dev)
            d_c = torch.full((ctotal,), CANARY, dtype=torch.uint8, device=dev)
            d_c_off = torch.from_numpy(c_off.view(np.int64).copy()).to(dev)
            d_csize = torch.zeros(len(ins), dtype=torch.int64, device=dev)
            d_cstat = torch.full((len(ins),), -1, dtype=torch.int32, device=dev)
            ctx.check(lib.bdf_compress_batch_device(ctx.handle, level, fmt, d_pin.data_ptr(), d_pin_off.data_ptr(),
                                                    len(ins), d_c.data_ptr(), d_c_off.data_ptr(), d_csize.data_ptr(),
                                                    d_cstat.data_ptr(), None))
            torch.cuda.synchronize(dev)
            cbuf = d_c.cpu().numpy()
            csize = d_csize.cpu().numpy()
            for i, p in enumerate(ins):
                tier.check_stream(cbuf[int(c_off[i]):int(c_off[i]) + int(csize[i])].tobytes(), p, o.compress(p, level, fmt), level, fmt, i)
            _check_gaps(cbuf, c_off, bounds, f"deflate level {level} fmt {fmt}")
            # pack the results densely: the pack kernel must stay inside its destination too
            dense_sizes = [int(x) for x in csize]
            d_off_np = np.zeros(len(ins) + 1, dtype=np.uint64)
            d_off_np[1:] = np.cumsum(dense_sizes)
            lead = 37
            d_dense = torch.full((int(d_off_np[-1]) + lead + 64,), CANARY, dtype=torch.uint8, device=dev)
            d_doff = torch.from_numpy((d_off_np + np.uint64(lead)).view(np.int64).copy()).to(dev)
            ctx.check(lib.bdf_gather_streams_device(ctx.handle, d_c.data_ptr(), d_c_off.data_ptr(), d_csize.data_ptr(),
                                                    len(ins), d_dense.data_ptr(), d_doff.data_ptr(), None))
            torch.cuda.synchronize(dev)
            dense = d_dense.cpu().numpy()
            assert (dense[:lead] == CANARY).all() and (dense[lead + int(d_off_np[-1]):] == CANARY).all()
            assert dense[lead:lead + int(d_off_np[-1])].tobytes() == b"".join(
                cbuf[int(c_off[i]):int(c_off[i]) + dense_sizes[i]].tobytes() for i in range(len(ins)))
