"""bdf_compress_batch_device_any: any stream length with every buffer on the device
(Compressor::compress takes any length, reference src/compress/mod.rs:699-772).  Byte-identical to the
oracle, like the host call."""
import ctypes as C
import os
import sys
import zlib

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as o  # noqa: E402
from test_gpu_checksum_compress import WBITS, large_inputs  # noqa: E402

pytestmark = pytest.mark.gpu


def compress_on_device(bdf, ins, level, fmt, entry="bdf_compress_batch_device_any"):
    import torch
    dev = torch.device("cuda", 0)
    ctx = bdf.default_context()
    n = len(ins)
    flat, off = bdf.flatten(ins)
    bounds = np.array([(bdf.compress_bound(fmt, len(s)) + 15) & ~15 for s in ins], dtype=np.uint64)
    out_off = np.zeros(n, dtype=np.uint64)
    out_off[1:] = np.cumsum(bounds)[:-1]
    d_in = torch.from_numpy(flat.copy()).to(dev)
    d_off = torch.from_numpy(off.view(np.int64).copy()).to(dev)
    d_out = torch.zeros(int(bounds.sum()) + 16, dtype=torch.uint8, device=dev)
    d_ooff = torch.from_numpy(out_off.view(np.int64).copy()).to(dev)
    d_size = torch.zeros(n, dtype=torch.int64, device=dev)
    d_stat = torch.full((n,), -1, dtype=torch.int32, device=dev)
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        ctx.check(getattr(ctx._lib, entry)(ctx.handle, level, fmt, d_in.data_ptr(), d_off.data_ptr(), n, d_out.data_ptr(),
                                           d_ooff.data_ptr(), d_size.data_ptr(), d_stat.data_ptr(), C.c_void_p(s.cuda_stream)))
    s.synchronize()
    sizes, stat, outb = d_size.cpu().numpy(), d_stat.cpu().numpy(), d_out.cpu().numpy()
    return [outb[int(out_off[i]):int(out_off[i]) + int(sizes[i])].tobytes() if stat[i] == 0 else None for i in range(n)], stat


def test_any_length_byte_identical():
    import libdeflate_rsx_b200 as bdf
    ins = large_inputs()
    for fmt, level in ((0, 1), (0, 6), (1, 2), (2, 9), (0, 0)):
        got, stat = compress_on_device(bdf, ins, level, fmt)
        for g, s in zip(got, ins):
            exp = o.compress(s, level, fmt)
            assert g == exp, (fmt, level, len(s))
            if exp is not None:
                assert zlib.decompress(g, WBITS[fmt]) == s


def test_short_batches_take_the_plain_path():
    import libdeflate_rsx_b200 as bdf
    ins = [b"abc" * 1000, b"", os.urandom(10) * 500]
    got, stat = compress_on_device(bdf, ins, 6, 1)
    assert got == [o.compress(s, 6, 1) for s in ins]
    # the call that never synchronises marks what it cannot take
    got, stat = compress_on_device(bdf, large_inputs()[:1] + [b"small"], 6, 0, entry="bdf_compress_batch_device")
    assert stat[0] == 100 and got[1] == o.compress(b"small", 6, 0)
