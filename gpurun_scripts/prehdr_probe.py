"""Whole decompress call on config 2 with and without the header pre-pass (CUDA events, median of 10).
usage: prehdr_probe.py [n_streams]"""
import ctypes as C
import os
import sys
import zlib

sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import torch

import corpus
import libdeflate_rsx_b200 as b

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
plain = [corpus.corpus_a_stream(k % 16) for k in range(64)]
comp = [zlib.compress(p, 6) for p in plain]
flat, off = b.flatten([comp[k % 64] for k in range(n)])
d_in = torch.from_numpy(flat).to(dev)
d_off = torch.from_numpy(off.view(np.int64)).to(dev)
d_out = torch.empty(n * 65536, dtype=torch.uint8, device=dev)
d_out_off = torch.arange(n, dtype=torch.int64, device=dev) * 65536
d_max = torch.full((n,), 65536, dtype=torch.int64, device=dev)
d_size = torch.zeros(n, dtype=torch.int64, device=dev)
d_stat = torch.zeros(n, dtype=torch.int32, device=dev)
d_sum = torch.zeros(n, dtype=torch.int32, device=dev)
for tag, env in (("default", {}), ("nopre", {"BDF_INFLATE_PREHDR": "0"}), ("default2", {})):
    for k in ("BDF_PREHDR_CARVEOUT", "BDF_INFLATE_PREHDR"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = b.Context(0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ts = []
    for it in range(12):
        ev[0].record(stream)
        ctx.check(ctx._lib.bdf_decompress_batch_device(
            ctx.handle, b.ZLIB, d_in.data_ptr(), d_off.data_ptr(), n, d_out.data_ptr(), d_out_off.data_ptr(),
            d_max.data_ptr(), d_size.data_ptr(), d_sum.data_ptr(), d_stat.data_ptr(), C.c_void_p(stream.cuda_stream)))
        ev[1].record(stream)
        torch.cuda.synchronize(dev)
        ts.append(ev[0].elapsed_time(ev[1]))
    ok = bool((d_stat == 0).all())
    ts = sorted(ts[2:])
    print(f"{tag:9s} call ms: best {ts[0]:.4f} median {ts[len(ts) // 2]:.4f}  -> {n * 65536 / ts[len(ts) // 2] / 1e6:7.1f} GB/s {'ok' if ok else 'BAD'}", flush=True)
