"""Device-resident inflate throughput of every engine mode on each corpus kind (CUDA events).
usage: python gpurun_scripts/inflate_modes.py [n_streams] [modes...]"""
import ctypes as C
import os
import sys
import zlib

sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import torch

import corpus
import oracle_lib as o
import libdeflate_rsx_b200 as b

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
want = sys.argv[2:] or ["group", "lane0", "lane1", "auto"]
ENV = {"group": {"BDF_INFLATE_MODE": "group"}, "lane0": {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "0"},
       "lane1": {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "1"}, "lane2": {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "2"}, "auto": {"BDF_INFLATE_MODE": "auto"}, "auto0": {"BDF_INFLATE_MODE": "auto", "BDF_LANE_CFG": "0"},
       # the same without the first-block header pre-pass (inflate_prehdr.cuh)
       "auto_nopre": {"BDF_INFLATE_MODE": "auto", "BDF_INFLATE_PREHDR": "0"},
       "group_nopre": {"BDF_INFLATE_MODE": "group", "BDF_INFLATE_PREHDR": "0"},
       "serial1": {"BDF_INFLATE_MODE": "auto", "BDF_INFLATE_SERIAL": "1"},
       "serial2": {"BDF_INFLATE_MODE": "auto", "BDF_INFLATE_SERIAL": "2"},
       "concurrent": {"BDF_INFLATE_MODE": "auto", "BDF_INFLATE_SERIAL": "0"},
       "lane3": {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "3"}, "lane4": {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "4"},
       "lane5": {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "5"}, "auto5": {"BDF_INFLATE_MODE": "auto", "BDF_LANE_CFG": "5"},
       "lane3w12": {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "3", "BDF_LANE_WARPS": "12"},
       "lane3w10": {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "3", "BDF_LANE_WARPS": "10"},
       "lane0_nopre": {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "0", "BDF_INFLATE_PREHDR": "0"}}
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
KINDS = {
    "text": lambda k: corpus.text_stream(k), "binary": lambda k: corpus.binary_stream(k),
    "lowent": lambda k: corpus.lowentropy_stream(k), "mixedB": lambda k: corpus.corpus_b_stream(k),
    "corpusA": lambda k: corpus.corpus_a_stream(k % 16),
}
ctxs = {}
for m in want:
    for k in ("BDF_INFLATE_MODE", "BDF_LANE_CFG", "BDF_INFLATE_PREHDR", "BDF_INFLATE_SERIAL", "BDF_LANE_WARPS"):
        os.environ.pop(k, None)
    os.environ.update(ENV[m])
    ctxs[m] = b.Context(0)
only = os.environ.get("KINDS")
for kind, gen in KINDS.items():
    if only and kind not in only.split(","):
        continue
    plain = [gen(k) for k in range(64)]
    for level, who in ((6, "oracle"), (6, "zlib"))[:int(os.environ.get("PRODUCERS", "2"))]:
        if who == "oracle":
            comp = [o.compress(p, level, b.ZLIB) for p in plain]
        else:
            comp = [zlib.compress(p, level) for p in plain]
        adl = np.array([zlib.adler32(p) for p in plain], dtype=np.uint32)
        flat, off = b.flatten([comp[k % 64] for k in range(n)])
        d_in = torch.from_numpy(flat).to(dev)
        d_off = torch.from_numpy(off.view(np.int64)).to(dev)
        d_out = torch.empty(n * 65536, dtype=torch.uint8, device=dev)
        d_out_off = torch.arange(n, dtype=torch.int64, device=dev) * 65536
        d_max = torch.full((n,), 65536, dtype=torch.int64, device=dev)
        d_size = torch.zeros(n, dtype=torch.int64, device=dev)
        d_stat = torch.zeros(n, dtype=torch.int32, device=dev)
        d_sum = torch.zeros(n, dtype=torch.int32, device=dev)
        exp = torch.from_numpy(np.tile(adl, n // 64).view(np.int32)).to(dev)
        res = []
        for m in want:
            ctx = ctxs[m]
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            best = None
            for it in range(4):
                d_sum.zero_(); d_stat.fill_(-1)
                ev[0].record(stream)
                ctx.check(ctx._lib.bdf_decompress_batch_device(
                    ctx.handle, b.ZLIB, d_in.data_ptr(), d_off.data_ptr(), n, d_out.data_ptr(), d_out_off.data_ptr(),
                    d_max.data_ptr(), d_size.data_ptr(), d_sum.data_ptr(), d_stat.data_ptr(), C.c_void_p(stream.cuda_stream)))
                ev[1].record(stream)
                torch.cuda.synchronize(dev)
                ms = ev[0].elapsed_time(ev[1])
                best = ms if best is None or ms < best else best
            ok = bool((d_stat == 0).all()) and bool((d_sum == exp).all()) and bool((d_size == 65536).all())
            res.append(f"{m} {n * 65536 / best / 1e6:8.1f} GB/s {'ok' if ok else 'BAD'}")
        ratio = n * 65536 / int(off[-1])
        print(f"{kind:8s} {who:6s} ratio {ratio:6.1f} | " + " | ".join(res), flush=True)
