# timing only (no parity checks): config-2 shape through the device entry point
import sys, zlib
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np, torch, ctypes as C
import corpus
import libdeflate_rsx_b200 as b
n = 65536
plain = [corpus.corpus_a_stream(k) for k in range(16)]
comp = []
for p in plain:
    c = zlib.compressobj(6, zlib.DEFLATED, 15); comp.append(c.compress(p) + c.flush())
flat = np.frombuffer(b"".join(comp) * (n // 16), dtype=np.uint8)
lens = np.array([len(comp[k % 16]) for k in range(n)], dtype=np.uint64)
off = np.zeros(n + 1, dtype=np.uint64); off[1:] = np.cumsum(lens)
dev = torch.device("cuda", 0); ctx = b.default_context(); lib = ctx._lib
d_in = torch.from_numpy(flat.copy()).to(dev); d_off = torch.from_numpy(off.view(np.int64)).to(dev)
d_out = torch.empty(n * 65536, dtype=torch.uint8, device=dev)
d_ooff = torch.arange(n, dtype=torch.int64, device=dev) * 65536
d_max = torch.full((n,), 65536, dtype=torch.int64, device=dev)
d_size = torch.zeros(n, dtype=torch.int64, device=dev); d_st = torch.zeros(n, dtype=torch.int32, device=dev); d_sum = torch.zeros(n, dtype=torch.int32, device=dev)
s = torch.cuda.Stream(dev); torch.cuda.set_stream(s)
def step():
    ctx.check(lib.bdf_decompress_batch_device(ctx.handle, 1, d_in.data_ptr(), d_off.data_ptr(), n, d_out.data_ptr(), d_ooff.data_ptr(), d_max.data_ptr(), d_size.data_ptr(), d_sum.data_ptr(), d_st.data_ptr(), C.c_void_p(s.cuda_stream)))
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(10): step()
e1.record(s); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{sys.argv[1] if len(sys.argv) > 1 else ''} ms {ms:.3f}  GB/s {n * 65536 / ms / 1e6:.1f}  status_ok {int((d_st == 0).sum())}")
