# write-only HBM bandwidth (the inflate kernel's traffic is 99 % writes): torch fill_ and cudaMemsetAsync on 4 GiB
import torch
n = 4 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d32 = d.view(torch.int32)
for name, fn in (("fill_ (int32)", lambda: d32.fill_(0x01020304)), ("zero_ (memset)", lambda: d.zero_()),
                 ("copy_ 2 GiB -> 2 GiB", lambda: d[: n // 2].copy_(d[n // 2:]))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    moved = n if "copy" not in name else n      # copy: 2 GiB read + 2 GiB written
    print(f"{name}: {moved / best / 1e6:.1f} GB/s ({best:.3f} ms)")
