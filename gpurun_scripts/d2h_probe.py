# D2H bandwidth with 1/2/4 concurrent copies (pinned destination), to see whether the end-to-end arm can gain
import time, torch
n = 4 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for k in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    step = n // k
    for rep in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                h[i * step:(i + 1) * step].copy_(d[i * step:(i + 1) * step], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    print(k, "copies:", round(n / dt / 1e9, 2), "GB/s")
