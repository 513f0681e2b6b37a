import torch, time
for mb in (256, 1024, 4096):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, src, dst in (("d2h", d, h), ("h2d", h, d)):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(name, mb, "MB: %.1f GB/s" % (n / dt / 1e9), flush=True)
# two streams both directions
h1 = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); d2 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1):
    h1.copy_(d1, non_blocking=True)
with torch.cuda.stream(s2):
    h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("2 concurrent d2h streams: %.1f GB/s total" % (2 * (1 << 30) / dt / 1e9))
import os
print("cpus", os.cpu_count())
