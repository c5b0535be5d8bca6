"""Pinned-host <-> device copy bandwidth of the box (floor of the e2e copies of mrp_step_host)."""
import torch
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, dst, src in (("H2D", d, h), ("D2H", h, d)):
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (5 * n / e0.elapsed_time(e1) / 1e6))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); e0.record()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
print("duplex each way %.1f GB/s" % (5 * n / e0.elapsed_time(e1) / 1e6))
