import torch, time
n = 1 << 29   # 2 GiB of float32 = 512M floats
d = torch.empty(n, dtype=torch.float32, device="cuda")
h = torch.empty(n, dtype=torch.float32).pin_memory()
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(name, n * 4 / dt / 1e9, "GB/s")
# two streams both directions
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.float32).pin_memory(); d2 = torch.empty(n, dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("bidir each", n * 4 / dt / 1e9, "GB/s")
