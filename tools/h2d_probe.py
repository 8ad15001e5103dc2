import torch, time
n = 185*1024*1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(10): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt=(time.perf_counter()-t0)/10
print("H2D pinned %.1f GB/s (%.2f ms per 185 MiB)" % (n/dt/1e9, dt*1e3))
# two streams concurrently
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2=torch.empty_like(d)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
print("H2D 2 streams %.1f GB/s" % (2*n/dt/1e9))
