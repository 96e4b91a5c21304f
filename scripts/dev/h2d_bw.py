import torch, time
n = 800_000_000
h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.fill_(1.0)
d = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(4)]
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); d[0].copy_(h, non_blocking=True); torch.cuda.synchronize()
    print("1 stream H2D GB/s", n * 4 / (time.perf_counter() - t) / 1e9)
ss = [torch.cuda.Stream() for _ in range(4)]
torch.cuda.synchronize(); t = time.perf_counter()
for s, x in zip(ss, d):
    with torch.cuda.stream(s):
        x.copy_(h, non_blocking=True)
torch.cuda.synchronize(); print("4 streams H2D GB/s", 4 * n * 4 / (time.perf_counter() - t) / 1e9)
hp = torch.empty(n, dtype=torch.float32); hp.fill_(1.0)
torch.cuda.synchronize(); t = time.perf_counter(); d[0].copy_(hp); torch.cuda.synchronize()
print("pageable H2D GB/s", n * 4 / (time.perf_counter() - t) / 1e9)
ho = torch.empty(n, dtype=torch.float32, pin_memory=True)
torch.cuda.synchronize(); t = time.perf_counter(); ho.copy_(d[0], non_blocking=True); torch.cuda.synchronize()
print("D2H pinned GB/s", n * 4 / (time.perf_counter() - t) / 1e9)
