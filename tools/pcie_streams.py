import torch, time
n = 3_600_000_000
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
def run(parts):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sz = n // parts
    for i, s in enumerate(streams):
        with torch.cuda.stream(s):
            h[i*sz:(i+1)*sz].copy_(d[i*sz:(i+1)*sz], non_blocking=True)
    torch.cuda.synchronize()
    return n / (time.perf_counter() - t0) / 1e9
for parts in (1, 2, 4, 1, 2):
    print(parts, "stream(s): %.1f GB/s" % run(parts))
