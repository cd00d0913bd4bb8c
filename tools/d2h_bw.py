import time, torch
n = 1 << 28  # 2 GiB of doubles
d = torch.empty(n, dtype=torch.float64, device="cuda")
h = torch.empty(n, dtype=torch.float64, pin_memory=True)
for chunks in (1, 6):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        step = n // chunks
        for c in range(chunks):
            h[c*step:(c+1)*step].copy_(d[c*step:(c+1)*step], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("chunks", chunks, "GB/s", n * 8 / dt / 1e9)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    half = n // 2
    with torch.cuda.stream(s1): h[:half].copy_(d[:half], non_blocking=True)
    with torch.cuda.stream(s2): h[half:].copy_(d[half:], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("two streams GB/s", n * 8 / dt / 1e9)
