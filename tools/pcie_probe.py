import time, torch
n = 64 * 14_400_000
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True); h_in.fill_(1.0)
h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
d_a = torch.empty(n, dtype=torch.float32, device="cuda"); d_b = torch.ones(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
gb = n * 4 / 1e9
a = t(lambda: d_a.copy_(h_in, non_blocking=True)); print(f"H2D {gb / a:.1f} GB/s ({a * 1e3:.1f} ms)")
b = t(lambda: h_out.copy_(d_b, non_blocking=True)); print(f"D2H {gb / b:.1f} GB/s ({b * 1e3:.1f} ms)")
def both():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
c = t(both); print(f"both directions concurrently: {2 * gb / c:.1f} GB/s total ({c * 1e3:.1f} ms)")
def chunks():
    k = 64; m = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d_a[i * m:(i + 1) * m].copy_(h_in[i * m:(i + 1) * m], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i * m:(i + 1) * m].copy_(d_b[i * m:(i + 1) * m], non_blocking=True)
d = t(chunks); print(f"both, 64 chunks each: {2 * gb / d:.1f} GB/s total ({d * 1e3:.1f} ms)")
