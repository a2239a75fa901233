import torch, time
dev = torch.device("cuda")
h_in = torch.empty(12_060_000, dtype=torch.uint8).pin_memory()
h_out = torch.empty(8_294_400, dtype=torch.uint8).pin_memory()
d_in = torch.empty_like(h_in, device=dev); d_out = torch.empty_like(h_out, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=50):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D 12.06 MB: {a:.3f} ms ({12.06/a:.1f} GB/s)  D2H 8.29 MB: {b:.3f} ms ({8.29/b:.1f} GB/s)  both concurrently: {c:.3f} ms per pair")
