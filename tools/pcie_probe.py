"""Measures pinned host<->device copy bandwidth (one direction and both at once)."""
import torch, time
n = 8 * 1024 * 1024  # 64 MiB of float64
h_in = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(2)]
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d[0].copy_(h_in[0], non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d[1], non_blocking=True)
def both():
    h2d(); d2h()
for name, fn in (("H2D", h2d), ("D2H", d2h), ("both", both)):
    fn(); dt = t(fn)
    print(f"{name}: {n*8/dt/1e9:.1f} GB/s per direction ({dt*1e3:.2f} ms for 64 MiB)")
for chunk_mb in (1, 2, 4, 8):
    m = chunk_mb * 1024 * 1024 // 8
    def chunks():
        with torch.cuda.stream(s1):
            for k in range(0, n, m): d[0][k:k+m].copy_(h_in[0][k:k+m], non_blocking=True)
    chunks(); dt = t(chunks)
    print(f"H2D in {chunk_mb} MiB pieces: {n*8/dt/1e9:.1f} GB/s")
