"""Host<->device copy bandwidth of the box (development tool): the floor of the e2e number."""
import time, torch
n = 16777216
h_in = torch.empty((n, 8), dtype=torch.float32, pin_memory=True); h_out = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
d_in = torch.empty((n, 8), dtype=torch.float32, device="cuda"); d_out = torch.empty((n, 4), dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print("H2D 537MB %.2f ms (%.1f GB/s); D2H 268MB %.2f ms (%.1f GB/s); concurrent %.2f ms => e2e floor %.0f Mrays/s" % (a, 0.537 / a * 1e3, b, 0.268 / b * 1e3, c, n / c * 1e-3))
