import torch, time
n = 8_000_000
h = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(4)]
d = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=20):
    fn(); torch.cuda.synchronize(); t0 = time.time()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.time() - t0) / reps
def d2h():
    with torch.cuda.stream(s1): h[0].copy_(d[0], non_blocking=True)
def h2d():
    with torch.cuda.stream(s2): d[1].copy_(h[1], non_blocking=True)
def both(): d2h(); h2d()
print("D2H GB/s", 8 * n / t(d2h) / 1e9, "H2D GB/s", 8 * n / t(h2d) / 1e9, "both (each) GB/s", 8 * n / t(both) / 1e9)
small = 1_000_000
hs = torch.empty(small, dtype=torch.float64).pin_memory(); ds = torch.empty(small, dtype=torch.float64, device="cuda")
def d2h_small():
    with torch.cuda.stream(s1): hs.copy_(ds, non_blocking=True)
print("8 MB D2H ms", t(d2h_small) * 1e3)
