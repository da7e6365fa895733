"""PCIe copy probe: H2D / D2H alone and concurrently (pinned host memory), to size the e2e pipeline."""
import torch, time, json
n = 99_679_725
dev = torch.device("cuda", 0)
hb = torch.empty(n, dtype=torch.float64).pin_memory()
hu = torch.empty(n, dtype=torch.float64).pin_memory()
db = torch.empty(n, dtype=torch.float64, device=dev)
du = torch.empty(n, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
def h2d():
    with torch.cuda.stream(s1): db.copy_(hb, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): hu.copy_(du, non_blocking=True)
def both():
    h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
gb = n * 8 / 1e9
print(json.dumps({"h2d_ms": a, "d2h_ms": b, "both_ms": c, "h2d_gbs": gb / a * 1e3, "d2h_gbs": gb / b * 1e3,
                  "both_gbs_each": gb / c * 1e3}))
