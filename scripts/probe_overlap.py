"""Does bulk PCIe traffic slow the HBM-bound kernels?  P4 apply loop and vector pass with / without a
concurrent pinned H2D or D2H copy on another stream."""
import sys, json, torch
sys.path.insert(0, ".")
from pmg_dolfinx_b200 import api
import numpy as np
ctx = api.Context(0)
P = 4
n = api.boxmesh_fit(100_000_000, P)
m = api.BoxMesh(n); sp = m.space(P)
dm, xg, gd = ctx.to_device(sp.dofmap), ctx.to_device(m.xgeom), ctx.to_device(m.geom_dofmap)
kap = torch.full((m.n_cells,), 2.0, dtype=torch.float64, device=ctx.device)
bc = ctx.to_device(sp.bc)
op = api.MatFreeLaplacian(ctx, P, kap, dm, xg, gd, m.lcells, m.bcells, bc, sp.n_owned)
x, y, r = api.Vector(ctx, sp.n_owned), api.Vector(ctx, sp.n_owned), api.Vector(ctx, sp.n_owned)
x.set(1.0)
hb = torch.empty(sp.n_owned, dtype=torch.float64).pin_memory()
db = torch.empty(sp.n_owned, dtype=torch.float64, device=ctx.device)
side = torch.cuda.Stream(device=ctx.device)
def run(fn, reps, traffic):
    for _ in range(3): fn()
    ctx.sync(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if traffic:
        with torch.cuda.stream(side):
            for _ in range(4):
                if traffic == "h2d": db.copy_(hb, non_blocking=True)
                else: hb.copy_(db, non_blocking=True)
    e0.record(ctx.stream)
    for _ in range(reps): fn()
    e1.record(ctx.stream)
    ctx.sync(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
for tr in (None, "h2d", "d2h"):
    out[f"apply_ms_{tr}"] = run(lambda: op(x, y), 20, tr)      # 20 x 2 ms < 4 copies x 14 ms
    out[f"axpy_ms_{tr}"] = run(lambda: api.axpy(r, 0.5, x, y), 100, tr)
print(json.dumps(out))
