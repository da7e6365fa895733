"""Quick perf probe: matrix-free apply throughput per degree at ~ndofs (single GPU)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from pmg_dolfinx_b200 import api

def b_apply(P, ncells, ndofs):
    return ncells * ((P + 1) ** 3 * 52 + 8) + ndofs * 17

def run(P, ndofs, reps=10, perturb=0.0):
    ctx = api.Context(0)
    n = api.boxmesh_fit(ndofs, P)
    t0 = time.time()
    m = api.BoxMesh(n, perturb=perturb)
    sp = m.space(P)
    t1 = time.time()
    d_dm, d_x, d_g = ctx.to_device(sp.dofmap), ctx.to_device(m.xgeom), ctx.to_device(m.geom_dofmap)
    d_k = torch.full((m.n_cells,), 2.0, dtype=torch.float64, device=ctx.device)
    d_bc = ctx.to_device(sp.bc)
    op = api.MatFreeLaplacian(ctx, P, d_k, d_dm, d_x, d_g, m.lcells, m.bcells, d_bc, sp.n_owned, 0, None, 2)
    x, y = api.Vector(ctx, sp.n_owned), api.Vector(ctx, sp.n_owned)
    x.set(1.0)
    for _ in range(3):
        op(x, y)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.stream)
    for _ in range(reps):
        op(x, y)
    e1.record(ctx.stream)
    ctx.sync()
    ms = e0.elapsed_time(e1) / reps
    B = b_apply(P, m.n_cells, sp.n_owned)
    out = dict(P=P, n=n, ndofs=sp.n_owned, ms=ms, gdofs=sp.n_owned / ms / 1e6, gbs=B / ms / 1e6,
               frac=B / ms / 1e6 / 6533.2, ynorm=api.norm(y), setup_s=round(t1 - t0, 1))
    print(json.dumps(out), flush=True)
    op.destroy()
    del d_dm, d_x, d_g, d_k, d_bc, x, y
    torch.cuda.empty_cache()

if __name__ == "__main__":
    nd = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    degs = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [3]
    perturb = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0   # > 0: non-affine cells, G streamed
    for P in degs:
        run(P, nd, perturb=perturb)
