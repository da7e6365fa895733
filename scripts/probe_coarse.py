"""Coarse-level probe: P1 CSR SpMV and the coarse PCG alone (per-kernel view with ncu)."""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from pmg_dolfinx_b200 import api

def main():
    ctx = api.Context(0)
    n = (115, 116, 116) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split(","))
    m = api.BoxMesh(n)
    sp = m.space(1)
    d_dm, d_x, d_g = ctx.to_device(sp.dofmap), ctx.to_device(m.xgeom), ctx.to_device(m.geom_dofmap)
    d_k = torch.full((m.n_cells,), 2.0, dtype=torch.float64, device=ctx.device)
    d_bc = ctx.to_device(sp.bc)
    op = api.MatFreeLaplacian(ctx, 1, d_k, d_dm, d_x, d_g, m.lcells, m.bcells, d_bc, sp.n_owned, 0, None)
    A = op.to_csr()
    import os
    amg = os.environ.get("PROBE_AMG", "1") != "0"
    nu = int(os.environ.get("PROBE_NU", "2"))
    rtol = float(os.environ.get("PROBE_RTOL", "1e-5"))
    cs = api.CoarseSolverType(ctx, A, 60, rtol, amg=amg, nu=nu)
    x, b, y = api.Vector(ctx, sp.n_owned), api.Vector(ctx, sp.n_owned), api.Vector(ctx, sp.n_owned)
    b.set(1.0)
    b.data[: sp.n_owned] *= (1 - d_bc[: sp.n_owned].double())
    def timeit(fn, reps=20):
        for _ in range(3): fn()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        for _ in range(reps): fn()
        e1.record(ctx.stream)
        ctx.sync()
        return e0.elapsed_time(e1) / reps
    t_spmv = timeit(lambda: A(b, y))
    def solve():
        x.set(0.0); cs.solve(x, b)
    t_cs = timeit(solve, reps=5)
    nnz = A.nnz()
    print(json.dumps(dict(n_rows=sp.n_owned, nnz=nnz, spmv_ms=t_spmv, spmv_gbs=(nnz * 12 + sp.n_owned * 24) / t_spmv / 1e6,
                          coarse_ms=t_cs, amg=amg, nu=nu, rtol=rtol, iterations=cs.last_iterations(), status=cs.last_status(),
                          levels=cs.levels())))
main()
