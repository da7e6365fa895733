"""Same-GPU comparator: the reference's own stiffness_operator kernel (oracle/_ref) vs ours."""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from pmg_dolfinx_b200 import api
from oracle import refkernels, operator as oo, gll

def run(P, ndofs, reps=5):
    ctx = api.Context(0)
    L = refkernels.load()
    n = api.boxmesh_fit(ndofs, P)
    m = api.BoxMesh(n)
    sp = m.space(P)
    nq = (P + 1) ** 3
    d_dm, d_x, d_g = ctx.to_device(sp.dofmap), ctx.to_device(m.xgeom), ctx.to_device(m.geom_dofmap)
    d_k = torch.full((m.n_cells,), 2.0, dtype=torch.float64, device=ctx.device)
    d_bc = ctx.to_device(sp.bc)
    op = api.MatFreeLaplacian(ctx, P, d_k, d_dm, d_x, d_g, m.lcells, m.bcells, d_bc, sp.n_owned, 0, None, 2)
    ent = ctx.to_device(np.arange(m.n_cells, dtype=np.int32))
    G = ctx.zeros(m.n_cells * nq * 6)
    d_dphi, d_w, d_D = ctx.to_device(oo.trilinear_dphi(P)), ctx.to_device(oo.weights_3d(P)), ctx.to_device(gll.tables(P)[2])
    ctx.sync()
    assert L.ref_geometry(P, d_x.data_ptr(), G.data_ptr(), d_g.data_ptr(), d_dphi.data_ptr(), d_w.data_ptr(), ent.data_ptr(), m.n_cells) == 0
    x = torch.rand(sp.n_owned, dtype=torch.float64, device=ctx.device)
    xv, yv = api.Vector(ctx, sp.n_owned), api.Vector(ctx, sp.n_owned)
    xv.data.copy_(x)
    yr = ctx.zeros(sp.n_owned)
    ctx.sync(); torch.cuda.synchronize()
    def ref():
        yr.zero_()
        assert L.ref_stiffness(P, x.data_ptr(), d_k.data_ptr(), yr.data_ptr(), G.data_ptr(), d_dm.data_ptr(), d_D.data_ptr(),
                               ent.data_ptr(), m.n_cells, d_bc.data_ptr(), 0) == 0
    # the reference launches on the default stream: time with default-stream events
    with torch.cuda.stream(torch.cuda.default_stream()):
        ref(); ref()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): ref()
        e1.record()
        torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / reps
    for _ in range(3): op(xv, yv)
    ctx.sync()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(ctx.stream)
    for _ in range(reps): op(xv, yv)
    a1.record(ctx.stream)
    ctx.sync()
    ms = a0.elapsed_time(a1) / reps
    err = float((yv.data - yr).norm() / yr.norm())
    print(json.dumps(dict(P=P, ndofs=sp.n_owned, ref_ms=ms_ref, ours_ms=ms, speedup=ms_ref / ms, ref_gdofs=sp.n_owned / ms_ref / 1e6,
                          ours_gdofs=sp.n_owned / ms / 1e6, rel_diff=err)), flush=True)
    op.destroy()

if __name__ == "__main__":
    nd = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    for P in [int(a) for a in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["3"])]:
        run(P, nd)
        torch.cuda.empty_cache()
