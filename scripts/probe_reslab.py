"""Two apply kernels of the same P4 operator in one process (for one ncu invocation): the broadcast-row
affine kernel and the re-slabbed one."""
import os, sys
import torch
sys.path.insert(0, ".")
from pmg_dolfinx_b200 import api

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctx = api.Context(0)
n = api.boxmesh_fit(int(1e8), P)
m = api.BoxMesh(n)
sp = m.space(P)
d_dm, d_x, d_g = ctx.to_device(sp.dofmap), ctx.to_device(m.xgeom), ctx.to_device(m.geom_dofmap)
d_k = torch.full((m.n_cells,), 2.0, dtype=torch.float64, device=ctx.device)
d_bc = ctx.to_device(sp.bc)
x, y = api.Vector(ctx, sp.n_owned), api.Vector(ctx, sp.n_owned)
x.set(1.0)
for mode in ("0", "2"):
    os.environ["PMGX_AFFINE_RESLAB"] = mode
    op = api.MatFreeLaplacian(ctx, P, d_k, d_dm, d_x, d_g, m.lcells, m.bcells, d_bc, sp.n_owned, 0, None, 2)
    for _ in range(2):
        op(x, y)
    ctx.sync()
    print(op.kernel_name(), api.norm(y))
    op.destroy()
