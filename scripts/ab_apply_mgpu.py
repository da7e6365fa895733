"""A/B of the multi-rank operator apply (run under torchrun): P4 apply time, max over ranks."""
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pmg_dolfinx_b200 import api
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
box = [api.Context.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
ctx = api.Context(local, rank, world, box[0])
out = {}
for P in (4, 2, 1):
    n = api.boxmesh_fit(100_000_000 * world, 4)
    mesh = api.BoxMesh(n, bench.PGRID[world], rank)
    sp = mesh.space(P)
    halo = api.Halo.from_space(ctx, sp)
    dm, bc = ctx.to_device(sp.dofmap), ctx.to_device(sp.bc)
    xg, gd = ctx.to_device(mesh.xgeom), ctx.to_device(mesh.geom_dofmap)
    kap = torch.full((mesh.n_cells,), 2.0, dtype=torch.float64, device=ctx.device)
    op = api.MatFreeLaplacian(ctx, P, kap, dm, xg, gd, mesh.lcells, mesh.bcells, bc, sp.n_owned, sp.n_ghost, halo)
    x, y = api.Vector(ctx, sp.n_owned, sp.n_ghost, halo), api.Vector(ctx, sp.n_owned, sp.n_ghost)
    x.set(1.0)
    for _ in range(5): op(x, y)
    ctx.sync(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.stream)
    for _ in range(30): op(x, y)
    e1.record(ctx.stream)
    ctx.sync(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 30], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[f"P{P}_apply_ms"] = round(float(t.item()), 4)
    out[f"P{P}_bcells"] = len(mesh.bcells)
    del op, x, y
if rank == 0:
    print(json.dumps(dict(world=world, side=not os.environ.get("PMGX_BOUNDARY_ON_COMPUTE"), **out)), flush=True)
dist.barrier(); dist.destroy_process_group()
