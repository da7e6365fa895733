"""Time the components of the bench V-cycle separately (CUDA events, max over ranks)."""
import os, sys, json
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pmg_dolfinx_b200 import api

def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [api.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    ctx = api.Context(local, rank, world, nccl_id)
    ndofs = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    n = api.boxmesh_fit(ndofs * world, 4)
    mesh = api.BoxMesh(n, bench.PGRID[world], rank)
    pmg, ops, b, eigs, keep = bench.build_problem(ctx, api, torch, mesh, world > 1)
    lv, _, _, _, A0, coarse, interps, smoothers = keep
    def barrier():
        ctx.sync()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
    def timeit(name, fn, reps=10):
        for _ in range(2): fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        for _ in range(reps): fn()
        e1.record(ctx.stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=ctx.device)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0: print(f"  {name:34s} {t.item():8.3f} ms", flush=True)
        return t.item()
    vec = lambda d: api.Vector(ctx, d["sp"].n_owned, d["sp"].n_ghost, d["halo"])
    if rank == 0: print(f"world {world} mesh {n} cells/rank {mesh.n_cells} (owned {mesh.n_owned_cells}) lcells {len(mesh.lcells)} bcells {len(mesh.bcells)}")
    tot = 0
    for li, d in enumerate(lv):
        x, y, bb = vec(d), vec(d), vec(d)
        x.set(1.0); bb.set(1.0)
        ta = timeit(f"P{d['P']} apply", lambda: ops[li](x, y))
        ts = timeit(f"P{d['P']} chebyshev(2) solve", lambda: smoothers[li].solve(ops[li], x, bb))
        if d["halo"] is not None:
            timeit(f"P{d['P']} halo fwd (begin+end)", lambda: x.scatter_fwd())
        timeit(f"P{d['P']} axpy", lambda: api.axpy(y, -1.0, y, bb))
        timeit(f"P{d['P']} dot (host result)", lambda: api.inner_product(x, bb), reps=5)
    for i, it in enumerate(interps):
        c, f = vec(lv[i]), vec(lv[i + 1])
        c.set(1.0); f.set(1.0)
        timeit(f"prolong P{lv[i]['P']}->P{lv[i+1]['P']}", lambda: it.interpolate(c, f))
        timeit(f"restrict P{lv[i+1]['P']}->P{lv[i]['P']}", lambda: it.reverse_interpolate(f, c))
    x0, b0 = vec(lv[0]), vec(lv[0])
    b0.set(1.0)
    def cs():
        x0.set(0.0); coarse.solve(x0, b0)
    timeit("coarse solve (b=1: up to 60 its)", cs, reps=3)
    u = vec(lv[-1])
    timeit("V-cycle", lambda: pmg.apply(b, u), reps=5)
    # python replica of the cycle with an event after every phase
    nl = len(lv)
    U = [vec(d) for d in lv]; R = [vec(d) for d in lv]; B = [vec(d) for d in lv]; DU = [vec(d) for d in lv]
    api.copy(B[-1], b)
    marks = []
    def mark(name):
        e = torch.cuda.Event(enable_timing=True); e.record(ctx.stream); marks.append((name, e))
    for rep in range(3):
        marks.clear()
        barrier()
        mark("start")
        for i in range(nl - 1): U[i].set(0.0)
        for i in range(nl - 1, 0, -1):
            smoothers[i].solve(ops[i], U[i], B[i]); mark(f"pre-smooth L{i}")
            ops[i](U[i], R[i]); api.axpy(R[i], -1.0, R[i], B[i]); mark(f"residual L{i}")
            interps[i - 1].reverse_interpolate(R[i], B[i - 1]); mark(f"restrict L{i}")
        coarse.solve(U[0], B[0]); mark("coarse")
        for i in range(nl - 1):
            interps[i].interpolate(U[i], DU[i + 1]); mark(f"prolong L{i+1}")
            api.axpy(U[i + 1], 1.0, U[i + 1], DU[i + 1])
            smoothers[i + 1].solve(ops[i + 1], U[i + 1], B[i + 1]); mark(f"post-smooth L{i+1}")
        barrier()
    if rank == 0:
        for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
            print(f"    phase {n1:20s} {e0.elapsed_time(e1):8.3f} ms")
        print(f"    total {marks[0][1].elapsed_time(marks[-1][1]):8.3f} ms")
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
main()
