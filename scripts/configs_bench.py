"""BASELINE configs 2, 3 and 4 on one GPU (the bench.py line is config 5).

  config 2  examples/vector-update: FP64 axpy/dot/... bandwidth sweep, 2^20 .. 2^30 entries
  config 3  examples/mat_free: P3 matrix-free apply at ~100 M dofs, compared with the reference's own
            stiffness_operator kernel on the same GPU when oracle/_ref is present
  config 4  examples/cg: Jacobi-CG on P6 at ~200 M dofs (20 its, b = 1) + 30 Chebyshev iterations

One JSON line per measurement; CUDA events on the library stream, >= 3 warm-ups, L2 flushed
between repetitions when the working set is smaller than 2x L2.
usage: python scripts/configs_bench.py [2|3|4 ...]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmg_dolfinx_b200 import api  # noqa: E402

PEAK = 6533.2
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
L2_BYTES = 126e6


def timed(ctx, fn, reps, flush=None):
    for _ in range(3):
        fn()
    ctx.sync()
    tot = 0.0
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        for _ in range(reps):
            fn()
        e1.record(ctx.stream)
        ctx.sync()
        return e0.elapsed_time(e1) / reps
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        fn()
        e1.record(ctx.stream)
        ctx.sync()
        tot += e0.elapsed_time(e1)
    return tot / reps


def config2(ctx):
    flush = torch.empty(int(3 * L2_BYTES) // 8, dtype=torch.float64, device=ctx.device)
    for lg in range(20, 31):
        n = 1 << lg
        x, y, r = api.Vector(ctx, n), api.Vector(ctx, n), api.Vector(ctx, n)
        x.data.copy_(torch.arange(n, device=ctx.device, dtype=torch.float64) % 7)   # x_i = i mod 7
        y.set(1.0)
        ops = {
            "axpy": (24, lambda: api.axpy(r, 0.5, x, y)),
            "dot": (16, lambda: api.inner_product(x, y)),      # returns to the host like the reference (:345-350)
            "norm": (8, lambda: api.norm(x)),
            "scale": (16, lambda: api.scale(r, 1.0000001)),
            "copy": (16, lambda: api.copy(r, x)),
            "pointwise_mult": (24, lambda: api.pointwise_mult(r, x, y)),
        }
        small = 3 * 8 * n < 2 * L2_BYTES
        out = {"config": 2, "n": n}
        for name, (bpe, fn) in ops.items():
            ms = timed(ctx, fn, 20 if lg < 28 else 10, flush if small else None)
            out[name + "_gbs"] = round(bpe * n / ms / 1e6, 1)
            out[name + "_frac"] = round(bpe * n / ms / 1e6 / PEAK, 3)
        out["l2_flushed"] = small
        # dot value check: sum_i (i mod 7) = closed form
        want = (n // 7) * 21 + sum(range(n % 7))
        out["dot_exact"] = bool(api.inner_product(x, y) == want)
        print(json.dumps(out), flush=True)
        del x, y, r
        torch.cuda.empty_cache()


PGRID = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
RANK, WORLD = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def make_level(ctx, P, ndofs, want_coords=False):
    """ndofs per rank (weak scaling: the mesh-fit routine runs on ndofs * world, examples/cg/main.cpp:288-310)"""
    n = api.boxmesh_fit(ndofs * WORLD, P)
    m = api.BoxMesh(n, PGRID[WORLD], RANK)
    sp = m.space(P, want_coords=want_coords)
    halo = api.Halo.from_space(ctx, sp) if WORLD > 1 else None
    d = dict(mesh=m, sp=sp, halo=halo, dm=ctx.to_device(sp.dofmap), xg=ctx.to_device(m.xgeom), gd=ctx.to_device(m.geom_dofmap),
             kap=torch.full((m.n_cells,), 2.0, dtype=torch.float64, device=ctx.device), bc=ctx.to_device(sp.bc))
    d["op"] = api.MatFreeLaplacian(ctx, P, d["kap"], d["dm"], d["xg"], d["gd"], m.lcells, m.bcells, d["bc"], sp.n_owned,
                                   sp.n_ghost, halo)
    return d


def rank_max(ctx, ms):
    if WORLD == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def b_apply(P, ncells, ndofs):
    return ncells * ((P + 1) ** 3 * 52 + 8) + ndofs * 17


def config3(ctx):
    P = 3
    d = make_level(ctx, P, 100_000_000)
    sp, m, op = d["sp"], d["mesh"], d["op"]
    x, y = api.Vector(ctx, sp.n_owned), api.Vector(ctx, sp.n_owned)
    x.set(1.0)                                                           # examples/mat_free/main.cpp:250-256
    ms = timed(ctx, lambda: op(x, y), 20)
    B = b_apply(P, m.n_cells, sp.n_owned)
    out = {"config": 3, "P": P, "cells": list(m.n), "ndofs": sp.n_owned, "apply_ms": round(ms, 4),
           "gdofs": round(sp.n_owned / ms / 1e6, 2), "algorithmic_gbs": round(B / ms / 1e6, 1),
           "frac_of_hbm_peak": round(B / ms / 1e6 / PEAK, 3), "ynorm_u1": api.norm(y)}
    xr = torch.from_numpy(np.random.default_rng(42).uniform(-1, 1, sp.n_owned)).to(ctx.device)
    x.data.copy_(xr)
    op(x, y)
    try:
        from oracle import refkernels, operator as oo, gll
        if refkernels.available():
            L = refkernels.load()
            nq = (P + 1) ** 3
            ent = ctx.to_device(np.arange(m.n_cells, dtype=np.int32))
            G = ctx.zeros(m.n_cells * nq * 6)
            dphi, w, D = ctx.to_device(oo.trilinear_dphi(P)), ctx.to_device(oo.weights_3d(P)), ctx.to_device(gll.tables(P)[2])
            ctx.sync()
            assert L.ref_geometry(P, d["xg"].data_ptr(), G.data_ptr(), d["gd"].data_ptr(), dphi.data_ptr(), w.data_ptr(),
                                  ent.data_ptr(), m.n_cells) == 0
            yr = ctx.zeros(sp.n_owned)

            def ref():
                yr.zero_()
                assert L.ref_stiffness(P, xr.data_ptr(), d["kap"].data_ptr(), yr.data_ptr(), G.data_ptr(), d["dm"].data_ptr(),
                                       D.data_ptr(), ent.data_ptr(), m.n_cells, d["bc"].data_ptr(), 0) == 0
            ctx.sync()
            torch.cuda.synchronize()
            with torch.cuda.stream(torch.cuda.default_stream()):
                ref(); ref()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    ref()
                e1.record()
                torch.cuda.synchronize()
            out["reference_kernel_ms"] = round(e0.elapsed_time(e1) / 5, 3)
            out["speedup_vs_reference_kernel"] = round(out["reference_kernel_ms"] / ms, 2)
            out["rel_diff_vs_reference_kernel"] = float((y.data[: sp.n_owned] - yr).norm() / yr.norm())
    except Exception as e:  # the comparator is optional (oracle/_ref is built only where /root/reference exists)
        out["reference_kernel"] = f"unavailable: {e}"
    print(json.dumps(out), flush=True)


def config4(ctx):
    """examples/cg at P6, ~200 M dofs per GPU on 1/2/4/8 GPUs (run under torchrun for more than one): Jacobi-CG with
    b = 1, x0 = 0, 20 its, rtol 1e-6 (examples/cg/main.cpp:238-249), then 30 Chebyshev iterations on the problem of
    :136-158,234-236,268-284: f = 1000 exp(-((x-.5)^2+(y-.5)^2)/0.02), g = 1.3 with lifting, x0 = 1 with the BC set."""
    P = 6
    d = make_level(ctx, P, 200_000_000, want_coords=True)
    sp, m, op = d["sp"], d["mesh"], d["op"]
    x, b = api.Vector(ctx, sp.n_owned, sp.n_ghost, d["halo"]), api.Vector(ctx, sp.n_owned, sp.n_ghost, d["halo"])
    b.set(1.0)
    cg = api.CGSolver(ctx, sp.n_owned, sp.n_ghost)
    cg.set_max_iterations(20)
    cg.set_tolerance(1e-6)
    cg.store_coefficients(True)
    cg.solve(op, x, b)                       # warm-up
    x.set(0.0)
    cg2 = api.CGSolver(ctx, sp.n_owned, sp.n_ghost)
    cg2.set_max_iterations(20)
    cg2.set_tolerance(1e-6)
    cg2.store_coefficients(True)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.stream)
    its = cg2.solve(op, x, b)
    e1.record(ctx.stream)
    ctx.sync()
    ms_cg = rank_max(ctx, e0.elapsed_time(e1))
    eig = cg2.compute_eigenvalues()
    ms_apply = rank_max(ctx, timed(ctx, lambda: op(b, x), 10))
    B = b_apply(P, m.n_owned_cells, sp.n_owned) * WORLD
    ch = api.Chebyshev(ctx, sp.n_owned, sp.n_ghost, (0.1 * eig[-1], 1.1 * eig[-1]))
    ch.set_max_iterations(30)
    X = sp.coords
    f = 1000.0 * np.exp(-((X[:, 0] - 0.5) ** 2 + (X[:, 1] - 0.5) ** 2) / 0.02)
    op.assemble_rhs(ctx.to_device(f), 1.3, b)                       # assemble + lifting + set_bc (g = 1.3)
    def start():
        x.set(1.0)
        x.data[: sp.n_owned + sp.n_ghost][d["bc"].bool()] = 1.3      # set_bc on the initial guess (:280-281)
    start()
    ch.solve(op, x, b)
    start()
    ctx.sync()
    e0.record(ctx.stream)
    ch.solve(op, x, b)
    e1.record(ctx.stream)
    ctx.sync()
    ms_ch = rank_max(ctx, e0.elapsed_time(e1))
    res = ch.residual(op, x, b)
    n = sp.n_global
    if RANK != 0:
        return
    print(json.dumps({"config": 4, "n_gpus": WORLD, "P": P, "cells": list(m.n), "ndofs": n, "cg_iterations": its,
                      "kernel": op.kernel_name(), "chebyshev_residual_after_30": res,
                      "cg_ms_per_iteration": round(ms_cg / its, 3), "cg_gdofs_per_iteration": round(n * its / ms_cg / 1e6, 2),
                      "cg_algorithmic_gbs": round((B + 104 * n) * its / ms_cg / 1e6, 1),
                      "cg_frac_of_hbm_peak": round((B + 104 * n) * its / ms_cg / 1e6 / PEAK / WORLD, 3),
                      "apply_ms": round(ms_apply, 3), "apply_frac_of_hbm_peak": round(B / ms_apply / 1e6 / PEAK / WORLD, 3),
                      "frac_note": "SURVEY 8d byte model (G streamed per quadrature point) per GPU; the affine kernel does not stream G, so > 1 is possible",
                      "lambda_max": float(eig[-1]),
                      "chebyshev_ms_per_iteration": round(ms_ch / 30, 3),
                      "chebyshev_algorithmic_gbs": round((B + 64 * n) * 30 / ms_ch / 1e6, 1),
                      "chebyshev_frac_of_hbm_peak": round((B + 64 * n) * 30 / ms_ch / 1e6 / PEAK / WORLD, 3)}), flush=True)


if __name__ == "__main__":
    which = [int(a) for a in sys.argv[1:]] or [2, 3, 4]
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    nccl_id = None
    if WORLD > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [api.Context.nccl_unique_id() if RANK == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    ctx = api.Context(local, RANK, WORLD, nccl_id)
    for c in which:
        {2: config2, 3: config3, 4: config4}[c](ctx)
        torch.cuda.empty_cache()
    if WORLD > 1:
        ctx.sync()
        dist.barrier()
        dist.destroy_process_group()
