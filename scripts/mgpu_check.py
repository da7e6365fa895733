"""Multi-GPU parity check (run under torchrun, one rank per GPU; also called in-process by bench.py
before its timed region).

Partitions a small box mesh over the ranks with the product's host partitioner, runs the
operator apply, Jacobi-CG, Chebyshev and the P4->P2->P1 V-cycle with the peer-memory / NCCL halo
exchange and all-reduced dots, gathers the owned values on rank 0 and compares them with the
single-domain oracle in the canonical global numbering (result must be partition
independent to 1e-12; histories 1e-10, identical CG iteration counts).  Exit code 0 = pass.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PGRID = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    backend = os.environ.get("PMGX_CHECK_BACKEND", "nccl")
    torch.cuda.set_device(local)
    dist.init_process_group(backend, device_id=torch.device("cuda", local) if backend == "nccl" else None)
    from pmg_dolfinx_b200 import api
    box = [api.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx = api.Context(local, rank, world, box[0])
    n = tuple(int(v) for v in os.environ.get("PMGX_CHECK_MESH", "9,8,7").split(","))
    perturb = float(os.environ.get("PMGX_CHECK_PERTURB", "0.15"))
    res = run_check(api, ctx, rank, world, n, perturb, general=os.environ.get("PMGX_CHECK_GENERAL", "0") == "1")
    ctx.sync()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print(f"[mgpu x{world}] {'PASS' if res['ok'] else 'FAIL'} (max rel err / tolerance = {res['max_err_over_tol']:.3e})",
              flush=True)
    sys.exit(0 if res["ok"] else 1)


def run_check(api, ctx, rank, world, n=(9, 8, 7), perturb=0.15, log=print, general=False):
    """Collective over all ranks of ctx (torch.distributed must be initialised when world > 1).
    Returns {"ok", "max_rel", "max_err_over_tol", "transport", "checks"} on every rank.
    general: the mesh is handed to the library as a GENERAL hex mesh -- vertices renumbered at random, cells
    shuffled and rotated, partition along skew planes -- through the ghost-layer builder
    (api.GhostLayerMesh, src/mesh.hpp:16-143) instead of the structured box partitioner."""
    degrees = (1, 2, 4)
    gen_oracle_mesh = None
    if general:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        from test_ghostmesh import scrambled_box
        gen_oracle_mesh, gcells, gowner, gcoords = scrambled_box(n, perturb, 5, world)
        mesh = api.GhostLayerMesh(gcells, gowner, gcoords, rank, world)
        full = mesh
    else:
        mesh = api.BoxMesh(n, PGRID[world], rank, perturb=perturb)
        # the same perturbed geometry for the oracle: take it from a single-domain product mesh
        full = api.BoxMesh(n, (1, 1, 1), 0, perturb=perturb)

    def gather_owned(vec, sp):
        """rank 0 gets the global vector in canonical numbering."""
        vals = vec.data[: sp.n_owned].cpu().numpy() if hasattr(vec, "data") else vec
        if world > 1:
            objs = [None] * world if rank == 0 else None
            dist.gather_object((sp.l2g[: sp.n_owned], vals), objs, dst=0)
        else:
            objs = [(sp.l2g[: sp.n_owned], vals)]
        if rank != 0:
            return None
        out = np.full(sp.n_global, np.nan)
        for l2g, v in objs:
            out[l2g] = v
        assert not np.isnan(out).any()
        return out

    def scatter_global(gvec, sp, halo):
        v = api.Vector(ctx, sp.n_owned, sp.n_ghost, halo)
        loc = np.zeros(sp.n_owned + sp.n_ghost)
        loc[: sp.n_owned] = gvec[sp.l2g[: sp.n_owned]]
        v.data.copy_(torch.from_numpy(loc))
        return v

    xgeom, gdm = ctx.to_device(mesh.xgeom), ctx.to_device(mesh.geom_dofmap)
    kappa = torch.full((mesh.n_cells,), 2.0, dtype=torch.float64, device=ctx.device)
    lv = []
    for P in degrees:
        sp = mesh.space(P, want_coords=True)
        if general:
            # canonical numbering = the structured oracle's, matched through the dof coordinates
            from scipy.spatial import cKDTree
            from oracle import mesh as om_
            dd, idx = cKDTree(om_.dof_coords(gen_oracle_mesh, P)).query(sp.coords)
            assert dd.max() < 1e-12
            sp.l2g, sp.n_global = idx.astype(np.int64), om_.num_dofs(gen_oracle_mesh, P)
        halo = api.Halo.from_space(ctx, sp) if world > 1 else None
        dm, bc = ctx.to_device(sp.dofmap), ctx.to_device(sp.bc)
        op = api.MatFreeLaplacian(ctx, P, kappa, dm, xgeom, gdm, mesh.lcells, mesh.bcells, bc, sp.n_owned, sp.n_ghost, halo)
        # exterior-facet marker built on the device (owned + ghost entries) against the host builder's
        bc_dev = api.exterior_bc_marker(ctx, P, gdm, dm, sp.n_owned, sp.n_ghost, halo)
        bad = torch.tensor([int((bc_dev != bc).sum().item())], device=ctx.device)
        if world > 1:
            dist.all_reduce(bad)
        lv.append(dict(P=P, sp=sp, halo=halo, dm=dm, bc=bc, op=op, bc_mismatch=int(bad.item())))

    transport = "none"
    if world > 1:
        paths = {bool(api.lib.pmgx_halo_uses_p2p(d["halo"].h)) for d in lv} | {bool(api.lib.pmgx_ctx_uses_p2p(ctx.h))}
        transport = "nvlink-p2p" if paths == {True} else "nccl" if paths == {False} else "mixed"
    if rank == 0:
        log("halo path: " + transport)

    # ---------------- oracle on rank 0 (single domain, same geometry)
    ok = True
    if rank == 0:
        from oracle import mesh as om, operator as oo, solvers as osol
        omesh = gen_oracle_mesh if general else om.BoxMesh(n, full.xgeom.copy(), full.geom_dofmap.copy())
        O = []
        for P in degrees:
            dm, bc, nd = om.dofmap(omesh, P), om.bc_marker(omesh, P), om.num_dofs(omesh, P)
            G, _ = oo.geometry_factors(omesh.verts, omesh.geom_dofmap, P)
            kap = np.full(omesh.ncells, 2.0)
            A = (lambda P, dm, G, kap, bc: lambda v: oo.apply(P, dm, G, kap, bc, v))(P, dm, G, kap, bc)
            O.append(dict(P=P, dm=dm, bc=bc, nd=nd, A=A, dinv=1.0 / oo.diagonal(P, dm, G, kap, bc, nd)))

    checks, worst = {}, [0.0, 0.0]

    def check(name, got, ref, tol):
        nonlocal ok
        err = float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300))
        good = err <= tol
        ok = ok and good
        checks[name] = err
        worst[0], worst[1] = max(worst[0], err), max(worst[1], err / tol)
        log(f"[mgpu x{world}] {name:38s} rel err {err:.3e}  {'ok' if good else 'FAIL'}")

    rng = np.random.default_rng(42)
    smoothers, eigs = [], []
    for li, L in enumerate(lv):
        sp, P = L["sp"], L["P"]
        xg = rng.uniform(-1, 1, sp.n_global)
        x = scatter_global(xg, sp, L["halo"])
        y = api.Vector(ctx, sp.n_owned, sp.n_ghost)
        L["op"](x, y)
        yg = gather_owned(y, sp)
        dv = api.Vector(ctx, sp.n_owned, sp.n_ghost)
        L["op"].get_diag_inverse(dv)
        dg = gather_owned(dv, sp)
        # CG with b = 1 (examples/pmg/main.cpp:306-330)
        cg = api.CGSolver(ctx, sp.n_owned, sp.n_ghost)
        cg.set_max_iterations(20)
        cg.set_tolerance(1e-6)
        cg.store_coefficients(True)
        xs, b1 = api.Vector(ctx, sp.n_owned, sp.n_ghost, L["halo"]), api.Vector(ctx, sp.n_owned, sp.n_ghost)
        b1.set(1.0)
        k = cg.solve(L["op"], xs, b1)
        eig = cg.compute_eigenvalues()
        r0, hist = cg.history()
        xsg = gather_owned(xs, sp)
        s = api.Chebyshev(ctx, sp.n_owned, sp.n_ghost, (0.1 * eig[-1], 1.1 * eig[-1]))
        s.set_max_iterations(2)
        smoothers.append(s)
        eigs.append(eig[-1])
        if rank == 0:
            o = O[li]
            good = L["bc_mismatch"] == 0
            ok = ok and good
            checks[f"P{P} device BC marker mismatches"] = L["bc_mismatch"]
            log(f"[mgpu x{world}] P{P} device exterior-facet BC marker vs host: {L['bc_mismatch']} mismatches  {'ok' if good else 'FAIL'}")
            check(f"P{P} apply", yg, o["A"](xg), 1e-12)
            check(f"P{P} diag inverse", dg, o["dinv"], 1e-13)
            xo, ko, al, be, ho, r0o = osol.cg(o["A"], o["dinv"], np.zeros(o["nd"]), np.ones(o["nd"]), 20, 1e-6)
            ok = ok and (k == ko)
            checks[f"P{P} CG iterations (ours, oracle)"] = [int(k), int(ko)]
            log(f"[mgpu x{world}] P{P} CG iterations {k} vs oracle {ko}")
            check(f"P{P} CG residual history", hist, ho, 1e-10)
            check(f"P{P} CG solution", xsg, xo, 1e-9)
            check(f"P{P} lambda_max", np.array([eig[-1]]), np.array([osol.lanczos_eigenvalues(al, be)[-1]]), 1e-9)
            o["lmax"] = 1.1 * osol.lanczos_eigenvalues(al, be)[-1]

    # ---------------- V-cycle
    interps = []
    for a, b in zip(lv[:-1], lv[1:]):
        interps.append(api.Interpolator(ctx, a["P"], b["P"], a["dm"], b["dm"], a["sp"].n_owned + a["sp"].n_ghost,
                                        b["sp"].n_owned + b["sp"].n_ghost, mesh.lcells, mesh.bcells, a["halo"], b["halo"]))
    A0 = lv[0]["op"].to_csr()
    # the coarse solver of the product: smoothed-aggregation PCG (distributed hierarchy; PMGX_CHECK_AMG=0:
    # Jacobi-PCG).  Both reach 1e-10, so the cycle's iterates equal the oracle's (Jacobi-CG to 1e-10)
    use_amg = os.environ.get("PMGX_CHECK_AMG", "1") != "0"
    # (1e-12 in the M^-1 norm: the stage residuals below are compared to 1e-9 in the 2-norm)
    coarse = api.CoarseSolverType(ctx, A0, 200 if not use_amg else 60, 1e-12, amg=use_amg, min_coarse=40)
    if rank == 0:
        log(f"[mgpu x{world}] coarse solver: {'SA-AMG PCG' if use_amg else 'Jacobi-PCG'}, levels on rank 0 "
            f"(rows, nnz, ghosts, dense): {coarse.levels()}")
    pmg = api.MultigridPreconditioner(ctx, [L["bc"] for L in lv], flags=2)
    pmg.set_solvers(smoothers)
    pmg.set_operators([L["op"] for L in lv])
    pmg.set_interpolators(interps)
    pmg.set_coarse_solver(coarse)
    top = lv[-1]
    sp = top["sp"]
    X = sp.coords
    f = 2.0 * np.pi ** 2 * 3 * np.sin(np.pi * X[:, 0]) * np.sin(np.pi * X[:, 1]) * np.sin(np.pi * X[:, 2]) + 1.0 + X[:, 0]
    bvec = api.Vector(ctx, sp.n_owned, sp.n_ghost, top["halo"])
    top["op"].assemble_rhs(ctx.to_device(f), 0.0, bvec)
    bg = gather_owned(bvec, sp)
    # inhomogeneous Dirichlet data: assemble + lifting + set_bc (examples/pmg/main.cpp:289-295)
    blift = api.Vector(ctx, sp.n_owned, sp.n_ghost, top["halo"])
    top["op"].assemble_rhs(ctx.to_device(f), 1.3, blift)
    blg = gather_owned(blift, sp)
    u = api.Vector(ctx, sp.n_owned, sp.n_ghost)
    if rank == 0:
        from oracle import mesh as om, operator as oo, solvers as osol
        olev = [osol.Level(o["A"], o["dinv"], o["bc"].astype(float), o["lmax"], 2) for o in O]
        pro, res = [], []
        for a, b in zip(O[:-1], O[1:]):
            pro.append((lambda a, b: lambda xc: oo.prolong(a["P"], b["P"], a["dm"], b["dm"], xc, b["nd"]))(a, b))
            res.append((lambda a, b: lambda xf: oo.restrict(a["P"], b["P"], a["dm"], b["dm"], xf, a["nd"]))(a, b))
        G0, _ = oo.geometry_factors(omesh.verts, omesh.geom_dofmap, degrees[0])
        A0o = oo.assemble_csr(degrees[0], O[0]["dm"], G0, np.full(omesh.ncells, 2.0), O[0]["bc"], O[0]["nd"])
        d0 = 1.0 / A0o.diagonal()
        cso = lambda u0, b0: osol.cg(lambda v: A0o @ v, d0, u0, b0, 400, 1e-12)[0]
        Xo = om.dof_coords(omesh, degrees[-1])
        fo = 2.0 * np.pi ** 2 * 3 * np.sin(np.pi * Xo[:, 0]) * np.sin(np.pi * Xo[:, 1]) * np.sin(np.pi * Xo[:, 2]) + 1.0 + Xo[:, 0]
        bo = oo.rhs_collocated(omesh, degrees[-1], lambda _: fo, O[-1]["bc"])
        check("rhs assembly", bg, bo, 1e-13)
        check("rhs assembly + lifting (g = 1.3)", blg,
              oo.rhs_collocated(omesh, degrees[-1], lambda _: fo, O[-1]["bc"], g=1.3, kappa=2.0), 1e-12)
        uo = np.zeros(O[-1]["nd"])
    for it in range(3):
        rn = pmg.apply(bvec, u, verbose=True)
        conv, crel = coarse.last_status()
        if rank == 0:
            good = conv and coarse.last_iterations() <= (16 if use_amg else 200)
            ok = ok and good
            checks[f"V-cycle {it} coarse solve (iterations, converged, rel)"] = [coarse.last_iterations(), bool(conv), crel]
            log(f"[mgpu x{world}] V-cycle {it} coarse solve: {coarse.last_iterations()} iterations, "
                f"converged {conv}, rel. residual {crel:.2e}  {'ok' if good else 'FAIL'}")
        hg = pmg.diagnostics()
        ug = gather_owned(u, sp)
        if rank == 0:
            ho = []
            uo = osol.vcycle(olev, pro, res, bo, uo, coarse_solve=cso, history=ho)
            hov = np.array([h[2] for h in ho])
            good = len(hg) == len(hov) and np.all(np.abs(hg - hov) <= 1e-9 * hov[0])
            ok = ok and good
            log(f"[mgpu x{world}] V-cycle {it} stage residuals {'ok' if good else 'FAIL'} (final {rn:.6e} vs {hov[-1]:.6e})")
            check(f"V-cycle {it} solution", ug, uo, 1e-9)
    res = [dict(ok=bool(ok), max_rel=worst[0], max_err_over_tol=worst[1], transport=transport, checks=checks,
                mesh=list(n), perturb=perturb) if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(res, src=0)
    ctx.sync()
    if full is not mesh:
        full.close()
    mesh.close()
    return res[0]


if __name__ == "__main__":
    main()
