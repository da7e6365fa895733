"""CPU, world_size 2 and 4 over gloo: the host side of the multi-GPU path (box partition, ghost
layer, lcells/bcells split, forward-scatter lists from the product's C++ partitioner) driven
through real send/recv, with the oracle standing in for the device kernels.  The distributed
result must equal the single-domain oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PGRID = {2: (2, 1, 1), 4: (2, 2, 1)}


def _halo_fwd(sp, v, world):
    """owner -> ghost update with the product's Scatterer-style lists (src/vector.hpp:186-238)."""
    reqs, bufs = [], []
    for i, q in enumerate(sp.recv_ranks):
        buf = torch.zeros(int(sp.recv_offsets[i + 1] - sp.recv_offsets[i]), dtype=torch.float64)
        bufs.append(buf)
        reqs.append(dist.irecv(buf, src=int(q)))
    for i, q in enumerate(sp.send_ranks):
        idx = sp.send_idx[sp.send_offsets[i]:sp.send_offsets[i + 1]]
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(v[idx])), dst=int(q)))
    for r in reqs:
        r.wait()
    for i, buf in enumerate(bufs):
        slots = sp.recv_idx[sp.recv_offsets[i]:sp.recv_offsets[i + 1]]
        v[sp.n_owned + slots] = buf.numpy()


def _worker(rank, world, port, n, P, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pmg_dolfinx_b200 import api
    from oracle import operator as oo
    mesh = api.BoxMesh(n, PGRID[world], rank, perturb=0.1)
    sp = mesh.space(P)
    rng = np.random.default_rng(7)
    xg = rng.uniform(-1, 1, sp.n_global)                 # same on every rank
    x = np.zeros(sp.n_owned + sp.n_ghost)
    x[: sp.n_owned] = xg[sp.l2g[: sp.n_owned]]
    G, _ = oo.geometry_factors(mesh.xgeom, mesh.geom_dofmap, P)
    kap = np.full(mesh.n_cells, 2.0)
    y = np.zeros_like(x)
    # begin -> interior cells -> end -> boundary cells (src/laplacian.hpp:378-455)
    oo.apply_cells(P, sp.dofmap, G, kap, sp.bc, x, y, mesh.lcells)   # touches owned dofs only
    assert (sp.dofmap[mesh.lcells] < sp.n_owned).all()
    _halo_fwd(sp, x, world)
    assert np.array_equal(x, xg[sp.l2g])
    oo.apply_cells(P, sp.dofmap, G, kap, sp.bc, x, y, mesh.bcells)
    dots = torch.tensor([float(np.dot(x[: sp.n_owned], y[: sp.n_owned]))], dtype=torch.float64)
    dist.all_reduce(dots)                                 # inner_product + allreduce (vector.hpp:333-352)
    objs = [None] * world if rank == 0 else None
    dist.gather_object((sp.l2g[: sp.n_owned], y[: sp.n_owned]), objs, dst=0)
    if rank == 0:
        yg = np.full(sp.n_global, np.nan)
        for l2g, v in objs:
            yg[l2g] = v
        np.save(out, np.concatenate([yg, dots.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_distributed_apply_matches_single_domain(world, tmp_path):
    from pmg_dolfinx_b200 import api
    from oracle import mesh as om, operator as oo
    n, P = (6, 5, 4), 2
    out = str(tmp_path / "y.npy")
    mp.spawn(_worker, args=(world, 29700 + world, n, P, out), nprocs=world, join=True)
    res = np.load(out)
    yg, dot = res[:-1], res[-1]
    full = api.BoxMesh(n, (1, 1, 1), 0, perturb=0.1)
    omesh = om.BoxMesh(n, full.xgeom.copy(), full.geom_dofmap.copy())
    dm, bc, nd = om.dofmap(omesh, P), om.bc_marker(omesh, P), om.num_dofs(omesh, P)
    G, _ = oo.geometry_factors(omesh.verts, omesh.geom_dofmap, P)
    xg = np.random.default_rng(7).uniform(-1, 1, nd)
    yo = oo.apply(P, dm, G, np.full(omesh.ncells, 2.0), bc, xg)
    assert not np.isnan(yg).any()
    assert np.linalg.norm(yg - yo) <= 1e-12 * np.linalg.norm(yo)
    assert abs(dot - np.dot(xg, yo)) <= 1e-12 * abs(np.dot(xg, yo))
