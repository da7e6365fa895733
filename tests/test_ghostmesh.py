"""CPU: the general-mesh ghost-layer builder (pmgx_ghostmesh_*, csrc/ghostmesh.cpp; SURVEY 8f-1, the
stand-in for create_mesh + ghost_layer_mesh + compute_boundary_cells, src/mesh.hpp:16-143).

The library is handed a mesh that is NOT a lexicographic box: the oracle's (perturbed) box with its
vertices renumbered by a random permutation, its cells shuffled, every cell's local frame rotated by a
random proper rotation of the cube (so neighbouring cells see shared edges and faces with different
orientations) and a partition along skew planes into 3 or 5 parts.  Each rank's arrays are then driven
through the oracle's element kernel with an emulated halo update; the assembled result must equal the
single-domain structured oracle on the same physical dofs (matched by coordinates) to 1e-12."""
import itertools

import numpy as np
import pytest
from scipy.spatial import cKDTree

from oracle import mesh as om, operator as oo


def _rotations():
    """The 24 orientation-preserving symmetries of the unit cube as (axis permutation, flips)."""
    out = []
    for perm in itertools.permutations(range(3)):
        for flips in itertools.product((0, 1), repeat=3):
            M = np.zeros((3, 3))
            for d in range(3):
                M[d, perm[d]] = -1.0 if flips[d] else 1.0
            if np.linalg.det(M) > 0:
                out.append((perm, flips))
    assert len(out) == 24
    return out


def scrambled_box(n, perturb, seed, nranks):
    m = om.create_box(*n, perturb=perturb)
    rng = np.random.default_rng(seed)
    nv, nc = len(m.verts), m.ncells
    vnew = rng.permutation(nv)                       # new id of old vertex
    coords = np.empty_like(m.verts)
    coords[vnew] = m.verts
    rots = _rotations()
    cells = np.empty((nc, 8), dtype=np.int64)
    for c in range(nc):
        perm, flips = rots[rng.integers(24)]
        for a, b, cc in itertools.product((0, 1), repeat=3):
            new = (a, b, cc)
            old = [0, 0, 0]
            for d in range(3):                       # old coordinate d = (flipped) new coordinate perm[d]
                old[d] = 1 - new[perm[d]] if flips[d] else new[perm[d]]
            cells[c, 4 * a + 2 * b + cc] = vnew[m.geom_dofmap[c, 4 * old[0] + 2 * old[1] + old[2]]]
    order = rng.permutation(nc)
    cells = cells[order]
    cen = coords[cells].mean(axis=1)
    frac = (cen[:, 0] + 0.7 * cen[:, 1] + 0.4 * cen[:, 2]) / 2.1
    owner = np.minimum((frac * nranks).astype(np.int32), nranks - 1)
    return m, cells, owner, coords


@pytest.mark.parametrize("n,perturb,nranks,degrees", [((4, 3, 3), 0.15, 3, (1, 2, 3)), ((3, 3, 2), 0.0, 5, (4,)),
                                                     ((3, 2, 2), 0.2, 1, (2, 5))])
def test_general_mesh_ghost_layer_apply_matches_structured_oracle(n, perturb, nranks, degrees):
    from pmg_dolfinx_b200 import api
    m, cells, owner, coords = scrambled_box(n, perturb, 5, nranks)
    meshes = [api.GhostLayerMesh(cells, owner, coords, r, nranks) for r in range(nranks)]
    # ghost layer: exactly the foreign cells sharing a vertex with an owned cell (src/mesh.hpp:25-46)
    for r, gm in enumerate(meshes):
        mine = np.flatnonzero(owner == r)
        assert np.array_equal(gm.cell_gid[: gm.n_owned_cells], mine)
        touched = np.zeros(len(coords), dtype=bool)
        touched[cells[mine].ravel()] = True
        expect = np.flatnonzero((owner != r) & touched[cells].any(axis=1))
        assert sorted(gm.cell_gid[gm.n_owned_cells:].tolist()) == expect.tolist()
        assert np.allclose(gm.xgeom[gm.geom_dofmap], coords[cells[gm.cell_gid]])
        assert sorted(np.concatenate([gm.lcells, gm.bcells]).tolist()) == list(range(gm.n_cells))
    for P in degrees:
        dm, bc, nd = om.dofmap(m, P), om.bc_marker(m, P), om.num_dofs(m, P)
        tree = cKDTree(om.dof_coords(m, P))
        G, _ = oo.geometry_factors(m.verts, m.geom_dofmap, P)
        kap_cell = 1.0 + np.arange(m.ncells) % 3          # per-cell coefficient, identified through the centroid
        ctree = cKDTree(m.verts[m.geom_dofmap].mean(axis=1))
        xg = np.random.default_rng(9).uniform(-1, 1, nd)
        yo = oo.apply(P, dm, G, kap_cell.astype(float), bc, xg)
        spaces = [gm.space(P) for gm in meshes]
        can, seen = [], np.zeros(nd, dtype=int)
        for r, (gm, sp) in enumerate(zip(meshes, spaces)):
            dist, idx = tree.query(sp.coords)
            assert dist.max() < 1e-12
            can.append(idx)
            assert len(np.unique(idx)) == len(idx)            # one local dof per physical dof
            assert np.array_equal(sp.bc, bc[idx])             # exterior-facet marker, on ghosts too
            seen[idx[: sp.n_owned]] += 1
            assert (sp.dofmap[gm.lcells] < sp.n_owned).all()  # lcells hold no ghost dof (src/mesh.hpp:119-138)
            if len(gm.bcells) and gm.n_owned_cells < gm.n_cells:
                assert all((sp.dofmap[c] >= sp.n_owned).any() or c >= gm.n_owned_cells for c in gm.bcells)
            assert sp.n_global == nd
        assert (seen == 1).all()                              # ownership partitions the dofs
        # the same physical dof carries the same library-global id on every rank
        gid_of_can = -np.ones(nd, dtype=np.int64)
        for sp, idx in zip(spaces, can):
            assert ((gid_of_can[idx] == -1) | (gid_of_can[idx] == sp.l2g)).all()
            gid_of_can[idx] = sp.l2g
        # Scatterer lists: what r sends to q is what q expects from r, in order
        for r, sp in enumerate(spaces):
            for k, q in enumerate(sp.send_ranks):
                sq = spaces[q]
                kk = list(sq.recv_ranks).index(r)
                sent = sp.l2g[sp.send_idx[sp.send_offsets[k]:sp.send_offsets[k + 1]]]
                slots = sq.recv_idx[sq.recv_offsets[kk]:sq.recv_offsets[kk + 1]]
                assert np.array_equal(sent, sq.l2g[sq.n_owned + slots])
            assert sorted(sp.recv_idx.tolist()) == list(range(sp.n_ghost))
        # operator through each rank's arrays with the halo update emulated by the lists
        owned_vals = [xg[idx[: sp.n_owned]] for sp, idx in zip(spaces, can)]
        for r, (gm, sp, idx) in enumerate(zip(meshes, spaces, can)):
            x = np.zeros(sp.n_owned + sp.n_ghost)
            x[: sp.n_owned] = owned_vals[r]
            for kk, q in enumerate(sp.recv_ranks):            # forward scatter (src/vector.hpp:186-238)
                sq = spaces[q]
                k = list(sq.send_ranks).index(r)
                vals = owned_vals[q][sq.send_idx[sq.send_offsets[k]:sq.send_offsets[k + 1]]]
                x[sp.n_owned + sp.recv_idx[sp.recv_offsets[kk]:sp.recv_offsets[kk + 1]]] = vals
            assert np.array_equal(x, xg[idx])
            Gl, detJ = oo.geometry_factors(gm.xgeom, gm.geom_dofmap, P)
            assert (detJ > 0).all()                           # proper rotations keep the orientation
            kap = kap_cell[ctree.query(gm.xgeom[gm.geom_dofmap].mean(axis=1))[1]].astype(float)
            y = np.zeros_like(x)
            oo.apply_cells(P, sp.dofmap, Gl, kap, sp.bc, x, y, gm.lcells)
            oo.apply_cells(P, sp.dofmap, Gl, kap, sp.bc, x, y, gm.bcells)
            ref = yo[idx[: sp.n_owned]]
            assert np.linalg.norm(y[: sp.n_owned] - ref) <= 1e-12 * np.linalg.norm(yo)
    for gm in meshes:
        gm.close()


def test_ghostmesh_rejects_bad_input():
    from pmg_dolfinx_b200.capi import lib, ptr
    import ctypes
    cells = np.arange(8, dtype=np.int64).reshape(1, 8)
    owner = np.zeros(1, dtype=np.int32)
    xs = np.zeros((8, 3))
    h = ctypes.c_void_p()
    assert lib.pmgx_ghostmesh_create(0, 1, 1, ptr(cells), ptr(owner), 7, ptr(xs), ctypes.addressof(h)) != 0  # vertex id 7 >= 7
    owner[0] = 3
    assert lib.pmgx_ghostmesh_create(0, 2, 1, ptr(cells), ptr(owner), 8, ptr(xs), ctypes.addressof(h)) != 0  # owner out of range


def test_ghostmesh_ranks_without_cells():
    """A partition may leave ranks empty (more ranks than parts): they get empty arrays and the global dof count,
    the rank that owns everything gets the single-domain arrays with no ghosts."""
    from pmg_dolfinx_b200 import api
    m = om.create_box(2, 2, 2)
    cv = m.geom_dofmap.astype(np.int64)
    owner = np.zeros(len(cv), dtype=np.int32)
    for rank in range(3):
        gm = api.GhostLayerMesh(cv, owner, m.verts, rank, 3)
        sp = gm.space(2)
        assert sp.n_global == 125
        if rank == 0:
            assert (gm.n_cells, gm.n_owned_cells, sp.n_owned, sp.n_ghost) == (8, 8, 125, 0)
            assert len(gm.lcells) == 8 and len(gm.bcells) == 0 and len(sp.send_ranks) == 0 and len(sp.recv_ranks) == 0
        else:
            assert (gm.n_cells, gm.n_owned_cells, sp.n_owned, sp.n_ghost) == (0, 0, 0, 0)
            assert len(gm.lcells) == 0 and len(gm.bcells) == 0
        gm.close()
