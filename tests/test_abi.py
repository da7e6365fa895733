"""CPU: the C-ABI library loads, exports every symbol include/pmgx.h declares, and its host-only
entry points (GLL tables, tqli, box mesh / partition / halo lists) agree with the oracle.
No compute call is made here (no GPU in this container)."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import gll, mesh as om, solvers as osol


def test_library_exports_every_declared_symbol():
    from pmg_dolfinx_b200 import capi
    protos = capi.parse_header()
    assert len(protos) >= 70
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIBPATH], capture_output=True, text=True, check=True)
    exported = {l.split()[-1] for l in out.stdout.splitlines() if " T " in l}
    missing = set(protos) - exported
    assert not missing, missing
    for name in protos:
        assert hasattr(capi.lib, name)
    assert capi.lib.pmgx_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch
    from pmg_dolfinx_b200 import api
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.PmgxError) as e:
        api.Context(0)
    assert e.value.code == 2  # PMGX_ERR_CUDA: loud failure, no silent CPU path


@pytest.mark.parametrize("P", range(1, 9))
def test_gll_tables_match_oracle(P):
    from pmg_dolfinx_b200 import api
    x, w, D = api.gll_tables(P)
    xo, wo, Do = gll.tables(P)
    assert np.abs(x - xo).max() < 1e-15 and np.abs(w - wo).max() < 1e-15
    assert np.abs(D - Do).max() < 1e-12 * np.abs(Do).max()
    for Pc in range(1, P + 1):
        assert np.abs(api.gll_interp_1d(Pc, P) - gll.interp_1d(Pc, P)).max() < 1e-13


def test_tqli_golden_and_error_behaviour():
    from pmg_dolfinx_b200 import api
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tqli.json")))
    assert np.allclose(np.sort(api.tqli(g["d"], g["e"])), g["eigs"], rtol=1e-13)
    do, eo = np.array(g["d"]), np.array(g["e"])
    osol.tqli(do, eo)
    assert np.allclose(np.sort(api.tqli(g["d"], g["e"])), np.sort(do), rtol=1e-14)
    with pytest.raises(api.PmgxError, match="Unsupported degree"):
        api.gll_tables(9)


def test_boxmesh_fit_reference_sizes():
    """Mesh-fit routine of the drivers (examples/pmg/main.cpp:412-435) at the BASELINE sizes."""
    from pmg_dolfinx_b200 import api
    assert api.boxmesh_fit(100_000_000, 3) == (154, 154, 155)
    assert api.boxmesh_fit(200_000_000, 6) == (94, 98, 100)
    assert api.boxmesh_fit(100_000_000, 4) == (115, 116, 116)


def test_boxmesh_single_rank_matches_oracle():
    from pmg_dolfinx_b200 import api
    m, o = api.BoxMesh((3, 4, 2)), om.create_box(3, 4, 2)
    assert np.allclose(m.xgeom, o.verts) and np.array_equal(m.geom_dofmap, o.geom_dofmap)
    assert len(m.lcells) == 24 and len(m.bcells) == 0
    for P in (1, 2, 3, 5):
        sp = m.space(P, True)
        assert np.array_equal(sp.dofmap, om.dofmap(o, P)) and np.array_equal(sp.bc, om.bc_marker(o, P))
        assert np.allclose(sp.coords, om.dof_coords(o, P))
        assert sp.n_ghost == 0 and sp.n_owned == om.num_dofs(o, P) == sp.n_global


@pytest.mark.parametrize("n,pg", [((5, 4, 6), (2, 2, 2)), ((4, 4, 4), (2, 1, 1)), ((6, 5, 3), (2, 2, 1)), ((3, 3, 3), (3, 1, 1))])
def test_boxmesh_partition_matches_oracle(n, pg):
    """Two independent implementations (C++ product, numpy oracle) of src/mesh.hpp:16-143 and the
    Scatterer index lists must agree entry for entry."""
    from pmg_dolfinx_b200 import api
    o = om.create_box(*n)
    degs = [1, 2, 4]
    parts = om.partition(o, pg, degs)
    tot = {P: 0 for P in degs}
    for r, p in enumerate(parts):
        m = api.BoxMesh(n, pg, r)
        assert m.n_cells == len(p.cells) and m.n_owned_cells == p.n_owned_cells
        assert np.array_equal(np.sort(m.lcells), np.sort(p.lcells)) and np.array_equal(np.sort(m.bcells), np.sort(p.bcells))
        assert np.allclose(m.xgeom[m.geom_dofmap], p.verts[p.geom_dofmap])
        for P in degs:
            sp, lv = m.space(P), p.levels[P]
            tot[P] += sp.n_owned
            assert (sp.n_owned, sp.n_ghost) == (lv.n_owned, lv.n_ghost)
            assert np.array_equal(sp.l2g, lv.l2g) and np.array_equal(sp.bc, lv.bc)
            assert np.array_equal(sp.l2g[sp.dofmap], lv.l2g[lv.dofmap])
            assert sp.recv_ranks.tolist() == [q for q, _ in lv.nbr_recv]
            for i, (_, slots) in enumerate(lv.nbr_recv):
                assert np.array_equal(sp.recv_idx[sp.recv_offsets[i]:sp.recv_offsets[i + 1]], slots)
            srt = sorted(lv.nbr_send, key=lambda t: t[0])
            assert sp.send_ranks.tolist() == [q for q, _ in srt]
            for i, (_, idx) in enumerate(srt):
                assert np.array_equal(sp.send_idx[sp.send_offsets[i]:sp.send_offsets[i + 1]], idx)
        m.close()
    for P in degs:
        assert tot[P] == om.num_dofs(o, P)   # every dof owned exactly once


def test_boxmesh_perturbation_is_rank_independent():
    from pmg_dolfinx_b200 import api
    n, pg = (4, 4, 4), (2, 2, 1)
    ref = api.BoxMesh(n, (1, 1, 1), 0, perturb=0.2)
    sp0 = ref.space(1, True)
    g = {int(k): sp0.coords[i] for i, k in enumerate(sp0.l2g)}
    assert np.abs(ref.xgeom - om.create_box(*n).verts).max() > 1e-3
    for r in range(4):
        m = api.BoxMesh(n, pg, r, perturb=0.2)
        sp = m.space(1, True)
        for i, k in enumerate(sp.l2g):
            assert np.array_equal(sp.coords[i], g[int(k)])
