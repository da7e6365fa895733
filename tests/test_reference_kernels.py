"""Pinning against the reference ITSELF: its own kernels (src/laplacian.hpp, vector.hpp,
interpolate.hpp, csr.hpp) and host tqli (src/cg.hpp), compiled from /root/reference into
oracle/_ref/libref_kernels.so (oracle/ref_build/Makefile), are run on the same inputs as the
numpy oracle and the CUDA path.  Launch lists use all cells in one list so the reference's
block_id-indexed G (quirk Q1) is consistent, and PMGX_LAP_LITERAL_DETJ reproduces its detJ
expression (quirk Q17)."""
import numpy as np
import pytest

from oracle import gll, mesh as om, operator as oo, solvers as osol, refkernels
from helpers import OracleLevel, GpuLevel, rel

needs_ref = pytest.mark.skipif(not refkernels.available(), reason="oracle/_ref not built")


@needs_ref
def test_reference_tqli_host():
    """CPU: reference tqli vs the oracle restatement and the product's pmgx_tqli."""
    from pmg_dolfinx_b200 import api
    L = refkernels.load()
    rng = np.random.default_rng(3)
    for n in (2, 5, 10, 19, 20):
        d, e = rng.uniform(0.3, 1.5, n), np.append(rng.uniform(0.1, 0.7, n - 1), 0.0)
        dr, er = d.copy(), e.copy()
        assert L.ref_tqli(dr.ctypes.data, er.ctypes.data, n) == 0
        do, eo = d.copy(), e.copy()
        assert osol.tqli(do, eo) == 0
        assert np.array_equal(np.sort(dr), np.sort(do))          # same algorithm, same arithmetic
        assert np.allclose(np.sort(api.tqli(d, e)), np.sort(dr), rtol=1e-14)


def _ref_G(ctx, L, ol):
    """G[c][q][6] from the reference's geometry_computation kernel."""
    import torch
    m, P = ol.mesh, ol.P
    nq = (P + 1) ** 3
    d_x, d_gd = ctx.to_device(m.verts), ctx.to_device(m.geom_dofmap)
    d_dphi = ctx.to_device(oo.trilinear_dphi(P))
    d_w = ctx.to_device(oo.weights_3d(P))
    ent = ctx.to_device(np.arange(m.ncells, dtype=np.int32))
    G = ctx.zeros(m.ncells * nq * 6)
    ctx.sync()
    assert L.ref_geometry(P, d_x.data_ptr(), G.data_ptr(), d_gd.data_ptr(), d_dphi.data_ptr(), d_w.data_ptr(),
                          ent.data_ptr(), m.ncells) == 0
    return G, ent


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("P", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("perturb", [0.0, 0.2])
def test_reference_geometry_and_stiffness(ctx, P, perturb):
    L = refkernels.load()
    ol = OracleLevel(om.create_box(3, 2, 4, perturb=perturb), P, literal_detj=True)
    G, ent = _ref_G(ctx, L, ol)
    Gr = G.cpu().numpy().reshape(ol.G.shape)
    assert np.abs(Gr - ol.G).max() <= 1e-13 * np.abs(ol.G).max()       # oracle == reference geometry
    gl = GpuLevel(ctx, ol, flags=1)
    assert np.abs(gl.op.geometry_factors() - Gr).max() <= 1e-13 * np.abs(Gr).max()
    x = np.random.default_rng(42).uniform(-1, 1, ol.nd)
    d_x, y = ctx.to_device(x), ctx.zeros(ol.nd)
    d_D = ctx.to_device(gll.tables(P)[2])
    ctx.sync()
    assert L.ref_stiffness(P, d_x.data_ptr(), gl.kappa.data_ptr(), y.data_ptr(), G.data_ptr(), gl.dofmap.data_ptr(),
                           d_D.data_ptr(), ent.data_ptr(), ol.mesh.ncells, gl.bc.data_ptr(), 1) == 0
    yr = y.cpu().numpy()
    e2, einf = rel(ol.A(x), yr)
    assert e2 < 1e-12 and einf < 1e-12                                   # oracle == reference operator
    e2, einf = rel(gl.apply(x), yr)
    assert e2 < 1e-12 and einf < 1e-12                                   # CUDA path == reference operator


@needs_ref
@pytest.mark.gpu
def test_exact_detj_equals_reference_on_axis_aligned_mesh(ctx):
    """On the benchmark meshes (uniform cube) the default (exact detJ) path is the reference's."""
    L = refkernels.load()
    ol = OracleLevel(om.create_box(4, 4, 4), 3)
    G, _ = _ref_G(ctx, L, ol)
    gl = GpuLevel(ctx, ol)
    Gr = G.cpu().numpy().reshape(ol.G.shape)
    assert np.abs(gl.op.geometry_factors() - Gr).max() <= 1e-14 * np.abs(Gr).max()


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("Pc,Pf", [(1, 3), (2, 4), (1, 2)])
def test_reference_interpolation_kernels(ctx, Pc, Pf):
    import scipy.sparse as sp
    from pmg_dolfinx_b200 import api
    L = refkernels.load()
    mesh = om.create_box(3, 3, 2)
    oc, of = OracleLevel(mesh, Pc), OracleLevel(mesh, Pf)
    M = sp.csr_matrix(oo.local_interp_matrix(Pc, Pf))
    MT = sp.csr_matrix(oo.local_interp_matrix(Pc, Pf).T)
    dev = lambda a, dt: ctx.to_device(np.ascontiguousarray(a, dtype=dt))
    dmc, dmf = dev(oc.dm, np.int32), dev(of.dm, np.int32)
    cells = dev(np.arange(mesh.ncells), np.int32)
    rng = np.random.default_rng(5)
    xc, xf = rng.uniform(-1, 1, oc.nd), rng.uniform(-1, 1, of.nd)
    vc, vf = dev(xc, np.float64), ctx.zeros(of.nd)
    mp_, mc_, mv_ = dev(M.indptr, np.int32), dev(M.indices, np.int32), dev(M.data, np.float64)
    ctx.sync()
    assert L.ref_interpolate_Q1Q2(mesh.ncells, cells.data_ptr(), dmc.data_ptr(), oc.dm.shape[1], dmf.data_ptr(),
                                  of.dm.shape[1], vc.data_ptr(), vf.data_ptr(), mp_.data_ptr(), mc_.data_ptr(),
                                  mv_.data_ptr()) == 0
    pr = vf.cpu().numpy()
    it = api.Interpolator(ctx, Pc, Pf, dmc, dmf, oc.nd, of.nd, np.arange(mesh.ncells, dtype=np.int32), np.zeros(0, np.int32))
    a, b = api.Vector(ctx, oc.nd), api.Vector(ctx, of.nd)
    a.copy_from_host(xc)
    it.interpolate(a, b)
    assert rel(b.data_copy(), pr)[1] < 1e-13 and rel(oo.prolong(Pc, Pf, oc.dm, of.dm, xc, of.nd), pr)[1] < 1e-13
    mult = dev(oo.multiplicity(of.dm, of.nd), np.float64)
    tp_, tc_, tv_ = dev(MT.indptr, np.int32), dev(MT.indices, np.int32), dev(MT.data, np.float64)
    wf, wc = dev(xf, np.float64), ctx.zeros(oc.nd)
    ctx.sync()
    assert L.ref_interpolate_Q2Q1(mesh.ncells, cells.data_ptr(), dmc.data_ptr(), oc.dm.shape[1], dmf.data_ptr(),
                                  of.dm.shape[1], wc.data_ptr(), wf.data_ptr(), tp_.data_ptr(), tc_.data_ptr(),
                                  tv_.data_ptr(), mult.data_ptr()) == 0
    rr = wc.cpu().numpy()
    b.copy_from_host(xf)
    it.reverse_interpolate(b, a)
    assert rel(a.data_copy(), rr)[1] < 1e-13 and rel(oo.restrict(Pc, Pf, oc.dm, of.dm, xf, oc.nd), rr)[1] < 1e-13


@needs_ref
@pytest.mark.gpu
def test_reference_spmv_and_pack(ctx):
    from pmg_dolfinx_b200 import api
    from pmg_dolfinx_b200.capi import lib, check, ptr
    L = refkernels.load()
    ol = OracleLevel(om.create_box(4, 3, 3, perturb=0.1), 1)
    A = oo.assemble_csr(1, ol.dm, ol.G, ol.kappa, ol.bc, ol.nd)
    dev = lambda a, dt: ctx.to_device(np.ascontiguousarray(a, dtype=dt))
    ip, ix, va = dev(A.indptr, np.int32), dev(A.indices, np.int32), dev(A.data, np.float64)
    x = np.random.default_rng(1).uniform(-1, 1, ol.nd)
    dx, y = dev(x, np.float64), ctx.zeros(ol.nd)
    ctx.sync()
    assert L.ref_spmv(ol.nd, va.data_ptr(), ip.data_ptr(), ip.data_ptr() + 4, ix.data_ptr(), dx.data_ptr(), y.data_ptr()) == 0
    yr = y.cpu().numpy()
    B = api.MatrixOperator(ctx, A.indptr, A.indptr[1:], A.indices, A.data)
    xv, yv = api.Vector(ctx, ol.nd), api.Vector(ctx, ol.nd)
    xv.copy_from_host(x)
    B(xv, yv)
    assert rel(yv.data_copy(), yr)[1] < 1e-14 and rel(A @ x, yr)[1] < 1e-14
    idx = dev(np.random.default_rng(2).integers(0, ol.nd, 300), np.int32)
    o1, o2 = ctx.zeros(300), ctx.zeros(300)
    ctx.sync()
    assert L.ref_pack(300, idx.data_ptr(), dx.data_ptr(), o1.data_ptr()) == 0
    check(lib.pmgx_pack(ctx.h, 300, ptr(idx), ptr(dx), ptr(o2)))
    ctx.sync()
    assert np.array_equal(o1.cpu().numpy(), o2.cpu().numpy())
