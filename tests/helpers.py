"""Shared builders for the parity tests: one problem, two implementations (oracle / CUDA)."""
import numpy as np

from oracle import mesh as om, operator as oo


class OracleLevel:
    """Oracle-side data of one degree on a (possibly perturbed) box mesh."""

    def __init__(self, mesh, P, kappa=2.0, literal_detj=False):
        self.mesh, self.P = mesh, P
        self.dm = om.dofmap(mesh, P)
        self.bc = om.bc_marker(mesh, P)
        self.nd = om.num_dofs(mesh, P)
        # scalar, or one value per cell (cell_constants of src/laplacian.hpp:230)
        self.kappa = np.full(mesh.ncells, kappa) if np.isscalar(kappa) else np.ascontiguousarray(kappa, dtype=np.float64)
        self.G, self.detJ = oo.geometry_factors(mesh.verts, mesh.geom_dofmap, P, literal_detj=literal_detj)

    def A(self, x):
        return oo.apply(self.P, self.dm, self.G, self.kappa, self.bc, x)

    def diag(self):
        return oo.diagonal(self.P, self.dm, self.G, self.kappa, self.bc, self.nd)


class GpuLevel:
    """CUDA-side objects for the same data (single rank), built through the C ABI."""

    def __init__(self, ctx, ol, flags=0, lcells=None, bcells=None):
        from pmg_dolfinx_b200 import api
        m = ol.mesh
        self.ctx, self.ol = ctx, ol
        self.dofmap = ctx.to_device(ol.dm)
        self.xgeom = ctx.to_device(m.verts)
        self.gdm = ctx.to_device(m.geom_dofmap)
        self.kappa = ctx.to_device(ol.kappa)
        self.bc = ctx.to_device(ol.bc)
        if lcells is None:
            lcells = np.arange(m.ncells, dtype=np.int32)
            bcells = np.zeros(0, dtype=np.int32)
        self.lcells, self.bcells = lcells, bcells
        self.op = api.MatFreeLaplacian(ctx, ol.P, self.kappa, self.dofmap, self.xgeom, self.gdm, lcells, bcells,
                                       self.bc, ol.nd, 0, None, flags)

    def vec(self, a=None):
        from pmg_dolfinx_b200 import api
        v = api.Vector(self.ctx, self.ol.nd, 0)
        if a is not None:
            v.copy_from_host(a)
        return v

    def apply(self, x):
        xv, yv = self.vec(x), self.vec()
        self.op(xv, yv)
        return yv.data_copy()


def rel(a, b):
    """(relative 2-norm error, relative max-norm error)."""
    return (np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300),
            np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
