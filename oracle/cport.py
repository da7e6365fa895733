"""ctypes front-end of oracle/c/pmg_oracle.c (oracle; TEST INFRASTRUCTURE ONLY).

The C/OpenMP restatement is the checker at sizes numpy cannot reach and the timed CPU stand-in
("port") of bench.py's cpu_baseline / --impl reference legs.  See the header of pmg_oracle.c for
the reference lines each function follows."""
import ctypes
import os
import subprocess

import numpy as np

from . import gll

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_DIR, "libpmg_oracle.so")


def _load():
    if not os.path.exists(_SO):
        subprocess.run(["make"], cwd=_DIR, check=True, capture_output=True)
    L = ctypes.CDLL(_SO)
    vp, i64, dbl, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int
    L.orc_num_threads.restype = i32
    L.orc_set_num_threads.argtypes = [i32]
    L.orc_geometry.argtypes = [i32, vp, vp, vp, vp, i64, vp, vp]
    L.orc_apply.argtypes = [i32, vp, vp, vp, vp, vp, i64, i64, vp, vp]
    L.orc_spmv.argtypes = [i64, vp, vp, vp, vp, vp]
    L.orc_axpy.argtypes = [i64, dbl, vp, vp, vp]
    L.orc_dot.argtypes = [i64, vp, vp]
    L.orc_dot.restype = dbl
    return L


L = _load()


def _p(a):
    return a.ctypes.data


def num_threads():
    return int(L.orc_num_threads())


def set_num_threads(n):
    """Explicit OpenMP thread count (torch.distributed.run exports OMP_NUM_THREADS=1)."""
    L.orc_set_num_threads(int(n))


def geometry(verts, geom_dofmap, P):
    x, w, _ = gll.tables(P)
    nc = geom_dofmap.shape[0]
    nq = (P + 1) ** 3
    G = np.empty((nc, nq, 6))
    detj = np.empty((nc, nq))
    v = np.ascontiguousarray(verts, dtype=np.float64)
    gd = np.ascontiguousarray(geom_dofmap, dtype=np.int32)
    L.orc_geometry(P + 1, _p(x), _p(w), _p(v), _p(gd), nc, _p(G), _p(detj))
    return G, detj


class Apply:
    """y = A x with everything pre-staged as contiguous arrays."""

    def __init__(self, P, dofmap, G, kappa, bc, ndofs):
        self.P, self.nd = P, int(ndofs)
        self.D = np.ascontiguousarray(gll.tables(P)[2])
        self.dm = np.ascontiguousarray(dofmap, dtype=np.int32)
        self.G = np.ascontiguousarray(G, dtype=np.float64)
        self.kappa = np.ascontiguousarray(kappa, dtype=np.float64)
        self.bc = np.ascontiguousarray(bc, dtype=np.int8)

    def __call__(self, x, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if y is None:
            y = np.empty(self.nd)
        L.orc_apply(self.P + 1, _p(self.D), _p(self.dm), _p(self.G), _p(self.kappa), _p(self.bc),
                    self.dm.shape[0], self.nd, _p(x), _p(y))
        return y


def spmv(A, x):
    y = np.empty(A.shape[0])
    ip = np.ascontiguousarray(A.indptr, dtype=np.int32)
    ix = np.ascontiguousarray(A.indices, dtype=np.int32)
    va = np.ascontiguousarray(A.data, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    L.orc_spmv(A.shape[0], _p(ip), _p(ix), _p(va), _p(x), _p(y))
    return y
