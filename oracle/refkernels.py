"""ctypes front-end of oracle/_ref/libref_kernels.so: the reference's OWN kernels, compiled from
/root/reference by oracle/ref_build/Makefile (oracle; TEST INFRASTRUCTURE ONLY).

Used to pin the numpy oracle and the CUDA path against outputs of the reference itself on the
GPU box, and as a same-GPU timing comparator in bench.py.  Device pointers are passed as ints
(torch ``data_ptr()``)."""
import ctypes
import os

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_kernels.so")


def available():
    return os.path.exists(_SO)


def load():
    import torch  # noqa: F401  (loads libcudart before the kernels library)
    L = ctypes.CDLL(_SO)
    vp, i = ctypes.c_void_p, ctypes.c_int
    L.ref_geometry.argtypes = [i, vp, vp, vp, vp, vp, vp, i]
    L.ref_stiffness.argtypes = [i, vp, vp, vp, vp, vp, vp, vp, i, vp, i]
    for f in (L.ref_pack, L.ref_unpack, L.ref_unpack_add):
        f.argtypes = [i, vp, vp, vp]
    L.ref_interpolate_Q1Q2.argtypes = [i, vp, vp, i, vp, i, vp, vp, vp, vp, vp]
    L.ref_interpolate_Q2Q1.argtypes = [i, vp, vp, i, vp, i, vp, vp, vp, vp, vp, vp]
    L.ref_spmv.argtypes = [i, vp, vp, vp, vp, vp, vp]
    L.ref_spmvT.argtypes = [i, vp, vp, vp, vp, vp, vp]
    L.ref_tqli.argtypes = [vp, vp, i]
    return L
