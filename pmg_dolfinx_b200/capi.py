"""ctypes binding of the C ABI declared in include/pmgx.h.

The prototypes are parsed from the header itself so the binding can never drift from the
declared boundary.  There is no fallback: if libpmgx.so (built by ``__graft_entry__.build()``
/ ``make -C pmg_dolfinx_b200/csrc``) is missing, importing this module raises.
"""
import ctypes
import os
import re

# torch bundles its own libnccl.so.2 (2.28.x); it must be loaded BEFORE libpmgx.so so that both
# share one NCCL (the soname resolves to the copy already in the process), otherwise the system
# 2.27.3 copy would shadow symbols torch needs.
import torch  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "pmgx.h")
# PMGX_LIB: an alternative build of the same library (A/B builds of kernel compile-time parameters)
LIBPATH = os.environ.get("PMGX_LIB") or os.path.join(_HERE, "libpmgx.so")

_SCALARS = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "double": ctypes.c_double,
    "uint64_t": ctypes.c_uint64,
    "size_t": ctypes.c_size_t,
}


def _ctype(decl):
    decl = decl.replace("const ", "").strip()
    if "*" in decl or decl.split(" ")[0].endswith("_fn"):  # pointers and callback typedefs
        return ctypes.c_void_p
    base = decl.rsplit(" ", 1)[0].strip() if " " in decl and decl not in _SCALARS else decl
    for k in ("long long", "uint64_t", "double", "size_t", "int"):
        if base == k or decl.startswith(k + " ") or decl == k:
            return _SCALARS[k]
    raise ValueError(f"cannot map C type of '{decl}'")


def parse_header(path=HEADER):
    """Return {name: (restype, [argtypes])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"#[^\n]*", "", src)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(pmgx_\w+)\s*\(([^;{}]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef") or ret.startswith("struct"):
            continue
        if ret == "const char*" or ret == "const char *":
            restype = ctypes.c_char_p
        elif "*" in ret:
            restype = ctypes.c_void_p
        else:
            restype = _ctype(ret)
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                argtypes.append(_ctype(" ".join(a.split())))
        protos[name] = (restype, argtypes)
    return protos


class PmgxError(RuntimeError):
    """Raised for any non-zero status; mirrors the std::runtime_error the reference throws
    (src/laplacian.hpp:346,479, src/vector.hpp:343, src/cg.hpp:125,138)."""

    def __init__(self, code, msg):
        super().__init__(f"pmgx error {code}: {msg}")
        self.code = code
        self.msg = msg


def load(path=LIBPATH):
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
            "g.build()' or make -C pmg_dolfinx_b200/csrc). There is no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = load()


def check(status):
    if status != 0:
        raise PmgxError(status, lib.pmgx_last_error_string().decode())


def ptr(x):
    """Device/host pointer of a torch tensor, numpy array, int or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    if isinstance(x, ctypes.c_void_p):
        return x.value
    raise TypeError(f"cannot take a pointer of {type(x)}")
