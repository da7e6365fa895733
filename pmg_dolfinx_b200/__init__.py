"""pmg_dolfinx_b200 -- B200-native p-multigrid hot path behind the pmg-dolfinx operator API.

The product is libpmgx.so (hand-written sm_100a CUDA + NCCL halo, C ABI in include/pmgx.h).
Importing the package loads it and fails loudly if it has not been built; there is no CPU
fallback.  ``api`` mirrors the reference's class names on top of the C ABI.
"""
from .capi import lib, PmgxError  # noqa: F401
from . import api  # noqa: F401
