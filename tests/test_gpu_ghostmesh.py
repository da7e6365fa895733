"""GPU: a NON-box mesh through the CUDA path.  The library is fed the arrays of the general
ghost-layer builder (random vertex numbering, shuffled and rotated cells: non-lexicographic dofmaps with
edge/face dofs seen in different orientations by neighbouring cells) and must reproduce the structured
single-domain oracle: apply 1e-12, diagonal 1e-13, CG histories 1e-10 with identical iteration counts,
RHS with lifting, three P4->P2->P1 V-cycles with the AMG coarse solver (scripts/mgpu_check.py logic;
the multi-rank version runs in tests/test_gpu_multi.py and in bench.py's parity gate)."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("general", [True, False], ids=["general-mesh", "box-mesh"])
@pytest.mark.parametrize("perturb", [0.0, 0.15], ids=["affine-cells", "perturbed"])
def test_general_mesh_single_gpu_matches_oracle(ctx, perturb, general):
    from pmg_dolfinx_b200 import api
    spec = importlib.util.spec_from_file_location("mgpu_check", os.path.join(ROOT, "scripts", "mgpu_check.py"))
    mc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mc)
    lines = []
    res = mc.run_check(api, ctx, 0, 1, n=(5, 4, 4), perturb=perturb, log=lines.append, general=general)
    assert res["ok"], "\n".join(lines)
    assert res["max_err_over_tol"] < 1.0
    # the run includes the device-built exterior-facet Dirichlet marker (pmgx_bc_marker_exterior) against the host one
    assert all(v == 0 for k, v in res["checks"].items() if "BC marker" in k) and any("BC marker" in k for k in res["checks"])
