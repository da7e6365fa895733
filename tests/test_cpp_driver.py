"""The C++ host drivers (examples/*.cpp) are written against the reference's class names through
include/pmgx/dolfinx_acc_compat.hpp: they must compile with a plain C++20 compiler, fail loudly
without a GPU (no CPU fallback) and, on a GPU, reproduce the python harness' results."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")


def _build():
    r = subprocess.run(["make"], cwd=EX, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return os.path.join(EX, "build", "pmg_main"), os.path.join(EX, "build", "cg_main")


def test_cpp_drivers_compile_against_the_shim():
    pmg, cg = _build()
    assert os.access(pmg, os.X_OK) and os.access(cg, os.X_OK)


def test_cpp_driver_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pmg, _ = _build()
    r = subprocess.run([pmg, "--ndofs", "1000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "CUDA" in r.stderr


@pytest.mark.gpu
def test_cpp_pmg_driver_converges_like_the_python_harness(ctx):
    """P1->P2->P4 on ~40k dofs: the residual falls monotonically and ends below 1e-6 relative."""
    pmg, _ = _build()
    r = subprocess.run([pmg, "--ndofs", "40000", "--niter", "10"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    rel = [float(m) for m in re.findall(r"relative ([0-9.eE+-]+)\)", r.stdout)]
    assert len(rel) == 10 and all(b < a for a, b in zip(rel[:-1], rel[1:])), r.stdout
    assert rel[-1] < 1e-6, r.stdout


@pytest.mark.gpu
def test_cpp_cg_driver_matches_oracle_iteration_count_and_eigs(ctx):
    """examples/cg mirror at P3 on a small box: 20 CG iterations, same Lanczos lambda_max as the oracle."""
    import numpy as np
    from oracle import mesh as om, solvers as osol
    from helpers import OracleLevel
    _, cg = _build()
    r = subprocess.run([cg, "--ndofs", str(19 ** 3), "--degree", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    n = [int(v) for v in re.search(r"mesh (\d+) x (\d+) x (\d+) cells", r.stdout).groups()]
    its = int(re.search(r"Number of iterations (\d+)", r.stdout).group(1))
    lmax = float(re.search(r"Computed eigs = \(([0-9.eE+-]+), ([0-9.eE+-]+)\)", r.stdout).group(2))
    ol = OracleLevel(om.create_box(*n), 3)
    _, k, al, be, _, _ = osol.cg(ol.A, 1.0 / ol.diag(), np.zeros(ol.nd), np.ones(ol.nd), 20, 1e-6)
    assert its == k
    assert abs(lmax - osol.lanczos_eigenvalues(al, be)[-1]) < 1e-8
    m = re.search(r"residual ([0-9.eE+-]+) -> ([0-9.eE+-]+)", r.stdout)
    assert float(m.group(2)) < 0.2 * float(m.group(1)), r.stdout
