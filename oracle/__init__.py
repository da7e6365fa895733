"""CPU oracle for the p-multigrid hot path of Wells-Group/pmg-dolfinx.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU stand-in.  The product path (``pmg_dolfinx_b200``) never imports
it and fails loudly when its CUDA library is missing.

Pinning status (see DESIGN.md section "Oracle"):
  * ``tqli``           -- pinned to the reference's golden vectors
                          (python_tests/tqli.py:63-99), tests/golden/tqli.json.
  * device kernels     -- the reference's own ``__global__`` kernels
                          (laplacian.hpp, interpolate.hpp, csr.hpp, vector.hpp)
                          are compiled from /root/reference into oracle/_ref/
                          (see oracle/ref_build/) and compared on the GPU box.
  * Chebyshev-4, CG,   -- the reference's own Python prototypes
    Lanczos estimate     (python_tests/chebyshev.py, cg.py, tqli.py) are run in the
                          build container on a small SPD matrix with only their
                          dolfinx/petsc4py IMPORTS stubbed
                          (scripts/make_golden_solvers.py); their outputs are the
                          fixture tests/golden/solvers_ref.npz.
  * V-cycle, partition -- parity unpinned by the reference (python_tests/pmg.py is a
                          DOLFINx script, not importable; no numeric fixtures);
                          pinned by the oracle and analytic known answers in tests/.

Numerical conventions: GLL nodes == GLL quadrature
points on [0, 1] in ascending order; hex local dof index = ix*nd^2 + iy*nd + iz
(src/laplacian.hpp:168-173); global numbering is lexicographic on the GLL grid.
"""
