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
  * everything else    -- parity unpinned by the reference (it holds no other
                          numeric fixtures and DOLFINx/Basix/PETSc are absent);
                          pinned by analytic known answers in tests/.

Numerical conventions: GLL nodes == GLL quadrature
points on [0, 1] in ascending order; hex local dof index = ix*nd^2 + iy*nd + iz
(src/laplacian.hpp:168-173); global numbering is lexicographic on the GLL grid.
"""
