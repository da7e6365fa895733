// C++ host driver mirroring examples/cg/main.cpp of the reference (BASELINE config 4): one
// matrix-free Laplacian of degree P (reference: hard-coded 3, quirk Q3; here --degree, default
// 6), Jacobi-preconditioned CG with b = 1, x0 = 0, 20 iterations, rtol 1e-6 (:238-249), Lanczos
// eigenvalue estimate (:251-253), then a 30-iteration Chebyshev solve of the Poisson problem
// with f = 1000 exp(-((x-1/2)^2 + (y-1/2)^2)/0.02), g = 1.3 on the boundary and x0 = 1 with the
// boundary values set (:136-148,158,268-284), written against the reference's class names.
//
//   cg_main [--ndofs N] [--degree P] [--idfile PATH]      (RANK/WORLD_SIZE/LOCAL_RANK from env)
#include "box_setup.hpp"

using T = double;
using namespace dolfinx;
using DeviceVector = acc::Vector<T, acc::Device::CUDA>;

int main(int argc, char** argv)
{
  long long ndofs = 50000;
  int degree = 6;
  std::string idfile;
  for (int i = 1; i + 1 < argc; i += 2)
  {
    if (!std::strcmp(argv[i], "--ndofs"))
      ndofs = (long long)std::atof(argv[i + 1]);
    else if (!std::strcmp(argv[i], "--degree"))
      degree = std::atoi(argv[i + 1]);
    else if (!std::strcmp(argv[i], "--idfile"))
      idfile = argv[i + 1];
    else
    {
      std::printf("usage: %s [--ndofs N] [--degree P] [--idfile path]\n", argv[0]);
      return 1;
    }
  }
  try
  {
    const int rank = box::env_int("RANK", 0), nranks = box::env_int("WORLD_SIZE", 1);
    const int local = box::env_int("LOCAL_RANK", 0);
    std::vector<char> id;
    box::exchange_nccl_id(rank, nranks, idfile, id);
    auto ctx = std::make_shared<pmgx::Context>(local, rank, nranks, nranks > 1 ? id.data() : nullptr);

    box::Mesh mesh(ndofs, degree, rank, nranks, 0.0, 2.0);
    box::Space V;
    mesh.make_space(ctx, degree, V);
    std::shared_ptr<const pmgx::IndexMap> map = V.map;
    const int n_owned = map->size_local(), nt = n_owned + map->num_ghosts();
    if (rank == 0)
      std::printf("mesh %d x %d x %d cells, P%d, %lld dofs on %d rank(s)\n", mesh.nxyz[0], mesh.nxyz[1], mesh.nxyz[2],
                  degree, V.n_global, nranks);

    acc::MatFreeLaplacian<T> op(degree, mesh.kappa.span(), V.dofmap.span(), mesh.xgeom.span(),
                                mesh.geometry_dofmap.span(), {}, {}, mesh.lcells, mesh.bcells, V.bc.span());

    // ---- CG with b = 1 (:238-249)
    DeviceVector b_d(map, 1), x(map, 1);
    b_d.set(T{1.0});
    b_d.scatter_fwd();
    x.set(T{0.0});
    acc::CGSolver<DeviceVector> cg(map, 1);
    cg.set_max_iterations(20);
    cg.set_tolerance(1e-6);
    cg.store_coefficients(true);
    ctx->synchronize();
    auto t0 = std::chrono::steady_clock::now();
    int its = cg.solve(op, x, b_d, true);
    ctx->synchronize();
    const double tcg = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<T> eign = cg.compute_eigenvalues();
    std::array<T, 2> eig_range = {0.1 * eign.back(), 1.1 * eign.back()};
    if (rank == 0)
    {
      std::printf("Number of iterations %d (%.3f s, %.2f Gdof/s per iteration)\n", its, tcg,
                  (double)V.n_global * its / tcg / 1e9);
      std::printf("Computed eigs = (%.10f, %.10f)\nUsing eig range: %.10f - %.10f\n", eign.front(), eign.back(),
                  eig_range[0], eig_range[1]);
    }

    // ---- Poisson right-hand side with lifting of g = 1.3 (:268-275): b = L - A_nobc g on free
    // rows, b = g on Dirichlet rows.  A_nobc is the same operator without a Dirichlet marker.
    const double g = 1.3;
    std::vector<double> f((size_t)nt);
    for (int i = 0; i < nt; ++i)
    {
      const double dx = V.coords[3 * i] - 0.5, dy = V.coords[3 * i + 1] - 0.5;
      f[i] = 1000.0 * std::exp(-(dx * dx + dy * dy) / 0.02);
    }
    pmgx::DeviceArray<double> fvals(f);
    pmgx::check(pmgx_laplacian_rhs(op.handle(map), fvals.p, g, b_d.mutable_array().data()));
    {
      std::vector<std::int8_t> bc_h((size_t)nt), none((size_t)nt, 0);
      cudaMemcpy(bc_h.data(), V.bc.p, nt, cudaMemcpyDeviceToHost);
      pmgx::DeviceArray<std::int8_t> nobc(none);
      acc::MatFreeLaplacian<T> op_nobc(degree, mesh.kappa.span(), V.dofmap.span(), mesh.xgeom.span(),
                                       mesh.geometry_dofmap.span(), {}, {}, mesh.lcells, mesh.bcells, nobc.span());
      std::vector<double> gh((size_t)nt);
      for (int i = 0; i < nt; ++i)
        gh[i] = bc_h[i] ? g : 0.0;
      DeviceVector gv(map, 1), Ag(map, 1);
      gv.copy_from_host(gh);
      op_nobc(gv, Ag);
      acc::axpy(Ag, T{-1.0}, Ag, b_d);                                     // b - A g
      pmgx::check(pmgx_vec_mask_bc(ctx->handle(), Ag.mutable_array().data(), V.bc.p, n_owned));
      acc::pointwise_mult(gv, gv, gv);                                     // g^2 at BC rows
      acc::axpy(b_d, T{1.0 / g}, gv, Ag);                                  // + g at BC rows
    }

    // ---- Chebyshev, 30 iterations, x0 = 1 with the boundary values set (:256-284)
    acc::Chebyshev<DeviceVector> cheb(map, 1, eig_range);
    cheb.set_max_iterations(30);
    {
      std::vector<std::int8_t> bc_h((size_t)nt);
      cudaMemcpy(bc_h.data(), V.bc.p, nt, cudaMemcpyDeviceToHost);
      std::vector<double> sol((size_t)nt, 1.0);
      for (int i = 0; i < nt; ++i)
        if (bc_h[i])
          sol[i] = g;
      x.copy_from_host(sol);
    }
    ctx->synchronize();
    t0 = std::chrono::steady_clock::now();
    cheb.solve(op, x, b_d, true);
    ctx->synchronize();
    const double tch = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rank == 0)
    {
      const std::vector<T>& h = cheb.residual_history();
      std::printf("Chebyshev: 30 iterations in %.3f s; residual %.6e -> %.6e\n", tch, h.front(), h.back());
    }
    return 0;
  }
  catch (const std::exception& e)
  {
    std::fprintf(stderr, "cg_main: %s\n", e.what());
    return 1;
  }
}
