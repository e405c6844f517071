// 'Solver type = Direct': band Cholesky on the device (direct_band.cuh) behind
// gf_nl_newton_solve / gf_lin_step with type_lin = 1. Replaces SparseDirectUMFPACK
// (nonlinear_elasticity.cc:1192-1200, linear_elasticity.cc:556-563).
#include <cstdlib>

#include "direct_band.cuh"
#include "gf_context.h"
#include "rcm.h"

namespace gf
{
  namespace
  {
    struct CudaLaunch
    {
      gf_context &c;
      template <class... KA, class... A>
      void operator()(unsigned grid, unsigned block, size_t smem, void (*kernel)(KA...), A... args)
      {
        kernel<<<grid, block, smem, c.stream>>>(args...);
        c.prof_sink->kernel_launches++;
      }
    };

    // ordering + memory check, once per handle
    void analyse(gf_context &c)
    {
      DirectBand &d = c.direct;
      d.analysed    = true;
      d.usable      = false;
      if (c.comm != nullptr || c.n_owned_nodes != c.n_nodes)
        {
          d.why = "partitioned handle";
          return;
        }
      const int64_t nn = c.n_owned_nodes;
      // n w^2 work and n w memory: a method for small / medium systems. Beyond this size the
      // ordering is not even attempted (GF_DIRECT_MAX_DOFS raises the limit)
      int64_t max_dofs = 400000;
      if (const char *env = std::getenv("GF_DIRECT_MAX_DOFS"))
        max_dofs = atoll(env);
      if (nn * c.dim > max_dofs)
        {
          d.why = std::to_string(nn * c.dim) + " DoFs exceed the direct solver's size limit of " +
                  std::to_string(max_dofs);
          return;
        }
      std::vector<int32_t> brow(nn + 1), bcol(c.bcol.n), node_new;
      c.brow_ptr.download(brow.data(), c.stream);
      c.bcol.download(bcol.data(), c.stream);
      const int64_t wn = rcm_order(nn, brow.data(), bcol.data(), node_new);
      d.n              = nn * c.dim;
      d.w              = wn * c.dim + c.dim - 1;
      d.ld             = d.w + DB_NB; // W + 1, W = w + DB_NB - 1
      size_t free_b = 0, total_b = 0;
      GF_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
      double budget = 0.4 * double(free_b);
      if (const char *env = std::getenv("GF_DIRECT_BUDGET_MB"))
        budget = std::min(budget, atof(env) * 1024.0 * 1024.0);
      const double need = double(d.n) * double(d.ld) * sizeof(double);
      if (need > budget)
        {
          d.why = "band of " + std::to_string(int64_t(need / (1024.0 * 1024.0))) +
                  " MB exceeds the memory budget";
          return;
        }
      d.node_new.upload(node_new.data(), node_new.size(), c.stream);
      d.band.alloc(size_t(d.n) * size_t(d.ld));
      d.xp.alloc(size_t(d.n));
      d.tmp.alloc_zero(DB_NB, c.stream);
      d.info.alloc_zero(1, c.stream);
      d.usable = true;
    }
  } // namespace

  bool direct_available(gf_context &c)
  {
    if (c.direct_mode == 2)
      return false;
    if (!c.direct.analysed)
      analyse(c);
    GF_REQUIRE(c.direct.usable || c.direct_mode != 1, GF_ERR_UNSUPPORTED,
               "direct solver not available: " + c.direct.why);
    return c.direct.usable;
  }

  // band <- A, A = L L^T. false: A is not positive definite (the caller falls back to the CG)
  bool direct_factor(gf_context &c, const double *A)
  {
    DirectBand &d = c.direct;
    ProfScope   ps(c, Profile::CG_VEC, 0);
    d.factored = false;
    GF_CUDA_CHECK(cudaMemsetAsync(d.band.p, 0, d.band.n * sizeof(double), c.stream));
    GF_CUDA_CHECK(cudaMemsetAsync(d.info.p, 0, sizeof(int), c.stream));
    CudaLaunch launch{c};
    launch(unsigned(std::min<int64_t>(c.n_owned_nodes, 65535)), 128u, size_t(0), db_fill_kernel,
           c.n_owned_nodes, c.dim, (const int32_t *)c.brow_ptr.p, (const int64_t *)c.val_ptr.p,
           (const int32_t *)c.bcol.p, A, (const int32_t *)d.node_new.p, d.ld, d.band.p);
    db_factor(launch, d.n, d.w, d.ld, d.band.p, d.info.p);
    GF_CUDA_CHECK(cudaGetLastError());
    int info = 0;
    GF_CUDA_CHECK(cudaMemcpyAsync(&info, d.info.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    d.factored = info == 0;
    return d.factored;
  }

  // x = A^-1 b with the current factor (internal dof order on both sides; b and x may alias)
  void direct_solve(gf_context &c, const double *b, double *x)
  {
    DirectBand &d = c.direct;
    GF_REQUIRE(d.factored, GF_ERR_INVALID_ARG, "direct solver: no factorisation");
    ProfScope     ps(c, Profile::CG_VEC, 0);
    CudaLaunch    launch{c};
    const unsigned grid = unsigned(std::min<int64_t>((d.n + 255) / 256, 65535));
    launch(grid, 256u, size_t(0), db_permute_kernel, c.n_owned_nodes, c.dim,
           (const int32_t *)d.node_new.p, true, const_cast<double *>(b), d.xp.p);
    db_solve(launch, d.n, d.w, d.ld, (const double *)d.band.p, d.xp.p, d.tmp.p);
    launch(grid, 256u, size_t(0), db_permute_kernel, c.n_owned_nodes, c.dim,
           (const int32_t *)d.node_new.p, false, x, d.xp.p);
    GF_CUDA_CHECK(cudaGetLastError());
  }
} // namespace gf
