// K5/K6: preconditioned conjugate gradients entirely on the device — deal.II's SolverCG
// recurrences (reference call sites nonlinear_elasticity.cc:1174-1190, linear_elasticity.cc:
// 544-554; algorithm: deal.II 9.5 SolverCG::solve / IterationWorker::do_iteration) with the
// serial SSOR sweeps (PreconditionSelector("ssor",.65) / PreconditionSSOR(1.2)) replaced by a
// node-block Jacobi preconditioner, as BASELINE north_star prescribes.
//
//   startup   r = b - A x ; z = M^-1 r ; res = ||r|| ; check(0,res)
//   iteration p = beta p + z ; v = A p ; alpha = r.z / p.v ;
//             x += alpha p ; r -= alpha v ; z = M^-1 r ; r.r, r.z in the same pass ; check(it,res)
// Scalars (alpha, beta, residual, iteration counter, SolverControl state) live in device memory;
// every kernel returns immediately once the state leaves `iterate`, so the host only polls the
// state every `cg_check_every` iterations — no per-iteration host round trip. Reductions are
// two-stage over a PARTITION-INDEPENDENT tree (reduce.cuh: one partial sum per reduction chunk,
// then a fixed tree over the global chunk list in ONE launch that also does the scalar step; with
// several ranks that launch first gathers all ranks' partials through the NVLink peer windows), so
// the CG history is bitwise reproducible run to run AND for every number of ranks. No FP atomics.
#include <cmath>

#include "reduce.cuh"

namespace gf
{
  namespace
  {
    constexpr int VEC_THREADS = 256;

    template <int DIM>
    __device__ __forceinline__ void apply_block(const double *__restrict__ dinv, const int64_t A,
                                                const double (&r)[DIM], double (&z)[DIM])
    {
#pragma unroll
      for (int i = 0; i < DIM; ++i)
        {
          double v = 0;
#pragma unroll
          for (int j = 0; j < DIM; ++j)
            v = fma(dinv[A * DIM * DIM + i * DIM + j], r[j], v);
          z[i] = v;
        }
    }

    // startup: r = b - v (v = A x) ; z = M^-1 r ; partial sums of r.r and r.z
    // update : x += alpha p ; r -= alpha v ; z = M^-1 r ; partial sums
    // One CTA per reduction chunk (reduce.cuh): the partial sums do not depend on the partition.
    template <int DIM, bool STARTUP>
    __global__ void __launch_bounds__(RED_THREADS)
      cg_update_kernel(const int32_t *__restrict__ chunk_ptr, const CGScalars *s,
                       const double *__restrict__ b, const double *__restrict__ p,
                       const double *__restrict__ v, const double *__restrict__ dinv,
                       double *__restrict__ x, double *__restrict__ r, double *__restrict__ z,
                       double *__restrict__ partials, const int pstride)
    {
      if (s->status != 0)
        return;
      __shared__ double sm[64];
      const double      alpha = STARTUP ? 0.0 : s->alpha;
      double            acc[2] = {0.0, 0.0};
      const int64_t     n0 = chunk_ptr[blockIdx.x], n1 = chunk_ptr[blockIdx.x + 1];
      for (int64_t A = n0 + threadIdx.x; A < n1; A += RED_THREADS)
        {
          double rl[DIM], zl[DIM];
#pragma unroll
          for (int i = 0; i < DIM; ++i)
            {
              const int64_t k = A * DIM + i;
              if (STARTUP)
                rl[i] = b[k] - v[k];
              else
                {
                  x[k] = fma(alpha, p[k], x[k]);
                  rl[i] = fma(-alpha, v[k], r[k]);
                }
              r[k] = rl[i];
            }
          apply_block<DIM>(dinv, A, rl, zl);
#pragma unroll
          for (int i = 0; i < DIM; ++i)
            {
              z[A * DIM + i] = zl[i];
              acc[0]         = fma(rl[i], rl[i], acc[0]);
              acc[1]         = fma(rl[i], zl[i], acc[1]);
            }
        }
      block_sum<2>(acc, sm);
      if (threadIdx.x == 0)
        {
          partials[blockIdx.x]           = acc[0];
          partials[pstride + blockIdx.x] = acc[1];
        }
    }

    // p = beta p + z (beta = 0 in the first iteration: p = z)
    __global__ void __launch_bounds__(VEC_THREADS)
      cg_direction_kernel(const int64_t n, const CGScalars *s, const double *__restrict__ z,
                          double *__restrict__ p)
    {
      if (s->status != 0)
        return;
      const double beta = s->beta;
      for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n;
           i += int64_t(gridDim.x) * blockDim.x)
        p[i] = fma(beta, p[i], z[i]);
    }

    int vec_grid(const gf_context &c, int64_t n)
    {
      const int64_t want = (n + VEC_THREADS - 1) / VEC_THREADS;
      return int(std::max<int64_t>(1, std::min<int64_t>(want, c.max_red_blocks)));
    }

    void launch_update(gf_context &c, bool startup, const double *b, double *x)
    {
      cudaStream_t s = c.stream;
      if (c.n_red_chunks == 0)
        return;
      const int g = c.n_red_chunks;
      if (c.dim == 3)
        {
          if (startup)
            cg_update_kernel<3, true><<<g, RED_THREADS, 0, s>>>(
              c.red_chunk_ptr.p, c.cg_scalars.p, b, nullptr, c.cg_v.p, c.dinv.p, x, c.cg_r.p,
              c.cg_z.p, c.partials.p, c.red_stride);
          else
            cg_update_kernel<3, false><<<g, RED_THREADS, 0, s>>>(
              c.red_chunk_ptr.p, c.cg_scalars.p, b, c.cg_p.p, c.cg_v.p, c.dinv.p, x, c.cg_r.p,
              c.cg_z.p, c.partials.p, c.red_stride);
        }
      else
        {
          if (startup)
            cg_update_kernel<2, true><<<g, RED_THREADS, 0, s>>>(
              c.red_chunk_ptr.p, c.cg_scalars.p, b, nullptr, c.cg_v.p, c.dinv.p, x, c.cg_r.p,
              c.cg_z.p, c.partials.p, c.red_stride);
          else
            cg_update_kernel<2, false><<<g, RED_THREADS, 0, s>>>(
              c.red_chunk_ptr.p, c.cg_scalars.p, b, c.cg_p.p, c.cg_v.p, c.dinv.p, x, c.cg_r.p,
              c.cg_z.p, c.partials.p, c.red_stride);
        }
    }
  } // namespace

  // The same recurrences with z = V-cycle(r). One CG iteration is now several milliseconds of
  // device work, so the host reads the SolverControl state after every residual update (before
  // the V-cycle is enqueued) instead of every cg_check_every iterations.
  static int cg_solve_mg(gf_context &c, const double *val, double *x, const double *b, double tol,
                         int64_t maxit, double bnorm, uint32_t *last_step, double *last_value)
  {
    GF_REQUIRE(mg_active(c), GF_ERR_INVALID_ARG,
               "GF_PRECOND_MULTIGRID selected but no coarse level is attached (gf_mg_attach)");
    for (gf_context *l = &c; l != nullptr; l = l->mg.coarse)
      GF_REQUIRE(l->mg_ops_valid, GF_ERR_INVALID_ARG,
                 "multigrid operators not built: attach the levels before assembling");
    cudaStream_t s = c.stream;
    CGScalars    init{};
    init.tol     = tol;
    init.maxit   = int(std::min<int64_t>(maxit, 2147483647));
    init.status  = 0;
    *c.h_scalars = init;
    GF_CUDA_CHECK(cudaMemcpyAsync(c.cg_scalars.p, c.h_scalars, sizeof(CGScalars),
                                  cudaMemcpyHostToDevice, s));
    GF_CUDA_CHECK(cudaMemsetAsync(c.cg_p.p, 0, c.n_local * sizeof(double), s));
    const int dg = vec_grid(c, c.n_owned);
    auto update = [&](bool startup) {
      ProfScope ps(c, Profile::CG_VEC, 2);
      launch_update(c, startup, b, x);
      reduce_sums(c, 1, startup ? 3 : 4, true); // r.r -> residual check
    };
    auto poll = [&]() {
      GF_CUDA_CHECK(cudaMemcpyAsync(c.h_scalars, c.cg_scalars.p, sizeof(CGScalars),
                                    cudaMemcpyDeviceToHost, s));
      GF_CUDA_CHECK(cudaStreamSynchronize(s));
      comm_check(c);
      for (gf_context *l = &c; l != nullptr; l = l->mg.coarse)
        GF_REQUIRE(*l->h_lmax > 0.0 && std::isfinite(*l->h_lmax), GF_ERR_NOT_CONVERGED,
                   "multigrid: eigenvalue estimate of the smoother failed");
      return c.h_scalars->status;
    };
    // ---- startup: r = b - A x ; check ----
    if (c.comm)
      halo_exchange(c, x);
    op_apply(c, val, x, c.cg_v.p, nullptr);
    update(true);
    GF_CUDA_CHECK(cudaGetLastError());
    bool first_poll = true;
    while (poll() == 0)
      {
        if (first_poll && bnorm >= 0.0 && c.h_scalars->res > bnorm)
          {
            // GF_OPT_CG_INITIAL_GUESS = 1: the caller's guess (the previous Newton update,
            // nonlinear_elasticity.cc:1184) is further from the solution than the zero vector
            // (||b - A x0|| > ||b||): start from zero instead. Same stopping criterion.
            vec_zero(c, x);
            vec_zero(c, c.cg_v.p);
            update(true);
            first_poll = false;
            continue;
          }
        first_poll = false;
        mg_vcycle(c, c.cg_r.p, c.cg_z.p); // z = M^-1 r
        {
          ProfScope ps(c, Profile::CG_VEC, 3);
          launch_dot_chunks(c, c.cg_r.p, c.cg_z.p, true);
          reduce_sums(c, 1, 5, true); // r.z, beta
          cg_direction_kernel<<<dg, VEC_THREADS, 0, s>>>(c.n_owned, c.cg_scalars.p, c.cg_z.p,
                                                         c.cg_p.p);
        }
        if (c.comm)
          halo_exchange(c, c.cg_p.p);
        op_apply(c, val, c.cg_p.p, c.cg_v.p, nullptr);
        {
          ProfScope ps(c, Profile::CG_VEC, 2);
          launch_dot_chunks(c, c.cg_p.p, c.cg_v.p, true);
          reduce_sums(c, 1, 1, true); // p.Ap, alpha
        }
        update(false);
        GF_CUDA_CHECK(cudaGetLastError());
      }
    *last_step  = uint32_t(c.h_scalars->it);
    *last_value = c.h_scalars->res;
    return c.h_scalars->status == 1 ? GF_OK : GF_ERR_NOT_CONVERGED;
  }

  // 'Solver type = Direct' on a handle WITH hanging-node constraint lines: the band Cholesky
  // holds the uncondensed A, the system is the condensed C^T A C (operator.cu). The factor is then
  // the preconditioner of a CG on the condensed operator, z = A^-1 r (restricted to the
  // unconstrained dofs this is the inverse of a Schur complement of A: SPD). C^T A C differs from
  // that Schur complement by a term whose rank is bounded by the number of constrained dofs, so
  // the iteration ends after a handful of steps. Same recurrences and device-resident
  // SolverControl state as cg_solve_mg; x starts from zero; tol is absolute.
  int cg_solve_direct_precond(gf_context &c, const double *val, double *x, const double *b,
                              double tol, int64_t maxit, uint32_t *last_step, double *last_value)
  {
    GF_REQUIRE(c.comm == nullptr && c.direct.factored, GF_ERR_INVALID_ARG,
               "cg_solve_direct_precond: serial handle with a factorisation");
    cudaStream_t s = c.stream;
    CGScalars    init{};
    init.tol     = tol;
    init.maxit   = int(std::min<int64_t>(maxit, 2147483647));
    init.status  = 0;
    *c.h_scalars = init;
    GF_CUDA_CHECK(cudaMemcpyAsync(c.cg_scalars.p, c.h_scalars, sizeof(CGScalars),
                                  cudaMemcpyHostToDevice, s));
    GF_CUDA_CHECK(cudaMemsetAsync(c.cg_p.p, 0, c.n_local * sizeof(double), s));
    const int dg = vec_grid(c, c.n_owned);
    auto update = [&](bool startup) {
      ProfScope ps(c, Profile::CG_VEC, 2);
      launch_update(c, startup, b, x);
      reduce_sums(c, 1, startup ? 3 : 4, true); // r.r -> residual check
    };
    auto poll = [&]() {
      GF_CUDA_CHECK(cudaMemcpyAsync(c.h_scalars, c.cg_scalars.p, sizeof(CGScalars),
                                    cudaMemcpyDeviceToHost, s));
      GF_CUDA_CHECK(cudaStreamSynchronize(s));
      return c.h_scalars->status;
    };
    vec_zero(c, x);
    vec_zero(c, c.cg_v.p);
    update(true); // r = b
    GF_CUDA_CHECK(cudaGetLastError());
    while (poll() == 0)
      {
        direct_solve(c, c.cg_r.p, c.cg_z.p); // z = A^-1 r
        {
          ProfScope ps(c, Profile::CG_VEC, 3);
          launch_dot_chunks(c, c.cg_r.p, c.cg_z.p, true);
          reduce_sums(c, 1, 5, true); // r.z, beta
          cg_direction_kernel<<<dg, VEC_THREADS, 0, s>>>(c.n_owned, c.cg_scalars.p, c.cg_z.p,
                                                         c.cg_p.p);
        }
        op_apply(c, val, c.cg_p.p, c.cg_v.p, nullptr);
        {
          ProfScope ps(c, Profile::CG_VEC, 2);
          launch_dot_chunks(c, c.cg_p.p, c.cg_v.p, true);
          reduce_sums(c, 1, 1, true); // p.Ap, alpha
        }
        update(false);
        GF_CUDA_CHECK(cudaGetLastError());
      }
    *last_step  = uint32_t(c.h_scalars->it);
    *last_value = c.h_scalars->res;
    return c.h_scalars->status == 1 ? GF_OK : GF_ERR_NOT_CONVERGED;
  }

  int cg_solve(gf_context &c, const double *val, double *x, const double *b, double tol,
               bool tol_relative_to_rhs, int64_t maxit, uint32_t *last_step, double *last_value)
  {
    cudaStream_t s = c.stream;
    double       bnorm = -1.0; // >= 0: the zero vector may replace a worse initial guess
    if (tol_relative_to_rhs)
      {
        const double bn = vec_masked_norm(c, b, false);
        tol *= bn; // tol_lin * system_rhs.l2_norm() (:1171-1172)
        if (c.cg_initial_guess == 1)
          bnorm = bn;
      }
    if (c.precond == GF_PRECOND_MULTIGRID)
      return cg_solve_mg(c, val, x, b, tol, maxit, bnorm, last_step, last_value);
    CGScalars init{};
    init.tol    = tol;
    init.maxit  = int(std::min<int64_t>(maxit, 2147483647));
    init.status = 0;
    *c.h_scalars = init;
    GF_CUDA_CHECK(cudaMemcpyAsync(c.cg_scalars.p, c.h_scalars, sizeof(CGScalars),
                                  cudaMemcpyHostToDevice, s));
    GF_CUDA_CHECK(cudaMemsetAsync(c.cg_p.p, 0, c.n_local * sizeof(double), s));
    const int dg = vec_grid(c, c.n_owned);

    // ---- startup ----
    if (c.comm)
      halo_exchange(c, x);
    op_apply(c, val, x, c.cg_v.p, nullptr);
    {
      ProfScope ps(c, Profile::CG_VEC, 2);
      launch_update(c, true, b, x);
      reduce_sums(c, 2, 0, true);
    }
    GF_CUDA_CHECK(cudaGetLastError());

    int64_t enqueued = 0;
    while (true)
      {
        GF_CUDA_CHECK(cudaMemcpyAsync(c.h_scalars, c.cg_scalars.p, sizeof(CGScalars),
                                      cudaMemcpyDeviceToHost, s));
        GF_CUDA_CHECK(cudaStreamSynchronize(s));
        comm_check(c);
        if (c.h_scalars->status != 0)
          break;
        if (enqueued == 0 && bnorm >= 0.0 && c.h_scalars->res > bnorm)
          {
            // the zero vector is the better initial guess (see cg_solve_mg)
            bnorm = -1.0;
            vec_zero(c, x);
            vec_zero(c, c.cg_v.p);
            ProfScope ps(c, Profile::CG_VEC, 2);
            launch_update(c, true, b, x);
            reduce_sums(c, 2, 0, true);
            continue;
          }
        GF_REQUIRE(enqueued <= int64_t(c.h_scalars->it) + c.cg_check_every, GF_ERR_CUDA,
                   "CG device state did not advance");
        for (int k = 0; k < c.cg_check_every; ++k, ++enqueued)
          {
            {
              ProfScope ps(c, Profile::CG_VEC);
              cg_direction_kernel<<<dg, VEC_THREADS, 0, s>>>(c.n_owned, c.cg_scalars.p, c.cg_z.p,
                                                             c.cg_p.p);
            }
            if (c.comm)
              halo_exchange(c, c.cg_p.p);
            op_apply(c, val, c.cg_p.p, c.cg_v.p, nullptr);
            {
              ProfScope ps(c, Profile::CG_VEC, 4);
              launch_dot_chunks(c, c.cg_p.p, c.cg_v.p, true);
              reduce_sums(c, 1, 1, true);
              launch_update(c, false, b, x);
              reduce_sums(c, 2, 2, true);
            }
          }
        GF_CUDA_CHECK(cudaGetLastError());
      }
    *last_step  = uint32_t(c.h_scalars->it);
    *last_value = c.h_scalars->res;
    return c.h_scalars->status == 1 ? GF_OK : GF_ERR_NOT_CONVERGED;
  }
} // namespace gf
