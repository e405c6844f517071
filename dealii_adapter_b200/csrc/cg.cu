// K5/K6: preconditioned conjugate gradients entirely on the device — deal.II's SolverCG
// recurrences (reference call sites nonlinear_elasticity.cc:1174-1190, linear_elasticity.cc:
// 544-554; algorithm: deal.II 9.5 SolverCG::solve / IterationWorker::do_iteration) with the
// serial SSOR sweeps (PreconditionSelector("ssor",.65) / PreconditionSSOR(1.2)) replaced by a
// node-block Jacobi preconditioner, as BASELINE north_star prescribes.
//
//   startup   r = b - A x ; z = M^-1 r ; res = ||r|| ; check(0,res)
//   iteration p = beta p + z ; v = A p (+ p.v fused in the SpMV) ; alpha = r.z / p.v ;
//             x += alpha p ; r -= alpha v ; z = M^-1 r ; r.r, r.z in the same pass ; check(it,res)
// Scalars (alpha, beta, residual, iteration counter, SolverControl state) live in device memory;
// every kernel returns immediately once the state leaves `iterate`, so the host only polls the
// state every `cg_check_every` iterations — no per-iteration host round trip. Reductions are
// two-stage with a fixed summation order (bitwise reproducible; no FP atomics). Multi-GPU: the
// local sums are all-reduced over NCCL before the scalar step.
#include "gf_context.h"
#include "kernel_utils.cuh"

namespace gf
{
  namespace
  {
    constexpr int VEC_THREADS = 256;

    // sums[k] = sum_j partials[k*stride + j], fixed order; single block
    __global__ void reduce_partials_kernel(const double *__restrict__ partials, const int n,
                                           const int stride, const int n_sums,
                                           double *__restrict__ sums, const int *status)
    {
      if (status != nullptr && *status != 0)
        return;
      __shared__ double sm[32];
      for (int k = 0; k < n_sums; ++k)
        {
          double v = 0;
          for (int j = threadIdx.x; j < n; j += blockDim.x)
            v += partials[k * stride + j];
          v = warp_sum(v);
          __syncthreads();
          if ((threadIdx.x & 31) == 0)
            sm[threadIdx.x >> 5] = v;
          __syncthreads();
          if (threadIdx.x < 32)
            {
              double w = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
              w        = warp_sum(w);
              if (threadIdx.x == 0)
                sums[k] = w;
            }
        }
    }

    // phase 0: after startup (sums = rr, rz) ; phase 1: after SpMV (sums = pAp) ;
    // phase 2: after the update (sums = rr, rz)
    __global__ void cg_scalar_kernel(CGScalars *s, const double *sums, const int phase)
    {
      if (s->status != 0)
        return;
      if (phase == 1)
        {
          s->pAp   = sums[0];
          s->alpha = s->rz / s->pAp;
          return;
        }
      if (phase == 3 || phase == 4) // multigrid CG: residual check BEFORE the preconditioner
        {
          const double rr = sums[0];
          s->it           = phase == 3 ? 0 : s->it + 1;
          if (phase == 3)
            {
              s->res0 = sqrt(fabs(rr));
              s->beta = 0.0;
              s->rz   = 0.0;
            }
          s->rr  = rr;
          s->res = sqrt(fabs(rr));
          if (s->res <= s->tol)
            s->status = 1;
          else if (s->it >= s->maxit || s->res != s->res)
            s->status = 2;
          return;
        }
      if (phase == 5) // multigrid CG: r.z after the V-cycle
        {
          const double rz = sums[0];
          s->beta         = s->rz != 0.0 ? rz / s->rz : 0.0;
          s->rz           = rz;
          return;
        }
      const double rr = sums[0], rz = sums[1];
      if (phase == 0)
        {
          s->it   = 0;
          s->res0 = sqrt(fabs(rr));
          s->beta = 0.0;
        }
      else
        {
          s->it += 1;
          s->beta = rz / s->rz; // r_dot_preconditioner_dot_r / previous
        }
      s->rr  = rr;
      s->rz  = rz;
      s->res = sqrt(fabs(rr));
      // SolverControl::check
      if (s->res <= s->tol)
        s->status = 1;
      else if (s->it >= s->maxit || s->res != s->res)
        s->status = 2;
    }

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
    template <int DIM, bool STARTUP>
    __global__ void __launch_bounds__(VEC_THREADS)
      cg_update_kernel(const int64_t n_nodes, const CGScalars *s, const double *__restrict__ b,
                       const double *__restrict__ p, const double *__restrict__ v,
                       const double *__restrict__ dinv, double *__restrict__ x,
                       double *__restrict__ r, double *__restrict__ z,
                       double *__restrict__ partials, const int pstride)
    {
      if (s->status != 0)
        return;
      __shared__ double sm[64];
      const double      alpha = STARTUP ? 0.0 : s->alpha;
      double            acc[2] = {0.0, 0.0};
      for (int64_t A = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; A < n_nodes;
           A += int64_t(gridDim.x) * blockDim.x)
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

    // partial sums of a . b (owned entries)
    __global__ void __launch_bounds__(VEC_THREADS)
      dot_kernel(const int64_t n, const double *__restrict__ a, const double *__restrict__ b,
                 double *__restrict__ partials)
    {
      __shared__ double sm[32];
      double            acc[1] = {0.0};
      for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n;
           i += int64_t(gridDim.x) * blockDim.x)
        acc[0] = fma(a[i], b[i], acc[0]);
      block_sum<1>(acc, sm);
      if (threadIdx.x == 0)
        partials[blockIdx.x] = acc[0];
    }

    int vec_grid(const gf_context &c, int64_t n)
    {
      const int64_t want = (n + VEC_THREADS - 1) / VEC_THREADS;
      return int(std::max<int64_t>(1, std::min<int64_t>(want, c.max_red_blocks)));
    }

    void reduce_and_scalar(gf_context &c, int n_partials, int n_sums, int phase)
    {
      const int *st   = &c.cg_scalars.p->status;
      double *   sums = c.partials.p + 3 * size_t(c.max_red_blocks);
      reduce_partials_kernel<<<1, 1024, 0, c.stream>>>(c.partials.p, n_partials, c.max_red_blocks,
                                                       n_sums, sums, st);
      if (c.comm)
        allreduce_sum(c, sums, n_sums);
      cg_scalar_kernel<<<1, 1, 0, c.stream>>>(c.cg_scalars.p, sums, phase);
    }
  } // namespace

  // The same recurrences with z = V-cycle(r). One CG iteration is now several milliseconds of
  // device work, so the host reads the SolverControl state after every residual update (before
  // the V-cycle is enqueued) instead of every cg_check_every iterations.
  static int cg_solve_mg(gf_context &c, const double *val, double *x, const double *b, double tol,
                         int64_t maxit, uint32_t *last_step, double *last_value)
  {
    GF_REQUIRE(mg_active(c), GF_ERR_INVALID_ARG,
               "GF_PRECOND_MULTIGRID selected but no coarse level is attached (gf_mg_attach)");
    for (gf_context *l = &c; l != nullptr; l = l->mg.coarse)
      GF_REQUIRE(l->mg_lmax > 0.0, GF_ERR_INVALID_ARG,
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
    const int64_t n_nodes = c.n_owned_nodes;
    const int     ug = vec_grid(c, n_nodes), dg = vec_grid(c, c.n_owned);
    const int     sg_rows = op_dot_partials(c);
    auto update = [&](bool startup) {
      ProfScope ps(c, Profile::CG_VEC, 3);
      if (c.dim == 3)
        {
          if (startup)
            cg_update_kernel<3, true><<<ug, VEC_THREADS, 0, s>>>(
              n_nodes, c.cg_scalars.p, b, nullptr, c.cg_v.p, c.dinv.p, x, c.cg_r.p, c.cg_z.p,
              c.partials.p, c.max_red_blocks);
          else
            cg_update_kernel<3, false><<<ug, VEC_THREADS, 0, s>>>(
              n_nodes, c.cg_scalars.p, b, c.cg_p.p, c.cg_v.p, c.dinv.p, x, c.cg_r.p, c.cg_z.p,
              c.partials.p, c.max_red_blocks);
        }
      else
        {
          if (startup)
            cg_update_kernel<2, true><<<ug, VEC_THREADS, 0, s>>>(
              n_nodes, c.cg_scalars.p, b, nullptr, c.cg_v.p, c.dinv.p, x, c.cg_r.p, c.cg_z.p,
              c.partials.p, c.max_red_blocks);
          else
            cg_update_kernel<2, false><<<ug, VEC_THREADS, 0, s>>>(
              n_nodes, c.cg_scalars.p, b, c.cg_p.p, c.cg_v.p, c.dinv.p, x, c.cg_r.p, c.cg_z.p,
              c.partials.p, c.max_red_blocks);
        }
      reduce_and_scalar(c, ug, 1, startup ? 3 : 4); // r.r -> residual check
    };
    auto poll = [&]() {
      GF_CUDA_CHECK(cudaMemcpyAsync(c.h_scalars, c.cg_scalars.p, sizeof(CGScalars),
                                    cudaMemcpyDeviceToHost, s));
      GF_CUDA_CHECK(cudaStreamSynchronize(s));
      comm_check(c);
      return c.h_scalars->status;
    };
    // ---- startup: r = b - A x ; check ----
    if (c.comm)
      halo_exchange(c, x);
    op_apply(c, val, x, c.cg_v.p, nullptr);
    update(true);
    GF_CUDA_CHECK(cudaGetLastError());
    while (poll() == 0)
      {
        mg_vcycle(c, c.cg_r.p, c.cg_z.p); // z = M^-1 r
        {
          ProfScope ps(c, Profile::CG_VEC, 4);
          dot_kernel<<<dg, VEC_THREADS, 0, s>>>(c.n_owned, c.cg_r.p, c.cg_z.p, c.partials.p);
          reduce_and_scalar(c, dg, 1, 5); // r.z, beta
          cg_direction_kernel<<<dg, VEC_THREADS, 0, s>>>(c.n_owned, c.cg_scalars.p, c.cg_z.p,
                                                         c.cg_p.p);
        }
        if (c.comm)
          halo_exchange(c, c.cg_p.p);
        op_apply(c, val, c.cg_p.p, c.cg_v.p, c.partials.p);
        {
          ProfScope ps(c, Profile::CG_VEC, 2);
          reduce_and_scalar(c, sg_rows, 1, 1);
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
    if (tol_relative_to_rhs)
      tol *= vec_masked_norm(c, b, false); // tol_lin * system_rhs.l2_norm() (:1171-1172)
    if (c.precond == GF_PRECOND_MULTIGRID)
      return cg_solve_mg(c, val, x, b, tol, maxit, last_step, last_value);
    CGScalars init{};
    init.tol    = tol;
    init.maxit  = int(std::min<int64_t>(maxit, 2147483647));
    init.status = 0;
    *c.h_scalars = init;
    GF_CUDA_CHECK(cudaMemcpyAsync(c.cg_scalars.p, c.h_scalars, sizeof(CGScalars),
                                  cudaMemcpyHostToDevice, s));
    GF_CUDA_CHECK(cudaMemsetAsync(c.cg_p.p, 0, c.n_local * sizeof(double), s));
    const int64_t n_nodes = c.n_owned_nodes;
    const int     ug      = vec_grid(c, n_nodes);
    const int     dg      = vec_grid(c, c.n_owned);
    const int     sg_rows = op_dot_partials(c);

    // ---- startup ----
    if (c.comm)
      halo_exchange(c, x);
    op_apply(c, val, x, c.cg_v.p, nullptr);
    {
      ProfScope ps(c, Profile::CG_VEC, 3);
      if (c.dim == 3)
        cg_update_kernel<3, true><<<ug, VEC_THREADS, 0, s>>>(n_nodes, c.cg_scalars.p, b, nullptr,
                                                             c.cg_v.p, c.dinv.p, x, c.cg_r.p,
                                                             c.cg_z.p, c.partials.p,
                                                             c.max_red_blocks);
      else
        cg_update_kernel<2, true><<<ug, VEC_THREADS, 0, s>>>(n_nodes, c.cg_scalars.p, b, nullptr,
                                                             c.cg_v.p, c.dinv.p, x, c.cg_r.p,
                                                             c.cg_z.p, c.partials.p,
                                                             c.max_red_blocks);
      reduce_and_scalar(c, ug, 2, 0);
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
            op_apply(c, val, c.cg_p.p, c.cg_v.p, c.partials.p);
            {
              ProfScope ps(c, Profile::CG_VEC, 5);
              reduce_and_scalar(c, sg_rows, 1, 1);
              if (c.dim == 3)
                cg_update_kernel<3, false><<<ug, VEC_THREADS, 0, s>>>(
                  n_nodes, c.cg_scalars.p, b, c.cg_p.p, c.cg_v.p, c.dinv.p, x, c.cg_r.p, c.cg_z.p,
                  c.partials.p, c.max_red_blocks);
              else
                cg_update_kernel<2, false><<<ug, VEC_THREADS, 0, s>>>(
                  n_nodes, c.cg_scalars.p, b, c.cg_p.p, c.cg_v.p, c.dinv.p, x, c.cg_r.p, c.cg_z.p,
                  c.partials.p, c.max_red_blocks);
              reduce_and_scalar(c, ug, 2, 2);
            }
          }
        GF_CUDA_CHECK(cudaGetLastError());
      }
    *last_step  = uint32_t(c.h_scalars->it);
    *last_value = c.h_scalars->res;
    return c.h_scalars->status == 1 ? GF_OK : GF_ERR_NOT_CONVERGED;
  }
} // namespace gf
