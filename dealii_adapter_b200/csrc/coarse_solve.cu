// Coarsest multigrid level in ONE launch: the whole Chebyshev iteration (degree ~80) of the
// block-Jacobi-preconditioned operator runs inside a persistent cooperative kernel.
//
// The coarsest level of the bench hierarchy has ~5 k DoFs (3x18x3 Q2 cells): every one of its 80
// smoother steps used to be two launches (SpMV + vector update) of a few microseconds each, i.e.
// pure launch latency (~1 ms per V-cycle, ~10 % of a Newton solve). Here every CTA owns a
// contiguous range of block rows, stages ITS slice of the matrix (values + column indices) in
// shared memory once, keeps its residual / solution rows in shared memory across all steps and
// exchanges only the direction vector d through L2, with one grid-wide barrier per step
// (monotonic 64-bit arrival counter, acquire/release at gpu scope). Same recurrence as
// multigrid.cu::smooth (Saad, alg. 12.1) with a zero initial guess:
//   r = b ; d = (1/theta) D^-1 r ; x = d
//   k >= 1:  r -= A d ; d = rho_k rho_{k-1} d + (2 rho_k / delta) D^-1 r ; x += d
// Fixed summation order per row (lane <-> pair of entries, shuffle tree): deterministic.
// Used when the level is the coarsest, is not partitioned (serial or replicated) and small.
#include <algorithm>

#include "gf_context.h"
#include "kernel_utils.cuh"

namespace gf
{
  namespace
  {
    constexpr int CS_THREADS = 256;

    __device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
    {
      unsigned long long v;
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
      return v;
    }

    // all CTAs of the (cooperative) grid arrive; `target` = arrivals expected so far
    __device__ __forceinline__ void grid_barrier(unsigned long long *counter,
                                                 const unsigned long long target)
    {
      __syncthreads();
      if (threadIdx.x == 0)
        {
          __threadfence();
          atomicAdd(counter, 1ull);
          while (ld_acquire_gpu(counter) < target)
            ;
          __threadfence();
        }
      __syncthreads();
    }

    template <int DIM>
    __global__ void __launch_bounds__(CS_THREADS, 1)
      coarse_cheb_kernel(const int n_rows, const int rows_per_cta,
                         const int32_t *__restrict__ brow_ptr, const int64_t *__restrict__ val_ptr,
                         const int32_t *__restrict__ bcol, const double *__restrict__ val,
                         const double *__restrict__ dinv, const double *__restrict__ b,
                         double *__restrict__ x_out, double *d0, double *d1, const int degree,
                         const double *__restrict__ lmax, const double ratio,
                         unsigned long long *counter, const unsigned long long counter_base,
                         const int stage_in_smem)
    {
      extern __shared__ __align__(16) unsigned char cs_smem[];
      const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = CS_THREADS / 32;
      // Chebyshev interval [lmax/ratio, lmax] from the device-resident eigenvalue estimate
      const double bb = *lmax, aa = bb / ratio;
      const double theta = 0.5 * (bb + aa), delta = 0.5 * (bb - aa), sigma = theta / delta;
      const int row_begin = blockIdx.x * rows_per_cta;
      const int row_end   = min(n_rows, row_begin + rows_per_cta);
      const int n_my      = max(0, row_end - row_begin);
      // shared layout: r[n_my*DIM] | x[n_my*DIM] | values slice | column slice
      double *sr = reinterpret_cast<double *>(cs_smem);
      double *sx = sr + rows_per_cta * DIM;
      double *sval = sx + rows_per_cta * DIM;
      const int64_t v0 = n_my > 0 ? val_ptr[row_begin] : 0, v1 = n_my > 0 ? val_ptr[row_end] : 0;
      const int32_t c0 = n_my > 0 ? brow_ptr[row_begin] : 0, c1 = n_my > 0 ? brow_ptr[row_end] : 0;
      int32_t *scol = reinterpret_cast<int32_t *>(sval + (stage_in_smem ? (v1 - v0) : 0));
      if (stage_in_smem)
        {
          for (int64_t i = tid; i < v1 - v0; i += CS_THREADS)
            sval[i] = val[v0 + i];
          for (int i = tid; i < c1 - c0; i += CS_THREADS)
            scol[i] = bcol[c0 + i];
        }
      const double * vbase_p = stage_in_smem ? sval - v0 : val;   // index with global offsets
      const int32_t *cbase_p = stage_in_smem ? scol - c0 : bcol;
      // ---- step 0: r = b ; d = D^-1 r / theta ; x = d -------------------------------------------
      for (int i = tid; i < n_my * DIM; i += CS_THREADS)
        sr[i] = b[int64_t(row_begin) * DIM + i];
      __syncthreads();
      for (int i = tid; i < n_my * DIM; i += CS_THREADS)
        {
          const int     lr = i / DIM, comp = i - lr * DIM;
          const int64_t A  = row_begin + lr;
          double        z  = 0.0;
#pragma unroll
          for (int j = 0; j < DIM; ++j)
            z = fma(dinv[A * DIM * DIM + comp * DIM + j], sr[lr * DIM + j], z);
          const double dn = (1.0 / theta) * z;
          d0[A * DIM + comp] = dn;
          sx[i]              = dn;
        }
      unsigned long long arrivals = counter_base;
      double             rho      = 1.0 / sigma;
      double *           d_old = d0, *d_new = d1;
      for (int k = 1; k < degree; ++k)
        {
          arrivals += gridDim.x;
          grid_barrier(counter, arrivals); // d_old complete on every CTA
          const double rho_new = 1.0 / (2.0 * sigma - rho);
          const double c_d = rho_new * rho, c_r = 2.0 * rho_new / delta;
          rho = rho_new;
          for (int lr = warp; lr < n_my; lr += n_warps)
            {
              const int64_t  A      = row_begin + lr;
              const int32_t  b0     = brow_ptr[A];
              const int      ne     = (brow_ptr[A + 1] - b0) * DIM;
              const int64_t  vb     = val_ptr[A];
              const int      stride = int((val_ptr[A + 1] - vb) / DIM);
              const double * vrow   = vbase_p + vb;
              const int32_t *crow   = cbase_p + b0;
              double         acc[DIM];
#pragma unroll
              for (int r = 0; r < DIM; ++r)
                acc[r] = 0.0;
              for (int e = 2 * lane; e < ne; e += 64)
                {
                  const int    e1 = e + 1;
                  const int    k0 = e / DIM, k1 = e1 / DIM;
                  // d_old was written by other CTAs in the previous step: read through L2
                  const double x0 = __ldcg(d_old + int64_t(crow[k0]) * DIM + (e - k0 * DIM));
                  double       x1 = 0.0;
                  if (e1 < ne)
                    x1 = __ldcg(d_old + int64_t(crow[k1]) * DIM + (e1 - k1 * DIM));
#pragma unroll
                  for (int r = 0; r < DIM; ++r)
                    {
                      const double2 v = *reinterpret_cast<const double2 *>(vrow + r * stride + e);
                      acc[r]          = fma(v.x, x0, acc[r]);
                      acc[r]          = fma(v.y, x1, acc[r]);
                    }
                }
#pragma unroll
              for (int r = 0; r < DIM; ++r)
                acc[r] = warp_sum(acc[r]);
              if (lane == 0)
                {
                  double rl[DIM];
#pragma unroll
                  for (int i = 0; i < DIM; ++i)
                    {
                      rl[i]             = sr[lr * DIM + i] - acc[i];
                      sr[lr * DIM + i] = rl[i];
                    }
#pragma unroll
                  for (int i = 0; i < DIM; ++i)
                    {
                      double z = 0.0;
#pragma unroll
                      for (int j = 0; j < DIM; ++j)
                        z = fma(dinv[A * DIM * DIM + i * DIM + j], rl[j], z);
                      const double dn = fma(c_d, __ldcg(d_old + A * DIM + i), c_r * z);
                      d_new[A * DIM + i] = dn;
                      sx[lr * DIM + i] += dn;
                    }
                }
            }
          double *t = d_old;
          d_old     = d_new;
          d_new     = t;
        }
      __syncthreads();
      for (int i = tid; i < n_my * DIM; i += CS_THREADS)
        x_out[int64_t(row_begin) * DIM + i] = sx[i];
    }
  } // namespace

  // decides once per level whether the single-launch solver applies and how it is launched
  static void coarse_solver_plan(gf_context &c)
  {
    c.cs_planned = true;
    c.cs_enabled = false;
    const char *env = getenv("GF_COARSE_SOLVER");
    if (env && atoi(env) == 0)
      return;
    if (c.comm != nullptr || c.n_owned_nodes <= 0 || c.n_owned_nodes > 40000)
      return;
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c.device);
    if (!coop)
      return;
    const int n_rows = int(c.n_owned_nodes);
    c.cs_rows_per_cta = (n_rows + c.sm_count - 1) / c.sm_count;
    c.cs_grid         = (n_rows + c.cs_rows_per_cta - 1) / c.cs_rows_per_cta;
    // largest matrix slice of a CTA
    std::vector<int32_t> brow(n_rows + 1);
    std::vector<int64_t> vptr(n_rows + 1);
    GF_CUDA_CHECK(cudaMemcpy(brow.data(), c.brow_ptr.p, (n_rows + 1) * sizeof(int32_t),
                             cudaMemcpyDeviceToHost));
    GF_CUDA_CHECK(cudaMemcpy(vptr.data(), c.val_ptr.p, (n_rows + 1) * sizeof(int64_t),
                             cudaMemcpyDeviceToHost));
    size_t slice = 0;
    for (int r0 = 0; r0 < n_rows; r0 += c.cs_rows_per_cta)
      {
        const int r1 = std::min(n_rows, r0 + c.cs_rows_per_cta);
        slice = std::max(slice, size_t(vptr[r1] - vptr[r0]) * 8 + size_t(brow[r1] - brow[r0]) * 4);
      }
    const size_t state = size_t(c.cs_rows_per_cta) * c.dim * 2 * sizeof(double);
    c.cs_stage = state + slice + 64 <= size_t(200) * 1024;
    c.cs_smem  = state + (c.cs_stage ? slice : 0) + 64;
    if (c.cs_smem > size_t(200) * 1024)
      return;
    const void *fn = c.dim == 3 ? (const void *)coarse_cheb_kernel<3> : (const void *)coarse_cheb_kernel<2>;
    GF_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(c.cs_smem)));
    int per_sm = 0;
    GF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, CS_THREADS, c.cs_smem));
    if (per_sm < 1 || c.cs_grid > per_sm * c.sm_count)
      return;
    c.cs_counter.alloc_zero(1, c.stream);
    c.cs_arrivals = 0;
    c.cs_enabled  = true;
  }

  // x = Chebyshev(degree) applied to b with a zero initial guess on this (coarsest) level;
  // returns false if the single-launch solver does not apply (caller falls back to smooth())
  bool coarse_solve_single_launch(gf_context &c, const double *val, const double *b, double *x,
                                  int degree, double ratio)
  {
    if (!c.cs_planned)
      coarse_solver_plan(c);
    if (!c.cs_enabled || c.operator_kind != 0)
      return false;
    ProfScope    ps(c, Profile::MG_SPMV);
    const double *lmax = c.mg_lmax_dev.p;
    int          n_rows = int(c.n_owned_nodes), rows_per_cta = c.cs_rows_per_cta,
        stage = c.cs_stage ? 1 : 0;
    const int32_t *brow = c.brow_ptr.p, *bcol = c.bcol.p;
    const int64_t *vptr = c.val_ptr.p;
    const double * dinv = c.dinv.p;
    double *       d0 = c.mg_d.p, *d1 = c.mg_v.p;
    unsigned long long *counter = c.cs_counter.p, base = c.cs_arrivals;
    void *args[] = {&n_rows, &rows_per_cta, &brow, &vptr, &bcol, &val, &dinv, &b, &x, &d0, &d1,
                    &degree, &lmax, &ratio, &counter, &base, &stage};
    const void *fn = c.dim == 3 ? (const void *)coarse_cheb_kernel<3> : (const void *)coarse_cheb_kernel<2>;
    GF_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(c.cs_grid), dim3(CS_THREADS), args,
                                              c.cs_smem, c.stream));
    c.cs_arrivals += (unsigned long long)(std::max(0, degree - 1)) * c.cs_grid;
    return true;
  }
} // namespace gf
