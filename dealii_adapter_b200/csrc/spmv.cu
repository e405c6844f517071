// K4: FP64 block-row SpMV y = A x with an optional fused dot product (x . A x) — the `vmult`
// inside SolverCG (nonlinear_elasticity.cc:1184, linear_elasticity.cc:551) and the explicit
// vmults of assemble_rhs (linear_elasticity.cc:411-419).
//
// Storage (gf_context.h): per node row A with nb blocks, `dim` scalar rows of nb*dim doubles
// (padded to an even count), all sharing ONE column-index stream: 8 B per value + 4/dim^2 B of
// index per value instead of CSR's 8+4. HBM-bound; algorithmic bytes per launch =
// 8*n_val + 4*n_blocks + 4*(n_rows+1) + 8*(n_rows+1) + 8*n_local(x) + 8*n_owned(y).
// One warp per node row; every lane streams 16-byte (double2) pieces of each scalar row with
// streaming loads and gathers x through the read-only path; rows are handed out by a
// persistent grid so the fused dot product needs only gridDim.x partial sums (fixed order).
#include "gf_context.h"
#include "kernel_utils.cuh"

namespace gf
{
  namespace
  {
    constexpr int SPMV_THREADS = 256;

    __device__ __forceinline__ double2 ld_stream2(const double *p)
    {
      return __ldcs(reinterpret_cast<const double2 *>(p));
    }

    template <int DIM, bool DOT>
    __global__ void __launch_bounds__(SPMV_THREADS)
      spmv_kernel(const int64_t n_rows, const int32_t *__restrict__ brow_ptr,
                  const int64_t *__restrict__ val_ptr, const int32_t *__restrict__ bcol,
                  const double *__restrict__ val, const double *__restrict__ x,
                  double *__restrict__ y, double *__restrict__ partials, const int *status)
    {
      if (status != nullptr && *status != 0)
        return;
      __shared__ double red[32];
      const int         lane = threadIdx.x & 31;
      const int64_t     warp_global = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
      const int64_t     n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
      double            dot = 0.0;
      for (int64_t A = warp_global; A < n_rows; A += n_warps)
        {
          const int32_t b0     = brow_ptr[A];
          const int     ne     = (brow_ptr[A + 1] - b0) * DIM;
          const int64_t vbase  = val_ptr[A];
          const int     stride = int((val_ptr[A + 1] - vbase) / DIM);
          const double *vrow   = val + vbase;
          const int32_t *crow  = bcol + b0;
          double        acc[DIM];
#pragma unroll
          for (int r = 0; r < DIM; ++r)
            acc[r] = 0.0;
#pragma unroll 2
          for (int e = 2 * lane; e < ne; e += 64)
            {
              const int    e1 = e + 1;
              const int    k0 = e / DIM, k1 = e1 / DIM;
              const double x0 = __ldg(x + int64_t(__ldcs(crow + k0)) * DIM + (e - k0 * DIM));
              double       x1 = 0.0;
              if (e1 < ne)
                x1 = __ldg(x + int64_t(__ldcs(crow + k1)) * DIM + (e1 - k1 * DIM));
#pragma unroll
              for (int r = 0; r < DIM; ++r)
                {
                  const double2 v = ld_stream2(vrow + r * stride + e);
                  acc[r]          = fma(v.x, x0, acc[r]);
                  acc[r]          = fma(v.y, x1, acc[r]);
                }
            }
#pragma unroll
          for (int r = 0; r < DIM; ++r)
            acc[r] = warp_sum(acc[r]);
          if (lane < DIM)
            {
              double yr = acc[0];
#pragma unroll
              for (int r = 1; r < DIM; ++r)
                if (lane == r)
                  yr = acc[r];
              y[A * DIM + lane] = yr;
              if (DOT)
                dot = fma(yr, x[A * DIM + lane], dot);
            }
        }
      if (DOT)
        {
          dot = warp_sum(dot);
          if (lane == 0)
            red[threadIdx.x >> 5] = dot;
          __syncthreads();
          if (threadIdx.x < 32)
            {
              double v = threadIdx.x < (SPMV_THREADS >> 5) ? red[threadIdx.x] : 0.0;
              v        = warp_sum(v);
              if (threadIdx.x == 0)
                partials[blockIdx.x] = v;
            }
        }
    }

    // y = M x with M = m_ab delta_cd (consistent mass, one scalar per block)
    template <int DIM>
    __global__ void spmv_mass_kernel(const int64_t n_rows, const int32_t *__restrict__ brow_ptr,
                                     const int32_t *__restrict__ bcol,
                                     const double *__restrict__ mass_blk,
                                     const double *__restrict__ x, double *__restrict__ y)
    {
      const int     lane = threadIdx.x & 31;
      const int64_t A    = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
      if (A >= n_rows)
        return;
      const int32_t b0 = brow_ptr[A], b1 = brow_ptr[A + 1];
      double        acc[DIM];
#pragma unroll
      for (int r = 0; r < DIM; ++r)
        acc[r] = 0.0;
      for (int32_t k = b0 + lane; k < b1; k += 32)
        {
          const double  m = mass_blk[k];
          const int64_t B = bcol[k];
#pragma unroll
          for (int r = 0; r < DIM; ++r)
            acc[r] = fma(m, x[B * DIM + r], acc[r]);
        }
#pragma unroll
      for (int r = 0; r < DIM; ++r)
        acc[r] = warp_sum(acc[r]);
      if (lane == 0)
#pragma unroll
        for (int r = 0; r < DIM; ++r)
          y[A * DIM + r] = acc[r];
    }
  } // namespace

  void launch_spmv(gf_context &c, const double *val, const double *x, double *y,
                   double *dot_partials)
  {
    ProfScope     ps(c, Profile::SPMV);
    const int64_t n_rows = c.n_owned_nodes;
    if (n_rows == 0)
      return;
    const int64_t want = (n_rows * 32 + SPMV_THREADS - 1) / SPMV_THREADS;
    const int     grid = int(std::min<int64_t>(want, c.max_red_blocks));
    const int *   st   = dot_partials ? &c.cg_scalars.p->status : nullptr;
    if (c.dim == 3)
      {
        if (dot_partials)
          spmv_kernel<3, true><<<grid, SPMV_THREADS, 0, c.stream>>>(
            n_rows, c.brow_ptr.p, c.val_ptr.p, c.bcol.p, val, x, y, dot_partials, st);
        else
          spmv_kernel<3, false><<<grid, SPMV_THREADS, 0, c.stream>>>(
            n_rows, c.brow_ptr.p, c.val_ptr.p, c.bcol.p, val, x, y, nullptr, nullptr);
      }
    else
      {
        if (dot_partials)
          spmv_kernel<2, true><<<grid, SPMV_THREADS, 0, c.stream>>>(
            n_rows, c.brow_ptr.p, c.val_ptr.p, c.bcol.p, val, x, y, dot_partials, st);
        else
          spmv_kernel<2, false><<<grid, SPMV_THREADS, 0, c.stream>>>(
            n_rows, c.brow_ptr.p, c.val_ptr.p, c.bcol.p, val, x, y, nullptr, nullptr);
      }
    GF_CUDA_CHECK(cudaGetLastError());
  }

  void launch_spmv_mass(gf_context &c, const double *x, double *y)
  {
    ProfScope     ps(c, Profile::SPMV);
    const int64_t n_rows = c.n_owned_nodes;
    if (n_rows == 0)
      return;
    const unsigned grid = unsigned((n_rows * 32 + 255) / 256);
    if (c.dim == 3)
      spmv_mass_kernel<3><<<grid, 256, 0, c.stream>>>(n_rows, c.brow_ptr.p, c.bcol.p,
                                                      c.mass_blk.p, x, y);
    else
      spmv_mass_kernel<2><<<grid, 256, 0, c.stream>>>(n_rows, c.brow_ptr.p, c.bcol.p,
                                                      c.mass_blk.p, x, y);
    GF_CUDA_CHECK(cudaGetLastError());
  }

  // bytes one SpMV launch must move in the stored format (the roofline numerator)
  double spmv_bytes(const gf_context &c)
  {
    return 8.0 * double(c.n_val) + 4.0 * double(c.n_blocks) + 12.0 * double(c.n_owned_nodes + 1) +
           8.0 * double(c.n_local) + 8.0 * double(c.n_owned);
  }
} // namespace gf
