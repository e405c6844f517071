// 'Solver type = Direct' (parameters.prm:43, the shipped default): the reference factorises the
// tangent / system matrix with SparseDirectUMFPACK and applies it once per solve
// (nonlinear_elasticity.cc:1192-1200, linear_elasticity.cc:556-563). Device replacement: the matrix
// is SPD, so it is factorised by a blocked BAND CHOLESKY (dpbtrf-style, right-looking) in a
// bandwidth-reducing ordering of the nodes (reverse Cuthill-McKee, rcm.h), followed by blocked
// forward / backward substitution.
//
// Storage: lower band, column-major: entry (i, j), 0 <= i - j <= W, at band[j * ld + (i - j)],
// ld = W + 1. W = w + DB_NB - 1 where w is the half bandwidth of the matrix, so that the panel
// under a DB_NB-wide diagonal block is a full RECTANGLE of m <= w rows (the entries beyond the
// band are zeros that stay zeros); the trailing update never leaves the band.
//
// Per block column (j0, nb):   potf2  A11 = L11 L11^T            one CTA
//                              trsm   P = A21 L11^-T             thread per panel row
//                              syrk   A22 -= P P^T               32 x 32 tiles, lower ones only
// Work n w^2 flops, FP64; launches 3 n / DB_NB. This is the small / medium problem path (2D, and
// 3D up to a few 10^5 DoFs: the band must fit the memory budget, else the tight CG stand-in runs).
//
// Emulation-ready (emu_compat.cuh): plain __syncthreads kernels and a driver templated on the
// launcher, so tests/test_cuda_emulation.py runs the SAME kernels and block loops on the CPU.
#pragma once
#include <cstdint>

#include "emu_compat.cuh"

namespace gf
{
  constexpr int DB_NB   = 32; // block columns
  constexpr int DB_TILE = 32; // syrk tile edge

  // band <- permuted lower triangle of the block-row matrix (format: gf_context.h); the band must
  // have been zeroed. node_new: old node -> new node; dof (node, c) -> node_new[node] * dim + c
  __global__ void db_fill_kernel(const int64_t n_nodes, const int dim,
                                 const int32_t *__restrict__ brow_ptr,
                                 const int64_t *__restrict__ val_ptr,
                                 const int32_t *__restrict__ bcol, const double *__restrict__ val,
                                 const int32_t *__restrict__ node_new, const int64_t ld,
                                 double *__restrict__ band)
  {
    for (int64_t A = blockIdx.x; A < n_nodes; A += gridDim.x)
      {
        const int32_t b0     = brow_ptr[A];
        const int     nb     = brow_ptr[A + 1] - b0;
        const int64_t vbase  = val_ptr[A];
        const int     stride = int((val_ptr[A + 1] - vbase) / dim);
        const int64_t ia     = int64_t(node_new[A]) * dim;
        for (int e = threadIdx.x; e < nb * dim * dim; e += blockDim.x)
          {
            const int     blk = e / (dim * dim), rc = e - blk * dim * dim;
            const int     r = rc / dim, cc = rc - r * dim;
            const int64_t i = ia + r, j = int64_t(node_new[bcol[b0 + blk]]) * dim + cc;
            if (i >= j)
              band[j * ld + (i - j)] = val[vbase + int64_t(r) * stride + blk * dim + cc];
          }
      }
  }

  // xp[new dof] = x[old dof] (to_new) or x[old dof] = xp[new dof]
  __global__ void db_permute_kernel(const int64_t n_nodes, const int dim,
                                    const int32_t *__restrict__ node_new, const bool to_new,
                                    double *__restrict__ x_old, double *__restrict__ x_new)
  {
    const int64_t n = n_nodes * dim;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n;
         i += int64_t(gridDim.x) * blockDim.x)
      {
        const int64_t node = i / dim, cc = i - node * dim;
        const int64_t k    = int64_t(node_new[node]) * dim + cc;
        if (to_new)
          x_new[k] = x_old[i];
        else
          x_old[i] = x_new[k];
      }
  }

  // A11 (nb x nb at (j0, j0)) = L11 L11^T in place; *info = 1-based column of a non-positive pivot
  __global__ void db_potf2_kernel(const int64_t j0, const int nb, const int64_t ld,
                                  double *__restrict__ band, int *info)
  {
    GF_DYN_SMEM(double, s); // [DB_NB][DB_NB + 1]
    constexpr int LS = DB_NB + 1;
    const int     tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < nb * nb; e += nt)
      {
        const int i = e / nb, j = e - i * nb;
        if (j <= i)
          s[i * LS + j] = band[(j0 + j) * ld + (i - j)];
      }
    __syncthreads();
    for (int c = 0; c < nb; ++c)
      {
        if (tid == 0)
          {
            double d = s[c * LS + c];
            if (!(d > 0.0))
              {
                if (*info == 0)
                  *info = int(j0 + c + 1);
                d = 1.0; // keep going with finite numbers; the caller discards the factor
              }
            s[c * LS + c] = sqrt(d);
          }
        __syncthreads();
        const double piv = s[c * LS + c];
        for (int i = c + 1 + tid; i < nb; i += nt)
          s[i * LS + c] /= piv;
        __syncthreads();
        const int r = nb - c - 1;
        for (int e = tid; e < r * r; e += nt)
          {
            const int i = c + 1 + e / r, j = c + 1 + e % r;
            if (j <= i)
              s[i * LS + j] -= s[i * LS + c] * s[j * LS + c];
          }
        __syncthreads();
      }
    for (int e = tid; e < nb * nb; e += nt)
      {
        const int i = e / nb, j = e - i * nb;
        if (j <= i)
          band[(j0 + j) * ld + (i - j)] = s[i * LS + j];
      }
  }

  // panel rows i = j0 + nb + t, t < m:  row := row L11^-T
  __global__ void db_trsm_kernel(const int64_t j0, const int nb, const int64_t m, const int64_t ld,
                                 double *__restrict__ band)
  {
    GF_DYN_SMEM(double, s); // L11 [DB_NB][DB_NB + 1]
    constexpr int LS = DB_NB + 1;
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x)
      {
        const int i = e / nb, j = e - i * nb;
        if (j <= i)
          s[i * LS + j] = band[(j0 + j) * ld + (i - j)];
      }
    __syncthreads();
    const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (t < m)
      {
        const int64_t i = j0 + nb + t;
        double        x[DB_NB];
        for (int c = 0; c < nb; ++c)
          {
            double v = band[(j0 + c) * ld + (i - j0 - c)];
            for (int k = 0; k < c; ++k)
              v -= x[k] * s[c * LS + k];
            x[c] = v / s[c * LS + c];
          }
        for (int c = 0; c < nb; ++c)
          band[(j0 + c) * ld + (i - j0 - c)] = x[c];
      }
  }

  // trailing window rows/cols r0 + [0, m), r0 = j0 + nb:  A22(i, i') -= sum_c P(i, c) P(i', c),
  // lower triangle; one CTA per pair of 32-row tiles (ti >= tj), 1-D grid over the pairs
  __global__ void db_syrk_kernel(const int64_t j0, const int nb, const int64_t m, const int64_t ld,
                                 double *__restrict__ band)
  {
    GF_DYN_SMEM(double, s); // Pi [DB_TILE][DB_NB + 1], Pj [DB_TILE][DB_NB + 1]
    constexpr int LS = DB_NB + 1;
    double *      Pi = s, *Pj = s + DB_TILE * LS;
    // pair index -> (ti, tj), tj <= ti
    const int64_t pair = blockIdx.x;
    int64_t       ti   = int64_t((sqrt(8.0 * double(pair) + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= pair)
      ++ti;
    while (ti * (ti + 1) / 2 > pair)
      --ti;
    const int64_t tj = pair - ti * (ti + 1) / 2;
    const int64_t r0 = j0 + nb;
    for (int e = threadIdx.x; e < DB_TILE * nb; e += blockDim.x)
      {
        const int     c = e / DB_TILE, l = e - c * DB_TILE; // consecutive threads: consecutive rows
        const int64_t a = ti * DB_TILE + l, b = tj * DB_TILE + l;
        Pi[l * LS + c]  = a < m ? band[(j0 + c) * ld + (r0 + a - j0 - c)] : 0.0;
        Pj[l * LS + c]  = b < m ? band[(j0 + c) * ld + (r0 + b - j0 - c)] : 0.0;
      }
    __syncthreads();
    for (int e = threadIdx.x; e < DB_TILE * DB_TILE; e += blockDim.x)
      {
        const int     lj = e / DB_TILE, li = e - lj * DB_TILE; // consecutive threads: consecutive i
        const int64_t a = ti * DB_TILE + li, b = tj * DB_TILE + lj;
        if (a < m && b <= a)
          {
            double v = 0.0;
            for (int c = 0; c < nb; ++c)
              v = fma(Pi[li * LS + c], Pj[lj * LS + c], v);
            band[(r0 + b) * ld + (a - b)] -= v;
          }
      }
  }

  // forward substitution, block (j0, nb): x_k = L11^-1 x_k (one CTA) ...
  __global__ void db_fwd_diag_kernel(const int64_t j0, const int nb, const int64_t ld,
                                     const double *__restrict__ band, double *__restrict__ x)
  {
    GF_DYN_SMEM(double, s); // x_k [DB_NB]
    const int tid = threadIdx.x;
    if (tid < nb)
      s[tid] = x[j0 + tid];
    __syncthreads();
    for (int c = 0; c < nb; ++c)
      {
        if (tid == 0)
          s[c] /= band[(j0 + c) * ld];
        __syncthreads();
        if (tid > c && tid < nb)
          s[tid] -= band[(j0 + c) * ld + (tid - c)] * s[c];
        __syncthreads();
      }
    if (tid < nb)
      x[j0 + tid] = s[tid];
  }
  // ... then x[panel rows] -= P x_k (thread per row)
  __global__ void db_fwd_update_kernel(const int64_t j0, const int nb, const int64_t m,
                                       const int64_t ld, const double *__restrict__ band,
                                       double *__restrict__ x)
  {
    const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (t >= m)
      return;
    const int64_t i = j0 + nb + t;
    double        v = x[i];
    for (int c = 0; c < nb; ++c)
      v -= band[(j0 + c) * ld + (i - j0 - c)] * x[j0 + c];
    x[i] = v;
  }

  // backward substitution, block (j0, nb): tmp[c] = sum_t P(t, c) x[panel row t] (CTA per column,
  // fixed-order tree) ...
  __global__ void db_bwd_gather_kernel(const int64_t j0, const int nb, const int64_t m,
                                       const int64_t ld, const double *__restrict__ band,
                                       const double *__restrict__ x, double *__restrict__ tmp)
  {
    GF_DYN_SMEM(double, s); // [blockDim.x]
    const int     c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int64_t r0 = j0 + nb;
    double        v  = 0.0;
    for (int64_t t = tid; t < m; t += nt)
      v += band[(j0 + c) * ld + (r0 + t - j0 - c)] * x[r0 + t];
    s[tid] = v;
    __syncthreads();
    for (int o = nt / 2; o > 0; o >>= 1) // nt is a power of two
      {
        if (tid < o)
          s[tid] += s[tid + o];
        __syncthreads();
      }
    if (tid == 0)
      tmp[c] = s[0];
  }
  // ... then x_k = L11^-T (x_k - tmp) (one CTA)
  __global__ void db_bwd_diag_kernel(const int64_t j0, const int nb, const int64_t ld,
                                     const double *__restrict__ band, double *__restrict__ x,
                                     const double *__restrict__ tmp)
  {
    GF_DYN_SMEM(double, s); // x_k [DB_NB]
    const int tid = threadIdx.x;
    if (tid < nb)
      s[tid] = x[j0 + tid] - tmp[tid];
    __syncthreads();
    for (int c = nb - 1; c >= 0; --c)
      {
        if (tid == 0)
          s[c] /= band[(j0 + c) * ld];
        __syncthreads();
        if (tid < c)
          s[tid] -= band[(j0 + tid) * ld + (c - tid)] * s[c];
        __syncthreads();
      }
    if (tid < nb)
      x[j0 + tid] = s[tid];
  }

  // ---- drivers: the block loops, templated on the launcher so that the emulation runs them too.
  // launch(grid, block, dynamic smem bytes, kernel, args...)
  inline int64_t db_panel_rows(int64_t n, int64_t w, int64_t j0, int nb)
  {
    const int64_t below = n - (j0 + nb);
    return below < w ? (below < 0 ? 0 : below) : w;
  }

  template <class Launch>
  void db_factor(Launch &&launch, const int64_t n, const int64_t w, const int64_t ld, double *band,
                 int *info)
  {
    const size_t smem_l = size_t(DB_NB) * (DB_NB + 1) * sizeof(double);
    for (int64_t j0 = 0; j0 < n; j0 += DB_NB)
      {
        const int     nb = int(n - j0 < DB_NB ? n - j0 : DB_NB);
        const int64_t m  = db_panel_rows(n, w, j0, nb);
        launch(1u, 256u, smem_l, db_potf2_kernel, j0, nb, ld, band, info);
        if (m == 0)
          continue;
        launch(unsigned((m + 127) / 128), 128u, smem_l, db_trsm_kernel, j0, nb, m, ld, band);
        const int64_t nt = (m + DB_TILE - 1) / DB_TILE;
        launch(unsigned(nt * (nt + 1) / 2), 256u, 2 * size_t(DB_TILE) * (DB_NB + 1) * sizeof(double),
               db_syrk_kernel, j0, nb, m, ld, band);
      }
  }

  // x (new ordering, length n) := (L L^T)^-1 x ; tmp: DB_NB doubles
  template <class Launch>
  void db_solve(Launch &&launch, const int64_t n, const int64_t w, const int64_t ld,
                const double *band, double *x, double *tmp)
  {
    const size_t smem_x = size_t(DB_NB) * sizeof(double);
    for (int64_t j0 = 0; j0 < n; j0 += DB_NB)
      {
        const int     nb = int(n - j0 < DB_NB ? n - j0 : DB_NB);
        const int64_t m  = db_panel_rows(n, w, j0, nb);
        launch(1u, unsigned(DB_NB), smem_x, db_fwd_diag_kernel, j0, nb, ld, band, x);
        if (m > 0)
          launch(unsigned((m + 127) / 128), 128u, size_t(0), db_fwd_update_kernel, j0, nb, m, ld,
                 band, x);
      }
    const int64_t last = ((n - 1) / DB_NB) * DB_NB;
    for (int64_t j0 = last; j0 >= 0; j0 -= DB_NB)
      {
        const int     nb = int(n - j0 < DB_NB ? n - j0 : DB_NB);
        const int64_t m  = db_panel_rows(n, w, j0, nb);
        // tmp = P^T x[panel] (zero without a panel: m = 0 leaves the sums empty)
        launch(unsigned(nb), 128u, 128 * sizeof(double), db_bwd_gather_kernel, j0, nb, m, ld, band,
               (const double *)x, tmp);
        launch(1u, unsigned(DB_NB), smem_x, db_bwd_diag_kernel, j0, nb, ld, band, x,
               (const double *)tmp);
      }
  }
} // namespace gf
