// K3: deterministic, atomic-free scatter of element matrices / vectors into the device BSR
// matrix and DoF vectors, with the homogeneous-Dirichlet handling of
// AffineConstraints::distribute_local_to_global fused in (reference: copy_local_to_global_ASM,
// nonlinear_elasticity.cc:760-774; linear: stiffness_matrix.add :327-334 and
// MatrixTools::apply_boundary_values :448-451).
//
// Gather formulation: every BSR block owns the ordered (ascending cell = WorkStream copier
// order, :1078-1084) list of element-matrix blocks that sum into it; one warp per block row, one
// lane per dim x dim block. HBM-bound: reads dpc^2*8 B per cell, writes the
// BSR values once.
#include "gf_context.h"
#include "kernel_utils.cuh"

namespace gf
{
  namespace
  {
    // SYM: the element buffer holds only the node blocks (a, b) with b <= a (the neo-Hookean cell
    // kernel computes the lower triangle, nonlinear_elasticity.cc:1003-1035); an upper block is read
    // as the transpose of its mirror, which is exactly the reference's copy lower -> upper.
    template <int DIM, bool SYM>
    __global__ void __launch_bounds__(256, 3) scatter_matrix_kernel(const int64_t n_rows, const int npc,
                                          const int32_t *__restrict__ brow_ptr,
                                          const int64_t *__restrict__ val_ptr,
                                          const int64_t *__restrict__ cand_ptr,
                                          const uint16_t *__restrict__ src_off,
                                          const int32_t *__restrict__ row_src,
                                          const int32_t *__restrict__ bcol,
                                          const uint8_t *__restrict__ constrained,
                                          const double *__restrict__ ke_buf, const int64_t c0,
                                          const int64_t c1, const bool first,
                                          const bool apply_constraints, double *__restrict__ val)
    {
      const int     lane = threadIdx.x & 31;
      const int64_t A    = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
      if (A >= n_rows)
        return;
      const int     dpc    = npc * DIM;
      const int32_t b0     = brow_ptr[A];
      const int     nb     = brow_ptr[A + 1] - b0;
      const int64_t vbase  = val_ptr[A];
      const int     stride = int((val_ptr[A + 1] - vbase) / DIM);
      const int64_t cbase  = cand_ptr[A];
      const int     ctotal = int(cand_ptr[A + 1] - cbase);
      const int     npc2   = npc * npc;
      bool          row_con[DIM];
#pragma unroll
      for (int r = 0; r < DIM; ++r)
        row_con[r] = apply_constraints && constrained[A * DIM + r] != 0;
      bool row_any = false;
#pragma unroll
      for (int r = 0; r < DIM; ++r)
        row_any |= row_con[r];
      if (first && lane == 0 && stride > nb * DIM) // padding double of every scalar row
#pragma unroll
        for (int r = 0; r < DIM; ++r)
          val[vbase + r * stride + nb * DIM] = 0.0;
      // one lane per dim x dim block: the index arithmetic and the source-list walk are shared
      // by the block's dim*dim entries; stores of consecutive lanes are contiguous per scalar row
      for (int blk = lane; blk < nb; blk += 32)
        {
          const int32_t B  = bcol[b0 + blk];
          const int     s0 = src_off[b0 + blk];
          const int     s1 = blk + 1 < nb ? int(src_off[b0 + blk + 1]) : ctotal;
          bool          col_con[DIM];
#pragma unroll
          for (int cc = 0; cc < DIM; ++cc)
            col_con[cc] = apply_constraints && constrained[int64_t(B) * DIM + cc] != 0;
          double sum[DIM][DIM];
          // running value continues across element-buffer chunks: same order as one pass
#pragma unroll
          for (int r = 0; r < DIM; ++r)
#pragma unroll
            for (int cc = 0; cc < DIM; ++cc)
              sum[r][cc] = first ? 0.0 : val[vbase + r * stride + blk * DIM + cc];
          bool col_any = false;
#pragma unroll
          for (int cc = 0; cc < DIM; ++cc)
            col_any |= col_con[cc];
          if (!row_any && !col_any)
            {
              // fast path (all but the clamped nodes): plain ordered sums of the source blocks
              for (int s = s0; s < s1; ++s)
                {
                  const int32_t src  = row_src[cbase + s];
                  const int64_t cell = src / npc2;
                  if (cell < c0 || cell >= c1)
                    continue;
                  const int     ab = src - int32_t(cell) * npc2;
                  const int     a = ab / npc, b = ab - a * npc;
                  const bool    tr = SYM && b > a;
                  const double *kb = ke_buf + (cell - c0) * int64_t(dpc) * dpc +
                                     ((tr ? b : a) * DIM) * dpc + (tr ? a : b) * DIM;
#pragma unroll
                  for (int r = 0; r < DIM; ++r)
#pragma unroll
                    for (int cc = 0; cc < DIM; ++cc)
                      sum[r][cc] += __ldg(tr ? kb + cc * dpc + r : kb + r * dpc + cc);
                }
            }
          else
            for (int s = s0; s < s1; ++s)
              {
                const int32_t src  = row_src[cbase + s];
                const int64_t cell = src / npc2;
                if (cell < c0 || cell >= c1)
                  continue;
                const int     ab = src - int32_t(cell) * npc2;
                const int     a = ab / npc, b = ab - a * npc;
                const double *ke = ke_buf + (cell - c0) * int64_t(dpc) * dpc;
                const bool    tr = SYM && b > a;
                const double *kb = ke + ((tr ? b : a) * DIM) * dpc + (tr ? a : b) * DIM;
#pragma unroll
                for (int r = 0; r < DIM; ++r)
#pragma unroll
                  for (int cc = 0; cc < DIM; ++cc)
                    {
                      // constrained rows/columns are dropped; a constrained diagonal collects
                      // |K_e(i,i)|
                      const bool use_abs = (B == A) && (r == cc) && row_con[r];
                      const bool drop    = (row_con[r] || col_con[cc]) && !use_abs;
                      if (drop)
                        continue;
                      double v = tr ? kb[cc * dpc + r] : kb[r * dpc + cc];
                      if (use_abs)
                        {
                          v = fabs(v);
                          if (v == 0.0) // deal.II: fall back to the cell's average |diagonal|
                            {
                              for (int i = 0; i < dpc; ++i)
                                v += fabs(ke[i * dpc + i]);
                              v /= double(dpc);
                            }
                        }
                      sum[r][cc] += v;
                    }
              }
#pragma unroll
          for (int r = 0; r < DIM; ++r)
#pragma unroll
            for (int cc = 0; cc < DIM; ++cc)
              val[vbase + r * stride + blk * DIM + cc] = sum[r][cc];
        }
    }

    __global__ void scatter_mass_kernel(const int64_t n_rows, const int npc,
                                        const int32_t *__restrict__ brow_ptr,
                                        const int64_t *__restrict__ cand_ptr,
                                        const uint16_t *__restrict__ src_off,
                                        const int32_t *__restrict__ row_src,
                                        const double *__restrict__ me_buf, const int64_t c0,
                                        const int64_t c1, const bool first,
                                        double *__restrict__ mass_blk)
    {
      const int     lane = threadIdx.x & 31;
      const int64_t A    = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
      if (A >= n_rows)
        return;
      const int32_t b0     = brow_ptr[A];
      const int     nb     = brow_ptr[A + 1] - b0;
      const int64_t cbase  = cand_ptr[A];
      const int     ctotal = int(cand_ptr[A + 1] - cbase);
      const int     npc2   = npc * npc;
      for (int blk = lane; blk < nb; blk += 32)
        {
          const int s0 = src_off[b0 + blk];
          const int s1 = blk + 1 < nb ? int(src_off[b0 + blk + 1]) : ctotal;
          double    sum = first ? 0.0 : mass_blk[b0 + blk];
          for (int s = s0; s < s1; ++s)
            {
              const int32_t src  = row_src[cbase + s];
              const int64_t cell = src / npc2;
              if (cell < c0 || cell >= c1)
                continue;
              sum += me_buf[(cell - c0) * npc2 + (src - int32_t(cell) * npc2)];
            }
          mass_blk[b0 + blk] = sum;
        }
    }

    // rhs[dof] = sum over the node's cells (ascending) of r_e ; constrained rows stay 0
    template <int DIM>
    __global__ void scatter_rhs_kernel(const int64_t n_owned, const int npc,
                                       const int64_t *__restrict__ nc_ptr,
                                       const int32_t *__restrict__ nc_src,
                                       const uint8_t *__restrict__ constrained,
                                       const double *__restrict__ re_buf, const bool mask,
                                       double *__restrict__ rhs)
    {
      const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (i >= n_owned)
        return;
      const int64_t A = i / DIM;
      const int     r = int(i - A * DIM);
      double        s = 0;
      for (int64_t k = nc_ptr[A]; k < nc_ptr[A + 1]; ++k)
        {
          const int32_t src  = nc_src[k]; // cell*npc + a
          const int32_t cell = src / npc;
          const int     a    = src - cell * npc;
          s += re_buf[int64_t(cell) * npc * DIM + a * DIM + r];
        }
      if (mask && constrained[i])
        s = 0.0;
      rhs[i] = s;
    }

    // system_matrix = mass + factor * stiffness with zero-Dirichlet rows/cols eliminated and the
    // diagonal kept (MatrixTools::apply_boundary_values, linear_elasticity.cc:426-451)
    template <int DIM>
    __global__ void build_system_kernel(const int64_t n_rows, const int32_t *__restrict__ brow_ptr,
                                        const int64_t *__restrict__ val_ptr,
                                        const int32_t *__restrict__ bcol,
                                        const uint8_t *__restrict__ constrained,
                                        const double *__restrict__ kval,
                                        const double *__restrict__ mass_blk, const double factor,
                                        double *__restrict__ aval)
    {
      const int     lane = threadIdx.x & 31;
      const int64_t A    = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
      if (A >= n_rows)
        return;
      const int32_t b0     = brow_ptr[A];
      const int     nb     = brow_ptr[A + 1] - b0;
      const int64_t vbase  = val_ptr[A];
      const int     stride = int((val_ptr[A + 1] - vbase) / DIM);
      for (int idx = lane; idx < DIM * stride; idx += 32)
        {
          const int r = idx / stride, rem = idx - r * stride;
          if (rem >= nb * DIM)
            {
              aval[vbase + idx] = 0.0;
              continue;
            }
          const int     blk = rem / DIM, cc = rem - blk * DIM;
          const int32_t B = bcol[b0 + blk];
          // stepping_matrix = K * (dt^2 theta^2) + M  (:348-353)
          double v = kval[vbase + idx];
          v *= factor;
          if (r == cc)
            v += mass_blk[b0 + blk];
          const bool rc = constrained[A * DIM + r] != 0;
          const bool cn = constrained[int64_t(B) * DIM + cc] != 0;
          if ((rc || cn) && !(B == A && r == cc))
            v = 0.0;
          aval[vbase + idx] = v;
        }
    }

    // preconditioner blocks from the diagonal blocks of the matrix
    template <int DIM>
    __global__ void build_precond_kernel(const int64_t n_rows, const int32_t *__restrict__ brow_ptr,
                                         const int64_t *__restrict__ val_ptr,
                                         const int32_t *__restrict__ bcol,
                                         const double *__restrict__ val, const int kind,
                                         double *__restrict__ dinv)
    {
      const int64_t A = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (A >= n_rows)
        return;
      const int32_t b0 = brow_ptr[A];
      const int     nb = brow_ptr[A + 1] - b0;
      // linear scan: the blocks of a row are ordered by the partition-independent node order
      // (pattern.cu), which is not the local index order on a rank with ghosts below its slab
      int pos = -1;
      for (int k = 0; k < nb && pos < 0; ++k)
        if (bcol[b0 + k] == A)
          pos = k;
      const int64_t vbase  = val_ptr[A];
      const int     stride = int((val_ptr[A + 1] - vbase) / DIM);
      double        M[DIM][DIM], R[DIM][DIM];
#pragma unroll
      for (int r = 0; r < DIM; ++r)
#pragma unroll
        for (int cc = 0; cc < DIM; ++cc)
          {
            M[r][cc] = pos >= 0 ? val[vbase + r * stride + pos * DIM + cc] : (r == cc ? 1.0 : 0.0);
            R[r][cc] = r == cc ? 1.0 : 0.0;
          }
      if (kind == GF_PRECOND_JACOBI)
        {
#pragma unroll
          for (int r = 0; r < DIM; ++r)
            R[r][r] = 1.0 / M[r][r];
        }
      else if (kind >= GF_PRECOND_BLOCK_JACOBI) // multigrid smoothers use the block inverse too
        {
          // symmetrise before inverting so that the preconditioner is exactly symmetric
#pragma unroll
          for (int r = 0; r < DIM; ++r)
#pragma unroll
            for (int cc = r + 1; cc < DIM; ++cc)
              M[r][cc] = M[cc][r] = 0.5 * (M[r][cc] + M[cc][r]);
          inverse<DIM>(M, det<DIM>(M), R);
#pragma unroll
          for (int r = 0; r < DIM; ++r)
#pragma unroll
            for (int cc = r + 1; cc < DIM; ++cc)
              R[r][cc] = R[cc][r] = 0.5 * (R[r][cc] + R[cc][r]);
        }
#pragma unroll
      for (int r = 0; r < DIM; ++r)
#pragma unroll
        for (int cc = 0; cc < DIM; ++cc)
          dinv[A * DIM * DIM + r * DIM + cc] = R[r][cc];
    }
  } // namespace

  void launch_scatter_matrix(gf_context &c, double *val, int64_t c0, int64_t c1, bool first,
                             bool apply_constraints)
  {
    ProfScope      ps(c, Profile::SCATTER);
    const int64_t  n_rows = c.n_owned_nodes;
    const int      nt     = 256;
    const unsigned grid   = unsigned((n_rows * 32 + nt - 1) / nt);
    if (n_rows == 0)
      return;
    // the neo-Hookean element buffer holds the lower node blocks only (assemble_nl.cu)
    const bool sym = c.model == GF_MODEL_NEO_HOOKEAN;
#define GF_SCATTER(D, S)                                                                         \
  scatter_matrix_kernel<D, S><<<grid, nt, 0, c.stream>>>(                                        \
    n_rows, c.npc, c.brow_ptr.p, c.val_ptr.p, c.cand_ptr.p, c.src_off.p, c.row_src.p, c.bcol.p,  \
    c.constrained.p, c.ke_buf.p, c0, c1, first, apply_constraints, val)
    if (c.dim == 3)
      {
        if (sym)
          GF_SCATTER(3, true);
        else
          GF_SCATTER(3, false);
      }
    else
      {
        if (sym)
          GF_SCATTER(2, true);
        else
          GF_SCATTER(2, false);
      }
#undef GF_SCATTER
    GF_CUDA_CHECK(cudaGetLastError());
  }

  void launch_scatter_mass(gf_context &c, int64_t c0, int64_t c1, bool first)
  {
    ProfScope      ps(c, Profile::SCATTER);
    const int64_t  n_rows = c.n_owned_nodes;
    const int      nt     = 256;
    const unsigned grid   = unsigned((n_rows * 32 + nt - 1) / nt);
    if (n_rows == 0)
      return;
    scatter_mass_kernel<<<grid, nt, 0, c.stream>>>(n_rows, c.npc, c.brow_ptr.p, c.cand_ptr.p,
                                                   c.src_off.p, c.row_src.p, c.me_buf.p, c0, c1,
                                                   first, c.mass_blk.p);
    GF_CUDA_CHECK(cudaGetLastError());
  }

  void launch_scatter_rhs(gf_context &c, double *rhs, bool mask_constrained)
  {
    ProfScope      ps(c, Profile::SCATTER);
    const int      nt   = 256;
    const unsigned grid = unsigned((c.n_owned + nt - 1) / nt);
    if (c.n_owned == 0)
      return;
    if (c.dim == 3)
      scatter_rhs_kernel<3><<<grid, nt, 0, c.stream>>>(c.n_owned, c.npc, c.nc_ptr.p, c.nc_src.p,
                                                       c.constrained.p, c.re_buf.p,
                                                       mask_constrained, rhs);
    else
      scatter_rhs_kernel<2><<<grid, nt, 0, c.stream>>>(c.n_owned, c.npc, c.nc_ptr.p, c.nc_src.p,
                                                       c.constrained.p, c.re_buf.p,
                                                       mask_constrained, rhs);
    GF_CUDA_CHECK(cudaGetLastError());
  }

  void launch_build_system_matrix(gf_context &c, double factor)
  {
    ProfScope      ps(c, Profile::SCATTER);
    const int64_t  n_rows = c.n_owned_nodes;
    const int      nt     = 256;
    const unsigned grid   = unsigned((n_rows * 32 + nt - 1) / nt);
    if (n_rows == 0)
      return;
    if (c.dim == 3)
      build_system_kernel<3><<<grid, nt, 0, c.stream>>>(
        n_rows, c.brow_ptr.p, c.val_ptr.p, c.bcol.p, c.constrained.p, c.mat[GF_MAT_STIFFNESS].val.p,
        c.mass_blk.p, factor, c.mat[GF_MAT_SYSTEM].val.p);
    else
      build_system_kernel<2><<<grid, nt, 0, c.stream>>>(
        n_rows, c.brow_ptr.p, c.val_ptr.p, c.bcol.p, c.constrained.p, c.mat[GF_MAT_STIFFNESS].val.p,
        c.mass_blk.p, factor, c.mat[GF_MAT_SYSTEM].val.p);
    GF_CUDA_CHECK(cudaGetLastError());
  }

  void launch_build_precond(gf_context &c, const double *val)
  {
    ProfScope      ps(c, Profile::SCATTER);
    const int64_t  n_rows = c.n_owned_nodes;
    const int      nt     = 128;
    const unsigned grid   = unsigned((n_rows + nt - 1) / nt);
    if (n_rows == 0)
      return;
    if (c.dim == 3)
      build_precond_kernel<3><<<grid, nt, 0, c.stream>>>(n_rows, c.brow_ptr.p, c.val_ptr.p,
                                                         c.bcol.p, val, c.precond, c.dinv.p);
    else
      build_precond_kernel<2><<<grid, nt, 0, c.stream>>>(n_rows, c.brow_ptr.p, c.val_ptr.p,
                                                         c.bcol.p, val, c.precond, c.dinv.p);
    GF_CUDA_CHECK(cudaGetLastError());
  }
} // namespace gf
