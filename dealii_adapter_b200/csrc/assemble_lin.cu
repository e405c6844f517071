// K10: linear-elastic element stiffness / consistent mass (once) and the interface load (per
// step). Replaces ElastoDynamics::assemble_system (linear_elasticity.cc:248-374: K_e :289-323,
// MatrixCreator::create_mass_matrix :340-345, body force :358-373) and
// assemble_consistent_loading (:458-521). Scatter: scatter.cu.
#include "assemble_general.cuh"
#include "gf_context.h"
#include "kernel_utils.cuh"

namespace gf
{
  namespace
  {
    // one CTA per cell; thread per (i,j) entry of K_e. FP64; run once, not performance critical.
    template <int DIM>
    __global__ void lin_cells_kernel(const int64_t c0, const int64_t c1, const int npc,
                                     const int nq, const double *__restrict__ geom,
                                     const double *__restrict__ tabdN,
                                     const double *__restrict__ tabw,
                                     const double *__restrict__ Mref, const double lambda,
                                     const double mu, const double rho,
                                     double *__restrict__ ke_buf, double *__restrict__ me_buf)
    {
      extern __shared__ double sg[]; // [nq][npc][DIM] real-space gradients, then [nq] JxW
      const int                dpc = npc * DIM;
      for (int64_t cell = c0 + blockIdx.x; cell < c1; cell += gridDim.x)
        {
          const double *gm = geom + cell * (DIM * DIM + 1);
          double        Jinv[DIM][DIM];
#pragma unroll
          for (int i = 0; i < DIM; ++i)
#pragma unroll
            for (int j = 0; j < DIM; ++j)
              Jinv[i][j] = gm[i * DIM + j];
          const double detJ = gm[DIM * DIM];
          __syncthreads();
          for (int k = threadIdx.x; k < nq * npc; k += blockDim.x)
            {
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                {
                  double v = 0;
#pragma unroll
                  for (int e = 0; e < DIM; ++e)
                    v += tabdN[k * DIM + e] * Jinv[e][d];
                  sg[k * DIM + d] = v;
                }
            }
          double *sw = sg + nq * npc * DIM;
          for (int q = threadIdx.x; q < nq; q += blockDim.x)
            sw[q] = detJ * tabw[q];
          __syncthreads();
          double *ke = ke_buf + (cell - c0) * int64_t(dpc) * dpc;
          for (int e = threadIdx.x; e < dpc * dpc; e += blockDim.x)
            {
              const int i = e / dpc, j = e - i * dpc;
              const int ai = i / DIM, ci = i - ai * DIM, aj = j / DIM, cj = j - aj * DIM;
              double    s = 0;
              for (int q = 0; q < nq; ++q) // :299-321
                {
                  const double *gi = sg + (q * npc + ai) * DIM;
                  const double *gj = sg + (q * npc + aj) * DIM;
                  double        gg = 0;
                  if (ci == cj)
                    {
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        gg += gi[d] * gj[d];
                      gg *= mu;
                    }
                  s += ((gi[ci] * gj[cj] * lambda) + (gi[cj] * gj[ci] * mu) + gg) * sw[q];
                }
              ke[e] = s;
            }
          double *me = me_buf + (cell - c0) * int64_t(npc) * npc;
          for (int e = threadIdx.x; e < npc * npc; e += blockDim.x)
            me[e] = rho * detJ * Mref[e];
        }
    }

    // body-force load per cell: r_e[a][c] = rho b_c detJ sum_q w N_a   (:358-373)
    template <int DIM>
    __global__ void body_force_kernel(const int64_t n_cells, const int npc,
                                      const double *__restrict__ geom,
                                      const double *__restrict__ Mref, const double rho,
                                      const double b0, const double b1, const double b2,
                                      double *__restrict__ re_buf)
    {
      const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (k >= n_cells * npc)
        return;
      const int64_t cell = k / npc;
      const int     a    = int(k - cell * npc);
      double        s    = 0;
      for (int b = 0; b < npc; ++b)
        s += Mref[a * npc + b]; // sum_q w N_a (partition of unity)
      const double detJ = geom[cell * (DIM * DIM + 1) + DIM * DIM];
      const double bf[3] = {b0, b1, b2};
#pragma unroll
      for (int cc = 0; cc < DIM; ++cc)
        re_buf[k * DIM + cc] = (rho * bf[cc]) * s * detJ;
    }

    // consistent interface load (:483-520), one CTA per cell with interface faces; overwrites
    // that cell's r_e (all other rows of re_buf stay zero)
    template <int DIM>
    __global__ void lin_faces_kernel(const int n_iface_cells, const int npc, const int nqf,
                                     const int32_t *__restrict__ cell_list,
                                     const int32_t *__restrict__ face_ptr,
                                     const int32_t *__restrict__ face_no,
                                     const int32_t *__restrict__ cell_nodes,
                                     const double *__restrict__ geom,
                                     const double *__restrict__ stress,
                                     const double *__restrict__ tabNf,
                                     const double *__restrict__ tabwf, double *__restrict__ re_buf)
    {
      extern __shared__ double sh[]; // ss[dpc], st[nqf*DIM]
      const int                dpc = npc * DIM;
      double *                 ss = sh, *st = sh + dpc;
      const int                ic = blockIdx.x;
      if (ic >= n_iface_cells)
        return;
      const int     tid  = threadIdx.x;
      const int64_t cell = cell_list[ic];
      const double *gm   = geom + cell * (DIM * DIM + 1);
      const double  detJ = gm[DIM * DIM];
      if (tid < dpc)
        ss[tid] = stress[int64_t(cell_nodes[cell * npc + tid / DIM]) * DIM + tid % DIM];
      __syncthreads();
      double r_add = 0;
      for (int fi = face_ptr[ic]; fi < face_ptr[ic + 1]; ++fi)
        {
          const int face = face_no[fi], fd = face / 2;
          double    len = 0;
#pragma unroll
          for (int i = 0; i < DIM; ++i)
            {
              const double v = detJ * gm[fd * DIM + i];
              len += v * v;
            }
          len = sqrt(len);
          for (int k = tid; k < nqf * DIM; k += blockDim.x)
            {
              const int q = k / DIM, cc = k - q * DIM;
              double    t = 0;
              for (int a = 0; a < npc; ++a) // get_function_values :499
                t += ss[a * DIM + cc] * tabNf[(face * nqf + q) * npc + a];
              st[k] = t;
            }
          __syncthreads();
          if (tid < dpc)
            {
              const int a = tid / DIM, cc = tid % DIM;
              for (int q = 0; q < nqf; ++q) // :508-510
                r_add += tabNf[(face * nqf + q) * npc + a] * st[q * DIM + cc] * (len * tabwf[q]);
            }
          __syncthreads();
        }
      if (tid < dpc)
        re_buf[cell * dpc + tid] = r_add;
    }
  } // namespace

  void launch_lin_cells(gf_context &c, int64_t c0, int64_t c1)
  {
    ProfScope     ps(c, Profile::ASM_CELLS);
    const double  lambda = 2 * c.desc.mu * c.desc.nu / (1 - 2 * c.desc.nu); // parameters.cc:189
    const int64_t n      = c1 - c0;
    if (n <= 0)
      return;
    const int    grid = int(std::min<int64_t>(n, int64_t(c.sm_count) * 8));
    const size_t smem = (size_t(c.tables.nq) * c.npc * c.dim + c.tables.nq) * sizeof(double);
    // beyond the 48 KB default from 3D Q3 on (64 q-points x 64 nodes x 3 gradients: 98 KB)
    static size_t configured[2] = {48 * 1024, 48 * 1024};
    if (smem > configured[c.dim - 2])
      {
        GF_REQUIRE(smem <= 227 * 1024, GF_ERR_UNSUPPORTED,
                   "polynomial degree too high for the linear cell kernel's shared memory");
        if (c.dim == 3)
          GF_CUDA_CHECK(cudaFuncSetAttribute(
            lin_cells_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        else
          GF_CUDA_CHECK(cudaFuncSetAttribute(
            lin_cells_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured[c.dim - 2] = smem;
      }
    if (!c.affine)
      {
        // general cells: per-q-point Jacobians (assemble_general.cuh); same shared-memory layout
        static size_t configured_g[2] = {48 * 1024, 48 * 1024};
        if (smem > configured_g[c.dim - 2])
          {
            GF_REQUIRE(smem <= 227 * 1024, GF_ERR_UNSUPPORTED,
                       "polynomial degree too high for the linear cell kernel's shared memory");
            if (c.dim == 3)
              GF_CUDA_CHECK(cudaFuncSetAttribute(lin_cells_general_kernel<3>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 int(smem)));
            else
              GF_CUDA_CHECK(cudaFuncSetAttribute(lin_cells_general_kernel<2>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 int(smem)));
            configured_g[c.dim - 2] = smem;
          }
        if (c.dim == 3)
          lin_cells_general_kernel<3><<<grid, 256, smem, c.stream>>>(
            c0, c1, c.npc, c.tables.nq, c.cell_verts.p, c.tables.dphi.p, c.tables.N.p, c.tables.dN.p,
            c.tables.w.p, lambda, c.desc.mu, c.desc.rho, c.ke_buf.p, c.me_buf.p, c.err_flag.p);
        else
          lin_cells_general_kernel<2><<<grid, 256, smem, c.stream>>>(
            c0, c1, c.npc, c.tables.nq, c.cell_verts.p, c.tables.dphi.p, c.tables.N.p, c.tables.dN.p,
            c.tables.w.p, lambda, c.desc.mu, c.desc.rho, c.ke_buf.p, c.me_buf.p, c.err_flag.p);
      }
    else if (c.dim == 3)
      lin_cells_kernel<3><<<grid, 256, smem, c.stream>>>(
        c0, c1, c.npc, c.tables.nq, c.geom.p, c.tables.dN.p, c.tables.w.p, c.tables.Mref.p, lambda,
        c.desc.mu, c.desc.rho, c.ke_buf.p, c.me_buf.p);
    else
      lin_cells_kernel<2><<<grid, 256, smem, c.stream>>>(
        c0, c1, c.npc, c.tables.nq, c.geom.p, c.tables.dN.p, c.tables.w.p, c.tables.Mref.p, lambda,
        c.desc.mu, c.desc.rho, c.ke_buf.p, c.me_buf.p);
    GF_CUDA_CHECK(cudaGetLastError());
  }

  void launch_body_force(gf_context &c, double *out)
  {
    const int64_t  n    = c.n_cells * c.npc;
    const unsigned grid = unsigned((n + 255) / 256);
    if (!c.affine)
      {
        const unsigned g = std::min(grid, 65535u);
        if (c.dim == 3)
          body_force_general_kernel<3><<<g, 256, 0, c.stream>>>(
            c.n_cells, c.npc, c.tables.nq, c.cell_verts.p, c.tables.dphi.p, c.tables.N.p,
            c.tables.w.p, c.desc.rho, c.desc.body_force[0], c.desc.body_force[1],
            c.desc.body_force[2], c.re_buf.p);
        else
          body_force_general_kernel<2><<<g, 256, 0, c.stream>>>(
            c.n_cells, c.npc, c.tables.nq, c.cell_verts.p, c.tables.dphi.p, c.tables.N.p,
            c.tables.w.p, c.desc.rho, c.desc.body_force[0], c.desc.body_force[1],
            c.desc.body_force[2], c.re_buf.p);
      }
    else if (c.dim == 3)
      body_force_kernel<3><<<grid, 256, 0, c.stream>>>(c.n_cells, c.npc, c.geom.p, c.tables.Mref.p,
                                                       c.desc.rho, c.desc.body_force[0],
                                                       c.desc.body_force[1], c.desc.body_force[2],
                                                       c.re_buf.p);
    else
      body_force_kernel<2><<<grid, 256, 0, c.stream>>>(c.n_cells, c.npc, c.geom.p, c.tables.Mref.p,
                                                       c.desc.rho, c.desc.body_force[0],
                                                       c.desc.body_force[1], c.desc.body_force[2],
                                                       c.re_buf.p);
    GF_CUDA_CHECK(cudaGetLastError());
    launch_scatter_rhs(c, out, false);
    GF_CUDA_CHECK(cudaMemsetAsync(c.re_buf.p, 0, c.re_buf.n * sizeof(double), c.stream));
  }

  void launch_lin_faces(gf_context &c, const double *stress, double *rhs_out)
  {
    {
      ProfScope ps(c, Profile::ASM_FACES);
      if (c.n_iface_cells > 0)
        {
          const int    nt   = ((c.dpc + 31) / 32) * 32;
          const size_t smem = (size_t(c.dpc) + size_t(c.tables.nqf) * c.dim) * sizeof(double);
          if (!c.affine)
            {
              const size_t smem_g =
                (2 * size_t(c.dpc) + size_t(c.tables.nqf) * (c.dim + 1)) * sizeof(double);
              if (c.dim == 3)
                lin_faces_general_kernel<3><<<unsigned(c.n_iface_cells), 128, smem_g, c.stream>>>(
                  int(c.n_iface_cells), c.npc, c.tables.nqf, c.iface_cell_list.p,
                  c.iface_face_ptr.p, c.iface_face_no.p, c.cell_nodes.p, c.cell_verts.p,
                  c.tables.dphif.p, stress, c.tables.Nf.p, c.tables.wf.p, c.re_buf.p);
              else
                lin_faces_general_kernel<2><<<unsigned(c.n_iface_cells), 128, smem_g, c.stream>>>(
                  int(c.n_iface_cells), c.npc, c.tables.nqf, c.iface_cell_list.p,
                  c.iface_face_ptr.p, c.iface_face_no.p, c.cell_nodes.p, c.cell_verts.p,
                  c.tables.dphif.p, stress, c.tables.Nf.p, c.tables.wf.p, c.re_buf.p);
            }
          else if (c.dim == 3)
            lin_faces_kernel<3><<<unsigned(c.n_iface_cells), nt, smem, c.stream>>>(
              int(c.n_iface_cells), c.npc, c.tables.nqf, c.iface_cell_list.p, c.iface_face_ptr.p,
              c.iface_face_no.p, c.cell_nodes.p, c.geom.p, stress, c.tables.Nf.p, c.tables.wf.p,
              c.re_buf.p);
          else
            lin_faces_kernel<2><<<unsigned(c.n_iface_cells), nt, smem, c.stream>>>(
              int(c.n_iface_cells), c.npc, c.tables.nqf, c.iface_cell_list.p, c.iface_face_ptr.p,
              c.iface_face_no.p, c.cell_nodes.p, c.geom.p, stress, c.tables.Nf.p, c.tables.wf.p,
              c.re_buf.p);
          GF_CUDA_CHECK(cudaGetLastError());
        }
    }
    launch_scatter_rhs(c, rhs_out, false);
  }
} // namespace gf
