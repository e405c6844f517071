// General (non-affine) cells for the linear model and the output path: what K10 (assemble_lin.cu)
// and K12 (postprocess.cu) do on parallelepipeds, with the MappingQ1 Jacobian evaluated per
// quadrature / patch point from the cell's vertices (q1_geometry, assemble_nl_generic.cuh) - the
// geometry a real deal.II host hands over once the mesh is not a box (MappingQGeneric on
// straight-sided cells, linear_elasticity.cc:60). Reference lines: local stiffness :289-323,
// MatrixCreator::create_mass_matrix :340-345, body force :358-373, consistent loading :483-520,
// output_results -> DataOut + Postprocessor (nonlinear:1215-1254, linear:590-629,
// postprocessor.h:44-76).
// Emulation-ready (emu_compat.cuh): the same source runs on the CPU in
// tests/test_cuda_emulation.py against the oracle on distorted meshes.
#pragma once
#include "assemble_nl_generic.cuh"

namespace gf
{
  // dynamic shared memory (doubles): real-space gradients [nq][npc][DIM], JxW [nq]
  template <int DIM>
  __global__ void lin_cells_general_kernel(const int64_t c0, const int64_t c1, const int npc,
                                           const int nq, const double *__restrict__ cell_verts,
                                           const double *__restrict__ tab_dphi,
                                           const double *__restrict__ tabN,
                                           const double *__restrict__ tabdN,
                                           const double *__restrict__ tabw, const double lambda,
                                           const double mu, const double rho,
                                           double *__restrict__ ke_buf, double *__restrict__ me_buf,
                                           int *err_flag)
  {
    GF_DYN_SMEM(double, sg);
    const int dpc = npc * DIM;
    double *  sw  = sg + nq * npc * DIM;
    for (int64_t cell = c0 + blockIdx.x; cell < c1; cell += gridDim.x)
      {
        const double *vx = cell_verts + cell * ((1 << DIM) * DIM);
        __syncthreads();
        for (int k = threadIdx.x; k < nq * npc; k += blockDim.x)
          {
            const int    q = k / npc;
            double       Jinv[DIM][DIM];
            const double detJ = q1_geometry<DIM>(vx, tab_dphi + q * ((1 << DIM) * DIM), Jinv);
            for (int d = 0; d < DIM; ++d)
              {
                double v = 0;
                for (int e = 0; e < DIM; ++e)
                  v += tabdN[k * DIM + e] * Jinv[e][d];
                sg[k * DIM + d] = v;
              }
            if (k == q * npc)
              {
                if (!(detJ > 0.0))
                  atomicExch(err_flag, 1); // inverted cell
                sw[q] = detJ * tabw[q];
              }
          }
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
                    for (int d = 0; d < DIM; ++d)
                      gg += gi[d] * gj[d];
                    gg *= mu;
                  }
                s += ((gi[ci] * gj[cj] * lambda) + (gi[cj] * gj[ci] * mu) + gg) * sw[q];
              }
            ke[e] = s;
          }
        // scalar mass blocks m_ab = rho sum_q N_a N_b JxW (:340-345)
        double *me = me_buf + (cell - c0) * int64_t(npc) * npc;
        for (int e = threadIdx.x; e < npc * npc; e += blockDim.x)
          {
            const int a = e / npc, b = e - a * npc;
            double    s = 0;
            for (int q = 0; q < nq; ++q)
              s += (rho * tabN[q * npc + a] * tabN[q * npc + b]) * sw[q];
            me[e] = s;
          }
      }
  }

  // body-force load per cell: r_e[a][c] = rho b_c sum_q N_a JxW   (:358-373)
  template <int DIM>
  __global__ void body_force_general_kernel(const int64_t n_cells, const int npc, const int nq,
                                            const double *__restrict__ cell_verts,
                                            const double *__restrict__ tab_dphi,
                                            const double *__restrict__ tabN,
                                            const double *__restrict__ tabw, const double rho,
                                            const double b0, const double b1, const double b2,
                                            double *__restrict__ re_buf)
  {
    const int64_t n = n_cells * npc;
    for (int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; k < n;
         k += int64_t(gridDim.x) * blockDim.x)
      {
        const int64_t cell = k / npc;
        const int     a    = int(k - cell * npc);
        const double *vx   = cell_verts + cell * ((1 << DIM) * DIM);
        double        s    = 0;
        for (int q = 0; q < nq; ++q)
          {
            double       Jinv[DIM][DIM];
            const double detJ = q1_geometry<DIM>(vx, tab_dphi + q * ((1 << DIM) * DIM), Jinv);
            s += tabN[q * npc + a] * (detJ * tabw[q]);
          }
        const double bf[3] = {b0, b1, b2};
        for (int cc = 0; cc < DIM; ++cc)
          re_buf[k * DIM + cc] = (rho * bf[cc]) * s;
      }
  }

  // consistent interface load (:483-520), one CTA per cell with interface faces; overwrites that
  // cell's r_e. dynamic shared memory (doubles): sigma [dpc], r_add [dpc], traction [nqf][DIM],
  // face JxW [nqf]
  template <int DIM>
  __global__ void lin_faces_general_kernel(const int n_iface_cells, const int npc, const int nqf,
                                           const int32_t *__restrict__ cell_list,
                                           const int32_t *__restrict__ face_ptr,
                                           const int32_t *__restrict__ face_no,
                                           const int32_t *__restrict__ cell_nodes,
                                           const double *__restrict__ cell_verts,
                                           const double *__restrict__ tab_dphif,
                                           const double *__restrict__ stress,
                                           const double *__restrict__ tabNf,
                                           const double *__restrict__ tabwf,
                                           double *__restrict__ re_buf)
  {
    GF_DYN_SMEM(double, sh);
    const int dpc = npc * DIM, tid = threadIdx.x, nt = blockDim.x;
    double *  ss = sh, *sadd = ss + dpc, *st = sadd + dpc, *sl = st + nqf * DIM;
    const int ic = blockIdx.x;
    if (ic >= n_iface_cells)
      return;
    const int64_t cell = cell_list[ic];
    const double *vx   = cell_verts + cell * ((1 << DIM) * DIM);
    for (int i = tid; i < dpc; i += nt)
      {
        ss[i]   = stress[int64_t(cell_nodes[cell * npc + i / DIM]) * DIM + i % DIM];
        sadd[i] = 0.0;
      }
    __syncthreads();
    for (int fi = face_ptr[ic]; fi < face_ptr[ic + 1]; ++fi)
      {
        const int face = face_no[fi], fd = face / 2;
        for (int k = tid; k < nqf * DIM; k += nt)
          {
            const int q = k / DIM, cc = k - q * DIM;
            double    t = 0;
            for (int a = 0; a < npc; ++a) // get_function_values :499
              t += ss[a * DIM + cc] * tabNf[(face * nqf + q) * npc + a];
            st[k] = t;
            if (cc == 0)
              {
                // |det(J) J^-T n_ref| dA_ref at the face quadrature point
                double       Jinv[DIM][DIM];
                const double detJ =
                  q1_geometry<DIM>(vx, tab_dphif + (face * nqf + q) * ((1 << DIM) * DIM), Jinv);
                double len = 0;
                for (int i = 0; i < DIM; ++i)
                  {
                    const double v = detJ * Jinv[fd][i];
                    len += v * v;
                  }
                sl[q] = sqrt(len) * tabwf[q];
              }
          }
        __syncthreads();
        for (int i = tid; i < dpc; i += nt)
          {
            const int a = i / DIM, cc = i - a * DIM;
            double    r = sadd[i];
            for (int q = 0; q < nqf; ++q) // :508-510
              r += tabNf[(face * nqf + q) * npc + a] * st[q * DIM + cc] * sl[q];
            sadd[i] = r;
          }
        __syncthreads();
      }
    for (int i = tid; i < dpc; i += nt)
      re_buf[cell * dpc + i] = sadd[i];
  }

  // K12 on general cells: one thread per (cell, patch point); see postprocess.cu
  template <int DIM>
  __global__ void postprocess_general_kernel(const int64_t c0, const int64_t n_cells, const int npc,
                                             const int32_t *__restrict__ cell_nodes,
                                             const double *__restrict__ cell_verts,
                                             const double *__restrict__ tab_dphip,
                                             const double *__restrict__ ppN,
                                             const double *__restrict__ ppdN,
                                             const double *__restrict__ u,
                                             double *__restrict__ fields,
                                             int *__restrict__ err_flag)
  {
    constexpr int NF = DIM + DIM * DIM;
    const int64_t n  = n_cells * npc;
    for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < n;
         t += int64_t(gridDim.x) * blockDim.x)
      {
        const int64_t  cell = c0 + t / npc; // fields is the chunk's buffer, indexed by t
        const int      pt   = int(t % npc);
        const int32_t *cn   = cell_nodes + cell * npc;
        double         gm[DIM][DIM];
        const double   detJ = q1_geometry<DIM>(cell_verts + cell * ((1 << DIM) * DIM),
                                               tab_dphip + pt * ((1 << DIM) * DIM), gm);
        double val[DIM], G[DIM][DIM]; // u and its unit-cell gradient at the patch point
        for (int c = 0; c < DIM; ++c)
          {
            val[c] = 0.0;
            for (int k = 0; k < DIM; ++k)
              G[c][k] = 0.0;
          }
        for (int a = 0; a < npc; ++a)
          {
            const double  Na = ppN[pt * npc + a];
            const double *dN = ppdN + (pt * npc + a) * DIM;
            const int64_t nd = cn[a];
            for (int c = 0; c < DIM; ++c)
              {
                const double ua = u[nd * DIM + c];
                val[c]          = fma(Na, ua, val[c]);
                for (int k = 0; k < DIM; ++k)
                  G[c][k] = fma(dN[k], ua, G[c][k]);
              }
          }
        double H[DIM][DIM], F[DIM][DIM], Finv[DIM][DIM]; // H = grad_X u
        for (int c = 0; c < DIM; ++c)
          for (int j = 0; j < DIM; ++j)
            {
              double s = 0.0;
              for (int k = 0; k < DIM; ++k)
                s = fma(G[c][k], gm[k][j], s);
              H[c][j] = s;
              F[c][j] = (c == j ? 1.0 : 0.0) + s;
            }
        const double dF = det<DIM>(F);
        if (!(dF > 0.0) || !(detJ > 0.0))
          {
            *err_flag = 1; // inverted (Eulerian) mapping: deal.II would assert as well
            continue;
          }
        inverse<DIM>(F, dF, Finv);
        double g[DIM][DIM]; // grad_x u = H F^-1
        for (int c = 0; c < DIM; ++c)
          for (int j = 0; j < DIM; ++j)
            {
              double s = 0.0;
              for (int k = 0; k < DIM; ++k)
                s = fma(H[c][k], Finv[k][j], s);
              g[c][j] = s;
            }
        double *out = fields + t * NF;
        for (int d = 0; d < DIM; ++d)
          {
            out[d] = val[d];
            for (int e = 0; e < DIM; ++e)
              out[DIM + d * DIM + e] = (g[d][e] + g[e][d]) / 2; // postprocessor.h:63-70
          }
      }
  }
} // namespace gf
