// K1g / K2g: the neo-Hookean cell and interface-face kernels for ANY polynomial degree
// (runtime npc / nq). They serve the degrees the tuned nl_cells_kernel<DIM, P> of assemble_nl.cu
// has no instantiation for - the reference's shipped parameter files ask for degree 3
// (parameters.prm:21) and 4 (nonlinear_elasticity.prm:24) - and follow the same reference lines:
// assemble_system_tangent_residual_one_cell (nonlinear_elasticity.cc:872-1036), the material
// closed forms (compressible_neo_hook_material.h:62-138), assemble_neumann_contribution_one_cell
// (nonlinear_elasticity.cc:791-859 incl. the cell-q-point quirk :825-827).
//
// Same outputs as the tuned kernels: K_e row-major dpc x dpc per cell in the INTERNAL local order
// (node a, component c) -> a * dim + c, only the node blocks b <= a (the reference's j <= i loop
// :1003-1035; scatter.cu mirrors them), r_e per cell.
//
// Structure (one CTA per cell, grid-stride; plain __syncthreads between the phases, no warp-level
// primitives - so the SAME source also compiles with g++ against tests/cuda_emu/, where
// tests/test_cuda_emulation.py runs it on CPU threads against the oracle):
//   per chunk of QC quadrature points
//     phase A  thread per q-point: F, J, F^-1, b-bar, tau, Jc scaled by JxW     (:902-934,:958-961)
//     phase B  thread per (q, node a): g_a = grad_X N_a F^-1 (:946-947), T_a = B_a^T (JxW D),
//              t_a = JxW tau g_a
//     residual thread per local dof, fixed q order (:984-995)
//     phase C  thread per node pair (a, b), b <= a: K_ab += T_a B_b (:1011), S_ab += t_a . g_b
//              (:1018-1019); the running K_e lives in the element buffer in HBM/L2 (a 3D Q3 cell
//              has 2080 node pairs x 9 entries - more than a CTA's registers)
//   tables N, dN are read from global memory (3D Q3: 256 KB, L2-resident).
// AFFINE = false: general (non-affine) hexahedra / quadrilaterals, i.e. what a real deal.II host
// hands over once the mesh is not a box (MappingQ1 geometry, as MappingQGeneric gives for
// straight-sided cells, linear_elasticity.cc:60): J, J^-1, det J are evaluated per quadrature
// point from the cell's vertices and the unit-cell gradients of the 2^dim vertex shape functions;
// the mass term is integrated per q-point instead of taken from the reference mass matrix.
// Not a hot path of the bench configurations (degrees 1, 2, box meshes); correctness first.
#pragma once
#include "emu_compat.cuh"
#include "kernel_utils.cuh"
#include "nl_material.cuh"

namespace gf
{
  // MappingQ1: J(i, j) = dX_i / dxi_j = sum_v X_v(i) dphi_v/dxi_j; returns det J, fills J^-1
  template <int DIM>
  __device__ inline double q1_geometry(const double *__restrict__ verts /* [2^DIM][DIM] */,
                                       const double *__restrict__ dphi /* [2^DIM][DIM] */,
                                       double (&Jinv)[DIM][DIM])
  {
    double J[DIM][DIM];
    for (int i = 0; i < DIM; ++i)
      for (int j = 0; j < DIM; ++j)
        {
          double v = 0;
          for (int k = 0; k < (1 << DIM); ++k)
            v += verts[k * DIM + i] * dphi[k * DIM + j];
          J[i][j] = v;
        }
    const double d = det<DIM>(J);
    inverse<DIM>(J, d, Jinv);
    return d;
  }

  template <int DIM>
  struct NLGen
  {
    static constexpr int VO    = DIM * (DIM + 1) / 2;
    static constexpr int TS    = DIM * VO + DIM; // T_a (DIM x VO), then t_a (DIM)
    static constexpr int QC    = 8;              // quadrature points per chunk
    // per-q record
    static constexpr int Q_D   = 0;                 // JxW * Jc (Voigt)        [VO*VO]
    static constexpr int Q_TAU = Q_D + VO * VO;     // JxW * tau (Voigt)       [VO]
    static constexpr int Q_C   = Q_TAU + VO;        // C = Jinv * Finv         [DIM*DIM]
    static constexpr int Q_A   = Q_C + DIM * DIM;   // rho * JxW * sumN * acc  [DIM]
    static constexpr int Q_W   = Q_A + DIM;         // JxW
    static constexpr int QS    = Q_W + 1;
    // dynamic shared memory (doubles): u, acc, r (dpc each), QC records, T ring, G ring
    __host__ __device__ static constexpr long smem_doubles(int npc)
    {
      return 3L * npc * DIM + long(QC) * QS + long(QC) * npc * TS + long(QC) * DIM * npc;
    }
  };

  template <int DIM, bool AFFINE>
  __global__ void nl_cells_generic_kernel(const int64_t c0, const int64_t c1, const int npc,
                                          const int nq, const int32_t *__restrict__ cell_nodes,
                                          const double *__restrict__ geom,
                                          const double *__restrict__ cell_verts,
                                          const double *__restrict__ tab_dphi,
                                          const double *__restrict__ u_total,
                                          const double *__restrict__ accel,
                                          const double *__restrict__ tabN,
                                          const double *__restrict__ tabdN,
                                          const double *__restrict__ tabw,
                                          const double *__restrict__ Mref, const NLParams prm,
                                          double *__restrict__ ke_buf, double *__restrict__ re_buf,
                                          int *err_flag)
  {
    using C = NLGen<DIM>;
    constexpr int VO = C::VO, TS = C::TS, QC = C::QC, QS = C::QS;
    GF_DYN_SMEM(double, sm);
    const int dpc = npc * DIM, tid = threadIdx.x, nt = blockDim.x;
    double *  su = sm, *sacc = su + dpc, *sr = sacc + dpc, *sQ = sr + dpc, *sT = sQ + QC * QS,
           *sG      = sT + QC * npc * TS;
    const int n_pairs = npc * (npc + 1) / 2;
    for (int64_t cell = c0 + blockIdx.x; cell < c1; cell += gridDim.x)
      {
        __syncthreads(); // the previous cell's shared data are no longer read
        // affine cell: one J^-1 / det J record; general cell: the vertices (J per q-point)
        const double *gm = AFFINE ? geom + cell * (DIM * DIM + 1) : nullptr;
        const double *vx = AFFINE ? nullptr : cell_verts + cell * ((1 << DIM) * DIM);
        for (int i = tid; i < dpc; i += nt)
          {
            const int32_t node = cell_nodes[cell * npc + i / DIM];
            su[i]              = u_total[int64_t(node) * DIM + i % DIM];
            sacc[i]            = accel[int64_t(node) * DIM + i % DIM];
            sr[i]              = 0.0;
          }
        __syncthreads();
        double *     ke   = ke_buf + (cell - c0) * int64_t(dpc) * dpc;
        const double mfac = prm.rho * prm.alpha_1 * (AFFINE ? gm[DIM * DIM] : 1.0);
        for (int qc = 0; qc < nq; qc += QC)
          {
            const int nqc = nq - qc < QC ? nq - qc : QC;
            // ------------- phase A: kinematics + material per quadrature point -----------------
            for (int ql = tid; ql < nqc; ql += nt)
              {
                const int q = qc + ql;
                double    Jinv[DIM][DIM], detJ;
                if (AFFINE)
                  {
                    for (int i = 0; i < DIM; ++i)
                      for (int j = 0; j < DIM; ++j)
                        Jinv[i][j] = gm[i * DIM + j];
                    detJ = gm[DIM * DIM];
                  }
                else
                  {
                    detJ = q1_geometry<DIM>(vx, tab_dphi + q * ((1 << DIM) * DIM), Jinv);
                    if (!(detJ > 0.0))
                      atomicExch(err_flag, 1); // inverted cell
                  }
                double Hr[DIM][DIM], acc[DIM], sumN = 0;
                for (int i = 0; i < DIM; ++i)
                  {
                    acc[i] = 0;
                    for (int j = 0; j < DIM; ++j)
                      Hr[i][j] = 0;
                  }
                for (int a = 0; a < npc; ++a)
                  {
                    const double Na = tabN[q * npc + a];
                    sumN += Na;
                    for (int cc = 0; cc < DIM; ++cc)
                      {
                        const double ua = su[a * DIM + cc];
                        acc[cc] += sacc[a * DIM + cc] * Na;
                        for (int e = 0; e < DIM; ++e)
                          Hr[cc][e] += ua * tabdN[(q * npc + a) * DIM + e];
                      }
                  }
                double F[DIM][DIM];
                for (int i = 0; i < DIM; ++i)
                  for (int j = 0; j < DIM; ++j)
                    {
                      double h = 0;
                      for (int e = 0; e < DIM; ++e)
                        h += Hr[i][e] * Jinv[e][j];
                      F[i][j] = (i == j ? 1.0 : 0.0) + h; // Kinematics::F :927
                    }
                const double detF = det<DIM>(F); // :929
                if (!(detF > 0.0))
                  atomicExch(err_flag, 1); // Assert :935
                double Finv[DIM][DIM];
                inverse<DIM>(F, detF, Finv);            // :934
                const double s = pow(detF, -1.0 / DIM); // Kinematics::F_iso :930
                double       bbar[VO];                  // Kinematics::b :932
                for (int k = 0; k < VO; ++k)
                  {
                    const int i = voigt_i<DIM>(k), j = voigt_j<DIM>(k);
                    double    v = 0;
                    for (int e = 0; e < DIM; ++e)
                      v += (s * F[i][e]) * (s * F[j][e]);
                    bbar[k] = v;
                  }
                double tau[VO], D[VO][VO];
                neo_hooke<DIM>(prm.kappa, prm.mu, detF, bbar, tau, D); // :958-961
                const double JxW = detJ * tabw[q];
                double *     rec = sQ + ql * QS;
                for (int e = 0; e < DIM; ++e)
                  for (int l = 0; l < DIM; ++l)
                    {
                      double v = 0;
                      for (int d = 0; d < DIM; ++d)
                        v += Jinv[e][d] * Finv[d][l];
                      rec[C::Q_C + e * DIM + l] = v;
                    }
                for (int k = 0; k < VO; ++k)
                  {
                    rec[C::Q_TAU + k] = tau[k] * JxW;
                    for (int l = 0; l < VO; ++l)
                      rec[C::Q_D + k * VO + l] = D[k][l] * JxW;
                  }
                for (int cc = 0; cc < DIM; ++cc)
                  rec[C::Q_A + cc] = prm.rho * sumN * acc[cc] * JxW; // :993-995 summed over j
                rec[C::Q_W] = JxW;
              }
            __syncthreads();
            // ------------- phase B: g_a, T_a = B_a^T (JxW D), t_a = JxW tau g_a ----------------
            for (int item = tid; item < nqc * npc; item += nt)
              {
                const int     ql = item / npc, a = item - ql * npc, q = qc + ql;
                const double *rec = sQ + ql * QS;
                double        g[DIM];
                for (int l = 0; l < DIM; ++l)
                  {
                    double v = 0;
                    for (int e = 0; e < DIM; ++e)
                      v += tabdN[(q * npc + a) * DIM + e] * rec[C::Q_C + e * DIM + l];
                    g[l]                         = v;
                    sG[(ql * DIM + l) * npc + a] = v;
                  }
                double *T = sT + (ql * npc + a) * TS;
                for (int ci = 0; ci < DIM; ++ci)
                  {
                    // engineering strain of dof (a,ci): eps_(ci,l) = g[l]
                    for (int k = 0; k < VO; ++k)
                      {
                        double v = 0;
                        for (int l = 0; l < DIM; ++l)
                          v += g[l] * rec[C::Q_D + voigt_index<DIM>(ci, l) * VO + k];
                        T[ci * VO + k] = v;
                      }
                    double t = 0;
                    for (int l = 0; l < DIM; ++l)
                      t += rec[C::Q_TAU + voigt_index<DIM>(ci, l)] * g[l];
                    T[DIM * VO + ci] = t;
                  }
              }
            __syncthreads();
            // ------------- residual (:984-995), fixed q order; entry i stays with its thread -----
            for (int i = tid; i < dpc; i += nt)
              {
                const int a = i / DIM, cc = i - a * DIM;
                double    r = sr[i];
                for (int ql = 0; ql < nqc; ++ql)
                  {
                    const double *rec = sQ + ql * QS;
                    const double  Na  = tabN[(qc + ql) * npc + a];
                    r -= (sT[(ql * npc + a) * TS + DIM * VO + cc] -
                          prm.body_force[cc] * prm.rho * Na * rec[C::Q_W]);
                    r -= Na * rec[C::Q_A + cc];
                  }
                sr[i] = r;
              }
            // ------------- phase C: node pairs b <= a; pair -> thread is the same in every chunk
            for (int pair = tid; pair < n_pairs; pair += nt)
              {
                // a = largest integer with a (a + 1) / 2 <= pair
                int a = int((sqrt(8.0 * double(pair) + 1.0) - 1.0) * 0.5);
                while ((a + 1) * (a + 2) / 2 <= pair)
                  ++a;
                while (a * (a + 1) / 2 > pair)
                  --a;
                const int b = pair - a * (a + 1) / 2;
                double    K[DIM][DIM], S = 0, mm = 0;
                for (int i = 0; i < DIM; ++i)
                  for (int j = 0; j < DIM; ++j)
                    K[i][j] = 0;
                for (int ql = 0; ql < nqc; ++ql)
                  {
                    if (!AFFINE) // mass term integrated per q-point: N_a N_b JxW
                      mm = fma(tabN[(qc + ql) * npc + a] * tabN[(qc + ql) * npc + b],
                               sQ[ql * QS + C::Q_W], mm);
                    const double *T = sT + (ql * npc + a) * TS;
                    double        gb[DIM];
                    for (int l = 0; l < DIM; ++l)
                      gb[l] = sG[(ql * DIM + l) * npc + b];
                    for (int ci = 0; ci < DIM; ++ci)
                      for (int cj = 0; cj < DIM; ++cj)
                        {
                          double v = K[ci][cj];
                          for (int l = 0; l < DIM; ++l)
                            v = fma(T[ci * VO + voigt_index<DIM>(cj, l)], gb[l], v);
                          K[ci][cj] = v;
                        }
                    for (int l = 0; l < DIM; ++l)
                      S = fma(T[DIM * VO + l], gb[l], S);
                  }
                // diagonal of the node block: geometric term + mass term rho alpha_1 M_ab
                // (:1018-1021); affine cells take M_ab = detJ * reference mass matrix, once
                const double dd = S + (AFFINE ? (qc == 0 ? mfac * Mref[a * npc + b] : 0.0) : mfac * mm);
                for (int ci = 0; ci < DIM; ++ci)
                  for (int cj = 0; cj < DIM; ++cj)
                    {
                      double *     e    = ke + int64_t(a * DIM + ci) * dpc + b * DIM + cj;
                      const double base = qc == 0 ? 0.0 : *e;
                      *e                = base + (K[ci][cj] + (ci == cj ? dd : 0.0));
                    }
              }
            __syncthreads(); // the next chunk overwrites sQ / sT / sG
          }
        for (int i = tid; i < dpc; i += nt)
          re_buf[cell * dpc + i] = sr[i];
      }
  }

  // K2g: one CTA per cell that owns interface faces; faces in ascending face number
  // (cell->face_iterators(), :804). Adds into r_e after the cell kernel (:755-756 order).
  // dynamic shared memory (doubles): u, sigma, r_add (dpc each), sfac[nqf], strac[nqf][DIM]
  template <int DIM>
  __host__ __device__ constexpr long nl_faces_generic_smem_doubles(int npc, int nqf)
  {
    return 3L * npc * DIM + long(nqf) * (1 + DIM);
  }

  template <int DIM, bool AFFINE>
  __global__ void nl_faces_generic_kernel(const int n_iface_cells, const int npc, const int nqf,
                                          const int32_t *__restrict__ cell_list,
                                          const int32_t *__restrict__ face_ptr,
                                          const int32_t *__restrict__ face_no,
                                          const int32_t *__restrict__ cell_nodes,
                                          const double *__restrict__ geom,
                                          const double *__restrict__ cell_verts,
                                          const double *__restrict__ tab_dphi,
                                          const double *__restrict__ tab_dphif,
                                          const double *__restrict__ u_total,
                                          const double *__restrict__ stress,
                                          const double *__restrict__ tabdN,
                                          const double *__restrict__ tabNf,
                                          const double *__restrict__ tabwf,
                                          double *__restrict__ re_buf, int *err_flag)
  {
    GF_DYN_SMEM(double, sm);
    const int dpc = npc * DIM, tid = threadIdx.x, nt = blockDim.x;
    double *  su = sm, *ss = su + dpc, *sadd = ss + dpc, *sfac = sadd + dpc, *strac = sfac + nqf;
    const int ic = blockIdx.x;
    if (ic >= n_iface_cells)
      return;
    const int64_t cell = cell_list[ic];
    const double *gm   = AFFINE ? geom + cell * (DIM * DIM + 1) : nullptr;
    const double *vx   = AFFINE ? nullptr : cell_verts + cell * ((1 << DIM) * DIM);
    for (int i = tid; i < dpc; i += nt)
      {
        const int32_t node = cell_nodes[cell * npc + i / DIM];
        su[i]              = u_total[int64_t(node) * DIM + i % DIM];
        ss[i]              = stress[int64_t(node) * DIM + i % DIM];
        sadd[i]            = 0.0;
      }
    __syncthreads();
    for (int fi = face_ptr[ic]; fi < face_ptr[ic + 1]; ++fi)
      {
        const int face = face_no[fi];
        const int fd   = face / 2;
        for (int q = tid; q < nqf; q += nt)
          {
            // geometry at the FACE quadrature point: n da = det(J) J^-T n_ref dA_ref (constant
            // per face on an affine cell); at the CELL quadrature point of the same index for
            // the deformation gradient (:825-827)
            double Jf[DIM][DIM], Jinv[DIM][DIM], detJf;
            if (AFFINE)
              {
                for (int i = 0; i < DIM; ++i)
                  for (int j = 0; j < DIM; ++j)
                    Jf[i][j] = Jinv[i][j] = gm[i * DIM + j];
                detJf = gm[DIM * DIM];
              }
            else
              {
                detJf = q1_geometry<DIM>(vx, tab_dphif + (face * nqf + q) * ((1 << DIM) * DIM), Jf);
                const double dc = q1_geometry<DIM>(vx, tab_dphi + q * ((1 << DIM) * DIM), Jinv);
                if (!(detJf > 0.0) || !(dc > 0.0))
                  atomicExch(err_flag, 1);
              }
            double nrm[DIM], len = 0;
            for (int i = 0; i < DIM; ++i)
              {
                nrm[i] = detJf * Jf[fd][i] * ((face & 1) ? 1.0 : -1.0);
                len += nrm[i] * nrm[i];
              }
            len = sqrt(len);
            for (int i = 0; i < DIM; ++i)
              nrm[i] /= len;
            // :825-827 - CELL quadrature gradient at index f_q_point
            double Hr[DIM][DIM];
            for (int i = 0; i < DIM; ++i)
              for (int j = 0; j < DIM; ++j)
                Hr[i][j] = 0;
            for (int a = 0; a < npc; ++a)
              for (int cc = 0; cc < DIM; ++cc)
                for (int e = 0; e < DIM; ++e)
                  Hr[cc][e] += su[a * DIM + cc] * tabdN[(q * npc + a) * DIM + e];
            double F[DIM][DIM];
            for (int i = 0; i < DIM; ++i)
              for (int j = 0; j < DIM; ++j)
                {
                  double h = 0;
                  for (int e = 0; e < DIM; ++e)
                    h += Hr[i][e] * Jinv[e][j];
                  F[i][j] = (i == j ? 1.0 : 0.0) + h;
                }
            const double detF = det<DIM>(F);
            if (!(detF > 0.0))
              atomicExch(err_flag, 1);
            double Finv[DIM][DIM];
            inverse<DIM>(F, detF, Finv);
            // n_star = det F F^-T N  (:831-833)
            double n2 = 0;
            for (int i = 0; i < DIM; ++i)
              {
                double v = 0;
                for (int j = 0; j < DIM; ++j)
                  v += (detF * Finv[j][i]) * nrm[j];
                n2 += v * v;
              }
            const double fac = sqrt(n2);
            // local_stress(q) = sum_k N_k(face q) sigma_k (:815-816)
            double t[DIM];
            for (int cc = 0; cc < DIM; ++cc)
              t[cc] = 0;
            for (int a = 0; a < npc; ++a)
              {
                const double Na = tabNf[(face * nqf + q) * npc + a];
                for (int cc = 0; cc < DIM; ++cc)
                  t[cc] += ss[a * DIM + cc] * Na;
              }
            sfac[q] = len * tabwf[q]; // JxW on the face
            for (int cc = 0; cc < DIM; ++cc)
              strac[q * DIM + cc] = t[cc] * fac; // referential_stress :836-837
          }
        __syncthreads();
        for (int i = tid; i < dpc; i += nt)
          {
            const int a = i / DIM, cc = i - a * DIM;
            double    r = sadd[i];
            for (int q = 0; q < nqf; ++q) // :853-854
              r += (tabNf[(face * nqf + q) * npc + a] * strac[q * DIM + cc]) * sfac[q];
            sadd[i] = r;
          }
        __syncthreads();
      }
    for (int i = tid; i < dpc; i += nt)
      re_buf[cell * dpc + i] += sadd[i];
  }
} // namespace gf
