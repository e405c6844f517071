// K11: matrix-free, sum-factorised tangent operator of the neo-Hookean solid — the variant
// BASELINE's north_star asks to be benchmarked beside the assembled SpMV (K4). It applies exactly
// the bilinear form the reference assembles (nonlinear_elasticity.cc:1011-1023):
//   (K x)_i = sum_q [ sym grad_x N_i : Jc : sym grad_x x  +  grad_x N_i . tau . grad_x x_(c_i)
//                     + N_i rho alpha_1 x_(c_i) ] JxW
// with the homogeneous-Dirichlet treatment of distribute_local_to_global (:769-773): constrained
// rows/columns dropped, their diagonal = sum over cells of |K_e(i,i)|.
//
//   mf_setup  (per Newton iteration, replaces K1 + matrix scatter on the finest level): per cell and
//             quadrature point the kinematics + material of :902-961 once; stores
//             C = J^-1 F^-1 (9), JxW*Jc (21, upper Voigt), JxW*tau (6) -> 36 doubles per point
//             ([cell][field][q] so that the apply kernel reads coalesced), the residual r_e
//             (:984-995) and the diagonal node blocks of K_e for the block-Jacobi / Chebyshev
//             smoothers. No element matrix is formed.
//   mf_apply  16 threads per cell (one per (qx,qy) column, qz in registers): x_e -> reference
//             gradients by 1D contractions (sum factorisation), the point operator above, the
//             transposed contractions -> y_e; a gather kernel sums y_e per node in ascending cell
//             order (no FP atomics) and fuses the CG dot product.
// Roofline: HBM-bound on the quadrature data, 36*8*nq B per cell = 18.4 KB (3D Q2) against 40 KB
// per cell for the assembled matrix; ~16 k FMA per cell keep the FP64 pipe below its limit.
#include "gf_context.h"
#include "kernel_utils.cuh"
#include "nl_material.cuh"

namespace gf
{
  namespace
  {
    template <int DIM>
    struct MFDim
    {
      static constexpr int VO = DIM * (DIM + 1) / 2;
      static constexpr int NF = DIM * DIM + VO * (VO + 1) / 2 + VO; // fields per quadrature point
      static constexpr int F_C = 0, F_D = DIM * DIM, F_TAU = DIM * DIM + VO * (VO + 1) / 2;
    };
    // index of (k,l), k <= l, in the packed upper triangle of a VO x VO symmetric matrix
    template <int VO>
    __host__ __device__ constexpr int upper_index(int k, int l)
    {
      return k <= l ? k * VO - k * (k - 1) / 2 + (l - k) : l * VO - l * (l - 1) / 2 + (k - l);
    }

    template <int P>
    struct Tab1D // passed by value, copied to shared memory
    {
      double N[(P + 2) * (P + 1)], D[(P + 2) * (P + 1)], w[P + 2];
      int    lex2hier[(P + 1) * (P + 1) * (P + 1)];
    };

    // ---------------------------------------------------------------------------------------------
    // set-up: quadrature-point data, residual, diagonal blocks. One CTA per cell.
    // ---------------------------------------------------------------------------------------------
    template <int DIM, int P>
    struct SetupCfg
    {
      static constexpr int NPC = ipow(P + 1, DIM), DPC = NPC * DIM, NQ = ipow(P + 2, DIM);
      static constexpr int VO  = DIM * (DIM + 1) / 2;
      static constexpr int QCH = 4; // q-chunks per node in the diagonal-block pass
      static constexpr int NEED = NQ > DPC ? (NQ > NPC * QCH ? NQ : NPC * QCH) :
                                             (DPC > NPC * QCH ? DPC : NPC * QCH);
      static constexpr int NT = ((NEED + 31) / 32) * 32;
      static constexpr int QS = DIM * DIM + VO * VO + VO + DIM + 2; // as in assemble_nl.cu
      static constexpr int Q_C = 0, Q_D = DIM * DIM, Q_TAU = Q_D + VO * VO, Q_A = Q_TAU + VO,
                           Q_W = Q_A + DIM;
      static constexpr int OFF_N = 0, OFF_DN = NQ * NPC, OFF_U = OFF_DN + NQ * NPC * DIM,
                           OFF_ACC = OFF_U + DPC, OFF_Q = OFF_ACC + DPC,
                           OFF_R = OFF_Q + NQ * QS, // diagonal partials [NPC][QCH][DIM*DIM+1]
                           SMEM_D = OFF_R + NPC * QCH * (DIM * DIM + 1);
      static constexpr size_t SMEM_BYTES = size_t(SMEM_D) * sizeof(double);
    };

    template <int DIM, int P>
    __global__ void __launch_bounds__(SetupCfg<DIM, P>::NT)
      mf_setup_kernel(const int64_t n_cells, const int32_t *__restrict__ cell_nodes,
                      const double *__restrict__ geom, const double *__restrict__ u_total,
                      const double *__restrict__ accel, const double *__restrict__ tabN,
                      const double *__restrict__ tabdN, const double *__restrict__ tabw,
                      const double *__restrict__ Mref, const NLParams prm,
                      double *__restrict__ qp_buf, double *__restrict__ re_buf,
                      double *__restrict__ diag_e, int *err_flag)
    {
      using C = SetupCfg<DIM, P>;
      using M = MFDim<DIM>;
      constexpr int NPC = C::NPC, DPC = C::DPC, NQ = C::NQ, VO = C::VO, QS = C::QS, QCH = C::QCH;
      extern __shared__ __align__(16) double sm[];
      double *sN = sm + C::OFF_N, *sdN = sm + C::OFF_DN, *su = sm + C::OFF_U,
             *sacc = sm + C::OFF_ACC, *sQ = sm + C::OFF_Q, *sR = sm + C::OFF_R;
      const int tid = threadIdx.x;
      for (int i = tid; i < NQ * NPC; i += C::NT)
        sN[i] = tabN[i];
      for (int i = tid; i < NQ * NPC * DIM; i += C::NT)
        sdN[i] = tabdN[i];
      for (int64_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x)
        {
          __syncthreads();
          const double *gm = geom + cell * (DIM * DIM + 1);
          double        Jinv[DIM][DIM];
#pragma unroll
          for (int i = 0; i < DIM; ++i)
#pragma unroll
            for (int j = 0; j < DIM; ++j)
              Jinv[i][j] = gm[i * DIM + j];
          const double detJ = gm[DIM * DIM];
          if (tid < DPC)
            {
              const int32_t node = cell_nodes[cell * NPC + tid / DIM];
              su[tid]            = u_total[int64_t(node) * DIM + tid % DIM];
              sacc[tid]          = accel[int64_t(node) * DIM + tid % DIM];
            }
          __syncthreads();
          // ---- kinematics + material per quadrature point (:902-961) -------------------------
          if (tid < NQ)
            {
              const int q = tid;
              double    Hr[DIM][DIM], acc[DIM], sumN = 0;
#pragma unroll
              for (int i = 0; i < DIM; ++i)
                {
                  acc[i] = 0;
#pragma unroll
                  for (int j = 0; j < DIM; ++j)
                    Hr[i][j] = 0;
                }
              for (int a = 0; a < NPC; ++a)
                {
                  const double Na = sN[q * NPC + a];
                  sumN += Na;
#pragma unroll
                  for (int cc = 0; cc < DIM; ++cc)
                    {
                      const double ua = su[a * DIM + cc];
                      acc[cc] += sacc[a * DIM + cc] * Na;
#pragma unroll
                      for (int e = 0; e < DIM; ++e)
                        Hr[cc][e] += ua * sdN[(q * NPC + a) * DIM + e];
                    }
                }
              double F[DIM][DIM];
#pragma unroll
              for (int i = 0; i < DIM; ++i)
#pragma unroll
                for (int j = 0; j < DIM; ++j)
                  {
                    double h = 0;
#pragma unroll
                    for (int e = 0; e < DIM; ++e)
                      h += Hr[i][e] * Jinv[e][j];
                    F[i][j] = (i == j ? 1.0 : 0.0) + h;
                  }
              const double detF = det<DIM>(F);
              if (!(detF > 0.0))
                atomicExch(err_flag, 1); // Assert :935
              double Finv[DIM][DIM];
              inverse<DIM>(F, detF, Finv);
              const double s = pow(detF, -1.0 / DIM);
              double       bbar[VO];
#pragma unroll
              for (int k = 0; k < VO; ++k)
                {
                  const int i = voigt_i<DIM>(k), j = voigt_j<DIM>(k);
                  double    v = 0;
#pragma unroll
                  for (int e = 0; e < DIM; ++e)
                    v += (s * F[i][e]) * (s * F[j][e]);
                  bbar[k] = v;
                }
              double tau[VO], D[VO][VO];
              neo_hooke<DIM>(prm.kappa, prm.mu, detF, bbar, tau, D);
              const double JxW = detJ * tabw[q];
              double *     rec = sQ + q * QS;
              double *     out = qp_buf + cell * int64_t(M::NF) * NQ + q; // [field][q]
#pragma unroll
              for (int e = 0; e < DIM; ++e)
#pragma unroll
                for (int l = 0; l < DIM; ++l)
                  {
                    double v = 0;
#pragma unroll
                    for (int d = 0; d < DIM; ++d)
                      v += Jinv[e][d] * Finv[d][l];
                    rec[C::Q_C + e * DIM + l]                  = v;
                    out[int64_t(M::F_C + e * DIM + l) * NQ] = v;
                  }
#pragma unroll
              for (int k = 0; k < VO; ++k)
                {
                  rec[C::Q_TAU + k]                  = tau[k] * JxW;
                  out[int64_t(M::F_TAU + k) * NQ] = tau[k] * JxW;
#pragma unroll
                  for (int l = 0; l < VO; ++l)
                    {
                      rec[C::Q_D + k * VO + l] = D[k][l] * JxW;
                      if (l >= k)
                        out[int64_t(M::F_D + upper_index<VO>(k, l)) * NQ] = D[k][l] * JxW;
                    }
                }
#pragma unroll
              for (int cc = 0; cc < DIM; ++cc)
                rec[C::Q_A + cc] = prm.rho * sumN * acc[cc] * JxW; // :993-995 summed over j
              rec[C::Q_W] = JxW;
            }
          __syncthreads();
          // ---- residual (:984-995), fixed q order --------------------------------------------
          if (tid < DPC)
            {
              const int a = tid / DIM, cc = tid % DIM;
              double    r_i = 0;
              for (int q = 0; q < NQ; ++q)
                {
                  const double *rec = sQ + q * QS;
                  double        g[DIM];
#pragma unroll
                  for (int l = 0; l < DIM; ++l)
                    {
                      double v = 0;
#pragma unroll
                      for (int e = 0; e < DIM; ++e)
                        v += sdN[(q * NPC + a) * DIM + e] * rec[C::Q_C + e * DIM + l];
                      g[l] = v;
                    }
                  double t = 0;
#pragma unroll
                  for (int l = 0; l < DIM; ++l)
                    t += rec[C::Q_TAU + voigt_index<DIM>(cc, l)] * g[l];
                  const double Na = sN[q * NPC + a];
                  r_i -= (t - prm.body_force[cc] * prm.rho * Na * rec[C::Q_W]);
                  r_i -= Na * rec[C::Q_A + cc];
                }
              re_buf[cell * DPC + tid] = r_i;
            }
          // ---- diagonal node blocks of K_e (:1011-1023 for i, j on the same node) --------------
          if (tid < NPC * QCH)
            {
              const int a = tid / QCH, ch = tid % QCH;
              double    Kd[DIM][DIM], sg = 0;
#pragma unroll
              for (int i = 0; i < DIM; ++i)
#pragma unroll
                for (int j = 0; j < DIM; ++j)
                  Kd[i][j] = 0;
              for (int q = ch; q < NQ; q += QCH)
                {
                  const double *rec = sQ + q * QS;
                  double        g[DIM];
#pragma unroll
                  for (int l = 0; l < DIM; ++l)
                    {
                      double v = 0;
#pragma unroll
                      for (int e = 0; e < DIM; ++e)
                        v += sdN[(q * NPC + a) * DIM + e] * rec[C::Q_C + e * DIM + l];
                      g[l] = v;
                    }
#pragma unroll
                  for (int ci = 0; ci < DIM; ++ci)
                    {
                      double T[VO];
#pragma unroll
                      for (int k = 0; k < VO; ++k)
                        {
                          double v = 0;
#pragma unroll
                          for (int l = 0; l < DIM; ++l)
                            v += g[l] * rec[C::Q_D + voigt_index<DIM>(ci, l) * VO + k];
                          T[k] = v;
                        }
#pragma unroll
                      for (int cj = 0; cj < DIM; ++cj)
#pragma unroll
                        for (int l = 0; l < DIM; ++l)
                          Kd[ci][cj] = fma(T[voigt_index<DIM>(cj, l)], g[l], Kd[ci][cj]);
                      double t = 0;
#pragma unroll
                      for (int l = 0; l < DIM; ++l)
                        t += rec[C::Q_TAU + voigt_index<DIM>(ci, l)] * g[l];
                      sg = fma(t, g[ci], sg);
                    }
                }
              double *o = sR + (a * QCH + ch) * (DIM * DIM + 1);
#pragma unroll
              for (int i = 0; i < DIM; ++i)
#pragma unroll
                for (int j = 0; j < DIM; ++j)
                  o[i * DIM + j] = Kd[i][j];
              o[DIM * DIM] = sg;
            }
          __syncthreads();
          for (int o = tid; o < NPC * DIM * DIM; o += C::NT)
            {
              const int a = o / (DIM * DIM), ij = o % (DIM * DIM);
              double    v = 0, sg = 0;
              for (int ch = 0; ch < QCH; ++ch)
                {
                  v += sR[(a * QCH + ch) * (DIM * DIM + 1) + ij];
                  sg += sR[(a * QCH + ch) * (DIM * DIM + 1) + DIM * DIM];
                }
              if (ij / DIM == ij % DIM)
                v += sg + prm.rho * prm.alpha_1 * detJ * Mref[a * NPC + a]; // :1018-1021
              diag_e[(cell * NPC + a) * (DIM * DIM) + ij] = v;
            }
        }
    }

    // node blocks: sum of the element diagonal blocks (ascending cell), Dirichlet treatment of
    // distribute_local_to_global, inversion for the block-Jacobi / Chebyshev smoothers
    template <int DIM>
    __global__ void mf_diag_kernel(const int64_t n_owned_nodes, const int npc,
                                   const int64_t *__restrict__ nc_ptr,
                                   const int32_t *__restrict__ nc_src,
                                   const uint8_t *__restrict__ constrained,
                                   const double *__restrict__ diag_e, const int kind,
                                   double *__restrict__ dinv, double *__restrict__ cdiag)
    {
      const int64_t A = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (A >= n_owned_nodes)
        return;
      double Mb[DIM][DIM], ad[DIM];
#pragma unroll
      for (int r = 0; r < DIM; ++r)
        {
          ad[r] = 0;
#pragma unroll
          for (int cc = 0; cc < DIM; ++cc)
            Mb[r][cc] = 0;
        }
      for (int64_t k = nc_ptr[A]; k < nc_ptr[A + 1]; ++k)
        {
          const int32_t src  = nc_src[k];
          const double *blk  = diag_e + int64_t(src) * (DIM * DIM);
          const int32_t cell = src / npc;
#pragma unroll
          for (int r = 0; r < DIM; ++r)
            {
#pragma unroll
              for (int cc = 0; cc < DIM; ++cc)
                Mb[r][cc] += blk[r * DIM + cc];
              double v = fabs(blk[r * DIM + r]);
              if (v == 0.0) // deal.II: the cell's average |diagonal|
                {
                  for (int i = 0; i < npc; ++i)
                    for (int d = 0; d < DIM; ++d)
                      v += fabs(diag_e[(int64_t(cell) * npc + i) * (DIM * DIM) + d * DIM + d]);
                  v /= double(npc * DIM);
                }
              ad[r] += v;
            }
        }
      bool con[DIM];
#pragma unroll
      for (int r = 0; r < DIM; ++r)
        con[r] = constrained[A * DIM + r] != 0;
#pragma unroll
      for (int r = 0; r < DIM; ++r)
#pragma unroll
        for (int cc = 0; cc < DIM; ++cc)
          if (con[r] || con[cc])
            Mb[r][cc] = (r == cc) ? ad[r] : 0.0;
#pragma unroll
      for (int r = 0; r < DIM; ++r)
        cdiag[A * DIM + r] = con[r] ? ad[r] : 0.0;
      double R[DIM][DIM];
#pragma unroll
      for (int r = 0; r < DIM; ++r)
#pragma unroll
        for (int cc = 0; cc < DIM; ++cc)
          R[r][cc] = r == cc ? 1.0 : 0.0;
      if (kind == GF_PRECOND_JACOBI)
        {
#pragma unroll
          for (int r = 0; r < DIM; ++r)
            R[r][r] = 1.0 / Mb[r][r];
        }
      else if (kind >= GF_PRECOND_BLOCK_JACOBI)
        {
#pragma unroll
          for (int r = 0; r < DIM; ++r)
#pragma unroll
            for (int cc = r + 1; cc < DIM; ++cc)
              Mb[r][cc] = Mb[cc][r] = 0.5 * (Mb[r][cc] + Mb[cc][r]);
          inverse<DIM>(Mb, det<DIM>(Mb), R);
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

    // ---------------------------------------------------------------------------------------------
    // apply (3D): (P+2)^2 threads per cell, thread = (qx, qy) column, qz loop in registers
    // ---------------------------------------------------------------------------------------------
    template <int P>
    struct ApplyCfg
    {
      static constexpr int N1 = P + 1, Q1 = P + 2, NPC = N1 * N1 * N1, DPC = 3 * NPC,
                           NQ = Q1 * Q1 * Q1, TPC = Q1 * Q1;
      static constexpr int CPB = 128 / TPC; // cells per CTA
      static constexpr int NT  = 128;
      static constexpr int SX = NPC * 3;             // x_e / y_e, lexicographic [k][j][i][c]
      static constexpr int SB = N1 * N1 * Q1 * 2 * 3; // [k][j][qx][type][c]
      static constexpr int SA = N1 * Q1 * Q1 * 3 * 3; // [k][qy][qx][type][c]
      static constexpr int CELL_D = SX + SB + SA;
      static constexpr int TAB_D  = 2 * Q1 * N1 + Q1;
      static constexpr size_t SMEM_BYTES = size_t(CPB * CELL_D + TAB_D) * sizeof(double) +
                                           NPC * sizeof(int);
    };

    template <int P>
    __global__ void __launch_bounds__(ApplyCfg<P>::NT)
      mf_apply_kernel(const int64_t n_cells, const int32_t *__restrict__ cell_nodes,
                      const double *__restrict__ geom, const uint8_t *__restrict__ constrained,
                      const double *__restrict__ qp_buf, const double *__restrict__ x,
                      const Tab1D<P> tab, const double rho_alpha1, double *__restrict__ ye,
                      const int *status)
    {
      using C = ApplyCfg<P>;
      using M = MFDim<3>;
      constexpr int N1 = C::N1, Q1 = C::Q1, NPC = C::NPC, NQ = C::NQ, TPC = C::TPC, CPB = C::CPB;
      if (status != nullptr && *status != 0)
        return;
      extern __shared__ __align__(16) double sm[];
      double *sN1 = sm, *sD1 = sm + Q1 * N1, *sw1 = sm + 2 * Q1 * N1;
      double *cells_sm = sm + C::TAB_D;
      int *   sl2h     = reinterpret_cast<int *>(cells_sm + CPB * C::CELL_D);
      const int tid = threadIdx.x;
      for (int i = tid; i < Q1 * N1; i += C::NT)
        {
          sN1[i] = tab.N[i];
          sD1[i] = tab.D[i];
        }
      if (tid < Q1)
        sw1[tid] = tab.w[tid];
      if (tid < NPC)
        sl2h[tid] = tab.lex2hier[tid];
      const int  lc = tid / TPC, t = tid % TPC; // local cell, thread in cell
      const bool thread_active = lc < CPB;
      const int  qx = t % Q1, qy = t / Q1;
      double *   sx = cells_sm + (thread_active ? lc : 0) * C::CELL_D;
      double *   sB = sx + C::SX;
      double *   sA = sB + C::SB;
      __syncthreads();

      for (int64_t cell0 = int64_t(blockIdx.x) * CPB; cell0 < n_cells;
           cell0 += int64_t(gridDim.x) * CPB)
        {
          const int64_t cell   = cell0 + lc;
          const bool    active = thread_active && cell < n_cells;
          // ---- gather x_e (lexicographic), constrained columns dropped ------------------------
          if (active)
            for (int idx = t; idx < NPC * 3; idx += TPC)
              {
                const int     l = idx / 3, cc = idx - l * 3;
                const int64_t d = int64_t(cell_nodes[cell * NPC + sl2h[l]]) * 3 + cc;
                sx[idx]         = constrained[d] ? 0.0 : x[d];
              }
          __syncthreads();
          // ---- stage X: contract i -> (k, j, qx) values and x-derivatives ----------------------
          if (active)
            for (int o = t; o < N1 * N1 * Q1 * 3; o += TPC)
              {
                const int cc = o % 3, q = (o / 3) % Q1, kj = o / (3 * Q1);
                double    vN = 0, vD = 0;
#pragma unroll
                for (int i = 0; i < N1; ++i)
                  {
                    const double xv = sx[(kj * N1 + i) * 3 + cc];
                    vN              = fma(sN1[q * N1 + i], xv, vN);
                    vD              = fma(sD1[q * N1 + i], xv, vD);
                  }
                sB[((kj * Q1 + q) * 2 + 0) * 3 + cc] = vN;
                sB[((kj * Q1 + q) * 2 + 1) * 3 + cc] = vD;
              }
          __syncthreads();
          double A0[N1][3], A1[N1][3], A2[N1][3];
          if (active)
            {
              // ---- stage Y in registers: this thread's (qx,qy) column for every k --------------
              double NN[N1][3], DN[N1][3], ND[N1][3];
#pragma unroll
              for (int k = 0; k < N1; ++k)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
                  {
                    double a = 0, b = 0, d = 0;
#pragma unroll
                    for (int j = 0; j < N1; ++j)
                      {
                        const double xn = sB[(((k * N1 + j) * Q1 + qx) * 2 + 0) * 3 + cc];
                        const double xd = sB[(((k * N1 + j) * Q1 + qx) * 2 + 1) * 3 + cc];
                        a               = fma(sN1[qy * N1 + j], xn, a);
                        b               = fma(sN1[qy * N1 + j], xd, b);
                        d               = fma(sD1[qy * N1 + j], xn, d);
                      }
                    NN[k][cc] = a;
                    DN[k][cc] = b;
                    ND[k][cc] = d;
                  }
#pragma unroll
              for (int k = 0; k < N1; ++k)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
                  A0[k][cc] = A1[k][cc] = A2[k][cc] = 0.0;
              const double  mfac = rho_alpha1 * geom[cell * 10 + 9] * sw1[qx] * sw1[qy];
              const double *qp   = qp_buf + cell * int64_t(M::NF) * NQ + qy * Q1 + qx;
#pragma unroll 1
              for (int qz = 0; qz < Q1; ++qz)
                {
                  const double *rec = qp + qz * TPC; // field f at rec[f*NQ]
                  // ---- stage Z: reference gradients and value at (qx,qy,qz) --------------------
                  double Gr[3][3], val[3];
#pragma unroll
                  for (int cc = 0; cc < 3; ++cc)
                    {
                      double v = 0, gx = 0, gy = 0, gz = 0;
#pragma unroll
                      for (int k = 0; k < N1; ++k)
                        {
                          const double nk = sN1[qz * N1 + k], dk = sD1[qz * N1 + k];
                          v               = fma(nk, NN[k][cc], v);
                          gx              = fma(nk, DN[k][cc], gx);
                          gy              = fma(nk, ND[k][cc], gy);
                          gz              = fma(dk, NN[k][cc], gz);
                        }
                      val[cc]   = v;
                      Gr[cc][0] = gx;
                      Gr[cc][1] = gy;
                      Gr[cc][2] = gz;
                    }
                  // ---- point operator -----------------------------------------------------------
                  double Cm[3][3];
#pragma unroll
                  for (int e = 0; e < 3; ++e)
#pragma unroll
                    for (int l = 0; l < 3; ++l)
                      Cm[e][l] = __ldcs(rec + int64_t(M::F_C + e * 3 + l) * NQ);
                  double H[3][3]; // spatial gradient of x
#pragma unroll
                  for (int cc = 0; cc < 3; ++cc)
#pragma unroll
                    for (int l = 0; l < 3; ++l)
                      {
                        double v = 0;
#pragma unroll
                        for (int e = 0; e < 3; ++e)
                          v = fma(Gr[cc][e], Cm[e][l], v);
                        H[cc][l] = v;
                      }
                  double eps[6], sig[6];
                  eps[0] = H[0][0];
                  eps[1] = H[1][1];
                  eps[2] = H[2][2];
                  eps[3] = H[0][1] + H[1][0];
                  eps[4] = H[0][2] + H[2][0];
                  eps[5] = H[1][2] + H[2][1];
#pragma unroll
                  for (int k = 0; k < 6; ++k)
                    sig[k] = 0.0;
#pragma unroll
                  for (int k = 0; k < 6; ++k)
#pragma unroll
                    for (int l = k; l < 6; ++l)
                      {
                        const double d = __ldcs(rec + int64_t(M::F_D + upper_index<6>(k, l)) * NQ);
                        sig[k]         = fma(d, eps[l], sig[k]);
                        if (l != k)
                          sig[l] = fma(d, eps[k], sig[l]);
                      }
                  double tau[6];
#pragma unroll
                  for (int k = 0; k < 6; ++k)
                    tau[k] = __ldcs(rec + int64_t(M::F_TAU + k) * NQ);
                  double S[3][3];
#pragma unroll
                  for (int cc = 0; cc < 3; ++cc)
#pragma unroll
                    for (int l = 0; l < 3; ++l)
                      {
                        double v = sig[voigt_index<3>(cc, l)];
#pragma unroll
                        for (int m = 0; m < 3; ++m)
                          v = fma(tau[voigt_index<3>(l, m)], H[cc][m], v);
                        S[cc][l] = v;
                      }
                  const double mw = mfac * sw1[qz];
#pragma unroll
                  for (int cc = 0; cc < 3; ++cc)
                    {
                      double Sr[3];
#pragma unroll
                      for (int e = 0; e < 3; ++e)
                        {
                          double v = 0;
#pragma unroll
                          for (int l = 0; l < 3; ++l)
                            v = fma(S[cc][l], Cm[e][l], v);
                          Sr[e] = v;
                        }
                      const double m = mw * val[cc];
                      // ---- transposed stage Z ---------------------------------------------------
#pragma unroll
                      for (int k = 0; k < N1; ++k)
                        {
                          const double nk = sN1[qz * N1 + k], dk = sD1[qz * N1 + k];
                          A0[k][cc]       = fma(nk, m, fma(dk, Sr[2], A0[k][cc]));
                          A1[k][cc]       = fma(nk, Sr[0], A1[k][cc]);
                          A2[k][cc]       = fma(nk, Sr[1], A2[k][cc]);
                        }
                    }
                }
#pragma unroll
              for (int k = 0; k < N1; ++k)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
                  {
                    double *o = sA + (((k * Q1 + qy) * Q1 + qx) * 3) * 3 + cc;
                    o[0]      = A0[k][cc];
                    o[3]      = A1[k][cc];
                    o[6]      = A2[k][cc];
                  }
            }
          __syncthreads();
          // ---- transposed stage Y: contract qy -> (k, j, qx) -----------------------------------
          if (active)
            for (int o = t; o < N1 * N1 * Q1 * 3; o += TPC)
              {
                const int cc = o % 3, q = (o / 3) % Q1, kj = o / (3 * Q1);
                const int k = kj / N1, j = kj % N1;
                double    b0 = 0, b1 = 0;
#pragma unroll
                for (int y = 0; y < Q1; ++y)
                  {
                    const double *a = sA + (((k * Q1 + y) * Q1 + q) * 3) * 3 + cc;
                    b0              = fma(sN1[y * N1 + j], a[0], fma(sD1[y * N1 + j], a[6], b0));
                    b1              = fma(sN1[y * N1 + j], a[3], b1);
                  }
                sB[((kj * Q1 + q) * 2 + 0) * 3 + cc] = b0;
                sB[((kj * Q1 + q) * 2 + 1) * 3 + cc] = b1;
              }
          __syncthreads();
          // ---- transposed stage X: contract qx -> y_e, written in FE_Q local order -------------
          if (active)
            for (int idx = t; idx < NPC * 3; idx += TPC)
              {
                const int l = idx / 3, cc = idx - l * 3;
                const int i = l % N1, kj = l / N1;
                double    v = 0;
#pragma unroll
                for (int q = 0; q < Q1; ++q)
                  v = fma(sN1[q * N1 + i], sB[((kj * Q1 + q) * 2 + 0) * 3 + cc],
                          fma(sD1[q * N1 + i], sB[((kj * Q1 + q) * 2 + 1) * 3 + cc], v));
                ye[cell * (NPC * 3) + sl2h[l] * 3 + cc] = v;
              }
          __syncthreads();
        }
    }

    // y = sum of the element results per node (ascending cell); constrained rows: cdiag * x;
    // optional fused partial sums of x . y
    template <int DIM, bool DOT>
    __global__ void __launch_bounds__(256)
      mf_gather_kernel(const int64_t n_owned_nodes, const int npc,
                       const int64_t *__restrict__ nc_ptr, const int32_t *__restrict__ nc_src,
                       const uint8_t *__restrict__ constrained, const double *__restrict__ cdiag,
                       const double *__restrict__ ye, const double *__restrict__ x,
                       double *__restrict__ y, double *__restrict__ partials, const int *status)
    {
      if (status != nullptr && *status != 0)
        return;
      __shared__ double sm[32];
      double            acc[1] = {0.0};
      for (int64_t A = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; A < n_owned_nodes;
           A += int64_t(gridDim.x) * blockDim.x)
        {
          double s[DIM];
#pragma unroll
          for (int r = 0; r < DIM; ++r)
            s[r] = 0.0;
          for (int64_t k = nc_ptr[A]; k < nc_ptr[A + 1]; ++k)
            {
              const double *e = ye + int64_t(nc_src[k]) * DIM;
#pragma unroll
              for (int r = 0; r < DIM; ++r)
                s[r] += e[r];
            }
#pragma unroll
          for (int r = 0; r < DIM; ++r)
            {
              const int64_t i  = A * DIM + r;
              const double  xi = x[i];
              const double  v  = constrained[i] ? cdiag[i] * xi : s[r];
              y[i]             = v;
              if (DOT)
                acc[0] = fma(v, xi, acc[0]);
            }
        }
      if (DOT)
        {
          block_sum<1>(acc, sm);
          if (threadIdx.x == 0)
            partials[blockIdx.x] = acc[0];
        }
    }

    template <int DIM, int P>
    void setup_t(gf_context &c, const double *u_total, const double *accel)
    {
      using C = SetupCfg<DIM, P>;
      static bool configured = false;
      if (!configured)
        {
          GF_CUDA_CHECK(cudaFuncSetAttribute(mf_setup_kernel<DIM, P>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             int(C::SMEM_BYTES)));
          configured = true;
        }
      int bps = 1;
      GF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, mf_setup_kernel<DIM, P>,
                                                                  C::NT, C::SMEM_BYTES));
      const int grid = int(std::min<int64_t>(c.n_cells, int64_t(c.sm_count) * std::max(1, bps)));
      mf_setup_kernel<DIM, P><<<grid, C::NT, C::SMEM_BYTES, c.stream>>>(
        c.n_cells, c.cell_nodes.p, c.geom.p, u_total, accel, c.tables.N.p, c.tables.dN.p,
        c.tables.w.p, c.tables.Mref.p, make_nl_params(c.desc), c.mf_qp.p, c.re_buf.p,
        c.mf_diag_e.p, c.err_flag.p);
      GF_CUDA_CHECK(cudaGetLastError());
    }

    template <int P>
    void apply_t(gf_context &c, const double *x, const int *st)
    {
      using C = ApplyCfg<P>;
      static bool configured = false;
      if (!configured)
        {
          GF_CUDA_CHECK(cudaFuncSetAttribute(mf_apply_kernel<P>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             int(C::SMEM_BYTES)));
          configured = true;
        }
      Tab1D<P> tab;
      for (int i = 0; i < (P + 2) * (P + 1); ++i)
        {
          tab.N[i] = c.tables.h1N[i];
          tab.D[i] = c.tables.h1D[i];
        }
      for (int i = 0; i < P + 2; ++i)
        tab.w[i] = c.tables.h1w[i];
      for (int i = 0; i < C::NPC; ++i)
        tab.lex2hier[i] = c.tables.lex2hier[i];
      int bps = 1;
      GF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, mf_apply_kernel<P>, C::NT,
                                                                  C::SMEM_BYTES));
      const int64_t n_groups = (c.n_cells + C::CPB - 1) / C::CPB;
      const int grid = int(std::min<int64_t>(n_groups, int64_t(c.sm_count) * std::max(1, bps)));
      const double rho_alpha1 = c.desc.rho / (c.desc.beta * c.desc.delta_t * c.desc.delta_t);
      mf_apply_kernel<P><<<grid, C::NT, C::SMEM_BYTES, c.stream>>>(
        c.n_cells, c.cell_nodes.p, c.geom.p, c.constrained.p, c.mf_qp.p, x, tab, rho_alpha1,
        c.mf_ye.p, st);
      GF_CUDA_CHECK(cudaGetLastError());
    }

    int gather_grid(const gf_context &c)
    {
      return int(
        std::max<int64_t>(1, std::min<int64_t>((c.n_owned_nodes + 255) / 256, c.max_red_blocks)));
    }
  } // namespace

  int mf_dot_partials(const gf_context &c) { return gather_grid(c); }

  // quadrature data + element results + vectors moved by one application
  double mf_bytes(const gf_context &c)
  {
    const double nf = c.dim == 3 ? MFDim<3>::NF : MFDim<2>::NF;
    return 8.0 * nf * double(c.tables.nq) * double(c.n_cells) // quadrature-point data
           + 2.0 * 8.0 * double(c.dpc) * double(c.n_cells)    // y_e written and read once
           + 8.0 * double(c.n_local) + 8.0 * double(c.n_owned); // x read, y written
  }

  void mf_setup(gf_context &c, const double *u_total, const double *accel)
  {
    GF_REQUIRE(c.model == GF_MODEL_NEO_HOOKEAN && c.dim == 3, GF_ERR_UNSUPPORTED,
               "the matrix-free operator is available for the 3D neo-Hookean model");
    if (!c.mf_qp.p)
      {
        c.mf_qp.alloc(size_t(c.n_cells) * MFDim<3>::NF * c.tables.nq);
        c.mf_ye.alloc(size_t(c.n_cells) * c.dpc);
        c.mf_diag_e.alloc(size_t(c.n_cells) * c.npc * c.dim * c.dim);
        c.mf_cdiag.alloc_zero(c.n_local, c.stream);
      }
    {
      ProfScope ps(c, Profile::ASM_CELLS);
      if (c.p == 2)
        setup_t<3, 2>(c, u_total, accel);
      else
        setup_t<3, 1>(c, u_total, accel);
    }
    {
      ProfScope      ps(c, Profile::SCATTER);
      const unsigned grid = unsigned((c.n_owned_nodes + 127) / 128);
      mf_diag_kernel<3><<<grid, 128, 0, c.stream>>>(c.n_owned_nodes, c.npc, c.nc_ptr.p,
                                                    c.nc_src.p, c.constrained.p, c.mf_diag_e.p,
                                                    c.precond, c.dinv.p, c.mf_cdiag.p);
      GF_CUDA_CHECK(cudaGetLastError());
    }
    c.mf_valid = true;
  }

  void mf_apply(gf_context &c, const double *x, double *y, double *dot_partials)
  {
    GF_REQUIRE(c.mf_valid, GF_ERR_INVALID_ARG, "matrix-free operator used before its set-up");
    ProfScope  ps(c, Profile::SPMV, 2);
    const int *st = dot_partials ? &c.cg_scalars.p->status : nullptr;
    if (c.p == 2)
      apply_t<2>(c, x, st);
    else
      apply_t<1>(c, x, st);
    const int g = gather_grid(c);
    if (dot_partials)
      mf_gather_kernel<3, true><<<g, 256, 0, c.stream>>>(
        c.n_owned_nodes, c.npc, c.nc_ptr.p, c.nc_src.p, c.constrained.p, c.mf_cdiag.p, c.mf_ye.p,
        x, y, dot_partials, st);
    else
      mf_gather_kernel<3, false><<<g, 256, 0, c.stream>>>(
        c.n_owned_nodes, c.npc, c.nc_ptr.p, c.nc_src.p, c.constrained.p, c.mf_cdiag.p, c.mf_ye.p,
        x, y, nullptr, nullptr);
    GF_CUDA_CHECK(cudaGetLastError());
  }
} // namespace gf
