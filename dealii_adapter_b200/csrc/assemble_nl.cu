// K1: batched neo-Hookean tangent + residual per cell; K2: interface traction (Neumann) term.
//
// Replaces Assembler<dim,double>::assemble_system_tangent_residual_one_cell
// (nonlinear_elasticity.cc:872-1036), the material closed forms
// (compressible_neo_hook_material.h:62-138) and assemble_neumann_contribution_one_cell
// (nonlinear_elasticity.cc:791-859). Output is the element matrix K_e (row-major dpc x dpc) and
// r_e per cell; the constrained, deterministic scatter is scatter.cu.
//
// One persistent CTA per SM, warp-specialised: PRODUCER warps run phases A and B for the chunk
// (QC quadrature points) after the one the CONSUMER warps are contracting in phase C; the T/G
// chunks travel through a 2-deep shared-memory ring guarded by mbarriers (full/empty), so the
// FP64 FMA stream of phase C is not interrupted by the latency-bound set-up phases or by
// CTA-wide barriers. Tables are staged once per CTA in shared memory.
//   phase A  one thread per quadrature point: H = grad u, F, J, F^-1, b_bar, tau, Jc (Voigt D),
//            all scaled by JxW                                              (:902-934,:958-961)
//   phase B  one thread per (q, node a): spatial gradient g_a = grad_X N_a F^-1 (:946-947),
//            T_a = B_a^T (JxW D)  (dim x V) and t_a = JxW tau g_a           (pre-contraction the
//            reference repeats per (i,j) at :1011-1012)
//   phase C  lane <-> 2x2 tile of node pairs (a, b) with b <= a ONLY - the lower triangle, as the
//            reference's j <= i loop (:1003-1035); rows u and NG-1-u of tiles share a unit so
//            all lanes carry the same load: K_ab += T_a B_b (material, :1011),
//            S_ab += t_a . g_b (geometric, :1018-1019); FP64 FMA pipe bound. The upper node
//            blocks are never computed or stored: the scatter reads K_e(b,a)^T (:1033-1035)
//   end      K_ab diag += S_ab + rho alpha_1 detJ M_ab (mass, :1020-1021); r_e from t_a, body
//            force and inertia (:984-995)
// Algorithmic flops / q-point (3D Q2): 378 node pairs (b <= a) * 30 FMA = 22.7 kflop.
#include "assemble_nl_generic.cuh"
#include "gf_context.h"
#include "kernel_utils.cuh"
#include "nl_material.cuh"

namespace gf
{
  namespace
  {
#ifndef GF_CUDA_EMULATION // tuned kernels (mbarrier pipelines, warp shuffles): hardware only; the
                          // CPU stand-in of tests/cuda_emu runs every (dim, degree) on the
                          // generic kernels below
    template <int DIM, int P>
    struct NLCfg
    {
      static constexpr int NPC  = ipow(P + 1, DIM);
      static constexpr int DPC  = NPC * DIM;
      static constexpr int NQ1  = P + 2;
      static constexpr int NQ   = ipow(NQ1, DIM);
      static constexpr int NQF  = NQ / NQ1;
      static constexpr int VO   = DIM * (DIM + 1) / 2;
      static constexpr int TS   = (DIM * VO + DIM + 1) & ~1; // T (DIM x VO) + t (DIM), even
      static constexpr int NPCP = NPC <= 4 ? 4 : (NPC <= 8 ? 8 : (NPC <= 16 ? 16 : 32));
      static constexpr int AT   = 1;                   // nodes a per lane (register tile rows)
      static constexpr int BT   = 2;                   // nodes b per lane (register tile columns)
      static constexpr int NG   = (NPC + AT - 1) / AT; // tile rows (a-groups)
      static constexpr int NBG  = (NPC + BT - 1) / BT; // b-pairs
      // Only the node blocks with b <= a are computed (the reference fills j <= i and mirrors,
      // nonlinear_elasticity.cc:1003-1035): tile row ag needs the b-pairs bg with BT*bg <= last a
      // of the row. A UNIT pairs row u with row NG-1-u so that every unit carries (nearly) the
      // same number of tiles on its LPU lanes; half the FMAs of the full matrix disappear. The
      // 1 x 2 tile (3D Q2: 196 tiles = 7 consumer warps) keeps two consumer warps per scheduler:
      // with the 2 x 2 tile (105 tiles = 4 warps) the FP64 pipe idled 64 % of the time on
      // dependent-issue latency (ncu, profiles/r02_assembly_ncu_summary.md).
      static constexpr int tiles_in_row(int ag)
      {
        return (AT * ag + AT - 1) / BT + 1 < NBG ? (AT * ag + AT - 1) / BT + 1 : NBG;
      }
      static constexpr int unit_tiles(int u)
      {
        return tiles_in_row(u) + (NG - 1 - u != u ? tiles_in_row(NG - 1 - u) : 0);
      }
      static constexpr int max_unit_tiles(int u = 0)
      {
        return u >= (NG + 1) / 2 ? 0 :
                                   (unit_tiles(u) > max_unit_tiles(u + 1) ? unit_tiles(u) :
                                                                            max_unit_tiles(u + 1));
      }
      static constexpr int NU   = (NG + 1) / 2;         // units
      static constexpr int MUT  = max_unit_tiles();
      static constexpr int LPU  = MUT <= 2 ? 2 : (MUT <= 4 ? 4 : (MUT <= 8 ? 8 : 16)); // lanes/unit
      static_assert(MUT <= 16, "tiles of a unit must fit half a warp");
      static constexpr int NSUB = 32 / LPU;             // units per consumer warp
      static constexpr int NW   = (NU + NSUB - 1) / NSUB; // consumer warps (phase C)
      // producer warps (phases A, B). ncu of the 4-warp version (profiles/r02_assembly_ncu_summary
      // .md): the consumers spent their time waiting for the `full` barrier - the producers, not
      // the FP64 contraction, bounded the kernel once only the lower triangle was computed.
      static constexpr int NPW  = (DIM == 3 && P == 2) ? 8 : 1;
      static constexpr int NT   = (NW + NPW) * 32;
      // phase A: PA threads share a quadrature point (each sums every PA-th node, then a
      // butterfly over the PA lanes), so all producer threads work on the 27-node loops
      static constexpr int PA   = (NPW * 32) / NQ >= 4 ? 4 : ((NPW * 32) / NQ >= 2 ? 2 : 1);
      static constexpr int RPT  = (DPC + NPW * 32 - 1) / (NPW * 32); // residual entries/producer
      static constexpr int QC   = (DIM == 3 && P == 2) ? 8 : NQ; // q-points per chunk
      static_assert(NBG <= 16 && BT * NBG <= NPCP + BT, "b-pair mapping");
      static constexpr int QS   = (((VO * VO + 1) & ~1) + ((VO + 1) & ~1) + DIM * DIM + DIM + 1 + 1) & ~1; // per-q scalars
      static_assert(NW * NSUB >= NU, "every unit needs its lanes");
      static_assert((TS % 2) == 0 && (NPCP % 2) == 0, "16-byte aligned rows");
      static_assert(NQ % QC == 0, "chunking");

      // shared memory (doubles)
      static constexpr int OFF_N   = 0;
      static constexpr int OFF_DN  = OFF_N + NQ * NPC;
      static constexpr int OFF_U   = OFF_DN + NQ * NPC * DIM;
      static constexpr int OFF_ACC = OFF_U + DPC;
      static constexpr int OFF_Q   = OFF_ACC + DPC;          // [NQ][QS]
      static constexpr int OFF_T   = OFF_Q + NQ * QS;            // 2 x [QC][NPC][TS]
      // ring depth: 3 buffers for 3D Q2 (218 KB of shared memory): the producers may run 24
      // q-points ahead, which covers phase A of the next cell
      static constexpr int NB      = (DIM == 3 && P == 2) ? 3 : 2;
      static constexpr int OFF_G   = OFF_T + NB * QC * NPC * TS;  // NB x [QC][DIM][NPCP]
      static constexpr int OFF_BAR = OFF_G + NB * QC * DIM * NPCP; // 2 NB mbarriers
      static constexpr int SMEM_D  = OFF_BAR + 2 * NB;
      static_assert(SMEM_D * 8 <= 227 * 1024, "shared memory budget");
      static constexpr size_t SMEM_BYTES = size_t(SMEM_D) * sizeof(double);
      // offsets inside a per-q record
      // (D and tau first: even offsets, so phase B fetches them as 16-byte loads)
      static constexpr int Q_D   = 0;                 // JxW * Jc (Voigt)           [VO*VO]
      static constexpr int Q_TAU = Q_D + ((VO * VO + 1) & ~1); // JxW * tau (Voigt)  [VO]
      static constexpr int Q_C   = Q_TAU + ((VO + 1) & ~1);    // C = Jinv * Finv    [DIM*DIM]
      static constexpr int Q_A   = Q_C + DIM * DIM;   // rho * JxW * sumN * acc     [DIM]
      static constexpr int Q_W   = Q_A + DIM;         // JxW
      static constexpr int Q_END = Q_W + 1;
    };

    template <int DIM, int P>
    __global__ void __launch_bounds__(NLCfg<DIM, P>::NT, 1)
      nl_cells_kernel(const int64_t c0, const int64_t c1, const int32_t *__restrict__ cell_nodes,
                      const double *__restrict__ geom, const double *__restrict__ u_total,
                      const double *__restrict__ accel, const double *__restrict__ tabN,
                      const double *__restrict__ tabdN, const double *__restrict__ tabw,
                      const double *__restrict__ Mref, const NLParams prm,
                      double *__restrict__ ke_buf, double *__restrict__ re_buf, int *err_flag)
    {
      using C = NLCfg<DIM, P>;
      constexpr int NPC = C::NPC, DPC = C::DPC, NQ = C::NQ, VO = C::VO, TS = C::TS, QC = C::QC,
                    QS = C::QS, NPCP = C::NPCP, AT = C::AT, BT = C::BT, NCT = C::NW * 32,
                    NPT = C::NPW * 32;
      extern __shared__ __align__(16) double sm[];
      double *sN = sm + C::OFF_N, *sdN = sm + C::OFF_DN, *su = sm + C::OFF_U,
             *sacc = sm + C::OFF_ACC, *sQ = sm + C::OFF_Q;
      uint64_t *bars = reinterpret_cast<uint64_t *>(sm + C::OFF_BAR);
      uint64_t *full = bars, *empty = bars + C::NB; // per T/G buffer
      const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
      for (int i = tid; i < NQ * NPC; i += C::NT)
        sN[i] = tabN[i];
      for (int i = tid; i < NQ * NPC * DIM; i += C::NT)
        sdN[i] = tabdN[i];
      if (tid == 0)
        {
          for (int k = 0; k < C::NB; ++k)
            {
              mbar_init(smem_u32(&full[k]), C::NPW);  // one arrive per producer warp
              mbar_init(smem_u32(&empty[k]), C::NW);  // one arrive per consumer warp
            }
          asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
      __syncthreads();

      if (warp >= C::NW)
        {
          // =============================== producer warps ======================================
          // phase A (per cell), phase B + residual (per chunk of QC q-points) into the T/G ring
          const int ptid = tid - NCT;
          int       it   = 0; // chunks produced by this CTA so far
          for (int64_t cell = c0 + blockIdx.x; cell < c1; cell += gridDim.x)
            {
              named_bar_sync(1, NPT); // previous cell's su / sQ fully consumed by the producers
              const double *gm = geom + cell * (DIM * DIM + 1);
              double        Jinv[DIM][DIM];
#pragma unroll
              for (int i = 0; i < DIM; ++i)
#pragma unroll
                for (int j = 0; j < DIM; ++j)
                  Jinv[i][j] = gm[i * DIM + j];
              const double detJ = gm[DIM * DIM];
              for (int i = ptid; i < DPC; i += NPT)
                {
                  const int32_t node = cell_nodes[cell * NPC + i / DIM];
                  su[i]              = u_total[int64_t(node) * DIM + i % DIM];
                  sacc[i]            = accel[int64_t(node) * DIM + i % DIM];
                }
              named_bar_sync(1, NPT);
              // ------------- phase A: kinematics + material per quadrature point ---------------
              // PA lanes per q-point: lane `part` sums the nodes a = part, part + PA, ...; the
              // partial sums are combined by a butterfly (fixed order), then every lane of the
              // group holds the same H / acc and evaluates the material; each writes 1/PA of the
              // record
              constexpr int PA = C::PA;
              constexpr int NQA = ((NQ * PA + 31) / 32) * 32; // whole warps take part in the shuffles
              for (int qq = ptid; qq < NQA; qq += NPT)
                {
                  const bool valid = qq < NQ * PA;
                  const int  q = valid ? qq / PA : NQ - 1, part = qq % PA;
                  double Hr[DIM][DIM], acc[DIM];
                  double sumN = 0;
#pragma unroll
                  for (int i = 0; i < DIM; ++i)
                    {
                      acc[i] = 0;
#pragma unroll
                      for (int j = 0; j < DIM; ++j)
                        Hr[i][j] = 0;
                    }
                  for (int a = part; a < NPC; a += PA)
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
#pragma unroll
                  for (int o = 1; o < PA; o <<= 1)
                    {
                      sumN += __shfl_xor_sync(0xffffffffu, sumN, o);
#pragma unroll
                      for (int i = 0; i < DIM; ++i)
                        {
                          acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
#pragma unroll
                          for (int j = 0; j < DIM; ++j)
                            Hr[i][j] += __shfl_xor_sync(0xffffffffu, Hr[i][j], o);
                        }
                    }
                  if (!valid)
                    continue;
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
                        F[i][j] = (i == j ? 1.0 : 0.0) + h; // Kinematics::F :927
                      }
                  const double detF = det<DIM>(F); // :929
                  if (!(detF > 0.0))
                    atomicExch(err_flag, 1); // Assert :935
                  double Finv[DIM][DIM];
                  inverse<DIM>(F, detF, Finv); // :934
                  const double s = pow(detF, -1.0 / DIM); // Kinematics::F_iso :930
                  double       bbar[VO];                  // Kinematics::b :932
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
                  neo_hooke<DIM>(prm.kappa, prm.mu, detF, bbar, tau, D); // :958-961
                  const double JxW = detJ * tabw[q];
                  double *     rec = sQ + q * QS;
                  // the PA lanes hold identical values; lane `part` stores every PA-th entry
#define GF_REC(idx, val)                                                                         \
  do                                                                                             \
    {                                                                                            \
      if (PA == 1 || ((idx) % PA) == part)                                                       \
        rec[idx] = (val);                                                                        \
    }                                                                                            \
  while (0)
#pragma unroll
                  for (int e = 0; e < DIM; ++e)
#pragma unroll
                    for (int l = 0; l < DIM; ++l)
                      {
                        double v = 0;
#pragma unroll
                        for (int d = 0; d < DIM; ++d)
                          v += Jinv[e][d] * Finv[d][l];
                        GF_REC(C::Q_C + e * DIM + l, v);
                      }
#pragma unroll
                  for (int k = 0; k < VO; ++k)
                    {
                      GF_REC(C::Q_TAU + k, tau[k] * JxW);
#pragma unroll
                      for (int l = 0; l < VO; ++l)
                        GF_REC(C::Q_D + k * VO + l, D[k][l] * JxW);
                    }
#pragma unroll
                  for (int cc = 0; cc < DIM; ++cc)
                    GF_REC(C::Q_A + cc, prm.rho * sumN * acc[cc] * JxW); // :993-995 summed over j
                  GF_REC(C::Q_W, JxW);
#undef GF_REC
                }
              named_bar_sync(1, NPT);

              double r_i[C::RPT]; // residual entries ptid, ptid + NPT, ... of this cell
#pragma unroll
              for (int k = 0; k < C::RPT; ++k)
                r_i[k] = 0;
              for (int qc = 0; qc < NQ; qc += QC, ++it)
                {
                  const int buf = it % C::NB;
                  double *  sT  = sm + C::OFF_T + buf * (QC * NPC * TS);
                  double *  sG  = sm + C::OFF_G + buf * (QC * DIM * NPCP);
                  if (it >= C::NB) // the consumers released this buffer (chunk it - NB)
                    mbar_wait(smem_u32(&empty[buf]), ((it / C::NB) - 1) & 1);
                  // ---------- phase B: g_a, T_a = B_a^T (JxW D), t_a = JxW tau g_a -------------
                  for (int item = ptid; item < QC * NPC; item += NPT)
                    {
                      const int     ql = item / NPC, a = item % NPC, q = qc + ql;
                      const double *rec = sQ + q * QS;
                      // every operand of the item is fetched up front (16-byte loads where the
                      // layout allows): the FMAs then run without waiting on shared memory (ncu of
                      // the previous version: 37 % of all stalls were short-scoreboard, i.e. LDS
                      // latency, in this block)
                      double Dq[VO * VO], tq[VO], Cq[DIM * DIM], dn[DIM];
                      if constexpr (((VO * VO) % 2) == 0)
                        {
#pragma unroll
                          for (int k = 0; k < VO * VO; k += 2)
                            {
                              const double2 v =
                                *reinterpret_cast<const double2 *>(rec + C::Q_D + k);
                              Dq[k]     = v.x;
                              Dq[k + 1] = v.y;
                            }
                        }
                      else
                        {
#pragma unroll
                          for (int k = 0; k < VO * VO; ++k)
                            Dq[k] = rec[C::Q_D + k];
                        }
#pragma unroll
                      for (int k = 0; k < VO; ++k)
                        tq[k] = rec[C::Q_TAU + k];
#pragma unroll
                      for (int k = 0; k < DIM * DIM; ++k)
                        Cq[k] = rec[C::Q_C + k];
#pragma unroll
                      for (int e = 0; e < DIM; ++e)
                        dn[e] = sdN[(q * NPC + a) * DIM + e];
                      double g[DIM];
#pragma unroll
                      for (int l = 0; l < DIM; ++l)
                        {
                          double v = 0;
#pragma unroll
                          for (int e = 0; e < DIM; ++e)
                            v += dn[e] * Cq[e * DIM + l];
                          g[l] = v;
                          sG[(ql * DIM + l) * NPCP + a] = v;
                        }
                      double *T = sT + (ql * NPC + a) * TS;
#pragma unroll
                      for (int ci = 0; ci < DIM; ++ci)
                        {
                          // engineering strain of dof (a,ci): eps_(ci,l) = g[l]
                          double Tr[VO];
#pragma unroll
                          for (int k = 0; k < VO; ++k)
                            {
                              double v = 0;
#pragma unroll
                              for (int l = 0; l < DIM; ++l)
                                v += g[l] * Dq[voigt_index<DIM>(ci, l) * VO + k];
                              Tr[k] = v;
                            }
                          if constexpr ((VO % 2) == 0)
                            {
#pragma unroll
                              for (int k = 0; k < VO; k += 2)
                                *reinterpret_cast<double2 *>(T + ci * VO + k) =
                                  make_double2(Tr[k], Tr[k + 1]);
                            }
                          else
                            {
#pragma unroll
                              for (int k = 0; k < VO; ++k)
                                T[ci * VO + k] = Tr[k];
                            }
                          double t = 0;
#pragma unroll
                          for (int l = 0; l < DIM; ++l)
                            t += tq[voigt_index<DIM>(ci, l)] * g[l];
                          T[DIM * VO + ci] = t;
                        }
                    }
                  __syncwarp();
                  if (lane == 0)
                    mbar_arrive(smem_u32(&full[buf])); // release: this warp's part of the chunk
                  named_bar_sync(1, NPT);              // all producer warps wrote the chunk
                  // ---------- residual (:984-995), fixed q order ----------------------------------
#pragma unroll
                  for (int k = 0; k < C::RPT; ++k)
                    {
                      const int i = ptid + k * NPT;
                      if (i < DPC)
                        {
                          const int a = i / DIM, cc = i % DIM;
                          double    r = r_i[k];
                          for (int ql = 0; ql < QC; ++ql)
                            {
                              const int     q   = qc + ql;
                              const double *rec = sQ + q * QS;
                              const double  Na  = sN[q * NPC + a];
                              r -= (sT[(ql * NPC + a) * TS + DIM * VO + cc] -
                                    prm.body_force[cc] * prm.rho * Na * rec[C::Q_W]);
                              r -= Na * rec[C::Q_A + cc];
                            }
                          r_i[k] = r;
                        }
                    }
                }
#pragma unroll
              for (int k = 0; k < C::RPT; ++k)
                if (ptid + k * NPT < DPC)
                  re_buf[cell * DPC + ptid + k * NPT] = r_i[k];
            }
        }
      else
        {
          // =============================== consumer warps ======================================
          // phase C: K_ab += T_a B_b (material, :1011) ; S_ab += t_a . g_b (geometric, :1018-1019)
          // Register tile: every lane owns AT nodes a x BT nodes b (2 x 2 x dim x dim accumulators;
          // the shared-memory traffic per FMA depends on BT only, and the small AT leaves room for two
          // consumer warps per scheduler, which hides the dependent-DFMA latency).
          // unit = (warp, sub-warp) <-> group of AT nodes a, lane-in-sub <-> pair of nodes b: the
          // T_a rows are broadcast loads inside a sub-warp and serve AT*BT pairs per lane, the
          // g_b pair is one 16-byte load - shared-memory traffic per FMA is half that of a
          // 3 x 1 tile, which left the kernel bound by the LDS pipe (ncu: 76 % LSU, 39 % FP64)
          // lane j of unit u: the tiles of row u first, then those of row NG-1-u; the middle row
          // of an odd NG is its own partner and is taken once
          const int  sub = lane / C::LPU, j = lane % C::LPU;
          const int  unit = warp * C::NSUB + sub;
          const int  row2 = C::NG - 1 - unit;
          const int  n_first = unit < C::NU ? C::tiles_in_row(unit) : 0;
          const bool first_row = j < n_first;
          const int  ag = first_row ? unit : row2;
          const int  bg = first_row ? j : j - n_first;
          const bool unit_active = unit < C::NU && (first_row || (row2 != unit && row2 >= 0 &&
                                                                   bg < C::tiles_in_row(row2)));
          const int  a_base = unit_active ? ag * AT : 0;
          const int  bgc = unit_active ? bg : 0; // clamped for loads
          int        it = 0;
          for (int64_t cell = c0 + blockIdx.x; cell < c1; cell += gridDim.x)
            {
              double Kacc[AT][BT][DIM][DIM], Sacc[AT][BT];
#pragma unroll
              for (int ai = 0; ai < AT; ++ai)
#pragma unroll
                for (int bi = 0; bi < BT; ++bi)
                  {
                    Sacc[ai][bi] = 0;
#pragma unroll
                    for (int i = 0; i < DIM; ++i)
#pragma unroll
                      for (int j = 0; j < DIM; ++j)
                        Kacc[ai][bi][i][j] = 0;
                  }
              for (int qc = 0; qc < NQ; qc += QC, ++it)
                {
                  const int     buf = it % C::NB;
                  const double *sT  = sm + C::OFF_T + buf * (QC * NPC * TS);
                  const double *sG  = sm + C::OFF_G + buf * (QC * DIM * NPCP);
                  mbar_wait(smem_u32(&full[buf]), (it / C::NB) & 1);
                  if (unit_active)
                    {
#pragma unroll 2
                      for (int ql = 0; ql < QC; ++ql)
                        {
                          double gb[BT][DIM];
#pragma unroll
                          for (int l = 0; l < DIM; ++l)
                            {
                              static_assert(BT == 2, "g_b pairs are fetched as one double2");
                              const double2 v = *reinterpret_cast<const double2 *>(
                                sG + (ql * DIM + l) * NPCP + BT * bgc);
                              gb[0][l] = v.x;
                              gb[1][l] = v.y;
                            }
#pragma unroll
                          for (int ai = 0; ai < AT; ++ai)
                            {
                              const int     a = min(a_base + ai, NPC - 1);
                              const double *T = sT + (ql * NPC + a) * TS;
#pragma unroll
                              for (int ci = 0; ci < DIM; ++ci)
                                {
                                  double Tr[VO]; // row ci of T_a
                                  if constexpr ((VO % 2) == 0)
                                    {
#pragma unroll
                                      for (int k = 0; k < VO; k += 2)
                                        {
                                          const double2 v =
                                            *reinterpret_cast<const double2 *>(T + ci * VO + k);
                                          Tr[k]     = v.x;
                                          Tr[k + 1] = v.y;
                                        }
                                    }
                                  else
                                    {
#pragma unroll
                                      for (int k = 0; k < VO; ++k)
                                        Tr[k] = T[ci * VO + k];
                                    }
#pragma unroll
                                  for (int bi = 0; bi < BT; ++bi)
#pragma unroll
                                    for (int cj = 0; cj < DIM; ++cj)
                                      {
                                        double v = Kacc[ai][bi][ci][cj];
#pragma unroll
                                        for (int l = 0; l < DIM; ++l)
                                          v = fma(Tr[voigt_index<DIM>(cj, l)], gb[bi][l], v);
                                        Kacc[ai][bi][ci][cj] = v;
                                      }
                                }
                              double tl[DIM];
#pragma unroll
                              for (int l = 0; l < DIM; ++l)
                                tl[l] = T[DIM * VO + l];
#pragma unroll
                              for (int bi = 0; bi < BT; ++bi)
                                {
                                  double sv = Sacc[ai][bi];
#pragma unroll
                                  for (int l = 0; l < DIM; ++l)
                                    sv = fma(tl[l], gb[bi][l], sv);
                                  Sacc[ai][bi] = sv;
                                }
                            }
                        }
                    }
                  __syncwarp();
                  if (lane == 0)
                    mbar_arrive(smem_u32(&empty[buf])); // this warp is done with the buffer
                }
              // ------------- write K_e (row-major); mass term :1020-1021 ------------------------
              if (unit_active)
                {
                  const double detJ = geom[cell * (DIM * DIM + 1) + DIM * DIM];
                  const double mfac = prm.rho * prm.alpha_1 * detJ;
                  double *     ke   = ke_buf + (cell - c0) * int64_t(DPC) * DPC;
#pragma unroll
                  for (int ai = 0; ai < AT; ++ai)
                    {
                      const int a = a_base + ai;
                      if (a < NPC)
                        {
#pragma unroll
                          for (int bi = 0; bi < BT; ++bi)
                            {
                              const int b = BT * bg + bi;
                              if (b <= a) // lower node blocks only; the scatter mirrors them
                                {
                                  const double dd = Sacc[ai][bi] + mfac * Mref[a * NPC + b];
#pragma unroll
                                  for (int ci = 0; ci < DIM; ++ci)
#pragma unroll
                                    for (int cj = 0; cj < DIM; ++cj)
                                      ke[(a * DIM + ci) * DPC + b * DIM + cj] =
                                        Kacc[ai][bi][ci][cj] + (ci == cj ? dd : 0.0);
                                }
                            }
                        }
                    }
                }
            }
        }
    }

    // K2: one CTA per cell that owns interface faces; faces in ascending face number
    // (cell->face_iterators(), :804). Adds into r_e after the cell kernel (:755-756 order).
    template <int DIM, int P>
    __global__ void nl_faces_kernel(const int n_iface_cells, const int32_t *__restrict__ cell_list,
                                    const int32_t *__restrict__ face_ptr,
                                    const int32_t *__restrict__ face_no,
                                    const int32_t *__restrict__ cell_nodes,
                                    const double *__restrict__ geom,
                                    const double *__restrict__ u_total,
                                    const double *__restrict__ stress,
                                    const double *__restrict__ tabdN,
                                    const double *__restrict__ tabNf,
                                    const double *__restrict__ tabwf, double *__restrict__ re_buf,
                                    int *err_flag)
    {
      using C = NLCfg<DIM, P>;
      constexpr int NPC = C::NPC, DPC = C::DPC, NQF = C::NQF;
      __shared__ double su[DPC], ss[DPC];
      __shared__ double sfac[NQF];          // ||det F F^-T N|| at "face" q (cell q index!)
      __shared__ double strac[NQF][DIM];    // local_stress(q) * factor * JxW_f
      const int ic = blockIdx.x;
      if (ic >= n_iface_cells)
        return;
      const int     tid  = threadIdx.x;
      const int64_t cell = cell_list[ic];
      const double *gm   = geom + cell * (DIM * DIM + 1);
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
          ss[tid]            = stress[int64_t(node) * DIM + tid % DIM];
        }
      __syncthreads();
      double r_add = 0;
      for (int fi = face_ptr[ic]; fi < face_ptr[ic + 1]; ++fi)
        {
          const int face = face_no[fi];
          const int fd = face / 2;
          // n da = det(J) J^-T n_ref dA_ref  (affine cell: constant per face)
          double nrm[DIM], len = 0;
#pragma unroll
          for (int i = 0; i < DIM; ++i)
            {
              nrm[i] = detJ * Jinv[fd][i] * ((face & 1) ? 1.0 : -1.0);
              len += nrm[i] * nrm[i];
            }
          len = sqrt(len);
#pragma unroll
          for (int i = 0; i < DIM; ++i)
            nrm[i] /= len;
          if (tid < NQF)
            {
              const int q = tid; // :825-827 — CELL quadrature gradient at index f_q_point
              double    Hr[DIM][DIM];
#pragma unroll
              for (int i = 0; i < DIM; ++i)
#pragma unroll
                for (int j = 0; j < DIM; ++j)
                  Hr[i][j] = 0;
              for (int a = 0; a < NPC; ++a)
#pragma unroll
                for (int cc = 0; cc < DIM; ++cc)
#pragma unroll
                  for (int e = 0; e < DIM; ++e)
                    Hr[cc][e] += su[a * DIM + cc] * tabdN[(q * NPC + a) * DIM + e];
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
                atomicExch(err_flag, 1);
              double Finv[DIM][DIM];
              inverse<DIM>(F, detF, Finv);
              // n_star = det F F^-T N  (:831-833)
              double n2 = 0;
#pragma unroll
              for (int i = 0; i < DIM; ++i)
                {
                  double v = 0;
#pragma unroll
                  for (int j = 0; j < DIM; ++j)
                    v += (detF * Finv[j][i]) * nrm[j];
                  n2 += v * v;
                }
              const double fac = sqrt(n2);
              // local_stress(q) = sum_k N_k(face q) sigma_k (:815-816)
              double t[DIM];
#pragma unroll
              for (int cc = 0; cc < DIM; ++cc)
                t[cc] = 0;
              for (int a = 0; a < NPC; ++a)
                {
                  const double Na = tabNf[(face * NQF + q) * NPC + a];
#pragma unroll
                  for (int cc = 0; cc < DIM; ++cc)
                    t[cc] += ss[a * DIM + cc] * Na;
                }
              sfac[q] = len * tabwf[q]; // JxW on the face
#pragma unroll
              for (int cc = 0; cc < DIM; ++cc)
                strac[q][cc] = t[cc] * fac; // referential_stress :836-837
            }
          __syncthreads();
          if (tid < DPC)
            {
              const int a = tid / DIM, cc = tid % DIM;
              for (int q = 0; q < NQF; ++q) // :853-854
                r_add += (tabNf[(face * NQF + q) * NPC + a] * strac[q][cc]) * sfac[q];
            }
          __syncthreads();
        }
      if (tid < DPC)
        re_buf[cell * DPC + tid] += r_add;
    }

    template <int DIM, int P>
    void launch_cells_t(gf_context &c, const double *u_total, const double *accel, int64_t c0,
                        int64_t c1)
    {
      using C = NLCfg<DIM, P>;
      static bool configured = false;
      if (!configured)
        {
          GF_CUDA_CHECK(cudaFuncSetAttribute(nl_cells_kernel<DIM, P>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             int(C::SMEM_BYTES)));
          configured = true;
        }
      const NLParams prm = make_nl_params(c.desc);
      int blocks_per_sm = 1;
      GF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &blocks_per_sm, nl_cells_kernel<DIM, P>, C::NT, C::SMEM_BYTES));
      const int64_t n    = c1 - c0;
      const int     grid = int(std::min<int64_t>(n, int64_t(c.sm_count) * std::max(1, blocks_per_sm)));
      if (grid <= 0)
        return;
      nl_cells_kernel<DIM, P><<<grid, C::NT, C::SMEM_BYTES, c.stream>>>(
        c0, c1, c.cell_nodes.p, c.geom.p, u_total, accel, c.tables.N.p, c.tables.dN.p,
        c.tables.w.p, c.tables.Mref.p, prm, c.ke_buf.p, c.re_buf.p, c.err_flag.p);
      GF_CUDA_CHECK(cudaGetLastError());
    }

    template <int DIM, int P>
    void launch_faces_t(gf_context &c, const double *u_total, const double *stress)
    {
      using C = NLCfg<DIM, P>;
      if (c.n_iface_cells == 0)
        return;
      const int nt = ((C::DPC + 31) / 32) * 32;
      nl_faces_kernel<DIM, P><<<unsigned(c.n_iface_cells), nt, 0, c.stream>>>(
        int(c.n_iface_cells), c.iface_cell_list.p, c.iface_face_ptr.p, c.iface_face_no.p,
        c.cell_nodes.p, c.geom.p, u_total, stress, c.tables.dN.p, c.tables.Nf.p, c.tables.wf.p,
        c.re_buf.p, c.err_flag.p);
      GF_CUDA_CHECK(cudaGetLastError());
    }
#endif // GF_CUDA_EMULATION
    // generic kernels (assemble_nl_generic.cuh): every (dim, degree) without a tuned
    // instantiation, i.e. degree >= 3, and every mesh with non-affine cells (AFFINE = false)
    template <int DIM, bool AFFINE>
    void launch_cells_generic(gf_context &c, const double *u_total, const double *accel, int64_t c0,
                              int64_t c1)
    {
      const size_t  smem = size_t(NLGen<DIM>::smem_doubles(c.npc)) * sizeof(double);
      static size_t configured = 48 * 1024;
      if (smem > configured)
        {
          GF_REQUIRE(smem <= 227 * 1024, GF_ERR_UNSUPPORTED,
                     "polynomial degree too high for the generic cell kernel's shared memory");
          GF_CUDA_CHECK(cudaFuncSetAttribute(nl_cells_generic_kernel<DIM, AFFINE>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
          configured = smem;
        }
      const int64_t n = c1 - c0;
      if (n <= 0)
        return;
      const int     nt   = 256;
      const int     grid = int(std::min<int64_t>(n, int64_t(c.sm_count) * 2));
      const NLParams prm = make_nl_params(c.desc);
      nl_cells_generic_kernel<DIM, AFFINE><<<grid, nt, smem, c.stream>>>(
        c0, c1, c.npc, c.tables.nq, c.cell_nodes.p, c.geom.p, c.cell_verts.p, c.tables.dphi.p,
        u_total, accel, c.tables.N.p, c.tables.dN.p, c.tables.w.p, c.tables.Mref.p, prm,
        c.ke_buf.p, c.re_buf.p, c.err_flag.p);
      GF_CUDA_CHECK(cudaGetLastError());
    }

    template <int DIM, bool AFFINE>
    void launch_faces_generic(gf_context &c, const double *u_total, const double *stress)
    {
      if (c.n_iface_cells == 0)
        return;
      const size_t smem =
        size_t(nl_faces_generic_smem_doubles<DIM>(c.npc, c.tables.nqf)) * sizeof(double);
      GF_REQUIRE(smem <= 48 * 1024, GF_ERR_UNSUPPORTED,
                 "polynomial degree too high for the generic face kernel's shared memory");
      nl_faces_generic_kernel<DIM, AFFINE><<<unsigned(c.n_iface_cells), 128, smem, c.stream>>>(
        int(c.n_iface_cells), c.npc, c.tables.nqf, c.iface_cell_list.p, c.iface_face_ptr.p,
        c.iface_face_no.p, c.cell_nodes.p, c.geom.p, c.cell_verts.p, c.tables.dphi.p,
        c.tables.dphif.p, u_total, stress, c.tables.dN.p, c.tables.Nf.p, c.tables.wf.p, c.re_buf.p,
        c.err_flag.p);
      GF_CUDA_CHECK(cudaGetLastError());
    }
  } // namespace

  void launch_nl_cells(gf_context &c, const double *u_total, const double *accel, int64_t c0,
                       int64_t c1)
  {
    ProfScope ps(c, Profile::ASM_CELLS);
    if (!c.affine)
      {
        if (c.dim == 3)
          launch_cells_generic<3, false>(c, u_total, accel, c0, c1);
        else
          launch_cells_generic<2, false>(c, u_total, accel, c0, c1);
      }
#ifndef GF_CUDA_EMULATION
    else if (c.dim == 3 && c.p == 2)
      launch_cells_t<3, 2>(c, u_total, accel, c0, c1);
    else if (c.dim == 3 && c.p == 1)
      launch_cells_t<3, 1>(c, u_total, accel, c0, c1);
    else if (c.dim == 2 && c.p == 2)
      launch_cells_t<2, 2>(c, u_total, accel, c0, c1);
    else if (c.dim == 2 && c.p == 1)
      launch_cells_t<2, 1>(c, u_total, accel, c0, c1);
#endif
    else if (c.dim == 3)
      launch_cells_generic<3, true>(c, u_total, accel, c0, c1);
    else
      launch_cells_generic<2, true>(c, u_total, accel, c0, c1);
  }

  void launch_nl_faces(gf_context &c, const double *u_total, const double *stress)
  {
    ProfScope ps(c, Profile::ASM_FACES);
    if (!c.affine)
      {
        if (c.dim == 3)
          launch_faces_generic<3, false>(c, u_total, stress);
        else
          launch_faces_generic<2, false>(c, u_total, stress);
      }
#ifndef GF_CUDA_EMULATION
    else if (c.dim == 3 && c.p == 2)
      launch_faces_t<3, 2>(c, u_total, stress);
    else if (c.dim == 3 && c.p == 1)
      launch_faces_t<3, 1>(c, u_total, stress);
    else if (c.dim == 2 && c.p == 2)
      launch_faces_t<2, 2>(c, u_total, stress);
    else if (c.dim == 2 && c.p == 1)
      launch_faces_t<2, 1>(c, u_total, stress);
#endif
    else if (c.dim == 3)
      launch_faces_generic<3, true>(c, u_total, stress);
    else
      launch_faces_generic<2, true>(c, u_total, stress);
  }
} // namespace gf
