// Partition-independent reductions: device pieces shared by cg.cu, vector_ops.cu, multigrid.cu,
// reduce.cu and comm.cu.
//
// Stage 1 (any kernel that needs a sum): one CTA of RED_THREADS threads per reduction CHUNK (a run
// of <= red_chunk_nodes owned nodes inside one node plane, gf_context.h). Thread t visits the
// nodes n0 + t, n0 + t + RED_THREADS, ... of its chunk and block_sum() combines the threads in a
// fixed tree, so a chunk's partial sum depends only on the chunk's content.
// Stage 2: tree_sum_1024 over the GLOBAL chunk list (all ranks' partials in slab order): thread t
// adds the chunks t, t + 1024, ... in ascending order, then the same fixed block tree. Every rank
// runs it on the same numbers => the same bits on every rank and for every number of ranks.
#pragma once
#include "gf_context.h"
#include "kernel_utils.cuh"

namespace gf
{
  constexpr int RED_THREADS  = 256;
  constexpr int TREE_THREADS = 1024;

  // sum of a[0..n) by the fixed tree; all TREE_THREADS threads of the (single) CTA call it; the
  // result is valid in thread 0. LDCG: the values may have been written by a peer over NVLink.
  __device__ __forceinline__ double tree_sum_1024(const double *a, const int n, double *smem32)
  {
    double v = 0.0;
    for (int j = threadIdx.x; j < n; j += TREE_THREADS)
      v += __ldcg(a + j);
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
      smem32[threadIdx.x >> 5] = v;
    __syncthreads();
    double w = 0.0;
    if (threadIdx.x < 32)
      w = warp_sum(smem32[threadIdx.x]);
    return w;
  }

  // SolverControl / CG scalar recurrences on the device (cg.cu explains the phases)
  __device__ __forceinline__ void cg_scalar_step(CGScalars *s, const double *sums, const int phase)
  {
    if (s->status != 0)
      return;
    if (phase == 1) // after v = A p: alpha = r.z / p.v
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
    // phase 0: after startup (sums = rr, rz) ; phase 2: after the update (sums = rr, rz)
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
    if (s->res <= s->tol) // SolverControl::check
      s->status = 1;
    else if (s->it >= s->maxit || s->res != s->res)
      s->status = 2;
  }

  // host side (reduce.cu)
  void    build_reduction_plan(gf_context &c);
  double *red_sums(gf_context &c); // device [4]: results of the last reduce_sums
  // sums[k] = tree over the global chunk list of partials[k*red_stride + chunk], k < n_sums <= 3;
  // cg_phase >= 0 additionally runs cg_scalar_step in the same launch. `status` (may be NULL):
  // skip everything once the CG has left the `iterate` state.
  void reduce_sums(gf_context &c, int n_sums, int cg_phase, bool check_status);
  // comm.cu: the cross-rank version (peer windows: one kernel; NCCL: all-gather + kernel)
  void comm_reduce_sums(gf_context &c, int n_sums, int cg_phase, bool check_status);
  // fixed-order dot product of two vectors over the owned nodes -> partials[chunk]
  void launch_dot_chunks(gf_context &c, const double *a, const double *b, bool check_status);
} // namespace gf
