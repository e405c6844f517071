// Geometric multigrid preconditioner for the device CG (K6'): replaces the serial SSOR sweeps of
// the reference (PreconditionSelector("ssor",.65), nonlinear_elasticity.cc:1180-1182;
// PreconditionSSOR(1.2), linear_elasticity.cc:548-549) by a V-cycle over the global-refinement
// hierarchy the host already owns (triangulation.refine_global, nonlinear_elasticity.cc:245-246,
// linear_elasticity.cc:150-151). Block-Jacobi CG needs ~2,400 iterations per Newton solve on the
// 2 M-DoF flap; the V-cycle makes the count mesh independent (~20).
//
//   levels     every level is an ordinary handle (gf_create on that level's cells); the host links
//              them with gf_mg_attach(fine, coarse, child_cells)
//   operators  re-discretised on every level by the SAME assembly kernels (K1/K10 + scatter):
//              nonlinear tangent at the injected state u_l (nested nodes), linear M + th^2 dt^2 K
//   smoother   Chebyshev polynomial of block-Jacobi, degree 3 on [lmax/40, 1.2 lmax]; lmax by a
//              warm-started power iteration after every assembly; coarsest level: degree 80 on
//              [lmax/1000, 1.2 lmax]. Polynomial smoothers are symmetric => the V-cycle is SPD
//              and plain CG applies. Only SpMV (K4) + streaming vector kernels: HBM-bound.
//   transfer   matrix-free FE embedding: prolongation evaluates the coarse shape functions at the
//              child-cell nodes; restriction is its transpose in gather form (fixed summation
//              order, no FP atomics => reproducible).
//   multi-GPU  partitioned levels have their own slab partition and halo lists; restriction sums
//              over owned fine nodes and sends the ghost partial sums to the owner
//              (halo_reduce_add). Small levels are REPLICATED: every rank holds the whole coarse
//              mesh, restriction and state injection end in one all-reduce over the coarse vector
//              and everything below (smoothing, the 80-step coarsest solve, re-discretisation)
//              runs redundantly with no communication at all.
#include <cmath>

#include "reduce.cuh"

namespace gf
{
  namespace
  {
    constexpr int NT = 256;

    double lagrange1d(int p, int i, double x)
    {
      if (p == 1)
        return i == 0 ? 1.0 - x : x;
      return i == 0 ? 2.0 * (x - 0.5) * (x - 1.0) :
                      (i == 1 ? -4.0 * x * (x - 1.0) : 2.0 * x * (x - 0.5));
    }

    inline unsigned blocks_for(int64_t n, int per_block = NT)
    {
      return unsigned(std::max<int64_t>(1, (n + per_block - 1) / per_block));
    }

    // first_owner[cell*npc + a] = 1 iff (cell, a) is the first appearance of its node in the
    // node -> cell adjacency and the node is owned
    __global__ void first_owner_kernel(const int64_t n_cn, const int64_t n_owned_nodes,
                                       const int32_t *__restrict__ cell_nodes,
                                       const int64_t *__restrict__ nc_ptr,
                                       const int32_t *__restrict__ nc_src,
                                       uint8_t *__restrict__ flag)
    {
      const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (i >= n_cn)
        return;
      const int32_t node = cell_nodes[i];
      flag[i]            = (node < n_owned_nodes && int64_t(nc_src[nc_ptr[node]]) == i) ? 1 : 0;
    }

    // coarse state <- fine state at the coinciding (nested) nodes
    template <int DIM>
    __global__ void inject_kernel(const int64_t n_coarse_nodes, const int npc, const int n_child,
                                  const int64_t *__restrict__ nc_ptr_c,
                                  const int32_t *__restrict__ nc_src_c,
                                  const int32_t *__restrict__ child_cells,
                                  const int32_t *__restrict__ inj,
                                  const int32_t *__restrict__ cell_nodes_f,
                                  const int64_t n_owned_nodes_f, const bool owned_only,
                                  const double *__restrict__ u_f, double *__restrict__ u_c)
    {
      const int64_t B = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (B >= n_coarse_nodes)
        return;
      if (owned_only) // replicated coarse level: exactly one rank contributes, the others add 0
#pragma unroll
        for (int cc = 0; cc < DIM; ++cc)
          u_c[B * DIM + cc] = 0.0;
      for (int64_t s = nc_ptr_c[B]; s < nc_ptr_c[B + 1]; ++s)
        {
          const int32_t src = nc_src_c[s];
          const int32_t pc = src / npc, b = src - pc * npc;
          const int32_t ka = inj[b];
          const int32_t fc = child_cells[int64_t(pc) * n_child + ka / npc];
          if (fc < 0)
            continue;
          const int64_t nf = cell_nodes_f[int64_t(fc) * npc + ka % npc];
          if (owned_only && nf >= n_owned_nodes_f)
            return; // the same fine node in every adjacent cell: its owner injects it
#pragma unroll
          for (int cc = 0; cc < DIM; ++cc)
            u_c[B * DIM + cc] = u_f[nf * DIM + cc];
          return;
        }
      // no local child (ghost of a ghost): filled by the halo exchange of the coarse level
    }

    // x_f += P x_c : fine node value = sum_b N_b^coarse(node) x_c[b], evaluated in the node's
    // first cell; constrained fine dofs stay untouched
    template <int DIM>
    __global__ void prolong_add_kernel(const int64_t n_fine_nodes, const int npc, const int n_child,
                                       const int64_t *__restrict__ nc_ptr_f,
                                       const int32_t *__restrict__ nc_src_f,
                                       const int32_t *__restrict__ parent,
                                       const double *__restrict__ E,
                                       const int32_t *__restrict__ cell_nodes_c,
                                       const uint8_t *__restrict__ constrained_f,
                                       const double *__restrict__ x_c, double *__restrict__ x_f)
    {
      const int64_t A = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (A >= n_fine_nodes)
        return;
      const int32_t src  = nc_src_f[nc_ptr_f[A]];
      const int32_t cell = src / npc, a = src - cell * npc;
      const int32_t par  = parent[cell];
      const int32_t pc = par / n_child, k = par - pc * n_child;
      const double *Ek = E + (int64_t(k) * npc + a) * npc;
      double        acc[DIM];
#pragma unroll
      for (int cc = 0; cc < DIM; ++cc)
        acc[cc] = 0.0;
      for (int b = 0; b < npc; ++b)
        {
          const double w = Ek[b];
          if (w != 0.0)
            {
              const int64_t nb = cell_nodes_c[int64_t(pc) * npc + b];
#pragma unroll
              for (int cc = 0; cc < DIM; ++cc)
                acc[cc] = fma(w, x_c[nb * DIM + cc], acc[cc]);
            }
        }
#pragma unroll
      for (int cc = 0; cc < DIM; ++cc)
        if (!constrained_f[A * DIM + cc])
          x_f[A * DIM + cc] += acc[cc];
    }

    // b_c = P^T r_f in gather form: one warp per coarse node; lanes sweep the non-zero
    // (child, fine local node) pairs of each parent cell around the node; fixed order.
    // Partition independence: the contributions of fine nodes BEYOND the coarse node's plane along
    // the slab axis (bit 30 of rl_ka) are summed separately from the rest and added last. At a
    // slab cut those are exactly the fine nodes the upper rank owns, so the owner's
    // "own part + received part" is the same two numbers added in the same order as here.
    template <int DIM>
    __global__ void __launch_bounds__(NT)
      restrict_kernel(const int64_t n_coarse_nodes, const int npc, const int n_child,
                      const int64_t *__restrict__ nc_ptr_c, const int32_t *__restrict__ nc_src_c,
                      const int32_t *__restrict__ child_cells, const int32_t *__restrict__ rl_ptr,
                      const int32_t *__restrict__ rl_ka, const double *__restrict__ rl_w,
                      const int32_t *__restrict__ cell_nodes_f,
                      const uint8_t *__restrict__ first_owner,
                      const uint8_t *__restrict__ constrained_c, const bool mask,
                      const double *__restrict__ r_f, double *__restrict__ b_c)
    {
      const int     lane = threadIdx.x & 31;
      const int64_t B    = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
      if (B >= n_coarse_nodes)
        return;
      double acc[DIM], acc_hi[DIM];
#pragma unroll
      for (int cc = 0; cc < DIM; ++cc)
        acc[cc] = acc_hi[cc] = 0.0;
      for (int64_t s = nc_ptr_c[B]; s < nc_ptr_c[B + 1]; ++s)
        {
          const int32_t src = nc_src_c[s];
          const int32_t pc = src / npc, b = src - pc * npc;
          const int     l0 = rl_ptr[b], l1 = rl_ptr[b + 1];
          for (int l = l0 + lane; l < l1; l += 32)
            {
              const int32_t kah = rl_ka[l];
              const bool    hi  = (kah >> 30) & 1;
              const int32_t ka  = kah & 0x3fffffff;
              const int32_t fc = child_cells[int64_t(pc) * n_child + ka / npc];
              if (fc < 0)
                continue;
              const int64_t fi = int64_t(fc) * npc + ka % npc;
              if (!first_owner[fi])
                continue;
              const int64_t nf = cell_nodes_f[fi];
              const double  w  = rl_w[l];
              if (hi)
                {
#pragma unroll
                  for (int cc = 0; cc < DIM; ++cc)
                    acc_hi[cc] = fma(w, r_f[nf * DIM + cc], acc_hi[cc]);
                }
              else
                {
#pragma unroll
                  for (int cc = 0; cc < DIM; ++cc)
                    acc[cc] = fma(w, r_f[nf * DIM + cc], acc[cc]);
                }
            }
        }
#pragma unroll
      for (int cc = 0; cc < DIM; ++cc)
        acc[cc] = warp_sum(acc[cc]) + warp_sum(acc_hi[cc]);
      if (lane == 0)
#pragma unroll
        for (int cc = 0; cc < DIM; ++cc)
          b_c[B * DIM + cc] = (mask && constrained_c[B * DIM + cc]) ? 0.0 : acc[cc];
    }

    // one Chebyshev step on node blocks:
    //   r = r_in - v (v optional) ; d = c_d d + c_r D^-1 r ; x = (x_zero ? 0 : x) + d
    // c_r = c_r_unit / lmax with lmax read from device memory: the eigenvalue estimate never
    // travels to the host (no stream synchronisation per level and assembly)
    template <int DIM>
    __global__ void __launch_bounds__(NT)
      cheb_step_kernel(const int64_t n_nodes, const double *__restrict__ r_in,
                       const double *__restrict__ v, const double *__restrict__ dinv,
                       const double c_d, const double c_r_unit, const double *__restrict__ lmax,
                       const bool x_zero, double *__restrict__ r_out, double *__restrict__ d,
                       double *__restrict__ x)
    {
      const double c_r = c_r_unit / *lmax;
      for (int64_t A = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; A < n_nodes;
           A += int64_t(gridDim.x) * blockDim.x)
        {
          double rl[DIM], z[DIM];
#pragma unroll
          for (int i = 0; i < DIM; ++i)
            {
              rl[i] = r_in[A * DIM + i];
              if (v != nullptr)
                rl[i] -= v[A * DIM + i];
              r_out[A * DIM + i] = rl[i];
            }
#pragma unroll
          for (int i = 0; i < DIM; ++i)
            {
              double s = 0;
#pragma unroll
              for (int j = 0; j < DIM; ++j)
                s = fma(dinv[A * DIM * DIM + i * DIM + j], rl[j], s);
              z[i] = s;
            }
#pragma unroll
          for (int i = 0; i < DIM; ++i)
            {
              const int64_t k  = A * DIM + i;
              double        dn = c_r * z[i];
              if (c_d != 0.0)
                dn = fma(c_d, d[k], dn);
              d[k] = dn;
              x[k] = x_zero ? dn : x[k] + dn;
            }
        }
    }

    // out = a - b
    __global__ void sub_kernel(const int64_t n, const double *__restrict__ a,
                               const double *__restrict__ b, double *__restrict__ out)
    {
      for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n;
           i += int64_t(gridDim.x) * blockDim.x)
        out[i] = a[i] - b[i];
    }

    // power iteration pieces: t = D^-1 v with partial sums of t.t ; e = t / sqrt(sum)
    template <int DIM>
    __global__ void __launch_bounds__(RED_THREADS)
      pw_apply_kernel(const int32_t *__restrict__ chunk_ptr, const double *__restrict__ v,
                      const double *__restrict__ dinv, double *__restrict__ t,
                      double *__restrict__ partials)
    {
      __shared__ double sm[32];
      double            acc[1] = {0.0};
      const int64_t     n0 = chunk_ptr[blockIdx.x], n1 = chunk_ptr[blockIdx.x + 1];
      for (int64_t A = n0 + threadIdx.x; A < n1; A += RED_THREADS)
        {
#pragma unroll
          for (int i = 0; i < DIM; ++i)
            {
              double s = 0;
#pragma unroll
              for (int j = 0; j < DIM; ++j)
                s = fma(dinv[A * DIM * DIM + i * DIM + j], v[A * DIM + j], s);
              t[A * DIM + i] = s;
              acc[0]         = fma(s, s, acc[0]);
            }
        }
      block_sum<1>(acc, sm);
      if (threadIdx.x == 0)
        partials[blockIdx.x] = acc[0];
    }
    // lmax = 1.2 * sqrt(||D^-1 A e||^2): the power iteration converges from below, safety factor
    // as in deal.II's Chebyshev preconditioner
    __global__ void lmax_store_kernel(const double *__restrict__ sumsq, double *lmax)
    {
      *lmax = 1.2 * sqrt(sumsq[0]);
    }
    __global__ void pw_scale_kernel(const int64_t n, const double *__restrict__ t,
                                    const double *__restrict__ sumsq, double *__restrict__ e)
    {
      const double s = 1.0 / sqrt(sumsq[0]);
      for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n;
           i += int64_t(gridDim.x) * blockDim.x)
        e[i] = t[i] * s;
    }
    // deterministic start vector (splitmix64 of the partition-independent id of the node and the
    // component), zero on constrained dofs
    __global__ void pw_init_kernel(const int64_t n, const int dim,
                                   const int64_t *__restrict__ node_gkey,
                                   const uint8_t *__restrict__ constrained, double *__restrict__ e)
    {
      for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n;
           i += int64_t(gridDim.x) * blockDim.x)
        {
          uint64_t z = uint64_t(node_gkey[i / dim]) * 4ull + uint64_t(i % dim) + 0x9E3779B97F4A7C15ull;
          z          = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
          z          = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
          z          = z ^ (z >> 31);
          e[i] = constrained[i] ? 0.0 : (double(z >> 11) * (1.0 / 9007199254740992.0) - 0.5);
        }
    }

    int vec_grid(const gf_context &c, int64_t n)
    {
      return int(std::max<int64_t>(1, std::min<int64_t>((n + NT - 1) / NT, c.max_red_blocks)));
    }

    // operator of a level: nonlinear tangent or the constrained linear system matrix
    const double *level_matrix(const gf_context &c)
    {
      return c.model == GF_MODEL_NEO_HOOKEAN ? c.mat[GF_MAT_TANGENT].val.p :
                                               c.mat[GF_MAT_SYSTEM].val.p;
    }

    // does this level apply its operator from the assembled matrix (not matrix-free)?
    bool level_is_assembled(const gf_context &c)
    {
      return !(c.operator_kind == 1 && c.mg_level == 0 && c.model == GF_MODEL_NEO_HOOKEAN);
    }

    void apply_operator(gf_context &c, double *x, double *y)
    {
      if (c.comm)
        halo_exchange(c, x);
      if (c.mg_matrix_precision == 2 && c.mg_val32_valid && level_is_assembled(c))
        launch_spmv_f32x(c, c.mg_val32.p, x, y);
      else if (c.mg_matrix_precision == 1 && c.mg_val32_valid && level_is_assembled(c))
        launch_spmv_f32(c, c.mg_val32.p, x, y);
      else
        op_apply(c, level_matrix(c), x, y, nullptr);
    }

    template <int DIM>
    void cheb_step(gf_context &c, const double *r_in, const double *v, double c_d, double c_r_unit,
                   bool x_zero, double *x)
    {
      ProfScope ps(c, Profile::MG_VEC);
      cheb_step_kernel<DIM><<<vec_grid(c, c.n_owned_nodes), NT, 0, c.stream>>>(
        c.n_owned_nodes, r_in, v, c.dinv.p, c_d, c_r_unit, c.mg_lmax_dev.p, x_zero, c.mg_r.p,
        c.mg_d.p, x);
      GF_CUDA_CHECK(cudaGetLastError());
    }

    // Chebyshev iteration for D^-1 A on [lmax/ratio, lmax] (Saad, Iterative Methods, alg. 12.1);
    // the coefficients are computed for lmax = 1 and scaled by the device-resident estimate
    void smooth(gf_context &c, const double *b, double *x, bool x_zero, int degree, double ratio)
    {
      const double bb = 1.0, aa = bb / ratio;
      const double theta = 0.5 * (bb + aa), delta = 0.5 * (bb - aa), sigma = theta / delta;
      double       rho = 1.0 / sigma;
      const double *v  = nullptr;
      if (!x_zero)
        {
          apply_operator(c, x, c.mg_v.p);
          v = c.mg_v.p;
        }
      if (c.dim == 3)
        cheb_step<3>(c, b, v, 0.0, 1.0 / theta, x_zero, x);
      else
        cheb_step<2>(c, b, v, 0.0, 1.0 / theta, x_zero, x);
      for (int k = 1; k < degree; ++k)
        {
          apply_operator(c, c.mg_d.p, c.mg_v.p);
          const double rho_new = 1.0 / (2.0 * sigma - rho);
          if (c.dim == 3)
            cheb_step<3>(c, c.mg_r.p, c.mg_v.p, rho_new * rho, 2.0 * rho_new / delta, false, x);
          else
            cheb_step<2>(c, c.mg_r.p, c.mg_v.p, rho_new * rho, 2.0 * rho_new / delta, false, x);
          rho = rho_new;
        }
    }

    void restrict_residual(gf_context &f, const double *r_f, double *b_c)
    {
      gf_context &co = *f.mg.coarse;
      {
        ProfScope ps(f, Profile::MG_VEC);
        // multi-GPU: partial sums for all local coarse nodes (ghosts go to their owner below;
        // replicated coarse level: all its nodes, summed over the ranks below)
        const bool    replicated = f.comm != nullptr && co.comm == nullptr;
        const int64_t n    = (co.comm || replicated) ? co.n_nodes : co.n_owned_nodes;
        const bool    mask = !co.comm && !replicated;
        const unsigned grid = blocks_for(n * 32);
        if (f.dim == 3)
          restrict_kernel<3><<<grid, NT, 0, f.stream>>>(
            n, f.npc, f.mg.n_child, co.nc_ptr.p, co.nc_src.p, f.mg.child_cells.p, f.mg.rl_ptr.p,
            f.mg.rl_ka.p, f.mg.rl_w.p, f.cell_nodes.p, f.mg.first_owner.p, co.constrained.p, mask,
            r_f, b_c);
        else
          restrict_kernel<2><<<grid, NT, 0, f.stream>>>(
            n, f.npc, f.mg.n_child, co.nc_ptr.p, co.nc_src.p, f.mg.child_cells.p, f.mg.rl_ptr.p,
            f.mg.rl_ka.p, f.mg.rl_w.p, f.cell_nodes.p, f.mg.first_owner.p, co.constrained.p, mask,
            r_f, b_c);
        GF_CUDA_CHECK(cudaGetLastError());
      }
      if (co.comm)
        {
          halo_reduce_add(co, b_c);
          vec_zero_constrained(co, b_c);
        }
      else if (f.comm)
        {
          allreduce_sum_vector(f, b_c, co.n_local);
          vec_zero_constrained(co, b_c);
        }
    }

    void prolong_add(gf_context &f, double *x_c, double *x_f)
    {
      gf_context &co = *f.mg.coarse;
      if (co.comm)
        halo_exchange(co, x_c);
      ProfScope ps(f, Profile::MG_VEC);
      const int64_t n = f.n_nodes; // ghosts too: x_f stays consistent without a fine halo exchange
      if (f.dim == 3)
        prolong_add_kernel<3><<<blocks_for(n), NT, 0, f.stream>>>(
          n, f.npc, f.mg.n_child, f.nc_ptr.p, f.nc_src.p, f.mg.parent.p, f.mg.E.p,
          co.cell_nodes.p, f.constrained.p, x_c, x_f);
      else
        prolong_add_kernel<2><<<blocks_for(n), NT, 0, f.stream>>>(
          n, f.npc, f.mg.n_child, f.nc_ptr.p, f.nc_src.p, f.mg.parent.p, f.mg.E.p,
          co.cell_nodes.p, f.constrained.p, x_c, x_f);
      GF_CUDA_CHECK(cudaGetLastError());
    }

    void inject_state(gf_context &f, const double *u_f, double *u_c)
    {
      gf_context &co = *f.mg.coarse;
      {
        ProfScope ps(f, Profile::MG_VEC);
        const int64_t n = co.n_nodes;
        const bool    replicated = f.comm != nullptr && co.comm == nullptr;
        if (f.dim == 3)
          inject_kernel<3><<<blocks_for(n), NT, 0, f.stream>>>(
            n, f.npc, f.mg.n_child, co.nc_ptr.p, co.nc_src.p, f.mg.child_cells.p, f.mg.inj.p,
            f.cell_nodes.p, f.n_owned_nodes, replicated, u_f, u_c);
        else
          inject_kernel<2><<<blocks_for(n), NT, 0, f.stream>>>(
            n, f.npc, f.mg.n_child, co.nc_ptr.p, co.nc_src.p, f.mg.child_cells.p, f.mg.inj.p,
            f.cell_nodes.p, f.n_owned_nodes, replicated, u_f, u_c);
        GF_CUDA_CHECK(cudaGetLastError());
      }
      if (co.comm)
        halo_exchange(co, u_c);
      else if (f.comm)
        allreduce_sum_vector(f, u_c, co.n_local);
    }

    // lambda_max(D^-1 A) by power iteration, warm-started from the previous operator's vector
    void estimate_lmax(gf_context &c)
    {
      const int64_t n = c.n_local;
      int           n_it = 4; // warm start: the operator changed little since the last assembly
      if (!c.mg_e_valid)
        {
          ProfScope ps(c, Profile::MG_VEC);
          pw_init_kernel<<<vec_grid(c, n), NT, 0, c.stream>>>(n, c.dim, c.node_gkey.p,
                                                              c.constrained.p, c.mg_e.p);
          c.mg_e_valid = true;
          n_it         = 24;
        }
      for (int k = 0; k < n_it; ++k)
        {
          apply_operator(c, c.mg_e.p, c.mg_v.p);
          ProfScope ps(c, Profile::MG_VEC, 3);
          if (c.n_red_chunks > 0)
            {
              if (c.dim == 3)
                pw_apply_kernel<3><<<c.n_red_chunks, RED_THREADS, 0, c.stream>>>(
                  c.red_chunk_ptr.p, c.mg_v.p, c.dinv.p, c.mg_r.p, c.partials.p);
              else
                pw_apply_kernel<2><<<c.n_red_chunks, RED_THREADS, 0, c.stream>>>(
                  c.red_chunk_ptr.p, c.mg_v.p, c.dinv.p, c.mg_r.p, c.partials.p);
            }
          reduce_sums(c, 1, -1, false);
          // the first pass only normalises the start vector; afterwards ||e|| = 1 and
          // ||D^-1 A e|| is the eigenvalue estimate
          pw_scale_kernel<<<vec_grid(c, c.n_owned), NT, 0, c.stream>>>(c.n_owned, c.mg_r.p,
                                                                       red_sums(c), c.mg_e.p);
          GF_CUDA_CHECK(cudaGetLastError());
        }
      // the estimate stays on the device (cheb_step_kernel / coarse_cheb_kernel read it there); a
      // copy goes to the pinned host slot asynchronously and is validated at the next CG poll
      lmax_store_kernel<<<1, 1, 0, c.stream>>>(red_sums(c), c.mg_lmax_dev.p);
      GF_CUDA_CHECK(cudaMemcpyAsync(c.h_lmax, c.mg_lmax_dev.p, sizeof(double),
                                    cudaMemcpyDeviceToHost, c.stream));
      c.mg_ops_valid = true;
    }

    void assemble_level_tangent(gf_context &c, const double *u_total)
    {
      double *K = c.mat[GF_MAT_TANGENT].val.p;
      for (int64_t c0 = 0; c0 < c.n_cells; c0 += c.ke_chunk_cells)
        {
          const int64_t c1 = std::min(c.n_cells, c0 + c.ke_chunk_cells);
          // the acceleration only enters the residual, which coarse levels do not use
          launch_nl_cells(c, u_total, c.vec[GF_NL_ACCELERATION].p, c0, c1);
          launch_scatter_matrix(c, K, c0, c1, c0 == 0, true);
        }
      launch_build_precond(c, K);
      c.mat[GF_MAT_TANGENT].valid = true;
    }
  } // namespace

  // (re)build the FP32 operator copies the V-cycle streams when GF_OPT_MG_MATRIX_PRECISION = 1.
  // The outer CG keeps applying the FP64 matrix, so the solve converges to the same tolerance on
  // the same residual; only the (fixed, SPD) preconditioner changes. A level whose copy does not
  // fit the free memory simply stays FP64.
  // one level (rank-local: a conversion kernel, no communication)
  void mg_refresh_f32_level(gf_context &lv)
  {
    gf_context *l     = &lv;
    l->mg_val32_valid = false;
    if (l->mg_matrix_precision == 0 || !level_is_assembled(*l))
      return;
    const double *A = level_matrix(*l);
    if (A == nullptr || l->n_val == 0)
      return;
    if (l->model == GF_MODEL_NEO_HOOKEAN ? !l->mat[GF_MAT_TANGENT].valid : !l->lin_assembled)
      return;
    if (!l->mg_val32.p)
      {
        size_t free_b = 0, total_b = 0;
        GF_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = size_t(l->n_val + 4) * sizeof(float);
        if (need + (size_t(1) << 30) > free_b)
          return;
        l->mg_val32.alloc_zero(size_t(l->n_val + 4), l->stream);
      }
    launch_convert_f32(*l, A, l->mg_val32.p);
    l->mg_val32_valid = true;
  }
  void mg_refresh_f32(gf_context &c)
  {
    for (gf_context *l = &c; l != nullptr; l = l->mg.coarse)
      mg_refresh_f32_level(*l);
  }

  bool mg_active(const gf_context &c)
  {
    return c.precond == GF_PRECOND_MULTIGRID && c.mg.coarse != nullptr;
  }

  void mg_attach(gf_context &f, gf_context &co, const int32_t *child_cells)
  {
    GF_REQUIRE(&f != &co, GF_ERR_INVALID_ARG, "a level cannot be its own coarse level");
    GF_REQUIRE(f.dim == co.dim && f.p == co.p && f.model == co.model && f.device == co.device,
               GF_ERR_INVALID_ARG,
               "multigrid levels must share dim, degree, model and device");
    // a partitioned level may sit on top of a REPLICATED (serial, whole-mesh on every rank) coarse
    // level: restriction / injection all-reduce over the whole coarse vector, everything below runs
    // redundantly without communication
    GF_REQUIRE(f.comm != nullptr || co.comm == nullptr, GF_ERR_INVALID_ARG,
               "a serial multigrid level cannot have a partitioned coarse level");
    GF_REQUIRE(child_cells != nullptr, GF_ERR_INVALID_ARG, "null child_cells");
    // the transfer below relies on nested support points (injection of the state, restriction
    // weights on the half-step grid): true for equidistant nodes only, i.e. FE_Q(1), FE_Q(2)
    GF_REQUIRE(f.lines.n == 0 && co.lines.n == 0, GF_ERR_UNSUPPORTED,
               "geometric multigrid does not take hanging-node constraints (block-Jacobi CG)");
    GF_REQUIRE(f.affine && co.affine, GF_ERR_UNSUPPORTED,
               "geometric multigrid needs parallelepiped cells (general cells: block-Jacobi CG)");
    GF_REQUIRE(f.p <= 2, GF_ERR_UNSUPPORTED,
               "geometric multigrid needs polynomial degree 1 or 2 (higher degrees: block-Jacobi CG)");
    GF_REQUIRE(f.mg.coarse == nullptr, GF_ERR_INVALID_ARG, "level already has a coarse level");
    const int dim = f.dim, p = f.p, npc = f.npc, n_child = 1 << dim;
    cudaStream_t s = f.stream;
    // ---- parent map -----------------------------------------------------------------------------
    std::vector<int32_t> parent(f.n_cells, -1);
    for (int64_t pc = 0; pc < co.n_cells; ++pc)
      for (int k = 0; k < n_child; ++k)
        {
          const int32_t fc = child_cells[pc * n_child + k];
          GF_REQUIRE(fc >= -1 && fc < f.n_cells, GF_ERR_INVALID_ARG, "child cell out of range");
          if (fc >= 0)
            {
              GF_REQUIRE(parent[fc] == -1, GF_ERR_INVALID_ARG, "fine cell has two parents");
              parent[fc] = int32_t(pc * n_child + k);
            }
        }
    for (int64_t fc = 0; fc < f.n_cells; ++fc)
      GF_REQUIRE(parent[fc] >= 0, GF_ERR_INVALID_ARG, "fine cell without a parent cell");
    f.mg.n_child = n_child;
    f.mg.parent.upload(parent.data(), parent.size(), s);
    f.mg.child_cells.upload(child_cells, size_t(co.n_cells) * n_child, s);
    // ---- which directions does this level pair refine? Isotropic refinement (refine_global)
    // halves every edge; a SEMI-COARSENED pair (the host halves only the even repetition counts
    // to get below a large coarsest level) keeps some directions. Read off the geometry: edge
    // length of a parent cell over that of one of its children, per reference direction.
    bool refined[3] = {true, true, true};
    {
      int64_t pc0 = -1, fc0 = -1;
      for (int64_t pc = 0; pc < co.n_cells && pc0 < 0; ++pc)
        for (int k = 0; k < n_child; ++k)
          if (child_cells[pc * n_child + k] >= 0)
            {
              pc0 = pc;
              fc0 = child_cells[pc * n_child + k];
              break;
            }
      GF_REQUIRE(pc0 >= 0, GF_ERR_INVALID_ARG, "no coarse cell has a local child");
      const int gs = dim * dim + 1;
      double    gc[10], gf_[10];
      GF_CUDA_CHECK(cudaMemcpy(gc, co.geom.p + pc0 * gs, gs * sizeof(double), cudaMemcpyDeviceToHost));
      GF_CUDA_CHECK(cudaMemcpy(gf_, f.geom.p + fc0 * gs, gs * sizeof(double), cudaMemcpyDeviceToHost));
      // rows of J^-1 are the gradients of the reference coordinates: |row d| = 1 / edge length
      for (int d = 0; d < dim; ++d)
        {
          double nc = 0, nf = 0;
          for (int j = 0; j < dim; ++j)
            {
              nc += gc[d * dim + j] * gc[d * dim + j];
              nf += gf_[d * dim + j] * gf_[d * dim + j];
            }
          refined[d] = std::sqrt(nf / nc) > 1.5;
        }
      int n_ref = 0;
      for (int d = 0; d < dim; ++d)
        n_ref += refined[d];
      GF_REQUIRE(n_ref >= 1, GF_ERR_INVALID_ARG, "coarse level is not coarser than the fine level");
      for (int64_t pc = 0; pc < co.n_cells; ++pc)
        for (int k = 0; k < n_child; ++k)
          for (int d = 0; d < dim; ++d)
            GF_REQUIRE(refined[d] || !((k >> d) & 1) || child_cells[pc * n_child + k] < 0,
                       GF_ERR_INVALID_ARG, "child index uses a direction that is not refined");
    }
    // ---- embedding matrices: coarse shape functions at the nodes of child k ---------------------
    const std::vector<int> &lex = f.tables.local_lex;
    std::vector<double>     E(size_t(n_child) * npc * npc, 0.0);
    for (int k = 0; k < n_child; ++k)
      for (int a = 0; a < npc; ++a)
        for (int b = 0; b < npc; ++b)
          {
            double w = 1.0;
            for (int d = 0; d < dim; ++d)
              {
                const double xi = refined[d] ?
                                    0.5 * (double(lex[a * 3 + d]) / p + double((k >> d) & 1)) :
                                    double(lex[a * 3 + d]) / p;
                w *= lagrange1d(p, lex[b * 3 + d], xi);
              }
            E[(size_t(k) * npc + a) * npc + b] = std::fabs(w) < 1e-14 ? 0.0 : w;
          }
    f.mg.E.upload(E.data(), E.size(), s);
    std::vector<int32_t> rl_ptr(npc + 1, 0), rl_ka, inj(npc, -1);
    std::vector<double>  rl_w;
    for (int b = 0; b < npc; ++b)
      {
        for (int k = 0; k < n_child; ++k)
          for (int a = 0; a < npc; ++a)
            {
              const double w = E[(size_t(k) * npc + a) * npc + b];
              if (w != 0.0)
                {
                  // beyond the coarse node's plane along the slab axis? (positions in the parent
                  // cell, in units of 1/(2p))
                  bool hi = false;
                  if (f.axis_dir >= 0)
                    hi = refined[f.axis_dir] ?
                           lex[a * 3 + f.axis_dir] + p * ((k >> f.axis_dir) & 1) >
                             2 * lex[b * 3 + f.axis_dir] :
                           lex[a * 3 + f.axis_dir] > lex[b * 3 + f.axis_dir];
                  rl_ka.push_back((k * npc + a) | (hi ? (1 << 30) : 0));
                  rl_w.push_back(w);
                }
            }
        rl_ptr[b + 1] = int32_t(rl_ka.size());
        // nested node: coarse local lex coordinate lb -> fine-node index 2*lb in 0..2p
        int k = 0, la[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d)
          {
            if (!refined[d])
              {
                la[d] = lex[b * 3 + d];
                continue;
              }
            const int t  = 2 * lex[b * 3 + d];
            const int kd = t > p ? 1 : 0;
            k |= kd << d;
            la[d] = t - kd * p;
          }
        for (int a = 0; a < npc; ++a)
          if (lex[a * 3] == la[0] && lex[a * 3 + 1] == la[1] && lex[a * 3 + 2] == la[2])
            inj[b] = k * npc + a;
        GF_REQUIRE(inj[b] >= 0, GF_ERR_INVALID_ARG, "internal: no nested node");
      }
    f.mg.rl_ptr.upload(rl_ptr.data(), rl_ptr.size(), s);
    f.mg.rl_ka.upload(rl_ka.data(), rl_ka.size(), s);
    f.mg.rl_w.upload(rl_w.data(), rl_w.size(), s);
    f.mg.inj.upload(inj.data(), inj.size(), s);
    const int64_t n_cn = f.n_cells * npc;
    f.mg.first_owner.alloc(n_cn);
    first_owner_kernel<<<blocks_for(n_cn), NT, 0, s>>>(n_cn, f.n_owned_nodes, f.cell_nodes.p,
                                                        f.nc_ptr.p, f.nc_src.p,
                                                        f.mg.first_owner.p);
    GF_CUDA_CHECK(cudaGetLastError());
    f.mg.coarse = &co;
    // ---- the whole chain below `f` runs on f's stream and accounts into f's profile -------------
    int level = f.mg_level;
    for (gf_context *l = &f; l != nullptr; l = l->mg.coarse)
      {
        l->mg_level = level++;
        if (l != &f)
          {
            GF_CUDA_CHECK(cudaStreamSynchronize(l->stream));
            comm_forget_stream(l->comm, l->stream);
            if (l->owns_stream && l->stream != f.stream)
              GF_CUDA_CHECK(cudaStreamDestroy(l->stream));
            l->stream      = f.stream;
            l->owns_stream = false;
            l->prof_sink   = f.prof_sink;
            // options set on the finest handle BEFORE the attach apply to the whole chain too
            l->mg_matrix_precision = f.mg_matrix_precision;
            l->spmv_kernel_kind    = f.spmv_kernel_kind;
            l->mg_smoother_degree  = f.mg_smoother_degree;
            l->mg_smoother_ratio   = f.mg_smoother_ratio;
            l->mg_coarse_degree    = f.mg_coarse_degree;
          }
        if (!l->mg_lmax_dev.p)
          {
            l->mg_lmax_dev.alloc_zero(1, s);
            GF_CUDA_CHECK(cudaMallocHost((void **)&l->h_lmax, sizeof(double)));
            *l->h_lmax = 1.0;
          }
        if (!l->mg_r.p)
          {
            l->mg_r.alloc_zero(l->n_local, s);
            l->mg_d.alloc_zero(l->n_local, s);
            l->mg_v.alloc_zero(l->n_local, s);
            l->mg_e.alloc_zero(l->n_local, s);
            l->mg_b.alloc_zero(l->n_local, s);
            l->mg_x.alloc_zero(l->n_local, s);
          }
        l->mg_e_valid = false;
      }
    GF_CUDA_CHECK(cudaStreamSynchronize(s));
  }

  // rebuild the coarse-level operators for the finest-level state and refresh the smoothers
  void mg_update_operators(gf_context &c, const double *u_total)
  {
    if (!c.mg.coarse)
      return;
    const double *u = u_total;
    for (gf_context *l = &c; l->mg.coarse != nullptr; l = l->mg.coarse)
      {
        gf_context &co = *l->mg.coarse;
        if (c.model == GF_MODEL_NEO_HOOKEAN)
          {
            inject_state(*l, u, co.tmp0.p);
            assemble_level_tangent(co, co.tmp0.p);
            u = co.tmp0.p;
          }
        else
          lin_assemble(co);
      }
    mg_refresh_f32(c);
    for (gf_context *l = &c; l != nullptr; l = l->mg.coarse)
      estimate_lmax(*l);
  }

  // x = V-cycle(b), zero initial guess; b and x are level vectors of `c` (x != b)
  void mg_vcycle(gf_context &c, const double *b, double *x)
  {
    if (!c.mg.coarse)
      {
        // coarsest level: one persistent cooperative kernel when it applies (small, not
        // partitioned), else the same Chebyshev iteration launch by launch
        if (!coarse_solve_single_launch(c, level_matrix(c), b, x, c.mg_coarse_degree,
                                        c.mg_coarse_ratio))
          smooth(c, b, x, true, c.mg_coarse_degree, c.mg_coarse_ratio);
        return;
      }
    gf_context &co = *c.mg.coarse;
    smooth(c, b, x, true, c.mg_smoother_degree, c.mg_smoother_ratio);
    apply_operator(c, x, c.mg_v.p);
    {
      ProfScope ps(c, Profile::MG_VEC);
      sub_kernel<<<vec_grid(c, c.n_owned), NT, 0, c.stream>>>(c.n_owned, b, c.mg_v.p, c.mg_r.p);
      GF_CUDA_CHECK(cudaGetLastError());
    }
    restrict_residual(c, c.mg_r.p, co.mg_b.p);
    mg_vcycle(co, co.mg_b.p, co.mg_x.p);
    prolong_add(c, co.mg_x.p, x);
    smooth(c, b, x, false, c.mg_smoother_degree, c.mg_smoother_ratio);
  }
} // namespace gf
