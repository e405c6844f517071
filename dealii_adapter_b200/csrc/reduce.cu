// Partition-independent reduction plan and the single-rank second stage (reduce.cuh).
#include <algorithm>

#include "reduce.cuh"

namespace gf
{
  namespace
  {
    __global__ void __launch_bounds__(TREE_THREADS)
      red_finish_kernel(const double *__restrict__ partials, const int stride, const int n_chunks,
                        const int n_sums, double *__restrict__ sums, CGScalars *s, const int phase,
                        const bool check_status)
    {
      if (check_status && s->status != 0)
        return;
      __shared__ double sm[32];
      for (int k = 0; k < n_sums; ++k)
        {
          const double w = tree_sum_1024(partials + size_t(k) * stride, n_chunks, sm);
          if (threadIdx.x == 0)
            sums[k] = w;
        }
      if (phase >= 0 && threadIdx.x == 0)
        cg_scalar_step(s, sums, phase);
    }

    template <int DIM>
    __global__ void __launch_bounds__(RED_THREADS)
      dot_chunks_kernel(const int32_t *__restrict__ chunk_ptr, const double *__restrict__ a,
                        const double *__restrict__ b, double *__restrict__ partials,
                        const int *status)
    {
      if (status != nullptr && *status != 0)
        return;
      __shared__ double sm[32];
      const int64_t     n0 = chunk_ptr[blockIdx.x], n1 = chunk_ptr[blockIdx.x + 1];
      double            acc[1] = {0.0};
      for (int64_t A = n0 + threadIdx.x; A < n1; A += RED_THREADS)
#pragma unroll
        for (int i = 0; i < DIM; ++i)
          acc[0] = fma(a[A * DIM + i], b[A * DIM + i], acc[0]);
      block_sum<1>(acc, sm);
      if (threadIdx.x == 0)
        partials[blockIdx.x] = acc[0];
    }
  } // namespace

  double *red_sums(gf_context &c) { return c.partials.p + 3 * size_t(c.red_stride); }

  // chunks of the owned node planes; the chunk size is a function of the GLOBAL node count only
  void build_reduction_plan(gf_context &c)
  {
    const int64_t n_global_nodes = c.n_global_dofs_for_maxit / c.dim;
    int           cn             = 1024;
    while (n_global_nodes / cn > 16384)
      cn *= 2;
    c.red_chunk_nodes = cn;
    std::vector<int32_t> ptr;
    for (size_t pl = 0; pl + 1 < c.h_plane_ptr.size(); ++pl)
      for (int32_t n0 = c.h_plane_ptr[pl]; n0 < c.h_plane_ptr[pl + 1]; n0 += cn)
        ptr.push_back(n0);
    ptr.push_back(int32_t(c.n_owned_nodes));
    c.n_red_chunks = int(ptr.size()) - 1;
    c.red_stride   = std::max(1, (c.n_red_chunks + 7) & ~7);
    c.red_chunk_ptr.upload(ptr.data(), ptr.size(), c.stream);
    c.partials.alloc_zero(3 * size_t(c.red_stride) + 16, c.stream);
    c.red_chunk_base      = 0;
    c.n_red_chunks_global = c.n_red_chunks;
    c.red_rank_chunks.assign(1, c.n_red_chunks);
    if (c.comm)
      {
        // every rank learns every rank's chunk count: all-reduce of a vector with one own entry
        const int           P = c.comm->n_ranks;
        std::vector<double> cnt(P2P_AR_MAX, 0.0);
        GF_REQUIRE(P <= P2P_AR_MAX, GF_ERR_UNSUPPORTED, "more than 8 ranks");
        cnt[c.comm->rank] = double(c.n_red_chunks);
        DevBuf<double> d;
        d.upload(cnt.data(), cnt.size(), c.stream);
        allreduce_sum(c, d.p, P);
        d.download(cnt.data(), c.stream);
        c.red_rank_chunks.assign(P, 0);
        int64_t total = 0;
        for (int r = 0; r < P; ++r)
          {
            c.red_rank_chunks[r] = int(cnt[r] + 0.5);
            if (r == c.comm->rank)
              c.red_chunk_base = total;
            total += c.red_rank_chunks[r];
          }
        c.n_red_chunks_global = total;
        GF_REQUIRE(total <= int64_t(P2P_GATHER_MAX), GF_ERR_UNSUPPORTED,
                   "too many reduction chunks for the gather window");
      }
  }

  void reduce_sums(gf_context &c, int n_sums, int cg_phase, bool check_status)
  {
    if (c.comm)
      {
        comm_reduce_sums(c, n_sums, cg_phase, check_status);
        return;
      }
    red_finish_kernel<<<1, TREE_THREADS, 0, c.stream>>>(c.partials.p, c.red_stride, c.n_red_chunks,
                                                        n_sums, red_sums(c), c.cg_scalars.p,
                                                        cg_phase, check_status);
    GF_CUDA_CHECK(cudaGetLastError());
  }

  void launch_dot_chunks(gf_context &c, const double *a, const double *b, bool check_status)
  {
    const int *st = check_status ? &c.cg_scalars.p->status : nullptr;
    if (c.n_red_chunks == 0)
      return;
    if (c.dim == 3)
      dot_chunks_kernel<3><<<c.n_red_chunks, RED_THREADS, 0, c.stream>>>(c.red_chunk_ptr.p, a, b,
                                                                        c.partials.p, st);
    else
      dot_chunks_kernel<2><<<c.n_red_chunks, RED_THREADS, 0, c.stream>>>(c.red_chunk_ptr.p, a, b,
                                                                        c.partials.p, st);
    GF_CUDA_CHECK(cudaGetLastError());
  }
} // namespace gf
