// K7/K8/K9: device-resident DoF-vector updates (Newmark nonlinear_elasticity.cc:592-622, theta
// scheme linear_elasticity.cc:390-420,584-585), masked l2 norms (nonlinear:549-576), and the
// interface gather/scatter bodies of the Adapter (adapter.h:401-416, 427-442). All HBM-bound
// streaming kernels; norms use the fixed-order two-stage reduction of cg.cu.
#include "reduce.cuh"

namespace gf
{
  namespace
  {
    constexpr int NT = 256;

    inline int grid_for(const gf_context &c, int64_t n)
    {
      return int(std::max<int64_t>(1, std::min<int64_t>((n + NT - 1) / NT, c.max_red_blocks)));
    }

    __global__ void permute_in_kernel(int64_t n, const int32_t *__restrict__ e2i,
                                      const double *__restrict__ ext, double *__restrict__ dst)
    {
      for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < n;
           e += int64_t(gridDim.x) * blockDim.x)
        dst[e2i[e]] = ext[e];
    }
    __global__ void permute_out_kernel(int64_t n, const int32_t *__restrict__ e2i,
                                       const double *__restrict__ src, double *__restrict__ ext)
    {
      for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < n;
           e += int64_t(gridDim.x) * blockDim.x)
        ext[e] = src[e2i[e]];
    }
    __global__ void axpby_kernel(int64_t n, double *__restrict__ y, double a,
                                 const double *__restrict__ x, double b)
    {
      for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n;
           i += int64_t(gridDim.x) * blockDim.x)
        y[i] = a * x[i] + b * y[i];
    }
    __global__ void lincomb3_kernel(int64_t n, double *__restrict__ out, double a,
                                    const double *__restrict__ x, double b,
                                    const double *__restrict__ y, double cc,
                                    const double *__restrict__ z)
    {
      for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n;
           i += int64_t(gridDim.x) * blockDim.x)
        {
          // deal.II order: out.equ(a,x); out.add(b,y,c,z)
          double v = a * x[i];
          v += b * y[i] + cc * z[i];
          out[i] = v;
        }
    }
    __global__ void zero_constrained_kernel(int64_t n, const uint8_t *__restrict__ con,
                                            double *__restrict__ v)
    {
      for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n;
           i += int64_t(gridDim.x) * blockDim.x)
        if (con[i])
          v[i] = 0.0;
    }
    // one CTA per reduction chunk (reduce.cuh): partition-independent partial sums
    template <int DIM>
    __global__ void __launch_bounds__(RED_THREADS)
      norm_chunks_kernel(const int32_t *__restrict__ chunk_ptr, const double *__restrict__ v,
                         const uint8_t *__restrict__ con, bool mask, double *__restrict__ partials)
    {
      __shared__ double sm[32];
      double            acc[1] = {0.0};
      const int64_t     n0 = chunk_ptr[blockIdx.x], n1 = chunk_ptr[blockIdx.x + 1];
      for (int64_t A = n0 + threadIdx.x; A < n1; A += RED_THREADS)
#pragma unroll
        for (int i = 0; i < DIM; ++i)
          {
            const double x = (mask && con[A * DIM + i]) ? 0.0 : v[A * DIM + i];
            acc[0]         = fma(x, x, acc[0]);
          }
      block_sum<1>(acc, sm);
      if (threadIdx.x == 0)
        partials[blockIdx.x] = acc[0];
    }
    __global__ void iface_scatter_kernel(int64_t n_nodes, int dim, const int32_t *__restrict__ dofs,
                                         const double *__restrict__ buf, double *__restrict__ v)
    {
      const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (k >= n_nodes * dim)
        return;
      const int64_t i = k / dim;
      const int     cc = int(k - i * dim);
      v[dofs[cc * n_nodes + i]] = buf[k]; // precice_to_deal[*comp] = read_data_buffer[dim*i+c]
    }
    __global__ void iface_gather_kernel(int64_t n_nodes, int dim, const int32_t *__restrict__ dofs,
                                        const double *__restrict__ v, double *__restrict__ buf)
    {
      const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (k >= n_nodes * dim)
        return;
      const int64_t i = k / dim;
      const int     cc = int(k - i * dim);
      buf[k] = v[dofs[cc * n_nodes + i]]; // write_data_buffer[dim*i+c] = deal_to_precice[*comp]
    }
  } // namespace

  void vec_permute_in(gf_context &c, const double *ext_host, double *dst)
  {
    GF_CUDA_CHECK(cudaMemcpyAsync(c.io_buf.p, ext_host, c.n_ext_dofs * sizeof(double),
                                  cudaMemcpyHostToDevice, c.stream));
    permute_in_kernel<<<grid_for(c, c.n_ext_dofs), NT, 0, c.stream>>>(c.n_ext_dofs, c.perm_e2i.p,
                                                                       c.io_buf.p, dst);
    GF_CUDA_CHECK(cudaGetLastError());
    GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  }
  void vec_permute_out(gf_context &c, const double *src, double *ext_host)
  {
    permute_out_kernel<<<grid_for(c, c.n_ext_dofs), NT, 0, c.stream>>>(c.n_ext_dofs, c.perm_e2i.p,
                                                                        src, c.io_buf.p);
    GF_CUDA_CHECK(cudaGetLastError());
    GF_CUDA_CHECK(cudaMemcpyAsync(ext_host, c.io_buf.p, c.n_ext_dofs * sizeof(double),
                                  cudaMemcpyDeviceToHost, c.stream));
    GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  }
  void vec_zero(gf_context &c, double *v)
  {
    GF_CUDA_CHECK(cudaMemsetAsync(v, 0, c.n_local * sizeof(double), c.stream));
  }
  void vec_copy(gf_context &c, double *dst, const double *src)
  {
    ProfScope ps(c, Profile::UPDATE);
    GF_CUDA_CHECK(cudaMemcpyAsync(dst, src, c.n_local * sizeof(double), cudaMemcpyDeviceToDevice,
                                  c.stream));
  }
  void vec_axpby(gf_context &c, double *y, double a, const double *x, double b)
  {
    ProfScope ps(c, Profile::UPDATE);
    axpby_kernel<<<grid_for(c, c.n_local), NT, 0, c.stream>>>(c.n_local, y, a, x, b);
    GF_CUDA_CHECK(cudaGetLastError());
  }
  void vec_lincomb3(gf_context &c, double *out, double a, const double *x, double b,
                    const double *y, double cc, const double *z)
  {
    ProfScope ps(c, Profile::UPDATE);
    lincomb3_kernel<<<grid_for(c, c.n_local), NT, 0, c.stream>>>(c.n_local, out, a, x, b, y, cc, z);
    GF_CUDA_CHECK(cudaGetLastError());
  }
  void vec_zero_constrained(gf_context &c, double *v)
  {
    ProfScope ps(c, Profile::UPDATE);
    zero_constrained_kernel<<<grid_for(c, c.n_local), NT, 0, c.stream>>>(c.n_local,
                                                                          c.constrained.p, v);
    GF_CUDA_CHECK(cudaGetLastError());
  }
  double vec_masked_norm(gf_context &c, const double *v, bool mask_constrained)
  {
    // constraints.is_constrained(i): Dirichlet dofs and, where the mesh has them, hanging nodes
    const uint8_t *mask = c.lines.n > 0 ? c.lines.norm_mask.p : c.constrained.p;
    {
      ProfScope ps(c, Profile::UPDATE, 2);
      if (c.n_red_chunks > 0)
        {
          if (c.dim == 3)
            norm_chunks_kernel<3><<<c.n_red_chunks, RED_THREADS, 0, c.stream>>>(
              c.red_chunk_ptr.p, v, mask, mask_constrained, c.partials.p);
          else
            norm_chunks_kernel<2><<<c.n_red_chunks, RED_THREADS, 0, c.stream>>>(
              c.red_chunk_ptr.p, v, mask, mask_constrained, c.partials.p);
        }
      GF_CUDA_CHECK(cudaGetLastError());
      reduce_sums(c, 1, -1, false);
    }
    GF_CUDA_CHECK(
      cudaMemcpyAsync(c.h_norm, red_sums(c), sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    comm_check(c);
    return std::sqrt(c.h_norm[0]);
  }
  void iface_scatter(gf_context &c, const double *host_buf, double *vec)
  {
    const int64_t n = c.n_iface_nodes * c.dim;
    if (n == 0)
      return;
    GF_CUDA_CHECK(cudaMemcpyAsync(c.iface_buf.p, host_buf, n * sizeof(double),
                                  cudaMemcpyHostToDevice, c.stream));
    iface_scatter_kernel<<<unsigned((n + NT - 1) / NT), NT, 0, c.stream>>>(
      c.n_iface_nodes, c.dim, c.iface_dofs_i.p, c.iface_buf.p, vec);
    GF_CUDA_CHECK(cudaGetLastError());
    GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  }
  void iface_gather(gf_context &c, const double *vec, double *host_buf)
  {
    const int64_t n = c.n_iface_nodes * c.dim;
    if (n == 0)
      return;
    iface_gather_kernel<<<unsigned((n + NT - 1) / NT), NT, 0, c.stream>>>(
      c.n_iface_nodes, c.dim, c.iface_dofs_i.p, vec, c.iface_buf.p);
    GF_CUDA_CHECK(cudaGetLastError());
    GF_CUDA_CHECK(cudaMemcpyAsync(host_buf, c.iface_buf.p, n * sizeof(double),
                                  cudaMemcpyDeviceToHost, c.stream));
    GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  }
} // namespace gf
