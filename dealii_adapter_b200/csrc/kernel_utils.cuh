// Small device helpers shared by the kernels: compile-time index maps of deal.II's
// SymmetricTensor storage order (00,11,[22],01,[02,12]), dim x dim determinant / inverse,
// warp and block reductions with a fixed (deterministic) combination order.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace gf
{
  __host__ __device__ constexpr int ipow(int b, int e) { return e == 0 ? 1 : b * ipow(b, e - 1); }

  template <int DIM>
  __host__ __device__ constexpr int voigt_index(int i, int j)
  {
    return i == j ? i : (DIM == 2 ? 2 : i + j + 2);
  }
  template <int DIM>
  __host__ __device__ constexpr int voigt_i(int k)
  {
    return k < DIM ? k : (DIM == 2 ? 0 : (k == 5 ? 1 : 0));
  }
  template <int DIM>
  __host__ __device__ constexpr int voigt_j(int k)
  {
    return k < DIM ? k : (DIM == 2 ? 1 : (k == 3 ? 1 : 2));
  }

  template <int DIM>
  __device__ __forceinline__ double det(const double (&F)[DIM][DIM])
  {
    if constexpr (DIM == 2)
      return F[0][0] * F[1][1] - F[1][0] * F[0][1];
    else
      return F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) -
             F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
             F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
  }
  template <int DIM>
  __device__ __forceinline__ void inverse(const double (&t)[DIM][DIM], const double d,
                                          double (&r)[DIM][DIM])
  {
    const double id = 1.0 / d;
    if constexpr (DIM == 2)
      {
        r[0][0] = t[1][1] * id;
        r[0][1] = -t[0][1] * id;
        r[1][0] = -t[1][0] * id;
        r[1][1] = t[0][0] * id;
      }
    else
      {
        r[0][0] = (t[1][1] * t[2][2] - t[1][2] * t[2][1]) * id;
        r[0][1] = (t[0][2] * t[2][1] - t[0][1] * t[2][2]) * id;
        r[0][2] = (t[0][1] * t[1][2] - t[0][2] * t[1][1]) * id;
        r[1][0] = (t[1][2] * t[2][0] - t[1][0] * t[2][2]) * id;
        r[1][1] = (t[0][0] * t[2][2] - t[0][2] * t[2][0]) * id;
        r[1][2] = (t[0][2] * t[1][0] - t[0][0] * t[1][2]) * id;
        r[2][0] = (t[1][0] * t[2][1] - t[1][1] * t[2][0]) * id;
        r[2][1] = (t[0][1] * t[2][0] - t[0][0] * t[2][1]) * id;
        r[2][2] = (t[0][0] * t[1][1] - t[0][1] * t[1][0]) * id;
      }
  }

#ifndef GF_CUDA_EMULATION // PTX helpers: not for kernels that also run in tests/cuda_emu
  // ---- shared-memory mbarriers (producer/consumer pipelines: spmv.cu, assemble_nl.cu) -----------
  __device__ __forceinline__ uint32_t smem_u32(const void *p)
  {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
  }
  __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  }
  __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
  {
    uint32_t done;
    do
      {
        asm volatile("{\n .reg .pred p;\n"
                     " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     " selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
      }
    while (!done);
  }
  __device__ __forceinline__ void mbar_arrive(uint32_t bar)
  {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
  }
  __device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
  }
  // named barrier among a subset of the CTA's warps (all `count` threads call it)
  __device__ __forceinline__ void named_bar_sync(int id, int count)
  {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
  }
#endif // GF_CUDA_EMULATION

  // (the emulation provides __shfl_xor_sync among the 32 CPU threads of a warp)
  __device__ __forceinline__ double warp_sum(double v)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }

  // Warp sums of DIM values at once by a TRANSPOSED butterfly: after the first exchanges every lane
  // carries one value instead of DIM, so DIM = 3 needs 6 shuffle+add rounds (DIM = 2: 5) instead
  // of 15 (10). Returns the sum of v[r] in the lanes [r*S, (r+1)*S), S = 8 (DIM 3) / 16 (DIM 2).
  // The partial sums are combined in the same order as in warp_sum (xor 16, 8, 4, 2, 1), only in
  // fewer lanes, so the result is BITWISE the one warp_sum(v[r]) gives
  // (tests/test_warp_reduction_emulation.py emulates both on the CPU).
  template <int DIM>
  __device__ __forceinline__ double warp_sum_rows(const double (&v)[DIM], const int lane)
  {
    constexpr unsigned full = 0xffffffffu;
    const bool         b4   = (lane & 16) != 0;
    double             t;
    if constexpr (DIM == 2)
      {
        const double keep = b4 ? v[1] : v[0], send = b4 ? v[0] : v[1];
        t                 = keep + __shfl_xor_sync(full, send, 16);
        t += __shfl_xor_sync(full, t, 8);
      }
    else
      {
        static_assert(DIM == 3, "warp_sum_rows: DIM 2 or 3");
        const bool   b3    = (lane & 8) != 0;
        const double keep0 = b4 ? v[2] : v[0], keep1 = b4 ? 0.0 : v[1];
        const double send0 = b4 ? v[0] : v[2], send1 = b4 ? v[1] : 0.0;
        const double r0    = keep0 + __shfl_xor_sync(full, send0, 16);
        const double r1    = keep1 + __shfl_xor_sync(full, send1, 16);
        t                  = (b3 ? r1 : r0) + __shfl_xor_sync(full, b3 ? r0 : r1, 8);
      }
    t += __shfl_xor_sync(full, t, 4);
    t += __shfl_xor_sync(full, t, 2);
    t += __shfl_xor_sync(full, t, 1);
    return t;
  }

  // block-wide sum of up to N values per thread, result valid in thread 0; fixed order
  template <int N>
  __device__ __forceinline__ void block_sum(double (&v)[N], double *smem /* [N*32] */)
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k)
      v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
      for (int k = 0; k < N; ++k)
        smem[k * 32 + warp] = v[k];
    __syncthreads();
    if (warp == 0)
      {
#pragma unroll
        for (int k = 0; k < N; ++k)
          {
            double x = lane < nw ? smem[k * 32 + lane] : 0.0;
            v[k]     = warp_sum(x);
          }
      }
  }
} // namespace gf
