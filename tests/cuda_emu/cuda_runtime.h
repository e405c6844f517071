// CPU stand-in for the few CUDA constructs the emulation-ready kernels use
// (dealii_adapter_b200/csrc/emu_compat.cuh): a CTA runs as `blockDim.x` std::threads, __syncthreads
// is a std::barrier, dynamic shared memory a per-block heap buffer, blocks run one after another.
// Test infrastructure only (tests/test_cuda_emulation.py); never part of the product.
#pragma once
#include <math.h>

#include <barrier>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <thread>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint3
{
  unsigned x, y, z;
};
struct dim3
{
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1)
    : x(a)
    , y(b)
    , z(c)
  {}
};

namespace gf_emu
{
  struct Block
  {
    std::barrier<> *barrier = nullptr;
    unsigned char * smem    = nullptr;
  };
  inline thread_local uint3 t_thread_idx{0, 0, 0}, t_block_idx{0, 0, 0};
  inline thread_local dim3  t_block_dim, t_grid_dim;
  inline thread_local Block t_block;
  inline void *             dynamic_smem() { return t_block.smem; }

  // kernel<<<grid, block, smem_bytes>>>(args...) -> launch(grid, block, smem_bytes, [&] { kernel(args...); })
  inline void launch(unsigned grid, unsigned block, size_t smem_bytes,
                     const std::function<void()> &kernel)
  {
    for (unsigned b = 0; b < grid; ++b)
      {
        // uninitialised shared memory must not look like zeros
        std::vector<double> smem(smem_bytes / sizeof(double) + 2, nan(""));
        std::barrier<>           bar(block);
        std::vector<std::thread> threads;
        for (unsigned t = 0; t < block; ++t)
          threads.emplace_back([&, t] {
            t_thread_idx = {t, 0, 0};
            t_block_idx  = {b, 0, 0};
            t_block_dim  = dim3(block);
            t_grid_dim   = dim3(grid);
            t_block      = {&bar, reinterpret_cast<unsigned char *>(smem.data())};
            kernel();
            bar.arrive_and_drop(); // a finished thread no longer takes part in barriers
          });
        for (auto &th : threads)
          th.join();
      }
  }
} // namespace gf_emu

#define threadIdx (gf_emu::t_thread_idx)
#define blockIdx (gf_emu::t_block_idx)
#define blockDim (gf_emu::t_block_dim)
#define gridDim (gf_emu::t_grid_dim)

inline void __syncthreads() { gf_emu::t_block.barrier->arrive_and_wait(); }
inline int  atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T>
inline T __ldg(const T *p)
{
  return *p;
}
