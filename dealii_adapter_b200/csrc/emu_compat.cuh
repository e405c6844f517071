// Kernels written against this header compile both with nvcc (the product) and, unchanged, with
// g++ against the CUDA stand-in of tests/cuda_emu/ (-DGF_CUDA_EMULATION), which runs a CTA as CPU
// threads with a barrier for __syncthreads: the generic-degree kernels were developed without a
// GPU at hand and are checked that way against the oracle (tests/test_cuda_emulation.py).
// Restrictions for such kernels: dynamic shared memory only (through GF_DYN_SMEM), no warp-level
// primitives, no mbarriers / TMA.
#pragma once
#ifdef GF_CUDA_EMULATION
#include <cuda_runtime.h> // the stand-in of tests/cuda_emu (-I tests/cuda_emu)
#define GF_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(gf_emu::dynamic_smem())
#else
#define GF_DYN_SMEM(type, name)                                                                  \
  extern __shared__ __align__(16) unsigned char name##_dyn_smem_raw[];                           \
  type *name = reinterpret_cast<type *>(name##_dyn_smem_raw)
#endif
