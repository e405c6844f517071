// Hanging-node constraint lines u_s = sum_j w_j u_m(j) (gf_desc.line_*): what
// AffineConstraints::condense / distribute_local_to_global and ::distribute do in the reference
// (linear_elasticity.cc:196-207 make_hanging_node_constraints, :355,:422 condense, :571-572
// distribute; nonlinear: constraints.distribute_local_to_global :769-773, distribute :1208).
// With C = the interpolation from the unconstrained dofs to all dofs, the condensed system is
// C^T A C; the library keeps A unconstrained and applies
//     x  ->  C x          lines_distribute_kernel   (thread per constrained dof)
//     y  ->  C^T y        lines_condense_kernel     (thread per master dof, gather form: fixed
//                         summation order, no atomics) followed by lines_zero_kernel
// around the SpMV inside the CG (matfree.cu: op_apply) and on the right-hand sides.
// Emulation-ready (emu_compat.cuh): tests/test_cuda_emulation.py runs them on the CPU.
#pragma once
#include <cstdint>

#include "emu_compat.cuh"

namespace gf
{
  // x[s] = sum_j w_j x[m_j] for every constrained dof s
  __global__ void lines_distribute_kernel(const int64_t n_lines, const int32_t *__restrict__ dof,
                                          const int64_t *__restrict__ ptr,
                                          const int32_t *__restrict__ master,
                                          const double *__restrict__ weight, double *__restrict__ x)
  {
    for (int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; k < n_lines;
         k += int64_t(gridDim.x) * blockDim.x)
      {
        double v = 0.0;
        for (int64_t j = ptr[k]; j < ptr[k + 1]; ++j)
          v += weight[j] * x[master[j]];
        x[dof[k]] = v;
      }
  }

  // y[m] += sum_j w_j y[s_j] for every master dof m (transposed lists, ascending constrained dof)
  __global__ void lines_condense_kernel(const int64_t n_masters, const int32_t *__restrict__ mdof,
                                        const int64_t *__restrict__ mptr,
                                        const int32_t *__restrict__ slave,
                                        const double *__restrict__ weight, double *__restrict__ y)
  {
    for (int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; k < n_masters;
         k += int64_t(gridDim.x) * blockDim.x)
      {
        double v = y[mdof[k]];
        for (int64_t j = mptr[k]; j < mptr[k + 1]; ++j)
          v += weight[j] * y[slave[j]];
        y[mdof[k]] = v;
      }
  }

  // y[s] = 0 for every constrained dof (after the masters have collected their shares)
  __global__ void lines_zero_kernel(const int64_t n_lines, const int32_t *__restrict__ dof,
                                    double *__restrict__ y)
  {
    for (int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; k < n_lines;
         k += int64_t(gridDim.x) * blockDim.x)
      y[dof[k]] = 0.0;
  }
} // namespace gf
