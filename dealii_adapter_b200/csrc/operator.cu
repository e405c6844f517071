// The operator the CG and the smoothers apply on one level: the assembled block-row SpMV (spmv.cu)
// or the matrix-free tangent (matfree.cu), wrapped in the hanging-node condensation C^T A C
// (constraints.cuh) where the handle carries constraint lines.
#include "gf_context.h"

namespace gf
{
  void op_apply(gf_context &c, const double *val, const double *x, double *y, double *dot_partials)
  {
    // hanging-node constraints: y = C^T A C x (constraints.cuh). The constrained entries of x are
    // (re)set from their masters first; they carry no information of their own
    if (c.lines.n > 0)
      lines_distribute(c, const_cast<double *>(x));
    if (c.operator_kind == 1 && c.mg_level == 0 && c.model == GF_MODEL_NEO_HOOKEAN)
      mf_apply(c, x, y, dot_partials);
    else
      launch_spmv(c, val, x, y, dot_partials);
    if (c.lines.n > 0)
      {
        lines_condense(c, y);
        vec_zero_constrained(c, y); // Dirichlet rows collect nothing from hanging neighbours
      }
  }
  int op_dot_partials(const gf_context &c)
  {
    if (c.operator_kind == 1 && c.mg_level == 0 && c.model == GF_MODEL_NEO_HOOKEAN)
      return mf_dot_partials(c);
    return spmv_dot_partials(c);
  }
} // namespace gf
