// Hanging-node constraint lines: set-up (caller numbering -> internal, transposed lists) and the
// launchers of constraints.cuh.
#include <algorithm>
#include <map>

#include "constraints.cuh"
#include "gf_context.h"

namespace gf
{
  void setup_lines(gf_context &c, const gf_desc &d)
  {
    ConstraintLines &L = c.lines;
    L.n                = 0;
    if (d.n_constraint_lines <= 0)
      return;
    GF_REQUIRE(d.line_dof && d.line_ptr && d.line_master && d.line_weight, GF_ERR_INVALID_ARG,
               "constraint lines: null arrays");
    GF_REQUIRE(c.comm == nullptr && c.n_owned == c.n_local, GF_ERR_UNSUPPORTED,
               "hanging-node constraints need a serial handle");
    const int64_t n = d.n_constraint_lines, nnz = d.line_ptr[n];
    GF_REQUIRE(d.line_ptr[0] == 0 && nnz >= 0, GF_ERR_INVALID_ARG, "constraint lines: bad line_ptr");
    std::vector<uint8_t> is_slave(c.n_local, 0);
    std::vector<int32_t> dof(n), master(nnz);
    std::vector<int64_t> ptr(d.line_ptr, d.line_ptr + n + 1);
    std::vector<double>  weight(d.line_weight, d.line_weight + nnz);
    for (int64_t k = 0; k < n; ++k)
      {
        const int32_t e = d.line_dof[k];
        GF_REQUIRE(e >= 0 && e < c.n_ext_dofs && ptr[k + 1] >= ptr[k], GF_ERR_INVALID_ARG,
                   "constraint lines: dof out of range");
        dof[k] = c.h_perm_e2i[e];
        GF_REQUIRE(!is_slave[dof[k]], GF_ERR_INVALID_ARG, "constraint lines: dof constrained twice");
        GF_REQUIRE(!d.constrained[e], GF_ERR_INVALID_ARG,
                   "constraint lines: a dof is both Dirichlet- and hanging-node-constrained");
        is_slave[dof[k]] = 1;
      }
    for (int64_t j = 0; j < nnz; ++j)
      {
        const int32_t e = d.line_master[j];
        GF_REQUIRE(e >= 0 && e < c.n_ext_dofs, GF_ERR_INVALID_ARG,
                   "constraint lines: master out of range");
        master[j] = c.h_perm_e2i[e];
      }
    for (int64_t j = 0; j < nnz; ++j)
      GF_REQUIRE(!is_slave[master[j]], GF_ERR_INVALID_ARG,
                 "constraint lines: chains must be resolved (AffineConstraints::close())");
    // transposed lists: per master the (constrained dof, weight) pairs in ascending line order
    std::map<int32_t, std::vector<std::pair<int32_t, double>>> by_master;
    for (int64_t k = 0; k < n; ++k)
      for (int64_t j = ptr[k]; j < ptr[k + 1]; ++j)
        by_master[master[j]].push_back({dof[k], weight[j]});
    std::vector<int32_t> mdof, slave;
    std::vector<int64_t> mptr{0};
    std::vector<double>  mweight;
    for (auto &kv : by_master)
      {
        mdof.push_back(kv.first);
        for (auto &sw : kv.second)
          {
            slave.push_back(sw.first);
            mweight.push_back(sw.second);
          }
        mptr.push_back(int64_t(slave.size()));
      }
    cudaStream_t s = c.stream;
    L.dof.upload(dof.data(), dof.size(), s);
    L.ptr.upload(ptr.data(), ptr.size(), s);
    L.master.upload(master.data(), std::max<size_t>(master.size(), 1), s);
    L.weight.upload(weight.data(), std::max<size_t>(weight.size(), 1), s);
    L.mdof.upload(mdof.data(), std::max<size_t>(mdof.size(), 1), s);
    L.mptr.upload(mptr.data(), mptr.size(), s);
    L.slave.upload(slave.data(), std::max<size_t>(slave.size(), 1), s);
    L.mweight.upload(mweight.data(), std::max<size_t>(mweight.size(), 1), s);
    L.n         = n;
    L.n_masters = int64_t(mdof.size());
    // the norms of the Newton loop skip every constrained dof (constraints.is_constrained(i),
    // nonlinear_elasticity.cc:553-575): Dirichlet or hanging
    std::vector<uint8_t> mask(c.n_local);
    c.constrained.download(mask.data(), s);
    for (int64_t i = 0; i < c.n_local; ++i)
      mask[i] = (mask[i] || is_slave[i]) ? 1 : 0;
    L.norm_mask.upload(mask.data(), mask.size(), s);
  }

  namespace
  {
    unsigned grid_of(int64_t n) { return unsigned(std::min<int64_t>((n + 127) / 128, 65535)); }
  } // namespace

  void lines_distribute(gf_context &c, double *x)
  {
    const ConstraintLines &L = c.lines;
    if (L.n == 0)
      return;
    ProfScope ps(c, Profile::UPDATE);
    lines_distribute_kernel<<<grid_of(L.n), 128, 0, c.stream>>>(L.n, L.dof.p, L.ptr.p, L.master.p,
                                                               L.weight.p, x);
    GF_CUDA_CHECK(cudaGetLastError());
  }

  void lines_condense(gf_context &c, double *y)
  {
    const ConstraintLines &L = c.lines;
    if (L.n == 0)
      return;
    ProfScope ps(c, Profile::UPDATE, 2);
    if (L.n_masters > 0)
      lines_condense_kernel<<<grid_of(L.n_masters), 128, 0, c.stream>>>(
        L.n_masters, L.mdof.p, L.mptr.p, L.slave.p, L.mweight.p, y);
    lines_zero_kernel<<<grid_of(L.n), 128, 0, c.stream>>>(L.n, L.dof.p, y);
    GF_CUDA_CHECK(cudaGetLastError());
  }
} // namespace gf
