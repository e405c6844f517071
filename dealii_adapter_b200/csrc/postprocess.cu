// K12: output-path post-processing on the device — what DataOut + Postprocessor compute at output
// steps (nonlinear_elasticity.cc:1215-1254 with include/postprocessor.h:44-76;
// linear_elasticity.cc:590-629 with its postprocessor.h). `build_patches(q_mapping, degree)`
// evaluates the displacement field and its gradient at the (degree+1)^dim equidistant patch points
// of every cell THROUGH MappingQEulerian, i.e. on the displaced configuration x = X + u:
//     grad_x u = grad_X u . (I + grad_X u)^-1
// and the Postprocessor stores  [ u_0..u_{dim-1} | strain_{de} = (grad_x u_de + grad_x u_ed)/2 ]
// with strain components unrolled row-major (component_to_unrolled_index(d,e) = d*dim + e).
// One thread per (cell, patch point); patch points in lexicographic order (x fastest) as in
// DataOutBase::Patch. HBM-bound and only run at output intervals: reads 8*dpc B/cell of nodal
// values (L2-resident neighbours), writes 8*(dim+dim^2)*npc B/cell.
#include "assemble_general.cuh"
#include "gf_context.h"
#include "kernel_utils.cuh"

namespace gf
{
  namespace
  {
    template <int DIM>
    __global__ void __launch_bounds__(256)
      postprocess_kernel(const int64_t c0, const int64_t n_cells, const int npc,
                         const int32_t *__restrict__ cell_nodes, const double *__restrict__ geom,
                         const double *__restrict__ ppN,  // [npc pts][npc]
                         const double *__restrict__ ppdN, // [npc pts][npc][DIM] unit-cell gradients
                         const double *__restrict__ u, double *__restrict__ fields,
                         int *__restrict__ err_flag)
    {
      constexpr int NF = DIM + DIM * DIM;
      const int64_t t  = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (t >= n_cells * npc)
        return;
      const int64_t  cell = c0 + t / npc; // fields is the chunk's buffer, indexed by t
      const int      pt   = int(t % npc);
      const int32_t *cn   = cell_nodes + cell * npc;
      const double * gm   = geom + cell * (DIM * DIM + 1);
      double         val[DIM], G[DIM][DIM]; // u and its unit-cell gradient at the patch point
#pragma unroll
      for (int c = 0; c < DIM; ++c)
        {
          val[c] = 0.0;
#pragma unroll
          for (int k = 0; k < DIM; ++k)
            G[c][k] = 0.0;
        }
      for (int a = 0; a < npc; ++a)
        {
          const double  Na = ppN[pt * npc + a];
          const double *dN = ppdN + (pt * npc + a) * DIM;
          const int64_t nd = cn[a];
#pragma unroll
          for (int c = 0; c < DIM; ++c)
            {
              const double ua = u[nd * DIM + c];
              val[c]          = fma(Na, ua, val[c]);
#pragma unroll
              for (int k = 0; k < DIM; ++k)
                G[c][k] = fma(dN[k], ua, G[c][k]);
            }
        }
      double H[DIM][DIM], F[DIM][DIM], Finv[DIM][DIM]; // H = grad_X u
#pragma unroll
      for (int c = 0; c < DIM; ++c)
#pragma unroll
        for (int j = 0; j < DIM; ++j)
          {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k)
              s = fma(G[c][k], gm[k * DIM + j], s);
            H[c][j] = s;
            F[c][j] = (c == j ? 1.0 : 0.0) + s;
          }
      const double dF = det<DIM>(F);
      if (!(dF > 0.0))
        {
          *err_flag = 1; // inverted Eulerian mapping: deal.II would assert as well
          return;
        }
      inverse<DIM>(F, dF, Finv);
      double g[DIM][DIM]; // grad_x u = H F^-1
#pragma unroll
      for (int c = 0; c < DIM; ++c)
#pragma unroll
        for (int j = 0; j < DIM; ++j)
          {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k)
              s = fma(H[c][k], Finv[k][j], s);
            g[c][j] = s;
          }
      double *out = fields + t * NF;
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          out[d] = val[d];
#pragma unroll
          for (int e = 0; e < DIM; ++e)
            out[DIM + d * DIM + e] = (g[d][e] + g[e][d]) / 2; // postprocessor.h:63-70
        }
    }
  } // namespace

  // cells [c0, c1) -> fields_dev: [c1 - c0][npc][dim + dim*dim] device buffer
  void launch_postprocess(gf_context &c, const double *u, int64_t c0, int64_t c1,
                          double *fields_dev)
  {
    const int dim = c.dim, p = c.p, npc = c.npc;
    if (!c.pp_N.p)
      {
        // shape values / unit-cell gradients of the hierarchical FE_Q nodes at the lexicographic
        // patch points xi = (i, j, k) / p
        const std::vector<int> &lex = c.tables.local_lex;
        const gf_fe::Basis1D    basis(p); // FE_Q(p) on its Gauss-Lobatto support points
        auto lag1  = [&](int, int i, double x) { return basis.value(i, x); };
        auto dlag1 = [&](int, int i, double x) { return basis.derivative(i, x); };
        std::vector<double>     N(size_t(npc) * npc), dN(size_t(npc) * npc * dim);
        for (int pt = 0; pt < npc; ++pt)
          {
            double xi[3] = {0, 0, 0};
            int    rem   = pt;
            for (int d = 0; d < dim; ++d)
              {
                xi[d] = double(rem % (p + 1)) / p;
                rem /= (p + 1);
              }
            for (int a = 0; a < npc; ++a)
              {
                double v = 1;
                for (int d = 0; d < dim; ++d)
                  v *= lag1(p, lex[a * 3 + d], xi[d]);
                N[size_t(pt) * npc + a] = v;
                for (int k = 0; k < dim; ++k)
                  {
                    double w = 1;
                    for (int d = 0; d < dim; ++d)
                      w *= (d == k) ? dlag1(p, lex[a * 3 + d], xi[d]) :
                                      lag1(p, lex[a * 3 + d], xi[d]);
                    dN[(size_t(pt) * npc + a) * dim + k] = w;
                  }
              }
          }
        c.pp_N.upload(N.data(), N.size(), c.stream);
        c.pp_dN.upload(dN.data(), dN.size(), c.stream);
      }
    if (c1 <= c0)
      return;
    ProfScope      ps(c, Profile::UPDATE);
    const int64_t  n    = (c1 - c0) * npc;
    const unsigned grid = unsigned((n + 255) / 256);
    if (!c.affine)
      {
        const unsigned g = std::min(grid, 65535u);
        if (dim == 3)
          postprocess_general_kernel<3><<<g, 256, 0, c.stream>>>(
            c0, c1 - c0, npc, c.cell_nodes.p, c.cell_verts.p, c.tables.dphip.p, c.pp_N.p,
            c.pp_dN.p, u, fields_dev, c.err_flag.p);
        else
          postprocess_general_kernel<2><<<g, 256, 0, c.stream>>>(
            c0, c1 - c0, npc, c.cell_nodes.p, c.cell_verts.p, c.tables.dphip.p, c.pp_N.p,
            c.pp_dN.p, u, fields_dev, c.err_flag.p);
      }
    else if (dim == 3)
      postprocess_kernel<3><<<grid, 256, 0, c.stream>>>(c0, c1 - c0, npc, c.cell_nodes.p, c.geom.p,
                                                        c.pp_N.p, c.pp_dN.p, u, fields_dev,
                                                        c.err_flag.p);
    else
      postprocess_kernel<2><<<grid, 256, 0, c.stream>>>(c0, c1 - c0, npc, c.cell_nodes.p, c.geom.p,
                                                        c.pp_N.p, c.pp_dN.p, u, fields_dev,
                                                        c.err_flag.p);
    GF_CUDA_CHECK(cudaGetLastError());
  }
} // namespace gf
