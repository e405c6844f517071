// Reference-cell tables on the host: FE_Q(p) shape values / unit-cell gradients in deal.II's
// hierarchical order at QGauss points, face tables in QProjector order; any polynomial degree
// (fe_basis.h). These replace FEValues / FEFaceValues (reference call sites:
// nonlinear_elasticity.cc:668-669,891,902-906,807,815; linear_elasticity.cc:254-258,280,467-470,
// 493,499).
//   quadrature: nonlinear QGauss(p+2) (nonlinear_elasticity.cc:74-75), linear QGauss(p+1)
//   (linear_elasticity.cc:61,252,465).
// Pure C++ (no CUDA): fe_tables.cu uploads the result; the CPU emulation of the generic-degree
// kernels (tests/cuda_emu) feeds the same tables to the same kernel source.
#pragma once
#include <cmath>
#include <vector>

#include "fe_basis.h"

namespace gf
{
  struct FETablesHost
  {
    int dim = 0, p = 0, npc = 0, dpc = 0, nq1 = 0, nq = 0, nqf = 0, nv = 0;
    std::vector<double> hN, hdN, hw, hNf, hwf; // [nq][npc], [nq][npc][dim], [nq], [2 dim][nqf][npc], [nqf]
    std::vector<double> hMref;                 // [npc][npc] sum_q w N_a N_b
    // MappingQ1 geometry of general (non-affine) cells: unit-cell gradients of the 2^dim vertex
    // shape functions at the cell / face quadrature points and at the output patch points
    std::vector<double> hdphi;  // [nq][nv][dim]
    std::vector<double> hdphif; // [2 dim][nqf][nv][dim]
    std::vector<double> hdphip; // [npc patch points (lexicographic, xi = i / p)][nv][dim]
    std::vector<int>    local_lex;             // [npc][3]
    // FESystem local DoF of (hierarchical node a, component c): loc_of[a * dim + c]; = a * dim + c
    // for p <= 2, entity-major / component / entity DoF for p >= 3 (fe_basis.h)
    std::vector<int>    loc_of;
    std::vector<double> support_1d; // the p + 1 support points of FE_Q(p) on [0, 1]
    // 1D factors of the tensor-product tables (matrix-free sum factorisation, matfree.cu)
    std::vector<double> h1N, h1D, h1w; // [nq1][p+1] values / derivatives at the 1D Gauss points, [nq1]
    std::vector<int>    lex2hier;      // lexicographic local node -> FE_Q hierarchical local node
  };

  inline void gauss01(int n, std::vector<double> &x, std::vector<double> &w)
  {
    // Gauss-Legendre on [0,1] by Newton iteration on P_n
    x.assign(n, 0.);
    w.assign(n, 0.);
    for (int i = 0; i < (n + 1) / 2; ++i)
      {
        long double z = cosl(3.14159265358979323846264338327950288L * (i + 0.75L) / (n + 0.5L));
        long double dp = 1;
        for (int iter = 0; iter < 100; ++iter)
          {
            long double p0 = 1, p1 = z;
            for (int k = 2; k <= n; ++k)
              {
                const long double pk = ((2 * k - 1) * z * p1 - (k - 1) * p0) / k;
                p0                   = p1;
                p1                   = pk;
              }
            if (n == 1)
              {
                p0 = 1;
                p1 = z;
              }
            dp                   = n * (z * p1 - p0) / (z * z - 1);
            const long double dz = p1 / dp;
            z -= dz;
            if (fabsl(dz) < 1e-19L)
              break;
          }
        // recompute derivative at converged z
        {
          long double p0 = 1, p1 = z;
          for (int k = 2; k <= n; ++k)
            {
              const long double pk = ((2 * k - 1) * z * p1 - (k - 1) * p0) / k;
              p0                   = p1;
              p1                   = pk;
            }
          dp = n * (z * p1 - p0) / (z * z - 1);
        }
        const double xx = double(0.5L * z);
        const double ww = double(1.0L / ((1.0L - z * z) * dp * dp));
        x[i]            = 0.5 - xx;
        x[n - 1 - i]    = 0.5 + xx;
        w[i] = w[n - 1 - i] = ww;
      }
  }

  inline void build_host_tables(FETablesHost &t, const int dim_, const int p_, const bool nonlinear)
  {
    t.dim = dim_;
    t.p   = p_;
    t.nv  = 1 << dim_;
    t.npc       = 1;
    for (int d = 0; d < t.dim; ++d)
      t.npc *= (t.p + 1);
    t.dpc = t.npc * t.dim;
    t.nq1 = nonlinear ? t.p + 2 : t.p + 1;
    t.nq  = 1;
    for (int d = 0; d < t.dim; ++d)
      t.nq *= t.nq1;
    t.nqf = t.nq / t.nq1;
    // FE_Q(p): hierarchical local order, Gauss-Lobatto support points, FESystem local numbering
    // (fe_basis.h restates deal.II's conventions for every degree)
    gf_fe::local_nodes(t.dim, t.p, t.local_lex);
    {
      std::vector<int> node_of, comp_of;
      gf_fe::system_numbering(t.dim, t.p, node_of, comp_of, t.loc_of);
    }
    const gf_fe::Basis1D basis(t.p);
    t.support_1d = basis.nodes;
    auto lag  = [&](int, int i, double x) { return basis.value(i, x); };
    auto dlag = [&](int, int i, double x) { return basis.derivative(i, x); };
    const int dim = t.dim, p = t.p, npc = t.npc, nq = t.nq, nq1 = t.nq1, nqf = t.nqf;
    std::vector<double> x1, w1;
    gauss01(nq1, x1, w1);
    auto shape = [&](int a, const double *xi) {
      double v = 1;
      for (int d = 0; d < dim; ++d)
        v *= lag(p, t.local_lex[a * 3 + d], xi[d]);
      return v;
    };
    t.h1N.assign(nq1 * (p + 1), 0.);
    t.h1D.assign(nq1 * (p + 1), 0.);
    t.h1w = w1;
    for (int q = 0; q < nq1; ++q)
      for (int i = 0; i <= p; ++i)
        {
          t.h1N[q * (p + 1) + i] = lag(p, i, x1[q]);
          t.h1D[q * (p + 1) + i] = dlag(p, i, x1[q]);
        }
    t.lex2hier.assign(npc, -1);
    for (int a = 0; a < npc; ++a)
      {
        int l = 0, mul = 1;
        for (int d = 0; d < dim; ++d)
          {
            l += t.local_lex[a * 3 + d] * mul;
            mul *= (p + 1);
          }
        t.lex2hier[l] = a;
      }
    // gradient of the vertex shape function v (MappingQ1) at xi
    const int nv          = t.nv;
    auto      vertex_grad = [&](int v, const double *xi, double *g) {
      for (int k = 0; k < dim; ++k)
        {
          double val = 1;
          for (int l = 0; l < dim; ++l)
            {
              const bool hi = (v >> l) & 1;
              val *= (l == k) ? (hi ? 1.0 : -1.0) : (hi ? xi[l] : 1.0 - xi[l]);
            }
          g[k] = val;
        }
    };
    t.hdphi.assign(size_t(nq) * nv * dim, 0.);
    t.hdphif.assign(size_t(2) * dim * nqf * nv * dim, 0.);
    t.hdphip.assign(size_t(npc) * nv * dim, 0.);
    for (int pt = 0; pt < npc; ++pt)
      {
        double xi[3] = {0, 0, 0};
        int    rem   = pt;
        for (int d = 0; d < dim; ++d)
          {
            xi[d] = double(rem % (p + 1)) / p;
            rem /= (p + 1);
          }
        for (int v = 0; v < nv; ++v)
          vertex_grad(v, xi, &t.hdphip[(size_t(pt) * nv + v) * dim]);
      }
    t.hN.assign(nq * npc, 0.);
    t.hdN.assign(nq * npc * dim, 0.);
    t.hw.assign(nq, 0.);
    for (int q = 0; q < nq; ++q)
      {
        double xi[3] = {0, 0, 0}, w = 1;
        int    rem   = q;
        for (int d = 0; d < dim; ++d)
          {
            xi[d] = x1[rem % nq1];
            w *= w1[rem % nq1];
            rem /= nq1;
          }
        t.hw[q] = w;
        for (int v = 0; v < nv; ++v)
          vertex_grad(v, xi, &t.hdphi[(size_t(q) * nv + v) * dim]);
        for (int a = 0; a < npc; ++a)
          {
            t.hN[q * npc + a] = shape(a, xi);
            for (int k = 0; k < dim; ++k)
              {
                double v = 1;
                for (int d = 0; d < dim; ++d)
                  v *= (d == k) ? dlag(p, t.local_lex[a * 3 + d], xi[d]) :
                                  lag(p, t.local_lex[a * 3 + d], xi[d]);
                t.hdN[(q * npc + a) * dim + k] = v;
              }
          }
      }
    // faces: QProjector::project_to_face (standard orientation). 3D y-faces use (z,x) ordering.
    t.hNf.assign(2 * dim * nqf * npc, 0.);
    t.hwf.assign(nqf, 0.);
    for (int q = 0; q < nqf; ++q)
      {
        double fq[2] = {0, 0}, w = 1;
        int    rem   = q;
        for (int d = 0; d < dim - 1; ++d)
          {
            fq[d] = x1[rem % nq1];
            w *= w1[rem % nq1];
            rem /= nq1;
          }
        t.hwf[q] = w;
        for (int f = 0; f < 2 * dim; ++f)
          {
            const int    d  = f / 2;
            const double cc = f % 2;
            double       xi[3] = {0, 0, 0};
            if (dim == 2)
              {
                xi[d]     = cc;
                xi[1 - d] = fq[0];
              }
            else if (d == 0)
              {
                xi[0] = cc;
                xi[1] = fq[0];
                xi[2] = fq[1];
              }
            else if (d == 1)
              {
                xi[1] = cc;
                xi[2] = fq[0];
                xi[0] = fq[1];
              }
            else
              {
                xi[2] = cc;
                xi[0] = fq[0];
                xi[1] = fq[1];
              }
            for (int a = 0; a < npc; ++a)
              t.hNf[(f * nqf + q) * npc + a] = shape(a, xi);
            for (int v = 0; v < nv; ++v)
              vertex_grad(v, xi, &t.hdphif[((size_t(f) * nqf + q) * nv + v) * dim]);
          }
      }
    t.hMref.assign(npc * npc, 0.);
    for (int a = 0; a < npc; ++a)
      for (int b = 0; b < npc; ++b)
        {
          double s = 0;
          for (int q = 0; q < nq; ++q)
            s += t.hw[q] * t.hN[q * npc + a] * t.hN[q * npc + b];
          t.hMref[a * npc + b] = s;
        }
  }
} // namespace gf
