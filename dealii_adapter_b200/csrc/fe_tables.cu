// Reference-cell tables: FE_Q(p) shape values / unit-cell gradients in deal.II's hierarchical
// order at QGauss points, face tables in QProjector order. These replace FEValues /
// FEFaceValues (reference call sites: nonlinear_elasticity.cc:668-669,891,902-906,807,815;
// linear_elasticity.cc:254-258,280,467-470,493,499).
//   quadrature: nonlinear QGauss(p+2) (nonlinear_elasticity.cc:74-75), linear QGauss(p+1)
//   (linear_elasticity.cc:61,252,465).
#include <cmath>

#include "gf_context.h"

namespace gf
{
  namespace
  {
    void gauss01(int n, std::vector<double> &x, std::vector<double> &w)
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
    double lag(int p, int i, double x)
    {
      if (p == 1)
        return i == 0 ? 1.0 - x : x;
      return i == 0 ? 2.0 * (x - 0.5) * (x - 1.0) : (i == 1 ? -4.0 * x * (x - 1.0) : 2.0 * x * (x - 0.5));
    }
    double dlag(int p, int i, double x)
    {
      if (p == 1)
        return i == 0 ? -1.0 : 1.0;
      return i == 0 ? 4.0 * x - 3.0 : (i == 1 ? -8.0 * x + 4.0 : 4.0 * x - 1.0);
    }
    // FE_Q hierarchical node -> lexicographic coordinates (vertices, lines, quads, hex), p <= 2
    void local_nodes(int dim, int p, std::vector<int> &lex)
    {
      lex.clear();
      auto push = [&](int x, int y, int z) {
        lex.push_back(x);
        lex.push_back(y);
        lex.push_back(z);
      };
      for (int v = 0; v < (1 << dim); ++v)
        push((v & 1) * p, ((v >> 1) & 1) * p, dim == 3 ? ((v >> 2) & 1) * p : 0);
      if (p < 2)
        return;
      if (dim == 2)
        {
          push(0, 1, 0);
          push(2, 1, 0);
          push(1, 0, 0);
          push(1, 2, 0);
          push(1, 1, 0);
          return;
        }
      for (int z = 0; z <= 2; z += 2)
        {
          push(0, 1, z);
          push(2, 1, z);
          push(1, 0, z);
          push(1, 2, z);
        }
      push(0, 0, 1);
      push(2, 0, 1);
      push(0, 2, 1);
      push(2, 2, 1);
      push(0, 1, 1);
      push(2, 1, 1);
      push(1, 0, 1);
      push(1, 2, 1);
      push(1, 1, 0);
      push(1, 1, 2);
      push(1, 1, 1);
    }
  } // namespace

  void build_tables(gf_context &c)
  {
    FETables &t = c.tables;
    t.dim       = c.dim;
    t.p         = c.p;
    t.nv        = 1 << c.dim;
    t.npc       = 1;
    for (int d = 0; d < c.dim; ++d)
      t.npc *= (c.p + 1);
    t.dpc = t.npc * c.dim;
    t.nq1 = c.model == GF_MODEL_NEO_HOOKEAN ? c.p + 2 : c.p + 1;
    t.nq  = 1;
    for (int d = 0; d < c.dim; ++d)
      t.nq *= t.nq1;
    t.nqf = t.nq / t.nq1;
    local_nodes(c.dim, c.p, t.local_lex);
    const int dim = c.dim, p = c.p, npc = t.npc, nq = t.nq, nq1 = t.nq1, nqf = t.nqf;
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
          }
      }
    std::vector<double> mref(npc * npc, 0.);
    for (int a = 0; a < npc; ++a)
      for (int b = 0; b < npc; ++b)
        {
          double s = 0;
          for (int q = 0; q < nq; ++q)
            s += t.hw[q] * t.hN[q * npc + a] * t.hN[q * npc + b];
          mref[a * npc + b] = s;
        }
    t.N.upload(t.hN.data(), t.hN.size(), c.stream);
    t.dN.upload(t.hdN.data(), t.hdN.size(), c.stream);
    t.w.upload(t.hw.data(), t.hw.size(), c.stream);
    t.Nf.upload(t.hNf.data(), t.hNf.size(), c.stream);
    t.wf.upload(t.hwf.data(), t.hwf.size(), c.stream);
    t.Mref.upload(mref.data(), mref.size(), c.stream);
  }
} // namespace gf
