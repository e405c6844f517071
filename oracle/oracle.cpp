// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h). Pinned in part by the reference's own code run here (oracle/_ref, tests/test_reference_pins.py); the deal.II-internal parts stay unpinned.
//
// Literal CPU/FP64 restatement of the dealii-adapter structural hot path. File:line citations are
// relative to /root/reference. deal.II (v9.5.0, CI pin .github/workflows/building.yml:27) is not
// in tree; its semantics are restated from the published algorithms and marked [deal.II].
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace
{
  thread_local std::string g_last_error;

  // ------------------------------------------------------------------------------------------
  // [deal.II] SymmetricTensor<2,dim> / <4,dim> with its storage order (00,11,[22],01,[02,12])
  // ------------------------------------------------------------------------------------------
  template <int dim>
  struct SymIndex
  {
    static constexpr int n = dim * (dim + 1) / 2;
    static int           of(int i, int j)
    {
      if (dim == 2)
        {
          static const int t[2][2] = {{0, 2}, {2, 1}};
          return t[i][j];
        }
      static const int t[3][3] = {{0, 3, 4}, {3, 1, 5}, {4, 5, 2}};
      return t[i][j];
    }
  };

  template <int dim>
  struct Sym2
  {
    static constexpr int n = SymIndex<dim>::n;
    double               d[n];
    Sym2()
    {
      for (int i = 0; i < n; ++i)
        d[i] = 0.;
    }
    double  operator()(int i, int j) const { return d[SymIndex<dim>::of(i, j)]; }
    double &operator()(int i, int j) { return d[SymIndex<dim>::of(i, j)]; }
  };
  template <int dim>
  struct Sym4
  {
    static constexpr int n = SymIndex<dim>::n;
    double               d[n][n];
    Sym4()
    {
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          d[i][j] = 0.;
    }
  };
  template <int dim>
  struct Ten2
  {
    double d[dim][dim];
    Ten2()
    {
      for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j)
          d[i][j] = 0.;
    }
  };
  template <int dim>
  struct Ten1
  {
    double d[dim];
    Ten1()
    {
      for (int i = 0; i < dim; ++i)
        d[i] = 0.;
    }
  };

  template <int dim>
  Sym2<dim> operator+(const Sym2<dim> &a, const Sym2<dim> &b)
  {
    Sym2<dim> r;
    for (int i = 0; i < Sym2<dim>::n; ++i)
      r.d[i] = a.d[i] + b.d[i];
    return r;
  }
  template <int dim>
  Sym2<dim> operator*(double s, const Sym2<dim> &a)
  {
    Sym2<dim> r;
    for (int i = 0; i < Sym2<dim>::n; ++i)
      r.d[i] = s * a.d[i];
    return r;
  }
  template <int dim>
  Sym4<dim> operator+(const Sym4<dim> &a, const Sym4<dim> &b)
  {
    Sym4<dim> r;
    for (int i = 0; i < Sym4<dim>::n; ++i)
      for (int j = 0; j < Sym4<dim>::n; ++j)
        r.d[i][j] = a.d[i][j] + b.d[i][j];
    return r;
  }
  template <int dim>
  Sym4<dim> operator-(const Sym4<dim> &a, const Sym4<dim> &b)
  {
    Sym4<dim> r;
    for (int i = 0; i < Sym4<dim>::n; ++i)
      for (int j = 0; j < Sym4<dim>::n; ++j)
        r.d[i][j] = a.d[i][j] - b.d[i][j];
    return r;
  }
  template <int dim>
  Sym4<dim> operator*(double s, const Sym4<dim> &a)
  {
    Sym4<dim> r;
    for (int i = 0; i < Sym4<dim>::n; ++i)
      for (int j = 0; j < Sym4<dim>::n; ++j)
        r.d[i][j] = s * a.d[i][j];
    return r;
  }
  // [deal.II] double contraction SymmetricTensor<4> * SymmetricTensor<2>
  template <int dim>
  Sym2<dim> operator*(const Sym4<dim> &t, const Sym2<dim> &s)
  {
    Sym2<dim> r;
    for (int i = 0; i < Sym2<dim>::n; ++i)
      {
        double sum = 0.;
        for (int j = 0; j < dim; ++j)
          sum += t.d[i][j] * s.d[j];
        for (int j = dim; j < Sym2<dim>::n; ++j)
          sum += 2.0 * t.d[i][j] * s.d[j];
        r.d[i] = sum;
      }
    return r;
  }
  // [deal.II] double contraction SymmetricTensor<2> * SymmetricTensor<4>
  template <int dim>
  Sym2<dim> operator*(const Sym2<dim> &s, const Sym4<dim> &t)
  {
    Sym2<dim> r;
    for (int j = 0; j < Sym2<dim>::n; ++j)
      {
        double sum = 0.;
        for (int i = 0; i < dim; ++i)
          sum += s.d[i] * t.d[i][j];
        for (int i = dim; i < Sym2<dim>::n; ++i)
          sum += 2.0 * s.d[i] * t.d[i][j];
        r.d[j] = sum;
      }
    return r;
  }
  // [deal.II] scalar product of two symmetric rank-2 tensors
  template <int dim>
  double operator*(const Sym2<dim> &a, const Sym2<dim> &b)
  {
    double sum = 0.;
    for (int i = 0; i < dim; ++i)
      sum += a.d[i] * b.d[i];
    for (int i = dim; i < Sym2<dim>::n; ++i)
      sum += 2.0 * a.d[i] * b.d[i];
    return sum;
  }
  // [deal.II] dev_P * c_bar * dev_P style rank-4 * rank-4 (needed literally for material.h:131-132)
  template <int dim>
  Sym4<dim> operator*(const Sym4<dim> &a, const Sym4<dim> &b)
  {
    Sym4<dim> r;
    for (int i = 0; i < Sym4<dim>::n; ++i)
      for (int j = 0; j < Sym4<dim>::n; ++j)
        {
          double sum = 0.;
          for (int k = 0; k < dim; ++k)
            sum += a.d[i][k] * b.d[k][j];
          for (int k = dim; k < Sym4<dim>::n; ++k)
            sum += 2.0 * a.d[i][k] * b.d[k][j];
          r.d[i][j] = sum;
        }
    return r;
  }
  template <int dim>
  Sym4<dim> outer_product(const Sym2<dim> &a, const Sym2<dim> &b)
  {
    Sym4<dim> r;
    for (int i = 0; i < Sym4<dim>::n; ++i)
      for (int j = 0; j < Sym4<dim>::n; ++j)
        r.d[i][j] = a.d[i] * b.d[j];
    return r;
  }
  template <int dim>
  double trace(const Sym2<dim> &a)
  {
    double t = a.d[0];
    for (int i = 1; i < dim; ++i)
      t += a.d[i];
    return t;
  }

  // [deal.II] Physics::Elasticity::StandardTensors<dim>
  template <int dim>
  struct StandardTensors
  {
    Sym2<dim> I;
    Sym4<dim> S, IxI, dev_P;
    StandardTensors()
    {
      for (int i = 0; i < dim; ++i)
        I.d[i] = 1.0;
      for (int i = 0; i < dim; ++i)
        S.d[i][i] = 1.0;
      for (int i = dim; i < Sym4<dim>::n; ++i)
        S.d[i][i] = 0.5;
      IxI   = outer_product(I, I);
      dev_P = S - (1.0 / dim) * IxI;
    }
  };

  template <int dim>
  double determinant(const Ten2<dim> &t)
  {
    if (dim == 2)
      return t.d[0][0] * t.d[1][1] - t.d[1][0] * t.d[0][1];
    return t.d[0][0] * (t.d[1][1] * t.d[2 % dim][2 % dim] - t.d[1][2 % dim] * t.d[2 % dim][1]) -
           t.d[0][1] * (t.d[1][0] * t.d[2 % dim][2 % dim] - t.d[1][2 % dim] * t.d[2 % dim][0]) +
           t.d[0][2 % dim] * (t.d[1][0] * t.d[2 % dim][1] - t.d[1][1] * t.d[2 % dim][0]);
  }
  template <int dim>
  Ten2<dim> invert(const Ten2<dim> &t)
  {
    Ten2<dim>    r;
    const double inv_det = 1.0 / determinant(t);
    if (dim == 2)
      {
        r.d[0][0] = t.d[1][1] * inv_det;
        r.d[0][1] = -t.d[0][1] * inv_det;
        r.d[1][0] = -t.d[1][0] * inv_det;
        r.d[1][1] = t.d[0][0] * inv_det;
        return r;
      }
    constexpr int X = 0, Y = 1, Z = 2 % dim;
    r.d[X][X] = (t.d[Y][Y] * t.d[Z][Z] - t.d[Y][Z] * t.d[Z][Y]) * inv_det;
    r.d[X][Y] = (t.d[X][Z] * t.d[Z][Y] - t.d[X][Y] * t.d[Z][Z]) * inv_det;
    r.d[X][Z] = (t.d[X][Y] * t.d[Y][Z] - t.d[X][Z] * t.d[Y][Y]) * inv_det;
    r.d[Y][X] = (t.d[Y][Z] * t.d[Z][X] - t.d[Y][X] * t.d[Z][Z]) * inv_det;
    r.d[Y][Y] = (t.d[X][X] * t.d[Z][Z] - t.d[X][Z] * t.d[Z][X]) * inv_det;
    r.d[Y][Z] = (t.d[X][Z] * t.d[Y][X] - t.d[X][X] * t.d[Y][Z]) * inv_det;
    r.d[Z][X] = (t.d[Y][X] * t.d[Z][Y] - t.d[Y][Y] * t.d[Z][X]) * inv_det;
    r.d[Z][Y] = (t.d[X][Y] * t.d[Z][X] - t.d[X][X] * t.d[Z][Y]) * inv_det;
    r.d[Z][Z] = (t.d[X][X] * t.d[Y][Y] - t.d[X][Y] * t.d[Y][X]) * inv_det;
    return r;
  }
  template <int dim>
  Ten2<dim> matmul(const Ten2<dim> &a, const Ten2<dim> &b)
  {
    Ten2<dim> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        {
          double s = 0.;
          for (int k = 0; k < dim; ++k)
            s += a.d[i][k] * b.d[k][j];
          r.d[i][j] = s;
        }
    return r;
  }
  template <int dim>
  Ten2<dim> transpose(const Ten2<dim> &a)
  {
    Ten2<dim> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        r.d[i][j] = a.d[j][i];
    return r;
  }
  template <int dim>
  Sym2<dim> symmetrize(const Ten2<dim> &a)
  {
    Sym2<dim> r;
    for (int i = 0; i < dim; ++i)
      r(i, i) = a.d[i][i];
    for (int i = 0; i < dim; ++i)
      for (int j = i + 1; j < dim; ++j)
        r(i, j) = (a.d[i][j] + a.d[j][i]) / 2.0;
    return r;
  }

  // ------------------------------------------------------------------------------------------
  // compressible_neo_hook_material.h:13-139
  // ------------------------------------------------------------------------------------------
  template <int dim>
  struct Material
  {
    double               kappa, c_1, rho;
    StandardTensors<dim> st;
    Material(double mu, double nu, double rho_)
      : kappa((2.0 * mu * (1.0 + nu)) / (3.0 * (1.0 - 2.0 * nu))) // :20
      , c_1(mu / 2.0)                                             // :21
      , rho(rho_)
    {}
    // :62-72
    double get_Psi(double det_F, const Sym2<dim> &b_bar) const
    {
      return (kappa / 4.0) * (det_F * det_F - 1.0 - 2.0 * std::log(det_F)) +
             c_1 * (trace(b_bar) - dim);
    }
    double    get_dPsi_vol_dJ(double det_F) const { return (kappa / 2.0) * (det_F - 1.0 / det_F); } // :74-78
    Sym2<dim> get_tau_vol(double det_F) const { return (get_dPsi_vol_dJ(det_F) * det_F) * st.I; }   // :80-85
    Sym2<dim> get_tau_bar(const Sym2<dim> &b_bar) const { return (2.0 * c_1) * b_bar; }            // :94-98
    Sym2<dim> get_tau_iso(const Sym2<dim> &b_bar) const { return st.dev_P * get_tau_bar(b_bar); }  // :87-92
    double    get_d2Psi_vol_dJ2(double det_F) const
    {
      return ((kappa / 2.0) * (1.0 + 1.0 / (det_F * det_F))); // :100-104
    }
    // :106-114
    Sym4<dim> get_Jc_vol(double det_F) const
    {
      return det_F * ((get_dPsi_vol_dJ(det_F) + det_F * get_d2Psi_vol_dJ2(det_F)) * st.IxI -
                      (2.0 * get_dPsi_vol_dJ(det_F)) * st.S);
    }
    // :116-133 (c_bar == 0, :135-139)
    Sym4<dim> get_Jc_iso(const Sym2<dim> &b_bar) const
    {
      const Sym2<dim> tau_bar     = get_tau_bar(b_bar);
      const Sym2<dim> tau_iso     = get_tau_iso(b_bar);
      const Sym4<dim> tau_iso_x_I = outer_product(tau_iso, st.I);
      const Sym4<dim> I_x_tau_iso = outer_product(st.I, tau_iso);
      const Sym4<dim> c_bar; // zero
      return ((2.0 / dim) * trace(tau_bar)) * st.dev_P - (2.0 / dim) * (tau_iso_x_I + I_x_tau_iso) +
             st.dev_P * c_bar * st.dev_P;
    }
    Sym2<dim> get_tau(double det_F, const Sym2<dim> &b_bar) const // :37-42
    {
      return get_tau_vol(det_F) + get_tau_iso(b_bar);
    }
    Sym4<dim> get_Jc(double det_F, const Sym2<dim> &b_bar) const // :44-49
    {
      return get_Jc_vol(det_F) + get_Jc_iso(b_bar);
    }
  };

  // ------------------------------------------------------------------------------------------
  // [deal.II] FE_Q(p) (Gauss-Lobatto support points, equidistant for p<=2), QGauss, QProjector
  // ------------------------------------------------------------------------------------------
  void gauss_legendre_01(int n, std::vector<double> &x, std::vector<double> &w)
  {
    // [deal.II] QGauss<1>(n): Newton iteration on Legendre P_n, mapped to [0,1], ascending
    x.assign(n, 0.);
    w.assign(n, 0.);
    const int m = (n + 1) / 2;
    for (int i = 1; i <= m; ++i)
      {
        long double z = std::cos(M_PIl * (i - 0.25L) / (n + 0.5L));
        long double pp, p1, p2, p3;
        do
          {
            p1 = 1.0L;
            p2 = 0.0L;
            for (int j = 0; j < n; ++j)
              {
                p3 = p2;
                p2 = p1;
                p1 = ((2.0L * j + 1.0L) * z * p2 - j * p3) / (j + 1);
              }
            pp = n * (z * p1 - p2) / (z * z - 1);
            z  = z - p1 / pp;
          }
        while (std::fabs((double)(p1 / pp)) > 1e-19);
        const double xx = double(0.5L * z);
        x[i - 1]        = 0.5 - xx;
        x[n - i]        = 0.5 + xx;
        const double ww = double(1.0L / ((1.0L - z * z) * pp * pp));
        w[i - 1] = w[n - i] = ww;
      }
  }
  inline double lagrange(int p, int i, double x)
  {
    if (p == 1)
      return i == 0 ? 1.0 - x : x;
    switch (i)
      {
        case 0: return 2.0 * (x - 0.5) * (x - 1.0);
        case 1: return -4.0 * x * (x - 1.0);
        default: return 2.0 * x * (x - 0.5);
      }
  }
  inline double dlagrange(int p, int i, double x)
  {
    if (p == 1)
      return i == 0 ? -1.0 : 1.0;
    switch (i)
      {
        case 0: return 4.0 * x - 3.0;
        case 1: return -8.0 * x + 4.0;
        default: return 4.0 * x - 1.0;
      }
  }
  // [deal.II] FE_Q(p), p >= 3: Lagrange polynomials on the Gauss-Lobatto points (QGaussLobatto(p+1)
  // mapped to [0,1]): 0, 1 and the roots of P'_p, each bracketed by two neighbouring Gauss points
  // (the roots of P_p interlace those of P'_p) and found by bisection
  std::vector<double> gauss_lobatto_01(int n)
  {
    const int           p = n - 1;
    std::vector<double> pts(n, 0.0), gx, gw;
    pts[n - 1] = 1.0;
    if (p < 2)
      return pts;
    gauss_legendre_01(p, gx, gw);
    auto dlegendre = [p](long double z) { // P'_p(z), z in (-1, 1)
      long double p0 = 1.0L, p1 = z;
      for (int k = 2; k <= p; ++k)
        {
          const long double pk = ((2 * k - 1) * z * p1 - (k - 1) * p0) / k;
          p0                   = p1;
          p1                   = pk;
        }
      return p * (z * p1 - p0) / (z * z - 1.0L);
    };
    for (int i = 1; i < p; ++i)
      {
        long double lo = 2.0L * gx[i - 1] - 1.0L, hi = 2.0L * gx[i] - 1.0L;
        const bool  lo_positive = dlegendre(lo) > 0;
        for (int it = 0; it < 200 && hi - lo > 1e-19L; ++it)
          {
            const long double mid = 0.5L * (lo + hi);
            if ((dlegendre(mid) > 0) == lo_positive)
              lo = mid;
            else
              hi = mid;
          }
        pts[i] = double(0.5L + 0.25L * (lo + hi));
      }
    return pts;
  }
  inline double lagrange_on(const std::vector<double> &nodes, int i, double x)
  {
    double v = 1.0;
    for (size_t j = 0; j < nodes.size(); ++j)
      if (int(j) != i)
        v *= (x - nodes[j]) / (nodes[i] - nodes[j]);
    return v;
  }
  inline double dlagrange_on(const std::vector<double> &nodes, int i, double x)
  {
    double s = 0.0;
    for (size_t k = 0; k < nodes.size(); ++k)
      if (int(k) != i)
        {
          double v = 1.0 / (nodes[i] - nodes[k]);
          for (size_t j = 0; j < nodes.size(); ++j)
            if (int(j) != i && j != k)
              v *= (x - nodes[j]) / (nodes[i] - nodes[j]);
          s += v;
        }
    return s;
  }
  // [deal.II] hierarchical FE_Q local node -> lexicographic (lx,ly,lz): vertices, lines, quads,
  // hex; the interior nodes of an entity in the entity's own lexicographic order
  // (FETools::hierarchic_to_lexicographic_numbering; y-faces run z fastest); SURVEY 8a
  void build_local_nodes(int dim, int p, std::vector<int> &lex)
  {
    lex.clear();
    auto push = [&](int x, int y, int z) {
      lex.push_back(x);
      lex.push_back(y);
      lex.push_back(z);
    };
    const int nv = 1 << dim;
    for (int v = 0; v < nv; ++v)
      push((v & 1) * p, ((v >> 1) & 1) * p, dim == 3 ? ((v >> 2) & 1) * p : 0);
    if (p < 2)
      return;
    if (dim == 2)
      {
        for (int m = 1; m < p; ++m)
          push(0, m, 0);
        for (int m = 1; m < p; ++m)
          push(p, m, 0);
        for (int m = 1; m < p; ++m)
          push(m, 0, 0);
        for (int m = 1; m < p; ++m)
          push(m, p, 0);
        for (int my = 1; my < p; ++my)
          for (int mx = 1; mx < p; ++mx)
            push(mx, my, 0);
      }
    else
      {
        for (int z = 0; z <= p; z += p)
          {
            for (int m = 1; m < p; ++m)
              push(0, m, z);
            for (int m = 1; m < p; ++m)
              push(p, m, z);
            for (int m = 1; m < p; ++m)
              push(m, 0, z);
            for (int m = 1; m < p; ++m)
              push(m, p, z);
          }
        for (int m = 1; m < p; ++m)
          push(0, 0, m);
        for (int m = 1; m < p; ++m)
          push(p, 0, m);
        for (int m = 1; m < p; ++m)
          push(0, p, m);
        for (int m = 1; m < p; ++m)
          push(p, p, m);
        for (int x = 0; x <= p; x += p)
          for (int mz = 1; mz < p; ++mz)
            for (int my = 1; my < p; ++my)
              push(x, my, mz);
        for (int y = 0; y <= p; y += p)
          for (int mx = 1; mx < p; ++mx)
            for (int mz = 1; mz < p; ++mz)
              push(mx, y, mz);
        for (int z = 0; z <= p; z += p)
          for (int my = 1; my < p; ++my)
            for (int mx = 1; mx < p; ++mx)
              push(mx, my, z);
        for (int mz = 1; mz < p; ++mz)
          for (int my = 1; my < p; ++my)
            for (int mx = 1; mx < p; ++mx)
              push(mx, my, mz);
      }
  }
  // [deal.II] FESystem(FE_Q(p), dim)::system_to_component_index: entity by entity, inside an
  // entity all DoFs of component 0, then component 1, ... (FESystem::build_cell_tables). For
  // p <= 2 (one scalar DoF per entity) this is i -> (i / dim, i % dim).
  void build_system_numbering(int dim, int p, std::vector<int> &node_of, std::vector<int> &comp_of,
                              std::vector<int> &loc_of)
  {
    node_of.clear();
    comp_of.clear();
    std::vector<int> entity_dofs;
    entity_dofs.insert(entity_dofs.end(), size_t(1) << dim, 1);
    entity_dofs.insert(entity_dofs.end(), dim == 2 ? 4 : 12, p - 1);
    entity_dofs.insert(entity_dofs.end(), dim == 2 ? 1 : 6, (p - 1) * (p - 1));
    if (dim == 3)
      entity_dofs.push_back((p - 1) * (p - 1) * (p - 1));
    int first = 0;
    for (int count : entity_dofs)
      {
        for (int c = 0; c < dim; ++c)
          for (int k = 0; k < count; ++k)
            {
              node_of.push_back(first + k);
              comp_of.push_back(c);
            }
        first += count;
      }
    loc_of.assign(node_of.size(), -1);
    for (size_t i = 0; i < node_of.size(); ++i)
      loc_of[node_of[i] * dim + comp_of[i]] = int(i);
  }

  struct CSR
  {
    int64_t              n = 0;
    std::vector<int64_t> rowptr;
    std::vector<int32_t> col;  // ascending
    std::vector<int64_t> diag; // position of the diagonal
    int64_t              find(int64_t r, int32_t c) const
    {
      const int32_t *b = col.data() + rowptr[r], *e = col.data() + rowptr[r + 1];
      const int32_t *p = std::lower_bound(b, e, c);
      return (p != e && *p == c) ? (p - col.data()) : -1;
    }
  };

  // [deal.II] SparseMatrix::vmult in stored order (diagonal first, then ascending columns)
  void vmult(const CSR &A, const std::vector<double> &val, const double *x, double *y)
  {
    for (int64_t r = 0; r < A.n; ++r)
      {
        double s = val[A.diag[r]] * x[r];
        for (int64_t j = A.rowptr[r]; j < A.rowptr[r + 1]; ++j)
          if (j != A.diag[r])
            s += val[j] * x[A.col[j]];
        y[r] = s;
      }
  }
  // [deal.II] SparseMatrix::precondition_SSOR : (D+wU)^-1 w(2-w) D (D+wL)^-1
  void precondition_SSOR(const CSR &A, const std::vector<double> &val, double *dst,
                         const double *src, double om)
  {
    const int64_t n = A.n;
    for (int64_t row = 0; row < n; ++row)
      {
        dst[row] = src[row];
        double s = 0;
        for (int64_t j = A.rowptr[row]; j < A.diag[row]; ++j)
          s += val[j] * dst[A.col[j]];
        dst[row] -= s * om;
        dst[row] /= val[A.diag[row]];
      }
    for (int64_t row = 0; row < n; ++row)
      dst[row] *= om * (2. - om) * val[A.diag[row]];
    for (int64_t row = n - 1; row >= 0; --row)
      {
        double s = 0;
        for (int64_t j = A.diag[row] + 1; j < A.rowptr[row + 1]; ++j)
          s += val[j] * dst[A.col[j]];
        dst[row] -= s * om;
        dst[row] /= val[A.diag[row]];
      }
  }
  double l2_norm(const std::vector<double> &v)
  {
    double s = 0;
    for (double x : v)
      s += x * x;
    return std::sqrt(s);
  }
  // [deal.II 9.5] SolverCG::solve with a generic preconditioner + SolverControl::check
  // returns 0 success, 1 failure (NoConvergence)
  int solver_cg_ssor(const CSR &A, const std::vector<double> &val, std::vector<double> &x,
                     const std::vector<double> &b, double om, int maxsteps, double tol,
                     uint32_t &last_step, double &last_value)
  {
    const int64_t       n = A.n;
    std::vector<double> r(n), p(n), v(n), z(n);
    bool                all_zero = true;
    for (double xi : x)
      if (xi != 0.)
        {
          all_zero = false;
          break;
        }
    if (!all_zero)
      {
        vmult(A, val, x.data(), r.data());
        for (int64_t i = 0; i < n; ++i)
          r[i] = -r[i] + b[i];
      }
    else
      r = b;
    double residual_norm = l2_norm(r);
    int    it            = 0;
    last_step            = 0;
    last_value           = residual_norm;
    if (residual_norm <= tol)
      return 0;
    if (it >= maxsteps)
      return 1;
    double r_dot_preconditioner_dot_r = 0;
    while (true)
      {
        ++it;
        const double previous_r_dot_preconditioner_dot_r = r_dot_preconditioner_dot_r;
        precondition_SSOR(A, val, z.data(), r.data(), om);
        double rz = 0;
        for (int64_t i = 0; i < n; ++i)
          rz += r[i] * z[i];
        r_dot_preconditioner_dot_r = rz;
        if (it > 1)
          {
            const double beta = r_dot_preconditioner_dot_r / previous_r_dot_preconditioner_dot_r;
            for (int64_t i = 0; i < n; ++i)
              p[i] = beta * p[i] + z[i];
          }
        else
          p = z;
        vmult(A, val, p.data(), v.data());
        double pAp = 0;
        for (int64_t i = 0; i < n; ++i)
          pAp += p[i] * v[i];
        const double alpha = r_dot_preconditioner_dot_r / pAp;
        for (int64_t i = 0; i < n; ++i)
          x[i] += alpha * p[i];
        double rr = 0;
        for (int64_t i = 0; i < n; ++i)
          {
            r[i] -= alpha * v[i];
            rr += r[i] * r[i];
          }
        residual_norm = std::sqrt(std::fabs(rr));
        last_step     = it;
        last_value    = residual_norm;
        if (residual_norm <= tol)
          return 0;
        if (it >= maxsteps || std::isnan(residual_norm))
          return 1;
      }
  }

  // ------------------------------------------------------------------------------------------
  struct ContextBase
  {
    virtual ~ContextBase() {}
    orc_desc             desc;
    int                  rt_dim, p, npc, dpc, nq1, nq, nqf1, nqf;
    std::vector<int32_t> cell_dofs;
    std::vector<double>  cell_vertices;
    std::vector<uint8_t> constrained;
    std::vector<int32_t> iface_cell, iface_face_no, iface_dofs;
    CSR                  pattern;
    std::vector<std::vector<double>> mats; // by ORC_MAT_*
    std::vector<std::vector<double>> vecs; // by vector id
    std::vector<std::vector<double>> old_state_data;
    // reference-cell tables
    std::vector<int>    local_lex;           // npc*3
    std::vector<int>    node_of, comp_of;    // system_to_component_index of local DoF i
    std::vector<int>    loc_of;              // local DoF of (node a, component c): [a*dim + c]
    std::vector<double> support_1d;          // FE_Q(p) support points on [0,1]
    std::vector<double> qx, qw;              // cell quadrature (unit cell) nq*dim, nq
    std::vector<double> N, dN;               // [nq*npc], [nq*npc*dim] reference gradients
    std::vector<double> fqx, fqw;            // face quadrature on unit face, nqf*(dim-1), nqf
    std::vector<double> Nf;                  // [2*dim][nqf][npc]
    std::vector<double> dphi_v;              // Q1 vertex basis gradients at cell q: [nq][nv][dim]
    std::vector<double> dphi_vf;             // at face q: [2*dim][nqf][nv][dim]
    std::vector<std::vector<int>> faces_of_cell_start; // unused
    std::vector<int64_t> iface_ptr;          // per cell -> range into sorted interface faces
    std::vector<int32_t> iface_sorted_face;  // face numbers grouped by cell

    std::vector<double> &V(int which) { return vecs[which]; }
  };

  template <int dim>
  struct Context : ContextBase
  {
    static constexpr int DIM = dim;
    Material<dim> material;
    double        alpha_1, alpha_2, alpha_3, alpha_4, alpha_5, alpha_6;
    Context(const orc_desc &d)
      : material(d.mu, d.nu, d.rho)
    {
      desc = d;
      // nonlinear_elasticity.h:242-250
      alpha_1 = 1. / (d.beta * std::pow(d.delta_t, 2));
      alpha_2 = 1. / (d.beta * d.delta_t);
      alpha_3 = (1 - (2 * d.beta)) / (2 * d.beta);
      alpha_4 = d.gamma / (d.beta * d.delta_t);
      alpha_5 = 1 - (d.gamma / d.beta);
      alpha_6 = (1 - (d.gamma / (2 * d.beta))) * d.delta_t;
      setup();
    }

    void setup()
    {
      const orc_desc &d = desc;
      rt_dim            = dim;
      p                 = d.degree;
      if (p < 1 || p > 8)
        throw std::invalid_argument("oracle: degree must be 1..8");
      npc = 1;
      for (int k = 0; k < dim; ++k)
        npc *= (p + 1);
      dpc = npc * dim;
      // quadrature orders: nonlinear QGauss(p+2) (nonlinear_elasticity.cc:74-75),
      //                    linear QGauss(p+1) (linear_elasticity.cc:61,252,465)
      nq1  = d.model == 1 ? p + 2 : p + 1;
      nqf1 = nq1;
      nq   = 1;
      for (int k = 0; k < dim; ++k)
        nq *= nq1;
      nqf = nq / nq1;
      cell_dofs.assign(d.cell_dofs, d.cell_dofs + d.n_cells * dpc);
      const int nv = 1 << dim;
      cell_vertices.assign(d.cell_vertices, d.cell_vertices + d.n_cells * nv * dim);
      constrained.assign(d.constrained, d.constrained + d.n_dofs);
      iface_cell.assign(d.iface_cell, d.iface_cell + d.n_iface_faces);
      iface_face_no.assign(d.iface_face_no, d.iface_face_no + d.n_iface_faces);
      iface_dofs.assign(d.iface_dofs, d.iface_dofs + d.n_iface_nodes * dim);
      build_local_nodes(dim, p, local_lex);
      build_system_numbering(dim, p, node_of, comp_of, loc_of);
      support_1d = gauss_lobatto_01(p + 1);
      build_tables();
      build_pattern();
      mats.assign(5, std::vector<double>());
      if (d.model == 1)
        mats[ORC_MAT_TANGENT].assign(pattern.col.size(), 0.);
      else
        for (int m = ORC_MAT_STIFFNESS; m <= ORC_MAT_SYSTEM; ++m)
          mats[m].assign(pattern.col.size(), 0.);
      vecs.assign(32, std::vector<double>());
      for (int v = 0; v < 32; ++v)
        vecs[v].assign(d.n_dofs, 0.);
      // interface faces grouped per cell, in face order (cell->face_iterators())
      iface_ptr.assign(d.n_cells + 1, 0);
      for (int64_t f = 0; f < d.n_iface_faces; ++f)
        iface_ptr[iface_cell[f] + 1]++;
      for (int64_t c = 0; c < d.n_cells; ++c)
        iface_ptr[c + 1] += iface_ptr[c];
      iface_sorted_face.resize(d.n_iface_faces);
      {
        std::vector<int64_t> fill(iface_ptr.begin(), iface_ptr.end() - 1);
        for (int64_t f = 0; f < d.n_iface_faces; ++f)
          iface_sorted_face[fill[iface_cell[f]]++] = iface_face_no[f];
        for (int64_t c = 0; c < d.n_cells; ++c)
          std::sort(iface_sorted_face.begin() + iface_ptr[c],
                    iface_sorted_face.begin() + iface_ptr[c + 1]);
      }
    }

    double shape(int a, const double *xi) const
    {
      double v = 1;
      for (int k = 0; k < dim; ++k)
        v *= p <= 2 ? lagrange(p, local_lex[a * 3 + k], xi[k]) :
                      lagrange_on(support_1d, local_lex[a * 3 + k], xi[k]);
      return v;
    }
    void shape_grad(int a, const double *xi, double *g) const
    {
      for (int k = 0; k < dim; ++k)
        {
          double v = 1;
          for (int l = 0; l < dim; ++l)
            {
              const int i = local_lex[a * 3 + l];
              if (p <= 2)
                v *= (l == k) ? dlagrange(p, i, xi[l]) : lagrange(p, i, xi[l]);
              else
                v *= (l == k) ? dlagrange_on(support_1d, i, xi[l]) :
                                lagrange_on(support_1d, i, xi[l]);
            }
          g[k] = v;
        }
    }
    // Q1 vertex basis (MappingQ1) gradient
    void vertex_grad(int v, const double *xi, double *g) const
    {
      for (int k = 0; k < dim; ++k)
        {
          double val = 1;
          for (int l = 0; l < dim; ++l)
            {
              const int bit = (v >> l) & 1;
              val *= (l == k) ? (bit ? 1.0 : -1.0) : (bit ? xi[l] : 1.0 - xi[l]);
            }
          g[k] = val;
        }
    }
    // [deal.II] QProjector<dim>::project_to_face for standard orientation
    void face_point(int face, const double *fq, double *xi) const
    {
      const int    d = face / 2;
      const double c = face % 2;
      if (dim == 2)
        {
          xi[d]     = c;
          xi[1 - d] = fq[0];
        }
      else
        {
          switch (d)
            {
              case 0: xi[0] = c; xi[1] = fq[0]; xi[2] = fq[1]; break;
              case 1: xi[1] = c; xi[2] = fq[0]; xi[0] = fq[1]; break; // (z,x) order on y-faces
              default: xi[2] = c; xi[0] = fq[0]; xi[1] = fq[1]; break;
            }
        }
    }

    void build_tables()
    {
      std::vector<double> x1, w1;
      gauss_legendre_01(nq1, x1, w1);
      qx.resize(nq * dim);
      qw.resize(nq);
      for (int q = 0; q < nq; ++q)
        {
          int    rem = q;
          double w   = 1;
          for (int k = 0; k < dim; ++k)
            {
              const int i     = rem % nq1;
              rem /= nq1;
              qx[q * dim + k] = x1[i];
              w *= w1[i];
            }
          qw[q] = w;
        }
      const int nv = 1 << dim;
      N.resize(nq * npc);
      dN.resize(nq * npc * dim);
      dphi_v.resize(nq * nv * dim);
      for (int q = 0; q < nq; ++q)
        {
          for (int a = 0; a < npc; ++a)
            {
              N[q * npc + a] = shape(a, &qx[q * dim]);
              shape_grad(a, &qx[q * dim], &dN[(q * npc + a) * dim]);
            }
          for (int v = 0; v < nv; ++v)
            vertex_grad(v, &qx[q * dim], &dphi_v[(q * nv + v) * dim]);
        }
      // faces
      gauss_legendre_01(nqf1, x1, w1);
      fqx.resize(nqf * (dim - 1));
      fqw.resize(nqf);
      for (int q = 0; q < nqf; ++q)
        {
          int    rem = q;
          double w   = 1;
          for (int k = 0; k < dim - 1; ++k)
            {
              const int i           = rem % nqf1;
              rem /= nqf1;
              fqx[q * (dim - 1) + k] = x1[i];
              w *= w1[i];
            }
          fqw[q] = w;
        }
      Nf.resize(2 * dim * nqf * npc);
      dphi_vf.resize(2 * dim * nqf * nv * dim);
      for (int f = 0; f < 2 * dim; ++f)
        for (int q = 0; q < nqf; ++q)
          {
            double xi[3];
            face_point(f, &fqx[q * (dim - 1)], xi);
            for (int a = 0; a < npc; ++a)
              Nf[(f * nqf + q) * npc + a] = shape(a, xi);
            for (int v = 0; v < nv; ++v)
              vertex_grad(v, xi, &dphi_vf[((f * nqf + q) * nv + v) * dim]);
          }
    }

    // [deal.II] DoFTools::make_sparsity_pattern: all dofs of a cell couple
    void build_pattern()
    {
      const int64_t                     n = desc.n_dofs;
      std::vector<std::vector<int32_t>> rows(n);
      for (int64_t c = 0; c < desc.n_cells; ++c)
        {
          const int32_t *ids = &cell_dofs[c * dpc];
          for (int i = 0; i < dpc; ++i)
            {
              auto &r = rows[ids[i]];
              r.insert(r.end(), ids, ids + dpc);
            }
          if ((c & 1023) == 1023)
            for (int i = 0; i < dpc; ++i)
              {
                auto &r = rows[ids[i]];
                std::sort(r.begin(), r.end());
                r.erase(std::unique(r.begin(), r.end()), r.end());
              }
        }
      pattern.n = n;
      pattern.rowptr.assign(n + 1, 0);
      for (int64_t r = 0; r < n; ++r)
        {
          auto &row = rows[r];
          std::sort(row.begin(), row.end());
          row.erase(std::unique(row.begin(), row.end()), row.end());
          if (row.empty())
            row.push_back(int32_t(r));
          pattern.rowptr[r + 1] = pattern.rowptr[r] + int64_t(row.size());
        }
      pattern.col.resize(pattern.rowptr[n]);
      pattern.diag.resize(n);
      for (int64_t r = 0; r < n; ++r)
        {
          std::copy(rows[r].begin(), rows[r].end(), pattern.col.begin() + pattern.rowptr[r]);
          pattern.diag[r] = pattern.find(r, int32_t(r));
          std::vector<int32_t>().swap(rows[r]);
        }
    }

    // [deal.II] FEValues::reinit with MappingQ1: per-q Jacobian, JxW and real-space gradients
    struct CellGeom
    {
      std::vector<double> JxW;    // nq
      std::vector<double> grad;   // [nq][npc][dim] real-space gradients of scalar shape fns
    };
    void jacobian(const double *verts, const double *dphi, double J[3][3]) const
    {
      const int nv = 1 << dim;
      for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j)
          {
            double s = 0;
            for (int v = 0; v < nv; ++v)
              s += verts[v * dim + i] * dphi[v * dim + j];
            J[i][j] = s;
          }
    }
    void reinit_cell(int64_t cell, CellGeom &g) const
    {
      const int nv = 1 << dim;
      g.JxW.resize(nq);
      g.grad.resize(nq * npc * dim);
      const double *verts = &cell_vertices[cell * nv * dim];
      for (int q = 0; q < nq; ++q)
        {
          double J[3][3];
          jacobian(verts, &dphi_v[q * nv * dim], J);
          Ten2<dim> Jt;
          for (int i = 0; i < dim; ++i)
            for (int j = 0; j < dim; ++j)
              Jt.d[i][j] = J[i][j];
          const double    det  = determinant(Jt);
          const Ten2<dim> Jinv = invert(Jt);
          g.JxW[q]             = det * qw[q];
          for (int a = 0; a < npc; ++a)
            for (int i = 0; i < dim; ++i)
              {
                // covariant transformation: grad_x = J^{-T} grad_xi
                double s = 0;
                for (int k = 0; k < dim; ++k)
                  s += Jinv.d[k][i] * dN[(q * npc + a) * dim + k];
                g.grad[(q * npc + a) * dim + i] = s;
              }
        }
    }
    // [deal.II] FEFaceValues::reinit: JxW on the face and outward unit normal
    void reinit_face(int64_t cell, int face, std::vector<double> &JxW,
                     std::vector<double> &normals) const
    {
      const int     nv    = 1 << dim;
      const double *verts = &cell_vertices[cell * nv * dim];
      JxW.resize(nqf);
      normals.resize(nqf * dim);
      const int    d    = face / 2;
      const double sign = face % 2 ? 1.0 : -1.0;
      for (int q = 0; q < nqf; ++q)
        {
          double J[3][3];
          jacobian(verts, &dphi_vf[((face * nqf + q) * nv) * dim], J);
          Ten2<dim> Jt;
          for (int i = 0; i < dim; ++i)
            for (int j = 0; j < dim; ++j)
              Jt.d[i][j] = J[i][j];
          const double    det  = determinant(Jt);
          const Ten2<dim> Jinv = invert(Jt);
          // n da = det(J) J^{-T} n_ref dA_ref
          double nvec[3], len = 0;
          for (int i = 0; i < dim; ++i)
            {
              nvec[i] = det * Jinv.d[d][i] * sign;
              len += nvec[i] * nvec[i];
            }
          len    = std::sqrt(len);
          JxW[q] = len * fqw[q];
          for (int i = 0; i < dim; ++i)
            normals[q * dim + i] = nvec[i] / len;
        }
    }

    // ----------------------------------------------------------------------------------------
    // Nonlinear solver
    // ----------------------------------------------------------------------------------------
    // nonlinear_elasticity.cc:872-1036
    void assemble_system_tangent_residual_one_cell(int64_t cell, const double *solution_total,
                                                   const double *acceleration, CellGeom &geom,
                                                   std::vector<Ten2<dim>> &solution_grads_u_total,
                                                   double *cell_matrix, double *cell_rhs,
                                                   const double *u_local_override   = nullptr,
                                                   const double *acc_local_override = nullptr) const
    {
      const double alpha_1 = this->alpha_1;
      double       body_force[3];
      for (int k = 0; k < 3; ++k)
        body_force[k] = desc.body_force[k];
      // data.reset(); scratch.reset(); fe_values_ref.reinit(cell) :889-892
      std::fill(cell_matrix, cell_matrix + dpc * dpc, 0.);
      std::fill(cell_rhs, cell_rhs + dpc, 0.);
      reinit_cell(cell, geom);
      const int32_t *local_dof_indices = &cell_dofs[cell * dpc];
      // get_function_gradients / get_function_values :902-906
      std::vector<Ten1<dim>> local_acceleration(nq);
      solution_grads_u_total.assign(nq, Ten2<dim>());
      for (int q = 0; q < nq; ++q)
        for (int k = 0; k < dpc; ++k)
          {
            const int    a = node_of[k], c = comp_of[k];
            const double uk =
              u_local_override ? u_local_override[k] : solution_total[local_dof_indices[k]];
            const double ak =
              acc_local_override ? acc_local_override[k] : acceleration[local_dof_indices[k]];
            for (int dd = 0; dd < dim; ++dd)
              solution_grads_u_total[q].d[c][dd] += uk * geom.grad[(q * npc + a) * dim + dd];
            local_acceleration[q].d[c] += ak * N[q * npc + a];
          }
      const double rho = material.rho; // :909

      std::vector<Ten2<dim>> grad_Nx(dpc);
      std::vector<Sym2<dim>> symm_grad_Nx(dpc);
      std::vector<Ten1<dim>> shape_value(dpc);
      for (int q_point = 0; q_point < nq; ++q_point) // :915
        {
          const Ten2<dim> &grad_u = solution_grads_u_total[q_point];
          const Ten1<dim> &acc    = local_acceleration[q_point];
          Ten2<dim>        F; // Kinematics::F :927
          for (int i = 0; i < dim; ++i)
            for (int j = 0; j < dim; ++j)
              F.d[i][j] = (i == j ? 1.0 : 0.0) + grad_u.d[i][j];
          const double det_F = determinant(F); // :929
          Ten2<dim>    F_bar; // Kinematics::F_iso :930
          {
            const double s = std::pow(det_F, -1.0 / dim);
            for (int i = 0; i < dim; ++i)
              for (int j = 0; j < dim; ++j)
                F_bar.d[i][j] = s * F.d[i][j];
          }
          const Sym2<dim> b_bar = symmetrize(matmul(F_bar, transpose(F_bar))); // :932
          const Ten2<dim> F_inv = invert(F);                                    // :934
          if (!(det_F > 0.0))
            throw std::runtime_error("det_F <= 0 (Assert nonlinear_elasticity.cc:935)");

          for (int k = 0; k < dpc; ++k) // :939-955
            {
              const int a = node_of[k], c = comp_of[k];
              Ten2<dim> grad_ref; // fe_values_ref[u_fe].gradient(k,q): only row c non-zero
              for (int dd = 0; dd < dim; ++dd)
                grad_ref.d[c][dd] = geom.grad[(q_point * npc + a) * dim + dd];
              grad_Nx[k]      = matmul(grad_ref, F_inv);
              symm_grad_Nx[k] = symmetrize(grad_Nx[k]);
              shape_value[k]  = Ten1<dim>();
              shape_value[k].d[c] = N[q_point * npc + a];
            }
          const Sym2<dim> tau = material.get_tau(det_F, b_bar); // :958
          const Sym4<dim> Jc  = material.get_Jc(det_F, b_bar);  // :960
          Ten2<dim>       tau_ns;
          for (int i = 0; i < dim; ++i)
            for (int j = 0; j < dim; ++j)
              tau_ns.d[i][j] = tau(i, j);
          const double JxW = geom.JxW[q_point];

          for (int i = 0; i < dpc; ++i) // :973
            {
              const int component_i = comp_of[i];
              // :984-988
              cell_rhs[i] -= ((symm_grad_Nx[i] * tau) -
                              (body_force[component_i] * rho * N[q_point * npc + node_of[i]])) *
                             JxW;
              // :993-995
              for (int j = 0; j < dpc; ++j)
                {
                  double dot = 0;
                  for (int dd = 0; dd < dim; ++dd)
                    dot += (shape_value[i].d[dd] * rho) * shape_value[j].d[dd];
                  cell_rhs[i] -= dot * acc.d[component_i] * JxW;
                }
              for (int j = 0; j <= i; ++j) // :1001
                {
                  const int component_j = comp_of[j];
                  // :1011-1012
                  cell_matrix[i * dpc + j] += ((symm_grad_Nx[i] * Jc) * symm_grad_Nx[j]) * JxW;
                  if (component_i == component_j) // :1015-1023
                    {
                      double gtg = 0;
                      for (int l = 0; l < dim; ++l)
                        {
                          double s = 0;
                          for (int k = 0; k < dim; ++k)
                            s += grad_Nx[i].d[component_i][k] * tau_ns.d[k][l];
                          gtg += s * grad_Nx[j].d[component_j][l];
                        }
                      cell_matrix[i * dpc + j] +=
                        (gtg + shape_value[i].d[component_i] * rho * alpha_1 *
                                 shape_value[j].d[component_j]) *
                        JxW;
                    }
                }
            }
        }
      for (int i = 0; i < dpc; ++i) // :1033-1035
        for (int j = i + 1; j < dpc; ++j)
          cell_matrix[i * dpc + j] = cell_matrix[j * dpc + i];
    }

    // nonlinear_elasticity.cc:791-859
    void assemble_neumann_contribution_one_cell(
      int64_t cell, const double *external_stress,
      const std::vector<Ten2<dim>> &solution_grads_u_total, double *cell_rhs) const
    {
      const int32_t *     local_dof_indices = &cell_dofs[cell * dpc];
      std::vector<double> JxWf, normals;
      for (int64_t fi = iface_ptr[cell]; fi < iface_ptr[cell + 1]; ++fi) // :804-805
        {
          const int face = iface_sorted_face[fi];
          reinit_face(cell, face, JxWf, normals); // :807
          std::vector<Ten1<dim>> local_stress(nqf);
          for (int q = 0; q < nqf; ++q) // :815-816
            for (int k = 0; k < dpc; ++k)
              local_stress[q].d[comp_of[k]] +=
                external_stress[local_dof_indices[k]] * Nf[(face * nqf + q) * npc + node_of[k]];
          for (int f_q_point = 0; f_q_point < nqf; ++f_q_point)
            {
              // :825-827 — cell-quadrature gradient indexed by the face q-point (as written)
              Ten2<dim> F;
              for (int i = 0; i < dim; ++i)
                for (int j = 0; j < dim; ++j)
                  F.d[i][j] = (i == j ? 1.0 : 0.0) + solution_grads_u_total[f_q_point].d[i][j];
              // :831-833 n_star = det F * F^{-T} * N
              const double    detF = determinant(F);
              const Ten2<dim> FinvT = transpose(invert(F));
              double          n_star[3], norm2 = 0;
              for (int i = 0; i < dim; ++i)
                {
                  double s = 0;
                  for (int j = 0; j < dim; ++j)
                    s += (detF * FinvT.d[i][j]) * normals[f_q_point * dim + j];
                  n_star[i] = s;
                  norm2 += s * s;
                }
              const double n_star_norm = std::sqrt(norm2);
              Ten1<dim>    referential_stress; // :836-837
              for (int i = 0; i < dim; ++i)
                referential_stress.d[i] = local_stress[f_q_point].d[i] * n_star_norm;
              for (int i = 0; i < dpc; ++i) // :839-856
                {
                  const int    component_i = comp_of[i];
                  const double Ni          = Nf[(face * nqf + f_q_point) * npc + node_of[i]];
                  const double JxW         = JxWf[f_q_point];
                  cell_rhs[i] += (Ni * referential_stress.d[component_i]) * JxW;
                }
            }
        }
    }

    // [deal.II] AffineConstraints::distribute_local_to_global, homogeneous Dirichlet only
    // (call site nonlinear_elasticity.cc:769-773)
    void distribute_local_to_global(const double *cell_matrix, const double *cell_rhs,
                                    const int32_t *ids, std::vector<double> &K,
                                    std::vector<double> &rhs) const
    {
      bool any_constrained = false;
      for (int i = 0; i < dpc; ++i)
        any_constrained |= (constrained[ids[i]] != 0);
      double average_diagonal = 0;
      if (any_constrained)
        {
          for (int i = 0; i < dpc; ++i)
            average_diagonal += std::fabs(cell_matrix[i * dpc + i]);
          average_diagonal /= double(dpc);
        }
      for (int i = 0; i < dpc; ++i)
        {
          const int32_t gi = ids[i];
          if (constrained[gi])
            {
              const double d = std::fabs(cell_matrix[i * dpc + i]);
              K[pattern.diag[gi]] += (d != 0 ? d : average_diagonal);
              continue;
            }
          rhs[gi] += cell_rhs[i];
          for (int j = 0; j < dpc; ++j)
            {
              const int32_t gj = ids[j];
              if (constrained[gj])
                continue;
              K[pattern.find(gi, gj)] += cell_matrix[i * dpc + j];
            }
        }
    }

    // nonlinear_elasticity.cc:1044-1087 (WorkStream: parallel worker, ordered serial copier)
    void nl_assemble_system(int n_threads)
    {
      std::vector<double> &K   = mats[ORC_MAT_TANGENT];
      std::vector<double> &rhs = V(ORC_NL_SYSTEM_RHS);
      std::fill(K.begin(), K.end(), 0.); // :1054-1055
      std::fill(rhs.begin(), rhs.end(), 0.);
      // get_total_solution :580-588
      std::vector<double> solution_total(V(ORC_NL_TOTAL_DISPLACEMENT));
      for (int64_t i = 0; i < desc.n_dofs; ++i)
        solution_total[i] += V(ORC_NL_SOLUTION_DELTA)[i];
      const std::vector<double> &acceleration    = V(ORC_NL_ACCELERATION);
      const std::vector<double> &external_stress = V(ORC_NL_EXTERNAL_STRESS);
      n_threads                                  = std::max(1, n_threads);
      const int64_t       chunk = int64_t(n_threads) * 8;
      std::vector<double> cm(chunk * dpc * dpc), cr(chunk * dpc);
      std::string         err;
      for (int64_t c0 = 0; c0 < desc.n_cells; c0 += chunk)
        {
          const int64_t c1     = std::min<int64_t>(desc.n_cells, c0 + chunk);
          auto          worker = [&](int t) {
            try
              {
                CellGeom               geom;
                std::vector<Ten2<dim>> grads;
                for (int64_t c = c0 + t; c < c1; c += n_threads)
                  {
                    double *m = &cm[(c - c0) * dpc * dpc], *r = &cr[(c - c0) * dpc];
                    assemble_system_tangent_residual_one_cell(
                      c, solution_total.data(), acceleration.data(), geom, grads, m, r); // :755
                    assemble_neumann_contribution_one_cell(c, external_stress.data(), grads,
                                                           r); // :756
                  }
              }
            catch (std::exception &e)
              {
                err = e.what();
              }
          };
          if (n_threads == 1)
            worker(0);
          else
            {
              std::vector<std::thread> th;
              for (int t = 0; t < n_threads; ++t)
                th.emplace_back(worker, t);
              for (auto &t : th)
                t.join();
            }
          if (!err.empty())
            throw std::runtime_error(err);
          for (int64_t c = c0; c < c1; ++c) // copier, in cell order :760-774
            distribute_local_to_global(&cm[(c - c0) * dpc * dpc], &cr[(c - c0) * dpc],
                                       &cell_dofs[c * dpc], K, rhs);
        }
    }

    void nl_update_acceleration() // :592-599
    {
      auto &a = V(ORC_NL_ACCELERATION);
      for (int64_t i = 0; i < desc.n_dofs; ++i)
        {
          a[i] = alpha_1 * V(ORC_NL_SOLUTION_DELTA)[i];
          a[i] += -alpha_2 * V(ORC_NL_VELOCITY_OLD)[i] + -alpha_3 * V(ORC_NL_ACCELERATION_OLD)[i];
        }
    }
    void nl_update_velocity() // :602-610
    {
      auto &v = V(ORC_NL_VELOCITY);
      for (int64_t i = 0; i < desc.n_dofs; ++i)
        {
          v[i] = alpha_4 * V(ORC_NL_SOLUTION_DELTA)[i];
          v[i] += alpha_5 * V(ORC_NL_VELOCITY_OLD)[i] + alpha_6 * V(ORC_NL_ACCELERATION_OLD)[i];
        }
    }
    void nl_update_old_variables() // :613-622
    {
      V(ORC_NL_TOTAL_DISPLACEMENT_OLD) = V(ORC_NL_TOTAL_DISPLACEMENT);
      V(ORC_NL_VELOCITY_OLD)           = V(ORC_NL_VELOCITY);
      V(ORC_NL_ACCELERATION_OLD)       = V(ORC_NL_ACCELERATION);
    }
    double masked_l2(const std::vector<double> &v) const // :549-576
    {
      double s = 0;
      for (int64_t i = 0; i < desc.n_dofs; ++i)
        if (!constrained[i])
          s += v[i] * v[i];
      return std::sqrt(s);
    }
    // :1153-1211
    int nl_solve_linear_system(uint32_t &lin_it, double &lin_res)
    {
      auto &      newton_update = V(ORC_NL_NEWTON_UPDATE);
      const auto &rhs           = V(ORC_NL_SYSTEM_RHS);
      int         status        = 0;
      const int   solver_its    = int(double(desc.n_dofs) * desc.max_iterations_lin); // :1169
      if (desc.type_lin == 0)
        {
          const double tol_sol = desc.tol_lin * l2_norm(rhs); // :1171
          status = solver_cg_ssor(pattern, mats[ORC_MAT_TANGENT], newton_update, rhs, .65,
                                  solver_its, tol_sol, lin_it, lin_res); // :1174-1190
        }
      else
        {
          // Direct (UMFPACK :1194-1199) stand-in: CG+SSOR to 1e-13 relative
          uint32_t it;
          double   res;
          std::fill(newton_update.begin(), newton_update.end(), 0.);
          status  = solver_cg_ssor(pattern, mats[ORC_MAT_TANGENT], newton_update, rhs, 1.0,
                                  10 * solver_its, 1e-13 * l2_norm(rhs), it, res);
          lin_it  = 1;
          lin_res = 0.0;
        }
      for (int64_t i = 0; i < desc.n_dofs; ++i) // constraints.distribute :1208
        if (constrained[i])
          newton_update[i] = 0.;
      return status;
    }
    // :410-499
    int nl_solve_nonlinear_timestep(int n_threads, double *hist, int hist_rows)
    {
      auto &newton_update = V(ORC_NL_NEWTON_UPDATE);
      std::fill(newton_update.begin(), newton_update.end(), 0.); // :419
      double error_residual = 1.0, error_residual_0 = 1.0, error_residual_norm = 1.0,
             error_update = 1.0, error_update_0 = 1.0, error_update_norm = 1.0; // :421-426
      unsigned int newton_iteration = 0;
      for (; newton_iteration < (unsigned)desc.max_iterations_NR; ++newton_iteration)
        {
          nl_update_acceleration();       // :444
          nl_assemble_system(n_threads);  // :446
          error_residual = masked_l2(V(ORC_NL_SYSTEM_RHS)); // :449
          if (newton_iteration == 0)
            error_residual_0 = error_residual;
          error_residual_norm = error_residual;
          if (error_residual_0 != 0.0) // Errors::normalise nonlinear_elasticity.h:304-309
            error_residual_norm /= error_residual_0;
          if (newton_iteration > 0 &&
              ((error_update_norm <= desc.tol_u || error_update <= 1e-15) &&
               (error_residual_norm <= desc.tol_f || error_residual <= 5e-9))) // :459-463
            break;
          uint32_t  lin_it  = 0;
          double    lin_res = 0;
          const int st      = nl_solve_linear_system(lin_it, lin_res); // :473
          if (st != 0)
            return -2;
          error_update = masked_l2(newton_update); // :476
          if (newton_iteration == 0)
            error_update_0 = error_update;
          error_update_norm = error_update;
          if (error_update_0 != 0.0)
            error_update_norm /= error_update_0;
          for (int64_t i = 0; i < desc.n_dofs; ++i) // :487
            V(ORC_NL_SOLUTION_DELTA)[i] += newton_update[i];
          if (hist && (int)newton_iteration < hist_rows)
            {
              double *row = hist + 6 * newton_iteration;
              row[0]      = lin_it;
              row[1]      = lin_res;
              row[2]      = error_residual_norm;
              row[3]      = error_residual;
              row[4]      = error_update_norm;
              row[5]      = error_update;
            }
        }
      if (!(newton_iteration < (unsigned)desc.max_iterations_NR)) // :497
        return -1;
      return int(newton_iteration);
    }
    // nonlinear_elasticity.cc:121,138-144
    int nl_timestep(int n_threads, double *hist, int hist_rows)
    {
      auto &delta = V(ORC_NL_SOLUTION_DELTA);
      std::fill(delta.begin(), delta.end(), 0.); // :121
      const int n = nl_solve_nonlinear_timestep(n_threads, hist, hist_rows); // :138
      if (n < 0)
        return n;
      for (int64_t i = 0; i < desc.n_dofs; ++i) // :139
        V(ORC_NL_TOTAL_DISPLACEMENT)[i] += delta[i];
      nl_update_acceleration();  // :142
      nl_update_velocity();      // :143
      nl_update_old_variables(); // :144
      return n;
    }

    // ----------------------------------------------------------------------------------------
    // Linear solver
    // ----------------------------------------------------------------------------------------
    // linear_elasticity.cc:248-374
    void lin_assemble_system()
    {
      const double lambda = 2 * desc.mu * desc.nu / (1 - 2 * desc.nu); // parameters.cc:189
      const double mu     = desc.mu;
      auto &       Kmat   = mats[ORC_MAT_STIFFNESS];
      auto &       Mmat   = mats[ORC_MAT_MASS];
      auto &       Smat   = mats[ORC_MAT_STEPPING];
      std::fill(Kmat.begin(), Kmat.end(), 0.);
      std::fill(Mmat.begin(), Mmat.end(), 0.);
      std::vector<double> cell_matrix(dpc * dpc), cell_mass(dpc * dpc);
      CellGeom            geom;
      auto &              bf = V(ORC_LIN_BODY_FORCE);
      std::fill(bf.begin(), bf.end(), 0.);
      double bnorm = 0;
      for (int k = 0; k < 3; ++k)
        bnorm += desc.body_force[k] * desc.body_force[k];
      const bool body_force_enabled = std::sqrt(bnorm) > 1e-15; // :62
      for (int64_t cell = 0; cell < desc.n_cells; ++cell)      // :276
        {
          std::fill(cell_matrix.begin(), cell_matrix.end(), 0.);
          std::fill(cell_mass.begin(), cell_mass.end(), 0.);
          reinit_cell(cell, geom);
          for (int i = 0; i < dpc; ++i) // :289-323
            {
              const int component_i = comp_of[i], ai = node_of[i];
              for (int j = 0; j < dpc; ++j)
                {
                  const int component_j = comp_of[j], aj = node_of[j];
                  for (int q = 0; q < nq; ++q)
                    {
                      const double *gi = &geom.grad[(q * npc + ai) * dim];
                      const double *gj = &geom.grad[(q * npc + aj) * dim];
                      double        gg = 0;
                      if (component_i == component_j)
                        {
                          for (int k = 0; k < dim; ++k)
                            gg += gi[k] * gj[k];
                          gg = gg * mu;
                        }
                      cell_matrix[i * dpc + j] +=
                        ((gi[component_i] * gj[component_j] * lambda) +
                         (gi[component_j] * gj[component_i] * mu) + gg) *
                        geom.JxW[q];
                      // MatrixCreator::create_mass_matrix with coefficient rho :341-344
                      if (component_i == component_j)
                        cell_mass[i * dpc + j] +=
                          desc.rho * N[q * npc + ai] * N[q * npc + aj] * geom.JxW[q];
                    }
                }
            }
          const int32_t *ids = &cell_dofs[cell * dpc];
          for (int i = 0; i < dpc; ++i) // :327-334
            for (int j = 0; j < dpc; ++j)
              {
                const int64_t pos = pattern.find(ids[i], ids[j]);
                Kmat[pos] += cell_matrix[i * dpc + j];
                Mmat[pos] += cell_mass[i * dpc + j];
              }
          if (body_force_enabled) // :358-373 VectorTools::create_right_hand_side
            for (int i = 0; i < dpc; ++i)
              {
                double s = 0;
                for (int q = 0; q < nq; ++q)
                  s += (desc.rho * desc.body_force[comp_of[i]]) * N[q * npc + node_of[i]] * geom.JxW[q];
                bf[ids[i]] += s;
              }
        }
      // :348-353
      const double f = desc.delta_t * desc.delta_t * desc.theta * desc.theta;
      for (size_t k = 0; k < Smat.size(); ++k)
        {
          Smat[k] = Kmat[k];
          Smat[k] *= f;
          Smat[k] += 1 * Mmat[k];
        }
    }
    // linear_elasticity.cc:458-521
    void lin_assemble_consistent_loading()
    {
      auto &      system_rhs = V(ORC_LIN_SYSTEM_RHS);
      const auto &stress     = V(ORC_LIN_STRESS);
      std::fill(system_rhs.begin(), system_rhs.end(), 0.); // :462
      std::vector<double> cell_rhs(dpc), JxWf, normals;
      for (int64_t cell = 0; cell < desc.n_cells; ++cell) // :483
        {
          if (iface_ptr[cell] == iface_ptr[cell + 1])
            continue; // cell_rhs == 0: adds exact zeros
          std::fill(cell_rhs.begin(), cell_rhs.end(), 0.);
          const int32_t *ids = &cell_dofs[cell * dpc];
          for (int64_t fi = iface_ptr[cell]; fi < iface_ptr[cell + 1]; ++fi)
            {
              const int face = iface_sorted_face[fi];
              reinit_face(cell, face, JxWf, normals);
              std::vector<double> local_stress(nqf * dim, 0.); // :499
              for (int q = 0; q < nqf; ++q)
                for (int k = 0; k < dpc; ++k)
                  local_stress[q * dim + comp_of[k]] +=
                    stress[ids[k]] * Nf[(face * nqf + q) * npc + node_of[k]];
              for (int q = 0; q < nqf; ++q) // :501-511
                for (int i = 0; i < dpc; ++i)
                  cell_rhs[i] += Nf[(face * nqf + q) * npc + node_of[i]] *
                                 local_stress[q * dim + comp_of[i]] * JxWf[q];
            }
          for (int i = 0; i < dpc; ++i) // :516-519
            system_rhs[ids[i]] += cell_rhs[i];
        }
    }
    // linear_elasticity.cc:378-454
    void lin_assemble_rhs()
    {
      const int64_t n          = desc.n_dofs;
      auto &        system_rhs = V(ORC_LIN_SYSTEM_RHS);
      const double  dt = desc.delta_t, theta = desc.theta;
      if (desc.data_consistent) // :385-388
        lin_assemble_consistent_loading();
      else
        system_rhs = V(ORC_LIN_STRESS);
      V(ORC_LIN_OLD_VELOCITY)     = V(ORC_LIN_VELOCITY); // :390-391
      V(ORC_LIN_OLD_DISPLACEMENT) = V(ORC_LIN_DISPLACEMENT);
      double bnorm = 0;
      for (int k = 0; k < 3; ++k)
        bnorm += desc.body_force[k] * desc.body_force[k];
      if (std::sqrt(bnorm) > 1e-15) // :394-395
        for (int64_t i = 0; i < n; ++i)
          system_rhs[i] += 1 * V(ORC_LIN_BODY_FORCE)[i];
      std::vector<double> tmp(system_rhs); // :402-405
      for (int64_t i = 0; i < n; ++i)      // :407-408
        {
          system_rhs[i] *= dt * theta;
          system_rhs[i] += dt * (1 - theta) * V(ORC_LIN_OLD_STRESS)[i];
        }
      V(ORC_LIN_OLD_STRESS) = tmp; // :409
      vmult(pattern, mats[ORC_MAT_MASS], V(ORC_LIN_OLD_VELOCITY).data(), tmp.data()); // :411
      for (int64_t i = 0; i < n; ++i)
        system_rhs[i] += 1 * tmp[i];
      vmult(pattern, mats[ORC_MAT_STIFFNESS], V(ORC_LIN_OLD_VELOCITY).data(), tmp.data()); // :414
      for (int64_t i = 0; i < n; ++i)
        system_rhs[i] += (-theta * dt * dt * (1 - theta)) * tmp[i];
      vmult(pattern, mats[ORC_MAT_STIFFNESS], V(ORC_LIN_OLD_DISPLACEMENT).data(),
            tmp.data()); // :419
      for (int64_t i = 0; i < n; ++i)
        system_rhs[i] += (-dt) * tmp[i];
      auto &A = mats[ORC_MAT_SYSTEM]; // :426-427
      A       = mats[ORC_MAT_STEPPING];
      // [deal.II] MatrixTools::apply_boundary_values with zero boundary values :431-451
      double first_nonzero_diagonal_entry = 1;
      for (int64_t i = 0; i < n; ++i)
        if (A[pattern.diag[i]] != 0.)
          {
            first_nonzero_diagonal_entry = A[pattern.diag[i]];
            break;
          }
      auto &velocity = V(ORC_LIN_VELOCITY);
      for (int64_t dof = 0; dof < n; ++dof)
        if (constrained[dof])
          {
            for (int64_t j = pattern.rowptr[dof]; j < pattern.rowptr[dof + 1]; ++j)
              if (j != pattern.diag[dof])
                A[j] = 0.;
            double new_rhs;
            if (A[pattern.diag[dof]] != 0.)
              new_rhs = 0. * A[pattern.diag[dof]];
            else
              {
                A[pattern.diag[dof]] = first_nonzero_diagonal_entry;
                new_rhs              = 0. * first_nonzero_diagonal_entry;
              }
            system_rhs[dof]             = new_rhs;
            const double diagonal_entry = A[pattern.diag[dof]];
            for (int64_t j = pattern.rowptr[dof]; j < pattern.rowptr[dof + 1]; ++j)
              if (j != pattern.diag[dof])
                {
                  const int64_t row = pattern.col[j];
                  const int64_t pos = pattern.find(row, int32_t(dof));
                  system_rhs[row] -= A[pos] / diagonal_entry * new_rhs;
                  A[pos] = 0.;
                }
            velocity[dof] = 0.;
          }
    }
    // linear_elasticity.cc:525-575
    int lin_solve(uint32_t &lin_it, double &lin_res)
    {
      lin_it           = 1;
      lin_res          = 0.0;
      int       status = 0;
      const int solver_its = int(double(desc.n_dofs) * desc.max_iterations_lin); // :540
      if (desc.type_lin == 0)
        status = solver_cg_ssor(pattern, mats[ORC_MAT_SYSTEM], V(ORC_LIN_VELOCITY),
                                V(ORC_LIN_SYSTEM_RHS), 1.2, solver_its, 1.e-10, lin_it,
                                lin_res); // :542-554
      else
        {
          uint32_t it;
          double   res; // Direct (UMFPACK :560-562) stand-in
          auto &   v = V(ORC_LIN_VELOCITY);
          std::fill(v.begin(), v.end(), 0.);
          status = solver_cg_ssor(pattern, mats[ORC_MAT_SYSTEM], v, V(ORC_LIN_SYSTEM_RHS), 1.0,
                                  10 * solver_its, 1e-13 * l2_norm(V(ORC_LIN_SYSTEM_RHS)), it, res);
        }
      return status;
    }
    void lin_update_displacement() // :579-586
    {
      const double dt = desc.delta_t, theta = desc.theta;
      auto &       d = V(ORC_LIN_DISPLACEMENT);
      for (int64_t i = 0; i < desc.n_dofs; ++i)
        {
          d[i] += dt * theta * V(ORC_LIN_VELOCITY)[i];
          d[i] += dt * (1 - theta) * V(ORC_LIN_OLD_VELOCITY)[i];
        }
    }

    // ----------------------------------------------------------------------------------------
    // Output path: output_results (nonlinear_elasticity.cc:1215-1254, linear_elasticity.cc:
    // 590-629) = DataOut::build_patches(MappingQEulerian(degree, dof_handler, displacement),
    // degree, curved_boundary) + Postprocessor::evaluate_vector_field (postprocessor.h:44-76).
    // [deal.II] build_patches evaluates solution values and gradients at the (degree+1)^dim
    // equidistant patch points (lexicographic, x fastest) with FEValues on the GIVEN mapping;
    // MappingQEulerian maps xi -> X(xi) + u_h(xi), so gradients are taken with respect to the
    // displaced coordinates: J_e = J + grad_xi u, grad_x u = grad_xi u . J_e^-1.
    //   fields[(cell*npts + pt)*(dim+dim*dim) + ...] = u, then strain unrolled (d*dim + e)
    //   points[(cell*npts + pt)*dim + ...]           = X + u (patch vertices on the displaced grid)
    // ----------------------------------------------------------------------------------------
    void postprocess(int which, double *points, double *fields) const
    {
      const std::vector<double> &sol = vecs[which];
      const int                  nv = 1 << dim, nf = dim + dim * dim;
      std::vector<double>        u_local(dpc);
      for (int64_t cell = 0; cell < desc.n_cells; ++cell)
        {
          const double * verts = &cell_vertices[cell * nv * dim];
          const int32_t *ldi   = &cell_dofs[cell * dpc];
          for (int k = 0; k < dpc; ++k)
            u_local[k] = sol[ldi[k]];
          for (int pt = 0; pt < npc; ++pt)
            {
              double xi[3] = {0, 0, 0};
              int    rem   = pt;
              for (int d = 0; d < dim; ++d)
                {
                  xi[d] = double(rem % (p + 1)) / p;
                  rem /= (p + 1);
                }
              // reference position and Jacobian of the Q1 geometry
              double X[3] = {0, 0, 0}, J[3][3];
              std::vector<double> dphi(nv * dim);
              for (int v = 0; v < nv; ++v)
                {
                  vertex_grad(v, xi, &dphi[v * dim]);
                  double phi = 1;
                  for (int l = 0; l < dim; ++l)
                    phi *= ((v >> l) & 1) ? xi[l] : 1.0 - xi[l];
                  for (int i = 0; i < dim; ++i)
                    X[i] += phi * verts[v * dim + i];
                }
              jacobian(verts, dphi.data(), J);
              // solution value and unit-cell gradient
              double    val[3] = {0, 0, 0};
              Ten2<dim> Gxi, Je;
              for (int a = 0; a < npc; ++a)
                {
                  const double Na = shape(a, xi);
                  double       g[3];
                  shape_grad(a, xi, g);
                  for (int c = 0; c < dim; ++c)
                    {
                      val[c] += Na * u_local[loc_of[a * dim + c]];
                      for (int k = 0; k < dim; ++k)
                        Gxi.d[c][k] += g[k] * u_local[loc_of[a * dim + c]];
                    }
                }
              for (int i = 0; i < dim; ++i)
                for (int j = 0; j < dim; ++j)
                  Je.d[i][j] = J[i][j] + Gxi.d[i][j];
              if (!(determinant(Je) > 0))
                throw std::runtime_error("oracle: inverted MappingQEulerian cell in output");
              const Ten2<dim> grad_u = matmul(Gxi, invert(Je));
              double *        out    = fields + (cell * npc + pt) * nf;
              for (int d = 0; d < dim; ++d) // postprocessor.h:61-72
                {
                  out[d] = val[d];
                  for (int e = 0; e < dim; ++e)
                    out[dim + d * dim + e] = (grad_u.d[d][e] + grad_u.d[e][d]) / 2;
                }
              if (points)
                for (int d = 0; d < dim; ++d)
                  points[(cell * npc + pt) * dim + d] = X[d] + val[d];
            }
        }
    }
  };

  template <typename F>
  auto dispatch(void *h, F &&f)
  {
    ContextBase *b = static_cast<ContextBase *>(h);
    if (b->rt_dim == 2)
      return f(*static_cast<Context<2> *>(b));
    return f(*static_cast<Context<3> *>(b));
  }
  const int nl_state[6]  = {ORC_NL_TOTAL_DISPLACEMENT, ORC_NL_TOTAL_DISPLACEMENT_OLD,
                           ORC_NL_VELOCITY,           ORC_NL_VELOCITY_OLD,
                           ORC_NL_ACCELERATION,       ORC_NL_ACCELERATION_OLD};
  const int lin_state[5] = {ORC_LIN_OLD_VELOCITY, ORC_LIN_VELOCITY, ORC_LIN_OLD_DISPLACEMENT,
                            ORC_LIN_DISPLACEMENT, ORC_LIN_OLD_STRESS};
} // namespace

extern "C"
{
  const char *orc_last_error(void) { return g_last_error.c_str(); }
  void *      orc_create(const orc_desc *d)
  {
    try
      {
        if (d->dim == 2)
          return static_cast<ContextBase *>(new Context<2>(*d));
        if (d->dim == 3)
          return static_cast<ContextBase *>(new Context<3>(*d));
        throw std::invalid_argument("oracle: dim must be 2 or 3");
      }
    catch (std::exception &e)
      {
        g_last_error = e.what();
        return nullptr;
      }
  }
  void    orc_destroy(void *h) { delete static_cast<ContextBase *>(h); }
  int64_t orc_nnz(void *h) { return int64_t(static_cast<ContextBase *>(h)->pattern.col.size()); }
  void    orc_get_pattern(void *h, int64_t *rowptr, int32_t *col)
  {
    ContextBase *b = static_cast<ContextBase *>(h);
    std::copy(b->pattern.rowptr.begin(), b->pattern.rowptr.end(), rowptr);
    std::copy(b->pattern.col.begin(), b->pattern.col.end(), col);
  }
  void orc_get_values(void *h, int which, double *val)
  {
    ContextBase *b = static_cast<ContextBase *>(h);
    std::copy(b->mats[which].begin(), b->mats[which].end(), val);
  }
  void orc_get_vector(void *h, int which, double *out)
  {
    ContextBase *b = static_cast<ContextBase *>(h);
    std::copy(b->vecs[which].begin(), b->vecs[which].end(), out);
  }
  void orc_set_vector(void *h, int which, const double *in)
  {
    ContextBase *b = static_cast<ContextBase *>(h);
    std::copy(in, in + b->desc.n_dofs, b->vecs[which].begin());
  }
  void orc_nl_update_acceleration(void *h)
  {
    dispatch(h, [](auto &c) { c.nl_update_acceleration(); return 0; });
  }
  void orc_nl_update_velocity(void *h)
  {
    dispatch(h, [](auto &c) { c.nl_update_velocity(); return 0; });
  }
  void orc_nl_update_old_variables(void *h)
  {
    dispatch(h, [](auto &c) { c.nl_update_old_variables(); return 0; });
  }
  void orc_nl_assemble_system(void *h, int n_threads)
  {
    try
      {
        dispatch(h, [&](auto &c) { c.nl_assemble_system(n_threads); return 0; });
      }
    catch (std::exception &e)
      {
        g_last_error = e.what();
      }
  }
  double orc_nl_error_residual(void *h)
  {
    return dispatch(h, [](auto &c) { return c.masked_l2(c.V(ORC_NL_SYSTEM_RHS)); });
  }
  int orc_nl_solve_linear_system(void *h, uint32_t *lin_it, double *lin_res)
  {
    return dispatch(h, [&](auto &c) { return c.nl_solve_linear_system(*lin_it, *lin_res); });
  }
  int orc_nl_solve_nonlinear_timestep(void *h, int n_threads, double *hist, int hist_rows)
  {
    try
      {
        return dispatch(
          h, [&](auto &c) { return c.nl_solve_nonlinear_timestep(n_threads, hist, hist_rows); });
      }
    catch (std::exception &e)
      {
        g_last_error = e.what();
        return -3;
      }
  }
  int orc_nl_timestep(void *h, int n_threads, double *hist, int hist_rows)
  {
    try
      {
        return dispatch(h, [&](auto &c) { return c.nl_timestep(n_threads, hist, hist_rows); });
      }
    catch (std::exception &e)
      {
        g_last_error = e.what();
        return -3;
      }
  }
  void orc_lin_assemble_system(void *h)
  {
    dispatch(h, [](auto &c) { c.lin_assemble_system(); return 0; });
  }
  void orc_lin_assemble_rhs(void *h)
  {
    dispatch(h, [](auto &c) { c.lin_assemble_rhs(); return 0; });
  }
  int orc_lin_solve(void *h, uint32_t *lin_it, double *lin_res)
  {
    return dispatch(h, [&](auto &c) { return c.lin_solve(*lin_it, *lin_res); });
  }
  void orc_lin_update_displacement(void *h)
  {
    dispatch(h, [](auto &c) { c.lin_update_displacement(); return 0; });
  }
  // adapter.h:421-443
  void orc_format_precice_to_deal(void *h, const double *read_data_buffer, int which)
  {
    ContextBase *       b   = static_cast<ContextBase *>(h);
    const int           dim = b->rt_dim;
    const int64_t       n   = b->desc.n_iface_nodes;
    std::vector<double> &v  = b->vecs[which];
    for (int64_t i = 0; i < n; ++i)
      for (int c = 0; c < dim; ++c)
        v[b->iface_dofs[c * n + i]] = read_data_buffer[dim * i + c];
  }
  // adapter.h:389-417
  void orc_format_deal_to_precice(void *h, int which, double *write_data_buffer)
  {
    ContextBase *             b   = static_cast<ContextBase *>(h);
    const int                 dim = b->rt_dim;
    const int64_t             n   = b->desc.n_iface_nodes;
    const std::vector<double> &v  = b->vecs[which];
    for (int64_t i = 0; i < n; ++i)
      for (int c = 0; c < dim; ++c)
        write_data_buffer[dim * i + c] = v[b->iface_dofs[c * n + i]];
  }
  // adapter.h:457-462
  void orc_save_state(void *h)
  {
    ContextBase *b     = static_cast<ContextBase *>(h);
    const bool   nl    = b->desc.model == 1;
    const int    count = nl ? 6 : 5;
    b->old_state_data.resize(count);
    for (int i = 0; i < count; ++i)
      b->old_state_data[i] = b->vecs[nl ? nl_state[i] : lin_state[i]];
  }
  // adapter.h:482-487
  void orc_reload_state(void *h)
  {
    ContextBase *b     = static_cast<ContextBase *>(h);
    const bool   nl    = b->desc.model == 1;
    const int    count = nl ? 6 : 5;
    for (int i = 0; i < count && i < (int)b->old_state_data.size(); ++i)
      b->vecs[nl ? nl_state[i] : lin_state[i]] = b->old_state_data[i];
  }
  void orc_material(int dim, double mu, double nu, double det_F, const double *b_bar, double *psi,
                    double *tau, double *Jc)
  {
    if (dim == 2)
      {
        Material<2> m(mu, nu, 0.);
        Sym2<2>     b;
        for (int i = 0; i < 3; ++i)
          b.d[i] = b_bar[i];
        *psi            = m.get_Psi(det_F, b);
        const Sym2<2> t = m.get_tau(det_F, b);
        const Sym4<2> J = m.get_Jc(det_F, b);
        for (int i = 0; i < 3; ++i)
          {
            tau[i] = t.d[i];
            for (int j = 0; j < 3; ++j)
              Jc[i * 3 + j] = J.d[i][j];
          }
      }
    else
      {
        Material<3> m(mu, nu, 0.);
        Sym2<3>     b;
        for (int i = 0; i < 6; ++i)
          b.d[i] = b_bar[i];
        *psi            = m.get_Psi(det_F, b);
        const Sym2<3> t = m.get_tau(det_F, b);
        const Sym4<3> J = m.get_Jc(det_F, b);
        for (int i = 0; i < 6; ++i)
          {
            tau[i] = t.d[i];
            for (int j = 0; j < 6; ++j)
              Jc[i * 6 + j] = J.d[i][j];
          }
      }
  }
  void orc_nl_cell(void *h, int64_t cell, const double *u_local, const double *acc_local,
                   double *cell_matrix, double *cell_rhs)
  {
    dispatch(h, [&](auto &c) {
      using C = typename std::remove_reference<decltype(c)>::type;
      typename C::CellGeom      geom;
      std::vector<Ten2<C::DIM>> grads;
      c.assemble_system_tangent_residual_one_cell(cell, nullptr, nullptr, geom, grads,
                                                  cell_matrix, cell_rhs, u_local, acc_local);
      return 0;
    });
  }
  void orc_postprocess(void *h, int which, double *points, double *fields)
  {
    dispatch(h, [&](auto &c) { c.postprocess(which, points, fields); return 0; });
  }
  void orc_vmult(void *h, int which, const double *x, double *y)
  {
    ContextBase *b = static_cast<ContextBase *>(h);
    vmult(b->pattern, b->mats[which], x, y);
  }
  int orc_threads_available(void) { return int(std::thread::hardware_concurrency()); }
}
