// FE_Q(p) on the unit interval / cell for every polynomial degree (host code, header only; shared
// by the device library's table builders and the host mesh stand-in).
//
// The reference takes the degree from the parameter file ("Polynomial degree",
// parameters.cc:110-113; shipped values: parameters.prm:21 = 3, nonlinear_elasticity.prm:24 = 4)
// and hands it to FESystem(FE_Q(degree), dim) (nonlinear_elasticity.cc:69-72,
// linear_elasticity.cc:55-60). What deal.II 9.5 does with it is restated here:
//   * support points: FE_Q(p) = Lagrange polynomials on the (p+1) Gauss-Lobatto points of [0,1]
//     (equidistant for p <= 2, where both families coincide);
//   * local scalar numbering ("hierarchical"): vertices, then lines, quads, hex, the interior
//     DoFs of each entity in the entity's lexicographic order
//     (FETools::hierarchic_to_lexicographic_numbering: lines 0..3 = x-, x+, y-, y+ of the z = 0
//     face, 4..7 the same at z = 1, 8..11 along z; quads x-, x+ (y fastest), y-, y+ (z fastest,
//     then x), z-, z+ (x fastest));
//   * FESystem local numbering: entity by entity, inside an entity component by component, inside
//     a component the entity's scalar DoFs (FESystem::build_cell_tables) - i.e. node-major /
//     component-minor only while every entity carries ONE scalar DoF (p <= 2).
#pragma once
#include <cmath>
#include <vector>

namespace gf_fe
{
  // the n Gauss-Lobatto points on [0,1], ascending (n >= 2): 0, 1 and the roots of P'_{n-1}
  inline std::vector<double> gauss_lobatto01(int n)
  {
    std::vector<double> x(n, 0.0);
    x[n - 1] = 1.0;
    if (n == 3)
      x[1] = 0.5;
    if (n <= 3)
      return x;
    const int         m  = n - 1; // roots of P'_m on (-1, 1): m - 1 of them
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 1; i < (n + 1) / 2; ++i)
      {
        // Chebyshev-Gauss-Lobatto start value, Newton on q(z) = P'_m(z) with
        // (1 - z^2) P''_m = 2 z P'_m - m (m + 1) P_m
        long double z = -cosl(pi * i / m);
        for (int it = 0; it < 100; ++it)
          {
            long double p0 = 1, p1 = z;
            for (int k = 2; k <= m; ++k)
              {
                const long double pk = ((2 * k - 1) * z * p1 - (k - 1) * p0) / k;
                p0                   = p1;
                p1                   = pk;
              }
            const long double dp  = m * (z * p1 - p0) / (z * z - 1);           // P'_m
            const long double ddp = (2 * z * dp - m * (m + 1) * p1) / (1 - z * z); // P''_m
            const long double dz  = dp / ddp;
            z -= dz;
            if (fabsl(dz) < 1e-19L)
              break;
          }
        x[i]         = double(0.5L + 0.5L * z);
        x[n - 1 - i] = double(0.5L - 0.5L * z);
      }
    if (n % 2 == 1)
      x[n / 2] = 0.5;
    return x;
  }

  // 1-D Lagrange basis of degree p on the FE_Q support points; closed forms for p <= 2 (kept
  // exactly as the tables of rounds 1/2 evaluated them), product form beyond
  struct Basis1D
  {
    int                 p;
    std::vector<double> nodes; // p + 1 support points, ascending
    explicit Basis1D(int p_)
      : p(p_)
      , nodes(gauss_lobatto01(p_ + 1))
    {}
    double value(int i, double x) const
    {
      if (p == 1)
        return i == 0 ? 1.0 - x : x;
      if (p == 2)
        return i == 0 ? 2.0 * (x - 0.5) * (x - 1.0) :
                        (i == 1 ? -4.0 * x * (x - 1.0) : 2.0 * x * (x - 0.5));
      double v = 1.0;
      for (int j = 0; j <= p; ++j)
        if (j != i)
          v *= (x - nodes[j]) / (nodes[i] - nodes[j]);
      return v;
    }
    double derivative(int i, double x) const
    {
      if (p == 1)
        return i == 0 ? -1.0 : 1.0;
      if (p == 2)
        return i == 0 ? 4.0 * x - 3.0 : (i == 1 ? -8.0 * x + 4.0 : 4.0 * x - 1.0);
      double s = 0.0;
      for (int k = 0; k <= p; ++k)
        if (k != i)
          {
            double v = 1.0 / (nodes[i] - nodes[k]);
            for (int j = 0; j <= p; ++j)
              if (j != i && j != k)
                v *= (x - nodes[j]) / (nodes[i] - nodes[j]);
            s += v;
          }
      return s;
    }
  };

  // hierarchical local scalar node a -> lexicographic (lx, ly, lz) in 0..p; lex[3 a + d]
  inline void local_nodes(int dim, int p, std::vector<int> &lex)
  {
    lex.clear();
    auto push = [&](int x, int y, int z) {
      lex.push_back(x);
      lex.push_back(y);
      lex.push_back(z);
    };
    for (int v = 0; v < (1 << dim); ++v)
      push((v & 1) * p, ((v >> 1) & 1) * p, dim == 3 ? ((v >> 2) & 1) * p : 0);
    const int zs = dim == 3 ? 2 : 1; // z planes that carry the "horizontal" lines
    for (int iz = 0; iz < zs; ++iz)
      {
        const int z = iz * p;
        for (int x = 0; x <= p; x += p) // lines x-, x+ : along y
          for (int i = 1; i < p; ++i)
            push(x, i, z);
        for (int y = 0; y <= p; y += p) // lines y-, y+ : along x
          for (int i = 1; i < p; ++i)
            push(i, y, z);
      }
    if (dim == 2)
      {
        for (int j = 1; j < p; ++j) // the quad, x fastest
          for (int i = 1; i < p; ++i)
            push(i, j, 0);
        return;
      }
    for (int y = 0; y <= p; y += p) // lines along z at (0,0), (p,0), (0,p), (p,p)
      for (int x = 0; x <= p; x += p)
        for (int i = 1; i < p; ++i)
          push(x, y, i);
    for (int x = 0; x <= p; x += p) // quads x-, x+ : y fastest, then z
      for (int k = 1; k < p; ++k)
        for (int j = 1; j < p; ++j)
          push(x, j, k);
    for (int y = 0; y <= p; y += p) // quads y-, y+ : z fastest, then x
      for (int i = 1; i < p; ++i)
        for (int k = 1; k < p; ++k)
          push(i, y, k);
    for (int z = 0; z <= p; z += p) // quads z-, z+ : x fastest, then y
      for (int j = 1; j < p; ++j)
        for (int i = 1; i < p; ++i)
          push(i, j, z);
    for (int k = 1; k < p; ++k) // hex, x fastest
      for (int j = 1; j < p; ++j)
        for (int i = 1; i < p; ++i)
          push(i, j, k);
  }

  // FESystem(FE_Q(p), dim) local DoF i -> scalar node (hierarchical index) and component:
  // node_of[i], comp_of[i]; and the inverse loc_of[a * dim + comp] = i
  inline void system_numbering(int dim, int p, std::vector<int> &node_of, std::vector<int> &comp_of,
                               std::vector<int> &loc_of)
  {
    int npc = 1;
    for (int d = 0; d < dim; ++d)
      npc *= p + 1;
    node_of.clear();
    comp_of.clear();
    // entities in hierarchical order with their number of scalar DoFs
    const int nv = 1 << dim, nl = dim == 2 ? 4 : 12, nqd = dim == 2 ? 1 : 6, nh = dim == 3 ? 1 : 0;
    const int per_line = p - 1, per_quad = (p - 1) * (p - 1), per_hex = (p - 1) * (p - 1) * (p - 1);
    int       next = 0;
    auto      entity = [&](int count) {
      for (int comp = 0; comp < dim; ++comp)
        for (int k = 0; k < count; ++k)
          {
            node_of.push_back(next + k);
            comp_of.push_back(comp);
          }
      next += count;
    };
    for (int v = 0; v < nv; ++v)
      entity(1);
    for (int l = 0; l < nl; ++l)
      entity(per_line);
    for (int q = 0; q < nqd; ++q)
      entity(per_quad);
    for (int h = 0; h < nh; ++h)
      entity(per_hex);
    loc_of.assign(size_t(npc) * dim, -1);
    for (int i = 0; i < int(node_of.size()); ++i)
      loc_of[node_of[i] * dim + comp_of[i]] = i;
  }
} // namespace gf_fe
