/* TEST INFRASTRUCTURE ONLY. Runs three header-only pieces of the reference, compiled IN PLACE from
 * /root/reference against oracle/ref_shim (see deal.II/base/dealii_min.h), and prints their outputs
 * for tests/golden/make_reference_vectors.py:
 *   material <dim> <mu> <nu> <rho> <det_F> <b_bar in deal.II order 00,11,[22],01,[02,12]>
 *       -> psi, tau (n), Jc (n x n, row-major) in the same component order
 *   strain <dim> : reads u (dim) and grad u (dim x dim, row-major) from stdin per point
 *       -> the dim + dim*dim computed_quantities of Postprocessor::evaluate_vector_field
 *   time <t_end> <dt> <n_increments> <t_reset>  -> timestep/current after increments and after
 *       set_absolute_time(t_reset) */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "compressible_neo_hook_material.h"
#include "postprocessor.h"
#include "adapter/time_handler.h"

using namespace dealii;

template <int dim>
int material(char **a)
{
  const double mu = atof(a[0]), nu = atof(a[1]), rho = atof(a[2]), det_F = atof(a[3]);
  constexpr int n = dim * (dim + 1) / 2;
  const int     vi[6] = {0, 1, dim == 3 ? 2 : 0, 0, 0, 1}, vj[6] = {0, 1, dim == 3 ? 2 : 1, 1, 2, 2};
  int           I[n], J[n];
  for (int k = 0; k < n; ++k)
    {
      I[k] = k < dim ? k : (dim == 2 ? 0 : vi[k]);
      J[k] = k < dim ? k : (dim == 2 ? 1 : vj[k]);
    }
  SymmetricTensor<2, dim> b;
  for (int k = 0; k < n; ++k)
    b.v[I[k]][J[k]] = b.v[J[k]][I[k]] = atof(a[4 + k]);
  Nonlinear_Elasticity::Material_Compressible_Neo_Hook_One_Field<dim, double> m(mu, nu, rho);
  const double                  psi = m.get_Psi(det_F, b);
  const SymmetricTensor<2, dim> tau = m.get_tau(det_F, b);
  const SymmetricTensor<4, dim> Jc  = m.get_Jc(det_F, b);
  printf("%.17g\n", psi);
  for (int k = 0; k < n; ++k)
    printf("%.17g ", tau.v[I[k]][J[k]]);
  printf("\n");
  for (int k = 0; k < n; ++k)
    {
      for (int l = 0; l < n; ++l)
        printf("%.17g ", Jc.v[I[k]][J[k]][I[l]][J[l]]);
      printf("\n");
    }
  // minor and major symmetry of what the reference formula produced (full storage shows it)
  double asym = 0;
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j < dim; ++j)
      for (int k = 0; k < dim; ++k)
        for (int l = 0; l < dim; ++l)
          {
            asym = std::max(asym, std::fabs(Jc.v[i][j][k][l] - Jc.v[j][i][k][l]));
            asym = std::max(asym, std::fabs(Jc.v[i][j][k][l] - Jc.v[i][j][l][k]));
            asym = std::max(asym, std::fabs(Jc.v[i][j][k][l] - Jc.v[k][l][i][j]));
          }
  printf("%.3g\n", asym);
  return 0;
}

template <int dim>
int strain()
{
  DataPostprocessorInputs::Vector<dim> in;
  double                               x;
  while (std::cin >> x)
    {
      Vector<double> u(dim);
      u[0] = x;
      for (int d = 1; d < dim; ++d)
        std::cin >> u[d];
      std::vector<Tensor<1, dim>> g(dim);
      for (int d = 0; d < dim; ++d)
        for (int e = 0; e < dim; ++e)
          std::cin >> g[d][e];
      in.solution_values.push_back(u);
      in.solution_gradients.push_back(g);
    }
  std::vector<Vector<double>> out(in.solution_values.size(), Vector<double>(dim * dim + dim));
  Nonlinear_Elasticity::Postprocessor<dim> pp;
  pp.evaluate_vector_field(in, out);
  for (const auto &n : pp.get_names())
    printf("%s ", n.c_str());
  printf("\n");
  for (const auto &q : out)
    {
      for (unsigned k = 0; k < q.size(); ++k)
        printf("%.17g ", q[k]);
      printf("\n");
    }
  return 0;
}

int main(int argc, char **argv)
{
  if (argc >= 3 && !strcmp(argv[1], "material"))
    return atoi(argv[2]) == 2 ? material<2>(argv + 3) : material<3>(argv + 3);
  if (argc >= 3 && !strcmp(argv[1], "strain"))
    return atoi(argv[2]) == 2 ? strain<2>() : strain<3>();
  if (argc >= 6 && !strcmp(argv[1], "time"))
    {
      Adapter::Time t(atof(argv[2]), atof(argv[3]));
      for (int k = 0; k < atoi(argv[4]); ++k)
        t.increment();
      printf("%u %.17g %.17g %.17g\n", t.get_timestep(), t.current(), t.end(), t.get_delta_t());
      t.set_absolute_time(atof(argv[5]));
      printf("%u %.17g\n", t.get_timestep(), t.current());
      return 0;
    }
  fprintf(stderr, "usage: ref_driver material|strain|time ...\n");
  return 2;
}
