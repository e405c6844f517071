/* TEST INFRASTRUCTURE ONLY. The reference's Newton driver — Solid::solve_nonlinear_timestep
 * (nonlinear_elasticity.cc:410-499) with its Errors struct (nonlinear_elasticity.h:293-313), both
 * cut out at build time — run against SCRIPTED residual / update norms: the members it calls
 * (make_constraints, update_acceleration, assemble_system, get_error_residual,
 * solve_linear_system, get_error_update) are stubs that hand out the next scripted value, so the
 * output is purely the reference's control flow: how many linear solves it performs, the
 * normalised errors it derives, and whether it ends with "No convergence in nonlinear solver!".
 *
 * stdin: max_iterations_NR tol_f tol_u n  residual[n]  update[n]
 * stdout: solves converged(1)/threw(0), then per solve: res_norm res_abs upd_norm upd_abs */
#include <cstdio>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "assembly_shim.h"
#include "adapter/time_handler.h"

namespace Nonlinear_Elasticity
{
  using namespace dealii;
  template <int dim, typename NumberType = double>
  class Solid
  {
  public:
#include "nl_errors_extract.inc"
    Errors error_residual, error_residual_0, error_residual_norm, error_update, error_update_0,
      error_update_norm;
    Parameters::AllParameters            parameters;
    Adapter::Time                        time{1e9, 0.01};
    std::vector<types::global_dof_index> dofs_per_block{4};
    BlockVector<double>                  acceleration, external_stress;
    // scripted norms and the record of what the driver derived from them
    std::vector<double>              script_residual, script_update;
    unsigned                         n_assemblies = 0, n_solves = 0;
    std::vector<std::vector<double>> rows;

    void print_conv_header() {}
    void print_conv_footer() {}
    void make_constraints(const int &) {}
    void update_acceleration(BlockVector<double> &) {}
    void assemble_system(const BlockVector<double> &, const BlockVector<double> &,
                         const BlockVector<double> &)
    {
      ++n_assemblies;
    }
    void get_error_residual(Errors &e) { e.u = script_residual.at(n_assemblies - 1); }
    std::pair<unsigned int, double> solve_linear_system(BlockVector<double> &)
    {
      ++n_solves;
      return {1u, 0.0};
    }
    void get_error_update(const BlockVector<double> &, Errors &e) { e.u = script_update.at(n_solves - 1); }
    void solve_nonlinear_timestep(BlockVector<double> &solution_delta);
  };
  // the driver prints its table to std::cout: keep it, the test reads the table from stderr-free
  // stdout AFTER a marker line
#include "nl_newton_extract.inc"
} // namespace Nonlinear_Elasticity

int main()
{
  using namespace Nonlinear_Elasticity;
  Solid<3, double> s;
  unsigned         n;
  std::cin >> s.parameters.max_iterations_NR >> s.parameters.tol_f >> s.parameters.tol_u >> n;
  s.script_residual.resize(n);
  s.script_update.resize(n);
  for (auto &x : s.script_residual)
    std::cin >> x;
  for (auto &x : s.script_update)
    std::cin >> x;
  dealii::BlockVector<double> delta(4);
  // silence the reference's own table; its last derived errors are read from the object
  std::ostringstream sink;
  auto *             old = std::cout.rdbuf(sink.rdbuf());
  int                converged = 1;
  try
    {
      s.solve_nonlinear_timestep(delta);
    }
  catch (const std::exception &e)
    {
      converged = 0;
    }
  std::cout.rdbuf(old);
  printf("%u %u %d\n", s.n_solves, s.n_assemblies, converged);
  printf("%.17g %.17g %.17g %.17g\n", s.error_residual_norm.u, s.error_residual.u,
         s.error_update_norm.u, s.error_update.u);
  // the table rows the reference printed (scientific, 3 digits): kept for the record
  std::istringstream in(sink.str());
  std::string        line;
  while (std::getline(in, line))
    if (line.find(" | ") != std::string::npos)
      printf("ROW%s\n", line.c_str());
  return 0;
}
