/* TEST INFRASTRUCTURE ONLY. The reference's own time-stepping lines, cut out at build time
 * (oracle/Makefile) and compiled against oracle/ref_shim:
 *   nl : Newmark coefficients (nonlinear_elasticity.h:242-250) and the member functions
 *        get_error_residual, get_error_update, get_total_solution, update_acceleration,
 *        update_velocity, update_old_variables (nonlinear_elasticity.cc:549-622)
 *   lin: the theta-scheme algebra of assemble_rhs (linear_elasticity.cc:384-420) and
 *        update_displacement (:579-586); Time is the reference's include/adapter/time_handler.h
 * Inputs on stdin, results on stdout (tests/golden/make_reference_vectors.py). */
#include <cstdio>
#include <iostream>
#include <string>

#include "assembly_shim.h"
#include "adapter/time_handler.h"

namespace Nonlinear_Elasticity
{
  using namespace dealii;
  template <int dim, typename NumberType = double>
  class Solid
  {
  public:
    struct Errors
    {
      double u = 1.0;
    };
    struct DoFs
    {
      unsigned n = 0;
      unsigned n_dofs() const { return n; }
    } dof_handler_ref;
    enum
    {
      u_dof = 0
    };
    Parameters::AllParameters            parameters;
    std::vector<types::global_dof_index> dofs_per_block;
    AffineConstraints<double>            constraints;
    BlockVector<double> system_rhs, total_displacement, total_displacement_old, velocity, velocity_old,
      acceleration, acceleration_old;
    explicit Solid(const Parameters::AllParameters &p)
      : parameters(p)
    {}
#include "nl_alpha_extract.inc"
    void                get_error_residual(Errors &error_residual);
    void                get_error_update(const BlockVector<double> &newton_update, Errors &error_update);
    BlockVector<double> get_total_solution(const BlockVector<double> &solution_delta) const;
    void                update_acceleration(BlockVector<double> &displacement_delta);
    void                update_velocity(BlockVector<double> &displacement_delta);
    void                update_old_variables();
  };
#include "nl_updates_extract.inc"
} // namespace Nonlinear_Elasticity

using namespace dealii;

static void read_vec(Vector<double> &v, unsigned n)
{
  v.reinit(n);
  for (unsigned i = 0; i < n; ++i)
    std::cin >> v[i];
}
static void print_vec(const Vector<double> &v)
{
  for (unsigned i = 0; i < v.size(); ++i)
    printf("%.17g ", v[i]);
  printf("\n");
}

static int run_nl()
{
  unsigned                  n;
  Parameters::AllParameters p;
  std::cin >> n >> p.beta >> p.gamma >> p.delta_t;
  Nonlinear_Elasticity::Solid<3, double> s(p);
  s.dof_handler_ref.n = n;
  s.dofs_per_block    = {n};
  s.constraints.constrained.resize(n);
  for (unsigned i = 0; i < n; ++i)
    {
      int c;
      std::cin >> c;
      s.constraints.constrained[i] = (unsigned char)c;
    }
  BlockVector<double> delta, newton_update;
  read_vec(delta, n);
  read_vec(s.total_displacement, n);
  read_vec(s.velocity_old, n);
  read_vec(s.acceleration_old, n);
  read_vec(s.system_rhs, n);
  read_vec(newton_update, n);
  s.acceleration.reinit(n);
  s.velocity.reinit(n);
  printf("%.17g %.17g %.17g %.17g %.17g %.17g\n", s.alpha_1, s.alpha_2, s.alpha_3, s.alpha_4,
         s.alpha_5, s.alpha_6);
  s.update_acceleration(delta);
  s.update_velocity(delta);
  print_vec(s.acceleration);
  print_vec(s.velocity);
  print_vec(s.get_total_solution(delta));
  Nonlinear_Elasticity::Solid<3, double>::Errors er, eu;
  s.get_error_residual(er);
  s.get_error_update(newton_update, eu);
  printf("%.17g %.17g\n", er.u, eu.u);
  s.update_old_variables();
  bool same = true;
  for (unsigned i = 0; i < n; ++i)
    same = same && s.total_displacement_old[i] == s.total_displacement[i] &&
           s.velocity_old[i] == s.velocity[i] && s.acceleration_old[i] == s.acceleration[i];
  printf("%d\n", same ? 1 : 0);
  return 0;
}

static int run_lin()
{
  unsigned                  n;
  int                       consistent, body_force_flag;
  double                    t_end = 1e9, dt;
  Parameters::AllParameters parameters;
  std::cin >> n >> parameters.theta >> dt >> consistent >> body_force_flag;
  parameters.data_consistent    = consistent != 0;
  const bool body_force_enabled = body_force_flag != 0;
  Adapter::Time        time(t_end, dt);
  SparseMatrix<double> stiffness_matrix, mass_matrix;
  stiffness_matrix.n = mass_matrix.n = n;
  stiffness_matrix.a.resize(size_t(n) * n);
  mass_matrix.a.resize(size_t(n) * n);
  for (auto &x : stiffness_matrix.a)
    std::cin >> x;
  for (auto &x : mass_matrix.a)
    std::cin >> x;
  Vector<double> loading, stress, old_stress, velocity, displacement, body_force_vector, new_velocity;
  read_vec(loading, n); // what assemble_consistent_loading() leaves in system_rhs
  read_vec(stress, n);
  read_vec(old_stress, n);
  read_vec(velocity, n);
  read_vec(displacement, n);
  read_vec(body_force_vector, n);
  read_vec(new_velocity, n);
  if (!std::cin)
    return 2;
  Vector<double> system_rhs(n), old_velocity(n), old_displacement(n);
  struct
  {
    unsigned n;
    unsigned n_dofs() const { return n; }
  } dof_handler{n};
  auto assemble_consistent_loading = [&]() { system_rhs = loading; };
  {
#include "lin_rhs_extract.inc"
  }
  print_vec(system_rhs);
  print_vec(old_stress);
  print_vec(old_velocity);
  print_vec(old_displacement);
  velocity = new_velocity; // the solve (:525-575) is not part of this block
  {
#include "lin_update_extract.inc"
  }
  print_vec(displacement);
  return 0;
}

int main()
{
  std::string mode;
  std::cin >> mode;
  return mode == "nl" ? run_nl() : run_lin();
}
