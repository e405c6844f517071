/* TEST INFRASTRUCTURE ONLY. How the reference configures its linear solves:
 * Solid::solve_linear_system (nonlinear_elasticity.cc:1153-1211) and ElastoDynamics::solve
 * (linear_elasticity.cc:525-575), member definitions cut out at build time, run against RECORDING
 * stand-ins of SolverControl / SolverCG / PreconditionSelector / PreconditionSSOR /
 * SparseDirectUMFPACK: prints the iteration limit and tolerance handed to SolverControl, the
 * preconditioner and its relaxation, the calls made, and what the function reports back.
 *   usage: ref_solver_driver nl|lin CG|Direct <n_dofs> <max_iterations_lin> <tol_lin> <rhs l2 norm> */
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

#include "assembly_shim.h"

namespace dealii
{
  inline const char *ExcNotImplemented() { return "ExcNotImplemented"; }
  struct SolverControl
  {
    SolverControl(int n, double t) { printf("SolverControl %d %.17g\n", n, t); }
    unsigned last_step() const { return 17; } // stands for whatever a run reports
    double   last_value() const { return 0.125; }
  };
  template <class V = Vector<double>>
  struct GrowingVectorMemory
  {};
  struct MatrixShim
  {
    unsigned          n = 0;
    unsigned          m() const { return n; }
    const MatrixShim &block(unsigned, unsigned) const { return *this; }
  };
  template <class M, class V>
  struct PreconditionSelector
  {
    PreconditionSelector(const std::string &name, double omega)
    {
      printf("PreconditionSelector %s %.17g\n", name.c_str(), omega);
    }
    template <class A>
    void use_matrix(const A &)
    {}
  };
  template <class M = MatrixShim>
  struct PreconditionSSOR
  {
    template <class A>
    void initialize(const A &, double omega)
    {
      printf("PreconditionSSOR %.17g\n", omega);
    }
  };
  template <class V = Vector<double>>
  struct SolverCG
  {
    template <class C, class G>
    SolverCG(C &, G &)
    {}
    template <class M, class X, class B, class P>
    void solve(const M &, X &, const B &, const P &)
    {
      printf("SolverCG::solve\n");
    }
  };
  struct SparseDirectUMFPACK
  {
    template <class M>
    void initialize(const M &)
    {
      printf("SparseDirectUMFPACK::initialize\n");
    }
    template <class X, class B>
    void vmult(X &, const B &)
    {
      printf("SparseDirectUMFPACK::vmult\n");
    }
  };
  struct ConstraintsShim
  {
    template <class V>
    void distribute(V &) const
    {
      printf("constraints.distribute\n");
    }
  };
  struct TimerShim
  {
    void enter_subsection(const std::string &) {}
    void leave_subsection(const std::string & = "") {}
  };
  struct VelocityShim : Vector<double>
  {
    using Vector<double>::Vector;
    double linfty_norm() const { return 0.0; }
  };
} // namespace dealii

struct SolverParameters
{
  std::string type_lin;
  double      max_iterations_lin = 1, tol_lin = 1e-6;
};

namespace Nonlinear_Elasticity
{
  using namespace dealii;
  template <int dim, typename NumberType = double>
  class Solid
  {
  public:
    SolverParameters                     parameters;
    std::vector<types::global_dof_index> dofs_per_block;
    enum
    {
      u_dof = 0
    };
    MatrixShim          tangent_matrix;
    BlockVector<double> system_rhs;
    ConstraintsShim     constraints;
    mutable TimerShim   timer;
    std::pair<unsigned int, double> solve_linear_system(BlockVector<double> &newton_update);
  };
#include "nl_solver_extract.inc"
} // namespace Nonlinear_Elasticity

namespace Linear_Elasticity
{
  using namespace dealii;
  template <int dim>
  class ElastoDynamics
  {
  public:
    SolverParameters  parameters;
    MatrixShim        system_matrix;
    VelocityShim      velocity;
    Vector<double>    system_rhs;
    ConstraintsShim   hanging_node_constraints;
    mutable TimerShim timer;
    void              solve();
  };
#include "lin_solver_extract.inc"
} // namespace Linear_Elasticity

int main(int argc, char **argv)
{
  if (argc < 7)
    return 2;
  const std::string  solver = argv[1];
  const unsigned     n      = unsigned(atoi(argv[3]));
  std::ostringstream sink;
  auto *             old = std::cout.rdbuf(sink.rdbuf());
  if (solver == "nl")
    {
      Nonlinear_Elasticity::Solid<3> s;
      s.parameters.type_lin           = argv[2];
      s.parameters.max_iterations_lin = atof(argv[4]);
      s.parameters.tol_lin            = atof(argv[5]);
      s.dofs_per_block                = {1};
      s.tangent_matrix.n              = n;
      s.system_rhs                    = dealii::BlockVector<double>(s.dofs_per_block);
      s.system_rhs[0]                 = atof(argv[6]); // l2 norm of a one-entry vector
      dealii::BlockVector<double> upd(s.dofs_per_block);
      const auto                  r = s.solve_linear_system(upd);
      printf("returns %u %.17g\n", r.first, r.second);
    }
  else
    {
      Linear_Elasticity::ElastoDynamics<3> s;
      s.parameters.type_lin           = argv[2];
      s.parameters.max_iterations_lin = atof(argv[4]);
      s.system_matrix.n               = n;
      s.solve();
    }
  std::cout.rdbuf(old);
  // the linear solver prints what it reports ("No of iterations", "Final residual")
  std::istringstream in(sink.str());
  std::string        line;
  while (std::getline(in, line))
    if (line.find("No of iterations") != std::string::npos || line.find("Final residual") != std::string::npos)
      printf("prints %s\n", line.c_str());
  return 0;
}
