/* TEST INFRASTRUCTURE ONLY. Which Dirichlet sets the reference asks deal.II for:
 * Solid::make_constraints (nonlinear_elasticity.cc:1094-1150, member definition cut out at build
 * time) and the boundary-value block of ElastoDynamics::assemble_rhs (linear_elasticity.cc:429-446,
 * statement block cut out), run against a RECORDING VectorTools::interpolate_boundary_values:
 * prints one line per call, `boundary_id component-mask-bits`, plus clear/close of the constraints.
 *   usage: ref_constraints_driver nl <dim> <newton iteration>  |  lin <dim> */
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <sstream>
#include <string>

#include "assembly_shim.h"

static void rec(const char *what, unsigned id, unsigned mask) { printf("%s %u %u\n", what, id, mask); }

namespace dealii
{
  struct ComponentMask
  {
    unsigned bits;
  };
  namespace FEValuesExtractors
  {
    struct Scalar
    {
      unsigned component;
      explicit Scalar(unsigned c)
        : component(c)
      {}
    };
  } // namespace FEValuesExtractors
  template <int dim>
  struct FEShim
  {
    ComponentMask component_mask(const FEValuesExtractors::Scalar &s) const { return {1u << s.component}; }
    ComponentMask component_mask(const FEValuesExtractors::Vector &v) const
    {
      return {((1u << dim) - 1u) << v.first_vector_component};
    }
  };
  namespace Functions
  {
    template <int dim>
    struct ZeroFunction
    {
      unsigned n_components;
      explicit ZeroFunction(unsigned n = 1)
        : n_components(n)
      {}
    };
  } // namespace Functions
  struct RecordingConstraints
  {
    void clear() { printf("clear\n"); }
    void close() { printf("close\n"); }
  };
  namespace VectorTools
  {
    // without a mask deal.II constrains every component
    template <class DH, int dim, class Target>
    void interpolate_boundary_values(const DH &, unsigned id, const Functions::ZeroFunction<dim> &f,
                                     Target &)
    {
      rec("interpolate_boundary_values", id, (1u << f.n_components) - 1u);
    }
    template <class DH, int dim, class Target>
    void interpolate_boundary_values(const DH &, unsigned id, const Functions::ZeroFunction<dim> &,
                                     Target &, const ComponentMask &m)
    {
      rec("interpolate_boundary_values", id, m.bits);
    }
  } // namespace VectorTools
} // namespace dealii

namespace Nonlinear_Elasticity
{
  using namespace dealii;
  template <int dim, typename NumberType = double>
  class Solid
  {
  public:
    int                              dof_handler_ref = 0;
    RecordingConstraints             constraints;
    FEShim<dim>                      fe;
    const FEValuesExtractors::Vector u_fe{0};
    static const unsigned int        n_components                 = dim;
    const unsigned int               clamped_boundary_id          = 1; // nonlinear_elasticity.h:255-256
    const unsigned int               out_of_plane_clamped_mesh_id = 8;
    void                             make_constraints(const int &it_nr);
  };
#include "nl_constraints_extract.inc"
} // namespace Nonlinear_Elasticity

template <int dim>
int run_lin()
{
  using namespace dealii;
  int          dof_handler = 0;
  FEShim<dim>  fe;
  unsigned int clamped_mesh_id = 0, out_of_plane_clamped_mesh_id = 4; // linear_elasticity.cc:157-158
#include "lin_bv_extract.inc"
  return 0;
}

int main(int argc, char **argv)
{
  if (argc < 3)
    return 2;
  const std::string solver = argv[1];
  const int         dim    = atoi(argv[2]);
  if (solver == "lin")
    return dim == 2 ? run_lin<2>() : run_lin<3>();
  const int          it = atoi(argv[3]);
  std::ostringstream sink; // " CST "
  auto *             old = std::cout.rdbuf(sink.rdbuf());
  if (dim == 2)
    {
      Nonlinear_Elasticity::Solid<2> s;
      s.make_constraints(it);
    }
  else
    {
      Nonlinear_Elasticity::Solid<3> s;
      s.make_constraints(it);
    }
  std::cout.rdbuf(old);
  return 0;
}
