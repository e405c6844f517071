/* TEST INFRASTRUCTURE ONLY. The reference's make_grid of both solver classes
 * (nonlinear_elasticity.cc:169-285, linear_elasticity.cc:79-187), member definitions cut out at
 * build time, run against a recording stand-in for Triangulation / GridGenerator: ONE cell whose
 * 2*dim faces carry the colorize ids 0..2*dim-1 of subdivided_hyper_rectangle. Prints what the
 * reference asked the grid generator for (repetitions, box corners, refinements) and the boundary
 * id every colorized face ends up with.
 *   usage: ref_grid_driver nl|lin <dim> FSI3|PF <flap_location> */
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

#include "assembly_shim.h"

namespace dealii
{
  template <int dim>
  class Point
  {
  public:
    double x[3] = {0, 0, 0};
    Point() = default;
    Point(double a, double b)
    {
      x[0] = a;
      x[1] = b;
    }
    Point(double a, double b, double c)
    {
      x[0] = a;
      x[1] = b;
      x[2] = c;
    }
  };
  template <int dim>
  class Triangulation
  {
  public:
    struct Face
    {
      unsigned id = 0;
      bool     at_boundary() const { return true; }
      unsigned boundary_id() const { return id; }
      void     set_boundary_id(unsigned i) { id = i; }
    };
    struct Cell
    {
      std::vector<Face>   faces;
      std::vector<Face *> face_iterators()
      {
        std::vector<Face *> r;
        for (auto &f : faces)
          r.push_back(&f);
        return r;
      }
    };
    Cell                      cell;
    std::vector<unsigned int> repetitions;
    Point<dim>                p1, p2;
    bool                      colorize = false;
    unsigned                  refinements = 0;
    void                      refine_global(unsigned n) { refinements += n; }
    std::vector<Cell *>       active_cell_iterators() { return {&cell}; }
  };
  namespace GridGenerator
  {
    template <int dim>
    void subdivided_hyper_rectangle(Triangulation<dim> &t, const std::vector<unsigned int> &reps,
                                    const Point<dim> &p1, const Point<dim> &p2, bool colorize)
    {
      t.repetitions = reps;
      t.p1          = p1;
      t.p2          = p2;
      t.colorize    = colorize;
      t.cell.faces.resize(2 * dim);
      for (unsigned f = 0; f < 2 * dim; ++f)
        t.cell.faces[f].id = colorize ? f : 0; // colorize: 0..5 = x-, x+, y-, y+, z-, z+
    }
  } // namespace GridGenerator
  namespace GridTools
  {
    template <int dim>
    double volume(const Triangulation<dim> &t)
    {
      double v = 1;
      for (int d = 0; d < dim; ++d)
        v *= t.p2.x[d] - t.p1.x[d];
      return v;
    }
  } // namespace GridTools
} // namespace dealii

struct AdapterShim
{
  unsigned deal_boundary_interface_id;
};

namespace Nonlinear_Elasticity
{
  using namespace dealii;
  template <int dim, typename NumberType = double>
  class Solid
  {
  public:
    Parameters::AllParameters parameters;
    Tensor<1, 3, double>      body_force;
    Triangulation<dim>        triangulation;
    const unsigned int        boundary_interface_id        = 7; // nonlinear_elasticity.cc:78
    const unsigned int        clamped_boundary_id          = 1; // nonlinear_elasticity.h:255-256
    const unsigned int        out_of_plane_clamped_mesh_id = 8;
    AdapterShim               adapter{7};
    double                    vol_reference = 0, vol_current = 0;
    void                      make_grid();
  };
#include "nl_grid_extract.inc"
} // namespace Nonlinear_Elasticity

namespace Linear_Elasticity
{
  using namespace dealii;
  template <int dim>
  class ElastoDynamics
  {
  public:
    Parameters::AllParameters parameters;
    Triangulation<dim>        triangulation;
    unsigned int              clamped_mesh_id = 99, out_of_plane_clamped_mesh_id = 99;
    const unsigned int        interface_boundary_id = 6; // linear_elasticity.cc:57
    AdapterShim               adapter{6};
    void                      make_grid();
  };
#include "lin_grid_extract.inc"
} // namespace Linear_Elasticity

template <class S, int dim>
int report(S &s, unsigned interface_id, unsigned clamped, unsigned zclamped)
{
  printf("%d %u %u %u %u\n", dim, s.triangulation.refinements, interface_id, clamped, zclamped);
  for (unsigned r : s.triangulation.repetitions)
    printf("%u ", r);
  printf("\n");
  for (int d = 0; d < dim; ++d)
    printf("%.17g ", s.triangulation.p1.x[d]);
  printf("\n");
  for (int d = 0; d < dim; ++d)
    printf("%.17g ", s.triangulation.p2.x[d]);
  printf("\n");
  for (const auto &f : s.triangulation.cell.faces)
    printf("%u ", f.id);
  printf("\n");
  return 0;
}

template <int dim>
int run(const std::string &solver, const std::string &scenario, double flap)
{
  if (solver == "nl")
    {
      Nonlinear_Elasticity::Solid<dim, double> s;
      s.parameters.scenario      = scenario;
      s.parameters.flap_location = flap;
      std::ostringstream sink; // "Grid: Reference volume ..." goes here
      auto *             old = std::cout.rdbuf(sink.rdbuf());
      s.make_grid();
      std::cout.rdbuf(old);
      printf("%.17g\n", s.vol_reference);
      return report<decltype(s), dim>(s, s.boundary_interface_id, s.clamped_boundary_id,
                                      s.out_of_plane_clamped_mesh_id);
    }
  Linear_Elasticity::ElastoDynamics<dim> s;
  s.parameters.scenario      = scenario;
  s.parameters.flap_location = flap;
  s.make_grid();
  printf("%.17g\n", dealii::GridTools::volume(s.triangulation));
  return report<decltype(s), dim>(s, s.interface_boundary_id, s.clamped_mesh_id,
                                  s.out_of_plane_clamped_mesh_id);
}

int main(int argc, char **argv)
{
  if (argc < 5)
    return 2;
  const int dim = atoi(argv[2]);
  return dim == 2 ? run<2>(argv[1], argv[3], atof(argv[4])) : run<3>(argv[1], argv[3], atof(argv[4]));
}
