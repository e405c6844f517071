/* TEST INFRASTRUCTURE ONLY. Runs the reference's own per-cell assembly —
 * Assembler<dim,double>::assemble_system_tangent_residual_one_cell (nonlinear_elasticity.cc:
 * 872-1036) and Assembler_Base::assemble_neumann_contribution_one_cell (:791-859), cut out of the
 * reference at build time into oracle/_ref/assembler_extract.inc, plus PointHistory
 * (nonlinear_elasticity.h:68-117) and the material header included in place — on ONE cell whose
 * finite-element tables, nodal vectors and parameters come from stdin, and prints the cell matrix
 * and the cell right-hand side (tests/golden/make_reference_vectors.py).
 *
 * stdin (whitespace separated):
 *   dim npc nq nqf n_faces   mu nu rho alpha_1   bx by bz   interface_id
 *   N[nq][npc]  gradN[nq][npc][dim]  JxW[nq]
 *   per face: number boundary_id  Nf[nqf][npc]  JxWf[nqf]  normal[nqf][dim]
 *   u_total[dpc]  acceleration[dpc]  external_stress[dpc] */
#include <cstdio>
#include <iostream>

#include "assembly_shim.h"
#include "compressible_neo_hook_material.h"

namespace Nonlinear_Elasticity
{
  using namespace dealii;
#include "point_history_extract.inc"

  template <int dim, typename NumberType>
  struct Assembler_Base;
  template <int dim, typename NumberType>
  struct Assembler;

  // the members of Solid the assembly block reads (nonlinear_elasticity.h:197-262)
  template <int dim, typename NumberType = double>
  class Solid
  {
  public:
    struct QPH
    {
      std::vector<std::shared_ptr<const PointHistory<dim, NumberType>>> data;
      template <class It>
      std::vector<std::shared_ptr<const PointHistory<dim, NumberType>>> get_data(const It &) const
      {
        return data;
      }
    } quadrature_point_history;
    FESystem<dim>              fe;
    unsigned int               dofs_per_cell = 0;
    FEValuesExtractors::Vector u_fe;
    enum
    {
      u_dof = 0
    };
    unsigned int              n_q_points = 0, n_q_points_f = 0;
    unsigned int              boundary_interface_id = 7;
    double                    alpha_1 = 0;
    Tensor<1, 3, double>      body_force;
    AffineConstraints<double> constraints;
    BlockSparseMatrix<double> tangent_matrix;
    BlockVector<double>       system_rhs;
  };

#include "assembler_extract.inc"
} // namespace Nonlinear_Elasticity

using namespace dealii;

template <int dim>
int run(unsigned npc, unsigned nq, unsigned nqf, unsigned n_faces)
{
  using namespace Nonlinear_Elasticity;
  const unsigned            dpc = npc * dim;
  Parameters::AllParameters prm;
  double                    alpha_1, b[3];
  unsigned                  interface_id;
  std::cin >> prm.mu >> prm.nu >> prm.rho >> alpha_1 >> b[0] >> b[1] >> b[2] >> interface_id;
  ShimTables<dim> &t = ShimTables<dim>::get();
  t.nq               = nq;
  t.nqf              = nqf;
  t.npc              = npc;
  auto read          = [](std::vector<double> &v, size_t n) {
    v.resize(n);
    for (auto &x : v)
      std::cin >> x;
  };
  read(t.N, size_t(nq) * npc);
  read(t.gradN, size_t(nq) * npc * dim);
  read(t.JxW, nq);
  typename DoFHandler<dim>::Cell cell;
  cell.dofs.resize(dpc);
  for (unsigned i = 0; i < dpc; ++i)
    cell.dofs[i] = i;
  cell.faces.resize(2 * dim);
  for (unsigned f = 0; f < 2 * dim; ++f)
    cell.faces[f].number = f;
  t.Nf.assign(2 * dim, {});
  t.JxWf.assign(2 * dim, {});
  t.normal.assign(2 * dim, {});
  for (unsigned k = 0; k < n_faces; ++k)
    {
      unsigned f, id;
      std::cin >> f >> id;
      cell.faces[f].boundary = true;
      cell.faces[f].id       = id;
      read(t.Nf[f], size_t(nqf) * npc);
      read(t.JxWf[f], nqf);
      read(t.normal[f], size_t(nqf) * dim);
    }
  BlockVector<double> u(dpc), acc(dpc), stress(dpc);
  for (unsigned i = 0; i < dpc; ++i)
    std::cin >> u[i];
  for (unsigned i = 0; i < dpc; ++i)
    std::cin >> acc[i];
  for (unsigned i = 0; i < dpc; ++i)
    std::cin >> stress[i];
  if (!std::cin)
    {
      fprintf(stderr, "ref_assembler_driver: short input\n");
      return 2;
    }
  read_shim_numbering(std::cin, dpc); // optional: FESystem numbering of degree >= 3

  Solid<dim, double> solid;
  solid.fe.dofs_per_cell      = dpc;
  solid.dofs_per_cell         = dpc;
  solid.n_q_points            = nq;
  solid.n_q_points_f          = nqf;
  solid.boundary_interface_id = interface_id;
  solid.alpha_1               = alpha_1;
  for (int d = 0; d < 3; ++d)
    solid.body_force[d] = b[d];
  for (unsigned q = 0; q < nq; ++q) // setup_qph: one PointHistory per quadrature point
    {
      auto ph = std::make_shared<PointHistory<dim, double>>();
      ph->setup_lqp(prm);
      solid.quadrature_point_history.data.push_back(ph);
    }
  QGauss<dim> qf_cell;
  qf_cell.n = nq;
  QGauss<dim - 1> qf_face;
  qf_face.n = nqf;
  typename Assembler_Base<dim, double>::PerTaskData_ASM per_task(&solid);
  typename Assembler_Base<dim, double>::ScratchData_ASM scratch(
    solid.fe, qf_cell, update_values | update_gradients, qf_face, update_values, u, acc, stress);
  Assembler<dim, double> assembler;
  assembler.assemble_system_one_cell(&cell, scratch, per_task); // :749-757
  for (unsigned i = 0; i < dpc; ++i)
    {
      for (unsigned j = 0; j < dpc; ++j)
        printf("%.17g ", per_task.cell_matrix(i, j));
      printf("\n");
    }
  for (unsigned i = 0; i < dpc; ++i)
    printf("%.17g ", per_task.cell_rhs(i));
  printf("\n");
  return 0;
}

int main()
{
  unsigned dim, npc, nq, nqf, n_faces;
  std::cin >> dim >> npc >> nq >> nqf >> n_faces;
  return dim == 2 ? run<2>(npc, nq, nqf, n_faces) : run<3>(npc, nq, nqf, n_faces);
}
