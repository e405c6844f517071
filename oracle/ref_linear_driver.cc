/* TEST INFRASTRUCTURE ONLY. Runs two statement blocks of the reference's linear solver, cut out of
 * linear_elasticity.cc at build time (oracle/Makefile) and compiled against oracle/ref_shim:
 *   the local stiffness loops of ElastoDynamics::assemble_system       (:289-323)
 *   the interface-face loop of ElastoDynamics::assemble_consistent_loading (:487-512)
 * on ONE cell whose finite-element tables and nodal stress come from stdin; prints the cell matrix
 * and the cell right-hand side (tests/golden/make_reference_vectors.py).
 *
 * stdin: dim npc nq nqf n_faces  lambda mu  interface_id
 *        gradN[nq][npc][dim] JxW[nq]
 *        per face: number boundary_id Nf[nqf][npc] JxWf[nqf]
 *        stress[dpc] */
#include <cstdio>
#include <iostream>

#include "assembly_shim.h"

using namespace dealii;

template <int dim>
int run(unsigned npc, unsigned nq, unsigned nqf, unsigned n_faces)
{
  double   lambda_value, mu_value;
  unsigned interface_boundary_id;
  std::cin >> lambda_value >> mu_value >> interface_boundary_id;
  ShimTables<dim> &t = ShimTables<dim>::get();
  t.nq               = nq;
  t.nqf              = nqf;
  t.npc              = npc;
  auto read          = [](std::vector<double> &v, size_t n) {
    v.resize(n);
    for (auto &x : v)
      std::cin >> x;
  };
  read(t.gradN, size_t(nq) * npc * dim);
  read(t.JxW, nq);
  t.N.assign(size_t(nq) * npc, 0.0);
  typename DoFHandler<dim>::Cell cell_object;
  const unsigned                 dofs_per_cell = npc * dim;
  cell_object.dofs.resize(dofs_per_cell);
  for (unsigned i = 0; i < dofs_per_cell; ++i)
    cell_object.dofs[i] = i;
  cell_object.faces.resize(2 * dim);
  for (unsigned f = 0; f < 2 * dim; ++f)
    cell_object.faces[f].number = f;
  t.Nf.assign(2 * dim, {});
  t.JxWf.assign(2 * dim, {});
  t.normal.assign(2 * dim, std::vector<double>(size_t(nqf) * dim, 0.0));
  for (unsigned k = 0; k < n_faces; ++k)
    {
      unsigned f, id;
      std::cin >> f >> id;
      cell_object.faces[f].boundary = true;
      cell_object.faces[f].id       = id;
      read(t.Nf[f], size_t(nqf) * npc);
      read(t.JxWf[f], nqf);
    }
  Vector<double> stress(dofs_per_cell);
  for (unsigned i = 0; i < dofs_per_cell; ++i)
    std::cin >> stress[i];
  if (!std::cin)
    {
      fprintf(stderr, "ref_linear_driver: short input\n");
      return 2;
    }
  read_shim_numbering(std::cin, dofs_per_cell); // optional: FESystem numbering of degree >= 3
  FESystem<dim> fe;
  fe.dofs_per_cell = dofs_per_cell;
  const typename DoFHandler<dim>::active_cell_iterator cell = &cell_object;

  // ---- assemble_system: the objects of :252-272, then the reference's loops -------------------
  {
    QGauss<dim> quadrature_formula;
    quadrature_formula.n = nq;
    FEValues<dim> fe_values(fe, quadrature_formula, update_values | update_gradients);
    const unsigned int  n_q_points = quadrature_formula.size();
    FullMatrix<double>  cell_matrix(dofs_per_cell, dofs_per_cell);
    std::vector<double> lambda_values(n_q_points, lambda_value); // ConstantFunction::value_list
    std::vector<double> mu_values(n_q_points, mu_value);
    cell_matrix = 0;
    fe_values.reinit(cell);
#include "lin_stiffness_extract.inc"
    for (unsigned i = 0; i < dofs_per_cell; ++i)
      {
        for (unsigned j = 0; j < dofs_per_cell; ++j)
          printf("%.17g ", cell_matrix(i, j));
        printf("\n");
      }
  }
  // ---- assemble_consistent_loading: the objects of :464-481, then the reference's face loop ---
  {
    QGauss<dim - 1> face_quadrature_formula;
    face_quadrature_formula.n = nqf;
    FEFaceValues<dim>           fe_face_values(fe, face_quadrature_formula, update_values);
    const unsigned int          n_face_q_points = face_quadrature_formula.size();
    Vector<double>              cell_rhs(dofs_per_cell);
    std::vector<Vector<double>> local_stress(n_face_q_points, Vector<double>(dim));
    cell_rhs = 0;
#include "lin_loading_extract.inc"
    for (unsigned i = 0; i < dofs_per_cell; ++i)
      printf("%.17g ", cell_rhs(i));
    printf("\n");
  }
  return 0;
}

int main()
{
  unsigned dim, npc, nq, nqf, n_faces;
  std::cin >> dim >> npc >> nq >> nqf >> n_faces;
  return dim == 2 ? run<2>(npc, nq, nqf, n_faces) : run<3>(npc, nq, nqf, n_faces);
}
