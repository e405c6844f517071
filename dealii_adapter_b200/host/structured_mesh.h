// Structured box mesh + DoF enumeration: the host-side stand-in for what deal.II's
// Triangulation / DoFHandler / FESystem(FE_Q(p),dim) hand to the solver classes.
//
// It reproduces, for GridGenerator::subdivided_hyper_rectangle(colorize=true)
// (reference call sites: nonlinear_elasticity.cc:237-241, linear_elasticity.cc:143-147):
//   * cells in lexicographic order (x fastest), vertices in deal.II order (x fastest),
//   * the FESystem(FE_Q(p),dim) LOCAL DoF order: vertex dofs (vertex-major, component-minor),
//     then line, quad, hex dofs (entity-major, then component, then the entity's scalar DoFs),
//     any polynomial degree (support points: Gauss-Lobatto for p >= 3 as FE_Q does),
//   * colorize boundary ids 0..5 = x-,x+,y-,y+,z-,z+  (used at nonlinear_elasticity.cc:200-230),
//   * DoFTools::extract_boundary_dofs per component as ascending index lists (adapter.h:247-275),
//   * boundary support points (dof_tools_extension.h:18-75).
// This is host scaffolding (deal.II is not available in the build image); the device library
// (include/graft_fem.h) is numbering-agnostic and only sees the arrays produced here.
#pragma once
#include <array>
#include <cstdint>
#include <vector>

namespace gfh
{
  enum Numbering
  {
    // analogue of DoFHandler::distribute_dofs: cell by cell, local order, first visit numbers
    numbering_cellwise = 0,
    // cellwise followed by DoFRenumbering::component_wise (nonlinear_elasticity.cc:318)
    numbering_component_wise = 1,
    // lexicographic grid nodes, components interleaved
    numbering_lexicographic = 2
  };

  struct StructuredMesh
  {
    int dim    = 0;
    int degree = 0;
    int reps[3]{1, 1, 1};
    double p0[3]{0, 0, 0}, p1[3]{0, 0, 0};
    int numbering = numbering_cellwise;

    int     nodes_per_cell = 0; // (p+1)^dim
    int     dofs_per_cell  = 0; // dim*(p+1)^dim
    int64_t n_cells = 0, n_nodes = 0, n_dofs = 0;
    int64_t grid_nodes[3]{1, 1, 1}; // reps*p+1 per direction

    // hierarchical (deal.II FE_Q) local scalar node -> local lexicographic (lx,ly,lz)
    std::vector<std::array<int, 3>> local_node_lex;
    // FESystem local DoF of (hierarchical node a, component c): local_dof_of[a * dim + c]
    // (= a * dim + c for p <= 2; entity-major / component / entity DoF beyond)
    std::vector<int> local_dof_of;
    // the p + 1 support points of FE_Q(p) on [0, 1] (Gauss-Lobatto; equidistant for p <= 2)
    std::vector<double> unit_support_1d;
    static constexpr int max_degree = 8;

    std::vector<int32_t> cell_dofs;     // [n_cells*dofs_per_cell]
    std::vector<double>  cell_vertices; // [n_cells*2^dim*dim]
    std::vector<int64_t> grid_to_node;  // lexicographic grid node -> node number
    std::vector<int64_t> node_to_grid;  // inverse
    std::vector<double>  support_points; // [n_dofs*dim]

    StructuredMesh(int dim, int degree, const int *reps, const double *p0, const double *p1,
                   int numbering);

    int32_t dof_of(int64_t node, int comp) const
    {
      return numbering == numbering_component_wise ? int32_t(comp * n_nodes + node) :
                                                     int32_t(node * dim + comp);
    }
    int64_t cell_index(int i, int j, int k) const
    {
      return (int64_t(k) * reps[1] + j) * reps[0] + i;
    }
    // colorize ids touching a grid node (bit f set if node lies on boundary face id f)
    unsigned node_face_mask(int64_t grid_node) const;

    // OR into mask[n_dofs] all dofs of components in comp_mask on faces in face_mask
    // (VectorTools::interpolate_boundary_values with ZeroFunction; nonlinear_elasticity.cc:1111-1146)
    int64_t boundary_dof_mask(unsigned face_mask, unsigned comp_mask, uint8_t *mask) const;
    // boundary faces (cell, face_no) with colorize id in face_mask, cell order then face order
    void boundary_faces(unsigned face_mask, std::vector<int32_t> &cells,
                        std::vector<int32_t> &faces) const;
    // per-component ascending dof lists of all nodes on faces in face_mask; out[c*n+i]
    int64_t interface_dofs(unsigned face_mask, std::vector<int32_t> &out) const;
  };

  // 1-D slab partition along `axis` into nparts; rank's local mesh arrays with
  // [owned | ghost] dof numbering and neighbour exchange lists.
  struct MeshPartition
  {
    int     rank = 0, nparts = 1, axis = 1;
    int64_t n_local_dofs = 0, n_owned_dofs = 0;
    int64_t n_local_cells = 0;
    std::vector<int32_t> cell_dofs;      // local cells (owned + ghost layer), local dof ids
    std::vector<double>  cell_vertices;
    std::vector<int64_t> local_cell_global; // global cell index of each local cell
    std::vector<int32_t> local_to_global;   // local dof -> global dof
    std::vector<int32_t> nbr_rank;          // neighbour ranks
    std::vector<int64_t> send_ptr, recv_ptr; // CSR offsets per neighbour
    std::vector<int32_t> send_dofs, recv_dofs; // local dof ids
  };
  MeshPartition partition_mesh(const StructuredMesh &mesh, int axis, int nparts, int rank);
} // namespace gfh

extern "C"
{
  void *      gfh_mesh_create(int dim, int degree, const int *reps, const double *p0,
                              const double *p1, int numbering);
  void        gfh_mesh_destroy(void *m);
  int64_t     gfh_mesh_n_cells(const void *m);
  int64_t     gfh_mesh_n_dofs(const void *m);
  int64_t     gfh_mesh_n_nodes(const void *m);
  int         gfh_mesh_dofs_per_cell(const void *m);
  const int32_t *gfh_mesh_cell_dofs(const void *m);
  const double * gfh_mesh_cell_vertices(const void *m);
  const double * gfh_mesh_support_points(const void *m);
  int64_t gfh_mesh_boundary_dof_mask(const void *m, unsigned face_mask, unsigned comp_mask,
                                     uint8_t *mask);
  // pass cells=faces=NULL to query the count
  int64_t gfh_mesh_boundary_faces(const void *m, unsigned face_mask, int32_t *cells,
                                  int32_t *faces);
  // pass out=NULL to query n_interface_nodes; out has dim*n entries, component-major
  int64_t gfh_mesh_interface_dofs(const void *m, unsigned face_mask, int32_t *out);

  void *  gfh_partition_create(const void *m, int axis, int nparts, int rank);
  void    gfh_partition_destroy(void *p);
  // which: 0 n_local_dofs, 1 n_owned_dofs, 2 n_local_cells, 3 n_neighbours, 4 n_send, 5 n_recv
  int64_t gfh_partition_size(const void *p, int which);
  // which: 0 cell_dofs(i32) 1 cell_vertices(f64) 2 local_cell_global(i64) 3 local_to_global(i32)
  //        4 nbr_rank(i32) 5 send_ptr(i64) 6 recv_ptr(i64) 7 send_dofs(i32) 8 recv_dofs(i32)
  const void *gfh_partition_array(const void *p, int which);

  // csrc/fe_basis.h for the tests: FE_Q(p) support points [p + 1]; basis values / derivatives
  // [n_points][p + 1]; hierarchical node -> lexicographic triple and FESystem local DoF ->
  // (node, component)
  void gfh_fe_support_points(int p, double *out);
  void gfh_fe_basis_eval(int p, int n_points, const double *x, double *values,
                         double *derivatives);
  int  gfh_fe_numbering(int dim, int p, int *lex, int *node_of, int *comp_of);
}
