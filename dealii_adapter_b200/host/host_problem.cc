#include "host_problem.h"

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>

#include "vtk_output.h"

namespace gfh
{
  void HostProblem::make_grid(const Parameters::AllParameters &prm, int dim,
                              const std::vector<int> &reps_override, int numbering)
  {
    // nonlinear_elasticity.cc:189-226 / linear_elasticity.cc:94-131
    int      reps[3];
    double   p0[3], p1[3];
    unsigned clamped, interface_faces;
    if (prm.scenario == "FSI3")
      {
        const int    r[3] = {18, 3, 1};
        const double a[3] = {0.24899, 0.19, -0.005}, b[3] = {0.6, 0.21, 0.005};
        for (int d = 0; d < 3; ++d)
          {
            reps[d] = r[d];
            p0[d]   = a[d];
            p1[d]   = b[d];
          }
        clamped         = 1u << 0;                             // id_flap_short_bottom = 0
        interface_faces = (1u << 2) | (1u << 3) | (1u << 1);   // long bottom/top, short top
      }
    else
      {
        const int    r[3] = {3, 18, 1};
        const double a[3] = {prm.flap_location - 0.05, 0, 0}, b[3] = {prm.flap_location + 0.05, 1, 0.3};
        for (int d = 0; d < 3; ++d)
          {
            reps[d] = r[d];
            p0[d]   = a[d];
            p1[d]   = b[d];
          }
        clamped         = 1u << 2;
        interface_faces = (1u << 0) | (1u << 1) | (1u << 3);
      }
    for (size_t d = 0; d < reps_override.size() && d < 3; ++d)
      reps[d] = reps_override[d];
    mesh.reset(new StructuredMesh(dim, int(prm.poly_degree), reps, p0, p1, numbering));
    constrained.assign(mesh->n_dofs, 0);
    mesh->boundary_dof_mask(clamped, (1u << dim) - 1, constrained.data());
    if (dim == 3) // out-of-plane faces clamped in z only (:1126-1147 / linear :436-446)
      mesh->boundary_dof_mask((1u << 4) | (1u << 5), 1u << 2, constrained.data());
    mesh->boundary_faces(interface_faces, iface_cell, iface_face_no);
    interface.dim               = dim;
    interface.n_interface_nodes = int(mesh->interface_dofs(interface_faces, iface_dofs));
    interface.interface_nodes_positions.resize(size_t(interface.n_interface_nodes) * dim);
    for (int i = 0; i < interface.n_interface_nodes; ++i)
      for (int d = 0; d < dim; ++d)
        interface.interface_nodes_positions[i * dim + d] =
          mesh->support_points[int64_t(iface_dofs[i]) * dim + d]; // adapter.h:313-321
    vol_reference = 1;
    for (int d = 0; d < dim; ++d)
      vol_reference *= (p1[d] - p0[d]);
  }

  void HostProblem::create_device(const Parameters::AllParameters &prm, int dim, int model)
  {
    gf_desc d{};
    d.dim           = dim;
    d.degree        = int(prm.poly_degree);
    d.model         = model;
    d.n_dofs        = mesh->n_dofs;
    d.n_cells       = mesh->n_cells;
    d.cell_dofs     = mesh->cell_dofs.data();
    d.cell_vertices = mesh->cell_vertices.data();
    d.constrained   = constrained.data();
    d.n_iface_faces = int64_t(iface_cell.size());
    d.iface_cell    = iface_cell.data();
    d.iface_face_no = iface_face_no.data();
    d.n_iface_nodes = interface.n_interface_nodes;
    d.iface_dofs    = iface_dofs.data();
    d.mu            = prm.mu;
    d.nu            = prm.nu;
    d.rho           = prm.rho;
    for (int k = 0; k < 3; ++k)
      d.body_force[k] = prm.body_force[k];
    d.beta            = prm.beta;
    d.gamma           = prm.gamma;
    d.theta           = prm.theta;
    d.delta_t         = prm.delta_t;
    d.data_consistent = prm.data_consistent ? 1 : 0;
    d.device          = 0;
    d.n_owned_dofs    = mesh->n_dofs;
    const int rc      = gf_create(&d, &handle);
    if (rc != GF_OK)
      throw std::runtime_error(gf_last_error(nullptr));
  }

  int HostProblem::create_multigrid(const Parameters::AllParameters &prm, int dim, int model)
  {
    const char *env = std::getenv("GF_PRECONDITIONER");
    if (env && std::string(env) == "block-jacobi")
      return 1;
    // the device's transfer operators need nested support points: FE_Q(1), FE_Q(2). The shipped
    // parameter files (degree 3 / 4) run the block-Jacobi CG
    if (prm.poly_degree > 2)
      return 1;
    HostProblem *fine = this;
    while (true)
      {
        std::vector<int> reps(dim);
        bool             ok = true;
        for (int d = 0; d < dim; ++d)
          {
            ok      = ok && fine->mesh->reps[d] % 2 == 0;
            reps[d] = fine->mesh->reps[d] / 2;
          }
        if (!ok)
          break;
        std::unique_ptr<HostProblem> co(new HostProblem);
        co->make_grid(prm, dim, reps, fine->mesh->numbering);
        co->create_device(prm, dim, model);
        // cell->child(k), k = kx + 2 ky + 4 kz; cells are lexicographic
        const int            nc = 1 << dim;
        std::vector<int32_t> child(size_t(co->mesh->n_cells) * nc);
        const int            rz = dim == 3 ? reps[2] : 1;
        for (int k = 0; k < rz; ++k)
          for (int j = 0; j < reps[1]; ++j)
            for (int i = 0; i < reps[0]; ++i)
              for (int c = 0; c < nc; ++c)
                child[size_t(co->mesh->cell_index(i, j, k)) * nc + c] = int32_t(fine->mesh->cell_index(
                  2 * i + (c & 1), 2 * j + ((c >> 1) & 1), dim == 3 ? 2 * k + ((c >> 2) & 1) : 0));
        if (gf_mg_attach(fine->handle, co->handle, child.data()) != GF_OK)
          throw std::runtime_error(gf_last_error(fine->handle));
        coarse_levels.push_back(std::move(co));
        fine = coarse_levels.back().get();
      }
    if (!coarse_levels.empty() &&
        gf_set_option(handle, GF_OPT_PRECONDITIONER, GF_PRECOND_MULTIGRID) != GF_OK)
      throw std::runtime_error(gf_last_error(handle));
    return 1 + int(coarse_levels.size());
  }

  void HostProblem::output_results(int which_vector, const std::string &folder,
                                   unsigned index) const
  {
    const int dim = mesh->dim, p = mesh->degree, nf = dim + dim * dim;
    std::vector<double> fields(size_t(mesh->n_cells) * mesh->nodes_per_cell * nf);
    if (gf_postprocess(handle, which_vector, fields.data()) != GF_OK)
      throw std::runtime_error(gf_last_error(handle));
    // DataOut::curved_boundary: only cells at the boundary take their inner patch points from
    // the (Eulerian) mapping
    std::vector<uint8_t> at_boundary(size_t(mesh->n_cells), 0);
    const int            rz = dim == 3 ? mesh->reps[2] : 1;
    for (int k = 0; k < rz; ++k)
      for (int j = 0; j < mesh->reps[1]; ++j)
        for (int i = 0; i < mesh->reps[0]; ++i)
          at_boundary[size_t(mesh->cell_index(i, j, k))] =
            (i == 0 || i == mesh->reps[0] - 1 || j == 0 || j == mesh->reps[1] - 1 ||
             (dim == 3 && (k == 0 || k == rz - 1))) ?
              1 :
              0;
    std::vector<double> points;
    patch_points(dim, p, mesh->n_cells, mesh->cell_vertices.data(), fields.data(),
                 at_boundary.data(), points);
    char name[32];
    std::snprintf(name, sizeof name, "solution-%03u.vtk", index); // Utilities::int_to_string(n, 3)
    const std::string path = (folder.empty() ? std::string(".") : folder) + "/" + name;
    if (!write_vtk(path, dim, p, mesh->n_cells, points.data(), fields.data()))
      throw std::runtime_error("cannot write " + path);
    std::cout << "\t Output written to " << name << " \n" << std::endl;
  }

  HostProblem::~HostProblem()
  {
    if (handle)
      gf_destroy(handle);
  }
} // namespace gfh
