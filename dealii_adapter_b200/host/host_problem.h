// Host-side set-up shared by the two solver classes: the part of make_grid / system_setup /
// make_constraints that stays on the host (nonlinear_elasticity.cc:171-380,1094-1150;
// linear_elasticity.cc:79-244,431-446) and fills the gf_desc handed to the device library.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "adapter/adapter.h"
#include "adapter/parameters.h"
#include "graft_fem.h"
#include "structured_mesh.h"

namespace gfh
{
  struct HostProblem
  {
    std::unique_ptr<StructuredMesh> mesh;
    std::vector<uint8_t>            constrained;
    std::vector<int32_t>            iface_cell, iface_face_no, iface_dofs;
    Adapter::InterfaceDescription   interface;
    gf_handle                       handle = nullptr;
    double                          vol_reference = 0;

    // scenario geometry + boundary roles; `reps_override` replaces the hard-coded repetitions
    void make_grid(const Parameters::AllParameters &prm, int dim,
                   const std::vector<int> &reps_override, int numbering);
    void create_device(const Parameters::AllParameters &prm, int dim, int model);
    // Geometric multigrid preconditioner: the coarser levels of the refinement hierarchy (every
    // repetition halved while all are even; in the reference: the levels of refine_global,
    // nonlinear_elasticity.cc:245-246) become handles linked below `handle` with gf_mg_attach.
    // Environment GF_PRECONDITIONER = multigrid (default when a coarser level exists) |
    // block-jacobi. Returns the number of levels in use.
    int create_multigrid(const Parameters::AllParameters &prm, int dim, int model);
    std::vector<std::unique_ptr<HostProblem>> coarse_levels;
    // output_results (nonlinear_elasticity.cc:1215-1254, linear_elasticity.cc:590-629): patch
    // fields from the device (gf_postprocess), points + VTK file on the host (vtk_output.h).
    // Writes <folder>/solution-<index, 3 digits>.vtk and prints the reference's message.
    void output_results(int which_vector, const std::string &folder, unsigned index) const;
    ~HostProblem();
  };
} // namespace gfh
