// Linear_Elasticity::ElastoDynamics<dim> — drop-in mirror of the reference class (public surface:
// ctor(parameter_file), run(); include/linear_elasticity.h:58-64); hot members forward to the
// device library.
#pragma once
#include <string>
#include <vector>

#include "adapter/adapter.h"
#include "adapter/parameters.h"
#include "adapter/time_handler.h"
#include "adapter/timer_output.h"
#include "host_problem.h"

namespace Linear_Elasticity
{
  template <int dim>
  class ElastoDynamics
  {
  public:
    ElastoDynamics(const std::string &parameter_file);
    ~ElastoDynamics();
    void run();

  private:
    using VectorType = Adapter::DeviceVector;
    void make_grid();
    void setup_system();
    void assemble_system();
    void assemble_rhs();
    void assemble_consistent_loading();
    void solve();
    void update_displacement();
    void output_results() const;

    Parameters::AllParameters parameters;
    const unsigned int        interface_boundary_id;
    unsigned int              clamped_mesh_id, out_of_plane_clamped_mesh_id;
    gfh::HostProblem          host;
    mutable Adapter::TimerOutput timer; // TimerOutput(std::cout, summary, wall_times) :63
    Adapter::Time             time;
    Adapter::Adapter<dim, VectorType, Parameters::AllParameters> adapter;
    VectorType old_velocity, velocity, old_displacement, displacement, old_stress, stress,
      system_rhs;
    std::vector<VectorType *> state_variables;
  };
} // namespace Linear_Elasticity
