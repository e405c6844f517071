// Host driver of the linear solver: control flow of ElastoDynamics::run()
// (linear_elasticity.cc:634-716); numerical members forwarded to the device library.
#include "linear_elasticity.h"

#include <cmath>
#include <iostream>
#include <stdexcept>

namespace Linear_Elasticity
{
  using Adapter::gf_check;

  template <int dim>
  ElastoDynamics<dim>::ElastoDynamics(const std::string &parameter_file)
    : parameters(parameter_file)
    , interface_boundary_id(6)
    , clamped_mesh_id(0)
    , out_of_plane_clamped_mesh_id(4)
    , timer(std::cout)
    , time(parameters.end_time, parameters.delta_t)
    , adapter(parameters, interface_boundary_id)
  {}

  template <int dim>
  ElastoDynamics<dim>::~ElastoDynamics()
  {}

  template <int dim>
  void ElastoDynamics<dim>::make_grid()
  {
    const std::string error_message("The interface_id cannot be the same as the clamped one");
    if (clamped_mesh_id == interface_boundary_id ||
        out_of_plane_clamped_mesh_id == interface_boundary_id) // :163-166
      throw std::runtime_error(error_message);
    if (interface_boundary_id != adapter.deal_boundary_interface_id)
      throw std::runtime_error("Wrong interface ID in the Adapter specified");
    host.make_grid(parameters, dim, adapter.precice.mesh_repetitions, gfh::numbering_cellwise);
  }

  template <int dim>
  void ElastoDynamics<dim>::setup_system()
  {
    std::cout << "Triangulation:"
              << "\n\t Number of active cells: " << host.mesh->n_cells
              << "\n\t Polynomial degree: " << parameters.poly_degree
              << "\n\t Number of degrees of freedom: " << host.mesh->n_dofs << std::endl;
    host.create_device(parameters, dim, GF_MODEL_LINEAR);
    const int n_mg_levels = host.create_multigrid(parameters, dim, GF_MODEL_LINEAR);
    std::cout << "\t CG preconditioner: "
              << (n_mg_levels > 1 ? "geometric multigrid, " + std::to_string(n_mg_levels) + " levels" :
                                    std::string("block-Jacobi"))
              << std::endl;
    if (parameters.type_lin == "Direct")
      {
        // SparseDirectUMFPACK's replacement: band Cholesky on the device where it fits
        int64_t   half_bandwidth = 0;
        const int rc = gf_direct_info(host.handle, nullptr, &half_bandwidth, nullptr);
        std::cout << "\t Direct solver: "
                  << (rc == GF_OK ? "band Cholesky on the device, half bandwidth " +
                                      std::to_string(half_bandwidth) :
                                    std::string("CG to 1e-13 (") + gf_last_error(host.handle) + ")")
                  << std::endl;
      }
    auto vec         = [&](int id) { return VectorType{host.handle, id}; };
    old_velocity     = vec(GF_LIN_OLD_VELOCITY);
    velocity         = vec(GF_LIN_VELOCITY);
    old_displacement = vec(GF_LIN_OLD_DISPLACEMENT);
    displacement     = vec(GF_LIN_DISPLACEMENT);
    old_stress       = vec(GF_LIN_OLD_STRESS);
    stress           = vec(GF_LIN_STRESS);
    system_rhs       = vec(GF_LIN_SYSTEM_RHS);
    state_variables  = {&old_velocity, &velocity, &old_displacement, &displacement,
                       &old_stress}; // :238-239
  }

  template <int dim>
  void ElastoDynamics<dim>::assemble_system()
  {
    gf_check(host.handle, gf_lin_assemble_once(host.handle)); // :248-374
  }

  template <int dim>
  void ElastoDynamics<dim>::assemble_consistent_loading()
  {} // face kernel inside gf_lin_step (:458-521)

  template <int dim>
  void ElastoDynamics<dim>::assemble_rhs()
  {
    timer.enter_subsection("Assemble rhs"); // fused into gf_lin_step (:378-454)
    timer.leave_subsection("Assemble rhs");
  }

  template <int dim>
  void ElastoDynamics<dim>::update_displacement()
  {} // fused into gf_lin_step (:579-586)

  template <int dim>
  void ElastoDynamics<dim>::solve()
  {
    uint32_t lin_it  = 1;
    double   lin_res = 0.0;
    if (parameters.type_lin == "CG")
      std::cout << "\t CG solver: " << std::endl;
    else if (parameters.type_lin == "Direct")
      std::cout << "\t Direct solver: " << std::endl;
    else
      throw std::runtime_error("Linear solver type not implemented");
    // assemble_rhs + solve + update_displacement (:680-686) are one device call; it is booked
    // under the reference's "Solve system" section (:529), "Assemble rhs" (:382) stays empty
    timer.enter_subsection("Solve system");
    gf_check(host.handle, gf_lin_step(host.handle, parameters.type_lin == "CG" ? 0 : 1,
                                      parameters.max_iterations_lin, &lin_it, &lin_res));
    timer.leave_subsection("Solve system");
    std::cout << "\t     No of iterations:\t" << lin_it << "\n \t     Final residual:\t" << lin_res
              << std::endl;
  }

  template <int dim>
  void ElastoDynamics<dim>::output_results() const
  {
    // DataOut + Postprocessor on the displaced grid (:590-629), file index as in :616-619
    timer.enter_subsection("Output results");
    host.output_results(GF_LIN_DISPLACEMENT, parameters.output_folder,
                        time.get_timestep() / parameters.output_interval);
    timer.leave_subsection("Output results");
  }

  template <int dim>
  void ElastoDynamics<dim>::run()
  {
    make_grid();
    setup_system();
    output_results();
    assemble_system();
    adapter.initialize(host.interface, displacement);
    while (adapter.precice.isCouplingOngoing())
      {
        adapter.save_current_state_if_required(state_variables, time);
        time.increment();
        std::cout << std::endl
                  << "Timestep " << time.get_timestep() << " @ " << std::fixed << time.current()
                  << "s" << std::endl;
        if (!(std::abs(time.get_delta_t() - adapter.precice.getMaxTimeStepSize()) < 1e-10))
          throw std::runtime_error(
            "This solver supports only constant time-step sizes."
            "Configured time step size in deal.II parameter file: " +
            std::to_string(time.get_delta_t()) + ". Time-window size from preCICE: " +
            std::to_string(adapter.precice.getMaxTimeStepSize()) + ".");
        adapter.read_data(time.get_delta_t(), stress);
        assemble_rhs();
        solve();
        update_displacement();
        timer.enter_subsection("Advance adapter"); // :696-698
        adapter.advance(displacement, time.get_delta_t());
        timer.leave_subsection("Advance adapter");
        adapter.reload_old_state_if_required(state_variables, time);
        if (adapter.precice.isTimeWindowComplete() &&
            time.get_timestep() % parameters.output_interval == 0)
          output_results();
      }
    adapter.precice.finalize();
  }

  template class ElastoDynamics<2>;
  template class ElastoDynamics<3>;
} // namespace Linear_Elasticity
