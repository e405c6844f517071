// Host driver of the nonlinear solver: control flow of the reference's Solid::run() and
// solve_nonlinear_timestep() (nonlinear_elasticity.cc:99-167, :410-499) with the numerical members
// forwarded to the device library.
#include "nonlinear_elasticity.h"

#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>

namespace Nonlinear_Elasticity
{
  using Adapter::gf_check;

  template <int dim, typename NumberType>
  Solid<dim, NumberType>::Solid(const std::string &parameter_file)
    : parameters(Parameters::AllParameters(parameter_file))
    , vol_reference(0.0)
    , vol_current(0.0)
    , boundary_interface_id(7)
    , timer(std::cout)
    , time(parameters.end_time, parameters.delta_t)
    , adapter(parameters, boundary_interface_id)
  {
    if (!parameters.data_consistent) // :83-87
      throw std::runtime_error(
        "The neo-Hookean solid doesn't support 'Force' data reading. Please switch to 'Stress' "
        "data on the Fluid side or use the linear model of the solid solver");
  }

  template <int dim, typename NumberType>
  Solid<dim, NumberType>::~Solid()
  {}

  template <int dim, typename NumberType>
  void Solid<dim, NumberType>::make_grid()
  {
    if (!((dim == 2 && parameters.body_force[2] == 0) || dim == 3)) // :177-180
      throw std::runtime_error(
        "Setting body forces in z-direction for a two dimensional simulation has no effect");
    // the reference hard-codes repetitions and refinement (:192-194,:209-211,:245); larger
    // configurations are a host-side choice carried by the participant config
    host.make_grid(parameters, dim, adapter.precice.mesh_repetitions, gfh::numbering_component_wise);
    if (boundary_interface_id != adapter.deal_boundary_interface_id) // :295-296
      throw std::runtime_error("Wrong interface ID in the Adapter.");
    vol_reference = host.vol_reference;
    vol_current   = vol_reference;
    std::cout << "Grid:\n\t Reference volume: " << vol_reference << std::endl;
  }

  template <int dim, typename NumberType>
  void Solid<dim, NumberType>::system_setup()
  {
    timer.enter_subsection("Setup system"); // :309
    std::cout << "Triangulation:"
              << "\n\t Number of active cells: " << host.mesh->n_cells
              << "\n\t Polynomial degree: " << parameters.poly_degree
              << "\n\t Number of degrees of freedom: " << host.mesh->n_dofs << std::endl;
    host.create_device(parameters, dim, GF_MODEL_NEO_HOOKEAN);
    const int n_mg_levels = host.create_multigrid(parameters, dim, GF_MODEL_NEO_HOOKEAN);
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
    auto vec = [&](int id) { return VectorType{host.handle, id}; };
    total_displacement     = vec(GF_NL_TOTAL_DISPLACEMENT);
    total_displacement_old = vec(GF_NL_TOTAL_DISPLACEMENT_OLD);
    velocity               = vec(GF_NL_VELOCITY);
    velocity_old           = vec(GF_NL_VELOCITY_OLD);
    acceleration           = vec(GF_NL_ACCELERATION);
    acceleration_old       = vec(GF_NL_ACCELERATION_OLD);
    external_stress        = vec(GF_NL_EXTERNAL_STRESS);
    system_rhs             = vec(GF_NL_SYSTEM_RHS);
    state_variables = {&total_displacement, &total_displacement_old, &velocity, &velocity_old,
                       &acceleration,       &acceleration_old}; // :370-375
    timer.leave_subsection();                                   // :379
  }

  template <int dim, typename NumberType>
  void Solid<dim, NumberType>::make_constraints(const int &)
  {
    // homogeneous Dirichlet set is static over the run (:1100-1149); it was handed to the device
    // as a mask in system_setup and is applied inside the scatter kernel
    std::cout << " CST " << std::flush;
  }

  template <int dim, typename NumberType>
  void Solid<dim, NumberType>::update_acceleration(VectorType &)
  {} // fused into gf_nl_newton_assemble (:444) and gf_nl_end_step (:142)

  template <int dim, typename NumberType>
  void Solid<dim, NumberType>::update_velocity(VectorType &)
  {} // gf_nl_end_step (:143)

  template <int dim, typename NumberType>
  void Solid<dim, NumberType>::update_old_variables()
  {} // gf_nl_end_step (:144)

  template <int dim, typename NumberType>
  void Solid<dim, NumberType>::assemble_system()
  {
    timer.enter_subsection("Assemble linear system"); // :1051
    std::cout << " ASM " << std::flush;
    gf_check(host.handle, gf_nl_newton_assemble(host.handle, &error_residual.u)); // :446-449
    timer.leave_subsection();                                                     // :1086
  }

  template <int dim, typename NumberType>
  std::pair<unsigned int, double> Solid<dim, NumberType>::solve_linear_system(VectorType &)
  {
    timer.enter_subsection("Linear solver"); // :1165
    std::cout << " SLV " << std::flush;
    uint32_t lin_it  = 0;
    double   lin_res = 0.0;
    if (parameters.type_lin != "CG" && parameters.type_lin != "Direct")
      throw std::runtime_error("Linear solver type not implemented");
    gf_check(host.handle,
             gf_nl_newton_solve(host.handle, parameters.type_lin == "CG" ? 0 : 1, parameters.tol_lin,
                                parameters.max_iterations_lin, &lin_it, &lin_res,
                                &last_update_norm)); // :1153-1211, :476, :487
    timer.leave_subsection();                        // :1205
    return std::make_pair(lin_it, lin_res);
  }

  template <int dim, typename NumberType>
  void Solid<dim, NumberType>::solve_nonlinear_timestep(VectorType &solution_delta)
  {
    std::cout << std::endl
              << "Timestep " << time.get_timestep() << " @ " << std::fixed << time.current() << "s"
              << std::endl;
    VectorType newton_update{host.handle, GF_NL_NEWTON_UPDATE};
    error_residual.reset();
    error_residual_0.reset();
    error_residual_norm.reset();
    error_update.reset();
    error_update_0.reset();
    error_update_norm.reset();
    print_conv_header();
    unsigned int newton_iteration = 0;
    for (; newton_iteration < parameters.max_iterations_NR; ++newton_iteration)
      {
        std::cout << " " << std::setw(2) << newton_iteration << " " << std::flush;
        make_constraints(newton_iteration);
        update_acceleration(solution_delta);
        assemble_system();
        if (newton_iteration == 0)
          error_residual_0 = error_residual;
        error_residual_norm = error_residual;
        error_residual_norm.normalise(error_residual_0);
        if (newton_iteration > 0 &&
            ((error_update_norm.u <= parameters.tol_u || error_update.u <= 1e-15) &&
             (error_residual_norm.u <= parameters.tol_f || error_residual.u <= 5e-9)))
          {
            std::cout << " CONVERGED! " << std::endl;
            print_conv_footer();
            break;
          }
        const std::pair<unsigned int, double> lin_solver_output = solve_linear_system(newton_update);
        error_update.u = last_update_norm;
        if (newton_iteration == 0)
          error_update_0 = error_update;
        error_update_norm = error_update;
        error_update_norm.normalise(error_update_0);
        std::cout << " | " << std::fixed << std::setprecision(3) << std::setw(7) << std::scientific
                  << lin_solver_output.first << "  " << lin_solver_output.second << "  "
                  << error_residual_norm.u << "  " << error_residual.u << "  "
                  << "  " << error_update_norm.u << "  " << error_update.u << "  " << std::endl;
      }
    if (!(newton_iteration < parameters.max_iterations_NR))
      throw std::runtime_error("No convergence in nonlinear solver!");
    newton_counts.push_back(newton_iteration);
  }

  template <int dim, typename NumberType>
  void Solid<dim, NumberType>::print_conv_header()
  {
    static const unsigned int l_width = 87;
    std::cout << std::string(l_width, '_') << std::endl;
    std::cout << "    SOLVER STEP    "
              << " |  LIN_IT   LIN_RES    RES_NORM   "
              << "RES_ABS      U_NORM    "
              << " U_ABS " << std::endl;
    std::cout << std::string(l_width, '_') << std::endl;
  }

  template <int dim, typename NumberType>
  void Solid<dim, NumberType>::print_conv_footer()
  {
    error_residual.normalise(error_residual_0);
    error_update.normalise(error_update_0);
    static const unsigned int l_width = 87;
    std::cout << std::string(l_width, '_') << std::endl;
    std::cout << "Relative errors:" << std::endl
              << "Displacement:\t" << error_update.u << std::endl
              << "Residual: \t" << error_residual.u << std::endl
              << "v / V_0:\t" << vol_current << " / " << vol_reference << std::endl;
  }

  template <int dim, typename NumberType>
  void Solid<dim, NumberType>::output_results() const
  {
    // DataOut + Postprocessor on the displaced grid (:1215-1254): patch fields on the device, file
    // on the host; file index as in :1240-1243
    timer.enter_subsection("Output results");
    host.output_results(GF_NL_TOTAL_DISPLACEMENT, parameters.output_folder,
                        time.get_timestep() / parameters.output_interval);
    timer.leave_subsection("Output results");
  }

  template <int dim, typename NumberType>
  void Solid<dim, NumberType>::run()
  {
    make_grid();
    system_setup();
    output_results();
    adapter.initialize(host.interface, total_displacement);
    VectorType solution_delta{host.handle, GF_NL_SOLUTION_DELTA};
    while (adapter.precice.isCouplingOngoing())
      {
        adapter.save_current_state_if_required(state_variables, time);
        gf_check(host.handle, gf_nl_begin_step(host.handle)); // solution_delta = 0.0 (:121)
        time.increment();
        if (!(std::abs(time.get_delta_t() - adapter.precice.getMaxTimeStepSize()) < 1e-10))
          throw std::runtime_error(
            "This solver supports only constant time-step sizes."
            "Configured time step size in deal.II parameter file: " +
            std::to_string(time.get_delta_t()) + ". Time-window size from preCICE: " +
            std::to_string(adapter.precice.getMaxTimeStepSize()) + ".");
        adapter.read_data(time.get_delta_t(), external_stress);
        solve_nonlinear_timestep(solution_delta);
        // total_displacement += solution_delta and the Newmark updates (:139-144)
        gf_check(host.handle, gf_nl_end_step(host.handle));
        timer.enter_subsection("Advance adapter"); // :149-154
        adapter.advance(total_displacement, time.get_delta_t());
        timer.leave_subsection("Advance adapter");
        adapter.reload_old_state_if_required(state_variables, time);
        if (adapter.precice.isTimeWindowComplete() &&
            time.get_timestep() % parameters.output_interval == 0)
          output_results();
      }
    adapter.precice.finalize();
  }

  template class Solid<2, double>;
  template class Solid<3, double>;
} // namespace Nonlinear_Elasticity
