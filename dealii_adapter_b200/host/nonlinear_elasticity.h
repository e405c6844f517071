// Nonlinear_Elasticity::Solid<dim> — drop-in mirror of the reference class (public surface:
// ctor(parameter_file), run(); include/nonlinear_elasticity.h:135-140). Its private hot members
// keep their names; their bodies forward to the sm_100a library through include/graft_fem.h.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "adapter/adapter.h"
#include "adapter/parameters.h"
#include "adapter/time_handler.h"
#include "adapter/timer_output.h"
#include "host_problem.h"

namespace Nonlinear_Elasticity
{
  template <int dim, typename NumberType = double>
  class Solid
  {
  public:
    Solid(const std::string &parameter_file);
    virtual ~Solid();
    void run();

    // observers used by the tests (not in the reference)
    const std::vector<unsigned int> &newton_iterations_per_step() const { return newton_counts; }

  private:
    using VectorType = Adapter::DeviceVector;
    void make_grid();
    void system_setup();
    void assemble_system(); // fused with update_acceleration + get_error_residual on the device
    void make_constraints(const int &it_nr);
    void solve_nonlinear_timestep(VectorType &solution_delta);
    std::pair<unsigned int, double> solve_linear_system(VectorType &newton_update);
    void update_acceleration(VectorType &displacement_delta);
    void update_velocity(VectorType &displacement_delta);
    void update_old_variables();
    void output_results() const;
    static void print_conv_header();
    void        print_conv_footer();

    const Parameters::AllParameters parameters;
    double                          vol_reference, vol_current;
    gfh::HostProblem                host;
    const unsigned int              boundary_interface_id;
    mutable Adapter::TimerOutput    timer; // TimerOutput(std::cout, summary, wall_times) :79
    Adapter::Time                   time;
    Adapter::Adapter<dim, VectorType, Parameters::AllParameters> adapter;

    VectorType total_displacement, total_displacement_old, velocity, velocity_old, acceleration,
      acceleration_old, external_stress, system_rhs;
    std::vector<VectorType *> state_variables;

    struct Errors
    {
      Errors()
        : u(1.0)
      {}
      void reset() { u = 1.0; }
      void normalise(const Errors &val)
      {
        if (val.u != 0.0)
          u /= val.u;
      }
      double u;
    };
    Errors error_residual, error_residual_0, error_residual_norm, error_update, error_update_0,
      error_update_norm;
    double                    last_update_norm = 0;
    std::vector<unsigned int> newton_counts;
  };
} // namespace Nonlinear_Elasticity
