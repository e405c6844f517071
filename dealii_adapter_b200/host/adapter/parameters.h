// Parameters::AllParameters of the host mirror: one flat record holding every value the reference
// reads from the .prm file (include/adapter/parameters.h:17-111, parameters.cc:3-206) under the
// SAME member names, so `parameters.mu`, `parameters.type_lin`, ... in the solver classes read
// unchanged. The reference splits them over five base structs (Time, System, Solver,
// Discretization, PreciceAdapterConfiguration) because ParameterHandler declares them per
// subsection; here host/parameters.cc is a small strict parser of the same .prm syntax
// (subsection / set / end, '#' comments; undeclared subsections or keys and out-of-pattern values
// are errors, as with ParameterHandler::parse_input, parameters.cc:187), and
// tests/test_reference_pins.py checks field by field that it reads what the reference reads.
#pragma once
#include <array>
#include <string>

namespace Parameters
{
  struct AllParameters
  {
    explicit AllParameters(const std::string &input_file);

    // subsection "Solver"
    std::string  model              = "linear"; // linear | neo-Hookean
    std::string  type_lin           = "Direct"; // CG | Direct
    double       tol_lin            = 1e-6;     // "Residual", times ||rhs||
    double       max_iterations_lin = 1;        // multiples of the matrix size
    unsigned int max_iterations_NR  = 10;
    double       tol_f = 1e-9, tol_u = 1e-6;

    // subsection "Discretization"
    unsigned int poly_degree = 3;
    double       theta       = 0.5;             // one-step-theta (linear model)
    double       beta = 0.25, gamma = 0.5;      // Newmark (neo-Hookean model)

    // subsection "System properties"; lambda is derived: 2 mu nu / (1 - 2 nu)
    double                mu = 1538462, nu = 0.3, lambda = -1, rho = 1000;
    std::array<double, 3> body_force{{0., 0., 0.}};

    // subsection "Time"
    double      end_time = 1, delta_t = 0.1;
    int         output_interval = 1;
    std::string output_folder;

    // subsection "precice configuration"; data_consistent is derived from the read data name
    std::string scenario = "FSI3", config_file = "precice-config.xml";
    std::string participant_name = "dealiisolver", mesh_name = "dealii-mesh";
    std::string read_data_name = "Stress", write_data_name = "Displacement";
    double      flap_location   = 0.0;
    bool        data_consistent = true;
  };
} // namespace Parameters
