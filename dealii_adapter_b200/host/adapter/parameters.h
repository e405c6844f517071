// Parameters::AllParameters — same structs, field names, prm subsections/keys, defaults and
// derived values as the reference (include/adapter/parameters.h:17-111, parameters.cc:3-206).
// deal.II's ParameterHandler is replaced by a small strict parser of the same .prm syntax
// (subsection/set/end, '#' comments); unknown subsections or keys are errors, as with
// ParameterHandler::parse_input (parameters.cc:187).
#pragma once
#include <array>
#include <string>

namespace Parameters
{
  struct Time
  {
    double      end_time        = 1;
    double      delta_t         = 0.1;
    int         output_interval = 1;
    std::string output_folder   = "";
  };
  struct System
  {
    double                nu     = 0.3;
    double                mu     = 1538462;
    double                lambda = -1;
    double                rho    = 1000;
    std::array<double, 3> body_force{{0., 0., 0.}};
  };
  struct Solver
  {
    std::string  model              = "linear";
    std::string  type_lin           = "Direct";
    double       tol_lin            = 1e-6;
    double       max_iterations_lin = 1;
    unsigned int max_iterations_NR  = 10;
    double       tol_f              = 1e-9;
    double       tol_u              = 1e-6;
  };
  struct Discretization
  {
    unsigned int poly_degree = 3;
    double       theta       = 0.5;
    double       beta        = 0.25;
    double       gamma       = 0.5;
  };
  struct PreciceAdapterConfiguration
  {
    std::string scenario         = "FSI3";
    std::string config_file      = "precice-config.xml";
    std::string participant_name = "dealiisolver";
    std::string mesh_name        = "dealii-mesh";
    std::string read_data_name   = "Stress";
    std::string write_data_name  = "Displacement";
    double      flap_location    = 0.0;
    bool        data_consistent  = true;
  };
  struct AllParameters : public Solver,
                         public Discretization,
                         public System,
                         public Time,
                         public PreciceAdapterConfiguration
  {
    AllParameters(const std::string &input_file);
  };
} // namespace Parameters
