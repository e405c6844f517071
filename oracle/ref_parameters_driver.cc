/* TEST INFRASTRUCTURE ONLY. Parameters::AllParameters of the reference — include/adapter/
 * parameters.h and parameters.cc compiled UNMODIFIED and in place against the ParameterHandler
 * stand-in of oracle/ref_shim — applied to a .prm file; prints every field as `key = value`
 * (or `ERROR: message` and exit code 1, the reference throws). */
#include <cstdio>
#include <iostream>

#include <adapter/parameters.h>

int main(int argc, char **argv)
{
  if (argc < 2)
    return 2;
  try
    {
      const Parameters::AllParameters p(argv[1]);
      printf("end_time = %.17g\ndelta_t = %.17g\noutput_interval = %d\noutput_folder = %s\n", p.end_time,
             p.delta_t, p.output_interval, p.output_folder.c_str());
      printf("nu = %.17g\nmu = %.17g\nlambda = %.17g\nrho = %.17g\nbody_force = %.17g,%.17g,%.17g\n", p.nu,
             p.mu, p.lambda, p.rho, p.body_force[0], p.body_force[1], p.body_force[2]);
      printf("model = %s\ntype_lin = %s\ntol_lin = %.17g\nmax_iterations_lin = %.17g\n"
             "max_iterations_NR = %u\ntol_f = %.17g\ntol_u = %.17g\n",
             p.model.c_str(), p.type_lin.c_str(), p.tol_lin, p.max_iterations_lin, p.max_iterations_NR,
             p.tol_f, p.tol_u);
      printf("poly_degree = %u\ntheta = %.17g\nbeta = %.17g\ngamma = %.17g\n", p.poly_degree, p.theta,
             p.beta, p.gamma);
      printf("scenario = %s\nconfig_file = %s\nparticipant_name = %s\nmesh_name = %s\n"
             "read_data_name = %s\nwrite_data_name = %s\nflap_location = %.17g\ndata_consistent = %d\n",
             p.scenario.c_str(), p.config_file.c_str(), p.participant_name.c_str(), p.mesh_name.c_str(),
             p.read_data_name.c_str(), p.write_data_name.c_str(), p.flap_location,
             p.data_consistent ? 1 : 0);
    }
  catch (const std::exception &e)
    {
      printf("ERROR: %s\n", e.what());
      return 1;
    }
  return 0;
}
