// Driver: same command line and dispatch as the reference's elasticity.cc:7-129 — argv[1] is the
// parameter file (default parameters.prm), "Solver/Model" selects the class, exceptions are printed
// and turn into exit code 1. DIM is a compile-time choice (-DDIM, CMakeLists.txt:15-18).
#include <sys/stat.h>

#include <cerrno>
#include <iostream>
#include <string>

#include "adapter/parameters.h"
#include "linear_elasticity.h"
#include "nonlinear_elasticity.h"

#ifndef DIM
#  define DIM 2
#endif

int main(int argc, char **argv)
{
  try
    {
      std::cout << "--------------------------------------------------\n"
                << "             Running deal.ii solver\n"
                << "   B200 device path (libgraftfem), dimension " << DIM << "\n"
                << "--------------------------------------------------\n"
                << std::endl;
      std::string parameter_file = argc > 1 ? argv[1] : "parameters.prm";
      const Parameters::AllParameters prm(parameter_file);
      // "Output folder" is created component by component; a component that can neither be
      // created nor exists ends the run with the reference's message (elasticity.cc:57-81)
      if (!prm.output_folder.empty())
        {
          std::string path = prm.output_folder;
          if (path.back() != '/')
            path += '/';
          for (size_t pos = path.find('/'); pos != std::string::npos; pos = path.find('/', pos + 1))
            {
              const std::string sub = path.substr(0, pos);
              if (sub.empty())
                continue; // leading '/'
              if (mkdir(sub.c_str(), S_IRWXU | S_IRGRP | S_IXGRP | S_IROTH | S_IXOTH) != 0 &&
                  errno != EEXIST)
                throw std::runtime_error("Can't create: " + path);
            }
        }
      if (prm.model == "neo-Hookean")
        {
          Nonlinear_Elasticity::Solid<DIM> solid(parameter_file);
          solid.run();
        }
      else if (prm.model == "linear")
        {
          Linear_Elasticity::ElastoDynamics<DIM> elastic_solver(parameter_file);
          elastic_solver.run();
        }
      else
        throw std::runtime_error("Unknown model '" + prm.model + "'");
    }
  catch (std::exception &exc)
    {
      std::cerr << std::endl
                << "----------------------------------------------------" << std::endl
                << "Exception on processing: " << std::endl
                << exc.what() << std::endl
                << "Aborting!" << std::endl
                << "----------------------------------------------------" << std::endl;
      return 1;
    }
  return 0;
}
