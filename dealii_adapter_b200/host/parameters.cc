// See adapter/parameters.h. Key names, patterns and defaults follow parameters.cc:8-171.
#include "adapter/parameters.h"

#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>
#include <vector>
#include <stdexcept>

namespace Parameters
{
  namespace
  {
    std::string trim(const std::string &s)
    {
      const auto b = s.find_first_not_of(" \t\r\n");
      if (b == std::string::npos)
        return "";
      const auto e = s.find_last_not_of(" \t\r\n");
      return s.substr(b, e - b + 1);
    }
    // collapse runs of whitespace (ParameterHandler treats "Time  step" like "Time step")
    std::string squeeze(const std::string &s)
    {
      std::string out;
      bool        sp = false;
      for (char ch : trim(s))
        {
          if (ch == ' ' || ch == '\t')
            {
              sp = true;
              continue;
            }
          if (sp && !out.empty())
            out += ' ';
          sp = false;
          out += ch;
        }
      return out;
    }
    double to_double(const std::string &key, const std::string &v, double lo, double hi)
    {
      size_t pos = 0;
      double x   = 0;
      try
        {
          x = std::stod(v, &pos);
        }
      catch (...)
        {
          throw std::runtime_error("parameter '" + key + "': '" + v + "' is not a number");
        }
      if (trim(v.substr(pos)) != "" || x < lo || x > hi)
        throw std::runtime_error("parameter '" + key + "': value '" + v + "' does not match its pattern");
      return x;
    }
    void check_selection(const std::string &key, const std::string &v,
                         std::initializer_list<const char *> allowed)
    {
      for (const char *a : allowed)
        if (v == a)
          return;
      throw std::runtime_error("parameter '" + key + "': value '" + v + "' is not in its selection");
    }
  } // namespace

  AllParameters::AllParameters(const std::string &input_file)
  {
    // declared entries: subsection -> key (parameters.cc:10-171)
    const std::map<std::string, std::vector<std::string>> declared = {
      {"Time", {"End time", "Time step size", "Output interval", "Output folder"}},
      {"System properties", {"Shear modulus", "Poisson's ratio", "rho", "body forces"}},
      {"Solver",
       {"Model", "Solver type", "Residual", "Max iteration multiplier",
        "Max iterations Newton-Raphson", "Tolerance force", "Tolerance displacement"}},
      {"Discretization", {"Polynomial degree", "theta", "beta", "gamma"}},
      {"precice configuration",
       {"Scenario", "precice config-file", "Participant name", "Mesh name", "Read data name",
        "Write data name", "Flap location"}}};
    std::ifstream in(input_file);
    if (!in)
      throw std::runtime_error("cannot open parameter file '" + input_file + "'");
    std::map<std::string, std::string> values; // "subsection/key" -> value
    std::string                        line, subsection;
    int                                lineno = 0;
    while (std::getline(in, line))
      {
        ++lineno;
        const auto hash = line.find('#');
        if (hash != std::string::npos)
          line = line.substr(0, hash);
        line = trim(line);
        if (line.empty())
          continue;
        const std::string where = input_file + ":" + std::to_string(lineno) + ": ";
        if (line.rfind("subsection", 0) == 0)
          {
            if (!subsection.empty())
              throw std::runtime_error(where + "nested subsections are not declared");
            subsection = squeeze(line.substr(10));
            if (!declared.count(subsection))
              throw std::runtime_error(where + "no such subsection '" + subsection + "'");
          }
        else if (line == "end")
          {
            if (subsection.empty())
              throw std::runtime_error(where + "'end' without subsection");
            subsection.clear();
          }
        else if (line.rfind("set", 0) == 0)
          {
            const auto eq = line.find('=');
            if (eq == std::string::npos || subsection.empty())
              throw std::runtime_error(where + "malformed 'set' line");
            const std::string key = squeeze(line.substr(3, eq - 3));
            const auto &      keys = declared.at(subsection);
            if (std::find(keys.begin(), keys.end(), key) == keys.end())
              throw std::runtime_error(where + "no entry '" + key + "' in subsection '" +
                                       subsection + "'");
            values[subsection + "/" + key] = trim(line.substr(eq + 1));
          }
        else
          throw std::runtime_error(where + "cannot parse line '" + line + "'");
      }
    if (!subsection.empty())
      throw std::runtime_error(input_file + ": missing 'end'");
    auto has = [&](const char *k) { return values.count(k) != 0; };
    auto get = [&](const char *k) { return values.at(k); };
    const double big = 1e300;
    // Time (parameters.cc:8-29)
    if (has("Time/End time"))
      end_time = to_double("End time", get("Time/End time"), -big, big);
    if (has("Time/Time step size"))
      delta_t = to_double("Time step size", get("Time/Time step size"), 0., big);
    if (has("Time/Output interval"))
      output_interval = int(to_double("Output interval", get("Time/Output interval"), 0, 2147483647));
    if (has("Time/Output folder"))
      output_folder = get("Time/Output folder");
    // System properties (:33-55)
    if (has("System properties/Shear modulus"))
      mu = to_double("Shear modulus", get("System properties/Shear modulus"), 0., big);
    if (has("System properties/Poisson's ratio"))
      nu = to_double("Poisson's ratio", get("System properties/Poisson's ratio"), -1.0, 0.5);
    if (has("System properties/rho"))
      rho = to_double("rho", get("System properties/rho"), 0., big);
    if (has("System properties/body forces"))
      {
        std::stringstream ss(get("System properties/body forces"));
        std::string       item;
        int               n = 0;
        while (std::getline(ss, item, ','))
          {
            if (n >= 3)
              throw std::runtime_error("parameter 'body forces': expected 3 values");
            body_force[n++] = to_double("body forces", item, -big, big);
          }
        if (n != 3)
          throw std::runtime_error("parameter 'body forces': expected 3 values");
      }
    // Solver (:59-104)
    if (has("Solver/Model"))
      {
        model = get("Solver/Model");
        check_selection("Model", model, {"linear", "neo-Hookean"});
      }
    if (has("Solver/Solver type"))
      {
        type_lin = get("Solver/Solver type");
        check_selection("Solver type", type_lin, {"CG", "Direct"});
      }
    if (has("Solver/Residual"))
      tol_lin = to_double("Residual", get("Solver/Residual"), 0., big);
    if (has("Solver/Max iteration multiplier"))
      max_iterations_lin =
        to_double("Max iteration multiplier", get("Solver/Max iteration multiplier"), 0., big);
    if (has("Solver/Max iterations Newton-Raphson"))
      max_iterations_NR = (unsigned int)to_double(
        "Max iterations Newton-Raphson", get("Solver/Max iterations Newton-Raphson"), 0, 2147483647);
    if (has("Solver/Tolerance force"))
      tol_f = to_double("Tolerance force", get("Solver/Tolerance force"), 0., big);
    if (has("Solver/Tolerance displacement"))
      tol_u = to_double("Tolerance displacement", get("Solver/Tolerance displacement"), 0., big);
    // Discretization (:108-130)
    if (has("Discretization/Polynomial degree"))
      poly_degree = (unsigned int)to_double("Polynomial degree",
                                            get("Discretization/Polynomial degree"), 0, 2147483647);
    if (has("Discretization/theta"))
      theta = to_double("theta", get("Discretization/theta"), 0., 1.);
    if (has("Discretization/beta"))
      beta = to_double("beta", get("Discretization/beta"), 0., 0.5);
    if (has("Discretization/gamma"))
      gamma = to_double("gamma", get("Discretization/gamma"), 0., 1.);
    // precice configuration (:134-177)
    if (has("precice configuration/Scenario"))
      {
        scenario = get("precice configuration/Scenario");
        check_selection("Scenario", scenario, {"FSI3", "PF"});
      }
    if (has("precice configuration/precice config-file"))
      config_file = get("precice configuration/precice config-file");
    if (has("precice configuration/Participant name"))
      participant_name = get("precice configuration/Participant name");
    if (has("precice configuration/Mesh name"))
      mesh_name = get("precice configuration/Mesh name");
    if (has("precice configuration/Read data name"))
      read_data_name = get("precice configuration/Read data name");
    if (has("precice configuration/Write data name"))
      write_data_name = get("precice configuration/Write data name");
    if (has("precice configuration/Flap location"))
      flap_location = to_double("Flap location", get("precice configuration/Flap location"), -3, 3);

    // derived values (parameters.cc:189-200)
    lambda = 2 * mu * nu / (1 - 2 * nu);
    if (read_data_name.find("Stress") == 0)
      data_consistent = true;
    else if (read_data_name.find("Force") == 0)
      data_consistent = false;
    else
      throw std::runtime_error("Unknown read data type. Please use 'Force' or 'Stress' in the read "
                               "data naming.");
  }
} // namespace Parameters
