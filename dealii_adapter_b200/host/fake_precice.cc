#include "fake_precice.h"

#include <cmath>
#include <iomanip>

namespace precice
{
  namespace
  {
    std::string trim(const std::string &s)
    {
      const auto b = s.find_first_not_of(" \t\r\n");
      if (b == std::string::npos)
        return "";
      return s.substr(b, s.find_last_not_of(" \t\r\n") - b + 1);
    }
    std::vector<double> list(const std::string &v)
    {
      std::vector<double> out;
      std::stringstream   ss(v);
      std::string         item;
      while (std::getline(ss, item, ','))
        out.push_back(std::stod(item));
      return out;
    }
  } // namespace

  Participant::Participant(const std::string &, const std::string &config_file, int rank, int size)
  {
    if (rank != 0 || size != 1)
      throw std::runtime_error("fake precice: serial participant only (adapter.h:152-154)");
    std::ifstream in(config_file);
    if (!in)
      throw std::runtime_error("fake precice: cannot open config file '" + config_file + "'");
    std::string line;
    while (std::getline(in, line))
      {
        const auto hash = line.find('#');
        if (hash != std::string::npos)
          line = line.substr(0, hash);
        const auto eq = line.find('=');
        if (eq == std::string::npos)
          continue;
        const std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
        if (key == "dimensions")
          dimensions = std::stoi(val);
        else if (key == "time-window-size")
          time_window_size = std::stod(val);
        else if (key == "max-time-windows")
          max_time_windows = std::stoi(val);
        else if (key == "sub-iterations")
          sub_iterations = std::stoi(val);
        else if (key == "traction")
          traction = list(val);
        else if (key == "ramp-time")
          ramp_time = std::stod(val);
        else if (key == "relaxation")
          relaxation = list(val);
        else if (key == "watch-point")
          watch_point = list(val);
        else if (key == "watch-point-file")
          watch_file = val;
        else if (key == "mesh-repetitions")
          for (double r : list(val))
            mesh_repetitions.push_back(int(r));
        else
          throw std::runtime_error("fake precice: unknown key '" + key + "'");
      }
    traction.resize(dimensions, 0.0);
  }

  void Participant::setMeshVertices(const std::string &, const std::vector<double> &pos,
                                    std::vector<int> &ids)
  {
    positions = pos;
    const int n = int(pos.size()) / dimensions;
    ids.resize(n);
    for (int i = 0; i < n; ++i)
      ids[i] = i;
    if (!watch_point.empty())
      {
        double best = 1e300;
        for (int i = 0; i < n; ++i)
          {
            double d2 = 0;
            for (int d = 0; d < dimensions; ++d)
              d2 += std::pow(pos[i * dimensions + d] - watch_point[d], 2);
            if (d2 < best)
              {
                best         = d2;
                watch_vertex = i;
              }
          }
        if (!watch_file.empty())
          {
            watch_out.open(watch_file);
            watch_out << "# Time Iteration Coordinate... Displacement...\n";
          }
      }
  }

  void Participant::readData(const std::string &, const std::string &, const std::vector<int> &ids,
                             double relative_read_time, std::vector<double> &values) const
  {
    const double t     = window * time_window_size + relative_read_time;
    double       scale = (ramp_time > 0) ? std::min(1.0, t / ramp_time) : 1.0;
    if (!relaxation.empty())
      scale *= relaxation[std::min<size_t>(iteration, relaxation.size() - 1)];
    values.resize(ids.size() * dimensions);
    for (size_t i = 0; i < ids.size(); ++i)
      for (int d = 0; d < dimensions; ++d)
        values[i * dimensions + d] = scale * traction[d];
  }

  void Participant::writeData(const std::string &, const std::string &, const std::vector<int> &,
                              const std::vector<double> &values)
  {
    if (watch_vertex >= 0 && watch_out.is_open())
      {
        watch_out << std::setprecision(17) << (window + 1) * time_window_size << " " << iteration;
        for (int d = 0; d < dimensions; ++d)
          watch_out << " " << positions[watch_vertex * dimensions + d];
        for (int d = 0; d < dimensions; ++d)
          watch_out << " " << values[watch_vertex * dimensions + d];
        watch_out << "\n";
      }
  }

  void Participant::advance(double)
  {
    ++iteration;
    if (iteration >= sub_iterations)
      {
        iteration       = 0;
        window_complete = true;
        ++window;
      }
    else
      window_complete = false;
  }

  void Participant::finalize()
  {
    if (watch_out.is_open())
      watch_out.close();
  }
} // namespace precice
