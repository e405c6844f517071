// precice::Participant stand-in with the 14 methods the adapter and the solver classes use
// (call sites adapter.h:217,236,324,329,334,341,354,377,384,455,476 and
// nonlinear_elasticity.cc:115,126,160,166). preCICE is a separate process/library in the reference
// and is OUT OF SCOPE; this scripted participant replays a coupling scheme so the drop-in host
// classes can run without it. With the real library present, the host classes compile against
// <precice/precice.hpp> unchanged (same signatures, precice::span replaced by std::vector here).
//
// "config file" (the prm key `precice config-file`): key = value lines
//   dimensions = 2|3            time-window-size = dt      max-time-windows = n
//   sub-iterations = k          (k > 1: implicit coupling, checkpoints requested)
//   traction = tx,ty[,tz]       ramp-time = t (linear ramp from 0)  relaxation = r0,r1,... per iteration
//   watch-point = x,y[,z]       watch-point-file = path    mesh-repetitions = nx,ny[,nz]
#pragma once
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace precice
{
  class Participant
  {
  public:
    Participant(const std::string &participant_name, const std::string &config_file, int rank,
                int size);
    int  getMeshDimensions(const std::string &mesh_name) const { return dimensions; }
    void setMeshVertices(const std::string &mesh_name, const std::vector<double> &positions,
                         std::vector<int> &ids);
    bool requiresInitialData() const { return false; }
    void writeData(const std::string &mesh_name, const std::string &data_name,
                   const std::vector<int> &ids, const std::vector<double> &values);
    void initialize() { initialized = true; }
    void readData(const std::string &mesh_name, const std::string &data_name,
                  const std::vector<int> &ids, double relative_read_time,
                  std::vector<double> &values) const;
    void   advance(double computed_time_step_size);
    bool   requiresWritingCheckpoint() const { return sub_iterations > 1 && iteration == 0; }
    bool   requiresReadingCheckpoint() const { return sub_iterations > 1 && !window_complete; }
    bool   isCouplingOngoing() const { return window < max_time_windows; }
    double getMaxTimeStepSize() const { return time_window_size; }
    bool   isTimeWindowComplete() const { return window_complete; }
    void   finalize();

    // host-side extras (not part of preCICE): scenario geometry overrides for the large configs
    std::vector<int> mesh_repetitions;

  private:
    int                 dimensions = 2, max_time_windows = 1, sub_iterations = 1;
    double              time_window_size = 0.1, ramp_time = 0.0;
    std::vector<double> traction, relaxation, watch_point, positions;
    std::string         watch_file;
    int                 window = 0, iteration = 0, watch_vertex = -1;
    bool                window_complete = false, initialized = false;
    std::ofstream       watch_out;
  };
} // namespace precice
