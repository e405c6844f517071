/* TEST INFRASTRUCTURE ONLY. The coupling loop of the reference — Solid::run()
 * (nonlinear_elasticity.cc:96-167) and ElastoDynamics::run() (linear_elasticity.cc:632-716), member
 * definitions cut out at build time — run against RECORDING stand-ins: every member the loop
 * calls only appends its name to an event list, and the preCICE participant is a scripted
 * serial-implicit / explicit coupling scheme (n windows, k sub-iterations per window). The output
 * is the order of events the reference's own loop produces. Time is the reference's time_handler.h.
 *   usage: ref_run_driver nl|lin <windows> <sub_iterations> <output_interval> <dt_solver> <dt_precice> */
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "assembly_shim.h"
#include "adapter/time_handler.h"

static std::vector<std::string> g_events;
static void                     ev(const std::string &s) { g_events.push_back(s); }

// preCICE itself is not part of the reference: a scripted coupling scheme (same rules as the
// host's scripted participant, host/fake_precice.h)
struct ParticipantScript
{
  int    windows = 1, sub = 1, window = 0, iteration = 0;
  bool   window_complete = false;
  double dt = 0.01;
  bool   isCouplingOngoing() const { return window < windows; }
  double getMaxTimeStepSize() const { return dt; }
  bool   isTimeWindowComplete() const { return window_complete; }
  bool   requiresWritingCheckpoint() const { return sub > 1 && iteration == 0; }
  bool   requiresReadingCheckpoint() const { return sub > 1 && !window_complete; }
  void   advance(double)
  {
    ++iteration;
    window_complete = iteration == sub;
    if (window_complete)
      {
        ++window;
        iteration = 0;
      }
  }
  void finalize() { ev("precice.finalize"); }
};

struct RecordingVector
{
  std::string      name;
  RecordingVector &operator=(double v)
  {
    ev(name + "=" + (v == 0.0 ? "0" : "x"));
    return *this;
  }
  RecordingVector &operator+=(const RecordingVector &o)
  {
    ev(name + "+=" + o.name);
    return *this;
  }
};
struct TimerShim
{
  void enter_subsection(const std::string &) {}
  void leave_subsection(const std::string & = "") {}
};
struct AdapterRecorder
{
  ParticipantScript precice;
  template <class D, class V>
  void initialize(const D &, const V &)
  {
    ev("adapter.initialize");
  }
  template <class S>
  void save_current_state_if_required(const S &, Adapter::Time &t)
  {
    const bool w = precice.requiresWritingCheckpoint();
    ev(std::string("adapter.save_current_state_if_required:") + (w ? "1" : "0"));
    if (w)
      saved_time = t.current(); // adapter.h:463
  }
  template <class S>
  void reload_old_state_if_required(S &, Adapter::Time &t)
  {
    const bool r = precice.requiresReadingCheckpoint();
    ev(std::string("adapter.reload_old_state_if_required:") + (r ? "1" : "0"));
    if (r)
      t.set_absolute_time(saved_time);
  }
  void read_data(double, RecordingVector &v) { ev("adapter.read_data:" + v.name); }
  void advance(const RecordingVector &v, double dt)
  {
    ev("adapter.advance:" + v.name);
    precice.advance(dt);
  }
  double saved_time = 0;
};

namespace Nonlinear_Elasticity
{
  using namespace dealii;
  // inside run() `BlockVector<NumberType> solution_delta(dofs_per_block)` is a local
  template <typename N>
  struct BlockVector : RecordingVector
  {
    BlockVector() = default;
    explicit BlockVector(int) { name = "solution_delta"; }
    using RecordingVector::operator=;
  };
  template <int dim, typename NumberType = double>
  class Solid
  {
  public:
    Parameters::AllParameters parameters;
    Adapter::Time             time;
    AdapterRecorder           adapter;
    mutable TimerShim         timer;
    int                       dof_handler_ref = 0, dofs_per_block = 0, state_variables = 0;
    explicit Solid(double dt)
      : time(1e9, dt)
    {}
    RecordingVector total_displacement{"total_displacement"}, external_stress{"external_stress"};
    void            make_grid() { ev("make_grid"); }
    void            system_setup() { ev("system_setup"); }
    void            output_results() const { ev("output_results@" + std::to_string(time.get_timestep())); }
    void            solve_nonlinear_timestep(RecordingVector &) { ev("solve_nonlinear_timestep"); }
    void            update_acceleration(RecordingVector &) { ev("update_acceleration"); }
    void            update_velocity(RecordingVector &) { ev("update_velocity"); }
    void            update_old_variables() { ev("update_old_variables"); }
    void            run();
  };
#include "nl_run_extract.inc"
} // namespace Nonlinear_Elasticity

namespace Linear_Elasticity
{
  using namespace dealii;
  template <int dim>
  class ElastoDynamics
  {
  public:
    Parameters::AllParameters parameters;
    Adapter::Time             time;
    AdapterRecorder           adapter;
    mutable TimerShim         timer;
    int                       dof_handler = 0, state_variables = 0;
    explicit ElastoDynamics(double dt)
      : time(1e9, dt)
    {}
    RecordingVector           displacement{"displacement"}, stress{"stress"};
    void                      make_grid() { ev("make_grid"); }
    void                      setup_system() { ev("setup_system"); }
    void                      assemble_system() { ev("assemble_system"); }
    void output_results() const { ev("output_results@" + std::to_string(time.get_timestep())); }
    void assemble_rhs() { ev("assemble_rhs"); }
    void solve() { ev("solve"); }
    void update_displacement() { ev("update_displacement"); }
    void run();
  };
#include "lin_run_extract.inc"
} // namespace Linear_Elasticity

int main(int argc, char **argv)
{
  if (argc < 7)
    return 2;
  const std::string  solver = argv[1];
  const int          windows = atoi(argv[2]), sub = atoi(argv[3]), interval = atoi(argv[4]);
  const double       dt_solver = atof(argv[5]), dt_precice = atof(argv[6]);
  std::ostringstream sink;
  auto *             old = std::cout.rdbuf(sink.rdbuf());
  int                rc  = 0;
  try
    {
      if (solver == "nl")
        {
          Nonlinear_Elasticity::Solid<3, double> s(dt_solver);
          s.parameters.output_interval = interval;
          s.adapter.precice            = ParticipantScript{windows, sub, 0, 0, false, dt_precice};
          s.run();
        }
      else
        {
          Linear_Elasticity::ElastoDynamics<3> s(dt_solver);
          s.parameters.output_interval = interval;
          s.adapter.precice            = ParticipantScript{windows, sub, 0, 0, false, dt_precice};
          s.run();
        }
    }
  catch (const std::exception &e)
    {
      ev(std::string("THROW:") + e.what());
      rc = 1;
    }
  std::cout.rdbuf(old);
  for (const auto &e : g_events)
    printf("%s\n", e.c_str());
  return rc;
}
