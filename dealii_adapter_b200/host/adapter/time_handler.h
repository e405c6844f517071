// Adapter::Time of the host mirror: step counter and current time with the accessors the solver
// classes and the Adapter use (reference: include/adapter/time_handler.h:21-84). It never leaves
// the host; the C-ABI receives delta_t by value. tests/test_reference_pins.py compares it with the
// reference's own class on increment / reset sequences.
#pragma once
#include <cmath>

namespace Adapter
{
  class Time
  {
    unsigned int n_steps = 0;
    double       now     = 0.0;
    const double t_end, dt;

  public:
    Time(const double time_end, const double delta_t)
      : t_end(time_end)
      , dt(delta_t)
    {}
    virtual ~Time() = default;

    unsigned int get_timestep() const { return n_steps; }
    double       current() const { return now; }
    double       end() const { return t_end; }
    double       get_delta_t() const { return dt; }

    void increment()
    {
      ++n_steps;
      now += dt;
    }
    // checkpoint reload: jump to an absolute time; the step counter follows from it (the ratio is
    // rounded at the 10th decimal first, as the reference does, then truncated)
    void set_absolute_time(const double new_time)
    {
      const double scale = std::pow(10, 10);
      n_steps            = static_cast<unsigned int>(std::round((new_time / dt) * scale) / scale);
      now                = new_time;
    }
  };
} // namespace Adapter
