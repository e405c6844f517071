// Adapter::Time — host-side time bookkeeping with the interface of the reference
// (include/adapter/time_handler.h:21-84); stays on the host, values cross the C-ABI by value.
#pragma once
#include <cmath>

namespace Adapter
{
  class Time
  {
  public:
    Time(const double time_end, const double delta_t)
      : timestep(0)
      , time_current(0.0)
      , time_end(time_end)
      , delta_t(delta_t)
    {}
    virtual ~Time() {}
    double       current() const { return time_current; }
    double       end() const { return time_end; }
    double       get_delta_t() const { return delta_t; }
    unsigned int get_timestep() const { return timestep; }
    // used when a checkpoint is reloaded: the step counter is recomputed from the absolute time
    void set_absolute_time(const double new_time)
    {
      const double factor = std::pow(10, 10);
      timestep            = (unsigned int)(std::round((new_time / delta_t) * factor) / factor);
      time_current        = new_time;
    }
    void increment()
    {
      time_current += delta_t;
      ++timestep;
    }

  private:
    unsigned int timestep;
    double       time_current;
    const double time_end;
    const double delta_t;
  };
} // namespace Adapter
