// TimerOutput(std::cout, summary, wall_times) as the reference's solver classes use it
// (nonlinear_elasticity.cc:79,149-154,309,379,1051,1086,1165,1205,1219,1253;
// linear_elasticity.cc:63,382,453,529,574,594,629,696-698): named sections accumulate wall time
// and a call count, the summary table is printed when the object is destroyed. The sections keep
// the reference's names; here they time the (synchronous) C-ABI calls, i.e. the device work.
// Table layout restated from deal.II 9.5's TimerOutput::print_summary.
#pragma once
#include <chrono>
#include <cstdio>
#include <list>
#include <map>
#include <ostream>
#include <string>

namespace Adapter
{
  class TimerOutput
  {
  public:
    explicit TimerOutput(std::ostream &stream)
      : out(stream)
      , start(clock::now())
    {}
    ~TimerOutput() { print_summary(); }

    void enter_subsection(const std::string &name)
    {
      if (!sections.count(name))
        order.push_back(name);
      sections[name].since = clock::now();
      sections[name].n_calls++;
      active.push_back(name);
    }
    // without a name: the section entered last (TimerOutput::leave_subsection("") semantics)
    void leave_subsection(const std::string &name = "")
    {
      const std::string which = name.empty() ? (active.empty() ? std::string() : active.back()) : name;
      auto              it    = sections.find(which);
      if (it == sections.end())
        return;
      it->second.total += std::chrono::duration<double>(clock::now() - it->second.since).count();
      active.remove(which);
    }
    void print_summary() const
    {
      const double total = std::chrono::duration<double>(clock::now() - start).count();
      char         line[160];
      out << "\n\n+---------------------------------------------+------------+------------+\n";
      std::snprintf(line, sizeof line, "| Total wallclock time elapsed since start    | %9.3gs |            |\n",
                    total);
      out << line;
      out << "|                                             |            |            |\n"
          << "| Section                         | no. calls |  wall time | % of total |\n"
          << "+---------------------------------+-----------+------------+------------+\n";
      for (const auto &name : order)
        {
          const Section &s = sections.at(name);
          std::snprintf(line, sizeof line, "| %-31.31s | %9u | %9.3gs | %9.2g%% |\n", name.c_str(),
                        s.n_calls, s.total, total > 0 ? 100.0 * s.total / total : 0.0);
          out << line;
        }
      out << "+---------------------------------+-----------+------------+------------+\n" << std::endl;
    }

  private:
    using clock = std::chrono::steady_clock;
    struct Section
    {
      double             total   = 0;
      unsigned           n_calls = 0;
      clock::time_point since;
    };
    std::ostream &                 out;
    clock::time_point              start;
    std::map<std::string, Section> sections;
    std::list<std::string>         order, active;
  };
} // namespace Adapter
