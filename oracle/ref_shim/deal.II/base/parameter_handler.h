/* TEST INFRASTRUCTURE ONLY — a small functional stand-in for dealii::ParameterHandler, enough for
 * the reference's include/adapter/parameters.{h,cc} to compile UNMODIFIED and in place:
 * add_parameter binds a variable, parse_input reads `subsection X` / `set Key = value` / `end`
 * files, values are checked against the declared pattern (deal.II 9.5 behaviour restated: an
 * undeclared entry or subsection and a value that does not match its pattern are errors). */
#ifndef PARAMETER_HANDLER_SHIM_H
#define PARAMETER_HANDLER_SHIM_H
#include <deal.II/base/dealii_min.h>

#include <cstdlib>
#include <fstream>
#include <functional>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef AssertThrow
#  define AssertThrow(cond, exc)             \
    do                                       \
      {                                      \
        if (!(cond))                         \
          throw std::runtime_error(exc);     \
      }                                      \
    while (0)
#endif

namespace dealii
{
  inline std::string ExcMessage(const std::string &m) { return m; }

  namespace Patterns
  {
    struct PatternBase
    {
      std::function<bool(const std::string &)> match;
      std::string                             description;
    };
    inline bool parse_double(const std::string &s, double &v)
    {
      char *end = nullptr;
      v         = std::strtod(s.c_str(), &end);
      while (end && *end == ' ')
        ++end;
      return end != s.c_str() && end && *end == '\0';
    }
    inline PatternBase Double(double lo = -std::numeric_limits<double>::max(),
                              double hi = std::numeric_limits<double>::max())
    {
      return {[lo, hi](const std::string &s) {
                double v;
                return parse_double(s, v) && v >= lo && v <= hi;
              },
              "[Double " + std::to_string(lo) + "..." + std::to_string(hi) + "]"};
    }
    inline PatternBase Integer(int lo = std::numeric_limits<int>::min(),
                               int hi = std::numeric_limits<int>::max())
    {
      return {[lo, hi](const std::string &s) {
                char *end = nullptr;
                long  v   = std::strtol(s.c_str(), &end, 10);
                return end != s.c_str() && *end == '\0' && v >= lo && v <= hi;
              },
              "[Integer]"};
    }
    inline PatternBase Anything()
    {
      return {[](const std::string &) { return true; }, "[Anything]"};
    }
    inline PatternBase Selection(const std::string &options)
    {
      return {[options](const std::string &s) {
                std::stringstream ss(options);
                std::string       o;
                while (std::getline(ss, o, '|'))
                  if (o == s)
                    return true;
                return false;
              },
              "[Selection " + options + "]"};
    }
    inline PatternBase List(const PatternBase &item)
    {
      return {[item](const std::string &s) {
                std::stringstream ss(s);
                std::string       o;
                while (std::getline(ss, o, ','))
                  {
                    const auto a = o.find_first_not_of(' '), b = o.find_last_not_of(' ');
                    if (a == std::string::npos || !item.match(o.substr(a, b - a + 1)))
                      return false;
                  }
                return true;
              },
              "[List]"};
    }
  } // namespace Patterns

  class ParameterHandler
  {
  public:
    void enter_subsection(const std::string &s) { path.push_back(s); }
    void leave_subsection() { path.pop_back(); }

    template <typename T>
    void add_parameter(const std::string &name, T &var, const std::string & /*doc*/,
                       const Patterns::PatternBase &pattern)
    {
      Entry e;
      e.pattern = pattern;
      e.assign  = [&var](const std::string &s) { convert(s, var); };
      entries[key(path, name)] = e;
      subsections_declared(path);
    }

    void parse_input(const std::string &filename)
    {
      std::ifstream in(filename);
      AssertThrow(bool(in), "ParameterHandler: cannot open " + filename);
      std::vector<std::string> cur;
      std::string              line;
      int                      lineno = 0;
      while (std::getline(in, line))
        {
          ++lineno;
          const auto hash = line.find('#');
          if (hash != std::string::npos)
            line = line.substr(0, hash);
          line = trim(line);
          if (line.empty())
            continue;
          const std::string where = " (line " + std::to_string(lineno) + " of " + filename + ")";
          if (line.rfind("subsection", 0) == 0)
            {
              cur.push_back(trim(line.substr(10)));
              AssertThrow(known_subsections.count(join(cur)),
                          "There is no such subsection to be entered: " + join(cur) + where);
            }
          else if (line == "end")
            {
              AssertThrow(!cur.empty(), "There is no subsection to leave here" + where);
              cur.pop_back();
            }
          else if (line.rfind("set", 0) == 0)
            {
              const auto eq = line.find('=');
              AssertThrow(eq != std::string::npos, "Invalid set line" + where);
              const std::string name = trim(line.substr(3, eq - 3)), value = trim(line.substr(eq + 1));
              auto              it   = entries.find(key(cur, name));
              AssertThrow(it != entries.end(),
                          "No entry with name <" + name + "> was declared in the current subsection" +
                            where);
              AssertThrow(it->second.pattern.match(value),
                          "The entry <" + name + "> does not match its pattern " +
                            it->second.pattern.description + ": <" + value + ">" + where);
              it->second.assign(value);
            }
          else
            AssertThrow(false, "Could not parse line: " + line + where);
        }
      AssertThrow(cur.empty(), "Missing 'end' in " + filename);
    }

  private:
    struct Entry
    {
      Patterns::PatternBase                    pattern;
      std::function<void(const std::string &)> assign;
    };
    std::vector<std::string>     path;
    std::map<std::string, Entry> entries;
    std::map<std::string, int>   known_subsections;

    static std::string trim(const std::string &s)
    {
      const auto a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
      return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    }
    static std::string join(const std::vector<std::string> &p)
    {
      std::string r;
      for (const auto &s : p)
        r += "/" + s;
      return r;
    }
    static std::string key(const std::vector<std::string> &p, const std::string &n)
    {
      return join(p) + "/" + n;
    }
    void subsections_declared(const std::vector<std::string> &p)
    {
      std::vector<std::string> q;
      for (const auto &s : p)
        {
          q.push_back(s);
          known_subsections[join(q)] = 1;
        }
    }
    static void convert(const std::string &s, double &v) { v = std::strtod(s.c_str(), nullptr); }
    static void convert(const std::string &s, int &v) { v = int(std::strtol(s.c_str(), nullptr, 10)); }
    static void convert(const std::string &s, unsigned int &v)
    {
      v = unsigned(std::strtoul(s.c_str(), nullptr, 10));
    }
    static void convert(const std::string &s, std::string &v) { v = s; }
    static void convert(const std::string &s, bool &v) { v = (s == "true"); }
    template <int dim>
    static void convert(const std::string &s, Tensor<1, dim, double> &v)
    {
      std::stringstream ss(s);
      std::string       o;
      int               k = 0;
      while (std::getline(ss, o, ','))
        {
          AssertThrow(k < dim, std::string("too many list entries for a Tensor<1,dim>"));
          v[k++] = std::strtod(o.c_str(), nullptr);
        }
      AssertThrow(k == dim, std::string("too few list entries for a Tensor<1,dim>"));
    }
  };
} // namespace dealii
#endif
