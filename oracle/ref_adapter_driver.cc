/* TEST INFRASTRUCTURE ONLY. The reference's Adapter bodies on the DoF vectors — the member
 * definitions format_deal_to_precice, format_precice_to_deal, save_current_state_if_required and
 * reload_old_state_if_required (include/adapter/adapter.h:389-489), cut out at build time and
 * compiled against a stand-in Adapter class (IndexSets = ascending index lists, the preCICE
 * participant = two scripted flags). Time is the reference's own time_handler.h.
 *
 * stdin: dim n_dofs n_iface  x_comp[n_iface] y_comp[n_iface] (z_comp[n_iface])
 *        deal_vector[n_dofs]  read_data_buffer[dim*n_iface]
 * stdout: write_data_buffer; the vector after format_precice_to_deal (started from deal_vector);
 *         checkpoint script: see main() */
#include <cstdio>
#include <iostream>
#include <set>

#include "assembly_shim.h"
#include "adapter/time_handler.h"

namespace Adapter
{
  using namespace dealii;
  struct IndexSetShim
  {
    std::set<unsigned> idx; // IndexSet iterates its elements in ascending order
    auto               begin() const { return idx.begin(); }
  };
  struct ParticipantShim
  {
    bool write_checkpoint = false, read_checkpoint = false;
    bool requiresWritingCheckpoint() const { return write_checkpoint; }
    bool requiresReadingCheckpoint() const { return read_checkpoint; }
    // read_data / advance (adapter.h:346-385): record the calls and their arguments
    void readData(const std::string &mesh, const std::string &data, const std::vector<int> &ids,
                  double relative_read_time, std::vector<double> &values) const
    {
      printf("EVENT readData %s %s %zu %.17g\n", mesh.c_str(), data.c_str(), ids.size(),
             relative_read_time);
      for (size_t i = 0; i < values.size(); ++i)
        values[i] = 100.0 + double(i);
    }
    void writeData(const std::string &mesh, const std::string &data, const std::vector<int> &ids,
                   const std::vector<double> &values)
    {
      printf("EVENT writeData %s %s %zu %.17g\n", mesh.c_str(), data.c_str(), ids.size(),
             values.empty() ? 0.0 : values.back());
    }
    void advance(double dt) { printf("EVENT advance %.17g\n", dt); }
  };
  template <int dim, typename VectorType, typename ParameterClass>
  class Adapter
  {
  public:
    ParticipantShim         precice;
    IndexSetShim            coupling_dofs_x_comp, coupling_dofs_y_comp, coupling_dofs_z_comp;
    int                     n_interface_nodes = 0;
    std::vector<double>     read_data_buffer, write_data_buffer;
    std::vector<VectorType> old_state_data;
    double                  old_time_value = 0;
    std::string mesh_name = "dealii-mesh", read_data_name = "Stress", write_data_name = "Displacement";
    std::vector<int>        interface_nodes_ids;
    void                    read_data(double relative_read_time, VectorType &precice_to_deal);
    void                    advance(const VectorType &deal_to_precice, const double computed_timestep_length);
    void                    format_deal_to_precice(const VectorType &deal_to_precice);
    void                    format_precice_to_deal(VectorType &precice_to_deal) const;
    void save_current_state_if_required(const std::vector<VectorType *> &state_variables,
                                        Time &                           time_class);
    void reload_old_state_if_required(std::vector<VectorType *> &state_variables, Time &time_class);
  };
#include "adapter_io_extract.inc"
#include "adapter_extract.inc"
} // namespace Adapter

using namespace dealii;

template <int dim>
int run(unsigned n_dofs, unsigned n_iface)
{
  Adapter::Adapter<dim, Vector<double>, Parameters::AllParameters> a;
  a.n_interface_nodes = int(n_iface);
  Adapter::IndexSetShim *sets[3] = {&a.coupling_dofs_x_comp, &a.coupling_dofs_y_comp,
                                    &a.coupling_dofs_z_comp};
  for (int c = 0; c < dim; ++c)
    for (unsigned i = 0; i < n_iface; ++i)
      {
        unsigned k;
        std::cin >> k;
        sets[c]->idx.insert(k);
      }
  Vector<double> v(n_dofs);
  for (unsigned i = 0; i < n_dofs; ++i)
    std::cin >> v[i];
  a.read_data_buffer.resize(size_t(dim) * n_iface);
  a.write_data_buffer.resize(size_t(dim) * n_iface);
  for (auto &x : a.read_data_buffer)
    std::cin >> x;
  if (!std::cin)
    return 2;
  a.format_deal_to_precice(v);
  for (double x : a.write_data_buffer)
    printf("%.17g ", x);
  printf("\n");
  Vector<double> w = v;
  a.format_precice_to_deal(w);
  for (unsigned i = 0; i < n_dofs; ++i)
    printf("%.17g ", w[i]);
  printf("\n");
  // checkpoint script: save at t = 3 dt (flag on), advance two steps and scale the state, then
  // reload (flag on): state and time must be those of the checkpoint; with the flags off nothing
  // happens
  Adapter::Time                 time(1e9, 0.01);
  Vector<double>                s0 = v, s1 = w;
  std::vector<Vector<double> *> state = {&s0, &s1};
  for (int k = 0; k < 3; ++k)
    time.increment();
  a.save_current_state_if_required(state, time); // flag off: no copy
  printf("%zu ", a.old_state_data.size());
  a.precice.write_checkpoint = true;
  a.save_current_state_if_required(state, time);
  a.precice.write_checkpoint = false;
  printf("%zu %.17g\n", a.old_state_data.size(), a.old_time_value);
  time.increment();
  time.increment();
  s0 *= 2.0;
  s1 *= -1.0;
  a.reload_old_state_if_required(state, time); // flag off: unchanged
  printf("%u %.17g %.17g\n", time.get_timestep(), time.current(), s0[0]);
  a.precice.read_checkpoint = true;
  a.reload_old_state_if_required(state, time);
  printf("%u %.17g\n", time.get_timestep(), time.current());
  bool same = true;
  for (unsigned i = 0; i < n_dofs; ++i)
    same = same && s0[i] == v[i] && s1[i] == w[i];
  printf("%d\n", same ? 1 : 0);
  // read_data / advance: order of the preCICE calls around the format functions
  a.interface_nodes_ids.assign(n_iface, 0);
  Vector<double> target(n_dofs);
  a.read_data(0.01, target);
  printf("EVENT read_value %.17g\n", target[*a.coupling_dofs_x_comp.begin()]);
  a.advance(v, 0.01);
  return 0;
}

int main()
{
  unsigned dim, n_dofs, n_iface;
  std::cin >> dim >> n_dofs >> n_iface;
  return dim == 2 ? run<2>(n_dofs, n_iface) : run<3>(n_dofs, n_iface);
}
