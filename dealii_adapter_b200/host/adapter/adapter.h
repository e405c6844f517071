// Adapter::Adapter<dim, VectorType, ParameterClass> — public API of the reference kept verbatim
// (include/adapter/adapter.h:62-143: initialize, read_data, advance,
// save_current_state_if_required, reload_old_state_if_required, public members `precice` and
// `deal_boundary_interface_id`). All precice::Participant calls stay here on the host; the bodies
// that touched DoF vectors (format_deal_to_precice :389-417, format_precice_to_deal :421-443,
// the checkpoint copies :457-462 / :482-487) forward to the device through the C-ABI.
//
// VectorType is a handle to a device-resident DoF vector (DeviceVector) instead of
// dealii::Vector<double> / BlockVector<double>; the DoFHandler argument of initialize() is the
// host mesh description that carries what the reference extracts from it (:247-321).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../fake_precice.h"
#include "graft_fem.h"
#include "time_handler.h"

namespace Adapter
{
  struct DeviceVector
  {
    gf_handle handle = nullptr;
    int       id     = -1; // GF_NL_* / GF_LIN_* vector id
  };

  // what Adapter::initialize reads off the DoFHandler: interface IndexSets and support points
  struct InterfaceDescription
  {
    int                 dim = 0;
    int                 n_interface_nodes = 0;
    std::vector<double> interface_nodes_positions; // [x0,y0,(z0),x1,...]
  };

  inline void gf_check(gf_handle h, int rc)
  {
    if (rc != GF_OK)
      throw std::runtime_error(gf_last_error(h)); // reference: AssertThrow(ExcMessage(...))
  }

  template <int dim, typename VectorType, typename ParameterClass>
  class Adapter
  {
  public:
    Adapter(const ParameterClass &parameters, const unsigned int deal_boundary_interface_id)
      : precice(parameters.participant_name, parameters.config_file, this_mpi_process,
                n_mpi_processes)
      , deal_boundary_interface_id(deal_boundary_interface_id)
      , mesh_name(parameters.mesh_name)
      , read_data_name(parameters.read_data_name)
      , write_data_name(parameters.write_data_name)
    {}

    void initialize(const InterfaceDescription &dof_handler, const VectorType &deal_to_precice)
    {
      if (dim != precice.getMeshDimensions(mesh_name))
        throw std::runtime_error(
          "The dimension of your solver needs to be consistent with the dimension specified in "
          "your precice-config file. In case you run one of the tutorials, the dimension can be "
          "specified via cmake -D DIM=dim .");
      n_interface_nodes = dof_handler.n_interface_nodes;
      write_data_buffer.resize(dim * n_interface_nodes);
      read_data_buffer.resize(dim * n_interface_nodes);
      interface_nodes_ids.resize(n_interface_nodes);
      precice.setMeshVertices(mesh_name, dof_handler.interface_nodes_positions,
                              interface_nodes_ids);
      if (precice.requiresInitialData())
        {
          format_deal_to_precice(deal_to_precice);
          precice.writeData(mesh_name, write_data_name, interface_nodes_ids, write_data_buffer);
        }
      precice.initialize();
    }

    void read_data(double relative_read_time, VectorType &precice_to_deal)
    {
      precice.readData(mesh_name, read_data_name, interface_nodes_ids, relative_read_time,
                       read_data_buffer);
      format_precice_to_deal(precice_to_deal);
    }

    void advance(const VectorType &deal_to_precice, const double computed_timestep_length)
    {
      format_deal_to_precice(deal_to_precice);
      precice.writeData(mesh_name, write_data_name, interface_nodes_ids, write_data_buffer);
      precice.advance(computed_timestep_length);
    }

    void save_current_state_if_required(const std::vector<VectorType *> &state_variables,
                                        Time &                           time_class)
    {
      if (precice.requiresWritingCheckpoint())
        {
          n_saved = state_variables.size();
          if (n_saved)
            gf_check(state_variables[0]->handle, gf_state_save(state_variables[0]->handle));
          old_time_value = time_class.current();
        }
    }

    void reload_old_state_if_required(std::vector<VectorType *> &state_variables, Time &time_class)
    {
      if (precice.requiresReadingCheckpoint())
        {
          if (state_variables.size() != n_saved)
            throw std::runtime_error("state_variables are not the same as previously saved.");
          if (n_saved)
            gf_check(state_variables[0]->handle, gf_state_restore(state_variables[0]->handle));
          time_class.set_absolute_time(old_time_value);
        }
    }

    precice::Participant precice;
    const unsigned int   deal_boundary_interface_id;

  private:
    const std::string mesh_name;
    const std::string read_data_name;
    const std::string write_data_name;

    static constexpr unsigned int this_mpi_process = 0;
    static constexpr unsigned int n_mpi_processes  = 1;

    int                 n_interface_nodes = 0;
    std::vector<int>    interface_nodes_ids;
    std::vector<double> read_data_buffer;
    std::vector<double> write_data_buffer;
    size_t              n_saved        = 0;
    double              old_time_value = 0;

    // gather on the device + one small D2H (K9)
    void format_deal_to_precice(const VectorType &deal_to_precice)
    {
      gf_check(deal_to_precice.handle,
               gf_get_interface_displacement(deal_to_precice.handle, write_data_buffer.data()));
    }
    // one small H2D + scatter on the device (K9)
    void format_precice_to_deal(VectorType &precice_to_deal) const
    {
      gf_check(precice_to_deal.handle,
               gf_set_traction(precice_to_deal.handle, read_data_buffer.data()));
    }
  };
} // namespace Adapter
