"""Python mirror of the reference's host classes for the hot path, driving the C-ABI:

  Time            include/adapter/time_handler.h:21-84
  Adapter         include/adapter/adapter.h:26-489 (initialize / read_data / advance /
                  save_current_state_if_required / reload_old_state_if_required)
  Solid           source/nonlinear_elasticity/nonlinear_elasticity.cc (run :99-167,
                  solve_nonlinear_timestep :410-499)
  ElastoDynamics  source/linear_elasticity/linear_elasticity.cc (run :634-716)
  FakeParticipant scripted stand-in for precice::Participant — preCICE is a
                  separate process in the reference and is out of scope; tests and bench use this.

The same classes exist in C++ (host/*.h) for the drop-in build; this mirror is what bench.py and the
pytest parity tests drive. The private hot members forward to libgraftfem.so; nothing numerical
happens in Python.
"""
import math

import numpy as np

from . import capi
from .problem import MODEL_NEO_HOOKEAN, Problem


class Time:
    def __init__(self, time_end, delta_t):
        self.timestep = 0
        self.time_current = 0.0
        self.time_end = time_end
        self.delta_t = delta_t

    def current(self):
        return self.time_current

    def end(self):
        return self.time_end

    def get_delta_t(self):
        return self.delta_t

    def get_timestep(self):
        return self.timestep

    def set_absolute_time(self, new_time):  # time_handler.h:63-70
        factor = math.pow(10, 10)
        self.timestep = int(round((new_time / self.delta_t) * factor) / factor)
        self.time_current = new_time

    def increment(self):
        self.time_current += self.delta_t
        self.timestep += 1


class FakeParticipant:
    """Scripted precice::Participant. `traction(t, iteration)` returns the interface read data
    [x0,y0,(z0),x1,...] for the window ending at time t; `n_subiterations` > 1 emulates an implicit
    coupling scheme that repeats every window (checkpoint write at the first pass, read-back after
    every non-final pass)."""

    def __init__(self, dim, n_windows, delta_t, traction, n_subiterations=1):
        self.dim = dim
        self.n_windows = n_windows
        self.delta_t = delta_t
        self.traction = traction
        self.n_sub = n_subiterations
        self.window = 0
        self.iteration = 0
        self.positions = None
        self.written = []          # (window, iteration, data) of every writeData
        self._window_complete = False
        self.finalized = False

    # -- precice::Participant subset
    def getMeshDimensions(self, mesh_name):
        return self.dim

    def setMeshVertices(self, mesh_name, positions):
        self.positions = np.array(positions, dtype=np.float64)
        return np.arange(len(positions) // self.dim, dtype=np.int32)

    def requiresInitialData(self):
        return False

    def writeData(self, mesh_name, data_name, ids, values):
        self.written.append((self.window, self.iteration, np.array(values, dtype=np.float64)))

    def initialize(self):
        pass

    def readData(self, mesh_name, data_name, ids, relative_read_time):
        t = (self.window * self.delta_t) + relative_read_time
        return np.asarray(self.traction(t, self.iteration), dtype=np.float64)

    def advance(self, dt):
        self.iteration += 1
        if self.iteration >= self.n_sub:
            self.iteration = 0
            self.window += 1
            self._window_complete = True
        else:
            self._window_complete = False

    def requiresWritingCheckpoint(self):
        return self.n_sub > 1 and self.iteration == 0

    def requiresReadingCheckpoint(self):
        return self.n_sub > 1 and not self._window_complete

    def isCouplingOngoing(self):
        return self.window < self.n_windows

    def getMaxTimeStepSize(self):
        return self.delta_t

    def isTimeWindowComplete(self):
        return self._window_complete

    def finalize(self):
        self.finalized = True


class Adapter:
    """adapter.h:26-489 with the copy/format bodies on the device."""

    def __init__(self, parameters, deal_boundary_interface_id, participant, handle):
        self.precice = participant
        self.deal_boundary_interface_id = deal_boundary_interface_id
        self.mesh_name = "Solid-Mesh"
        self.read_data_name = parameters.read_data_name
        self.write_data_name = "Displacement"
        self._h = handle
        self.n_interface_nodes = 0
        self.old_time_value = 0.0

    def initialize(self, problem: Problem):  # :229-342
        if self.precice.getMeshDimensions(self.mesh_name) != problem.dim:
            raise RuntimeError("The dimension of your solver needs to be consistent with the "
                               "dimension specified in your precice-config file.")
        self.n_interface_nodes = problem.n_iface_nodes
        self.interface_nodes_ids = self.precice.setMeshVertices(self.mesh_name,
                                                                problem.interface_positions())
        if self.precice.requiresInitialData():
            self.precice.writeData(self.mesh_name, self.write_data_name, self.interface_nodes_ids,
                                   self._h.get_interface_displacement())
        self.precice.initialize()

    def read_data(self, relative_read_time):  # :346-361
        buf = self.precice.readData(self.mesh_name, self.read_data_name, self.interface_nodes_ids,
                                    relative_read_time)
        self._h.set_traction(buf)  # format_precice_to_deal :421-443

    def advance(self, computed_timestep_length):  # :365-385
        buf = self._h.get_interface_displacement()  # format_deal_to_precice :389-417
        self.precice.writeData(self.mesh_name, self.write_data_name, self.interface_nodes_ids, buf)
        self.precice.advance(computed_timestep_length)

    def save_current_state_if_required(self, time_class):  # :447-464
        if self.precice.requiresWritingCheckpoint():
            self._h.state_save()
            self.old_time_value = time_class.current()

    def reload_old_state_if_required(self, time_class):  # :468-489
        if self.precice.requiresReadingCheckpoint():
            self._h.state_restore()
            time_class.set_absolute_time(self.old_time_value)


class Solid:
    """Nonlinear_Elasticity::Solid<dim>: Newmark + Newton, neo-Hooke."""

    def __init__(self, problem: Problem, participant, device=0, handle=None):
        assert problem.model == MODEL_NEO_HOOKEAN
        self.problem = problem
        self.parameters = problem.params
        if not self.parameters.data_consistent:  # :83-87
            raise RuntimeError("The neo-Hookean solid doesn't support 'Force' data reading.")
        self.boundary_interface_id = 7
        self.time = Time(self.parameters.end_time, self.parameters.delta_t)
        self.handle = handle if handle is not None else capi.Handle(problem, device=device)
        self.adapter = Adapter(self.parameters, self.boundary_interface_id, participant, self.handle)
        self.history = []  # per timestep: list of Newton table rows
        self.newton_solves = 0
        self.assemblies = 0

    def solve_nonlinear_timestep(self):  # :410-499
        p, h = self.parameters, self.handle
        type_lin = 0 if p.type_lin == "CG" else 1
        error_residual = error_residual_0 = error_residual_norm = 1.0
        error_update = error_update_0 = error_update_norm = 1.0
        rows = []
        newton_iteration = 0
        while newton_iteration < p.max_iterations_NR:
            error_residual = h.nl_newton_assemble()  # make_constraints, update_acceleration, assemble
            self.assemblies += 1
            if newton_iteration == 0:
                error_residual_0 = error_residual
            error_residual_norm = error_residual
            if error_residual_0 != 0.0:
                error_residual_norm /= error_residual_0
            if newton_iteration > 0 and ((error_update_norm <= p.tol_u or error_update <= 1e-15) and
                                         (error_residual_norm <= p.tol_f or error_residual <= 5e-9)):
                break
            lin_it, lin_res, error_update = h.nl_newton_solve(type_lin, p.tol_lin, p.max_iterations_lin)
            self.newton_solves += 1
            if newton_iteration == 0:
                error_update_0 = error_update
            error_update_norm = error_update
            if error_update_0 != 0.0:
                error_update_norm /= error_update_0
            rows.append((lin_it, lin_res, error_residual_norm, error_residual, error_update_norm,
                         error_update))
            newton_iteration += 1
        if not newton_iteration < p.max_iterations_NR:
            raise RuntimeError("No convergence in nonlinear solver!")
        self.history.append(rows)
        return rows

    def step(self):
        """One pass of the coupling loop body (:117-162) without output_results."""
        a, t = self.adapter, self.time
        a.save_current_state_if_required(t)
        self.handle.nl_begin_step()  # solution_delta = 0 :121
        t.increment()
        if abs(t.get_delta_t() - a.precice.getMaxTimeStepSize()) >= 1e-10:  # :125-133
            raise RuntimeError("This solver supports only constant time-step sizes.")
        a.read_data(t.get_delta_t())
        self.solve_nonlinear_timestep()
        self.handle.nl_end_step()  # :139-144
        a.advance(t.get_delta_t())
        a.reload_old_state_if_required(t)

    def output_results(self):  # :1215-1254
        """solution-<n>.vtk in parameters.output_folder (skipped when no folder is configured):
        patch fields from the device (gf_postprocess), points and file by the C++ host writer."""
        folder = getattr(self.parameters, "output_folder", "")
        if not folder:
            return None
        import os
        os.makedirs(folder, exist_ok=True)
        name = os.path.join(folder, "solution-%03d.vtk"
                            % (self.time.get_timestep() // self.parameters.output_interval))
        self.problem.mesh.write_vtk(name, self.handle.postprocess(capi.NL_TOTAL_DISPLACEMENT))
        return name

    def run(self):  # :99-167
        self.output_results()  # :103
        self.adapter.initialize(self.problem)
        while self.adapter.precice.isCouplingOngoing():
            self.step()
            if self.adapter.precice.isTimeWindowComplete() and \
                    self.time.get_timestep() % self.parameters.output_interval == 0:  # :150-153
                self.output_results()
        self.adapter.precice.finalize()


class ElastoDynamics:
    """Linear_Elasticity::ElastoDynamics<dim>: one-step-theta in the velocity unknown."""

    def __init__(self, problem: Problem, participant, device=0, handle=None):
        assert problem.model != MODEL_NEO_HOOKEAN
        self.problem = problem
        self.parameters = problem.params
        self.interface_boundary_id = 6
        self.time = Time(self.parameters.end_time, self.parameters.delta_t)
        self.handle = handle if handle is not None else capi.Handle(problem, device=device)
        self.adapter = Adapter(self.parameters, self.interface_boundary_id, participant, self.handle)
        self.history = []

    def step(self):  # :654-710
        a, t, p = self.adapter, self.time, self.parameters
        a.save_current_state_if_required(t)
        t.increment()
        if abs(t.get_delta_t() - a.precice.getMaxTimeStepSize()) >= 1e-10:
            raise RuntimeError("This solver supports only constant time-step sizes.")
        a.read_data(t.get_delta_t())
        # assemble_rhs :680, solve :683, update_displacement :686
        self.history.append(self.handle.lin_step(0 if p.type_lin == "CG" else 1, p.max_iterations_lin))
        a.advance(t.get_delta_t())
        a.reload_old_state_if_required(t)

    def output_results(self):  # :590-629
        folder = getattr(self.parameters, "output_folder", "")
        if not folder:
            return None
        import os
        os.makedirs(folder, exist_ok=True)
        name = os.path.join(folder, "solution-%03d.vtk"
                            % (self.time.get_timestep() // self.parameters.output_interval))
        self.problem.mesh.write_vtk(name, self.handle.postprocess(capi.LIN_DISPLACEMENT))
        return name

    def run(self):  # :634-716
        self.output_results()  # :640
        self.handle.lin_assemble_once()  # assemble_system :642
        self.adapter.initialize(self.problem)
        while self.adapter.precice.isCouplingOngoing():
            self.step()
            if self.adapter.precice.isTimeWindowComplete() and \
                    self.time.get_timestep() % self.parameters.output_interval == 0:  # :699-702
                self.output_results()
        self.adapter.precice.finalize()
