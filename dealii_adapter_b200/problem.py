"""Problem description shared by the C-ABI binding and the tests: the data the reference's
`make_grid` / `system_setup` / `make_constraints` / `Adapter::initialize` produce on the host
(nonlinear_elasticity.cc:171-380,1094-1150; linear_elasticity.cc:79-244,431-446; adapter.h:229-342)
"""
from dataclasses import dataclass, field

import numpy as np

from .mesh import StructuredMesh, scenario_geometry

MODEL_LINEAR = 0
MODEL_NEO_HOOKEAN = 1


@dataclass
class SolverParameters:
    """parameters.prm keys (parameters.cc:10-171) with the defaults of parameters.h:17-96."""
    # Time
    end_time: float = 1.0
    delta_t: float = 0.1
    output_interval: int = 1
    output_folder: str = ""        # "Output folder"; empty: the Python mirror writes no VTK files
    # System properties
    nu: float = 0.3
    mu: float = 1538462.0
    rho: float = 1000.0
    body_force: tuple = (0.0, 0.0, 0.0)
    # Solver
    model: str = "linear"          # linear | neo-Hookean
    type_lin: str = "Direct"       # CG | Direct
    tol_lin: float = 1e-6
    max_iterations_lin: float = 1.0
    max_iterations_NR: int = 10
    tol_f: float = 1e-9
    tol_u: float = 1e-6
    # Discretization
    poly_degree: int = 3
    theta: float = 0.5
    beta: float = 0.25
    gamma: float = 0.5
    # precice configuration
    scenario: str = "FSI3"
    read_data_name: str = "Stress"
    flap_location: float = 0.0

    @property
    def lam(self):  # parameters.cc:189
        return 2 * self.mu * self.nu / (1 - 2 * self.nu)

    @property
    def data_consistent(self):  # parameters.cc:192-200
        if self.read_data_name.startswith("Stress"):
            return True
        if self.read_data_name.startswith("Force"):
            return False
        raise ValueError("Unknown read data type. Please use 'Force' or 'Stress' in the read data "
                         "naming.")


@dataclass
class Problem:
    dim: int
    degree: int
    model: int
    params: SolverParameters
    mesh: StructuredMesh
    constrained: np.ndarray      # uint8 [n_dofs]
    iface_cell: np.ndarray       # int32 [n_iface_faces]
    iface_face_no: np.ndarray    # int32 [n_iface_faces]
    iface_dofs: np.ndarray       # int32 [dim, n_iface_nodes] (the IndexSets of adapter.h:161-163)
    extra: dict = field(default_factory=dict)

    @property
    def n_dofs(self):
        return self.mesh.n_dofs

    @property
    def n_iface_nodes(self):
        return self.iface_dofs.shape[1]

    def interface_positions(self):
        """[x0,y0,(z0),x1,...] as passed to setMeshVertices (adapter.h:313-326)."""
        return self.mesh.support_points[self.iface_dofs[0]].reshape(-1).copy()


def make_problem(params: SolverParameters, dim: int, reps=None, refinements: int = 0,
                 numbering: str = "cellwise", box=None) -> Problem:
    """Grid + boundary roles of `make_grid` for params.scenario. `reps` overrides the hard-coded
    repetitions; `refinements` multiplies them by 2^r (refine_global, nonlinear:245-246)."""
    p0, p1, base_reps, clamped, interface, zclamp = scenario_geometry(
        params.scenario, dim, params.flap_location)
    if box is not None:  # (p0, p1) override, e.g. a shortened / lengthened flap of equal cell size
        p0, p1 = list(box[0]), list(box[1])
    reps = list(reps) if reps is not None else list(base_reps)
    reps = [r * (1 << refinements) for r in reps]
    mesh = StructuredMesh(dim, params.poly_degree, reps, p0, p1, numbering)
    all_comps = (1 << dim) - 1
    constrained = mesh.boundary_dof_mask(clamped, all_comps)
    if dim == 3:
        mesh.boundary_dof_mask(zclamp, 1 << 2, constrained)
    iface_cell, iface_face_no = mesh.boundary_faces(interface)
    iface_dofs = mesh.interface_dofs(interface)
    model = MODEL_NEO_HOOKEAN if params.model == "neo-Hookean" else MODEL_LINEAR
    return Problem(dim, params.poly_degree, model, params, mesh, constrained, iface_cell,
                   iface_face_no, iface_dofs)
