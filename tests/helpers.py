"""Shared problem builders for the tests (parameter values: parameters.prm:26-60)."""
import numpy as np

from dealii_adapter_b200.problem import SolverParameters, make_problem


def nl_params(**kw):
    d = dict(model="neo-Hookean", type_lin="CG", poly_degree=2, scenario="PF", delta_t=0.01,
             mu=0.5e6, nu=0.4, rho=1000.0, tol_lin=1e-6, max_iterations_lin=2.0)
    d.update(kw)
    return SolverParameters(**d)


def lin_params(**kw):
    d = dict(model="linear", type_lin="CG", poly_degree=2, scenario="PF", delta_t=0.005,
             mu=0.5e6, nu=0.4, rho=1000.0, max_iterations_lin=2.0)
    d.update(kw)
    return SolverParameters(**d)


def smooth_field(problem, amp, seed=0):
    """Deterministic smooth displacement-like field on the dofs (zero on constrained dofs)."""
    x = problem.mesh.support_points
    dim = problem.dim
    n_nodes_lookup = np.zeros(problem.n_dofs)
    rng = np.random.RandomState(seed)
    k = rng.uniform(1.0, 3.0, size=(dim, dim))
    ph = rng.uniform(0, 1.0, size=dim)
    # component of each dof: recover from iface-independent rule via cell_dofs local order
    comp = dof_components(problem)
    L = np.array(problem.mesh.p1) - np.array(problem.mesh.p0)
    xi = (x - np.array(problem.mesh.p0)) / L
    for c in range(dim):
        sel = comp == c
        n_nodes_lookup[sel] = amp * np.sin(ph[c] + (xi[sel] * k[c]).sum(axis=1)) * xi[sel, 1]
    n_nodes_lookup[problem.constrained != 0] = 0.0
    return n_nodes_lookup


def local_components(dim, degree):
    """component of every FESystem(FE_Q(degree), dim) local DoF: entity by entity, inside an entity
    component by component (node-major / component-minor only for degree <= 2)."""
    counts = [1] * (1 << dim) + [degree - 1] * (4 if dim == 2 else 12) + \
        [(degree - 1) ** 2] * (1 if dim == 2 else 6) + ([(degree - 1) ** 3] if dim == 3 else [])
    return np.array([c for cnt in counts for c in range(dim) for _ in range(cnt)], dtype=np.int64)


def dof_components(problem):
    comp = np.zeros(problem.n_dofs, dtype=np.int64)
    cd = problem.mesh.cell_dofs.reshape(-1, problem.mesh.dofs_per_cell)
    comp[cd.reshape(-1)] = np.tile(local_components(problem.dim, problem.degree), len(cd))
    return comp


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def distort_mesh(problem, amp=0.12, seed=0):
    """Move every vertex by a smooth non-linear map X -> X + delta(X) (a fraction `amp` of the cell
    size), so that the cells are general (non-affine) quadrilaterals / hexahedra with a positive
    Jacobian. cell_vertices and support_points (= MappingQ1 image of the unit support points) of
    the problem's mesh are updated consistently; the topology and boundary roles stay."""
    import ref_formulas as rf
    mesh, dim, p = problem.mesh, problem.dim, problem.degree
    nv = 1 << dim
    p0, p1 = np.array(mesh.p0), np.array(mesh.p1)
    hcell = (p1 - p0) / np.array(mesh.reps[:dim])
    rng = np.random.RandomState(seed)
    k = rng.uniform(2.0, 5.0, size=(dim, dim))
    ph = rng.uniform(0.0, 2 * np.pi, size=dim)
    verts = mesh.cell_vertices.reshape(-1, dim)
    xi = (verts - p0) / (p1 - p0)
    delta = np.stack([amp * hcell[i] * np.sin(ph[i] + xi @ k[i]) for i in range(dim)], axis=1)
    new = (verts + delta).reshape(mesh.n_cells, nv, dim)
    mesh.cell_vertices = new.reshape(-1).copy()
    nodes = rf.hierarchical_nodes(dim, p)
    s2c = rf.system_to_node_component(dim, p)
    x1 = rf.support_points_1d(p)
    cd = mesh.cell_dofs.reshape(mesh.n_cells, -1)
    sp = mesh.support_points.copy()
    for i, (a, comp) in enumerate(s2c):
        unit = [x1[nodes[a][d]] for d in range(dim)]
        w = np.array([np.prod([unit[d] if (v >> d) & 1 else 1.0 - unit[d] for d in range(dim)])
                      for v in range(nv)])
        sp[cd[:, i]] = np.einsum("v,cvd->cd", w, new)
    mesh.support_points = sp
    return problem
