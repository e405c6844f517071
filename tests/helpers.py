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


def hanging_node_problem(params, degree, dim=2, clamp_axis=0):
    """A mesh WITH hanging nodes, which the structured stand-in cannot produce: the domain
    [0,2] x [0,1] (x [0,1]), left half one cell, right half refined once (2^dim cells). The nodes
    of the fine cells on the face x = 1 that the coarse cell does not own hang; their constraint
    lines are the coarse face's shape functions at their positions (what
    DoFTools::make_hanging_node_constraints produces, linear_elasticity.cc:196-207).
    Roles: x = 0 clamped (clamp_axis = 1: y = 0 instead, which makes masters of the hanging
    nodes Dirichlet dofs), y = 1 and x = 2 interface, the rest free. degree 1 or 2 (equidistant
    nodes). Returns a Problem whose `extra["constraint_lines"]` = (dof, ptr, master, weight)."""
    import itertools
    from types import SimpleNamespace
    import ref_formulas as rf
    from dealii_adapter_b200.problem import MODEL_LINEAR, MODEL_NEO_HOOKEAN, Problem
    p = degree
    assert p in (1, 2) and dim in (2, 3)
    boxes = [((0.0,) * dim, (1.0,) * dim)]
    fine_index = {}
    for idx in itertools.product(range(2), repeat=dim - 1):          # (k,) j slowest .. x fastest
        for i in range(2):
            ijk = (i,) + tuple(reversed(idx))                        # (i, j[, k])
            fine_index[ijk] = len(boxes)
            lo = (1.0 + 0.5 * ijk[0],) + tuple(0.5 * t for t in ijk[1:])
            boxes.append((lo, tuple(x + 0.5 for x in lo)))
    nodes = rf.hierarchical_nodes(dim, p)
    s2c = rf.system_to_node_component(dim, p)
    ids, coords, cell_dofs, cell_vertices = {}, [], [], []
    for lo, hi in boxes:
        for v in range(1 << dim):
            cell_vertices += [hi[d] if (v >> d) & 1 else lo[d] for d in range(dim)]
        row = []
        for a, c in s2c:
            x = tuple(round(lo[d] + (hi[d] - lo[d]) * nodes[a][d] / p, 12) for d in range(dim))
            if x not in ids:
                ids[x] = len(ids)
                coords.append(x)
            row.append(ids[x] * dim + c)
        cell_dofs += row
    n_nodes = len(ids)
    n_dofs = n_nodes * dim
    coords = np.array(coords)
    support_points = np.repeat(coords, dim, axis=0)
    # hanging nodes: on the face x = 1 and not a node of the coarse cell
    coarse_1d = np.linspace(0.0, 1.0, p + 1)
    on_grid = lambda t: bool(np.any(np.abs(coarse_1d - t) < 1e-12))
    dof, ptr, master, weight = [], [0], [], []
    for x, n in ids.items():
        if abs(x[0] - 1.0) < 1e-12 and not all(on_grid(t) for t in x[1:]):
            w1 = [rf.lagrange_1d(p, np.array([t]))[0][0] for t in x[1:]]   # coarse face basis at x
            for c in range(dim):
                dof.append(n * dim + c)
                for ks in itertools.product(range(p + 1), repeat=dim - 1):
                    w = float(np.prod([w1[d][k] for d, k in enumerate(ks)]))
                    if abs(w) > 1e-14:
                        m = (1.0,) + tuple(round(float(coarse_1d[k]), 12) for k in ks)
                        master.append(ids[m] * dim + c)
                        weight.append(w)
                ptr.append(len(master))
    constrained = np.zeros(n_dofs, dtype=np.uint8)
    constrained[np.repeat(np.abs(coords[:, clamp_axis]) < 1e-12, dim)] = 1
    # AffineConstraints::close(): chains are resolved - a master that is a Dirichlet dof (value 0)
    # drops out of the line; a hanging dof ON the Dirichlet boundary ends up with an empty line,
    # i.e. it is an ordinary Dirichlet dof (such lines go into `constrained`)
    dof2, ptr2, master2, weight2 = [], [0], [], []
    for k, s_ in enumerate(dof):
        if constrained[s_]:
            continue
        for j in range(ptr[k], ptr[k + 1]):
            if not constrained[master[j]]:
                master2.append(master[j])
                weight2.append(weight[j])
        dof2.append(s_)
        ptr2.append(len(master2))
    dof, ptr, master, weight = dof2, ptr2, master2, weight2
    # interface faces: y = 1 (face 3) of the coarse cell and of the fine cells with j = 1, x = 2
    # (face 1) of the fine cells with i = 1
    iface = [(0, 3)] + [(c, 3) for ijk, c in fine_index.items() if ijk[1] == 1] + \
        [(c, 1) for ijk, c in fine_index.items() if ijk[0] == 1]
    iface.sort()
    on_if = (np.abs(coords[:, 1] - 1.0) < 1e-12) | (np.abs(coords[:, 0] - 2.0) < 1e-12)
    if_nodes = np.nonzero(on_if)[0]
    iface_dofs = np.stack([if_nodes * dim + c for c in range(dim)]).astype(np.int32)
    mesh = SimpleNamespace(dim=dim, degree=p, n_cells=len(boxes), n_dofs=n_dofs, n_nodes=n_nodes,
                           dofs_per_cell=dim * (p + 1) ** dim, reps=[2] + [1] * (dim - 1),
                           p0=[0.0] * dim, p1=[2.0] + [1.0] * (dim - 1), numbering="custom",
                           cell_dofs=np.array(cell_dofs, dtype=np.int32),
                           cell_vertices=np.array(cell_vertices, dtype=np.float64),
                           support_points=support_points)
    model = MODEL_NEO_HOOKEAN if params.model == "neo-Hookean" else MODEL_LINEAR
    prob = Problem(dim, p, model, params, mesh, constrained,
                   np.array([c for c, f in iface], dtype=np.int32),
                   np.array([f for c, f in iface], dtype=np.int32), iface_dofs)
    prob.extra["constraint_lines"] = (np.array(dof, dtype=np.int32), np.array(ptr, dtype=np.int64),
                                      np.array(master, dtype=np.int32), np.array(weight))
    return prob


def constraint_matrix(prob):
    """C [n_dofs, n_dofs] (scipy): identity on unconstrained dofs, the weights in the rows of the
    hanging dofs (whose own columns are empty)."""
    import scipy.sparse as sp
    dof, ptr, master, weight = prob.extra["constraint_lines"]
    n = prob.n_dofs
    C = sp.lil_matrix((n, n))
    hanging = set(dof.tolist())
    for i in range(n):
        if i not in hanging:
            C[i, i] = 1.0
    for k, s in enumerate(dof):
        for j in range(ptr[k], ptr[k + 1]):
            C[s, master[j]] = weight[j]
    return C.tocsr()


def reference_linear_steps(orc, prob, buffers):
    """ElastoDynamics time steps 'by definition' with scipy on top of the oracle's UNcondensed
    K, M and consistent loading: assemble_rhs (:378-454) incl. hanging_node_constraints.condense,
    apply_boundary_values, the solve as a sparse LU, distribute (:571-572), update_displacement
    (:579-586). `buffers`: one interface traction buffer per step. Returns the displacement vectors.
    Without constraint lines this is the oracle's own lin_step (a CPU test checks that)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    p, n = prob.params, prob.n_dofs
    dt, th = p.delta_t, p.theta
    o = orc.Oracle(prob)
    o.lin_assemble_system()
    K, M = o.csr(orc.MAT_STIFFNESS), o.csr(orc.MAT_MASS)
    body = o.get(orc.LIN_BODY_FORCE)
    Cm = constraint_matrix(prob) if prob.extra.get("constraint_lines") is not None else sp.identity(n, format="csr")
    hanging = np.zeros(n, dtype=bool)
    if prob.extra.get("constraint_lines") is not None:
        hanging[prob.extra["constraint_lines"][0]] = True
    fixed = (prob.constrained != 0) | hanging
    free = np.nonzero(~fixed)[0]
    A = (Cm.T @ (M + (th * dt) ** 2 * K) @ Cm).tocsr()[free][:, free].tocsc()
    lu = spla.splu(A)
    v, d, F_old = np.zeros(n), np.zeros(n), np.zeros(n)
    out = []
    for buf in buffers:
        o.format_precice_to_deal(buf, orc.LIN_STRESS)
        o.set(orc.LIN_OLD_STRESS, np.zeros(n))
        o.set(orc.LIN_VELOCITY, np.zeros(n))
        o.set(orc.LIN_DISPLACEMENT, np.zeros(n))
        o.lin_assemble_rhs()
        F = o.get(orc.LIN_OLD_STRESS)          # consistent loading + body force of this step
        rhs = dt * th * F + dt * (1 - th) * F_old + M @ v - th * (1 - th) * dt * dt * (K @ v) - dt * (K @ d)
        rhs = Cm.T @ rhs
        v_new = np.zeros(n)
        v_new[free] = lu.solve(rhs[free])
        v_new = Cm @ v_new
        d = d + dt * th * v_new + dt * (1 - th) * v
        v, F_old = v_new, F
        out.append(d.copy())
    return out


def reference_nonlinear_step(orc, prob, buf, n_newton=8):
    """One Solid time step from rest 'by definition': Newton iterations on top of the oracle's
    UNcondensed tangent / residual (Dirichlet rows already eliminated), condensed with the
    constraint matrix, solved by a sparse LU, distributed. Returns the total displacement."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    n = prob.n_dofs
    o = orc.Oracle(prob)
    o.format_precice_to_deal(buf, orc.NL_EXTERNAL_STRESS)
    Cm = constraint_matrix(prob) if prob.extra.get("constraint_lines") is not None else sp.identity(n, format="csr")
    hanging = np.zeros(n, dtype=bool)
    if prob.extra.get("constraint_lines") is not None:
        hanging[prob.extra["constraint_lines"][0]] = True
    free = np.nonzero(~((prob.constrained != 0) | hanging))[0]
    delta = np.zeros(n)
    for it in range(n_newton):
        o.set(orc.NL_SOLUTION_DELTA, delta)
        o.nl_update_acceleration()
        o.nl_assemble_system()
        A, b = o.csr(orc.MAT_TANGENT), o.get(orc.NL_SYSTEM_RHS)
        At = (Cm.T @ A @ Cm).tocsr()[free][:, free].tocsc()
        x = np.zeros(n)
        x[free] = spla.splu(At).solve((Cm.T @ b)[free])
        delta = delta + Cm @ x
    return delta
