"""The 'by definition' references used for meshes WITH hanging nodes (tests/helpers.py:
reference_linear_steps, reference_nonlinear_step - condensation with the constraint matrix, sparse
LU) are checked here where the answer is known: without constraint lines they must reproduce the
oracle's own time steps; with lines the solution must be continuous across the refined edge."""
import numpy as np
import pytest

from helpers import (constraint_matrix, hanging_node_problem, lin_params, nl_params, rel_err,
                     reference_linear_steps, reference_nonlinear_step)
from dealii_adapter_b200.problem import make_problem


@pytest.fixture(scope="module")
def orc(native_libs):
    from oracle import oracle_py
    return oracle_py


def test_references_reproduce_the_oracle_without_constraint_lines(orc):
    p = lin_params(poly_degree=2, type_lin="Direct", body_force=(0.0, -9.81, 0.0))
    prob = make_problem(p, 2, reps=[3, 6])
    bufs = [np.tile([50.0 * (k + 1), -20.0], prob.n_iface_nodes) for k in range(3)]
    ref = reference_linear_steps(orc, prob, bufs)
    o = orc.Oracle(prob)
    o.lin_assemble_system()
    for k, buf in enumerate(bufs):
        o.format_precice_to_deal(buf, orc.LIN_STRESS)
        o.lin_step()
        assert rel_err(ref[k], o.get(orc.LIN_DISPLACEMENT)) < 1e-9
    pn = nl_params(poly_degree=2, type_lin="Direct", scenario="PF")
    prob = make_problem(pn, 2, reps=[3, 6])
    buf = np.tile([900.0, 0.0], prob.n_iface_nodes)
    ref = reference_nonlinear_step(orc, prob, buf)
    o = orc.Oracle(prob)
    o.format_precice_to_deal(buf, orc.NL_EXTERNAL_STRESS)
    o.nl_timestep()
    assert rel_err(ref, o.get(orc.NL_TOTAL_DISPLACEMENT)) < 1e-8


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_hanging_node_solution_is_conforming(orc, dim, degree):
    prob = hanging_node_problem(lin_params(poly_degree=degree, type_lin="Direct"), degree, dim)
    dof, ptr, master, weight = prob.extra["constraint_lines"]
    # fine nodes of the face x = 1 minus the coarse ones, dim components each
    assert len(dof) == dim * ((2 * degree + 1) ** (dim - 1) - (degree + 1) ** (dim - 1))
    bufs = [np.tile([0.0, -200.0, 50.0][:dim], prob.n_iface_nodes)] * 2
    d = reference_linear_steps(orc, prob, bufs)[-1]
    Cm = constraint_matrix(prob)
    masters_only = d.copy()
    masters_only[dof] = 0.0
    assert np.abs(Cm @ masters_only - d).max() <= 1e-15 * np.abs(d).max()   # u = C u_masters
    assert np.abs(d[dof]).max() > 0
    pn = nl_params(poly_degree=degree, type_lin="Direct")
    prob = hanging_node_problem(pn, degree, dim)
    u = reference_nonlinear_step(orc, prob, np.tile([0.0, -1500.0, 300.0][:dim], prob.n_iface_nodes))
    m = u.copy()
    m[dof] = 0.0
    assert np.abs(Cm @ m - u).max() <= 1e-15 * np.abs(u).max() and np.abs(u[dof]).max() > 0
