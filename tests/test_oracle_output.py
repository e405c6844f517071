"""Known-answer pins of the oracle's output path (orc_postprocess): output_results
(nonlinear_elasticity.cc:1215-1254, linear_elasticity.cc:590-629) builds DataOut patches through
MappingQEulerian and the Postprocessor (postprocessor.h:44-76) stores displacement + sym(grad u),
the gradient being taken on the DISPLACED configuration: grad_x u = H (I + H)^-1, H = grad_X u.
The reference ships no output fixtures; these closed forms pin the restatement instead."""
import numpy as np
import pytest

from helpers import dof_components, lin_params, nl_params
from dealii_adapter_b200.problem import make_problem


@pytest.fixture(scope="module")
def orc(native_libs):
    from oracle import oracle_py
    return oracle_py


def field_on_dofs(prob, fn):
    comp = dof_components(prob)
    X = prob.mesh.support_points
    U = fn(X)                      # [n_dofs, dim] evaluated at every dof's support point
    return U[np.arange(prob.n_dofs), comp]


def patch_reference_points(prob):
    """X of the lexicographic patch points of every cell (Q1 geometry)."""
    dim, p = prob.dim, prob.params.poly_degree
    nv = 1 << dim
    verts = np.asarray(prob.mesh.cell_vertices).reshape(prob.mesh.n_cells, nv, dim)
    pts = np.zeros((prob.mesh.n_cells, (p + 1) ** dim, dim))
    for pt in range((p + 1) ** dim):
        xi, rem = [], pt
        for d in range(dim):
            xi.append((rem % (p + 1)) / p)
            rem //= (p + 1)
        for v in range(nv):
            w = 1.0
            for d in range(dim):
                w *= xi[d] if (v >> d) & 1 else 1.0 - xi[d]
            pts[:, pt, :] += w * verts[:, v, :]
    return pts


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (3, 1), (3, 2), (2, 3), (2, 4), (3, 3)])
def test_uniform_stretch_and_rigid_rotation(orc, dim, degree):
    p = nl_params(poly_degree=degree)
    prob = make_problem(p, dim, reps=[2, 3, 2][:dim])
    o = orc.Oracle(prob)
    Xp = patch_reference_points(prob)
    # stretch along x: H = a e_x e_x  ->  grad_x u = a/(1+a) e_x e_x
    a = 0.25
    o.set(orc.NL_TOTAL_DISPLACEMENT, field_on_dofs(prob, lambda X: np.c_[a * X[:, 0], np.zeros((len(X), dim - 1))]))
    pts, fld = o.postprocess(orc.NL_TOTAL_DISPLACEMENT)
    assert np.allclose(fld[..., 0], a * Xp[..., 0], atol=1e-14)
    assert np.allclose(fld[..., 1:dim], 0.0, atol=1e-15)
    strain = fld[..., dim:].reshape(fld.shape[0], fld.shape[1], dim, dim)
    want = np.zeros((dim, dim))
    want[0, 0] = a / (1 + a)
    assert np.allclose(strain, want, atol=1e-13)
    assert np.allclose(pts, Xp + fld[..., :dim], atol=1e-15)
    # rigid rotation about z by theta: grad_x u = I - R^T, strain = (1 - cos) on the rotated plane
    th = 0.3
    R = np.eye(dim)
    R[:2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
    o.set(orc.NL_TOTAL_DISPLACEMENT, field_on_dofs(prob, lambda X: X @ (R - np.eye(dim)).T))
    pts, fld = o.postprocess(orc.NL_TOTAL_DISPLACEMENT)
    strain = fld[..., dim:].reshape(fld.shape[0], fld.shape[1], dim, dim)
    want = np.eye(dim) - 0.5 * (R + R.T)
    assert np.allclose(strain, want, atol=1e-13)
    assert np.allclose(pts, Xp @ R.T, atol=1e-14)


def test_quadratic_field_q2_matches_the_closed_form(orc):
    dim = 3
    prob = make_problem(lin_params(poly_degree=2), dim, reps=[2, 2, 3])
    o = orc.Oracle(prob)
    A = np.array([[0.03, -0.02, 0.01], [0.015, 0.02, -0.01], [-0.02, 0.01, 0.025]])
    B = np.array([0.2, -0.1, 0.15])

    def u(X):      # u_c = A_c . X + B_c x y  (in Q2)
        return X @ A.T + np.outer(X[:, 0] * X[:, 1], B)

    o.set(orc.LIN_DISPLACEMENT, field_on_dofs(prob, u))
    pts, fld = o.postprocess(orc.LIN_DISPLACEMENT)
    Xp = patch_reference_points(prob).reshape(-1, dim)
    H = np.repeat(A[None], len(Xp), axis=0)
    H[:, :, 0] += np.outer(Xp[:, 1], B)
    H[:, :, 1] += np.outer(Xp[:, 0], B)
    g = H @ np.linalg.inv(np.eye(dim) + H)
    want = 0.5 * (g + np.transpose(g, (0, 2, 1)))
    got = fld[..., dim:].reshape(-1, dim, dim)
    assert np.abs(got - want).max() < 1e-12
    assert np.abs(fld[..., :dim].reshape(-1, dim) - u(Xp)).max() < 1e-14
    # strain components are stored row-major: index dim + d*dim + e, symmetric
    assert np.array_equal(fld[..., dim + 1], fld[..., dim + dim])
