"""GPU parity of the output-path kernel (gf_postprocess, csrc/postprocess.cu) against the oracle's
restatement of output_results + Postprocessor (nonlinear_elasticity.cc:1215-1254,
linear_elasticity.cc:590-629, postprocessor.h:44-76). FP64, tolerance 1e-12 relative to the
largest entry of each field class (displacement, strain)."""
import numpy as np
import pytest

from helpers import lin_params, nl_params, smooth_field
from dealii_adapter_b200.problem import make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs(native_libs):
    from dealii_adapter_b200 import build, capi
    from oracle import oracle_py as orc
    build.build_cuda()
    capi.lib()
    return capi, orc


@pytest.mark.parametrize("model,dim,degree,reps,numbering", [
    ("nl", 3, 2, [3, 6, 2], "lexicographic"),
    ("nl", 3, 1, [4, 5, 3], "cellwise"),
    ("nl", 2, 2, [6, 9], "component_wise"),
    ("lin", 2, 1, [5, 7], "cellwise"),
    ("lin", 3, 2, [2, 4, 2], "component_wise"),
])
def test_postprocess_matches_oracle(libs, model, dim, degree, reps, numbering):
    capi, orc = libs
    p = nl_params(poly_degree=degree) if model == "nl" else lin_params(poly_degree=degree)
    prob = make_problem(p, dim, reps=reps, numbering=numbering)
    u = smooth_field(prob, 0.02, seed=4)
    which_g = capi.NL_TOTAL_DISPLACEMENT if model == "nl" else capi.LIN_DISPLACEMENT
    which_o = orc.NL_TOTAL_DISPLACEMENT if model == "nl" else orc.LIN_DISPLACEMENT
    h = capi.Handle(prob)
    h.set_vector(which_g, u)
    got = h.postprocess(which_g)
    o = orc.Oracle(prob)
    o.set(which_o, u)
    pts, ref = o.postprocess(which_o)
    assert got.shape == ref.shape
    assert np.abs(got[..., :dim] - ref[..., :dim]).max() <= 1e-12 * np.abs(ref[..., :dim]).max()
    assert np.abs(got[..., dim:] - ref[..., dim:]).max() <= 1e-12 * np.abs(ref[..., dim:]).max()
    assert np.abs(ref[..., dim:]).max() > 1e-4       # the field is not trivial
    h.close()


def test_postprocess_rejects_an_inverted_displaced_cell(libs):
    capi, orc = libs
    prob = make_problem(nl_params(poly_degree=1), 2, reps=[2, 2])
    h = capi.Handle(prob)
    X = prob.mesh.support_points
    from helpers import dof_components
    comp = dof_components(prob)
    u = np.where(comp == 0, -2.0 * X[:, 0], 0.0)      # F_xx = -1 everywhere
    h.set_vector(capi.NL_TOTAL_DISPLACEMENT, u)
    with pytest.raises(capi.GraftError):
        h.postprocess(capi.NL_TOTAL_DISPLACEMENT)
    h.close()
