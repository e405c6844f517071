"""CUDA path against outputs of THE REFERENCE'S OWN CODE: the element tangent / residual that the
reference's assembly block (nonlinear_elasticity.cc:791-859, 872-1036, cut out and compiled against
oracle/ref_shim in the build container; vectors in tests/golden/reference_vectors.npz) computes for
one cell, against the device assembling the same one-cell problem through the C-ABI
(gf_nl_newton_assemble -> cell kernel, face kernel, scatter). FP64, 1e-12 relative to the largest
entry, as for the oracle parity tests."""
import os

import numpy as np
import pytest

from helpers import nl_params
from dealii_adapter_b200.problem import make_problem

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def libs(native_libs):
    from dealii_adapter_b200 import build, capi
    build.build_cuda()
    capi.lib()
    return capi


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_device_cell_assembly_equals_the_reference_assembly_block(libs, case):
    capi = libs
    ref = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    meta = ref["asm%d_meta" % case]
    dim, degree = int(meta[0]), int(meta[1])
    h, body_force = meta[2:5], meta[5:8]
    mu, nu, rho, beta, dt = meta[8:13]
    assert sorted(ref["asm%d_faces" % case].tolist()) == [0, 1, 3]      # the PF interface faces
    p = nl_params(poly_degree=degree, scenario="PF", mu=mu, nu=nu, rho=rho, beta=beta, delta_t=dt,
                  body_force=tuple(body_force))
    prob = make_problem(p, dim, reps=[1] * dim, numbering="cellwise",
                        box=([0.0] * dim, list(h[:dim])))
    prob.constrained = np.zeros_like(prob.constrained)
    assert np.array_equal(prob.mesh.cell_dofs.reshape(-1), np.arange(prob.n_dofs))
    hd = capi.Handle(prob)
    hd.set_vector(capi.NL_TOTAL_DISPLACEMENT, ref["asm%d_u" % case])
    hd.set_vector(capi.NL_EXTERNAL_STRESS, ref["asm%d_stress" % case])
    # gf_nl_newton_assemble first runs update_acceleration (:592-599):
    # a = alpha_1*0 - alpha_2*v_old - alpha_3*a_old with alpha_3 = (1-2 beta)/(2 beta)
    alpha_3 = (1 - 2 * beta) / (2 * beta)
    hd.set_vector(capi.NL_ACCELERATION_OLD, -ref["asm%d_acc" % case] / alpha_3)
    hd.nl_begin_step()
    hd.nl_newton_assemble()
    assert np.abs(hd.get_vector(capi.NL_ACCELERATION) - ref["asm%d_acc" % case]).max() \
        <= 1e-14 * np.abs(ref["asm%d_acc" % case]).max()
    rowptr, col, val = hd.export_csr(capi.MAT_TANGENT)
    import scipy.sparse as sp
    K = sp.csr_matrix((val, col, rowptr), shape=(prob.n_dofs, prob.n_dofs)).toarray()
    r = hd.get_vector(capi.NL_SYSTEM_RHS)
    K_ref, r_ref = ref["asm%d_K" % case], ref["asm%d_r" % case]
    assert np.abs(K - K_ref).max() <= 1e-12 * np.abs(K_ref).max()
    assert np.abs(r - r_ref).max() <= 1e-12 * np.abs(r_ref).max()
    hd.close()


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_device_linear_stiffness_and_loading_equal_the_reference_loops(libs, case):
    """linear_elasticity.cc:289-323 (local stiffness) and :487-512 (consistent loading): the
    reference's own statements on one cell against gf_lin_assemble_once / gf_lin_step."""
    capi = libs
    from helpers import lin_params
    ref = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    meta = ref["lin%d_meta" % case]
    dim, degree, h, mu, nu = int(meta[0]), int(meta[1]), meta[2:5], meta[5], meta[6]
    p = lin_params(poly_degree=degree, scenario="PF", mu=mu, nu=nu, type_lin="CG")
    prob = make_problem(p, dim, reps=[1] * dim, numbering="cellwise",
                        box=([0.0] * dim, list(h[:dim])))
    prob.constrained = np.zeros_like(prob.constrained)
    assert sorted(prob.iface_face_no.tolist()) == [0, 1, 3]
    hd = capi.Handle(prob)
    hd.lin_assemble_once()
    rowptr, col, val = hd.export_csr(capi.MAT_STIFFNESS)
    import scipy.sparse as sp
    K = sp.csr_matrix((val, col, rowptr), shape=(prob.n_dofs, prob.n_dofs)).toarray()
    K_ref, F_ref = ref["lin%d_K" % case], ref["lin%d_F" % case]
    assert np.abs(K - K_ref).max() <= 1e-12 * np.abs(K_ref).max()
    hd.set_vector(capi.LIN_STRESS, ref["lin%d_stress" % case])
    hd.lin_step(0, 10.0)                      # old_stress <- the consistent loading (:405-409)
    F = hd.get_vector(capi.LIN_OLD_STRESS)
    assert np.abs(F - F_ref).max() <= 1e-12 * np.abs(F_ref).max()
    hd.close()


@pytest.mark.parametrize("case", [0, 1])
def test_device_newmark_updates_equal_the_reference_members(libs, case):
    """gf_nl_end_step against the reference's own update_acceleration / update_velocity /
    update_old_variables lines (nonlinear_elasticity.cc:592-622, coefficients .h:242-250)."""
    capi = libs
    ref = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    beta, gamma, dt = ref["upd%d_in" % case]
    prob = make_problem(nl_params(poly_degree=1, beta=beta, gamma=gamma, delta_t=dt), 2, reps=[3, 2])
    delta, total, v_old, a_old, rhs, upd = ref["upd%d_vecs" % case]
    hd = capi.Handle(prob)
    hd.set_vector(capi.NL_SOLUTION_DELTA, delta)
    hd.set_vector(capi.NL_TOTAL_DISPLACEMENT, total)
    hd.set_vector(capi.NL_VELOCITY_OLD, v_old)
    hd.set_vector(capi.NL_ACCELERATION_OLD, a_old)
    hd.nl_end_step()          # :139-144
    for which, name in ((capi.NL_ACCELERATION, "acc"), (capi.NL_VELOCITY, "vel"),
                        (capi.NL_TOTAL_DISPLACEMENT, "total")):
        w = ref["upd%d_%s" % (case, name)]
        assert np.abs(hd.get_vector(which) - w).max() <= 1e-14 * np.abs(w).max(), name
    assert np.array_equal(hd.get_vector(capi.NL_VELOCITY_OLD), hd.get_vector(capi.NL_VELOCITY))
    assert np.array_equal(hd.get_vector(capi.NL_ACCELERATION_OLD), hd.get_vector(capi.NL_ACCELERATION))
    hd.close()


@pytest.mark.parametrize("case", [0, 1])
def test_device_theta_scheme_rhs_equals_the_reference_block(libs, case):
    """gf_lin_step's right-hand side against the reference's own assemble_rhs algebra
    (linear_elasticity.cc:384-420; 'Force' data + body force, and 'Stress' data)."""
    capi = libs
    from helpers import lin_params
    ref = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    theta, dt, consistent = ref["rhs%d_in" % case][:3]
    bf = tuple(ref["rhs%d_in" % case][3:6])
    p = lin_params(poly_degree=1, theta=theta, delta_t=dt, type_lin="CG",
                   read_data_name="Stress" if consistent else "Force", body_force=bf)
    prob = make_problem(p, 2, reps=[3, 2])
    loading, stress, old_stress, vel, disp, bfv, new_vel = ref["rhs%d_vecs" % case]
    hd = capi.Handle(prob)
    hd.lin_assemble_once()
    hd.set_vector(capi.LIN_STRESS, stress)
    hd.set_vector(capi.LIN_OLD_STRESS, old_stress)
    hd.set_vector(capi.LIN_VELOCITY, vel)
    hd.set_vector(capi.LIN_DISPLACEMENT, disp)
    hd.lin_step(0, 10.0)
    free = prob.constrained == 0
    want = ref["rhs%d_system_rhs" % case]
    assert np.abs(hd.get_vector(capi.LIN_SYSTEM_RHS) - want)[free].max() <= 1e-11 * np.abs(want).max()
    for name, which in (("old_stress", capi.LIN_OLD_STRESS), ("old_velocity", capi.LIN_OLD_VELOCITY),
                        ("old_displacement", capi.LIN_OLD_DISPLACEMENT)):
        w = ref["rhs%d_%s" % (case, name)]
        assert np.abs(hd.get_vector(which) - w).max() <= 1e-12 * np.abs(w).max(), name
    hd.close()


@pytest.mark.parametrize("case", [0, 1])
def test_device_adapter_bodies_equal_the_reference_members(libs, case):
    """gf_get_interface_displacement / gf_set_traction against the reference's own
    format_deal_to_precice / format_precice_to_deal (adapter.h:389-443), bit for bit."""
    capi = libs
    ref = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    meta = ref["adp%d_meta" % case]
    dim, degree, reps = int(meta[0]), int(meta[1]), [int(x) for x in meta[2:]]
    prob = make_problem(nl_params(poly_degree=degree), dim, reps=reps, numbering="component_wise")
    vec, buf = ref["adp%d_vec" % case], ref["adp%d_buf" % case]
    hd = capi.Handle(prob)
    hd.set_vector(capi.NL_TOTAL_DISPLACEMENT, vec)
    assert np.array_equal(hd.get_interface_displacement(), ref["adp%d_write" % case])
    hd.set_vector(capi.NL_EXTERNAL_STRESS, vec)
    hd.set_traction(buf)
    assert np.array_equal(hd.get_vector(capi.NL_EXTERNAL_STRESS), ref["adp%d_after_read" % case])
    hd.close()
