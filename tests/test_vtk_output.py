"""Host side of output_results (nonlinear_elasticity.cc:1215-1254, linear_elasticity.cc:590-629):
the C++ writer host/vtk_output.cc fed with the oracle's patch fields. Checks the VTK Lagrange
connectivity against the published VTK node ordering, the file layout of deal.II's write_vtk with
write_higher_order_cells, the `curved_boundary` point rule and the data round trip."""
import numpy as np
import pytest

from helpers import dof_components, nl_params
from dealii_adapter_b200.mesh import host_lib
from dealii_adapter_b200.problem import make_problem


def vtk_index(dim, p, i, j, k=0, legacy=True):
    return host_lib().gfh_vtk_point_index_from_ijk(dim, p, i, j, k, 1 if legacy else 0)


def test_vtk_lagrange_node_order():
    # biquadratic quadrilateral (VTK_LAGRANGE_QUADRILATERAL order 2 == VTK_BIQUADRATIC_QUAD):
    # corners counter-clockwise, edge mid nodes bottom/right/top/left, centre
    quad9 = [(0, 0), (2, 0), (2, 2), (0, 2), (1, 0), (2, 1), (1, 2), (0, 1), (1, 1)]
    assert [vtk_index(2, 2, i, j) for i, j in quad9] == list(range(9))
    # triquadratic hexahedron, legacy edge order (files of version < 5, what write_vtk emits)
    hex27 = [(0, 0, 0), (2, 0, 0), (2, 2, 0), (0, 2, 0), (0, 0, 2), (2, 0, 2), (2, 2, 2), (0, 2, 2),
             (1, 0, 0), (2, 1, 0), (1, 2, 0), (0, 1, 0), (1, 0, 2), (2, 1, 2), (1, 2, 2), (0, 1, 2),
             (0, 0, 1), (2, 0, 1), (0, 2, 1), (2, 2, 1),
             (0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2), (1, 1, 1)]
    assert [vtk_index(3, 2, *ijk) for ijk in hex27] == list(range(27))
    # the non-legacy order only swaps the last two vertical edges
    assert vtk_index(3, 2, 2, 2, 1, legacy=False) == 18 and vtk_index(3, 2, 0, 2, 1, legacy=False) == 19
    # always a permutation, also for order 1 and 3
    for dim, p in ((2, 1), (2, 3), (3, 1), (3, 3)):
        n1 = p + 1
        idx = sorted(vtk_index(dim, p, i, j, k) for k in range(n1 if dim == 3 else 1)
                     for j in range(n1) for i in range(n1))
        assert idx == list(range(n1 ** dim))
    assert [vtk_index(3, 1, *v) for v in [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)]] == [0, 1, 2, 3]


def parse_vtk(path):
    tok = open(path).read().split("\n")
    assert tok[0] == "# vtk DataFile Version 3.0" and tok[2] == "ASCII"
    assert tok[3] == "DATASET UNSTRUCTURED_GRID"
    words = " ".join(tok[4:]).split()
    out, pos = {}, 0
    while pos < len(words):
        w = words[pos]
        if w == "POINTS":
            n = int(words[pos + 1])
            out["points"] = np.array(words[pos + 3:pos + 3 + 3 * n], dtype=float).reshape(n, 3)
            pos += 3 + 3 * n
        elif w == "CELLS":
            nc, tot = int(words[pos + 1]), int(words[pos + 2])
            out["cells"] = np.array(words[pos + 3:pos + 3 + tot], dtype=int).reshape(nc, tot // nc)
            pos += 3 + tot
        elif w == "CELL_TYPES":
            nc = int(words[pos + 1])
            out["types"] = np.array(words[pos + 2:pos + 2 + nc], dtype=int)
            pos += 2 + nc
        elif w == "POINT_DATA":
            out["n_point_data"] = int(words[pos + 1])
            pos += 2
        elif w == "VECTORS":
            n = out["n_point_data"]
            out["vec_" + words[pos + 1]] = np.array(words[pos + 3:pos + 3 + 3 * n], dtype=float).reshape(n, 3)
            pos += 3 + 3 * n
        elif w == "SCALARS":
            n = out["n_point_data"]
            assert words[pos + 4:pos + 6] == ["LOOKUP_TABLE", "default"]
            out.setdefault("scalars", []).append(words[pos + 1])
            out["sc_" + words[pos + 1]] = np.array(words[pos + 6:pos + 6 + n], dtype=float)
            pos += 6 + n
        else:
            raise AssertionError("unexpected token %r" % w)
    return out


@pytest.mark.parametrize("dim,degree,reps", [(2, 2, [4, 3]), (3, 2, [3, 3, 3]), (3, 1, [3, 4, 3]),
                                           (2, 3, [4, 3]), (2, 4, [3, 3]), (3, 3, [3, 3, 3])])
def test_written_file_round_trips_the_oracle_fields(native_libs, tmp_path, dim, degree, reps):
    from oracle import oracle_py as orc
    prob = make_problem(nl_params(poly_degree=degree), dim, reps=reps)
    comp = dof_components(prob)
    X = prob.mesh.support_points
    u = 0.02 * np.sin(3.0 * X[:, 0] + 2.0 * X[:, 1] + comp) * X[:, 1]
    o = orc.Oracle(prob)
    o.set(orc.NL_TOTAL_DISPLACEMENT, u)
    pts, fld = o.postprocess(orc.NL_TOTAL_DISPLACEMENT)
    name = tmp_path / "solution-000.vtk"
    prob.mesh.write_vtk(name, fld)
    got = parse_vtk(name)
    nc, npts = prob.mesh.n_cells, (degree + 1) ** dim
    assert got["points"].shape == (nc * npts, 3) and got["n_point_data"] == nc * npts
    assert got["cells"].shape == (nc, npts + 1) and np.all(got["cells"][:, 0] == npts)
    assert np.all(got["types"] == (70 if dim == 2 else 72))
    # every patch owns its points: the connectivity of cell c is a permutation of its block
    conn = got["cells"][:, 1:]
    assert np.array_equal(np.sort(conn, axis=1), np.arange(nc * npts).reshape(nc, npts))
    # first connectivity entries are the corners in VTK order = lexicographic (0,0),(p,0),(p,p),(0,p)
    n1 = degree + 1
    first4 = [0, degree, degree + degree * n1, degree * n1]
    assert np.array_equal(conn[:, :4] - np.arange(nc)[:, None] * npts, np.tile(first4, (nc, 1)))
    # data: displacement vectors (z padded in 2D) and the dim*dim strain scalars, 12 digits
    assert np.allclose(got["vec_displacement"][:, :dim], fld[..., :dim].reshape(-1, dim), rtol=1e-11, atol=1e-14)
    if dim == 2:
        assert np.all(got["vec_displacement"][:, 2] == 0) and np.all(got["points"][:, 2] == 0)
    names = ["strain_" + a + b for a in "xyz"[:dim] for b in "xyz"[:dim]]
    assert got["scalars"] == names
    for q, nm in enumerate(names):
        assert np.allclose(got["sc_" + nm], fld[..., dim + q].reshape(-1), rtol=1e-11, atol=1e-14)
    # points: boundary cells follow the Eulerian mapping everywhere, interior cells only at their
    # vertices (curved_boundary); interior mid points are the multilinear blend of the corners
    P = got["points"][:, :dim].reshape(nc, npts, dim)
    at = prob.mesh.cells_at_boundary().astype(bool)
    assert at.any() and (~at).any()
    assert np.allclose(P[at], pts[at], rtol=1e-11, atol=1e-14)
    corner = [i + n1 * j + n1 * n1 * k for k in ((0, degree) if dim == 3 else (0,))
              for j in (0, degree) for i in (0, degree)]
    assert np.allclose(P[~at][:, corner], pts[~at][:, corner], rtol=1e-11, atol=1e-14)
    centre = (npts - 1) // 2 if degree == 2 else None
    if centre is not None:
        assert np.allclose(P[~at][:, centre], P[~at][:, corner].mean(axis=1), rtol=1e-11, atol=1e-14)
        assert not np.allclose(P[~at][:, centre], pts[~at][:, centre], rtol=1e-9, atol=0)
