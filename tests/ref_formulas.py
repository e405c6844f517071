"""Independent numpy transcription of the reference's closed-form pieces, used to pin the C++ oracle
(the reference has no golden vectors; deal.II cannot run here).

 * material: /root/reference/source/nonlinear_elasticity/include/compressible_neo_hook_material.h
   :20-21 (kappa, c_1), :62-72 (Psi), :75-98 (tau), :101-138 (Jc), written with FULL 4th-order
   tensors (I, S, IxI, dev_P as dense dim^4 arrays) instead of the oracle's Voigt storage;
 * element energy  Pi(u) = sum_q Psi(F_q) JxW  on one Cartesian cell, with its own Lagrange /
   Gauss tables, so that -dPi/du and d2Pi/du2 (finite differences) check the oracle's residual and
   tangent (nonlinear_elasticity.cc:984-985, :1011-1023) against the reference's own Psi.
"""
import numpy as np


# ------------------------------------------------------------------------------------------------
# dense standard tensors (Physics::Elasticity::StandardTensors)
# ------------------------------------------------------------------------------------------------
def standard_tensors(dim):
    I = np.eye(dim)
    IxI = np.einsum("ij,kl->ijkl", I, I)
    S = 0.5 * (np.einsum("ik,jl->ijkl", I, I) + np.einsum("il,jk->ijkl", I, I))
    dev_P = S - IxI / dim
    return I, S, IxI, dev_P


def kappa_c1(mu, nu):
    return (2.0 * mu * (1.0 + nu)) / (3.0 * (1.0 - 2.0 * nu)), mu / 2.0


def psi(F, mu, nu):
    dim = F.shape[0]
    kappa, c1 = kappa_c1(mu, nu)
    J = np.linalg.det(F)
    Fb = J ** (-1.0 / dim) * F
    bb = Fb @ Fb.T
    return (kappa / 4.0) * (J * J - 1.0 - 2.0 * np.log(J)) + c1 * (np.trace(bb) - dim)


def tau_Jc(F, mu, nu):
    """Kirchhoff stress (dim,dim) and spatial tangent J*c as dense (dim,)*4 array."""
    dim = F.shape[0]
    kappa, c1 = kappa_c1(mu, nu)
    I, S, IxI, dev_P = standard_tensors(dim)
    J = np.linalg.det(F)
    Fb = J ** (-1.0 / dim) * F
    b_bar = 0.5 * (Fb @ Fb.T + (Fb @ Fb.T).T)
    dPsi = (kappa / 2.0) * (J - 1.0 / J)
    d2Psi = (kappa / 2.0) * (1.0 + 1.0 / (J * J))
    tau_vol = dPsi * J * I
    tau_bar = 2.0 * c1 * b_bar
    tau_iso = np.einsum("ijkl,kl->ij", dev_P, tau_bar)
    Jc_vol = J * ((dPsi + J * d2Psi) * IxI - (2.0 * dPsi) * S)
    Jc_iso = ((2.0 / dim) * np.trace(tau_bar) * dev_P
              - (2.0 / dim) * (np.einsum("ij,kl->ijkl", tau_iso, I) + np.einsum("ij,kl->ijkl", I, tau_iso)))
    return tau_vol + tau_iso, Jc_vol + Jc_iso


def sym_index(dim):
    """deal.II SymmetricTensor<2,dim> component order."""
    return [(0, 0), (1, 1), (0, 1)] if dim == 2 else [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]


def to_voigt2(t):
    return np.array([t[i, j] for (i, j) in sym_index(t.shape[0])])


def to_voigt4(t):
    idx = sym_index(t.shape[0])
    return np.array([[t[i, j, k, l] for (k, l) in idx] for (i, j) in idx])


# ------------------------------------------------------------------------------------------------
# FE_Q(p) on the unit cell in deal.II hierarchical order, QGauss(n)
# ------------------------------------------------------------------------------------------------
def gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def support_points_1d(p):
    """FE_Q(p) support points on [0,1]: Gauss-Lobatto (deal.II: QGaussLobatto(p+1)); for p <= 2 these
    are the equidistant points."""
    if p <= 2:
        return np.linspace(0.0, 1.0, p + 1)
    inner = np.polynomial.legendre.Legendre.basis(p).deriv().roots()
    return np.concatenate([[0.0], 0.5 * (np.sort(inner.real) + 1.0), [1.0]])


def lagrange_1d(p, x):
    """values and derivatives (n_pts, p+1) of the Lagrange basis on the FE_Q support points."""
    nodes = support_points_1d(p)
    x = np.atleast_1d(x)
    val = np.ones((len(x), p + 1))
    der = np.zeros((len(x), p + 1))
    for i in range(p + 1):
        for j in range(p + 1):
            if j != i:
                val[:, i] *= (x - nodes[j]) / (nodes[i] - nodes[j])
        for k in range(p + 1):
            if k == i:
                continue
            term = np.ones(len(x)) / (nodes[i] - nodes[k])
            for j in range(p + 1):
                if j != i and j != k:
                    term *= (x - nodes[j]) / (nodes[i] - nodes[j])
            der[:, i] += term
    return val, der


def hierarchical_nodes(dim, p):
    """deal.II FE_Q local node order as lexicographic (lx,ly,lz): vertices, lines, quads, hex, written
    with the index formulas of FETools::hierarchic_to_lexicographic_numbering (n = p + 1 points per
    direction, lexicographic index = lx + n ly + n^2 lz)."""
    n, m = p + 1, p - 1
    h2l = []
    if dim == 2:
        h2l += [0, n - 1, n * (n - 1), n * n - 1]
        h2l += [(1 + i) * n for i in range(m)]                  # line 0: x = 0
        h2l += [(2 + i) * n - 1 for i in range(m)]              # line 1: x = 1
        h2l += [1 + i for i in range(m)]                        # line 2: y = 0
        h2l += [n * (n - 1) + i + 1 for i in range(m)]          # line 3: y = 1
        h2l += [n * (i + 1) + j + 1 for i in range(m) for j in range(m)]
    else:
        h2l += [0, n - 1, n * (n - 1), n * n - 1]
        h2l += [x + n * n * (n - 1) for x in h2l[:4]]
        for off in (0, n * n * (n - 1)):                        # bottom, top face lines
            h2l += [off + (1 + i) * n for i in range(m)]
            h2l += [off + n - 1 + (i + 1) * n for i in range(m)]
            h2l += [off + 1 + i for i in range(m)]
            h2l += [off + 1 + i + n * (n - 1) for i in range(m)]
        h2l += [(1 + i) * n * n for i in range(m)]              # lines in z direction
        h2l += [n - 1 + (i + 1) * n * n for i in range(m)]
        h2l += [n * (n - 1) + (i + 1) * n * n for i in range(m)]
        h2l += [n * n - 1 + (i + 1) * n * n for i in range(m)]
        h2l += [(i + 1) * n * n + n * (j + 1) for i in range(m) for j in range(m)]            # x = 0
        h2l += [(i + 1) * n * n + n - 1 + n * (j + 1) for i in range(m) for j in range(m)]    # x = 1
        h2l += [(j + 1) * n * n + i + 1 for i in range(m) for j in range(m)]                  # y = 0
        h2l += [(j + 1) * n * n + i + 1 + n * (n - 1) for i in range(m) for j in range(m)]    # y = 1
        h2l += [n * (i + 1) + j + 1 for i in range(m) for j in range(m)]                      # z = 0
        h2l += [n * n * (n - 1) + n * (i + 1) + j + 1 for i in range(m) for j in range(m)]    # z = 1
        h2l += [n * n * (i + 1) + n * (j + 1) + k + 1
                for i in range(m) for j in range(m) for k in range(m)]
    assert sorted(h2l) == list(range(n ** dim))
    return [tuple((l // n ** d) % n for d in range(dim)) for l in h2l]


def system_to_node_component(dim, p):
    """FESystem(FE_Q(p), dim) local DoF -> (hierarchical scalar node, component): entity by entity
    (vertices, lines, quads, hex), inside an entity component by component
    (FESystem::build_cell_tables)."""
    counts = [1] * (1 << dim) + [p - 1] * (4 if dim == 2 else 12) + \
        [(p - 1) ** 2] * (1 if dim == 2 else 6) + ([(p - 1) ** 3] if dim == 3 else [])
    out, first = [], 0
    for cnt in counts:
        out += [(first + k, c) for c in range(dim) for k in range(cnt)]
        first += cnt
    return out


def cell_tables(dim, p, nq1):
    """N[q,a], dN[q,a,d] (unit-cell gradients), w[q]; q lexicographic with x fastest."""
    x1, w1 = gauss01(nq1)
    v1, d1 = lagrange_1d(p, x1)
    nodes = hierarchical_nodes(dim, p)
    nq = nq1 ** dim
    N = np.zeros((nq, len(nodes)))
    dN = np.zeros((nq, len(nodes), dim))
    w = np.zeros(nq)
    for q in range(nq):
        qi = [(q // nq1 ** d) % nq1 for d in range(dim)]
        w[q] = np.prod([w1[i] for i in qi])
        for a, lex in enumerate(nodes):
            N[q, a] = np.prod([v1[qi[d], lex[d]] for d in range(dim)])
            for k in range(dim):
                N_k = 1.0
                for d in range(dim):
                    N_k *= d1[qi[d], lex[d]] if d == k else v1[qi[d], lex[d]]
                dN[q, a, k] = N_k
    return N, dN, w


def element_energy(u_local, h, dim, p, nq1, mu, nu):
    """Pi(u) = sum_q Psi(I + grad u) JxW on a Cartesian cell with edge lengths h.
    u_local in FESystem local order (system_to_node_component)."""
    N, dN, w = cell_tables(dim, p, nq1)
    u = np.zeros((N.shape[1], dim))                     # [a, c]
    for i, (a, c) in enumerate(system_to_node_component(dim, p)):
        u[a, c] = u_local[i]
    grad = dN / np.asarray(h)[None, None, :]            # real-space gradients
    H = np.einsum("ac,qad->qcd", u, grad)               # H[q,c,d] = du_c/dX_d
    vol = np.prod(h)
    return sum(psi(np.eye(dim) + H[q], mu, nu) * w[q] * vol for q in range(len(w)))
