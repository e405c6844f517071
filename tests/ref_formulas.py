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


def lagrange_1d(p, x):
    """values and derivatives (n_pts, p+1) of the equidistant Lagrange basis on [0,1]."""
    nodes = np.linspace(0.0, 1.0, p + 1)
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
    """deal.II FE_Q local node order as lexicographic (lx,ly,lz): vertices, lines, quads, hex."""
    assert p in (1, 2)
    nodes = []
    for v in range(1 << dim):
        nodes.append(tuple(((v >> d) & 1) * p for d in range(dim)))
    if p == 2:
        if dim == 2:
            nodes += [(0, 1), (2, 1), (1, 0), (1, 2), (1, 1)]
        else:
            for z in (0, 2):
                nodes += [(0, 1, z), (2, 1, z), (1, 0, z), (1, 2, z)]
            nodes += [(0, 0, 1), (2, 0, 1), (0, 2, 1), (2, 2, 1)]
            nodes += [(0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2), (1, 1, 1)]
    return nodes


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
    u_local in FESystem order (node-major, component-minor)."""
    N, dN, w = cell_tables(dim, p, nq1)
    u = np.asarray(u_local).reshape(-1, dim)            # [a, c]
    grad = dN / np.asarray(h)[None, None, :]            # real-space gradients
    H = np.einsum("ac,qad->qcd", u, grad)               # H[q,c,d] = du_c/dX_d
    vol = np.prod(h)
    return sum(psi(np.eye(dim) + H[q], mu, nu) * w[q] * vol for q in range(len(w)))
