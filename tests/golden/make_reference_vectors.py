"""Golden vectors produced by THE REFERENCE'S OWN CODE, run in this container.

Three header-only pieces of /root/reference compile without the real deal.II against the small
stand-in oracle/ref_shim (`make -C oracle ref` -> oracle/_ref/ref_driver, sources taken in place):
  source/nonlinear_elasticity/include/compressible_neo_hook_material.h  (Psi, tau, Jc)
  source/nonlinear_elasticity/include/postprocessor.h                   (evaluate_vector_field)
  include/adapter/time_handler.h                                        (Time)
This script runs them on deterministic inputs and writes tests/golden/reference_vectors.npz;
tests/test_reference_pins.py checks the oracle (and the host Time mirror) against the file, and —
where /root/reference exists — that the file is what the reference produces.

usage: python tests/golden/make_reference_vectors.py [--check]
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
OUT = os.path.join(HERE, "reference_vectors.npz")


def build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"],
                          stdout=subprocess.DEVNULL)


def run(args, stdin=None):
    return subprocess.run([DRIVER] + [repr(float(a)) if isinstance(a, float) else str(a) for a in args],
                          input=stdin, capture_output=True, text=True, check=True).stdout


def material_cases():
    rng = np.random.RandomState(20260117)
    cases = []
    for dim in (2, 3):
        n = dim * (dim + 1) // 2
        for mu, nu in ((0.5e6, 0.4), (1538462.0, 0.3), (2.0e6, 0.0), (1.0e5, 0.49)):
            for _ in range(6):
                # b_bar = F_bar F_bar^T of a random deformation gradient (SPD, det 1), det F apart
                F = np.eye(dim) + 0.35 * rng.uniform(-1, 1, (dim, dim))
                if np.linalg.det(F) <= 0.2:
                    F = np.eye(dim) + 0.1 * rng.uniform(-1, 1, (dim, dim))
                J = np.linalg.det(F)
                Fb = J ** (-1.0 / dim) * F
                b = Fb @ Fb.T
                idx = [(0, 0), (1, 1), (0, 1)] if dim == 2 else \
                    [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]
                bv = [float(b[i, j]) for i, j in idx]
                cases.append((dim, mu, nu, float(J), bv))
    return cases


def generate():
    out = {}
    # ---- material -------------------------------------------------------------------------
    rows = []
    for k, (dim, mu, nu, J, bv) in enumerate(material_cases()):
        n = dim * (dim + 1) // 2
        txt = run(["material", dim, mu, nu, 1000.0, J] + bv).split("\n")
        psi = float(txt[0])
        tau = np.array(txt[1].split(), dtype=float)
        Jc = np.array([l.split() for l in txt[2:2 + n]], dtype=float)
        asym = float(txt[2 + n])
        assert tau.shape == (n,) and Jc.shape == (n, n)
        out["mat%02d_in" % k] = np.array([dim, mu, nu, J] + bv)
        out["mat%02d_psi" % k] = np.array(psi)
        out["mat%02d_tau" % k] = tau
        out["mat%02d_Jc" % k] = Jc
        out["mat%02d_asym" % k] = np.array(asym)
        rows.append(k)
    out["n_material"] = np.array(len(rows))
    # ---- Postprocessor: u and grad_x u = A (I + A)^-1 of the linear field u = A X ------------
    rng = np.random.RandomState(7)
    for dim in (2, 3):
        A = 0.08 * rng.uniform(-1, 1, (dim, dim))
        g = A @ np.linalg.inv(np.eye(dim) + A)
        X = rng.uniform(0, 1, (5, dim))
        lines = []
        for x in X:
            u = A @ x
            lines.append(" ".join(repr(float(v)) for v in list(u) + list(g.reshape(-1))))
        txt = run(["strain", dim], stdin="\n".join(lines) + "\n").split("\n")
        out["pp%d_A" % dim] = A
        out["pp%d_X" % dim] = X
        out["pp%d_names" % dim] = np.array(txt[0].split())
        out["pp%d_out" % dim] = np.array([l.split() for l in txt[1:1 + len(X)]], dtype=float)
    # ---- Time -----------------------------------------------------------------------------------
    tcases = [(1.0, 0.01, 7, 0.03), (10.0, 0.005, 13, 0.045), (0.5, 0.1, 3, 0.1), (2.0, 0.0025, 400, 0.9975)]
    tout = []
    for t_end, dt, n_inc, t_reset in tcases:
        txt = run(["time", t_end, dt, n_inc, t_reset]).split("\n")
        a, b = txt[0].split(), txt[1].split()
        tout.append([t_end, dt, n_inc, t_reset, float(a[0]), float(a[1]), float(b[0]), float(b[1])])
    out["time_cases"] = np.array(tout)
    return out


if __name__ == "__main__":
    build()
    data = generate()
    if "--check" in sys.argv:
        old = np.load(OUT)
        bad = [k for k in data if not np.array_equal(np.asarray(data[k]), old[k])]
        print("differences:", bad)
        sys.exit(1 if bad else 0)
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, "with", len(data), "arrays")
