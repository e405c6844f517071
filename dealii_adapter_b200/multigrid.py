"""Host side of the geometric multigrid preconditioner (C-ABI: gf_mg_attach).

In the reference the mesh is `subdivided_hyper_rectangle` followed by `refine_global(n)`
(nonlinear_elasticity.cc:237-246, linear_elasticity.cc:143-151), so the host owns a refinement
hierarchy: level cells and cell->child(k). This module builds the same information for the
structured stand-in mesh: one Problem per level (the same box with halved repetitions, same
boundary roles) and, per level pair, the child table in deal.II child order (k = kx + 2 ky + 4 kz).
"""
import numpy as np

from . import capi
from .problem import make_problem


def coarsen_problem(problem):
    """The next coarser level of `problem` (every direction halved) or None if a repetition is odd."""
    mesh = problem.mesh
    if any(r % 2 for r in mesh.reps) or min(mesh.reps) < 2:
        return None
    reps = [r // 2 for r in mesh.reps]
    return make_problem(problem.params, problem.dim, reps=reps, numbering=mesh.numbering,
                        box=(mesh.p0, mesh.p1))


def child_table(coarse_mesh, fine_mesh):
    """[n_coarse_cells, 2^dim] GLOBAL fine cell index of child k (cells are lexicographic)."""
    dim = coarse_mesh.dim
    rc = list(coarse_mesh.reps) + [1] * (3 - dim)
    rf = list(fine_mesh.reps) + [1] * (3 - dim)
    k, j, i = np.meshgrid(np.arange(rc[2]), np.arange(rc[1]), np.arange(rc[0]), indexing="ij")
    i, j, k = i.reshape(-1), j.reshape(-1), k.reshape(-1)   # coarse cell index = (k*ry + j)*rx + i
    out = np.zeros((len(i), 1 << dim), dtype=np.int64)
    for c in range(1 << dim):
        fi, fj = 2 * i + (c & 1), 2 * j + ((c >> 1) & 1)
        fk = 2 * k + ((c >> 2) & 1) if dim == 3 else k
        out[:, c] = (fk * rf[1] + fj) * rf[0] + fi
    return out


def plan_levels(problem, n_levels=None, world=1, axis=1, min_cells=1, replicate_below_dofs=80000):
    """Levels of the hierarchy below `problem` and, for a partitioned run, which of them are
    replicated on every rank instead of slab-partitioned: ([problems], [replicated flags]).
    A level can stay partitioned only while the slabs of consecutive levels coincide (the finer
    level's layers per rank are even); once a level is replicated, all coarser ones are."""
    problems, replicated = [problem], [False]
    while n_levels is None or len(problems) < n_levels:
        nxt = coarsen_problem(problems[-1])
        if nxt is None or nxt.mesh.n_cells < min_cells:
            break
        if world > 1:
            aligned = (not replicated[-1]) and problems[-1].mesh.reps[axis] % (2 * world) == 0
            replicated.append(replicated[-1] or not aligned or nxt.n_dofs <= replicate_below_dofs)
        else:
            replicated.append(False)
        problems.append(nxt)
    return problems, replicated


class Hierarchy:
    """Handles of all levels (levels[0] = finest), linked with gf_mg_attach.

    Partitioned run (world > 1): the fine levels are slab-partitioned with the same (axis, world,
    rank) — their slab boundaries must coincide, i.e. the layers per rank stay even — and the
    remaining small levels are REPLICATED: every rank holds the whole coarse mesh as a serial
    handle, so the coarse smoothing / coarsest solve need no halo exchange (the library
    all-reduces the restricted residual). A level is replicated when it has at most
    `replicate_below_dofs` DoFs or when its slabs could not be aligned any more."""

    def __init__(self, problem, device=0, n_levels=None, world=1, rank=0, comm=None, axis=1,
                 min_cells=1, replicate_below_dofs=80000):
        self.problems, self.replicated = plan_levels(problem, n_levels, world, axis, min_cells,
                                                     replicate_below_dofs)
        self.partitions = [p.mesh.partition(axis, world, rank) if (world > 1 and not rep) else None
                           for p, rep in zip(self.problems, self.replicated)]
        # slab_axis on every level and for every world size (1 included): all sums then run over
        # the same partition-independent tree, so the run is bitwise the same for any `world`
        self.handles = [capi.Handle(p, device=device, partition=part,
                                    comm=comm if part is not None else None, slab_axis=axis)
                        for p, part in zip(self.problems, self.partitions)]
        for l in range(len(self.handles) - 1):
            fine, coarse = self.problems[l], self.problems[l + 1]
            tab = child_table(coarse.mesh, fine.mesh)
            pf, pc = self.partitions[l], self.partitions[l + 1]
            if pf is not None:
                g2l = -np.ones(fine.mesh.n_cells, dtype=np.int64)
                g2l[pf.local_cell_global] = np.arange(pf.n_local_cells)
                # replicated coarse level: all coarse cells, children that are not local -> -1
                tab = g2l[tab[pc.local_cell_global]] if pc is not None else g2l[tab]
                covered = np.zeros(pf.n_local_cells, dtype=bool)
                covered[tab[tab >= 0]] = True
                if not covered.all():
                    raise ValueError("slab boundaries of multigrid levels %d and %d do not coincide"
                                     % (l, l + 1))
            self.handles[l].mg_attach(self.handles[l + 1], tab)
        self.fine = self.handles[0]
        self.fine.set_option(capi.OPT_PRECONDITIONER, capi.PRECOND_MULTIGRID)

    @property
    def n_levels(self):
        return len(self.handles)

    def close(self):
        for h in self.handles:
            h.close()
