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


# A coarsest level larger than this is semi-coarsened further (only the directions with an even
# repetition count are halved). Every rank solves the coarsest level redundantly, and beyond ~2 k
# nodes its Chebyshev(80) no longer fits the single-launch solver's shared-memory staging: on the
# 8-GPU weak-scaling mesh (24x288x96 -> 3x36x12 = 12,775 nodes) it cost 1.9 ms per V-cycle.
SEMI_COARSEN_ABOVE_NODES = 3000


def coarsen_problem(problem, allow_semi=True):
    """The next coarser level of `problem`: every direction halved (what refine_global undoes) while
    all repetition counts are even; else, if the level is still large, the even directions only
    (semi-coarsening); None when nothing can or needs to be coarsened."""
    mesh = problem.mesh
    even = [r % 2 == 0 and r >= 2 for r in mesh.reps]
    if all(even):
        reps = [r // 2 for r in mesh.reps]
    elif allow_semi and any(even) and mesh.n_nodes > SEMI_COARSEN_ABOVE_NODES:
        reps = [r // 2 if e else r for r, e in zip(mesh.reps, even)]
    else:
        return None
    return make_problem(problem.params, problem.dim, reps=reps, numbering=mesh.numbering,
                        box=(mesh.p0, mesh.p1))


def child_table(coarse_mesh, fine_mesh):
    """[n_coarse_cells, 2^dim] GLOBAL fine cell index of child k = kx + 2 ky + 4 kz (cells are
    lexicographic); -1 where the level pair does not refine that direction."""
    dim = coarse_mesh.dim
    rc = list(coarse_mesh.reps) + [1] * (3 - dim)
    rf = list(fine_mesh.reps) + [1] * (3 - dim)
    fac = [f // c for f, c in zip(rf, rc)]                   # 2 (refined) or 1 per direction
    assert all(x in (1, 2) for x in fac) and [c * x for c, x in zip(rc, fac)] == rf
    k, j, i = np.meshgrid(np.arange(rc[2]), np.arange(rc[1]), np.arange(rc[0]), indexing="ij")
    i, j, k = i.reshape(-1), j.reshape(-1), k.reshape(-1)   # coarse cell index = (k*ry + j)*rx + i
    out = -np.ones((len(i), 1 << dim), dtype=np.int64)
    for c in range(1 << dim):
        bits = [c & 1, (c >> 1) & 1, (c >> 2) & 1]
        if any(b and fac[d] == 1 for d, b in enumerate(bits)):
            continue
        fi, fj, fk = fac[0] * i + bits[0], fac[1] * j + bits[1], fac[2] * k + bits[2]
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
                sel = tab[pc.local_cell_global] if pc is not None else tab
                tab = np.where(sel >= 0, g2l[np.maximum(sel, 0)], -1)
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
