// See structured_mesh.h. Host scaffolding standing in for deal.II mesh/DoF setup
// (reference: nonlinear_elasticity.cc:171-380, linear_elasticity.cc:79-244).
#include "structured_mesh.h"

#include "fe_basis.h"

#include <algorithm>
#include <cassert>
#include <map>
#include <stdexcept>

namespace gfh
{
  // deal.II FE_Q hierarchical local node order (vertices, lines, quads, hex; every entity with
  // its interior nodes in lexicographic order) expressed in local lexicographic coordinates
  // 0..p per direction: csrc/fe_basis.h restates FETools::hierarchic_to_lexicographic_numbering
  static void build_local_nodes(int dim, int p, std::vector<std::array<int, 3>> &out)
  {
    std::vector<int> lex;
    gf_fe::local_nodes(dim, p, lex);
    out.clear();
    for (size_t a = 0; a < lex.size() / 3; ++a)
      out.push_back({lex[3 * a], lex[3 * a + 1], lex[3 * a + 2]});
  }

  StructuredMesh::StructuredMesh(int dim_, int degree_, const int *reps_, const double *p0_,
                                 const double *p1_, int numbering_)
    : dim(dim_)
    , degree(degree_)
    , numbering(numbering_)
  {
    if (dim != 2 && dim != 3)
      throw std::invalid_argument("StructuredMesh: dim must be 2 or 3");
    if (degree < 1 || degree > max_degree)
      throw std::invalid_argument("StructuredMesh: polynomial degree must be 1.." +
                                  std::to_string(max_degree));
    for (int d = 0; d < dim; ++d)
      {
        reps[d] = reps_[d];
        p0[d]   = p0_[d];
        p1[d]   = p1_[d];
        if (reps[d] < 1)
          throw std::invalid_argument("StructuredMesh: repetitions must be >= 1");
      }
    const int p = degree;
    nodes_per_cell = 1;
    for (int d = 0; d < dim; ++d)
      nodes_per_cell *= (p + 1);
    dofs_per_cell = nodes_per_cell * dim;
    n_cells       = 1;
    n_nodes       = 1;
    for (int d = 0; d < dim; ++d)
      {
        n_cells *= reps[d];
        grid_nodes[d] = int64_t(reps[d]) * p + 1;
        n_nodes *= grid_nodes[d];
      }
    n_dofs = n_nodes * dim;
    if (n_dofs >= (int64_t(1) << 31))
      throw std::invalid_argument("StructuredMesh: more than 2^31 dofs");
    build_local_nodes(dim, p, local_node_lex);
    {
      std::vector<int> node_of, comp_of;
      gf_fe::system_numbering(dim, p, node_of, comp_of, local_dof_of);
    }
    unit_support_1d = gf_fe::gauss_lobatto01(p + 1); // FE_Q(p): equidistant for p <= 2

    // node numbering
    grid_to_node.assign(n_nodes, -1);
    node_to_grid.assign(n_nodes, -1);
    const int nz = dim == 3 ? reps[2] : 1;
    if (numbering == numbering_lexicographic)
      {
        for (int64_t g = 0; g < n_nodes; ++g)
          grid_to_node[g] = node_to_grid[g] = g;
      }
    else
      {
        int64_t next = 0;
        for (int k = 0; k < nz; ++k)
          for (int j = 0; j < reps[1]; ++j)
            for (int i = 0; i < reps[0]; ++i)
              for (int a = 0; a < nodes_per_cell; ++a)
                {
                  const auto &  l  = local_node_lex[a];
                  const int64_t gx = int64_t(i) * p + l[0], gy = int64_t(j) * p + l[1],
                                gz = int64_t(k) * p + l[2];
                  const int64_t g  = (gz * grid_nodes[1] + gy) * grid_nodes[0] + gx;
                  if (grid_to_node[g] < 0)
                    {
                      grid_to_node[g]    = next;
                      node_to_grid[next] = g;
                      ++next;
                    }
                }
        assert(next == n_nodes);
      }

    // cell arrays
    const int nv = 1 << dim;
    cell_dofs.resize(n_cells * dofs_per_cell);
    cell_vertices.resize(n_cells * nv * dim);
    double h[3]{0, 0, 0};
    for (int d = 0; d < dim; ++d)
      h[d] = (p1[d] - p0[d]) / reps[d];
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < reps[1]; ++j)
        for (int i = 0; i < reps[0]; ++i)
          {
            const int64_t c      = cell_index(i, j, k);
            const int     ijk[3] = {i, j, k};
            for (int a = 0; a < nodes_per_cell; ++a)
              {
                const auto &  l  = local_node_lex[a];
                const int64_t gx = int64_t(i) * p + l[0], gy = int64_t(j) * p + l[1],
                              gz = int64_t(k) * p + l[2];
                const int64_t g  = (gz * grid_nodes[1] + gy) * grid_nodes[0] + gx;
                for (int comp = 0; comp < dim; ++comp)
                  cell_dofs[c * dofs_per_cell + local_dof_of[a * dim + comp]] =
                    dof_of(grid_to_node[g], comp);
              }
            for (int v = 0; v < nv; ++v)
              for (int d = 0; d < dim; ++d)
                {
                  const int idx = ijk[d] + ((v >> d) & 1);
                  // last vertex hits p1 exactly (as subdivided_hyper_rectangle does)
                  cell_vertices[(c * nv + v) * dim + d] =
                    idx == reps[d] ? p1[d] : p0[d] + idx * h[d];
                }
          }

    // support points of all dofs
    support_points.resize(n_dofs * dim);
    for (int64_t n = 0; n < n_nodes; ++n)
      {
        const int64_t g     = node_to_grid[n];
        const int64_t gi[3] = {g % grid_nodes[0], (g / grid_nodes[0]) % grid_nodes[1],
                               g / (grid_nodes[0] * grid_nodes[1])};
        for (int comp = 0; comp < dim; ++comp)
          for (int d = 0; d < dim; ++d)
            support_points[int64_t(dof_of(n, comp)) * dim + d] =
              gi[d] == grid_nodes[d] - 1 ?
                p1[d] :
                (p <= 2 ? p0[d] + gi[d] * (h[d] / p) :
                          p0[d] + (double(gi[d] / p) + unit_support_1d[gi[d] % p]) * h[d]);
      }
  }

  unsigned StructuredMesh::node_face_mask(int64_t g) const
  {
    const int64_t gi[3] = {g % grid_nodes[0], (g / grid_nodes[0]) % grid_nodes[1],
                           g / (grid_nodes[0] * grid_nodes[1])};
    unsigned      m     = 0;
    for (int d = 0; d < dim; ++d)
      {
        if (gi[d] == 0)
          m |= 1u << (2 * d);
        if (gi[d] == grid_nodes[d] - 1)
          m |= 1u << (2 * d + 1);
      }
    return m;
  }

  int64_t StructuredMesh::boundary_dof_mask(unsigned face_mask, unsigned comp_mask,
                                            uint8_t *mask) const
  {
    int64_t count = 0;
    for (int64_t n = 0; n < n_nodes; ++n)
      if (node_face_mask(node_to_grid[n]) & face_mask)
        for (int comp = 0; comp < dim; ++comp)
          if (comp_mask & (1u << comp))
            {
              const int32_t dof = dof_of(n, comp);
              if (!mask[dof])
                ++count;
              mask[dof] = 1;
            }
    return count;
  }

  void StructuredMesh::boundary_faces(unsigned face_mask, std::vector<int32_t> &cells,
                                      std::vector<int32_t> &faces) const
  {
    cells.clear();
    faces.clear();
    const int nz = dim == 3 ? reps[2] : 1;
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < reps[1]; ++j)
        for (int i = 0; i < reps[0]; ++i)
          {
            const int ijk[3] = {i, j, k};
            for (int f = 0; f < 2 * dim; ++f)
              {
                const int  d           = f / 2;
                const bool at_boundary = (f % 2 == 0) ? ijk[d] == 0 : ijk[d] == reps[d] - 1;
                if (at_boundary && (face_mask & (1u << f)))
                  {
                    cells.push_back(int32_t(cell_index(i, j, k)));
                    faces.push_back(f);
                  }
              }
          }
  }

  int64_t StructuredMesh::interface_dofs(unsigned face_mask, std::vector<int32_t> &out) const
  {
    std::vector<int64_t> nodes;
    for (int64_t n = 0; n < n_nodes; ++n)
      if (node_face_mask(node_to_grid[n]) & face_mask)
        nodes.push_back(n);
    // IndexSets are ascending in the global dof index; node order is identical for every
    // component in all numberings offered here (adapter.h:394-399 relies on this)
    std::sort(nodes.begin(), nodes.end(),
              [&](int64_t a, int64_t b) { return dof_of(a, 0) < dof_of(b, 0); });
    const int64_t n = int64_t(nodes.size());
    out.resize(n * dim);
    for (int comp = 0; comp < dim; ++comp)
      for (int64_t i = 0; i < n; ++i)
        out[comp * n + i] = dof_of(nodes[i], comp);
    return n;
  }

  MeshPartition partition_mesh(const StructuredMesh &mesh, int axis, int nparts, int rank)
  {
    if (axis < 0 || axis >= mesh.dim)
      throw std::invalid_argument("partition_mesh: bad axis");
    if (nparts < 1 || rank < 0 || rank >= nparts || mesh.reps[axis] < nparts)
      throw std::invalid_argument("partition_mesh: need at least one cell layer per part");
    const int p   = mesh.degree;
    const int dim = mesh.dim;
    MeshPartition part;
    part.rank   = rank;
    part.nparts = nparts;
    part.axis   = axis;
    auto layer_begin = [&](int r) { return int((int64_t(mesh.reps[axis]) * r) / nparts); };
    const int l0 = layer_begin(rank), l1 = layer_begin(rank + 1);
    // node planes (grid index along axis) owned by part r: (l0*p, l1*p], part 0 also owns plane 0
    auto owner_of_plane = [&](int64_t g) {
      if (g == 0)
        return 0;
      // smallest r with g <= layer_begin(r+1)*p
      int r = 0;
      while (g > int64_t(layer_begin(r + 1)) * p)
        ++r;
      return r;
    };
    // local cells: owned layers plus one ghost layer above (rows on the shared plane are
    // owned by the lower part, which assembles the neighbouring layer redundantly)
    const int lc0 = l0, lc1 = (rank + 1 < nparts) ? l1 + 1 : l1;
    const int nz  = dim == 3 ? mesh.reps[2] : 1;
    std::vector<int64_t> cells;
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < mesh.reps[1]; ++j)
        for (int i = 0; i < mesh.reps[0]; ++i)
          {
            const int ijk[3] = {i, j, k};
            if (ijk[axis] >= lc0 && ijk[axis] < lc1)
              cells.push_back(mesh.cell_index(i, j, k));
          }
    part.n_local_cells     = int64_t(cells.size());
    part.local_cell_global = cells;
    // collect local nodes, split owned/ghost
    std::vector<int64_t> owned, ghost;
    {
      std::vector<uint8_t> seen(mesh.n_nodes, 0);
      for (int64_t c : cells)
        for (int a = 0; a < mesh.nodes_per_cell; ++a)
          {
            const int32_t d0 = mesh.cell_dofs[c * mesh.dofs_per_cell + mesh.local_dof_of[a * dim]];
            const int64_t n  = mesh.numbering == numbering_component_wise ? d0 : d0 / dim;
            if (seen[n])
              continue;
            seen[n]            = 1;
            const int64_t g    = mesh.node_to_grid[n];
            const int64_t gi[3] = {g % mesh.grid_nodes[0],
                                   (g / mesh.grid_nodes[0]) % mesh.grid_nodes[1],
                                   g / (mesh.grid_nodes[0] * mesh.grid_nodes[1])};
            (owner_of_plane(gi[axis]) == rank ? owned : ghost).push_back(n);
          }
    }
    std::sort(owned.begin(), owned.end());
    std::sort(ghost.begin(), ghost.end());
    std::map<int64_t, int64_t> local_node;
    for (int64_t n : owned)
      local_node.emplace(n, int64_t(local_node.size()));
    for (int64_t n : ghost)
      local_node.emplace(n, int64_t(local_node.size()));
    part.n_owned_dofs = int64_t(owned.size()) * dim;
    part.n_local_dofs = int64_t(local_node.size()) * dim;
    part.local_to_global.resize(part.n_local_dofs);
    for (auto &kv : local_node)
      for (int comp = 0; comp < dim; ++comp)
        part.local_to_global[kv.second * dim + comp] = mesh.dof_of(kv.first, comp);
    const int nv = 1 << dim;
    part.cell_dofs.resize(cells.size() * mesh.dofs_per_cell);
    part.cell_vertices.resize(cells.size() * nv * dim);
    for (size_t lc = 0; lc < cells.size(); ++lc)
      {
        const int64_t c = cells[lc];
        for (int a = 0; a < mesh.nodes_per_cell; ++a)
          {
            const int32_t d0 = mesh.cell_dofs[c * mesh.dofs_per_cell + mesh.local_dof_of[a * dim]];
            const int64_t n  = mesh.numbering == numbering_component_wise ? d0 : d0 / dim;
            for (int comp = 0; comp < dim; ++comp)
              part.cell_dofs[lc * mesh.dofs_per_cell + mesh.local_dof_of[a * dim + comp]] =
                int32_t(local_node[n] * dim + comp);
          }
        std::copy(mesh.cell_vertices.begin() + c * nv * dim,
                  mesh.cell_vertices.begin() + (c + 1) * nv * dim,
                  part.cell_vertices.begin() + lc * nv * dim);
      }
    // exchange lists: receive my ghosts from their owners; send my owned nodes that are ghosts
    // of a neighbour. With slabs only ranks rank-1 and rank+1 are involved.
    auto plane_of = [&](int64_t n) {
      const int64_t g     = mesh.node_to_grid[n];
      const int64_t gi[3] = {g % mesh.grid_nodes[0],
                             (g / mesh.grid_nodes[0]) % mesh.grid_nodes[1],
                             g / (mesh.grid_nodes[0] * mesh.grid_nodes[1])};
      return gi[axis];
    };
    part.send_ptr.push_back(0);
    part.recv_ptr.push_back(0);
    for (int nbr : {rank - 1, rank + 1})
      {
        if (nbr < 0 || nbr >= nparts)
          continue;
        std::vector<int64_t> send_nodes, recv_nodes;
        for (int64_t n : ghost)
          if (owner_of_plane(plane_of(n)) == nbr)
            recv_nodes.push_back(n);
        if (nbr == rank + 1)
          {
            // upper neighbour's ghosts on my side: my top plane l1*p
            for (int64_t n : owned)
              if (plane_of(n) == int64_t(l1) * p)
                send_nodes.push_back(n);
          }
        else
          {
            // lower neighbour assembles my first cell layer: it needs planes l0*p+1 .. l0*p+p
            for (int64_t n : owned)
              {
                const int64_t g = plane_of(n);
                if (g > int64_t(l0) * p && g <= int64_t(l0) * p + p)
                  send_nodes.push_back(n);
              }
          }
        part.nbr_rank.push_back(nbr);
        for (int64_t n : send_nodes)
          for (int comp = 0; comp < dim; ++comp)
            part.send_dofs.push_back(int32_t(local_node[n] * dim + comp));
        for (int64_t n : recv_nodes)
          for (int comp = 0; comp < dim; ++comp)
            part.recv_dofs.push_back(int32_t(local_node[n] * dim + comp));
        part.send_ptr.push_back(int64_t(part.send_dofs.size()));
        part.recv_ptr.push_back(int64_t(part.recv_dofs.size()));
      }
    return part;
  }
} // namespace gfh

using gfh::MeshPartition;
using gfh::StructuredMesh;

extern "C"
{
  void *gfh_mesh_create(int dim, int degree, const int *reps, const double *p0, const double *p1,
                        int numbering)
  {
    try
      {
        return new StructuredMesh(dim, degree, reps, p0, p1, numbering);
      }
    catch (...)
      {
        return nullptr;
      }
  }
  void    gfh_mesh_destroy(void *m) { delete static_cast<StructuredMesh *>(m); }
  int64_t gfh_mesh_n_cells(const void *m) { return static_cast<const StructuredMesh *>(m)->n_cells; }
  int64_t gfh_mesh_n_dofs(const void *m) { return static_cast<const StructuredMesh *>(m)->n_dofs; }
  int64_t gfh_mesh_n_nodes(const void *m) { return static_cast<const StructuredMesh *>(m)->n_nodes; }
  int     gfh_mesh_dofs_per_cell(const void *m)
  {
    return static_cast<const StructuredMesh *>(m)->dofs_per_cell;
  }
  const int32_t *gfh_mesh_cell_dofs(const void *m)
  {
    return static_cast<const StructuredMesh *>(m)->cell_dofs.data();
  }
  const double *gfh_mesh_cell_vertices(const void *m)
  {
    return static_cast<const StructuredMesh *>(m)->cell_vertices.data();
  }
  const double *gfh_mesh_support_points(const void *m)
  {
    return static_cast<const StructuredMesh *>(m)->support_points.data();
  }
  int64_t gfh_mesh_boundary_dof_mask(const void *m, unsigned face_mask, unsigned comp_mask,
                                     uint8_t *mask)
  {
    return static_cast<const StructuredMesh *>(m)->boundary_dof_mask(face_mask, comp_mask, mask);
  }
  int64_t gfh_mesh_boundary_faces(const void *m, unsigned face_mask, int32_t *cells,
                                  int32_t *faces)
  {
    std::vector<int32_t> c, f;
    static_cast<const StructuredMesh *>(m)->boundary_faces(face_mask, c, f);
    if (cells && faces)
      {
        std::copy(c.begin(), c.end(), cells);
        std::copy(f.begin(), f.end(), faces);
      }
    return int64_t(c.size());
  }
  int64_t gfh_mesh_interface_dofs(const void *m, unsigned face_mask, int32_t *out)
  {
    std::vector<int32_t> v;
    const int64_t n = static_cast<const StructuredMesh *>(m)->interface_dofs(face_mask, v);
    if (out)
      std::copy(v.begin(), v.end(), out);
    return n;
  }

  void *gfh_partition_create(const void *m, int axis, int nparts, int rank)
  {
    try
      {
        return new MeshPartition(
          gfh::partition_mesh(*static_cast<const StructuredMesh *>(m), axis, nparts, rank));
      }
    catch (...)
      {
        return nullptr;
      }
  }
  void    gfh_partition_destroy(void *p) { delete static_cast<MeshPartition *>(p); }
  int64_t gfh_partition_size(const void *p_, int which)
  {
    const MeshPartition *p = static_cast<const MeshPartition *>(p_);
    switch (which)
      {
        case 0: return p->n_local_dofs;
        case 1: return p->n_owned_dofs;
        case 2: return p->n_local_cells;
        case 3: return int64_t(p->nbr_rank.size());
        case 4: return int64_t(p->send_dofs.size());
        case 5: return int64_t(p->recv_dofs.size());
      }
    return -1;
  }
  const void *gfh_partition_array(const void *p_, int which)
  {
    const MeshPartition *p = static_cast<const MeshPartition *>(p_);
    switch (which)
      {
        case 0: return p->cell_dofs.data();
        case 1: return p->cell_vertices.data();
        case 2: return p->local_cell_global.data();
        case 3: return p->local_to_global.data();
        case 4: return p->nbr_rank.data();
        case 5: return p->send_ptr.data();
        case 6: return p->recv_ptr.data();
        case 7: return p->send_dofs.data();
        case 8: return p->recv_dofs.data();
      }
    return nullptr;
  }

  // ---- csrc/fe_basis.h seen from the tests (tests/test_fe_basis.py compares it with numpy) ----
  void gfh_fe_support_points(int p, double *out)
  {
    const std::vector<double> x = gf_fe::gauss_lobatto01(p + 1);
    std::copy(x.begin(), x.end(), out);
  }
  void gfh_fe_basis_eval(int p, int n_points, const double *x, double *values, double *derivatives)
  {
    const gf_fe::Basis1D b(p);
    for (int k = 0; k < n_points; ++k)
      for (int i = 0; i <= p; ++i)
        {
          values[k * (p + 1) + i]      = b.value(i, x[k]);
          derivatives[k * (p + 1) + i] = b.derivative(i, x[k]);
        }
  }
  // lex[npc * 3], node_of / comp_of [npc * dim]; returns npc
  int gfh_fe_numbering(int dim, int p, int *lex, int *node_of, int *comp_of)
  {
    std::vector<int> l, n, c, inv;
    gf_fe::local_nodes(dim, p, l);
    gf_fe::system_numbering(dim, p, n, c, inv);
    std::copy(l.begin(), l.end(), lex);
    std::copy(n.begin(), n.end(), node_of);
    std::copy(c.begin(), c.end(), comp_of);
    return int(l.size() / 3);
  }
}
