// Set-up phase: internal node-block numbering, node->cell adjacency, BSR sparsity pattern and the
// deterministic scatter map. Replaces DoFTools::make_sparsity_pattern + the binary searches of
// AffineConstraints::distribute_local_to_global (reference: nonlinear_elasticity.cc:330-349,
// 769-773; linear_elasticity.cc:203-214,327-334). Runs once per gf_create; thrust is used only for
// the set-up sorts/scans, never on the per-timestep path.
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>

#include <algorithm>
#include <cmath>

#include "gf_context.h"

namespace gf
{
  namespace
  {
    constexpr int MAX_CELLS_PER_NODE = 16;

    __global__ void count_nodes_kernel(const int32_t *cell_nodes, int64_t n, int *count)
    {
      const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (i < n)
        atomicAdd(&count[cell_nodes[i]], 1);
    }

    // one warp per owned node; candidates = all nodes of all cells around the node
    template <bool FILL>
    __global__ void pattern_kernel(int64_t n_owned_nodes, int npc, const int64_t *nc_ptr,
                                   const int32_t *nc_src, const int32_t *cell_nodes,
                                   const int32_t *node_rank, int *row_nb,
                                   const int32_t *brow_ptr, const int64_t *cand_ptr, int32_t *bcol,
                                   uint16_t *src_off, int32_t *row_src, int *overflow)
    {
      extern __shared__ int32_t smem[];
      const int                 warps_per_block = blockDim.x / 32;
      const int                 warp = threadIdx.x / 32, lane = threadIdx.x % 32;
      const int64_t             A = blockIdx.x * int64_t(warps_per_block) + warp;
      if (A >= n_owned_nodes)
        return;
      const int max_cand = MAX_CELLS_PER_NODE * npc;
      int32_t * cand     = smem + warp * 3 * max_cand;
      const int64_t s0   = nc_ptr[A];
      const int     ncell = int(nc_ptr[A + 1] - s0);
      if (ncell > MAX_CELLS_PER_NODE)
        {
          if (lane == 0)
            atomicExch(overflow, 1);
          return;
        }
      const int ncand = ncell * npc;
      for (int j = lane; j < ncand; j += 32)
        {
          const int32_t s    = nc_src[s0 + j / npc];
          const int32_t cell = s / npc;
          cand[j]            = cell_nodes[int64_t(cell) * npc + j % npc];
        }
      __syncwarp();
      // blocks of a row are ordered by node_rank = the node's position in the partition-independent
      // (node plane, global id) order, NOT by the local index (ghosts are numbered last locally)
      int32_t *rk = cand + 2 * max_cand;
      for (int j = lane; j < ncand; j += 32)
        rk[j] = node_rank[cand[j]];
      __syncwarp();
      int32_t *is_first = cand + max_cand;
      int      n_first  = 0;
      // pass 1: per candidate, number of smaller candidates and of earlier duplicates
      for (int j = lane; j < ncand; j += 32)
        {
          const int32_t v = cand[j];
          int           dup = 0;
          for (int i = 0; i < j; ++i)
            dup += (cand[i] == v);
          is_first[j] = (dup == 0);
          n_first += (dup == 0);
        }
      if (FILL)
        {
          __syncwarp();
          for (int j = lane; j < ncand; j += 32)
            {
              const int32_t v = rk[j];
              int           less = 0, dup = 0, first_less = 0;
              for (int i = 0; i < ncand; ++i)
                {
                  const int32_t u = rk[i];
                  less += (u < v);
                  first_less += (u < v) & is_first[i];
                  dup += (u == v) & (i < j);
                }
              if (dup == 0)
                {
                  // rank among distinct values = number of distinct values < v
                  const int64_t blk = int64_t(brow_ptr[A]) + first_less;
                  bcol[blk]         = cand[j];
                  src_off[blk]      = uint16_t(less);
                }
              const int32_t s                    = nc_src[s0 + j / npc];
              row_src[cand_ptr[A] + less + dup] = s * npc + j % npc;
            }
        }
      if (!FILL)
        {
          for (int o = 16; o > 0; o >>= 1)
            n_first += __shfl_xor_sync(0xffffffffu, n_first, o);
          if (lane == 0)
            row_nb[A] = n_first;
        }
    }

    __global__ void row_sizes_kernel(int64_t n_rows, int dim, const int *row_nb,
                                     const int64_t *nc_ptr, int npc, int64_t *val_sz,
                                     int64_t *cand_sz)
    {
      const int64_t A = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (A >= n_rows)
        return;
      const int64_t stride = (int64_t(row_nb[A]) * dim + 1) & ~int64_t(1);
      val_sz[A]            = stride * dim;
      cand_sz[A]           = (nc_ptr[A + 1] - nc_ptr[A]) * npc;
    }
  } // namespace

  void build_numbering_and_pattern(gf_context &c, const gf_desc &d)
  {
    const int     dim = c.dim, npc = c.npc, dpc = c.dpc;
    const int64_t n_ext = d.n_dofs, n_cells = d.n_cells;
    cudaStream_t  s = c.stream;

    // ---- host: nodes = distinct x-component dofs. Internal order: owned nodes first; inside
    // each group by (node plane orthogonal to the slab axis, partition-independent id), so that a
    // node plane is a contiguous range (reduction chunks, x-gather locality of the SpMV) ----------
    GF_REQUIRE(d.slab_axis >= 0 && d.slab_axis <= dim, GF_ERR_INVALID_ARG,
               "slab_axis must be 0 (not given) or 1 + a coordinate axis");
    c.slab_axis = d.slab_axis;
    std::vector<int32_t> node_of_x(n_ext, -1);
    std::vector<double>  coord_of_x; // coordinate of the node along the slab axis
    int                  axis_dir = -1;
    double               h_min    = 0.0;
    if (c.slab_axis > 0)
      coord_of_x.assign(n_ext, 0.0);
    const int              ax  = c.slab_axis - 1;
    const std::vector<int> &lex = c.tables.local_lex;
    // cell_dofs is in the caller's FESystem local order: entry of (local node a, component comp)
    const std::vector<int> &loc_of = c.tables.loc_of;
    for (int64_t cell = 0; cell < n_cells; ++cell)
      {
        double Jax[3] = {0, 0, 0}, v0 = 0;
        if (c.slab_axis > 0)
          {
            const double *v = d.cell_vertices + cell * (int64_t(1) << dim) * dim;
            v0              = v[ax];
            int    best  = 0;
            double scale = 0;
            for (int j = 0; j < dim; ++j)
              {
                Jax[j] = v[(int64_t(1) << j) * dim + ax] - v[ax];
                if (std::fabs(Jax[j]) > std::fabs(Jax[best]))
                  best = j;
                for (int i = 0; i < dim; ++i)
                  scale = std::max(scale, std::fabs(v[(int64_t(1) << j) * dim + i] - v[i]));
              }
            bool ok = Jax[best] > 0 && (axis_dir < 0 || axis_dir == best);
            for (int j = 0; j < dim; ++j)
              ok = ok && (j == best || std::fabs(Jax[j]) <= 1e-10 * scale);
            GF_REQUIRE(ok, GF_ERR_UNSUPPORTED,
                       "slab_axis needs cells of one common orientation with edges along the axis");
            axis_dir = best;
            h_min    = (cell == 0) ? Jax[best] : std::min(h_min, Jax[best]);
          }
        for (int a = 0; a < npc; ++a)
          {
            const int32_t xd = d.cell_dofs[cell * dpc + loc_of[a * dim]];
            GF_REQUIRE(xd >= 0 && xd < n_ext, GF_ERR_INVALID_ARG, "cell_dofs entry out of range");
            if (node_of_x[xd] < 0 && c.slab_axis > 0)
              coord_of_x[xd] =
                v0 + Jax[axis_dir] * (c.p <= 2 ? double(lex[a * 3 + axis_dir]) / c.p :
                                                 c.tables.support_1d[lex[a * 3 + axis_dir]]);
            node_of_x[xd] = 0;
          }
      }
    c.axis_dir = axis_dir;
    std::vector<int32_t> xdofs; // distinct x-dofs, ascending caller id
    for (int64_t i = 0; i < n_ext; ++i)
      if (node_of_x[i] == 0)
        xdofs.push_back(int32_t(i));
    const int64_t n_nodes = int64_t(xdofs.size());
    GF_REQUIRE(n_nodes * dim == n_ext, GF_ERR_INVALID_ARG,
               "n_dofs is not dim * (number of distinct nodes in cell_dofs)");
    // node planes: runs of (nearly) equal coordinates
    std::vector<int32_t> plane_of(n_nodes, 0);
    if (c.slab_axis > 0)
      {
        std::vector<int32_t> by_coord(n_nodes);
        for (int64_t k = 0; k < n_nodes; ++k)
          by_coord[k] = int32_t(k);
        std::sort(by_coord.begin(), by_coord.end(), [&](int32_t a, int32_t b) {
          return coord_of_x[xdofs[a]] < coord_of_x[xdofs[b]];
        });
        // the closest two support points of a cell are more than h / p^2 apart (Gauss-Lobatto)
        const double tol = c.p <= 2 ? 1e-6 * h_min / c.p : 1e-6 * h_min / (c.p * c.p);
        int32_t      pl  = 0;
        for (int64_t k = 0; k < n_nodes; ++k)
          {
            if (k > 0 && coord_of_x[xdofs[by_coord[k]]] - coord_of_x[xdofs[by_coord[k - 1]]] > tol)
              ++pl;
            plane_of[by_coord[k]] = pl;
          }
      }
    // order of all local nodes by (plane, global id): the partition-independent order
    struct NodeKey
    {
      int32_t plane;
      int64_t gkey;
      int32_t k; // index into xdofs
    };
    std::vector<NodeKey> keys(n_nodes);
    for (int64_t k = 0; k < n_nodes; ++k)
      keys[k] = {plane_of[k], d.dof_global ? d.dof_global[xdofs[k]] : int64_t(xdofs[k]), int32_t(k)};
    std::sort(keys.begin(), keys.end(), [](const NodeKey &a, const NodeKey &b) {
      return a.plane != b.plane ? a.plane < b.plane : a.gkey < b.gkey;
    });
    for (int64_t k = 1; k < n_nodes; ++k)
      GF_REQUIRE(keys[k].plane != keys[k - 1].plane || keys[k].gkey != keys[k - 1].gkey,
                 GF_ERR_INVALID_ARG, "dof_global is not unique");
    // internal node index: owned nodes in that order, then the ghosts in that order
    int64_t              n_owned_nodes = 0;
    std::vector<int32_t> h_node_rank(n_nodes);
    std::vector<int64_t> h_node_gkey(n_nodes);
    c.h_plane_ptr.clear();
    for (int pass = 0; pass < 2; ++pass)
      {
        int64_t next = pass == 0 ? 0 : n_owned_nodes;
        int32_t last_plane = -1;
        for (int64_t r = 0; r < n_nodes; ++r)
          {
            const int32_t xd    = xdofs[keys[r].k];
            const bool    owned = xd < c.n_ext_owned;
            if (owned != (pass == 0))
              continue;
            if (pass == 0 && keys[r].plane != last_plane)
              {
                c.h_plane_ptr.push_back(int32_t(next));
                last_plane = keys[r].plane;
              }
            node_of_x[xd]     = int32_t(next);
            h_node_rank[next] = int32_t(r);
            h_node_gkey[next] = keys[r].gkey;
            ++next;
          }
        if (pass == 0)
          {
            n_owned_nodes = next;
            c.h_plane_ptr.push_back(int32_t(next));
          }
      }
    DevBuf<int32_t> node_rank;
    node_rank.upload(h_node_rank.data(), h_node_rank.size(), s);
    c.node_gkey.upload(h_node_gkey.data(), h_node_gkey.size(), s);
    c.n_nodes       = n_nodes;
    c.n_owned_nodes = n_owned_nodes;
    c.n_local       = n_nodes * dim;
    c.n_owned       = n_owned_nodes * dim;
    GF_REQUIRE(c.n_owned == c.n_ext_owned, GF_ERR_INVALID_ARG,
               "n_owned_dofs must cover whole nodes and precede ghost dofs");
    c.h_perm_e2i.assign(n_ext, -1);
    c.h_perm_i2e.assign(n_ext, -1);
    std::vector<int32_t> cell_nodes(n_cells * npc);
    for (int64_t cell = 0; cell < n_cells; ++cell)
      for (int a = 0; a < npc; ++a)
        {
          const int32_t node          = node_of_x[d.cell_dofs[cell * dpc + loc_of[a * dim]]];
          cell_nodes[cell * npc + a] = node;
          for (int comp = 0; comp < dim; ++comp)
            {
              const int32_t e = d.cell_dofs[cell * dpc + loc_of[a * dim + comp]];
              GF_REQUIRE(e >= 0 && e < n_ext, GF_ERR_INVALID_ARG, "cell_dofs entry out of range");
              const int32_t i = node * dim + comp;
              GF_REQUIRE(c.h_perm_e2i[e] == -1 || c.h_perm_e2i[e] == i, GF_ERR_INVALID_ARG,
                         "inconsistent node/component structure in cell_dofs");
              c.h_perm_e2i[e] = i;
              c.h_perm_i2e[i] = e;
            }
        }
    for (int64_t i = 0; i < n_ext; ++i)
      GF_REQUIRE(c.h_perm_e2i[i] >= 0 && c.h_perm_i2e[i] >= 0, GF_ERR_INVALID_ARG,
                 "dof not referenced by any cell");
    for (int64_t i = 0; i < n_ext; ++i)
      GF_REQUIRE((c.h_perm_i2e[i] < c.n_ext_owned) == (i < c.n_owned), GF_ERR_INVALID_ARG,
                 "owned node has a non-owned component dof");
    c.perm_e2i.upload(c.h_perm_e2i.data(), n_ext, s);
    c.perm_i2e.upload(c.h_perm_i2e.data(), n_ext, s);
    c.cell_nodes.upload(cell_nodes.data(), cell_nodes.size(), s);
    {
      std::vector<uint8_t> con(n_ext);
      for (int64_t i = 0; i < n_ext; ++i)
        con[i] = d.constrained[c.h_perm_i2e[i]] ? 1 : 0;
      c.constrained.upload(con.data(), n_ext, s);
    }

    // ---- per-cell affine geometry: J^{-1} (row-major) and det J -------------------------------
    {
      const int           nv = 1 << dim, gs = dim * dim + 1;
      std::vector<double> geom(n_cells * gs);
      bool                all_affine = true;
      for (int64_t cell = 0; cell < n_cells; ++cell)
        {
          const double *v = d.cell_vertices + cell * nv * dim;
          double        J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
          for (int i = 0; i < dim; ++i)
            for (int j = 0; j < dim; ++j)
              J[i][j] = v[(1 << j) * dim + i] - v[i];
          // affine check: every vertex = v0 + sum_j bit_j * column_j
          double scale = 0;
          for (int i = 0; i < dim; ++i)
            for (int j = 0; j < dim; ++j)
              scale = std::max(scale, std::fabs(J[i][j]));
          for (int vtx = 0; vtx < nv; ++vtx)
            for (int i = 0; i < dim; ++i)
              {
                double x = v[i];
                for (int j = 0; j < dim; ++j)
                  if ((vtx >> j) & 1)
                    x += J[i][j];
                if (std::fabs(x - v[vtx * dim + i]) > 1e-10 * scale)
                  all_affine = false; // general cell: Jacobian per quadrature point
              }
          const double det =
            dim == 2 ? J[0][0] * J[1][1] - J[0][1] * J[1][0] :
                       J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
                         J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                         J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
          GF_REQUIRE(det > 0, GF_ERR_INVALID_ARG, "cell with non-positive Jacobian determinant");
          double inv[3][3];
          if (dim == 2)
            {
              inv[0][0] = J[1][1] / det;
              inv[0][1] = -J[0][1] / det;
              inv[1][0] = -J[1][0] / det;
              inv[1][1] = J[0][0] / det;
            }
          else
            {
              inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
              inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
              inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
              inv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
              inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
              inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
              inv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
              inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
              inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
            }
          for (int i = 0; i < dim; ++i)
            for (int j = 0; j < dim; ++j)
              geom[cell * gs + i * dim + j] = inv[i][j];
          geom[cell * gs + dim * dim] = det;
        }
      c.geom.upload(geom.data(), geom.size(), s);
      c.affine = all_affine;
      if (!all_affine)
        {
          GF_REQUIRE(c.slab_axis == 0, GF_ERR_UNSUPPORTED,
                     "slab_axis (partition-independent node planes) needs parallelepiped cells");
          c.cell_verts.upload(d.cell_vertices, size_t(n_cells) * nv * dim, s);
        }
    }

    // ---- device: node -> (cell, local node) adjacency, ascending cell ---------------------------
    const int64_t n_cn = n_cells * npc;
    GF_REQUIRE(n_cn * npc < (int64_t(1) << 31), GF_ERR_UNSUPPORTED,
               "mesh too large for 32-bit scatter indices on one rank");
    {
      DevBuf<int32_t> keys;
      keys.alloc(n_cn);
      GF_CUDA_CHECK(cudaMemcpyAsync(keys.p, c.cell_nodes.p, n_cn * sizeof(int32_t),
                                    cudaMemcpyDeviceToDevice, s));
      c.nc_src.alloc(n_cn);
      auto pol = thrust::cuda::par.on(s);
      thrust::sequence(pol, thrust::device_pointer_cast(c.nc_src.p),
                       thrust::device_pointer_cast(c.nc_src.p + n_cn));
      thrust::stable_sort_by_key(pol, thrust::device_pointer_cast(keys.p),
                                 thrust::device_pointer_cast(keys.p + n_cn),
                                 thrust::device_pointer_cast(c.nc_src.p));
      DevBuf<int> counts;
      counts.alloc_zero(n_nodes + 1, s);
      count_nodes_kernel<<<unsigned((n_cn + 255) / 256), 256, 0, s>>>(c.cell_nodes.p, n_cn,
                                                                       counts.p);
      c.nc_ptr.alloc(n_nodes + 1);
      thrust::exclusive_scan(pol, thrust::device_pointer_cast(counts.p),
                             thrust::device_pointer_cast(counts.p + n_nodes + 1),
                             thrust::device_pointer_cast(c.nc_ptr.p), int64_t(0));
      GF_CUDA_CHECK(cudaStreamSynchronize(s));
    }

    // ---- device: BSR pattern + scatter map for owned rows ---------------------------------------
    const int64_t n_rows = n_owned_nodes;
    DevBuf<int>   row_nb, overflow;
    row_nb.alloc_zero(n_rows + 1, s);
    overflow.alloc_zero(1, s);
    const int    wpb  = 4;
    const size_t smem = size_t(wpb) * 3 * MAX_CELLS_PER_NODE * npc * sizeof(int32_t);
    const unsigned grid = unsigned((n_rows + wpb - 1) / wpb);
    pattern_kernel<false><<<grid, wpb * 32, smem, s>>>(n_rows, npc, c.nc_ptr.p, c.nc_src.p,
                                                       c.cell_nodes.p, node_rank.p, row_nb.p, nullptr,
                                                       nullptr, nullptr, nullptr, nullptr,
                                                       overflow.p);
    GF_CUDA_CHECK(cudaGetLastError());
    int h_overflow = 0;
    GF_CUDA_CHECK(
      cudaMemcpyAsync(&h_overflow, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    GF_CUDA_CHECK(cudaStreamSynchronize(s));
    GF_REQUIRE(h_overflow == 0, GF_ERR_UNSUPPORTED, "a node is shared by more than 16 cells");
    auto pol = thrust::cuda::par.on(s);
    c.brow_ptr.alloc(n_rows + 1);
    thrust::exclusive_scan(pol, thrust::device_pointer_cast(row_nb.p),
                           thrust::device_pointer_cast(row_nb.p + n_rows + 1),
                           thrust::device_pointer_cast(c.brow_ptr.p), int32_t(0));
    DevBuf<int64_t> val_sz, cand_sz;
    val_sz.alloc_zero(n_rows + 1, s);
    cand_sz.alloc_zero(n_rows + 1, s);
    row_sizes_kernel<<<unsigned((n_rows + 255) / 256), 256, 0, s>>>(
      n_rows, dim, row_nb.p, c.nc_ptr.p, npc, val_sz.p, cand_sz.p);
    c.val_ptr.alloc(n_rows + 1);
    c.cand_ptr.alloc(n_rows + 1);
    thrust::exclusive_scan(pol, thrust::device_pointer_cast(val_sz.p),
                           thrust::device_pointer_cast(val_sz.p + n_rows + 1),
                           thrust::device_pointer_cast(c.val_ptr.p), int64_t(0));
    thrust::exclusive_scan(pol, thrust::device_pointer_cast(cand_sz.p),
                           thrust::device_pointer_cast(cand_sz.p + n_rows + 1),
                           thrust::device_pointer_cast(c.cand_ptr.p), int64_t(0));
    int32_t nb_total = 0;
    int64_t nval = 0, ncand = 0;
    GF_CUDA_CHECK(cudaMemcpyAsync(&nb_total, c.brow_ptr.p + n_rows, sizeof(int32_t),
                                  cudaMemcpyDeviceToHost, s));
    GF_CUDA_CHECK(
      cudaMemcpyAsync(&nval, c.val_ptr.p + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    GF_CUDA_CHECK(
      cudaMemcpyAsync(&ncand, c.cand_ptr.p + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    GF_CUDA_CHECK(cudaStreamSynchronize(s));
    c.n_blocks = nb_total;
    c.n_val    = nval;
    c.n_cand   = ncand;
    c.bcol.alloc_zero(c.n_blocks + 8, s); // padded: tile copies fetch 16-byte aligned supersets
    c.src_off.alloc(c.n_blocks);
    c.row_src.alloc(c.n_cand);
    pattern_kernel<true><<<grid, wpb * 32, smem, s>>>(n_rows, npc, c.nc_ptr.p, c.nc_src.p,
                                                      c.cell_nodes.p, node_rank.p, nullptr,
                                                      c.brow_ptr.p, c.cand_ptr.p, c.bcol.p,
                                                      c.src_off.p, c.row_src.p, overflow.p);
    GF_CUDA_CHECK(cudaGetLastError());
    GF_CUDA_CHECK(cudaStreamSynchronize(s));

    // ---- host: SpMV tiles (greedy runs of consecutive rows) -------------------------------------
    {
      std::vector<int32_t> brow(n_rows + 1);
      std::vector<int64_t> vptr(n_rows + 1);
      GF_CUDA_CHECK(cudaMemcpy(brow.data(), c.brow_ptr.p, (n_rows + 1) * sizeof(int32_t),
                               cudaMemcpyDeviceToHost));
      GF_CUDA_CHECK(cudaMemcpy(vptr.data(), c.val_ptr.p, (n_rows + 1) * sizeof(int64_t),
                               cudaMemcpyDeviceToHost));
      const int             tile_v = spmv_tile_v(dim);
      const int             tile_c = ((tile_v / (dim * dim)) + 8 + 3) & ~3;
      std::vector<TileDesc> descs;
      std::vector<uint2>    meta;
      bool                  ok   = n_rows > 0;
      int64_t               row0 = 0;
      while (ok && row0 < n_rows)
        {
          const int32_t c0   = brow[row0] & ~3;
          int64_t       row1 = row0;
          while (row1 < n_rows && row1 - row0 < SPMV_TILE_ROWS &&
                 vptr[row1 + 1] - vptr[row0] <= tile_v &&
                 ((brow[row1 + 1] + 3) & ~3) - c0 <= tile_c && brow[row1 + 1] - brow[row1] < 65536)
            ++row1;
          if (row1 == row0)
            {
              ok = false; // a single row exceeds the tile: keep the LDG kernel for this matrix
              break;
            }
          TileDesc d;
          d.val_off   = vptr[row0];
          d.val_count = int(vptr[row1] - vptr[row0]);
          d.col_off   = c0;
          d.col_count = ((brow[row1] + 3) & ~3) - c0;
          d.row0      = int(row0);
          d.n_rows    = int(row1 - row0);
          d.pad       = 0;
          descs.push_back(d);
          const size_t base = meta.size();
          meta.resize(base + SPMV_META, make_uint2(0, 0));
          for (int64_t r = row0; r < row1; ++r)
            meta[base + (r - row0)] =
              make_uint2(unsigned(vptr[r] - vptr[row0]),
                         unsigned(brow[r] - c0) | (unsigned(brow[r + 1] - brow[r]) << 16));
          meta[base + SPMV_TILE_ROWS]     = make_uint2(unsigned(row0), unsigned(row1 - row0));
          // .y: elements the FP32 bulk copy of this tile starts early to be 16-byte aligned
          meta[base + SPMV_TILE_ROWS + 1] =
            make_uint2(unsigned(d.col_count), unsigned(vptr[row0] & 3));
          row0 = row1;
        }
      if (ok)
        {
          c.n_tiles = int64_t(descs.size());
          c.tile_desc.upload(descs.data(), descs.size(), s);
          c.tile_meta.upload(meta.data(), meta.size(), s);
        }
      else
        c.n_tiles = 0;
    }
  }
} // namespace gf
