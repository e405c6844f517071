// Reverse Cuthill-McKee ordering of the node graph (host, once per handle): the band Cholesky of
// direct_band.cuh costs n w^2, so the half bandwidth w is what matters. Input: the block-row
// pattern (row_ptr, cols; symmetric, self-loops allowed). Output: node_new[old] = new index,
// deterministic (the identity if the given numbering already has the narrower band); returns the
// half bandwidth in nodes, max |new(A) - new(B)| over the blocks.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace gf
{
  inline int64_t rcm_order(const int64_t n, const int32_t *row_ptr, const int32_t *cols,
                           std::vector<int32_t> &node_new)
  {
    std::vector<int32_t> degree(n), order, level(n, -1), queue;
    for (int64_t a = 0; a < n; ++a)
      degree[a] = row_ptr[a + 1] - row_ptr[a];
    order.reserve(n);
    std::vector<uint8_t> visited(n, 0);
    // BFS from `start` over the not yet ordered nodes; returns the nodes in visiting order with
    // the neighbours of a node taken by ascending (degree, index); fills level[]
    auto bfs = [&](int32_t start, std::vector<int32_t> &out, std::vector<uint8_t> &seen) {
      out.clear();
      out.push_back(start);
      seen[start]  = 1;
      level[start] = 0;
      std::vector<int32_t> nbrs;
      for (size_t head = 0; head < out.size(); ++head)
        {
          const int32_t u = out[head];
          nbrs.clear();
          for (int32_t k = row_ptr[u]; k < row_ptr[u + 1]; ++k)
            {
              const int32_t v = cols[k];
              if (!seen[v])
                {
                  seen[v]  = 1;
                  level[v] = level[u] + 1;
                  nbrs.push_back(v);
                }
            }
          std::sort(nbrs.begin(), nbrs.end(), [&](int32_t a, int32_t b) {
            return degree[a] != degree[b] ? degree[a] < degree[b] : a < b;
          });
          out.insert(out.end(), nbrs.begin(), nbrs.end());
        }
    };
    std::vector<int32_t> comp;
    for (int64_t s0 = 0; s0 < n; ++s0)
      {
        if (visited[s0])
          continue;
        // pseudo-peripheral start node of this component: repeat BFS from a minimum-degree node
        // of the last level while the eccentricity grows
        int32_t start = int32_t(s0);
        int     ecc   = -1;
        for (int iter = 0; iter < 8; ++iter)
          {
            std::vector<uint8_t> seen(visited);
            bfs(start, comp, seen);
            const int last = level[comp.back()];
            if (last <= ecc)
              break;
            ecc          = last;
            int32_t best = comp.back();
            for (auto it = comp.rbegin(); it != comp.rend() && level[*it] == last; ++it)
              if (degree[*it] < degree[best] || (degree[*it] == degree[best] && *it < best))
                best = *it;
            if (best == start)
              break;
            start = best;
          }
        bfs(start, comp, visited);
        order.insert(order.end(), comp.begin(), comp.end());
      }
    node_new.assign(n, -1);
    for (int64_t k = 0; k < n; ++k)
      node_new[order[n - 1 - k]] = int32_t(k); // reversed Cuthill-McKee
    int64_t w = 0, w_given = 0;
    for (int64_t a = 0; a < n; ++a)
      for (int32_t k = row_ptr[a]; k < row_ptr[a + 1]; ++k)
        {
          w       = std::max<int64_t>(w, std::abs(int64_t(node_new[a]) - int64_t(node_new[cols[k]])));
          w_given = std::max<int64_t>(w_given, std::abs(a - int64_t(cols[k])));
        }
    // the level sets of a high-order stencil are two node layers thick, so a given lexicographic
    // numbering across the short side of a slender mesh can beat RCM: keep the better one
    if (w_given <= w)
      {
        for (int64_t a = 0; a < n; ++a)
          node_new[a] = int32_t(a);
        w = w_given;
      }
    return w;
  }
} // namespace gf
