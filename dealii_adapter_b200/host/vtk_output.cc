#include "vtk_output.h"

#include <cstdio>
#include <ctime>
#include <fstream>
#include <iomanip>

namespace gfh
{
  // [VTK] vtkLagrangeQuadrilateral / vtkLagrangeHexahedron::PointIndexFromIJK, as restated by
  // deal.II's DataOutBase (vtk_point_index_from_ijk): vertices, then edge interiors, then face
  // interiors, then the body interior.
  int vtk_point_index_from_ijk(int dim, int p, int i, int j, int k, bool legacy)
  {
    const bool ibdy = (i == 0 || i == p), jbdy = (j == 0 || j == p);
    if (dim == 2)
      {
        const int nbdy = (ibdy ? 1 : 0) + (jbdy ? 1 : 0);
        if (nbdy == 2)
          return i ? (j ? 2 : 1) : (j ? 3 : 0);
        int offset = 4;
        if (nbdy == 1)
          {
            if (!ibdy) // edge along i
              return (i - 1) + (j ? (p - 1) + (p - 1) : 0) + offset;
            return (j - 1) + (i ? (p - 1) : 2 * (p - 1) + (p - 1)) + offset; // edge along j
          }
        offset += 4 * (p - 1);
        return offset + (i - 1) + (p - 1) * (j - 1);
      }
    const bool kbdy = (k == 0 || k == p);
    const int  nbdy = (ibdy ? 1 : 0) + (jbdy ? 1 : 0) + (kbdy ? 1 : 0);
    if (nbdy == 3)
      return (i ? (j ? 2 : 1) : (j ? 3 : 0)) + (k ? 4 : 0);
    int offset = 8;
    if (nbdy == 2)
      {
        if (!ibdy)
          return (i - 1) + (j ? (p - 1) + (p - 1) : 0) + (k ? 2 * ((p - 1) + (p - 1)) : 0) + offset;
        if (!jbdy)
          return (j - 1) + (i ? (p - 1) : 2 * (p - 1) + (p - 1)) +
                 (k ? 2 * ((p - 1) + (p - 1)) : 0) + offset;
        offset += 8 * (p - 1);
        if (legacy)
          return (k - 1) + (p - 1) * (i ? (j ? 3 : 1) : (j ? 2 : 0)) + offset;
        return (k - 1) + (p - 1) * (i ? (j ? 2 : 1) : (j ? 3 : 0)) + offset;
      }
    offset += 12 * (p - 1);
    if (nbdy == 1)
      {
        const int f = (p - 1) * (p - 1);
        if (ibdy)
          return (j - 1) + (p - 1) * (k - 1) + (i ? f : 0) + offset;
        offset += 2 * f;
        if (jbdy)
          return (i - 1) + (p - 1) * (k - 1) + (j ? f : 0) + offset;
        offset += 2 * f;
        return (i - 1) + (p - 1) * (j - 1) + (k ? f : 0) + offset;
      }
    offset += 6 * (p - 1) * (p - 1);
    return offset + (i - 1) + (p - 1) * ((j - 1) + (p - 1) * (k - 1));
  }

  void patch_points(int dim, int degree, int64_t n_cells, const double *cell_vertices,
                    const double *fields, const uint8_t *at_boundary, std::vector<double> &points)
  {
    const int p = degree, n1 = p + 1, nv = 1 << dim, nf = dim + dim * dim;
    int       npts = 1;
    for (int d = 0; d < dim; ++d)
      npts *= n1;
    points.assign(size_t(n_cells) * npts * dim, 0.0);
    for (int64_t cell = 0; cell < n_cells; ++cell)
      {
        const double *verts = cell_vertices + cell * nv * dim;
        const double *fld   = fields + cell * npts * nf;
        const bool    curved = at_boundary == nullptr || at_boundary[cell] != 0;
        // displaced vertices: patch point of vertex v is (bit_d(v) * p) in every direction
        double xv[8][3];
        for (int v = 0; v < nv; ++v)
          {
            int pt = 0, mul = 1;
            for (int d = 0; d < dim; ++d)
              {
                pt += ((v >> d) & 1) * p * mul;
                mul *= n1;
              }
            for (int d = 0; d < dim; ++d)
              xv[v][d] = verts[v * dim + d] + fld[pt * nf + d];
          }
        for (int pt = 0; pt < npts; ++pt)
          {
            double xi[3] = {0, 0, 0};
            int    rem   = pt;
            for (int d = 0; d < dim; ++d)
              {
                xi[d] = double(rem % n1) / p;
                rem /= n1;
              }
            double *out = &points[(size_t(cell) * npts + pt) * dim];
            for (int v = 0; v < nv; ++v)
              {
                double w = 1;
                for (int d = 0; d < dim; ++d)
                  w *= ((v >> d) & 1) ? xi[d] : 1.0 - xi[d];
                for (int d = 0; d < dim; ++d)
                  out[d] += w * (curved ? verts[v * dim + d] : xv[v][d]);
              }
            if (curved) // X(xi) + u(xi): the Eulerian mapping itself
              for (int d = 0; d < dim; ++d)
                out[d] += fld[pt * nf + d];
          }
      }
  }

  bool write_vtk(const std::string &filename, int dim, int degree, int64_t n_cells,
                 const double *points, const double *fields)
  {
    std::ofstream out(filename);
    if (!out)
      return false;
    const int p = degree, n1 = p + 1, nf = dim + dim * dim;
    int       npts = 1;
    for (int d = 0; d < dim; ++d)
      npts *= n1;
    const int64_t n_nodes = n_cells * npts;
    std::time_t   t       = std::time(nullptr);
    std::tm       tm_buf{};
    localtime_r(&t, &tm_buf);
    char stamp[64];
    std::snprintf(stamp, sizeof stamp, "%d/%d/%d at %d:%02d:%02d", tm_buf.tm_year + 1900,
                  tm_buf.tm_mon + 1, tm_buf.tm_mday, tm_buf.tm_hour, tm_buf.tm_min, tm_buf.tm_sec);
    out << "# vtk DataFile Version 3.0\n"
        << "#This file was generated by the deal.II-adapter B200 host (deal.II write_vtk layout) on "
        << stamp << "\nASCII\nDATASET UNSTRUCTURED_GRID\n\n";
    out << std::setprecision(12);
    out << "POINTS " << n_nodes << " double\n";
    for (int64_t n = 0; n < n_nodes; ++n)
      {
        for (int d = 0; d < 3; ++d)
          out << (d < dim ? points[n * dim + d] : 0.0) << (d < 2 ? ' ' : '\n');
      }
    out << "\nCELLS " << n_cells << ' ' << n_cells * (npts + 1) << '\n';
    // connectivity position -> lexicographic patch point
    std::vector<int> lex_of_vtk(npts);
    for (int k = 0; k < (dim == 3 ? n1 : 1); ++k)
      for (int j = 0; j < n1; ++j)
        for (int i = 0; i < n1; ++i)
          lex_of_vtk[vtk_point_index_from_ijk(dim, p, i, j, k, true)] = (k * n1 + j) * n1 + i;
    for (int64_t cell = 0; cell < n_cells; ++cell)
      {
        out << npts;
        for (int v = 0; v < npts; ++v)
          out << '\t' << cell * npts + lex_of_vtk[v];
        out << '\n';
      }
    out << "\nCELL_TYPES " << n_cells << '\n';
    for (int64_t cell = 0; cell < n_cells; ++cell)
      out << ' ' << (dim == 2 ? 70 : 72);
    out << "\nPOINT_DATA " << n_nodes << '\n';
    out << "VECTORS displacement double\n";
    for (int64_t n = 0; n < n_nodes; ++n)
      for (int d = 0; d < 3; ++d)
        out << (d < dim ? fields[n * nf + d] : 0.0) << (d < 2 ? ' ' : '\n');
    static const char suffixes[] = {'x', 'y', 'z'};
    for (int d = 0; d < dim; ++d)
      for (int e = 0; e < dim; ++e)
        {
          out << "SCALARS strain_" << suffixes[d] << suffixes[e] << " double 1\nLOOKUP_TABLE default\n";
          for (int64_t n = 0; n < n_nodes; ++n)
            out << fields[n * nf + dim + d * dim + e] << ' ';
          out << '\n';
        }
    out.flush();
    return bool(out);
  }
} // namespace gfh

extern "C"
{
  int gfh_write_vtk(const char *filename, int dim, int degree, int64_t n_cells,
                    const double *cell_vertices, const double *fields, const uint8_t *at_boundary)
  {
    std::vector<double> pts;
    gfh::patch_points(dim, degree, n_cells, cell_vertices, fields, at_boundary, pts);
    return gfh::write_vtk(filename, dim, degree, n_cells, pts.data(), fields) ? 0 : 1;
  }
  int gfh_vtk_point_index_from_ijk(int dim, int p, int i, int j, int k, int legacy)
  {
    return gfh::vtk_point_index_from_ijk(dim, p, i, j, k, legacy != 0);
  }
}
