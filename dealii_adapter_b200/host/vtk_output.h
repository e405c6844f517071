// Host side of output_results (nonlinear_elasticity.cc:1215-1254, linear_elasticity.cc:590-629):
// writes what `DataOut::build_patches(MappingQEulerian, degree, curved_boundary)` +
// `write_vtk` with `write_higher_order_cells = true` produce, from the patch fields the device
// library computes (gf_postprocess: displacement | strain per patch point, postprocessor.h:44-76).
// File layout follows DataOutBase::write_vtk of deal.II 9.5 (restated, deal.II is not in the tree):
// legacy ASCII UNSTRUCTURED_GRID, one VTK_LAGRANGE_QUADRILATERAL (70) / VTK_LAGRANGE_HEXAHEDRON
// (72) cell of order `degree` per patch, points not shared between patches, point data
// "displacement" as VECTORS followed by the scalars strain_xx, strain_xy, ...
// (Postprocessor::get_names, postprocessor.h:81-97).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace gfh
{
  // connectivity position of the lexicographic patch point (i, j[, k]) inside a VTK Lagrange
  // quadrilateral / hexahedron of order `p` per direction ([deal.II] vtk_point_index_from_ijk;
  // legacy = true is the edge order of VTK file version < 5 that write_vtk uses)
  int vtk_point_index_from_ijk(int dim, int p, int i, int j, int k, bool legacy);

  // points [n_cells][npts][dim] on the displaced grid: boundary cells (`at_boundary[cell]` != 0) get
  // X + u at every patch point, interior cells the multilinear interpolation of their displaced
  // vertices (DataOut::curved_boundary). fields: gf_postprocess layout.
  void patch_points(int dim, int degree, int64_t n_cells, const double *cell_vertices,
                    const double *fields, const uint8_t *at_boundary, std::vector<double> &points);

  // returns false if the file cannot be opened
  bool write_vtk(const std::string &filename, int dim, int degree, int64_t n_cells,
                 const double *points, const double *fields);
} // namespace gfh

extern "C"
{
  // C entry for the Python mirror / tests: computes the patch points and writes the file;
  // at_boundary may be NULL (all cells use the mapping). Returns 0 on success.
  int gfh_write_vtk(const char *filename, int dim, int degree, int64_t n_cells,
                    const double *cell_vertices, const double *fields, const uint8_t *at_boundary);
  int gfh_vtk_point_index_from_ijk(int dim, int p, int i, int j, int k, int legacy);
}
