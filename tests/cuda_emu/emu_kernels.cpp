// The generic-degree kernels of the product (dealii_adapter_b200/csrc/assemble_nl_generic.cuh),
// compiled UNCHANGED with g++ against the CUDA stand-in in this directory and driven with the
// product's own host tables (csrc/fe_tables_host.h). Test infrastructure only.
#include <cstring>

#include "assemble_general.cuh"
#include "assemble_nl_generic.cuh"
#include "constraints.cuh"
#include "direct_band.cuh"
#include "fe_tables_host.h"
#include "rcm.h"

namespace
{
  // verts == nullptr: affine cells described by geom (J^-1, det J); else general cells
  template <int DIM>
  int run_cells(int p, int64_t n_cells, const int32_t *cell_nodes, const double *geom,
                const double *verts, const double *u_total, const double *accel,
                const gf::NLParams &prm, unsigned grid, unsigned block, double *ke, double *re)
  {
    gf::FETablesHost t;
    gf::build_host_tables(t, DIM, p, true);
    int          err  = 0;
    const size_t smem = size_t(gf::NLGen<DIM>::smem_doubles(t.npc)) * sizeof(double);
    gf_emu::launch(grid, block, smem, [&] {
      if (verts == nullptr)
        gf::nl_cells_generic_kernel<DIM, true>(0, n_cells, t.npc, t.nq, cell_nodes, geom, nullptr,
                                               nullptr, u_total, accel, t.hN.data(), t.hdN.data(),
                                               t.hw.data(), t.hMref.data(), prm, ke, re, &err);
      else
        gf::nl_cells_generic_kernel<DIM, false>(0, n_cells, t.npc, t.nq, cell_nodes, nullptr, verts,
                                                t.hdphi.data(), u_total, accel, t.hN.data(),
                                                t.hdN.data(), t.hw.data(), nullptr, prm, ke, re,
                                                &err);
    });
    return err;
  }

  template <int DIM>
  int run_faces(int p, int n_iface_cells, const int32_t *cell_list, const int32_t *face_ptr,
                const int32_t *face_no, const int32_t *cell_nodes, const double *geom,
                const double *verts, const double *u_total, const double *stress, unsigned block,
                double *re)
  {
    gf::FETablesHost t;
    gf::build_host_tables(t, DIM, p, true);
    int          err  = 0;
    const size_t smem = size_t(gf::nl_faces_generic_smem_doubles<DIM>(t.npc, t.nqf)) * sizeof(double);
    gf_emu::launch(unsigned(n_iface_cells), block, smem, [&] {
      if (verts == nullptr)
        gf::nl_faces_generic_kernel<DIM, true>(n_iface_cells, t.npc, t.nqf, cell_list, face_ptr,
                                               face_no, cell_nodes, geom, nullptr, nullptr, nullptr,
                                               u_total, stress, t.hdN.data(), t.hNf.data(),
                                               t.hwf.data(), re, &err);
      else
        gf::nl_faces_generic_kernel<DIM, false>(n_iface_cells, t.npc, t.nqf, cell_list, face_ptr,
                                                face_no, cell_nodes, nullptr, verts, t.hdphi.data(),
                                                t.hdphif.data(), u_total, stress, t.hdN.data(),
                                                t.hNf.data(), t.hwf.data(), re, &err);
    });
    return err;
  }

  // linear model on general cells: K_e [n_cells][dpc][dpc], scalar mass blocks [n_cells][npc][npc],
  // body-force load and interface load [n_cells][dpc] (internal local order)
  template <int DIM>
  int run_linear(int p, int64_t n_cells, const double *verts, double lambda, double mu, double rho,
                 const double *bf, int n_iface_cells, const int32_t *cell_list,
                 const int32_t *face_ptr, const int32_t *face_no, const int32_t *cell_nodes,
                 const double *stress, unsigned block, double *ke, double *me, double *re_body,
                 double *re_face)
  {
    gf::FETablesHost t;
    gf::build_host_tables(t, DIM, p, false);
    int          err  = 0;
    const size_t smem = (size_t(t.nq) * t.npc * DIM + t.nq) * sizeof(double);
    gf_emu::launch(2u, block, smem, [&] {
      gf::lin_cells_general_kernel<DIM>(0, n_cells, t.npc, t.nq, verts, t.hdphi.data(), t.hN.data(),
                                        t.hdN.data(), t.hw.data(), lambda, mu, rho, ke, me, &err);
    });
    gf_emu::launch(2u, block, 0, [&] {
      gf::body_force_general_kernel<DIM>(n_cells, t.npc, t.nq, verts, t.hdphi.data(), t.hN.data(),
                                         t.hw.data(), rho, bf[0], bf[1], bf[2], re_body);
    });
    const size_t smem_f = (2 * size_t(t.dpc) + size_t(t.nqf) * (DIM + 1)) * sizeof(double);
    gf_emu::launch(unsigned(n_iface_cells), block, smem_f, [&] {
      gf::lin_faces_general_kernel<DIM>(n_iface_cells, t.npc, t.nqf, cell_list, face_ptr, face_no,
                                        cell_nodes, verts, t.hdphif.data(), stress, t.hNf.data(),
                                        t.hwf.data(), re_face);
    });
    return err;
  }

  template <int DIM>
  int run_postprocess(int p, int64_t n_cells, const int32_t *cell_nodes, const double *verts,
                      const double *u, unsigned block, double *fields)
  {
    gf::FETablesHost t;
    gf::build_host_tables(t, DIM, p, true);
    // shape values / unit-cell gradients at the patch points, as launch_postprocess builds them
    const gf_fe::Basis1D basis(p);
    const int            npc = t.npc;
    std::vector<double>  N(size_t(npc) * npc), dN(size_t(npc) * npc * DIM);
    for (int pt = 0; pt < npc; ++pt)
      {
        double xi[3] = {0, 0, 0};
        int    rem   = pt;
        for (int d = 0; d < DIM; ++d)
          {
            xi[d] = double(rem % (p + 1)) / p;
            rem /= (p + 1);
          }
        for (int a = 0; a < npc; ++a)
          {
            double v = 1;
            for (int d = 0; d < DIM; ++d)
              v *= basis.value(t.local_lex[a * 3 + d], xi[d]);
            N[size_t(pt) * npc + a] = v;
            for (int k = 0; k < DIM; ++k)
              {
                double w = 1;
                for (int d = 0; d < DIM; ++d)
                  w *= (d == k) ? basis.derivative(t.local_lex[a * 3 + d], xi[d]) :
                                  basis.value(t.local_lex[a * 3 + d], xi[d]);
                dN[(size_t(pt) * npc + a) * DIM + k] = w;
              }
          }
      }
    int err = 0;
    gf_emu::launch(2u, block, 0, [&] {
      gf::postprocess_general_kernel<DIM>(0, n_cells, npc, cell_nodes, verts, t.hdphip.data(),
                                          N.data(), dN.data(), u, fields, &err);
    });
    return err;
  }
} // namespace

namespace
{
  // kernel<<<grid, block, smem>>>(args...) of the product's drivers (db_factor, db_solve)
  struct EmuLaunch
  {
    template <class... KA, class... A>
    void operator()(unsigned grid, unsigned block, size_t smem, void (*kernel)(KA...), A... args)
    {
      gf_emu::launch(grid, block, smem, [&] { kernel(args...); });
    }
  };
} // namespace

extern "C"
{
  // 'Solver type = Direct': reverse Cuthill-McKee + band Cholesky + substitution on a matrix in the
  // library's block-row format. Returns the info flag of the factorisation (0 = positive definite);
  // w_out: [0] = half bandwidth (scalar) in the RCM ordering, [1] = in the given ordering.
  int emu_direct_solve(int64_t n_nodes, int dim, const int32_t *brow_ptr, const int64_t *val_ptr,
                       const int32_t *bcol, const double *val, const double *b, double *x,
                       int64_t *w_out)
  {
    std::vector<int32_t> node_new;
    const int64_t        wn = gf::rcm_order(n_nodes, brow_ptr, bcol, node_new);
    int64_t              w0 = 0;
    for (int64_t a = 0; a < n_nodes; ++a)
      for (int32_t k = brow_ptr[a]; k < brow_ptr[a + 1]; ++k)
        w0 = std::max<int64_t>(w0, std::abs(a - int64_t(bcol[k])));
    const int64_t n = n_nodes * dim, w = wn * dim + dim - 1, ld = w + gf::DB_NB;
    w_out[0]        = w;
    w_out[1]        = w0 * dim + dim - 1;
    std::vector<double> band(size_t(n) * ld, 0.0), xp(n, 0.0), tmp(gf::DB_NB, 0.0);
    std::vector<double> bb(b, b + n);
    int                 info = 0;
    EmuLaunch           launch;
    launch(unsigned(std::min<int64_t>(n_nodes, 7)), 64u, size_t(0), gf::db_fill_kernel, n_nodes, dim,
           brow_ptr, val_ptr, bcol, val, (const int32_t *)node_new.data(), ld, band.data());
    gf::db_factor(launch, n, w, ld, band.data(), &info);
    launch(3u, 64u, size_t(0), gf::db_permute_kernel, n_nodes, dim,
           (const int32_t *)node_new.data(), true, bb.data(), xp.data());
    gf::db_solve(launch, n, w, ld, (const double *)band.data(), xp.data(), tmp.data());
    launch(3u, 64u, size_t(0), gf::db_permute_kernel, n_nodes, dim,
           (const int32_t *)node_new.data(), false, x, xp.data());
    return info;
  }

  // params: kappa, mu, rho, alpha_1, body_force[3]. ke: [n_cells][dpc][dpc] (only the node blocks
  // b <= a are written), re: [n_cells][dpc]. Returns the det F error flag.
  int emu_nl_cells(int dim, int p, int64_t n_cells, const int32_t *cell_nodes, const double *geom,
                   const double *verts, const double *u_total, const double *accel,
                   const double *params, unsigned grid, unsigned block, double *ke, double *re)
  {
    gf::NLParams prm;
    prm.kappa   = params[0];
    prm.mu      = params[1];
    prm.rho     = params[2];
    prm.alpha_1 = params[3];
    for (int k = 0; k < 3; ++k)
      prm.body_force[k] = params[4 + k];
    return dim == 3 ?
             run_cells<3>(p, n_cells, cell_nodes, geom, verts, u_total, accel, prm, grid, block, ke, re) :
             run_cells<2>(p, n_cells, cell_nodes, geom, verts, u_total, accel, prm, grid, block, ke, re);
  }
  int emu_nl_faces(int dim, int p, int n_iface_cells, const int32_t *cell_list,
                   const int32_t *face_ptr, const int32_t *face_no, const int32_t *cell_nodes,
                   const double *geom, const double *verts, const double *u_total,
                   const double *stress, unsigned block, double *re)
  {
    return dim == 3 ? run_faces<3>(p, n_iface_cells, cell_list, face_ptr, face_no, cell_nodes, geom,
                                   verts, u_total, stress, block, re) :
                      run_faces<2>(p, n_iface_cells, cell_list, face_ptr, face_no, cell_nodes, geom,
                                   verts, u_total, stress, block, re);
  }
  int emu_linear_general(int dim, int p, int64_t n_cells, const double *verts, double lambda,
                         double mu, double rho, const double *bf, int n_iface_cells,
                         const int32_t *cell_list, const int32_t *face_ptr, const int32_t *face_no,
                         const int32_t *cell_nodes, const double *stress, unsigned block, double *ke,
                         double *me, double *re_body, double *re_face)
  {
    return dim == 3 ? run_linear<3>(p, n_cells, verts, lambda, mu, rho, bf, n_iface_cells, cell_list,
                                    face_ptr, face_no, cell_nodes, stress, block, ke, me, re_body,
                                    re_face) :
                      run_linear<2>(p, n_cells, verts, lambda, mu, rho, bf, n_iface_cells, cell_list,
                                    face_ptr, face_no, cell_nodes, stress, block, ke, me, re_body,
                                    re_face);
  }
  int emu_postprocess_general(int dim, int p, int64_t n_cells, const int32_t *cell_nodes,
                              const double *verts, const double *u, unsigned block, double *fields)
  {
    return dim == 3 ? run_postprocess<3>(p, n_cells, cell_nodes, verts, u, block, fields) :
                      run_postprocess<2>(p, n_cells, cell_nodes, verts, u, block, fields);
  }
  // hanging-node constraint lines: x <- C x (distribute), y <- C^T y (condense + zero)
  void emu_lines(int64_t n_lines, const int32_t *dof, const int64_t *ptr, const int32_t *master,
                 const double *weight, int64_t n_masters, const int32_t *mdof, const int64_t *mptr,
                 const int32_t *slave, const double *mweight, double *x, double *y)
  {
    gf_emu::launch(2u, 32u, 0, [&] { gf::lines_distribute_kernel(n_lines, dof, ptr, master, weight, x); });
    gf_emu::launch(2u, 32u, 0,
                   [&] { gf::lines_condense_kernel(n_masters, mdof, mptr, slave, mweight, y); });
    gf_emu::launch(2u, 32u, 0, [&] { gf::lines_zero_kernel(n_lines, dof, y); });
  }
  // the product's host tables, for the tests: loc_of [npc * dim]
  int emu_loc_of(int dim, int p, int *loc_of)
  {
    gf::FETablesHost t;
    gf::build_host_tables(t, dim, p, true);
    std::memcpy(loc_of, t.loc_of.data(), t.loc_of.size() * sizeof(int));
    return t.npc;
  }
}
