// C-ABI implementation (include/graft_fem.h): context life cycle and the step functions that the
// host classes Solid / ElastoDynamics / Adapter forward to.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "reduce.cuh"

namespace
{
  thread_local std::string g_create_error;

  template <typename F>
  int guarded(gf_handle h, F &&f)
  {
    if (!h)
      return GF_ERR_INVALID_ARG;
    try
      {
        GF_CUDA_CHECK(cudaSetDevice(h->device));
        return f(*h);
      }
    catch (gf::Error &e)
      {
        h->last_error = e.msg;
        return e.code;
      }
    catch (std::exception &e)
      {
        h->last_error = e.what();
        return GF_ERR_INVALID_ARG;
      }
  }

  double *vec_ptr(gf_context &c, int which)
  {
    GF_REQUIRE(which >= 0 && which < gf::MAX_VECTORS && c.vec[which].p != nullptr,
               GF_ERR_INVALID_ARG, "vector id not available for this model");
    return c.vec[which].p;
  }
  // Deferred half of assemble_system. The reference assembles tangent AND residual in every
  // Newton pass, also in the last one, where the convergence test (:459-463) only needs the
  // residual and no solve follows. gf_nl_newton_assemble therefore runs the cell kernel (element
  // matrices + residual) and stops; scattering K_e into the matrix, the block-Jacobi inverse and
  // the multigrid re-discretisation happen here, on the first use of the tangent - the solve, or
  // an export / vmult. The element buffer keeps the K_e of the whole mesh until then (single
  // chunk on every rank: agreed at gf_create, since the multigrid update is collective).
  // Two halves, because the second one is COLLECTIVE on a partitioned handle (halo exchanges and
  // reductions of the coarse levels) and must only run inside entry points that every rank calls:
  //   finish_tangent_local  scatter + block-Jacobi inverse: rank-local, legal anywhere (a single
  //                         rank may export rows or time the SpMV of its own matrix)
  //   finish_tangent        + multigrid update: gf_nl_newton_solve / gf_mg_vcycle only
  void finish_tangent_local(gf_context &c)
  {
    if (!c.tangent_pending)
      return;
    c.tangent_pending = false;
    double *K = c.mat[GF_MAT_TANGENT].val.p;
    gf::launch_scatter_matrix(c, K, 0, c.n_cells, true, true);
    c.mat[GF_MAT_TANGENT].valid = true;
    gf::launch_build_precond(c, K);
    c.mg_update_pending = gf::mg_active(c);
  }
  void finish_tangent(gf_context &c)
  {
    finish_tangent_local(c);
    if (c.mg_update_pending)
      {
        c.mg_update_pending = false;
        gf::mg_update_operators(c, c.tmp0.p); // coarse tangents at the injected state + smoothers
      }
  }

  // x = A^-1 b by the band Cholesky (+ one step of iterative refinement with the FP64 matrix, as
  // UMFPACK does by default). false: not available here (partitioned handle, band beyond the
  // memory budget, matrix-free operator, GF_OPT_DIRECT_SOLVER = 2, A not positive definite) - the
  // caller then runs the tight CG. `constant_matrix`: reuse the factor of the previous call.
  bool direct_apply(gf_context &c, const double *A, const double *b, double *x,
                    const bool constant_matrix)
  {
    if (c.operator_kind != 0 || !gf::direct_available(c))
      return false;
    try
      {
        if (!(constant_matrix && c.direct.factored && c.direct.factor_is_system_matrix))
          {
            if (!gf::direct_factor(c, A))
              {
                GF_REQUIRE(c.direct_mode != 1, GF_ERR_NOT_CONVERGED,
                           "direct solver: the matrix is not positive definite");
                return false;
              }
            c.direct.factor_is_system_matrix = constant_matrix;
          }
        if (c.lines.n > 0)
          {
            // hanging-node constraint lines: the band holds A, the system is the condensed
            // C^T A C - the factor preconditions a CG on the condensed operator (cg.cu), which
            // ends after a handful of iterations; tolerance as tight as the stand-in's
            const double bn = gf::vec_masked_norm(c, b, false);
            uint32_t     it  = 0;
            double       res = 0;
            const int    rc  = gf::cg_solve_direct_precond(c, A, x, b, 1e-13 * bn, 200, &it, &res);
            if (rc != GF_OK && !(res <= 1e-10 * bn))
              {
                GF_REQUIRE(c.direct_mode != 1, GF_ERR_NOT_CONVERGED,
                           "direct solver: the condensed iteration did not converge");
                return false;
              }
            c.direct.last_residual = bn > 0 ? res / bn : res;
            c.direct.n_solves++;
            return true;
          }
        gf::direct_solve(c, b, x);
        gf::launch_spmv(c, A, x, c.cg_v.p, nullptr);
        gf::vec_copy(c, c.cg_r.p, b);
        gf::vec_axpby(c, c.cg_r.p, -1.0, c.cg_v.p, 1.0); // r = b - A x
        gf::direct_solve(c, c.cg_r.p, c.cg_z.p);
        gf::vec_axpby(c, x, 1.0, c.cg_z.p, 1.0);
        // the answer is checked before it is used. Forward error: the refinement step must have
        // been a small correction, ||dx|| <= 1e-6 ||x|| (then the refined x is good to ~1e-12;
        // a broken or hopelessly ill-conditioned factorisation gives dx ~ x). Residual: the
        // attainable ||b - A x|| / ||b|| is eps ||A|| ||x|| / ||b||, 1e-14 .. 1e-12 on the test
        // meshes and growing like h^-2, so only a loose bound is imposed on it.
        gf::launch_spmv(c, A, x, c.cg_v.p, nullptr);
        gf::vec_copy(c, c.cg_r.p, b);
        gf::vec_axpby(c, c.cg_r.p, -1.0, c.cg_v.p, 1.0);
        const double rn = gf::vec_masked_norm(c, c.cg_r.p, false);
        const double bn = gf::vec_masked_norm(c, b, false);
        const double zn = gf::vec_masked_norm(c, c.cg_z.p, false);
        const double xn = gf::vec_masked_norm(c, x, false);
        c.direct.last_residual = bn > 0 ? rn / bn : rn;
        if (!(zn <= 1e-6 * xn && rn <= 1e-6 * bn))
          {
            GF_REQUIRE(c.direct_mode != 1, GF_ERR_NOT_CONVERGED,
                       "direct solver: refinement / residual check failed");
            return false;
          }
        c.direct.n_solves++;
        return true;
      }
    catch (gf::Error &e)
      {
        if (c.direct_mode == 1)
          throw;
        // first run of this path on a new device / size: never let it cost the solve
        cudaGetLastError();
        c.direct.usable   = false;
        c.direct.factored = false;
        c.direct.why      = e.msg;
        return false;
      }
  }

  double *mat_ptr(gf_context &c, int which)
  {
    if (which == GF_MAT_TANGENT)
      finish_tangent_local(c);
    GF_REQUIRE(which >= 0 && which < gf::N_MATRICES && which != GF_MAT_MASS &&
                 c.mat[which].val.p != nullptr,
               GF_ERR_INVALID_ARG, "matrix id not available for this model");
    return c.mat[which].val.p;
  }

  void setup_interface(gf_context &c, const gf_desc &d)
  {
    cudaStream_t s  = c.stream;
    c.n_iface_nodes = d.n_iface_nodes;
    if (d.n_iface_nodes > 0)
      {
        std::vector<int32_t> idofs(size_t(d.n_iface_nodes) * c.dim);
        for (size_t k = 0; k < idofs.size(); ++k)
          {
            const int32_t e = d.iface_dofs[k];
            GF_REQUIRE(e >= 0 && e < c.n_ext_dofs, GF_ERR_INVALID_ARG, "iface_dofs out of range");
            idofs[k] = c.h_perm_e2i[e];
          }
        c.iface_dofs_i.upload(idofs.data(), idofs.size(), s);
        c.iface_buf.alloc(idofs.size());
      }
    // interface faces grouped by cell, ascending face number (cell->face_iterators())
    std::vector<std::pair<int32_t, int32_t>> cf(d.n_iface_faces);
    for (int64_t f = 0; f < d.n_iface_faces; ++f)
      {
        GF_REQUIRE(d.iface_cell[f] >= 0 && d.iface_cell[f] < d.n_cells && d.iface_face_no[f] >= 0 &&
                     d.iface_face_no[f] < 2 * c.dim,
                   GF_ERR_INVALID_ARG, "interface face out of range");
        cf[f] = {d.iface_cell[f], d.iface_face_no[f]};
      }
    std::sort(cf.begin(), cf.end());
    std::vector<int32_t> cells, ptr{0}, faces;
    for (size_t k = 0; k < cf.size(); ++k)
      {
        if (k == 0 || cf[k].first != cf[k - 1].first)
          {
            if (k)
              ptr.push_back(int32_t(faces.size()));
            cells.push_back(cf[k].first);
          }
        faces.push_back(cf[k].second);
      }
    ptr.push_back(int32_t(faces.size()));
    c.n_iface_cells = int64_t(cells.size());
    if (c.n_iface_cells)
      {
        c.iface_cell_list.upload(cells.data(), cells.size(), s);
        c.iface_face_ptr.upload(ptr.data(), ptr.size(), s);
        c.iface_face_no.upload(faces.data(), faces.size(), s);
      }
  }

  void setup_halo(gf_context &c, const gf_desc &d)
  {
    c.comm = d.comm;
    if (!d.comm || d.n_neighbors == 0)
      return;
    c.nbr_rank.assign(d.nbr_rank, d.nbr_rank + d.n_neighbors);
    c.send_ptr.assign(d.send_ptr, d.send_ptr + d.n_neighbors + 1);
    c.recv_ptr.assign(d.recv_ptr, d.recv_ptr + d.n_neighbors + 1);
    std::vector<int32_t> si(c.send_ptr.back()), ri(c.recv_ptr.back());
    for (size_t k = 0; k < si.size(); ++k)
      {
        GF_REQUIRE(d.send_dofs[k] >= 0 && d.send_dofs[k] < c.n_ext_owned, GF_ERR_INVALID_ARG,
                   "send_dofs must be owned dofs");
        si[k] = c.h_perm_e2i[d.send_dofs[k]];
      }
    for (size_t k = 0; k < ri.size(); ++k)
      {
        GF_REQUIRE(d.recv_dofs[k] >= c.n_ext_owned && d.recv_dofs[k] < c.n_ext_dofs,
                   GF_ERR_INVALID_ARG, "recv_dofs must be ghost dofs");
        ri[k] = c.h_perm_e2i[d.recv_dofs[k]];
      }
    c.send_idx.upload(si.data(), si.size(), c.stream);
    c.recv_idx.upload(ri.data(), ri.size(), c.stream);
    c.send_buf.alloc(si.size());
    c.recv_buf.alloc(ri.size());
  }

  void allocate_state(gf_context &c)
  {
    cudaStream_t s        = c.stream;
    const bool   nl       = c.model == GF_MODEL_NEO_HOOKEAN;
    const int    first_id = nl ? GF_NL_TOTAL_DISPLACEMENT : GF_LIN_OLD_VELOCITY;
    const int    last_id  = nl ? GF_NL_NEWTON_UPDATE : GF_LIN_BODY_FORCE;
    for (int v = first_id; v <= last_id; ++v)
      c.vec[v].alloc_zero(c.n_local, s);
    c.vec[GF_VEC_SCRATCH0].alloc_zero(c.n_local, s);
    c.vec[GF_VEC_SCRATCH1].alloc_zero(c.n_local, s);
    c.cg_r.alloc_zero(c.n_local, s);
    c.cg_p.alloc_zero(c.n_local, s);
    c.cg_v.alloc_zero(c.n_local, s);
    c.cg_z.alloc_zero(c.n_local, s);
    c.tmp0.alloc_zero(c.n_local, s);
    c.tmp1.alloc_zero(c.n_local, s);
    c.io_buf.alloc(c.n_ext_dofs);
    c.dinv.alloc_zero(size_t(c.n_owned_nodes) * c.dim * c.dim, s);
    c.max_red_blocks = c.sm_count * 8;
    c.cg_scalars.alloc_zero(1, s);
    c.norm_out.alloc_zero(4, s);
    c.err_flag.alloc_zero(1, s);
    GF_CUDA_CHECK(cudaMallocHost((void **)&c.h_scalars, sizeof(gf::CGScalars)));
    GF_CUDA_CHECK(cudaMallocHost((void **)&c.h_norm, 4 * sizeof(double)));
    GF_CUDA_CHECK(cudaMallocHost((void **)&c.h_err, sizeof(int)));
    if (nl)
      c.mat[GF_MAT_TANGENT].val.alloc_zero(c.n_val, s);
    else
      {
        c.mat[GF_MAT_STIFFNESS].val.alloc_zero(c.n_val, s);
        c.mat[GF_MAT_SYSTEM].val.alloc_zero(c.n_val, s);
        c.mass_blk.alloc_zero(c.n_blocks, s);
      }
    // element buffers; K_e is chunked so that huge meshes never materialise all element matrices
    // default budget: 8 GB, or up to 35 % of the memory still free on a large mesh (every extra
    // chunk makes the scatter re-walk the whole value array)
    const char *env       = getenv("GF_KE_BUDGET_MB");
    double      budget_mb = 8192.0;
    if (env)
      budget_mb = atof(env);
    else
      {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
          budget_mb = std::max(budget_mb, 0.35 * double(free_b) / 1048576.0);
      }
    const double per_cell  = double(c.dpc) * c.dpc * 8.0;
    c.ke_chunk_cells =
      std::max<int64_t>(1, std::min<int64_t>(c.n_cells, int64_t(budget_mb * 1048576.0 / per_cell)));
    c.ke_buf.alloc(size_t(c.ke_chunk_cells) * c.dpc * c.dpc);
    if (!nl)
      c.me_buf.alloc(size_t(c.ke_chunk_cells) * c.npc * c.npc);
    c.re_buf.alloc_zero(size_t(c.n_cells) * c.dpc, s);
  }

  // update_acceleration (nonlinear_elasticity.cc:592-599)
  void nl_update_acceleration(gf_context &c)
  {
    const double dt = c.desc.delta_t, beta = c.desc.beta;
    const double alpha_1 = 1. / (beta * std::pow(dt, 2)), alpha_2 = 1. / (beta * dt),
                 alpha_3 = (1 - (2 * beta)) / (2 * beta); // nonlinear_elasticity.h:242-245
    gf::vec_lincomb3(c, c.vec[GF_NL_ACCELERATION].p, alpha_1, c.vec[GF_NL_SOLUTION_DELTA].p,
                     -alpha_2, c.vec[GF_NL_VELOCITY_OLD].p, -alpha_3,
                     c.vec[GF_NL_ACCELERATION_OLD].p);
  }
  // update_velocity (:602-610)
  void nl_update_velocity(gf_context &c)
  {
    const double dt = c.desc.delta_t, beta = c.desc.beta, gamma = c.desc.gamma;
    const double alpha_4 = gamma / (beta * dt), alpha_5 = 1 - (gamma / beta),
                 alpha_6 = (1 - (gamma / (2 * beta))) * dt; // nonlinear_elasticity.h:246-250
    gf::vec_lincomb3(c, c.vec[GF_NL_VELOCITY].p, alpha_4, c.vec[GF_NL_SOLUTION_DELTA].p, alpha_5,
                     c.vec[GF_NL_VELOCITY_OLD].p, alpha_6, c.vec[GF_NL_ACCELERATION_OLD].p);
  }

  int64_t cg_max_iterations(const gf_context &c, double multiplier)
  {
    return int64_t(int(double(c.n_global_dofs_for_maxit) * multiplier));
  }
} // namespace

namespace gf
{
  ProfScope::ProfScope(gf_context &ctx, int kind, int n_kernels)
    : c(ctx)
  {
    Profile &p = *c.prof_sink;
    p.launches[kind]++;
    p.kernel_launches += n_kernels;
    if (!p.enabled || (p.only_kind >= 0 && kind != p.only_kind))
      return;
    if (p.next_event + 2 > p.pool.size())
      {
        if (p.pool.size() >= 16384)
          profile_collect(c);
        if (p.next_event + 2 > p.pool.size())
          for (int k = 0; k < 256; ++k)
            {
              cudaEvent_t e;
              cudaEventCreate(&e);
              p.pool.push_back(e);
            }
      }
    idx = int(p.pending.size());
    p.pending.push_back({kind, int(p.next_event), int(p.next_event + 1)});
    cudaEventRecord(p.pool[p.next_event], c.stream);
    p.next_event += 2;
  }
  ProfScope::~ProfScope()
  {
    if (idx >= 0)
      cudaEventRecord(c.prof_sink->pool[c.prof_sink->pending[idx].e1], c.stream);
  }
  void profile_collect(gf_context &c)
  {
    Profile &p = *c.prof_sink;
    if (p.pending.empty())
      return;
    cudaStreamSynchronize(c.stream);
    for (auto &pe : p.pending)
      {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.pool[pe.e0], p.pool[pe.e1]) == cudaSuccess)
          p.ms[pe.kind] += ms;
      }
    p.pending.clear();
    p.next_event = 0;
  }
} // namespace gf

namespace gf
{
  // ElastoDynamics::assemble_system (linear_elasticity.cc:248-374) for one level
  void lin_assemble(gf_context &c)
  {
    for (int64_t c0 = 0; c0 < c.n_cells; c0 += c.ke_chunk_cells)
      {
        const int64_t c1 = std::min(c.n_cells, c0 + c.ke_chunk_cells);
        launch_lin_cells(c, c0, c1);
        launch_scatter_matrix(c, c.mat[GF_MAT_STIFFNESS].val.p, c0, c1, c0 == 0, false);
        launch_scatter_mass(c, c0, c1, c0 == 0);
      }
    const double dt = c.desc.delta_t, theta = c.desc.theta;
    launch_build_system_matrix(c, dt * dt * theta * theta); // :348-353, :426-451
    launch_build_precond(c, c.mat[GF_MAT_SYSTEM].val.p);
    double bn = 0;
    for (int k = 0; k < 3; ++k)
      bn += c.desc.body_force[k] * c.desc.body_force[k];
    if (std::sqrt(bn) > 1e-15) // body_force_enabled :62
      launch_body_force(c, c.vec[GF_LIN_BODY_FORCE].p);
    c.lin_assembled = true;
    c.direct.factored = false; // a new system matrix: the band Cholesky factor is stale
  }
} // namespace gf

extern "C"
{
  const char *gf_last_error(gf_handle h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

  int gf_create(const gf_desc *d, gf_handle *out)
  {
    if (!d || !out)
      return GF_ERR_INVALID_ARG;
    *out          = nullptr;
    gf_context *c = new gf_context;
    try
      {
        GF_REQUIRE(d->dim == 2 || d->dim == 3, GF_ERR_INVALID_ARG, "dim must be 2 or 3");
        // "Polynomial degree" (parameters.cc:110-113): the shipped files ask for 3 and 4
        // (parameters.prm:21, nonlinear_elasticity.prm:24). Degrees 1 and 2 run the tuned kernels,
        // higher ones the generic-degree kernels; the bound keeps the per-CTA shared-memory
        // tables and the 16-bit scatter offsets in range.
        GF_REQUIRE(d->degree >= 1 && d->degree <= (d->dim == 2 ? GF_MAX_DEGREE_2D : GF_MAX_DEGREE_3D),
                   GF_ERR_UNSUPPORTED, "polynomial degree must be 1..6 (2D) / 1..3 (3D)");
        GF_REQUIRE(d->model == GF_MODEL_LINEAR || d->model == GF_MODEL_NEO_HOOKEAN,
                   GF_ERR_INVALID_ARG, "unknown model");
        GF_REQUIRE(d->n_dofs > 0 && d->n_cells > 0 && d->cell_dofs && d->cell_vertices &&
                     d->constrained,
                   GF_ERR_INVALID_ARG, "empty mesh description");
        GF_REQUIRE(d->model != GF_MODEL_NEO_HOOKEAN || d->data_consistent, GF_ERR_INVALID_ARG,
                   "The neo-Hookean solid doesn't support 'Force' data reading. Please switch to "
                   "'Stress' data on the Fluid side or use the linear model of the solid solver");
        GF_REQUIRE(d->nu < 0.5 && d->delta_t > 0, GF_ERR_INVALID_ARG, "bad material/time parameters");
        int n_dev = 0;
        GF_CUDA_CHECK(cudaGetDeviceCount(&n_dev));
        GF_REQUIRE(d->device >= 0 && d->device < n_dev, GF_ERR_CUDA, "no such CUDA device");
        c->device = d->device;
        GF_CUDA_CHECK(cudaSetDevice(c->device));
        cudaDeviceProp prop;
        GF_CUDA_CHECK(cudaGetDeviceProperties(&prop, c->device));
        c->sm_count = prop.multiProcessorCount;
        GF_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->desc       = *d;
        c->dim        = d->dim;
        c->p          = d->degree;
        c->model      = d->model;
        c->n_cells    = d->n_cells;
        c->n_ext_dofs = d->n_dofs;
        c->n_ext_owned = d->n_owned_dofs > 0 ? d->n_owned_dofs : d->n_dofs;
        gf::build_tables(*c);
        c->npc = c->tables.npc;
        c->dpc = c->tables.dpc;
        gf::build_numbering_and_pattern(*c, *d);
        setup_interface(*c, *d);
        setup_halo(*c, *d);
        gf::comm_setup_context(*c);
        allocate_state(*c);
        gf::setup_lines(*c, *d);
        c->n_global_dofs_for_maxit = c->n_owned;
        if (c->comm)
          {
            *c->h_norm = double(c->n_owned);
            GF_CUDA_CHECK(cudaMemcpyAsync(c->norm_out.p, c->h_norm, sizeof(double),
                                          cudaMemcpyHostToDevice, c->stream));
            gf::allreduce_sum(*c, c->norm_out.p, 1);
            GF_CUDA_CHECK(cudaMemcpyAsync(c->h_norm, c->norm_out.p, sizeof(double),
                                          cudaMemcpyDeviceToHost, c->stream));
            GF_CUDA_CHECK(cudaStreamSynchronize(c->stream));
            c->n_global_dofs_for_maxit = int64_t(*c->h_norm + 0.5);
          }
        gf::build_reduction_plan(*c); // chunk size is a function of the GLOBAL size
        // deferred tangent completion needs the K_e of the whole mesh in ONE chunk on every rank
        {
          double multi = c->ke_chunk_cells < c->n_cells ? 1.0 : 0.0;
          if (c->comm)
            {
              *c->h_norm = multi;
              GF_CUDA_CHECK(cudaMemcpyAsync(c->norm_out.p, c->h_norm, sizeof(double),
                                            cudaMemcpyHostToDevice, c->stream));
              gf::allreduce_sum(*c, c->norm_out.p, 1);
              GF_CUDA_CHECK(cudaMemcpyAsync(c->h_norm, c->norm_out.p, sizeof(double),
                                            cudaMemcpyDeviceToHost, c->stream));
              GF_CUDA_CHECK(cudaStreamSynchronize(c->stream));
              multi = *c->h_norm;
            }
          c->defer_tangent = c->model == GF_MODEL_NEO_HOOKEAN && multi < 0.5 &&
                             getenv("GF_NO_DEFERRED_TANGENT") == nullptr;
        }
        // GF_DIRECT_SOLVER = auto | cholesky | cg presets GF_OPT_DIRECT_SOLVER (0 | 1 | 2)
        if (const char *env = getenv("GF_DIRECT_SOLVER"))
          c->direct_mode = std::string(env) == "cg" ? 2 : (std::string(env) == "cholesky" ? 1 : 0);
        // the pointers inside desc are caller-owned: never dereference them after create
        c->desc.cell_dofs = nullptr;
        c->desc.cell_vertices = nullptr;
        c->desc.constrained = nullptr;
        c->desc.iface_cell = c->desc.iface_face_no = c->desc.iface_dofs = nullptr;
        c->desc.nbr_rank = c->desc.send_dofs = c->desc.recv_dofs = nullptr;
        c->desc.send_ptr = c->desc.recv_ptr = nullptr;
        c->desc.dof_global = nullptr;
        c->desc.line_dof = c->desc.line_master = nullptr;
        c->desc.line_ptr = nullptr;
        c->desc.line_weight = nullptr;
        GF_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        *out = c;
        return GF_OK;
      }
    catch (gf::Error &e)
      {
        g_create_error = e.msg;
        gf_destroy(c);
        return e.code;
      }
    catch (std::exception &e)
      {
        g_create_error = e.what();
        gf_destroy(c);
        return GF_ERR_INVALID_ARG;
      }
  }

  void gf_destroy(gf_handle h)
  {
    if (!h)
      return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    // unlink from a multigrid chain: the coarser levels below stop using this level's stream
    if (h->mg_finer)
      h->mg_finer->mg.coarse = nullptr;
    if (h->mg.coarse)
      {
        h->mg.coarse->mg_finer = nullptr;
        for (gf_context *l = h->mg.coarse; l != nullptr; l = l->mg.coarse)
          if (l->stream == h->stream && !l->owns_stream)
            {
              l->stream    = nullptr; // legacy default stream; only used by its own gf_destroy
              l->prof_sink = &l->prof;
            }
      }
    for (auto e : h->prof.pool)
      cudaEventDestroy(e);
    for (auto e : h->prof.user_events)
      if (e)
        cudaEventDestroy(e);
    if (h->h_scalars)
      cudaFreeHost(h->h_scalars);
    if (h->h_norm)
      cudaFreeHost(h->h_norm);
    if (h->h_err)
      cudaFreeHost(h->h_err);
    if (h->h_lmax)
      cudaFreeHost(h->h_lmax);
    cudaStream_t s = h->owns_stream ? h->stream : nullptr;
    if (h->comm && h->stream)
      {
        cudaStreamSynchronize(h->stream);
        gf::comm_forget_stream(h->comm, h->stream);
      }
    delete h;
    if (s)
      cudaStreamDestroy(s);
  }

  int gf_set_option(gf_handle h, int option, int64_t value)
  {
    return guarded(h, [&](gf_context &c) {
      switch (option)
        {
          case GF_OPT_PRECONDITIONER:
            GF_REQUIRE(value >= GF_PRECOND_NONE && value <= GF_PRECOND_MULTIGRID,
                       GF_ERR_INVALID_ARG, "unknown preconditioner");
            c.precond = int(value);
            if (c.model == GF_MODEL_LINEAR && c.lin_assembled)
              gf::launch_build_precond(c, c.mat[GF_MAT_SYSTEM].val.p);
            break;
          case GF_OPT_CG_CHECK_INTERVAL:
            GF_REQUIRE(value >= 1 && value <= 4096, GF_ERR_INVALID_ARG, "bad check interval");
            c.cg_check_every = int(value);
            break;
          case GF_OPT_PROFILE:
            gf::profile_collect(c);
            c.prof.enabled   = value != 0;
            c.prof.only_kind = value == 2 ? int(gf::Profile::SPMV) : -1;
            break;
          case GF_OPT_OPERATOR:
            GF_REQUIRE(value == 0 || value == 1, GF_ERR_INVALID_ARG, "unknown operator kind");
            GF_REQUIRE(value == 0 || (c.model == GF_MODEL_NEO_HOOKEAN && c.dim == 3 && c.p <= 2 &&
                                      c.affine),
                       GF_ERR_UNSUPPORTED,
                       "the matrix-free operator is available for the 3D neo-Hookean model, "
                       "polynomial degree 1 or 2, parallelepiped cells");
            if (c.operator_kind != int(value))
              {
                c.mf_valid                     = false; // re-assemble before the next solve
                c.mat[GF_MAT_TANGENT].valid    = false;
                c.tangent_pending              = false;
                c.mg_update_pending            = false;
              }
            c.operator_kind = int(value);
            break;
          case GF_OPT_DIRECT_SOLVER:
            GF_REQUIRE(value >= 0 && value <= 2, GF_ERR_INVALID_ARG, "unknown direct solver mode");
            c.direct_mode = int(value);
            break;
          case GF_OPT_SPMV_KERNEL:
            GF_REQUIRE(value == 0 || value == 1 || value == 3 || value == 5 || value == 6,
                       GF_ERR_INVALID_ARG, "unknown SpMV kernel");
            for (gf_context *l = &c; l != nullptr; l = l->mg.coarse)
              l->spmv_kernel_kind = int(value);
            break;
          case GF_OPT_MG_SMOOTHER_DEGREE:
            GF_REQUIRE(value >= 1 && value <= 16, GF_ERR_INVALID_ARG, "bad smoother degree");
            for (gf_context *l = &c; l != nullptr; l = l->mg.coarse)
              l->mg_smoother_degree = int(value);
            break;
          case GF_OPT_MG_COARSE_DEGREE:
            GF_REQUIRE(value >= 1 && value <= 1000, GF_ERR_INVALID_ARG, "bad coarse degree");
            for (gf_context *l = &c; l != nullptr; l = l->mg.coarse)
              l->mg_coarse_degree = int(value);
            break;
          case GF_OPT_MG_SMOOTHER_RATIO:
            GF_REQUIRE(value >= 2 && value <= 1000, GF_ERR_INVALID_ARG, "bad smoother ratio");
            for (gf_context *l = &c; l != nullptr; l = l->mg.coarse)
              l->mg_smoother_ratio = double(value);
            break;
          case GF_OPT_CG_INITIAL_GUESS:
            GF_REQUIRE(value == 0 || value == 1, GF_ERR_INVALID_ARG, "initial guess: 0 or 1");
            c.cg_initial_guess = int(value);
            break;
          case GF_OPT_MG_MATRIX_PRECISION:
            GF_REQUIRE(value >= 0 && value <= 2, GF_ERR_INVALID_ARG,
                       "matrix precision of the V-cycle: 0 (FP64), 1 (FP32 copy) or 2 (all FP32)");
            for (gf_context *l = &c; l != nullptr; l = l->mg.coarse)
              {
                l->mg_matrix_precision = int(value);
                if (value == 0)
                  {
                    l->mg_val32_valid = false;
                    l->mg_val32.release();
                  }
              }
            if (value != 0)
              gf::mg_refresh_f32(c); // operators that are already assembled; else at assembly
            break;
          default:
            throw gf::Error{GF_ERR_INVALID_ARG, "unknown option"};
        }
      return GF_OK;
    });
  }

  int gf_mg_attach(gf_handle fine, gf_handle coarse, const int32_t *child_cells)
  {
    return guarded(fine, [&](gf_context &c) {
      GF_REQUIRE(coarse != nullptr, GF_ERR_INVALID_ARG, "null coarse handle");
      gf::mg_attach(c, *coarse, child_cells);
      coarse->mg_finer = &c;
      return GF_OK;
    });
  }

  int gf_mg_vcycle(gf_handle h, int which_b, int which_x)
  {
    return guarded(h, [&](gf_context &c) {
      double *b = vec_ptr(c, which_b), *x = vec_ptr(c, which_x);
      GF_REQUIRE(b != x, GF_ERR_INVALID_ARG, "b and x must differ");
      finish_tangent(c);
      GF_REQUIRE(c.mg.coarse != nullptr && c.mg_ops_valid, GF_ERR_INVALID_ARG,
                 "multigrid hierarchy not attached / operators not assembled");
      gf::mg_vcycle(c, b, x);
      GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
      return GF_OK;
    });
  }

  // ------------------------------------------------------------------------------------------
  int gf_set_traction(gf_handle h, const double *iface_buf)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(iface_buf != nullptr, GF_ERR_INVALID_ARG, "null buffer");
      gf::iface_scatter(c, iface_buf,
                        vec_ptr(c, c.model == GF_MODEL_NEO_HOOKEAN ? GF_NL_EXTERNAL_STRESS :
                                                                     GF_LIN_STRESS));
      return GF_OK;
    });
  }
  int gf_get_interface_displacement(gf_handle h, double *iface_buf)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(iface_buf != nullptr, GF_ERR_INVALID_ARG, "null buffer");
      gf::iface_gather(c,
                       vec_ptr(c, c.model == GF_MODEL_NEO_HOOKEAN ? GF_NL_TOTAL_DISPLACEMENT :
                                                                    GF_LIN_DISPLACEMENT),
                       iface_buf);
      return GF_OK;
    });
  }
  int gf_state_save(gf_handle h)
  {
    return guarded(h, [&](gf_context &c) {
      const bool nl = c.model == GF_MODEL_NEO_HOOKEAN;
      const int  n = nl ? 6 : 5, first = nl ? GF_NL_TOTAL_DISPLACEMENT : GF_LIN_OLD_VELOCITY;
      gf::ProfScope ps(c, gf::Profile::UPDATE);
      for (int i = 0; i < n; ++i)
        {
          if (!c.saved[i].p)
            c.saved[i].alloc(c.n_local);
          GF_CUDA_CHECK(cudaMemcpyAsync(c.saved[i].p, c.vec[first + i].p,
                                        c.n_local * sizeof(double), cudaMemcpyDeviceToDevice,
                                        c.stream));
        }
      // hidden solver state: the warm-start vector of the smoothers' eigenvalue estimate, so
      // that a restored window replays bit for bit
      for (gf_context *l = &c; l != nullptr; l = l->mg.coarse)
        if (l->mg_e.p)
          {
            if (!l->mg_e_saved.p)
              l->mg_e_saved.alloc(l->n_local);
            GF_CUDA_CHECK(cudaMemcpyAsync(l->mg_e_saved.p, l->mg_e.p, l->n_local * sizeof(double),
                                          cudaMemcpyDeviceToDevice, c.stream));
            l->mg_e_saved_valid = l->mg_e_valid;
          }
      c.has_saved        = true;
      return GF_OK;
    });
  }
  int gf_state_restore(gf_handle h)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(c.has_saved, GF_ERR_INVALID_ARG,
                 "state_variables are not the same as previously saved.");
      const bool nl = c.model == GF_MODEL_NEO_HOOKEAN;
      const int  n = nl ? 6 : 5, first = nl ? GF_NL_TOTAL_DISPLACEMENT : GF_LIN_OLD_VELOCITY;
      gf::ProfScope ps(c, gf::Profile::UPDATE);
      for (int i = 0; i < n; ++i)
        GF_CUDA_CHECK(cudaMemcpyAsync(c.vec[first + i].p, c.saved[i].p, c.n_local * sizeof(double),
                                      cudaMemcpyDeviceToDevice, c.stream));
      for (gf_context *l = &c; l != nullptr; l = l->mg.coarse)
        if (l->mg_e.p && l->mg_e_saved.p)
          {
            GF_CUDA_CHECK(cudaMemcpyAsync(l->mg_e.p, l->mg_e_saved.p, l->n_local * sizeof(double),
                                          cudaMemcpyDeviceToDevice, c.stream));
            l->mg_e_valid = l->mg_e_saved_valid;
          }
      return GF_OK;
    });
  }

  // ------------------------------------------------------------------------------------------
  int gf_nl_begin_step(gf_handle h)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(c.model == GF_MODEL_NEO_HOOKEAN, GF_ERR_INVALID_ARG, "not a neo-Hookean handle");
      gf::vec_zero(c, c.vec[GF_NL_SOLUTION_DELTA].p);
      gf::vec_zero(c, c.vec[GF_NL_NEWTON_UPDATE].p);
      return GF_OK;
    });
  }

  int gf_nl_newton_assemble(gf_handle h, double *res_abs)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(c.model == GF_MODEL_NEO_HOOKEAN, GF_ERR_INVALID_ARG, "not a neo-Hookean handle");
      nl_update_acceleration(c); // :444
      // solution_total = total_displacement + solution_delta (:1062-1063, :580-588)
      gf::vec_copy(c, c.tmp0.p, c.vec[GF_NL_TOTAL_DISPLACEMENT].p);
      gf::vec_axpby(c, c.tmp0.p, 1.0, c.vec[GF_NL_SOLUTION_DELTA].p, 1.0);
      GF_CUDA_CHECK(cudaMemsetAsync(c.err_flag.p, 0, sizeof(int), c.stream));
      double *K = c.mat[GF_MAT_TANGENT].val.p;
      c.tangent_pending   = false;
      c.mg_update_pending = false;
      if (c.operator_kind == 1)
        // matrix-free: quadrature-point data, r_e and the diagonal blocks; no element matrices
        gf::mf_setup(c, c.tmp0.p, c.vec[GF_NL_ACCELERATION].p);
      else
        {
          if (c.defer_tangent)
            {
              gf::launch_nl_cells(c, c.tmp0.p, c.vec[GF_NL_ACCELERATION].p, 0, c.n_cells);
              c.mat[GF_MAT_TANGENT].valid = false;
              c.tangent_pending           = true; // finish_tangent() on first use
            }
          else
            {
              for (int64_t c0 = 0; c0 < c.n_cells; c0 += c.ke_chunk_cells)
                {
                  const int64_t c1 = std::min(c.n_cells, c0 + c.ke_chunk_cells);
                  gf::launch_nl_cells(c, c.tmp0.p, c.vec[GF_NL_ACCELERATION].p, c0, c1);
                  gf::launch_scatter_matrix(c, K, c0, c1, c0 == 0, true);
                }
              c.mat[GF_MAT_TANGENT].valid = true;
            }
        }
      gf::launch_nl_faces(c, c.tmp0.p, c.vec[GF_NL_EXTERNAL_STRESS].p);
      gf::launch_scatter_rhs(c, c.vec[GF_NL_SYSTEM_RHS].p, true);
      if (!c.tangent_pending)
        {
          if (c.operator_kind == 0)
            gf::launch_build_precond(c, K);
          if (gf::mg_active(c))
            gf::mg_update_operators(c, c.tmp0.p); // coarse tangents at the injected state + smoothers
        }
      if (c.lines.n > 0)
        {
          // condensed right-hand side C^T r (distribute_local_to_global :769-773)
          gf::lines_condense(c, c.vec[GF_NL_SYSTEM_RHS].p);
          gf::vec_zero_constrained(c, c.vec[GF_NL_SYSTEM_RHS].p);
        }
      const double r = gf::vec_masked_norm(c, c.vec[GF_NL_SYSTEM_RHS].p, true); // :449
      GF_CUDA_CHECK(
        cudaMemcpyAsync(c.h_err, c.err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
      GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
      if (c.comm)
        {
          // every rank must agree on the error state
          *c.h_norm = double(*c.h_err);
          GF_CUDA_CHECK(cudaMemcpyAsync(c.norm_out.p, c.h_norm, sizeof(double),
                                        cudaMemcpyHostToDevice, c.stream));
          gf::allreduce_sum(c, c.norm_out.p, 1);
          GF_CUDA_CHECK(cudaMemcpyAsync(c.h_norm, c.norm_out.p, sizeof(double),
                                        cudaMemcpyDeviceToHost, c.stream));
          GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
          *c.h_err = *c.h_norm > 0.5;
        }
      GF_REQUIRE(*c.h_err == 0, GF_ERR_DET_F,
                 "det F <= 0 at a quadrature point (Assert nonlinear_elasticity.cc:935)");
      if (res_abs)
        *res_abs = r;
      return GF_OK;
    });
  }

  int gf_nl_newton_solve(gf_handle h, int type_lin, double tol_lin, double max_iterations_lin,
                         uint32_t *lin_it, double *lin_res, double *upd_abs)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(c.model == GF_MODEL_NEO_HOOKEAN, GF_ERR_INVALID_ARG, "not a neo-Hookean handle");
      GF_REQUIRE(type_lin == 0 || type_lin == 1, GF_ERR_INVALID_ARG,
                 "Linear solver type not implemented");
      finish_tangent(c); // scatter + preconditioner + multigrid update of the last assembly
      double * x   = c.vec[GF_NL_NEWTON_UPDATE].p;
      uint32_t it  = 0;
      double   res = 0;
      int      rc;
      if (type_lin == 0)
        rc = gf::cg_solve(c, c.mat[GF_MAT_TANGENT].val.p, x, c.vec[GF_NL_SYSTEM_RHS].p, tol_lin,
                          true, cg_max_iterations(c, max_iterations_lin), &it, &res);
      else
        {
          // SparseDirectUMFPACK (:1192-1200): band Cholesky of the tangent on the device where
          // the band fits (direct.cu), else the CG run to 1e-13 from a zero guess
          rc = GF_OK;
          if (!direct_apply(c, c.mat[GF_MAT_TANGENT].val.p, c.vec[GF_NL_SYSTEM_RHS].p, x, false))
            {
              gf::vec_zero(c, x);
              rc = gf::cg_solve(c, c.mat[GF_MAT_TANGENT].val.p, x, c.vec[GF_NL_SYSTEM_RHS].p, 1e-13,
                                true, 10 * cg_max_iterations(c, std::max(1.0, max_iterations_lin)),
                                &it, &res);
            }
          it  = 1; // :1198-1199
          res = 0.0;
        }
      if (lin_it)
        *lin_it = it;
      if (lin_res)
        *lin_res = res;
      if (rc != GF_OK)
        throw gf::Error{rc, "Iterative method reported convergence failure in step " +
                              std::to_string(it) + ". The residual in the last step was " +
                              std::to_string(res) + "."};
      gf::vec_zero_constrained(c, x); // constraints.distribute :1208
      if (c.comm)
        gf::halo_exchange(c, x);
      const double u = gf::vec_masked_norm(c, x, true); // :476
      gf::lines_distribute(c, x);                        // hanging nodes: constraints.distribute
      if (upd_abs)
        *upd_abs = u;
      gf::vec_axpby(c, c.vec[GF_NL_SOLUTION_DELTA].p, 1.0, x, 1.0); // :487
      return GF_OK;
    });
  }

  int gf_nl_end_step(gf_handle h)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(c.model == GF_MODEL_NEO_HOOKEAN, GF_ERR_INVALID_ARG, "not a neo-Hookean handle");
      gf::vec_axpby(c, c.vec[GF_NL_TOTAL_DISPLACEMENT].p, 1.0, c.vec[GF_NL_SOLUTION_DELTA].p,
                    1.0);       // :139
      nl_update_acceleration(c); // :142
      nl_update_velocity(c);     // :143
      gf::vec_copy(c, c.vec[GF_NL_TOTAL_DISPLACEMENT_OLD].p, c.vec[GF_NL_TOTAL_DISPLACEMENT].p);
      gf::vec_copy(c, c.vec[GF_NL_VELOCITY_OLD].p, c.vec[GF_NL_VELOCITY].p); // :619-621
      gf::vec_copy(c, c.vec[GF_NL_ACCELERATION_OLD].p, c.vec[GF_NL_ACCELERATION].p);
      return GF_OK;
    });
  }

  // ------------------------------------------------------------------------------------------
  int gf_lin_assemble_once(gf_handle h)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(c.model == GF_MODEL_LINEAR, GF_ERR_INVALID_ARG, "not a linear handle");
      gf::lin_assemble(c);
      gf::mg_update_operators(c, nullptr); // coarser levels (if linked) assemble their own A
      GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
      return GF_OK;
    });
  }

  int gf_lin_step(gf_handle h, int type_lin, double max_iterations_lin, uint32_t *lin_it,
                  double *lin_res)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(c.model == GF_MODEL_LINEAR, GF_ERR_INVALID_ARG, "not a linear handle");
      GF_REQUIRE(c.lin_assembled, GF_ERR_INVALID_ARG, "gf_lin_assemble_once has not been called");
      GF_REQUIRE(type_lin == 0 || type_lin == 1, GF_ERR_INVALID_ARG,
                 "Linear solver type not implemented");
      const double dt = c.desc.delta_t, theta = c.desc.theta;
      double *     rhs = c.vec[GF_LIN_SYSTEM_RHS].p;
      double *     vel = c.vec[GF_LIN_VELOCITY].p;
      // ---- assemble_rhs :378-454
      if (c.desc.data_consistent)
        gf::launch_lin_faces(c, c.vec[GF_LIN_STRESS].p, rhs); // :386
      else
        gf::vec_copy(c, rhs, c.vec[GF_LIN_STRESS].p); // :388
      gf::vec_copy(c, c.vec[GF_LIN_OLD_VELOCITY].p, vel); // :390-391
      gf::vec_copy(c, c.vec[GF_LIN_OLD_DISPLACEMENT].p, c.vec[GF_LIN_DISPLACEMENT].p);
      double bn = 0;
      for (int k = 0; k < 3; ++k)
        bn += c.desc.body_force[k] * c.desc.body_force[k];
      if (std::sqrt(bn) > 1e-15)
        gf::vec_axpby(c, rhs, 1.0, c.vec[GF_LIN_BODY_FORCE].p, 1.0); // :394-395
      gf::vec_copy(c, c.tmp0.p, rhs);                                 // tmp = system_rhs :405
      gf::vec_axpby(c, rhs, dt * (1 - theta), c.vec[GF_LIN_OLD_STRESS].p, dt * theta); // :407-408
      gf::vec_copy(c, c.vec[GF_LIN_OLD_STRESS].p, c.tmp0.p);                           // :409
      if (c.comm)
        {
          gf::halo_exchange(c, c.vec[GF_LIN_OLD_VELOCITY].p);
          gf::halo_exchange(c, c.vec[GF_LIN_OLD_DISPLACEMENT].p);
        }
      gf::launch_spmv_mass(c, c.vec[GF_LIN_OLD_VELOCITY].p, c.tmp1.p); // :411-412
      gf::vec_axpby(c, rhs, 1.0, c.tmp1.p, 1.0);
      gf::launch_spmv(c, c.mat[GF_MAT_STIFFNESS].val.p, c.vec[GF_LIN_OLD_VELOCITY].p, c.tmp1.p,
                      nullptr); // :414-417
      gf::vec_axpby(c, rhs, -theta * dt * dt * (1 - theta), c.tmp1.p, 1.0);
      gf::launch_spmv(c, c.mat[GF_MAT_STIFFNESS].val.p, c.vec[GF_LIN_OLD_DISPLACEMENT].p, c.tmp1.p,
                      nullptr); // :419-420
      gf::vec_axpby(c, rhs, -dt, c.tmp1.p, 1.0);
      // MatrixTools::apply_boundary_values with zero values (:448-451): rhs and solution
      if (c.lines.n > 0)
        gf::lines_condense(c, rhs); // hanging_node_constraints.condense(system_rhs) :422
      gf::vec_zero_constrained(c, rhs);
      gf::vec_zero_constrained(c, vel);
      // ---- solve :525-575
      uint32_t it  = 1;
      double   res = 0.0;
      int      rc;
      if (type_lin == 0)
        rc = gf::cg_solve(c, c.mat[GF_MAT_SYSTEM].val.p, vel, rhs, 1.e-10, false,
                          cg_max_iterations(c, max_iterations_lin), &it, &res); // :540-551
      else if (direct_apply(c, c.mat[GF_MAT_SYSTEM].val.p, rhs, vel, true))
        rc = GF_OK; // SparseDirectUMFPACK (:556-563); A is constant: factorised once
      else
        {
          uint32_t it2;
          double   res2;
          gf::vec_zero(c, vel);
          rc = gf::cg_solve(c, c.mat[GF_MAT_SYSTEM].val.p, vel, rhs, 1e-13, true,
                            10 * cg_max_iterations(c, std::max(1.0, max_iterations_lin)), &it2,
                            &res2);
          if (rc != GF_OK)
            {
              it  = it2;
              res = res2;
            }
        }
      if (lin_it)
        *lin_it = it;
      if (lin_res)
        *lin_res = res;
      if (rc != GF_OK)
        throw gf::Error{rc, "Iterative method reported convergence failure in step " +
                              std::to_string(it) + ". The residual in the last step was " +
                              std::to_string(res) + "."};
      if (c.comm)
        gf::halo_exchange(c, vel);
      gf::lines_distribute(c, vel); // hanging_node_constraints.distribute(velocity) :571-572
      // ---- update_displacement :579-586
      gf::vec_axpby(c, c.vec[GF_LIN_DISPLACEMENT].p, dt * theta, vel, 1.0);
      gf::vec_axpby(c, c.vec[GF_LIN_DISPLACEMENT].p, dt * (1 - theta),
                    c.vec[GF_LIN_OLD_VELOCITY].p, 1.0);
      return GF_OK;
    });
  }

  // ------------------------------------------------------------------------------------------
  int gf_get_vector(gf_handle h, int which, double *out)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(out != nullptr, GF_ERR_INVALID_ARG, "null buffer");
      gf::vec_permute_out(c, vec_ptr(c, which), out);
      return GF_OK;
    });
  }
  int gf_set_vector(gf_handle h, int which, const double *in)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(in != nullptr, GF_ERR_INVALID_ARG, "null buffer");
      gf::vec_permute_in(c, in, vec_ptr(c, which));
      return GF_OK;
    });
  }
  int64_t gf_nnz(gf_handle h) { return h ? h->n_blocks * h->dim * h->dim : -1; }

  int gf_export_csr(gf_handle h, int which, int64_t *rowptr, int32_t *col, double *val)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(rowptr && col && val, GF_ERR_INVALID_ARG, "null buffer");
      const int            dim = c.dim;
      std::vector<int32_t> brow(c.n_owned_nodes + 1), bcol(c.bcol.n); // bcol is padded (tiles)
      std::vector<int64_t> vptr(c.n_owned_nodes + 1);
      c.brow_ptr.download(brow.data(), c.stream);
      c.bcol.download(bcol.data(), c.stream);
      c.val_ptr.download(vptr.data(), c.stream);
      std::vector<double> v, mb;
      if (which == GF_MAT_MASS)
        {
          GF_REQUIRE(c.mass_blk.p != nullptr, GF_ERR_INVALID_ARG, "no mass matrix for this model");
          mb.resize(c.n_blocks);
          c.mass_blk.download(mb.data(), c.stream);
        }
      else
        {
          mat_ptr(c, which);
          GF_REQUIRE(which != GF_MAT_TANGENT || c.operator_kind == 0, GF_ERR_INVALID_ARG,
                     "no assembled tangent in matrix-free mode (GF_OPT_OPERATOR = 1)");
          v.resize(c.n_val);
          c.mat[which].val.download(v.data(), c.stream);
        }
      // rows in caller numbering (owned rows [0, n_ext_owned)), ascending caller columns
      std::vector<std::pair<int32_t, double>> row;
      int64_t                                 pos = 0;
      rowptr[0]                                   = 0;
      for (int64_t e = 0; e < c.n_ext_owned; ++e)
        {
          const int32_t i = c.h_perm_e2i[e];
          const int64_t A = i / dim;
          const int     r = i % dim;
          const int     nb = brow[A + 1] - brow[A];
          const int64_t stride = (vptr[A + 1] - vptr[A]) / dim;
          row.clear();
          for (int blk = 0; blk < nb; ++blk)
            for (int cc = 0; cc < dim; ++cc)
              {
                const int32_t ci = bcol[brow[A] + blk] * dim + cc;
                const double  x  = which == GF_MAT_MASS ?
                                     (r == cc ? mb[brow[A] + blk] : 0.0) :
                                     v[vptr[A] + r * stride + blk * dim + cc];
                row.emplace_back(c.h_perm_i2e[ci], x);
              }
          std::sort(row.begin(), row.end(),
                    [](const std::pair<int32_t, double> &a, const std::pair<int32_t, double> &b) {
                      return a.first < b.first;
                    });
          for (auto &kv : row)
            {
              col[pos] = kv.first;
              val[pos] = kv.second;
              ++pos;
            }
          rowptr[e + 1] = pos;
        }
      return GF_OK;
    });
  }

  int gf_export_rows(gf_handle h, int which, int64_t n_rows, const int32_t *rows, int64_t *rowptr,
                     int32_t *col, double *val)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(n_rows >= 0 && rows && rowptr, GF_ERR_INVALID_ARG, "null buffer");
      GF_REQUIRE((col == nullptr) == (val == nullptr), GF_ERR_INVALID_ARG,
                 "col and val must both be given or both be NULL");
      const int dim = c.dim;
      if (which == GF_MAT_MASS)
        GF_REQUIRE(c.mass_blk.p != nullptr, GF_ERR_INVALID_ARG, "no mass matrix for this model");
      else
        {
          mat_ptr(c, which);
          GF_REQUIRE(which != GF_MAT_TANGENT || c.operator_kind == 0, GF_ERR_INVALID_ARG,
                     "no assembled tangent in matrix-free mode (GF_OPT_OPERATOR = 1)");
        }
      GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
      std::vector<int32_t>                    bcol;
      std::vector<double>                     v;
      std::vector<std::pair<int32_t, double>> row;
      int64_t                                 pos = 0;
      rowptr[0]                                   = 0;
      for (int64_t k = 0; k < n_rows; ++k)
        {
          const int32_t e = rows[k];
          GF_REQUIRE(e >= 0 && e < c.n_ext_owned, GF_ERR_INVALID_ARG, "row is not an owned dof");
          const int32_t i = c.h_perm_e2i[e];
          const int64_t A = i / dim;
          const int     r = i % dim;
          int32_t       b[2];
          int64_t       vp[2];
          GF_CUDA_CHECK(cudaMemcpy(b, c.brow_ptr.p + A, sizeof(b), cudaMemcpyDeviceToHost));
          GF_CUDA_CHECK(cudaMemcpy(vp, c.val_ptr.p + A, sizeof(vp), cudaMemcpyDeviceToHost));
          const int nb = b[1] - b[0];
          if (col != nullptr)
            {
              const int64_t stride = (vp[1] - vp[0]) / dim;
              bcol.resize(nb);
              GF_CUDA_CHECK(cudaMemcpy(bcol.data(), c.bcol.p + b[0], nb * sizeof(int32_t),
                                       cudaMemcpyDeviceToHost));
              if (which == GF_MAT_MASS)
                {
                  v.resize(nb);
                  GF_CUDA_CHECK(cudaMemcpy(v.data(), c.mass_blk.p + b[0], nb * sizeof(double),
                                           cudaMemcpyDeviceToHost));
                }
              else
                {
                  v.resize(size_t(nb) * dim);
                  GF_CUDA_CHECK(cudaMemcpy(v.data(), c.mat[which].val.p + vp[0] + r * stride,
                                           size_t(nb) * dim * sizeof(double),
                                           cudaMemcpyDeviceToHost));
                }
              row.clear();
              for (int blk = 0; blk < nb; ++blk)
                for (int cc = 0; cc < dim; ++cc)
                  row.emplace_back(c.h_perm_i2e[bcol[blk] * dim + cc],
                                   which == GF_MAT_MASS ? (r == cc ? v[blk] : 0.0) :
                                                          v[size_t(blk) * dim + cc]);
              std::sort(row.begin(), row.end(),
                        [](const std::pair<int32_t, double> &x, const std::pair<int32_t, double> &y) {
                          return x.first < y.first;
                        });
              for (auto &kv : row)
                {
                  col[pos] = kv.first;
                  val[pos] = kv.second;
                  ++pos;
                }
            }
          else
            pos += int64_t(nb) * dim;
          rowptr[k + 1] = pos;
        }
      return GF_OK;
    });
  }

  // ------------------------------------------------------------------------------------------
  int gf_postprocess(gf_handle h, int which_vector, double *fields)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(fields != nullptr, GF_ERR_INVALID_ARG, "null buffer");
      double *u = vec_ptr(c, which_vector);
      if (c.comm)
        gf::halo_exchange(c, u);
      const int64_t per_cell = int64_t(c.npc) * (c.dim + c.dim * c.dim);
      // staging buffer of at most 256 MB: output volumes of large meshes go out in chunks
      const int64_t chunk =
        std::max<int64_t>(1, std::min<int64_t>(c.n_cells, (int64_t(32) << 20) / per_cell));
      gf::DevBuf<double> stage;
      stage.alloc(size_t(chunk * per_cell));
      GF_CUDA_CHECK(cudaMemsetAsync(c.err_flag.p, 0, sizeof(int), c.stream));
      for (int64_t c0 = 0; c0 < c.n_cells; c0 += chunk)
        {
          const int64_t c1 = std::min(c.n_cells, c0 + chunk);
          gf::launch_postprocess(c, u, c0, c1, stage.p);
          GF_CUDA_CHECK(cudaMemcpyAsync(fields + c0 * per_cell, stage.p,
                                        size_t((c1 - c0) * per_cell) * sizeof(double),
                                        cudaMemcpyDeviceToHost, c.stream));
          GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        }
      GF_CUDA_CHECK(
        cudaMemcpyAsync(c.h_err, c.err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
      GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
      GF_REQUIRE(*c.h_err == 0, GF_ERR_DET_F,
                 "output: the displaced (MappingQEulerian) cell is inverted at a patch point");
      return GF_OK;
    });
  }

  // ------------------------------------------------------------------------------------------
  int gf_spmv(gf_handle h, int which_matrix, int which_x, int which_y)
  {
    return guarded(h, [&](gf_context &c) {
      double *x = vec_ptr(c, which_x), *y = vec_ptr(c, which_y);
      GF_REQUIRE(x != y, GF_ERR_INVALID_ARG, "x and y must differ");
      finish_tangent_local(c);
      if (which_matrix == GF_MAT_MG_F32 && (c.mg_update_pending || !c.mg_val32_valid))
        gf::mg_refresh_f32_level(c); // this level's FP32 copy follows the matrix (rank-local);
                                     // the coarse levels follow with the multigrid update
      if (c.comm)
        gf::halo_exchange(c, x);
      if (which_matrix == GF_MAT_MASS)
        {
          GF_REQUIRE(c.mass_blk.p != nullptr, GF_ERR_INVALID_ARG, "no mass matrix for this model");
          gf::launch_spmv_mass(c, x, y);
        }
      else if (which_matrix == GF_MAT_MG_F32)
        {
          GF_REQUIRE(c.mg_val32_valid, GF_ERR_INVALID_ARG,
                     "no FP32 operator copy (GF_OPT_MG_MATRIX_PRECISION = 1, then assemble)");
          // the kernel the V-cycle would launch at the current GF_OPT_MG_MATRIX_PRECISION
          if (c.mg_matrix_precision == 2)
            gf::launch_spmv_f32x(c, c.mg_val32.p, x, y);
          else
            gf::launch_spmv_f32(c, c.mg_val32.p, x, y);
        }
      else if (which_matrix == GF_MAT_TANGENT)
        gf::op_apply(c, mat_ptr(c, which_matrix), x, y, nullptr); // assembled or matrix-free
      else
        gf::launch_spmv(c, mat_ptr(c, which_matrix), x, y, nullptr);
      GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
      return GF_OK;
    });
  }

  int gf_spmv_timed(gf_handle h, int which_matrix, int n_reps, double *ms_per_launch,
                    double *bytes_per_launch)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(n_reps >= 1, GF_ERR_INVALID_ARG, "n_reps must be >= 1");
      finish_tangent_local(c); // rank-local: bench.py times the SpMV on rank 0 only
      if (which_matrix == GF_MAT_MG_F32 && (c.mg_update_pending || !c.mg_val32_valid))
        gf::mg_refresh_f32_level(c);
      const bool f32 = which_matrix == GF_MAT_MG_F32;
      GF_REQUIRE(!f32 || c.mg_val32_valid, GF_ERR_INVALID_ARG,
                 "no FP32 operator copy (GF_OPT_MG_MATRIX_PRECISION = 1, then assemble)");
      double *A = f32 ? nullptr : mat_ptr(c, which_matrix);
      double *x = vec_ptr(c, GF_VEC_SCRATCH0), *y = vec_ptr(c, GF_VEC_SCRATCH1);
      cudaEvent_t e0, e1;
      GF_CUDA_CHECK(cudaEventCreate(&e0));
      GF_CUDA_CHECK(cudaEventCreate(&e1));
      const bool prof = c.prof.enabled;
      c.prof.enabled  = false;
      const bool tangent = which_matrix == GF_MAT_TANGENT;
      auto       apply   = [&]() {
        if (f32 && c.mg_matrix_precision == 2)
          gf::launch_spmv_f32x(c, c.mg_val32.p, x, y);
        else if (f32)
          gf::launch_spmv_f32(c, c.mg_val32.p, x, y);
        else if (tangent)
          gf::op_apply(c, A, x, y, nullptr);
        else
          gf::launch_spmv(c, A, x, y, nullptr);
      };
      for (int k = 0; k < 3; ++k)
        apply();
      GF_CUDA_CHECK(cudaEventRecord(e0, c.stream));
      for (int k = 0; k < n_reps; ++k)
        apply();
      GF_CUDA_CHECK(cudaEventRecord(e1, c.stream));
      GF_CUDA_CHECK(cudaEventSynchronize(e1));
      c.prof.enabled = prof;
      float ms       = 0;
      GF_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      if (ms_per_launch)
        *ms_per_launch = double(ms) / n_reps;
      if (bytes_per_launch)
        *bytes_per_launch = f32 ? gf::spmv_bytes_f32(c) :
                                  ((tangent && c.operator_kind == 1) ? gf::mf_bytes(c) :
                                                                       gf::spmv_bytes(c));
      return GF_OK;
    });
  }

  int gf_comm_timed(gf_handle h, int n_reps, double *halo_us, double *allreduce_us)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(n_reps >= 1, GF_ERR_INVALID_ARG, "n_reps must be >= 1");
      GF_REQUIRE(c.comm != nullptr, GF_ERR_INVALID_ARG, "handle has no communicator");
      double *    x = vec_ptr(c, GF_VEC_SCRATCH0);
      cudaEvent_t e0, e1, e2;
      GF_CUDA_CHECK(cudaEventCreate(&e0));
      GF_CUDA_CHECK(cudaEventCreate(&e1));
      GF_CUDA_CHECK(cudaEventCreate(&e2));
      const bool prof = c.prof.enabled;
      c.prof.enabled  = false;
      for (int k = 0; k < 5; ++k)
        {
          gf::halo_exchange(c, x);
          gf::allreduce_sum(c, c.norm_out.p + 2, 2);
        }
      GF_CUDA_CHECK(cudaEventRecord(e0, c.stream));
      for (int k = 0; k < n_reps; ++k)
        gf::halo_exchange(c, x);
      GF_CUDA_CHECK(cudaEventRecord(e1, c.stream));
      for (int k = 0; k < n_reps; ++k)
        gf::allreduce_sum(c, c.norm_out.p + 2, 2);
      GF_CUDA_CHECK(cudaEventRecord(e2, c.stream));
      GF_CUDA_CHECK(cudaEventSynchronize(e2));
      gf::comm_check(c);
      c.prof.enabled = prof;
      float ms_h = 0, ms_a = 0;
      GF_CUDA_CHECK(cudaEventElapsedTime(&ms_h, e0, e1));
      GF_CUDA_CHECK(cudaEventElapsedTime(&ms_a, e1, e2));
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      cudaEventDestroy(e2);
      if (halo_us)
        *halo_us = 1e3 * double(ms_h) / n_reps;
      if (allreduce_us)
        *allreduce_us = 1e3 * double(ms_a) / n_reps;
      return GF_OK;
    });
  }

  int gf_direct_info(gf_handle h, int64_t *n_solves, int64_t *half_bandwidth,
                     double *last_residual)
  {
    return guarded(h, [&](gf_context &c) {
      const bool ok = c.operator_kind == 0 && gf::direct_available(c);
      GF_REQUIRE(ok, GF_ERR_UNSUPPORTED,
                 "direct solver not in use: " +
                   (c.direct_mode == 2 ? std::string("GF_OPT_DIRECT_SOLVER = 2") :
                                         (c.operator_kind != 0 ? std::string("matrix-free operator") :
                                                                 c.direct.why)));
      if (n_solves)
        *n_solves = c.direct.n_solves;
      if (half_bandwidth)
        *half_bandwidth = c.direct.w;
      if (last_residual)
        *last_residual = c.direct.last_residual;
      return GF_OK;
    });
  }

  int gf_profile_get(gf_handle h, gf_profile *out, int reset)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(out != nullptr, GF_ERR_INVALID_ARG, "null buffer");
      gf::profile_collect(c);
      const gf::Profile &p         = c.prof;
      out->assemble_cells_ms       = p.ms[gf::Profile::ASM_CELLS];
      out->assemble_faces_ms       = p.ms[gf::Profile::ASM_FACES];
      out->scatter_ms              = p.ms[gf::Profile::SCATTER];
      out->spmv_ms                 = p.ms[gf::Profile::SPMV];
      out->cg_vector_ms            = p.ms[gf::Profile::CG_VEC];
      out->update_ms               = p.ms[gf::Profile::UPDATE];
      out->halo_ms                 = p.ms[gf::Profile::HALO];
      out->assemble_cells_launches = p.launches[gf::Profile::ASM_CELLS];
      out->assemble_faces_launches = p.launches[gf::Profile::ASM_FACES];
      out->scatter_launches        = p.launches[gf::Profile::SCATTER];
      out->spmv_launches           = p.launches[gf::Profile::SPMV];
      out->cg_vector_launches      = p.launches[gf::Profile::CG_VEC];
      out->update_launches         = p.launches[gf::Profile::UPDATE];
      out->halo_launches           = p.launches[gf::Profile::HALO];
      out->kernel_launches         = p.kernel_launches;
      out->mg_spmv_ms              = p.ms[gf::Profile::MG_SPMV];
      out->mg_vector_ms            = p.ms[gf::Profile::MG_VEC];
      out->mg_spmv_launches        = p.launches[gf::Profile::MG_SPMV];
      out->mg_vector_launches      = p.launches[gf::Profile::MG_VEC];
      if (reset)
        {
          for (int k = 0; k < gf::Profile::N_KINDS; ++k)
            {
              c.prof.ms[k]       = 0;
              c.prof.launches[k] = 0;
            }
          c.prof.kernel_launches = 0;
        }
      return GF_OK;
    });
  }

  int gf_event_record(gf_handle h, int slot)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(slot >= 0 && slot < 8, GF_ERR_INVALID_ARG, "event slot out of range");
      if (!c.prof.user_events[slot])
        GF_CUDA_CHECK(cudaEventCreate(&c.prof.user_events[slot]));
      GF_CUDA_CHECK(cudaEventRecord(c.prof.user_events[slot], c.stream));
      return GF_OK;
    });
  }
  int gf_event_elapsed_ms(gf_handle h, int s0, int s1, double *ms)
  {
    return guarded(h, [&](gf_context &c) {
      GF_REQUIRE(s0 >= 0 && s0 < 8 && s1 >= 0 && s1 < 8 && ms && c.prof.user_events[s0] &&
                   c.prof.user_events[s1],
                 GF_ERR_INVALID_ARG, "event slots not recorded");
      GF_CUDA_CHECK(cudaEventSynchronize(c.prof.user_events[s1]));
      float f = 0;
      GF_CUDA_CHECK(cudaEventElapsedTime(&f, c.prof.user_events[s0], c.prof.user_events[s1]));
      *ms = f;
      return GF_OK;
    });
  }

  int gf_synchronize(gf_handle h)
  {
    return guarded(h, [&](gf_context &c) {
      GF_CUDA_CHECK(cudaStreamSynchronize(c.stream));
      gf::comm_check(c);
      return GF_OK;
    });
  }
}
