// Internal context of the sm_100a structural hot-path library (C-ABI: include/graft_fem.h).
//
// HBM layout (all FP64 / int32, internal numbering = node blocks: dof = node*dim + comp,
// owned nodes first):
//   vectors      [n_local_dofs]                      state / work vectors, AoS per node
//   BSR matrix   brow_ptr[n_rows+1], bcol[nblocks]   block rows = owned nodes, dim x dim blocks
//                val: per block row A with nb blocks: dim "scalar rows", each nb*dim doubles padded
//                to an even count -> val[val_ptr[A] + r*stride + blk*dim + c]; the three scalar
//                rows of a node share one column-index stream (4 B per dim*dim values)
//   scatter map  for every block the ordered list (ascending cell) of element-matrix blocks
//                that sum into it: deterministic, atomic-free (nonlinear_elasticity.cc:760-774)
//   element buf  K_e row-major [chunk_cells][dpc][dpc], r_e [n_cells][dpc]
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "fe_tables_host.h"
#include "graft_fem.h"

namespace gf
{
  struct Error
  {
    int         code;
    std::string msg;
  };

#define GF_CUDA_CHECK(call)                                                                      \
  do                                                                                             \
    {                                                                                            \
      cudaError_t err__ = (call);                                                                \
      if (err__ != cudaSuccess)                                                                  \
        throw gf::Error{GF_ERR_CUDA, std::string(#call) + " failed: " +                          \
                                       cudaGetErrorString(err__) + " (" + __FILE__ + ":" +       \
                                       std::to_string(__LINE__) + ")"};                          \
    }                                                                                            \
  while (0)

#define GF_REQUIRE(cond, code, message)                                                          \
  do                                                                                             \
    {                                                                                            \
      if (!(cond))                                                                               \
        throw gf::Error{code, std::string(message)};                                             \
    }                                                                                            \
  while (0)

  template <typename T>
  struct DevBuf
  {
    T *    p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release()
    {
      if (p)
        cudaFree(p);
      p = nullptr;
      n = 0;
    }
    void alloc(size_t count)
    {
      release();
      n = count;
      if (count)
        GF_CUDA_CHECK(cudaMalloc((void **)&p, count * sizeof(T)));
    }
    void alloc_zero(size_t count, cudaStream_t s)
    {
      alloc(count);
      if (count)
        GF_CUDA_CHECK(cudaMemsetAsync(p, 0, count * sizeof(T), s));
    }
    void upload(const T *host, size_t count, cudaStream_t s)
    {
      alloc(count);
      if (count)
        {
          GF_CUDA_CHECK(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
          GF_CUDA_CHECK(cudaStreamSynchronize(s));
        }
    }
    void download(T *host, cudaStream_t s) const
    {
      if (n)
        {
          GF_CUDA_CHECK(cudaMemcpyAsync(host, p, n * sizeof(T), cudaMemcpyDeviceToHost, s));
          GF_CUDA_CHECK(cudaStreamSynchronize(s));
        }
    }
  };

  constexpr int MAX_VECTORS = 32;
  constexpr int N_MATRICES  = 4;

  // reference-cell tables (device) shared by all kernels
  struct FETables : FETablesHost
  {
    DevBuf<double> N;    // [nq][npc]
    DevBuf<double> dN;   // [nq][npc][dim]  unit-cell gradients
    DevBuf<double> w;    // [nq]
    DevBuf<double> Nf;   // [2*dim][nqf][npc]
    DevBuf<double> wf;   // [nqf]
    DevBuf<double> Mref; // [npc][npc] sum_q w N_a N_b
    // MappingQ1 vertex shape function gradients (general cells): cell q-points, face q-points,
    // output patch points
    DevBuf<double> dphi, dphif, dphip;
  };

  // hanging-node constraint lines (constraints.cu): internal dof numbering
  struct ConstraintLines
  {
    int64_t         n = 0, n_masters = 0;
    DevBuf<int32_t> dof, master;    // [n], [ptr[n]]
    DevBuf<int64_t> ptr;            // [n + 1]
    DevBuf<double>  weight;
    DevBuf<int32_t> mdof, slave;    // transposed: distinct masters, their constrained dofs
    DevBuf<int64_t> mptr;
    DevBuf<double>  mweight;
    DevBuf<uint8_t> norm_mask;      // Dirichlet or hanging: skipped by the Newton norms
  };

  // band Cholesky for 'Solver type = Direct' (direct.cu, direct_band.cuh)
  struct DirectBand
  {
    bool        analysed = false, usable = false, factored = false;
    bool        factor_is_system_matrix = false; // linear model: A is constant, factor once
    std::string why;                             // why it is not usable
    int64_t     n = 0, w = 0, ld = 0;            // scalar size, half bandwidth, band leading dim
    int64_t     n_solves = 0;                    // solves answered by the band Cholesky
    double      last_residual = 0;               // ||b - A x|| / ||b|| of the last one
    DevBuf<int32_t> node_new;                    // [n_nodes] reverse Cuthill-McKee: old -> new
    DevBuf<double>  band, xp, tmp;
    DevBuf<int>     info;
  };

  // CG scalars living on the device; the host polls `status` every check interval
  struct CGScalars
  {
    double rz, rz_old, pAp, rr, alpha, beta, res, tol, res0;
    int    it, maxit, status; // status: 0 iterate, 1 success, 2 failure
    int    pad;
  };

  // SpMV tile = run of consecutive block rows whose values (<= TILE_V doubles), column indices and
  // per-row records are fetched by three TMA bulk copies into one pipeline stage (spmv.cu)
  struct TileDesc
  {
    long long val_off;   // first value (doubles) in the value array
    int       val_count; // doubles, multiple of 2
    int       col_off;   // first column index, rounded down to a multiple of 4
    int       col_count; // ints, multiple of 4
    int       row0, n_rows, pad;
  };
  constexpr int spmv_tile_v(int dim) { return dim == 3 ? 6144 : 4096; } // value doubles per tile
  constexpr int SPMV_TILE_ROWS = 32;   // max rows per tile
  constexpr int SPMV_META      = SPMV_TILE_ROWS + 2; // uint2 records per tile (+ header), 272 B

  struct BsrMatrix
  {
    DevBuf<double> val;
    bool           valid = false;
  };

  struct Comm; // comm.cu

  // Geometric multigrid (multigrid.cu): data one level keeps about the next COARSER level.
  struct MGTransfer
  {
    gf_context *        coarse = nullptr; // not owned (the host destroys every level handle)
    int                 n_child = 0;      // 2^dim
    DevBuf<int32_t>     parent;           // [n_cells] coarse_cell*n_child + child index
    DevBuf<int32_t>     child_cells;      // [coarse n_cells*n_child] local fine cell or -1
    DevBuf<uint8_t>     first_owner;      // [n_cells*npc] 1: this (cell, a) is the node's first
                                          // appearance AND the node is owned by this rank
    DevBuf<double>      E;                // [n_child][npc][npc] embedding N_b^coarse(node a of child k)
    DevBuf<int32_t>     rl_ptr;           // [npc+1] per coarse local node b: non-zero (k, a) list
    DevBuf<int32_t>     rl_ka;            // k*npc + a
    DevBuf<double>      rl_w;             // E[k][a][b]
    DevBuf<int32_t>     inj;              // [npc] k*npc + a of the fine node coinciding with b
  };

  struct Profile
  {
    enum Kind
    {
      ASM_CELLS = 0,
      ASM_FACES,
      SCATTER,
      SPMV,
      CG_VEC,
      UPDATE,
      HALO,
      MG_SPMV, // SpMV launches on the coarser multigrid levels
      MG_VEC,  // multigrid smoother / transfer vector kernels (all levels)
      N_KINDS
    };
    bool                     enabled = false;
    int                      only_kind = -1; // >= 0: only this kind is bracketed with events
    double                   ms[N_KINDS]{};
    int64_t                  launches[N_KINDS]{};
    int64_t                  kernel_launches = 0;
    cudaEvent_t              user_events[8]{};
    std::vector<cudaEvent_t> pool;
    struct Pending
    {
      int kind;
      int e0, e1;
    };
    std::vector<Pending> pending;
    size_t               next_event = 0;
  };
} // namespace gf

// peer-window layout (comm.cu): flags, all-reduce slots, halo mailboxes, all double-buffered by
// the parity of the communicator-wide epoch
namespace gf
{
  constexpr int    P2P_MAX_RANKS  = 8;
  constexpr int    P2P_AR_MAX     = 8;       // doubles per all-reduce
  constexpr size_t P2P_HALO_FLAG  = 0;       // uint64 [2][P2P_MAX_RANKS] indexed by SENDER rank
  constexpr size_t P2P_AR_FLAG    = 256;     // uint64 [2][P2P_MAX_RANKS]
  constexpr size_t P2P_AR_SLOT    = 1024;    // double [2][P2P_MAX_RANKS][P2P_AR_MAX]
  constexpr size_t P2P_MAILBOX    = 4096;    // double [2][P2P_MAX_RANKS][P2P_HALO_CAP]
  constexpr size_t P2P_HALO_CAP   = size_t(1) << 19; // doubles per (parity, sender): 4 MiB
  // all-gather of the reduction-chunk partial sums (reduce.cuh): every rank stores its segment of
  // the GLOBAL chunk list into every window, double [2][3][P2P_GATHER_MAX]
  constexpr size_t P2P_GATHER_FLAG = 512;    // uint64 [2][P2P_MAX_RANKS]
  constexpr size_t P2P_GATHER_MAX  = 32768;  // chunks of all ranks together
  constexpr size_t P2P_GATHER =
    P2P_MAILBOX + 2 * size_t(P2P_MAX_RANKS) * P2P_HALO_CAP * sizeof(double);
  constexpr size_t P2P_WINDOW_BYTES = P2P_GATHER + 2 * 3 * P2P_GATHER_MAX * sizeof(double);
} // namespace gf

struct gf_comm_s
{
  void *nccl_comm = nullptr;
  int   rank = 0, n_ranks = 1, device = 0;
  // ---- peer windows over NVLink (cudaIpc); p2p == false: NCCL transport ----
  bool           p2p = false;
  unsigned char *win[gf::P2P_MAX_RANKS] = {}; // win[rank] = own allocation, others IPC-mapped
  unsigned long long halo_epoch[gf::P2P_MAX_RANKS] = {}; // per neighbour pair (symmetric counts)
  unsigned long long ar_epoch = 0, gather_epoch = 0;
  bool stream_waits = false; // ranks share a device: flag waits as stream memory operations
  unsigned *     blk_counter = nullptr; // device [P2P_MAX_RANKS]: last-block detection of the push
  int *          h_err = nullptr;       // mapped pinned: set by a kernel whose flag wait timed out
  int *          d_err = nullptr;       // device alias of h_err
  unsigned long long timeout_ns = 30ull * 1000000000ull;
  cudaStream_t   last_stream = nullptr; // ops of one communicator are ordered on one stream
  int64_t        n_halo = 0, n_allreduce = 0; // operations issued (diagnostics)
};

struct gf_context
{
  gf_desc      desc{}; // pointers inside are NOT valid after gf_create
  std::string  last_error;
  cudaStream_t stream = nullptr;
  bool         owns_stream = true; // coarser multigrid levels run on the finest level's stream
  int          dim = 0, p = 0, npc = 0, dpc = 0, model = 0, device = 0;
  int          sm_count = 148;

  int64_t n_ext_dofs = 0;   // caller-visible dofs (owned + ghost)
  int64_t n_ext_owned = 0;  // caller owned dofs
  int64_t n_nodes = 0;      // local nodes (owned + ghost)
  int64_t n_owned_nodes = 0;
  int64_t n_local = 0;      // n_nodes*dim
  int64_t n_owned = 0;      // n_owned_nodes*dim
  int64_t n_cells = 0;
  int64_t n_blocks = 0;     // BSR blocks
  int64_t n_val = 0;        // doubles per BSR value array
  int64_t n_cand = 0;       // scatter sources
  int64_t ke_chunk_cells = 0;
  int64_t n_global_dofs_for_maxit = 0; // global matrix size m() used for CG max iterations

  gf::FETables tables;

  // numbering
  gf::DevBuf<int32_t> perm_e2i; // [n_ext_dofs]
  gf::DevBuf<int32_t> perm_i2e; // [n_local]
  std::vector<int32_t> h_perm_e2i, h_perm_i2e;
  gf::DevBuf<uint8_t> constrained; // [n_local] internal numbering
  gf::DevBuf<int32_t> cell_nodes;  // [n_cells*npc]
  gf::DevBuf<double>  geom;        // [n_cells*(dim*dim+1)] : J^{-1} row-major, det J
  // general (non-affine) cells: the kernels of assemble_nl_generic.cuh / assemble_general.cuh
  // evaluate the MappingQ1 Jacobian per quadrature point from the vertices; geom then holds the
  // Jacobian at vertex 0 (diagnostics only)
  bool                affine = true;
  gf::DevBuf<double>  cell_verts;  // [n_cells][2^dim][dim], only when !affine
  // node -> (cell, local node) adjacency, ascending cell
  gf::DevBuf<int64_t> nc_ptr; // [n_nodes+1]
  gf::DevBuf<int32_t> nc_src; // [n_cells*npc] = cell*npc + a
  // BSR pattern
  gf::DevBuf<int32_t>  brow_ptr; // [n_owned_nodes+1]
  gf::DevBuf<int32_t>  bcol;     // [n_blocks]
  gf::DevBuf<int64_t>  val_ptr;  // [n_owned_nodes+1]
  gf::DevBuf<int64_t>  cand_ptr; // [n_owned_nodes+1]
  gf::DevBuf<int32_t>  row_src;  // [n_cand] = (cell*npc + a)*npc + b
  gf::DevBuf<uint16_t> src_off;  // [n_blocks] offset of the block's sources in its row list
  gf::DevBuf<gf::TileDesc> tile_desc; // [n_tiles]
  gf::DevBuf<uint2>        tile_meta; // [n_tiles*SPMV_META]: per row {val offset, col offset | nb<<16}
  int64_t                  n_tiles = 0; // 0: a row does not fit a tile -> LDG kernel only
  int                      spmv_kernel_kind = 0; // GF_OPT_SPMV_KERNEL: 0 auto, 1 LDG, 2-4 two-ring
                                                 // TMA variants, 5 single-ring TMA (spmv.cu)
  gf::BsrMatrix        mat[gf::N_MATRICES];
  gf::DevBuf<double>   mass_blk; // linear: scalar mass value per block (M = m_ab delta_cd)
  gf::DevBuf<double>   dinv;     // [n_owned_nodes*dim*dim] preconditioner blocks
  // interface
  int64_t             n_iface_nodes = 0, n_iface_cells = 0;
  gf::DevBuf<int32_t> iface_dofs_i;    // [dim*n_iface_nodes] internal dof ids
  gf::DevBuf<int32_t> iface_cell_list; // [n_iface_cells] cells owning >=1 interface face
  gf::DevBuf<int32_t> iface_face_ptr;  // [n_iface_cells+1]
  gf::DevBuf<int32_t> iface_face_no;   // faces grouped by cell, ascending face number
  gf::DevBuf<double>  iface_buf;       // [dim*n_iface_nodes] staging
  // element buffers
  gf::DevBuf<double> ke_buf; // [ke_chunk_cells*dpc*dpc]
  gf::DevBuf<double> me_buf; // linear: scalar mass blocks [ke_chunk_cells*npc*npc]
  gf::DevBuf<double> re_buf; // [n_cells*dpc]
  // vectors (internal numbering)
  gf::DevBuf<double> vec[gf::MAX_VECTORS];
  gf::DevBuf<double> saved[6];
  bool               has_saved = false;
  gf::DevBuf<double> cg_r, cg_p, cg_v, cg_z, tmp0, tmp1;
  gf::DevBuf<double> io_buf; // [n_ext_dofs] permutation staging
  // reductions
  gf::DevBuf<double>        partials; // [3*max_blocks]
  gf::DevBuf<gf::CGScalars> cg_scalars;
  gf::CGScalars *           h_scalars = nullptr; // pinned
  gf::DevBuf<double>        norm_out;            // [4]
  double *                  h_norm = nullptr;    // pinned [4]
  gf::DevBuf<int>           err_flag;            // det F <= 0 detection
  int *                     h_err = nullptr;     // pinned
  int                       max_red_blocks = 0;

  // options
  int  precond        = GF_PRECOND_BLOCK_JACOBI;
  int  cg_check_every = 32;
  int  cg_initial_guess = 1; // GF_OPT_CG_INITIAL_GUESS
  int  operator_kind  = 0;
  bool lin_assembled  = false;
  gf::ConstraintLines lines;   // gf_desc.line_*: hanging-node constraints (empty on box meshes)
  int  direct_mode    = 0;     // GF_OPT_DIRECT_SOLVER: 0 auto, 1 band Cholesky or error, 2 tight CG
  gf::DirectBand direct;
  bool defer_tangent  = false; // scatter / preconditioner / multigrid update on first use
  bool tangent_pending = false; // K_e of the last assembly not yet scattered (api.cu)
  bool mg_update_pending = false; // tangent scattered, multigrid update (collective) still due

  // partition-independent reductions (pattern.cu, reduce.cu): owned nodes are grouped into node
  // planes orthogonal to the slab axis (contiguous in the internal numbering) and into chunks of
  // <= red_chunk_nodes nodes that never cross a plane; every dot product / norm is the fixed tree
  // over the GLOBAL chunk list, so its bits do not depend on the number of ranks
  int                  slab_axis = 0;  // gf_desc.slab_axis (0: not given -> one plane per rank)
  int                  axis_dir  = -1; // reference-cell direction of the slab axis
  std::vector<int32_t> h_plane_ptr;    // [n_planes+1] owned node ranges
  gf::DevBuf<int64_t>  node_gkey;      // [n_nodes] partition-independent id of a node
  gf::DevBuf<int32_t>  red_chunk_ptr;  // [n_red_chunks+1] owned node ranges
  int                  red_chunk_nodes = 1024;
  int                  n_red_chunks = 0;        // this rank
  int                  red_stride = 0;          // distance between the partial sums of two quantities
  int64_t              red_chunk_base = 0;      // global index of this rank's first chunk
  int64_t              n_red_chunks_global = 0;
  std::vector<int>     red_rank_chunks;         // chunks per rank (NCCL transport: padded all-gather)
  gf::DevBuf<double>   red_gather;              // NCCL transport: gathered partials
  gf::DevBuf<int>      red_rank_off;            // NCCL transport: chunk offsets of the ranks

  // multi-GPU
  gf_comm             comm = nullptr;
  std::vector<int>    nbr_rank;
  std::vector<int64_t> send_ptr, recv_ptr;
  gf::DevBuf<int32_t> send_idx, recv_idx; // internal dof ids
  gf::DevBuf<double>  send_buf, recv_buf;
  bool                halo_p2p = false; // this handle's halo lists fit the peer mailboxes

  // matrix-free tangent operator (matfree.cu), GF_OPT_OPERATOR = 1
  gf::DevBuf<double> mf_qp;     // [n_cells][NF][nq]: C = J^-1 F^-1, JxW*Jc (upper Voigt), JxW*tau
  gf::DevBuf<double> mf_ye;     // [n_cells][dpc] element results of one operator application
  gf::DevBuf<double> mf_diag_e; // [n_cells][npc][dim*dim] diagonal node blocks of K_e
  gf::DevBuf<double> mf_cdiag;  // [n_local] diagonal of the constrained dofs (sum |K_e(i,i)|)
  bool               mf_valid = false;

  // geometric multigrid preconditioner (multigrid.cu)
  gf::MGTransfer     mg;             // link to the next coarser level (mg.coarse == nullptr: none)
  gf_context *       mg_finer = nullptr; // back link (for safe destruction in any order)
  int                mg_level = 0;   // 0 = finest
  gf::DevBuf<double> mg_b, mg_x, mg_r, mg_d, mg_v, mg_e; // rhs, solution, residual, direction, A d, eigvec
  gf::DevBuf<double> mg_lmax_dev;    // estimate of lambda_max(D^-1 A) of the current operator: lives
                                     // on the device, never waited for by the host
  double *           h_lmax = nullptr; // pinned copy (async), validated at the next CG poll
  bool               mg_ops_valid = false; // operators and smoothers built since the attach
  bool               mg_e_valid = false;
  gf::DevBuf<double> mg_e_saved;     // checkpoint copy of mg_e (gf_state_save / gf_state_restore)
  bool               mg_e_saved_valid = false;
  // FP32 copy of this level's operator for the V-cycle (GF_OPT_MG_MATRIX_PRECISION = 1)
  int                mg_matrix_precision = 0; // 0: FP64 values; 1: FP32 copy inside the V-cycle;
                                              // 2: FP32 copy, x staged / accumulated in FP32 too
  gf::DevBuf<float>  mg_val32;                // [n_val + 4]
  bool               mg_val32_valid = false;
  int                mg_smoother_degree = 3, mg_coarse_degree = 80;
  double             mg_smoother_ratio = 40.0, mg_coarse_ratio = 1000.0;

  // single-launch coarsest-level solver (coarse_solve.cu)
  bool   cs_planned = false, cs_enabled = false, cs_stage = false;
  int    cs_rows_per_cta = 0, cs_grid = 0;
  size_t cs_smem = 0;
  gf::DevBuf<unsigned long long> cs_counter;
  unsigned long long             cs_arrivals = 0;

  // output-path post-processing (postprocess.cu): shape tables at the patch points
  gf::DevBuf<double> pp_N, pp_dN;

  gf::Profile  prof;
  gf::Profile *prof_sink = &prof; // coarser multigrid levels account into the finest level
};

namespace gf
{
  // fe_tables.cu
  void build_tables(gf_context &c);
  // pattern.cu
  void build_numbering_and_pattern(gf_context &c, const gf_desc &d);
  // assemble_nl.cu
  void launch_nl_cells(gf_context &c, const double *u_total, const double *accel, int64_t c0,
                       int64_t c1);
  void launch_nl_faces(gf_context &c, const double *u_total, const double *stress);
  // assemble_lin.cu
  void launch_lin_cells(gf_context &c, int64_t c0, int64_t c1);
  void launch_lin_faces(gf_context &c, const double *stress, double *rhs_out);
  void launch_body_force(gf_context &c, double *out);
  // scatter.cu
  void launch_scatter_matrix(gf_context &c, double *val, int64_t c0, int64_t c1, bool first,
                             bool apply_constraints);
  void launch_scatter_mass(gf_context &c, int64_t c0, int64_t c1, bool first);
  void launch_scatter_rhs(gf_context &c, double *rhs, bool mask_constrained);
  void launch_build_system_matrix(gf_context &c, double factor);
  void launch_build_precond(gf_context &c, const double *val);
  // spmv.cu
  void launch_spmv(gf_context &c, const double *val, const double *x, double *y,
                   double *dot_partials);
  void launch_spmv_f32(gf_context &c, const float *val32, const double *x, double *y);
  void launch_spmv_f32x(gf_context &c, const float *val32, const double *x, double *y);
  void launch_convert_f32(gf_context &c, const double *val, float *val32);
  void launch_spmv_mass(gf_context &c, const double *x, double *y);
  double spmv_bytes(const gf_context &c);
  double spmv_bytes_f32(const gf_context &c);
  int    spmv_dot_partials(const gf_context &c);
  // cg.cu
  int cg_solve(gf_context &c, const double *val, double *x, const double *b, double tol,
               bool tol_relative_to_rhs, int64_t maxit, uint32_t *last_step, double *last_value);
  int cg_solve_direct_precond(gf_context &c, const double *val, double *x, const double *b,
                              double tol, int64_t maxit, uint32_t *last_step, double *last_value);
  // vector_ops.cu
  void   vec_permute_in(gf_context &c, const double *ext_host, double *dst);
  void   vec_permute_out(gf_context &c, const double *src, double *ext_host);
  void   vec_zero(gf_context &c, double *v);
  void   vec_copy(gf_context &c, double *dst, const double *src);
  void   vec_axpby(gf_context &c, double *y, double a, const double *x, double b); // y = a x + b y
  void   vec_lincomb3(gf_context &c, double *out, double a, const double *x, double b,
                      const double *y, double cc, const double *z);
  double vec_masked_norm(gf_context &c, const double *v, bool mask_constrained);
  void   vec_zero_constrained(gf_context &c, double *v);
  void   iface_scatter(gf_context &c, const double *host_buf, double *vec);
  void   iface_gather(gf_context &c, const double *vec, double *host_buf);
  // api.cu: body of ElastoDynamics::assemble_system for one level
  void lin_assemble(gf_context &c);
  // matfree.cu
  void mf_setup(gf_context &c, const double *u_total, const double *accel); // qp data, r_e, D^-1
  void mf_apply(gf_context &c, const double *x, double *y, double *dot_partials);
  int  mf_dot_partials(const gf_context &c);
  double mf_bytes(const gf_context &c);
  // operator.cu: operator of the CG / smoothers on this level (assembled SpMV or the matrix-free
  // tangent, condensed where the handle has hanging-node constraint lines)
  void   op_apply(gf_context &c, const double *val, const double *x, double *y,
                  double *dot_partials);
  int    op_dot_partials(const gf_context &c);
  // multigrid.cu
  void mg_attach(gf_context &fine, gf_context &coarse, const int32_t *child_cells);
  void mg_update_operators(gf_context &c, const double *u_total); // after the finest assembly
  void mg_vcycle(gf_context &c, const double *b, double *x);      // x = MG(b)
  void mg_refresh_f32(gf_context &c); // FP32 operator copies of all levels below and incl. c
  void mg_refresh_f32_level(gf_context &c); // this level only (rank-local)
  bool mg_active(const gf_context &c);
  // constraints.cu: hanging-node constraint lines (no-ops without lines)
  void setup_lines(gf_context &c, const gf_desc &d);
  void lines_distribute(gf_context &c, double *x); // x_s = sum w x_m
  void lines_condense(gf_context &c, double *y);   // y_m += sum w y_s ; y_s = 0
  // direct.cu: 'Solver type = Direct'
  bool direct_available(gf_context &c);                  // ordering + memory check (cached)
  bool direct_factor(gf_context &c, const double *A);    // false: not positive definite
  void direct_solve(gf_context &c, const double *b, double *x);
  // coarse_solve.cu
  bool coarse_solve_single_launch(gf_context &c, const double *val, const double *b, double *x,
                                  int degree, double ratio);
  // postprocess.cu
  void launch_postprocess(gf_context &c, const double *u, int64_t c0, int64_t c1,
                          double *fields_dev);
  // comm.cu
  void halo_reduce_add(gf_context &c, double *v); // ghost partial sums -> owners (+=)
  void halo_exchange(gf_context &c, double *v);
  void allreduce_sum(gf_context &c, double *dev_values, int count);
  void allreduce_sum_vector(gf_context &c, double *dev_values, int64_t count); // long vectors
  void comm_setup_context(gf_context &c);  // after setup_halo: agree on the transport (collective)
  void comm_check(gf_context &c);          // throws if a peer-window wait timed out
  void comm_forget_stream(gf_comm cm, cudaStream_t s); // before a stream is destroyed
  // profile helpers
  struct ProfScope
  {
    gf_context &c;
    int         idx = -1;
    ProfScope(gf_context &ctx, int kind, int n_kernels = 1);
    ~ProfScope();
  };
  void profile_collect(gf_context &c);
} // namespace gf
