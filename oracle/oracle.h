/* ORACLE — TEST INFRASTRUCTURE ONLY. NOT PRODUCT CODE.
 *
 * CPU (FP64) restatement of the structural hot path of precice/dealii-adapter, used as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
 * Nothing under dealii_adapter_b200/ may include, link or call this.
 *
 * PARITY PINNED IN PART: the reference ships no tests, golden vectors or fixtures for this path,
 * and as a whole it cannot be built here (deal.II >= 9.2 [CI: 9.5.0] and preCICE >= 3.0 are
 * absent, no network). Three header-only pieces DO compile in place against a small stand-in
 * (oracle/ref_shim, `make ref` -> oracle/_ref/): the neo-Hookean material, the Postprocessor
 * and Time headers in place, and - cut out of the .cc/.h by marker at build time - the
 * cell-assembly structs (tangent, residual, Neumann term), PointHistory, the Newmark coefficients
 * and update/norm members, the linear model's local stiffness loops, consistent-loading face
 * loop, theta-scheme right-hand side and displacement update. Their outputs are committed as
 * tests/golden/reference_vectors.npz and pin the oracle (tests/test_reference_pins.py) and the
 * device (tests/test_gpu_zz_reference_pins.py).
 * What lives INSIDE deal.II - FE tables, distribute_local_to_global, SolverCG/SSOR,
 * apply_boundary_values, create_mass_matrix - remains UNPINNED by the reference itself: deal.II semantics (FE_Q local order, QGauss, QProjector face order,
 * AffineConstraints::distribute_local_to_global, SolverCG, precondition_SSOR,
 * MatrixTools::apply_boundary_values) are restated from the published deal.II 9.5 algorithms.
 * The substitutes for golden vectors are the analytic known-answer tests in
 * tests/test_oracle_*.py (energy finite differences against the reference's own Psi,
 * nonlinear(u=0) == linear cross check, patch / rigid-motion tests) and an independent numpy
 * transcription of the material (tests/ref_formulas.py).
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

  typedef struct
  {
    int32_t dim;    /* 2 | 3 */
    int32_t degree; /* 1 | 2 */
    int32_t model;  /* 0 linear (ElastoDynamics), 1 neo-Hookean (Solid) */
    int64_t n_dofs;
    int64_t n_cells;
    const int32_t *cell_dofs;     /* [n_cells*dpc], FESystem(FE_Q(p),dim) local order */
    const double * cell_vertices; /* [n_cells*2^dim*dim] */
    const uint8_t *constrained;   /* [n_dofs] zero-Dirichlet mask */
    int64_t        n_iface_faces;
    const int32_t *iface_cell;    /* [n_iface_faces] */
    const int32_t *iface_face_no; /* [n_iface_faces] */
    int64_t        n_iface_nodes;
    const int32_t *iface_dofs; /* [dim*n_iface_nodes] component-major ascending lists */
    double         mu, nu, rho;
    double         body_force[3];
    double         beta, gamma, theta, delta_t;
    /* solver parameters (parameters.prm "Solver" subsection) */
    int32_t type_lin; /* 0 CG (+SSOR), 1 Direct stand-in (CG to 1e-14 relative) */
    double  tol_lin;
    double  max_iterations_lin;
    int32_t max_iterations_NR;
    double  tol_f, tol_u;
    int32_t data_consistent; /* parameters.cc:192-200 */
  } orc_desc;

  enum
  {
    /* nonlinear (nonlinear_elasticity.cc:370-375 order first) */
    ORC_NL_TOTAL_DISPLACEMENT = 0,
    ORC_NL_TOTAL_DISPLACEMENT_OLD,
    ORC_NL_VELOCITY,
    ORC_NL_VELOCITY_OLD,
    ORC_NL_ACCELERATION,
    ORC_NL_ACCELERATION_OLD,
    ORC_NL_EXTERNAL_STRESS,
    ORC_NL_SYSTEM_RHS,
    ORC_NL_SOLUTION_DELTA,
    ORC_NL_NEWTON_UPDATE,
    /* linear (linear_elasticity.cc:238-239 order first) */
    ORC_LIN_OLD_VELOCITY = 16,
    ORC_LIN_VELOCITY,
    ORC_LIN_OLD_DISPLACEMENT,
    ORC_LIN_DISPLACEMENT,
    ORC_LIN_OLD_STRESS,
    ORC_LIN_STRESS,
    ORC_LIN_SYSTEM_RHS,
    ORC_LIN_BODY_FORCE
  };
  enum
  {
    ORC_MAT_TANGENT = 0, /* nonlinear tangent_matrix */
    ORC_MAT_STIFFNESS,   /* linear K */
    ORC_MAT_MASS,        /* linear M */
    ORC_MAT_STEPPING,    /* linear M + theta^2 dt^2 K */
    ORC_MAT_SYSTEM       /* linear system_matrix after apply_boundary_values */
  };

  void *      orc_create(const orc_desc *desc);
  void        orc_destroy(void *h);
  const char *orc_last_error(void);

  int64_t orc_nnz(void *h);
  /* sorted-column CSR of the shared sparsity pattern, and the values of one matrix */
  void orc_get_pattern(void *h, int64_t *rowptr, int32_t *col);
  void orc_get_values(void *h, int which_matrix, double *val);
  void orc_get_vector(void *h, int which, double *out);
  void orc_set_vector(void *h, int which, const double *in);

  /* ---- nonlinear (Solid) ---- */
  void   orc_nl_update_acceleration(void *h); /* :592-599 */
  void   orc_nl_update_velocity(void *h);     /* :602-610 */
  void   orc_nl_update_old_variables(void *h); /* :613-622 */
  void   orc_nl_assemble_system(void *h, int n_threads); /* :1044-1087 */
  double orc_nl_error_residual(void *h);                 /* :549-560 */
  /* :1153-1211 ; returns 0 ok, 1 CG not converged */
  int orc_nl_solve_linear_system(void *h, uint32_t *lin_it, double *lin_res);
  /* :410-499 ; table rows written to hist[it*6 + {lin_it,lin_res,res_norm,res_abs,u_norm,u_abs}];
     returns number of Newton iterations performed (linear solves) or -1 if not converged,
     -2 if CG failed. n_assemblies = solves + 1 on convergence. */
  int orc_nl_solve_nonlinear_timestep(void *h, int n_threads, double *hist, int hist_rows);
  /* body of the coupling loop without adapter calls, :121,:138-144 */
  int orc_nl_timestep(void *h, int n_threads, double *hist, int hist_rows);

  /* ---- linear (ElastoDynamics) ---- */
  void orc_lin_assemble_system(void *h); /* linear_elasticity.cc:248-374 */
  void orc_lin_assemble_rhs(void *h);    /* :378-454 */
  int  orc_lin_solve(void *h, uint32_t *lin_it, double *lin_res); /* :525-575 */
  void orc_lin_update_displacement(void *h);                      /* :579-586 */

  /* ---- adapter bodies (adapter.h) ---- */
  void orc_format_precice_to_deal(void *h, const double *read_data_buffer, int which); /* :421-443 */
  void orc_format_deal_to_precice(void *h, int which, double *write_data_buffer);     /* :389-417 */
  void orc_save_state(void *h);   /* :457-462 */
  void orc_reload_state(void *h); /* :482-487 */

  /* ---- output path: output_results (nonlinear:1215-1254, linear:590-629) + Postprocessor
     (postprocessor.h:44-76): per cell and (degree+1)^dim lexicographic patch point,
     fields = [u | strain(d*dim+e)] on the MappingQEulerian (displaced) configuration,
     points = X + u (may be NULL) ---- */
  void orc_postprocess(void *h, int which, double *points, double *fields);

  /* ---- pieces exposed for known-answer tests ---- */
  /* material.h:37-49,62-138 : tau (n_indep) and Jc (n_indep^2) from det_F and b_bar (n_indep);
     deal.II SymmetricTensor component order (00,11,[22],01,[02,12]) */
  void orc_material(int dim, double mu, double nu, double det_F, const double *b_bar,
                    double *psi, double *tau, double *Jc);
  /* element tangent/residual of one cell for given local vectors (no Neumann term) */
  void orc_nl_cell(void *h, int64_t cell, const double *u_local, const double *acc_local,
                   double *cell_matrix, double *cell_rhs);
  /* stand-alone kernels for the CPU baseline */
  void   orc_vmult(void *h, int which_matrix, const double *x, double *y);
  int    orc_threads_available(void);

#ifdef __cplusplus
}
#endif
#endif
