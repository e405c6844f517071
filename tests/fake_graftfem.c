/* TEST INFRASTRUCTURE ONLY — a scripted stand-in for libgraftfem.so, LD_PRELOADed under the C++
 * host driver (elasticity_2d / elasticity_3d) so that the host mirror of Solid / ElastoDynamics /
 * Adapter / Time can be exercised on a machine WITHOUT a GPU: every C-ABI call is logged to
 * $GF_FAKE_LOG, the Newton residual / update norms come from $GF_FAKE_SCRIPT (two lines of
 * numbers, consumed in order). No arithmetic of the hot path happens here and the product never
 * links this file. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "graft_fem.h"

struct gf_context
{
  int     dim, degree, model;
  int64_t n_cells, n_iface_nodes;
  int     step;
};

static FILE * g_log;
static double g_res[256], g_upd[256];
static int    g_nres, g_nupd, g_ires, g_iupd;

static void init(void)
{
  if (g_log)
    return;
  const char *p = getenv("GF_FAKE_LOG");
  g_log         = fopen(p ? p : "/dev/null", "w");
  const char *s = getenv("GF_FAKE_SCRIPT");
  FILE *      f = s ? fopen(s, "r") : NULL;
  if (f)
    {
      char line[8192];
      for (int k = 0; k < 2 && fgets(line, sizeof line, f); ++k)
        for (char *t = strtok(line, " \n"); t; t = strtok(NULL, " \n"))
          {
            if (k == 0 && g_nres < 256)
              g_res[g_nres++] = atof(t);
            if (k == 1 && g_nupd < 256)
              g_upd[g_nupd++] = atof(t);
          }
      fclose(f);
    }
}
#define LOG(...)                 \
  do                             \
    {                            \
      init();                    \
      fprintf(g_log, __VA_ARGS__); \
      fflush(g_log);             \
    }                            \
  while (0)

int gf_create(const gf_desc *d, gf_handle *out)
{
  struct gf_context *c = calloc(1, sizeof *c);
  c->dim               = d->dim;
  c->degree            = d->degree;
  c->model             = d->model;
  c->n_cells           = d->n_cells;
  c->n_iface_nodes     = d->n_iface_nodes;
  *out                 = c;
  LOG("create dim=%d degree=%d model=%d n_dofs=%lld n_cells=%lld n_iface=%lld consistent=%d\n",
      d->dim, d->degree, d->model, (long long)d->n_dofs, (long long)d->n_cells,
      (long long)d->n_iface_nodes, d->data_consistent);
  return GF_OK;
}
void        gf_destroy(gf_handle h) { LOG("destroy\n"); free(h); }
const char *gf_last_error(gf_handle h) { (void)h; return "fake device error"; }
int gf_set_option(gf_handle h, int o, int64_t v) { (void)h; LOG("set_option %d %lld\n", o, (long long)v); return GF_OK; }
int gf_direct_info(gf_handle h, int64_t *n, int64_t *w, double *r)
{
  (void)h;
  if (n) *n = 0;
  if (w) *w = 7;
  if (r) *r = 0.0;
  return GF_OK;
}
int gf_mg_attach(gf_handle f, gf_handle c, const int32_t *t) { (void)f; (void)c; (void)t; LOG("mg_attach\n"); return GF_OK; }
int gf_set_traction(gf_handle h, const double *b) { LOG("set_traction %.17g\n", h->n_iface_nodes ? b[h->dim > 1 ? 1 : 0] : 0.0); return GF_OK; }
int gf_get_interface_displacement(gf_handle h, double *b)
{
  for (int64_t i = 0; i < h->n_iface_nodes * h->dim; ++i)
    b[i] = 1e-3 * h->step;
  LOG("get_interface_displacement\n");
  return GF_OK;
}
int gf_state_save(gf_handle h) { (void)h; LOG("state_save\n"); return GF_OK; }
int gf_state_restore(gf_handle h) { (void)h; LOG("state_restore\n"); return GF_OK; }
int gf_nl_begin_step(gf_handle h) { (void)h; LOG("nl_begin_step\n"); return GF_OK; }
int gf_nl_newton_assemble(gf_handle h, double *r)
{
  (void)h;
  init();
  *r = g_nres ? g_res[g_ires < g_nres ? g_ires : g_nres - 1] : 0.0;
  ++g_ires;
  LOG("nl_newton_assemble\n");
  return GF_OK;
}
int gf_nl_newton_solve(gf_handle h, int type, double tol, double maxit, uint32_t *it, double *res, double *upd)
{
  (void)h;
  init();
  *it  = 7;
  *res = 1e-7;
  *upd = g_nupd ? g_upd[g_iupd < g_nupd ? g_iupd : g_nupd - 1] : 0.0;
  ++g_iupd;
  LOG("nl_newton_solve type=%d tol=%g maxit=%g\n", type, tol, maxit);
  return GF_OK;
}
int gf_nl_end_step(gf_handle h) { h->step++; LOG("nl_end_step\n"); return GF_OK; }
int gf_lin_assemble_once(gf_handle h) { (void)h; LOG("lin_assemble_once\n"); return GF_OK; }
int gf_lin_step(gf_handle h, int type, double maxit, uint32_t *it, double *res)
{
  h->step++;
  *it  = 3;
  *res = 1e-11;
  LOG("lin_step type=%d maxit=%g\n", type, maxit);
  return GF_OK;
}
int gf_postprocess(gf_handle h, int which, double *fields)
{
  int npts = 1;
  for (int d = 0; d < h->dim; ++d)
    npts *= h->degree + 1;
  memset(fields, 0, sizeof(double) * (size_t)h->n_cells * npts * (h->dim + h->dim * h->dim));
  LOG("postprocess %d\n", which);
  return GF_OK;
}
