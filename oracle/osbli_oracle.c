/* osbli_oracle.c -- see osbli_oracle.h.  TEST INFRASTRUCTURE ONLY (CPU oracle, "port" of the
 * reference algorithm; every function cites the reference file:line it restates).
 * Build: gcc -O2 -std=c99 -ffp-contract=off -fPIC -shared osbli_oracle.c -o libosbli_oracle.so -lm
 */
#include "osbli_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NVMAX 5

typedef struct {
  int ndim, nv, h;
  int np[3], pd[3];
  long s[3]; /* strides */
  long n;    /* padded size */
} grid_t;

static void grid_init(const osbo_cfg *c, grid_t *g) {
  g->ndim = c->ndim; g->nv = c->ndim + 2; g->h = c->halo;
  long n = 1;
  for (int d = 0; d < 3; d++) {
    g->np[d] = d < c->ndim ? c->np[d] : 1;
    g->pd[d] = d < c->ndim ? c->np[d] + 2 * c->halo : 1;
    g->s[d] = n;
    n *= g->pd[d];
  }
  g->n = n;
}
static inline long gidx(const grid_t *g, int i, int j, int k) {
  long r = (long)(i + g->h);
  if (g->ndim > 1) r += g->s[1] * (j + g->h);
  if (g->ndim > 2) r += g->s[2] * (k + g->h);
  return r;
}
long osbo_padded_size(const osbo_cfg *c) { grid_t g; grid_init(c, &g); return g.n; }

/* scheme halos: WENO/TENO [-3,4] (weno.py:17-32, teno.py:18-36), central +-2 (scheme.py:32-44) */
static void scheme_halos(const osbo_cfg *c, int *hm, int *hp) {
  if (c->halo_m > 0) { *hm = c->halo_m; *hp = c->halo_p; return; }   /* block with further consumers of the halos (block.shock_filter: 3/4 on a central scheme) */
  if (c->conv == OSBO_CONV_CENTRAL) { *hm = 2; *hp = 2; } else { *hm = 3; *hp = 4; }
}

/* ---------------------------------------------------------------------------------------------
 * Boundary conditions
 * ------------------------------------------------------------------------------------------- */
/* periodic.py:42-56 + exchange.py:9-57 + opsc.py:555-593: buffered slab copy.
 * side 0: from [0,hm) to [np,np+hm);  side 1: from [np-hm, np-hm+hp) to [-hm, -hm+hp);
 * tangential extent = full scheme-halo padded range. */
static void bc_periodic(const osbo_cfg *c, const grid_t *g, double *const *q, int dir, int side) {
  int hm, hp; scheme_halos(c, &hm, &hp);
  int lo[3] = {0, 0, 0}, sz[3] = {1, 1, 1}, from[3], to[3];
  for (int d = 0; d < g->ndim; d++) { lo[d] = -hm; sz[d] = g->np[d] + hm + hp; }
  for (int d = 0; d < 3; d++) { from[d] = lo[d]; to[d] = lo[d]; }
  if (side == 0) { from[dir] = 0; to[dir] = g->np[dir]; sz[dir] = hm; }
  else { from[dir] = g->np[dir] - hm; to[dir] = -hm; sz[dir] = hp; }
  long cnt = (long)sz[0] * sz[1] * sz[2];
  double *tmp = (double *)malloc(sizeof(double) * cnt);
  for (int m = 0; m < g->nv; m++) {
    long n = 0;
    for (int k = 0; k < sz[2]; k++) for (int j = 0; j < sz[1]; j++) for (int i = 0; i < sz[0]; i++)
      tmp[n++] = q[m][gidx(g, from[0] + i, from[1] + j, from[2] + k)];
    n = 0;
    for (int k = 0; k < sz[2]; k++) for (int j = 0; j < sz[1]; j++) for (int i = 0; i < sz[0]; i++)
      q[m][gidx(g, to[0] + i, to[1] + j, to[2] + k)] = tmp[n++];
  }
  free(tmp);
}
/* part of a split face being applied (bc_core.py:110-127: arbitrary_bc_plane_kernel takes its range from run-time arrays
 * instead of the whole plane); NULL: whole plane.  The oracle is single-threaded test infrastructure. */
static const int *g_part_lo = NULL, *g_part_hi = NULL;
static const double *g_part_state = NULL;
/* dirichlet.py:28-41 + bc_core.py:158-198: boundary plane and the halo planes of that side get the
 * imposed state; tangential range = block range + scheme halos. */
static void bc_dirichlet(const osbo_cfg *c, const grid_t *g, double *const *q, int dir, int side) {
  int hm, hp; scheme_halos(c, &hm, &hp);
  int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  for (int d = 0; d < g->ndim; d++) { lo[d] = -hm; hi[d] = g->np[d] + hp; }
  if (g_part_lo) for (int d = 0; d < g->ndim; d++) { lo[d] = g_part_lo[d]; hi[d] = g_part_hi[d]; }
  if (side == 0) { lo[dir] = -hm; hi[dir] = 1; } else { lo[dir] = g->np[dir] - 1; hi[dir] = g->np[dir] + hp; }
  const double *state = g_part_state ? g_part_state : c->bc_q[dir][side];
  for (int m = 0; m < g->nv; m++)
    for (int k = lo[2]; k < hi[2]; k++) for (int j = lo[1]; j < hi[1]; j++) for (int i = lo[0]; i < hi[0]; i++)
      q[m][gidx(g, i, j, k)] = state[m];
}
/* plane loop helper: boundary plane of (dir, side), tangential range = block + scheme halos (bc_core.py:158-198) */
#define PLANE_LOOP(c, g, dir, side, ...)                                                               \
  {                                                                                                    \
    int hm_, hp_; scheme_halos(c, &hm_, &hp_);                                                         \
    int lo_[3] = {0, 0, 0}, hi_[3] = {1, 1, 1};                                                        \
    for (int d_ = 0; d_ < (g)->ndim; d_++) { lo_[d_] = -hm_; hi_[d_] = (g)->np[d_] + hp_; }            \
    if (g_part_lo) for (int d_ = 0; d_ < (g)->ndim; d_++) { lo_[d_] = g_part_lo[d_]; hi_[d_] = g_part_hi[d_]; } \
    lo_[dir] = (side) == 0 ? 0 : (g)->np[dir] - 1; hi_[dir] = lo_[dir] + 1;                            \
    for (int k = lo_[2]; k < hi_[2]; k++) for (int j = lo_[1]; j < hi_[1]; j++) for (int i = lo_[0]; i < hi_[0]; i++) { \
      const long x = gidx(g, i, j, k); (void)x; __VA_ARGS__                                                   \
    }                                                                                                  \
  }

/* dirichlet.py:28-41 with equations that depend on the tangential position (e.g. the shock generator of
 * apps/katzer_SBLI/katzer_SBLI.py:101-106): values come from a per-face table [nv][padded tangential extent]. */
static void bc_dirichlet_field(const osbo_cfg *c, const grid_t *g, double *const *q, int dir, int side) {
  int hm, hp; scheme_halos(c, &hm, &hp);
  const int n = side == 0 ? hm : hp;
  const long sd = g->s[dir];
  /* tangential linear index = padded index with dimension dir removed */
  long ts[3] = {0, 0, 0}, acc = 1;
  for (int d = 0; d < g->ndim; d++) if (d != dir) { ts[d] = acc; acc *= g->pd[d]; }
  const double *tab = c->bc_face[dir][side];
  PLANE_LOOP(c, g, dir, side, {
    int id[3] = {i, j, k};
    long t = 0;
    for (int d = 0; d < g->ndim; d++) if (d != dir) t += ts[d] * (id[d] + g->h);
    const int fr = c->bc_free[dir][side];
    for (int h = 0; h <= n; h++) {
      const long xo = x + (side == 0 ? -h : h) * sd;
      for (int m = 0; m < g->nv; m++) if (!(fr >> m & 1)) q[m][xo] = tab[m * acc + t];
      if (fr >> 8 & 1) {
        double ke = 0.0;
        for (int m = 1; m <= g->ndim; m++) if (fr >> m & 1) ke += q[m][xo] * q[m][xo];
        q[g->nv - 1][xo] = tab[(g->nv - 1) * acc + t] + 0.5 * ke / q[0][xo];
      }
    }
  })
}
/* extrapolation.py:29-58 */
static void bc_extrapolation(const osbo_cfg *c, const grid_t *g, double *const *q, int dir, int side) {
  int hm, hp; scheme_halos(c, &hm, &hp);
  const int n = side == 0 ? hm : hp;
  const long out = (side == 0 ? -1 : 1) * g->s[dir], in = -out;
  PLANE_LOOP(c, g, dir, side, {
    for (int m = 0; m < g->nv; m++) {
      if (c->extrap_order[dir][side] == 0) {
        for (int h = 0; h <= n; h++) q[m][x + h * out] = q[m][x + in];
      } else {
        for (int h = 1; h <= n; h++) q[m][x + h * out] = 2.0 * q[m][x + (h - 1) * out] - q[m][x + (h - 2) * out];
      }
    }
  })
}
/* inlet_pressure_extrapolate.py:32-66 (side 0 only) */
static void bc_inlet_pressure(const osbo_cfg *c, const grid_t *g, double *const *q, int dir, int side) {
  int hm, hp; scheme_halos(c, &hm, &hp);
  const int nd = g->ndim;
  const long sd = g->s[dir];
  (void)side; (void)hp;
  PLANE_LOOP(c, g, dir, 0, {
    double rhob = q[0][x], ub[3] = {0, 0, 0}, ke = 0.0;
    for (int d = 0; d < nd; d++) { ub[d] = fabs(q[1 + d][x] / q[0][x]); ke += ub[d] * ub[d]; }
    double pb = (c->gama - 1.0) * (-0.5 * rhob * ke + q[nd + 1][x]);
    double ab = sqrt(c->gama * pb / rhob);
    int sup = ub[dir] >= ab;
    for (int m = 0; m < g->nv; m++) q[m][x] = sup ? q[m][x - sd] : q[m][x];
    for (int h = 1; h <= hm; h++) q[nd + 1][x - h * sd] = sup ? q[nd + 1][x - h * sd] : q[nd + 1][x];
  })
}
/* isothermal_wall.py:32-88: no-slip wall at fixed temperature; halo states from wall pressure and a linearly
 * extrapolated temperature, velocities reflected. */
static void bc_isothermal_wall(const osbo_cfg *c, const grid_t *g, double *const *q, int dir, int side) {
  int hm, hp; scheme_halos(c, &hm, &hp);
  const int nd = g->ndim, n = side == 0 ? hm : hp;
  const long out = (side == 0 ? -1 : 1) * g->s[dir], in = -out;
  const double gm = c->gama, M2 = c->Minf * c->Minf;
  PLANE_LOOP(c, g, dir, side, {
    for (int d = 0; d < nd; d++) q[1 + d][x] = 0.0;
    q[nd + 1][x] = q[0][x] * c->Twall / (gm * (gm - 1.0) * M2);
    double ke = 0.0;
    for (int d = 0; d < nd; d++) ke += 0.5 * q[1 + d][x] * q[1 + d][x];
    const double Pw = (gm - 1.0) * (-ke / q[0][x] + q[nd + 1][x]);
    const long xa = x + in;
    double kea = 0.0;
    for (int d = 0; d < nd; d++) kea += 0.5 * q[1 + d][xa] * q[1 + d][xa];
    const double Ta = M2 * gm * (gm - 1.0) * (-kea / q[0][xa] + q[nd + 1][xa]) / q[0][xa];
    for (int h = 1; h <= n; h++) {
      const long xi = x + h * in, xo = x + h * out;
      const double Th = (h + 1) * c->Twall - h * Ta;
      const double rh = M2 * gm * Pw / Th;
      double u2 = 0.0;
      double uu[3];
      for (int d = 0; d < nd; d++) { uu[d] = q[1 + d][xi] / q[0][xi]; u2 += uu[d] * uu[d]; }
      q[0][xo] = rh;
      for (int d = 0; d < nd; d++) q[1 + d][xo] = -rh * uu[d];
      q[nd + 1][xo] = Pw / (gm - 1.0) + 0.5 * rh * u2;
    }
  })
}
/* adiabatic_wall.py:28-79: no-slip wall with dT/dn = 0 to fourth order, T_wall = 6/11 (3 T_1 - 3/2 T_2 + 1/3 T_3) from the
 * three points above the wall; wall energy = rho_wall T_wall / (gama (gama-1) Minf^2); halos mirror rho and rhoE and
 * reverse every momentum component.  Statement order as in the generated kernel (density halos, wall momentum, momentum
 * halos, wall energy, energy halos). */
static void bc_adiabatic_wall(const osbo_cfg *c, const grid_t *g, double *const *q, int dir, int side) {
  int hm, hp; scheme_halos(c, &hm, &hp);
  const int nd = g->ndim, n = side == 0 ? hm : hp;
  const long out = (side == 0 ? -1 : 1) * g->s[dir], in = -out;
  const double gm = c->gama, M2 = c->Minf * c->Minf;
  PLANE_LOOP(c, g, dir, side, {
    for (int h = 1; h <= n; h++) q[0][x + h * out] = q[0][x + h * in];
    for (int d = 0; d < nd; d++) q[1 + d][x] = 0.0;
    for (int h = 1; h <= n; h++) for (int d = 0; d < nd; d++) q[1 + d][x + h * out] = -q[1 + d][x + h * in];
    double T[4];
    for (int h = 1; h <= 3; h++) {
      const long xi = x + h * in;
      double ke = 0.0;
      for (int d = 0; d < nd; d++) ke += 0.5 * q[1 + d][xi] * q[1 + d][xi];
      T[h] = gm * M2 * (gm - 1.0) * (q[nd + 1][xi] - ke / q[0][xi]) / q[0][xi];
    }
    const double Tw = (6.0 / 11.0) * (3.0 * T[1] + (1.0 / 3.0) * T[3] - 1.5 * T[2]);
    q[nd + 1][x] = q[0][x] * Tw / (gm * (gm - 1.0) * M2);
    for (int h = 1; h <= n; h++) q[nd + 1][x + h * out] = q[nd + 1][x + h * in];
  })
}
/* symmetry.py:23-50 (cartesian: unit normal e_dir): halos mirror the interior with the normal momentum reversed */
static void bc_symmetry(const osbo_cfg *c, const grid_t *g, double *const *q, int dir, int side) {
  int hm, hp; scheme_halos(c, &hm, &hp);
  const int n = side == 0 ? hm : hp;
  const long out = (side == 0 ? -1 : 1) * g->s[dir], in = -out;
  PLANE_LOOP(c, g, dir, side, {
    for (int h = 1; h <= n; h++)
      for (int m = 0; m < g->nv; m++) {
        const double v = q[m][x + h * in];
        q[m][x + h * out] = (m == 1 + dir) ? v - 2.0 * v : v;
      }
  })
}
/* zero_gradient_outlet.py:12-23: the boundary point takes the value one point inside, the halos mirror the interior */
static void bc_zero_gradient_outlet(const osbo_cfg *c, const grid_t *g, double *const *q, int dir, int side) {
  int hm, hp; scheme_halos(c, &hm, &hp);
  const int n = side == 0 ? hm : hp;
  const long out = (side == 0 ? -1 : 1) * g->s[dir], in = -out;
  PLANE_LOOP(c, g, dir, side, {
    for (int m = 0; m < g->nv; m++) q[m][x] = q[m][x + in];
    for (int h = 1; h <= n; h++)
      for (int m = 0; m < g->nv; m++) q[m][x + h * out] = q[m][x + h * in];
  })
}
/* pressure_outlet.py:33-52 (side 1 only): density and momentum one point inside are copied to the boundary point and the
 * halos; the energy there is back_pressure/(gama-1) + 1/2 m.m/rho of that inner point */
static void bc_pressure_outlet(const osbo_cfg *c, const grid_t *g, double *const *q, int dir, int side) {
  int hm, hp; scheme_halos(c, &hm, &hp);
  const int nd = g->ndim, n = side == 0 ? hm : hp;
  const long out = (side == 0 ? -1 : 1) * g->s[dir], in = -out;
  PLANE_LOOP(c, g, dir, side, {
    for (int h = 0; h <= n; h++) {
      const long xo = x + h * out, xi = x + in;
      double mm = 0.0;
      for (int d = 0; d < nd; d++) mm += q[1 + d][xi] * q[1 + d][xi];
      const double rho = q[0][xi];
      const double E = c->back_pressure / (c->gama - 1.0) + 0.5 * mm / rho;
      q[0][xo] = rho;
      for (int d = 0; d < nd; d++) q[1 + d][xo] = q[1 + d][xi];
      q[nd + 1][xo] = E;
    }
  })
}
/* inviscid_wall.py:24-52 on a Cartesian block (unit normal e_dir): halos mirror the interior with the normal momentum reversed,
 * the boundary point takes the state one point inside with the normal momentum removed */
static void bc_inviscid_wall(const osbo_cfg *c, const grid_t *g, double *const *q, int dir, int side) {
  int hm, hp; scheme_halos(c, &hm, &hp);
  const int n = side == 0 ? hm : hp;
  const long out = (side == 0 ? -1 : 1) * g->s[dir], in = -out;
  PLANE_LOOP(c, g, dir, side, {
    for (int h = 1; h <= n; h++)
      for (int m = 0; m < g->nv; m++) {
        const double v = q[m][x + h * in];
        q[m][x + h * out] = (m == 1 + dir) ? v - 2.0 * v : v;
      }
    for (int m = 0; m < g->nv; m++) {
      const double v = q[m][x + in];
      q[m][x] = (m == 1 + dir) ? v - 1.0 * v : v;
    }
  })
}
/* order: dir0 side0, dir0 side1, dir1 side0 ...  (block.py:199-210, algorithm.py:440-442) */
static void bc_apply_kind(const osbo_cfg *c, const grid_t *g, double *const *q, int kind, int d, int s) {
  switch (kind) {
    case OSBO_BC_PERIODIC: bc_periodic(c, g, q, d, s); break;
    case OSBO_BC_DIRICHLET: bc_dirichlet(c, g, q, d, s); break;
    case OSBO_BC_DIRICHLET_FIELD: bc_dirichlet_field(c, g, q, d, s); break;
    case OSBO_BC_EXTRAPOLATION: bc_extrapolation(c, g, q, d, s); break;
    case OSBO_BC_INLET_PRESSURE_EXTRAPOLATE: bc_inlet_pressure(c, g, q, d, s); break;
    case OSBO_BC_ISOTHERMAL_WALL: bc_isothermal_wall(c, g, q, d, s); break;
    case OSBO_BC_SYMMETRY: bc_symmetry(c, g, q, d, s); break;
    case OSBO_BC_ADIABATIC_WALL: bc_adiabatic_wall(c, g, q, d, s); break;
    case OSBO_BC_ZERO_GRADIENT_OUTLET: bc_zero_gradient_outlet(c, g, q, d, s); break;
    case OSBO_BC_PRESSURE_OUTLET: bc_pressure_outlet(c, g, q, d, s); break;
    case OSBO_BC_INVISCID_WALL: bc_inviscid_wall(c, g, q, d, s); break;
    default: break;
  }
}
void osbo_apply_bcs(const osbo_cfg *c, double *const *q) {
  grid_t g; grid_init(c, &g);
  for (int d = 0; d < c->ndim; d++)
    for (int s = 0; s < 2; s++) {
      if (c->bc[d][s] == OSBO_BC_SPLIT) {
        /* bc_core.py:200-217: every part is the boundary class's own kernel over the part's range, in the order given */
        for (int n = 0; n < c->split_n[d][s]; n++) {
          osbo_cfg part = *c;                       /* the part's parameters where the class reads them from the face */
          part.extrap_order[d][s] = c->split_order[d][s][n];
          g_part_lo = c->split_lo[d][s][n]; g_part_hi = c->split_hi[d][s][n]; g_part_state = c->split_q[d][s][n];
          bc_apply_kind(&part, &g, q, c->split_kind[d][s][n], d, s);
          g_part_lo = g_part_hi = NULL; g_part_state = NULL;
        }
      } else {
        bc_apply_kind(c, &g, q, c->bc[d][s], d, s);
      }
    }
}

/* ---------------------------------------------------------------------------------------------
 * Constituent relations (app strings, e.g. apps/Sod_shock_tube/Sod_shock_tube.py:26-28):
 *   u_i = rhou_i/rho ; p = (gama-1)(rhoE - 1/2 rho u_i u_i) ; a = sqrt(gama p/rho) ; T = gama Minf^2 p/rho
 * evaluated over grid + scheme halos (opensbliequations.py:297-321).
 * ------------------------------------------------------------------------------------------- */
typedef struct { double *u[3], *p, *a, *T, *mu; } prim_t;

static void constituent(const osbo_cfg *c, const grid_t *g, double *const *q, prim_t *P) {
  int hm, hp; scheme_halos(c, &hm, &hp);
  int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  for (int d = 0; d < g->ndim; d++) { lo[d] = -hm; hi[d] = g->np[d] + hp; }
  const int nd = g->ndim;
  for (int k = lo[2]; k < hi[2]; k++) for (int j = lo[1]; j < hi[1]; j++) for (int i = lo[0]; i < hi[0]; i++) {
    long x = gidx(g, i, j, k);
    double rho = q[0][x], ke = 0.0;
    for (int d = 0; d < nd; d++) { double u = q[1 + d][x] / rho; P->u[d][x] = u; ke += 0.5 * rho * u * u; }
    double p = (c->gama - 1.0) * (q[nd + 1][x] - ke);
    P->p[x] = p;
    P->a[x] = sqrt(c->gama * p / rho);
    P->T[x] = c->Minf * c->Minf * c->gama * p / rho;
    /* viscosity laws of the apps: Sutherland katzer_SBLI.py:25, power law compressible_TCF_TENO/turbulent_channel.py:29 */
    if (c->visc_law == OSBO_MU_SUTHERLAND) P->mu[x] = (c->SuthT / c->RefT + 1.0) * pow(P->T[x], 1.5) / (c->SuthT / c->RefT + P->T[x]);
    else if (c->visc_law == OSBO_MU_POWER) P->mu[x] = pow(P->T[x], c->mu_exp);
    else P->mu[x] = 1.0;
  }
}

/* ---------------------------------------------------------------------------------------------
 * Non-linear reconstructions.  All work on g[0..4] = f(-2..2) for the right-biased (f+) side;
 * the left-biased (f-) side is the mirror image about i+1/2 (p -> 1-p), cf. teno.py:102-122,159-177.
 * ------------------------------------------------------------------------------------------- */
static inline double sq(double x) { return x * x; }
static inline double pow6(double x) { double x2 = x * x; return x2 * x2 * x2; }

/* TENO5: teno.py:113-133 (ENO coeffs, d = 11/20, 4/10, 1/20), 159-177 (beta), 195-213 (alpha, tau5,
 * C=1, q=6), 445-465 (cut-off delta_r = alpha_r*inv_alpha_sum < CT ? 0 : 1), 407-428 (omega). */
static double teno5_side(const double *f, double eps, double ct) {
  const double fm2 = f[0], fm1 = f[1], f0 = f[2], f1 = f[3], f2 = f[4];
  double b0 = 0.25 * sq(fm1 - f1) + (13.0 / 12.0) * sq(fm1 - 2.0 * f0 + f1);
  double b1 = 0.25 * sq(3.0 * f0 - 4.0 * f1 + f2) + (13.0 / 12.0) * sq(f0 - 2.0 * f1 + f2);
  double b2 = 0.25 * sq(fm2 - 4.0 * fm1 + 3.0 * f0) + (13.0 / 12.0) * sq(fm2 - 2.0 * fm1 + f0);
  double tau = fabs(b0 - b2);
  double a0 = pow6(1.0 + tau / (eps + b0)), a1 = pow6(1.0 + tau / (eps + b1)), a2 = pow6(1.0 + tau / (eps + b2));
  double ias = 1.0 / (a0 + a1 + a2);
  double d0 = (ct > a0 * ias) ? 0.0 : 1.0, d1 = (ct > a1 * ias) ? 0.0 : 1.0, d2 = (ct > a2 * ias) ? 0.0 : 1.0;
  double w = 1.0 / ((11.0 / 20.0) * d0 + (2.0 / 5.0) * d1 + (1.0 / 20.0) * d2);
  return (11.0 / 20.0) * d0 * w * (-(1.0 / 6.0) * fm1 + (5.0 / 6.0) * f0 + (1.0 / 3.0) * f1)
       + (2.0 / 5.0) * d1 * w * ((1.0 / 3.0) * f0 + (5.0 / 6.0) * f1 - (1.0 / 6.0) * f2)
       + (1.0 / 20.0) * d2 * w * ((1.0 / 3.0) * fm2 - (7.0 / 6.0) * fm1 + (11.0 / 6.0) * f0);
}
double osbo_recon_teno5(const double *fp, const double *fm, double eps, double ct) {
  double gm[5] = {fm[5], fm[4], fm[3], fm[2], fm[1]};
  return teno5_side(fp, eps, ct) + teno5_side(gm, eps, ct);
}

/* TENO6: teno.py:85,102-104 (4th stencil), 120-122 (ENO 3/12,13/12,-5/12,1/12), 136 (d = 231/500,
 * 3/10, 27/500, 23/125), 165-177 (beta_3), 216-234 (tau6 = |b3 - 1/6(b0 + b2 - 4 b1)|).
 * Reference quirk reproduced: on the right-biased side the last term of beta_3 is NOT squared
 * (teno.py:166-167); on the left-biased side it is (teno.py:175-177). */
static double teno6_side(const double *f, double eps, double ct, int square_last) {
  const double fm2 = f[0], fm1 = f[1], f0 = f[2], f1 = f[3], f2 = f[4], f3 = f[5];
  double b0 = 0.25 * sq(fm1 - f1) + (13.0 / 12.0) * sq(fm1 - 2.0 * f0 + f1);
  double b1 = 0.25 * sq(3.0 * f0 - 4.0 * f1 + f2) + (13.0 / 12.0) * sq(f0 - 2.0 * f1 + f2);
  double b2 = 0.25 * sq(fm2 - 4.0 * fm1 + 3.0 * f0) + (13.0 / 12.0) * sq(fm2 - 2.0 * fm1 + f0);
  double l3 = -f0 + 3.0 * f1 - 3.0 * f2 + f3;
  double b3 = (1.0 / 36.0) * sq(-11.0 * f0 + 18.0 * f1 - 9.0 * f2 + 2.0 * f3)
            + (13.0 / 12.0) * sq(2.0 * f0 - 5.0 * f1 + 4.0 * f2 - f3)
            + (781.0 / 720.0) * (square_last ? l3 * l3 : l3);
  double tau = fabs(b3 - (1.0 / 6.0) * (b0 + b2 - 4.0 * b1));
  double a0 = pow6(1.0 + tau / (eps + b0)), a1 = pow6(1.0 + tau / (eps + b1));
  double a2 = pow6(1.0 + tau / (eps + b2)), a3 = pow6(1.0 + tau / (eps + b3));
  double ias = 1.0 / (a0 + a1 + a2 + a3);
  double d0 = (ct > a0 * ias) ? 0.0 : 1.0, d1 = (ct > a1 * ias) ? 0.0 : 1.0;
  double d2 = (ct > a2 * ias) ? 0.0 : 1.0, d3 = (ct > a3 * ias) ? 0.0 : 1.0;
  const double c0 = 231.0 / 500.0, c1 = 3.0 / 10.0, c2 = 27.0 / 500.0, c3 = 23.0 / 125.0;
  double w = 1.0 / (c0 * d0 + c1 * d1 + c2 * d2 + c3 * d3);
  return c0 * d0 * w * (-(1.0 / 6.0) * fm1 + (5.0 / 6.0) * f0 + (1.0 / 3.0) * f1)
       + c1 * d1 * w * ((1.0 / 3.0) * f0 + (5.0 / 6.0) * f1 - (1.0 / 6.0) * f2)
       + c2 * d2 * w * ((1.0 / 3.0) * fm2 - (7.0 / 6.0) * fm1 + (11.0 / 6.0) * f0)
       + c3 * d3 * w * ((3.0 / 12.0) * f0 + (13.0 / 12.0) * f1 - (5.0 / 12.0) * f2 + (1.0 / 12.0) * f3);
}
double osbo_recon_teno6(const double *fp, const double *fm, double eps, double ct) {
  double gm[6] = {fm[5], fm[4], fm[3], fm[2], fm[1], fm[0]};
  /* mirrored last term: -(..) ; squared on this side so the sign is immaterial */
  return teno6_side(fp, eps, ct, 0) + teno6_side(gm, eps, ct, 1);
}

/* WENO5 (k=3): weno.py:72-98 (ENO c_rj, Shu 1997 table 2.1), 100-120 (d = 3/10, 6/10, 1/10),
 * 138-205 (Jiang-Shu beta), 341-369 (JS: alpha = d/(eps+beta)^2, eps = 1e-6),
 * 283-338 (Z: tau = |b0-b2|, alpha = d(1 + (tau/(eps+beta))^2), eps = 1e-14). Stencil r = {-r..-r+2}. */
static double weno5_side(const double *f, int z) {
  const double fm2 = f[0], fm1 = f[1], f0 = f[2], f1 = f[3], f2 = f[4];
  double b0 = (13.0 / 12.0) * sq(f0 - 2.0 * f1 + f2) + 0.25 * sq(3.0 * f0 - 4.0 * f1 + f2);
  double b1 = (13.0 / 12.0) * sq(fm1 - 2.0 * f0 + f1) + 0.25 * sq(fm1 - f1);
  double b2 = (13.0 / 12.0) * sq(fm2 - 2.0 * fm1 + f0) + 0.25 * sq(fm2 - 4.0 * fm1 + 3.0 * f0);
  double a0, a1, a2;
  if (!z) {
    const double e = 1.0e-6;
    a0 = (3.0 / 10.0) / sq(b0 + e); a1 = (3.0 / 5.0) / sq(b1 + e); a2 = (1.0 / 10.0) / sq(b2 + e);
  } else {
    const double e = 1.0e-14;
    double tau = fabs(b0 - b2);
    a0 = 0.3 + (3.0 / 10.0) * sq(tau) / sq(b0 + e);
    a1 = 0.6 + (3.0 / 5.0) * sq(tau) / sq(b1 + e);
    a2 = 0.1 + (1.0 / 10.0) * sq(tau) / sq(b2 + e);
  }
  double ias = 1.0 / (a0 + a1 + a2);
  return a0 * ias * ((1.0 / 3.0) * f0 + (5.0 / 6.0) * f1 - (1.0 / 6.0) * f2)
       + a1 * ias * (-(1.0 / 6.0) * fm1 + (5.0 / 6.0) * f0 + (1.0 / 3.0) * f1)
       + a2 * ias * ((1.0 / 3.0) * fm2 - (7.0 / 6.0) * fm1 + (11.0 / 6.0) * f0);
}
double osbo_recon_weno5(const double *fp, const double *fm, int z) {
  double gm[5] = {fm[5], fm[4], fm[3], fm[2], fm[1]};
  return weno5_side(fp, z) + weno5_side(gm, z);
}

/* ---------------------------------------------------------------------------------------------
 * Eigensystems for direction d with identity direction cosines (k~ = e_d; the branch taken when
 * the FD metric matrix is diagonal, euler_eigensystem.py:50-54).
 * 1-D: :57-75   2-D: :77-105   3-D: :107-135.   L = LEV, Rm = REV (nv x nv, row major, stride 5).
 * ------------------------------------------------------------------------------------------- */
static void eigensystem(int nd, int d, const double *kk, double gama, double rho, const double *u, double a,
                        double L[5][5], double Rm[5][5]) {
  const double gm1 = gama - 1.0;
  memset(L, 0, sizeof(double) * 25); memset(Rm, 0, sizeof(double) * 25);
  if (nd == 1) {
    double u0 = u[0], H = a * a / gm1 + 0.5 * u0 * u0, D = 2.0 * H - u0 * u0;
    L[0][0] = u0 * (2.0 * H + a * u0 - u0 * u0) / (2.0 * a * D); L[0][1] = (-H - a * u0 + 0.5 * u0 * u0) / (a * D); L[0][2] = 1.0 / D;
    L[1][0] = 2.0 * (H - u0 * u0) / D; L[1][1] = 2.0 * u0 / D; L[1][2] = -2.0 / D;
    L[2][0] = u0 * (-2.0 * H + a * u0 + u0 * u0) / (2.0 * a * D); L[2][1] = (H - a * u0 - 0.5 * u0 * u0) / (a * D); L[2][2] = 1.0 / D;
    Rm[0][0] = 1; Rm[0][1] = 1; Rm[0][2] = 1;
    Rm[1][0] = u0 - a; Rm[1][1] = u0; Rm[1][2] = u0 + a;
    Rm[2][0] = H - u0 * a; Rm[2][1] = 0.5 * u0 * u0; Rm[2][2] = H + u0 * a;
    return;
  }
  const double s2 = sqrt(2.0);
  double al = rho / (a * s2), bt = 1.0 / (rho * a * s2);
  if (nd == 2) {
    double k0 = kk ? kk[0] : (d == 0), k1 = kk ? kk[1] : (d == 1), u0 = u[0], u1 = u[1];
    double th = k0 * u0 + k1 * u1, ph = gm1 * 0.5 * (u0 * u0 + u1 * u1), a2 = a * a;
    L[0][0] = 1.0 - ph / a2; L[0][1] = gm1 * u0 / a2; L[0][2] = gm1 * u1 / a2; L[0][3] = -gm1 / a2;
    L[1][0] = -(k1 * u0 - k0 * u1) / rho; L[1][1] = k1 / rho; L[1][2] = -k0 / rho; L[1][3] = 0.0;
    L[2][0] = bt * (ph - a * th); L[2][1] = bt * (k0 * a - gm1 * u0); L[2][2] = bt * (k1 * a - gm1 * u1); L[2][3] = bt * gm1;
    L[3][0] = bt * (ph + a * th); L[3][1] = -bt * (k0 * a + gm1 * u0); L[3][2] = -bt * (k1 * a + gm1 * u1); L[3][3] = bt * gm1;
    Rm[0][0] = 1; Rm[0][1] = 0; Rm[0][2] = al; Rm[0][3] = al;
    Rm[1][0] = u0; Rm[1][1] = k1 * rho; Rm[1][2] = al * (u0 + k0 * a); Rm[1][3] = al * (u0 - k0 * a);
    Rm[2][0] = u1; Rm[2][1] = -k0 * rho; Rm[2][2] = al * (u1 + k1 * a); Rm[2][3] = al * (u1 - k1 * a);
    Rm[3][0] = ph / gm1; Rm[3][1] = rho * (k1 * u0 - k0 * u1);
    Rm[3][2] = al * ((ph + a2) / gm1 + a * th); Rm[3][3] = al * ((ph + a2) / gm1 - a * th);
    return;
  }
  double k0 = kk ? kk[0] : (d == 0), k1 = kk ? kk[1] : (d == 1), k2 = kk ? kk[2] : (d == 2), u0 = u[0], u1 = u[1], u2 = u[2];
  double th = k0 * u0 + k1 * u1 + k2 * u2, ph = gm1 * 0.5 * (u0 * u0 + u1 * u1 + u2 * u2), a2 = a * a;
  L[0][0] = k0 * (1.0 - ph / a2) - (k2 * u1 - k1 * u2) / rho; L[0][1] = k0 * gm1 * u0 / a2;
  L[0][2] = k0 * gm1 * u1 / a2 + k2 / rho; L[0][3] = k0 * gm1 * u2 / a2 - k1 / rho; L[0][4] = -k0 * gm1 / a2;
  L[1][0] = k1 * (1.0 - ph / a2) - (k0 * u2 - k2 * u0) / rho; L[1][1] = k1 * gm1 * u0 / a2 - k2 / rho;
  L[1][2] = k1 * gm1 * u1 / a2; L[1][3] = k1 * gm1 * u2 / a2 + k0 / rho; L[1][4] = -k1 * gm1 / a2;
  L[2][0] = k2 * (1.0 - ph / a2) - (k1 * u0 - k0 * u1) / rho; L[2][1] = k2 * gm1 * u0 / a2 + k1 / rho;
  L[2][2] = k2 * gm1 * u1 / a2 - k0 / rho; L[2][3] = k2 * gm1 * u2 / a2; L[2][4] = -k2 * gm1 / a2;
  L[3][0] = bt * (ph - th * a); L[3][1] = -bt * (gm1 * u0 - k0 * a); L[3][2] = -bt * (gm1 * u1 - k1 * a);
  L[3][3] = -bt * (gm1 * u2 - k2 * a); L[3][4] = bt * gm1;
  L[4][0] = bt * (ph + th * a); L[4][1] = -bt * (gm1 * u0 + k0 * a); L[4][2] = -bt * (gm1 * u1 + k1 * a);
  L[4][3] = -bt * (gm1 * u2 + k2 * a); L[4][4] = bt * gm1;
  Rm[0][0] = k0; Rm[0][1] = k1; Rm[0][2] = k2; Rm[0][3] = al; Rm[0][4] = al;
  Rm[1][0] = k0 * u0; Rm[1][1] = k1 * u0 - k2 * rho; Rm[1][2] = k2 * u0 + k1 * rho; Rm[1][3] = al * (u0 + k0 * a); Rm[1][4] = al * (u0 - k0 * a);
  Rm[2][0] = k0 * u1 + k2 * rho; Rm[2][1] = k1 * u1; Rm[2][2] = k2 * u1 - k0 * rho; Rm[2][3] = al * (u1 + k1 * a); Rm[2][4] = al * (u1 - k1 * a);
  Rm[3][0] = k0 * u2 - k1 * rho; Rm[3][1] = k1 * u2 + k0 * rho; Rm[3][2] = k2 * u2; Rm[3][3] = al * (u2 + k2 * a); Rm[3][4] = al * (u2 - k2 * a);
  Rm[4][0] = k0 * ph / gm1 + rho * (k2 * u1 - k1 * u2); Rm[4][1] = k1 * ph / gm1 + rho * (k0 * u2 - k2 * u0);
  Rm[4][2] = k2 * ph / gm1 + rho * (k1 * u0 - k0 * u1);
  Rm[4][3] = al * ((ph + a2) / gm1 + th * a); Rm[4][4] = al * ((ph + a2) / gm1 - th * a);
}

/* eigenvalue j at a stencil point (euler_eigensystem.py ev; shock_capturing.py:512-536) */
static inline double eigenvalue(int nd, int j, double ud, double a) {
  if (nd == 1) return j == 0 ? ud - a : (j == 1 ? ud : ud + a);
  if (j < nd) return ud;
  return j == nd ? ud + a : ud - a;
}

/* ---------------------------------------------------------------------------------------------
 * LLF characteristic flux reconstruction in direction d (shock_capturing.py:357-398 pre_process,
 * 450-477 CF/CS, 479-495 split, 512-536 max wave speed, 418-448 post_process), evaluated for the
 * interfaces i+1/2, i = -1 .. np_d-1 (the reference also evaluates i = np_d, which nothing consumes).
 * Output wk[m] = flux component m at interface i+1/2 stored at point i.
 * ------------------------------------------------------------------------------------------- */
/* adaptive TENO cut-off from the sensor (teno.py:430-443): C_T = 10^-floor(a1 - a2 (1 - (1-theta)^4 (1+4 theta))) */
static double adaptive_ct(const osbo_cfg *c, double th) {
  double om = 1.0 - th;
  return pow(10.0, -floor(c->teno_a1 - c->teno_a2 * (-(om * om) * (om * om) * (4.0 * th + 1.0) + 1.0)));
}

/* one interface: stencil data for the 6 points p = 0..5 <-> offsets -2..3 */
/* met != NULL (curvilinear): met[p][0..nd-1] = D_dir,j and met[p][3] = detJ at the stencil points.  Then (euler_wave.py:12-18,
 * shock_capturing.py:357-536 with the metric-aware eigensystem) the direction cosines are the simple average of D_dir. over
 * the two interface points, normalised; the flux vector is detJ (U q + p (0, D_dir., U)) with U = D_dir,j u_j; the wave speeds
 * are U, U +- a |D_dir.| at each stencil point. */
static void interface_flux(const osbo_cfg *c, int nd, int dir, double qs[6][5], double us[6][3],
                           const double *ps, const double *as, double teno_ct, double *flux, double (*met)[4]) {
  const int nv = nd + 2;
  const double gm1 = c->gama - 1.0;
  /* interface state between points 2 and 3 (averaging.py:31-59 simple, :62-114 Roe) */
  double rho, u[3] = {0, 0, 0}, a;
  if (c->averaging == OSBO_AVG_ROE) {
    double sl = sqrt(qs[2][0]), sr = sqrt(qs[3][0]);
    rho = sqrt(qs[2][0] * qs[3][0]);
    double w = 1.0 / (sr + sl), ke = 0.0;
    for (int d = 0; d < nd; d++) { u[d] = w * (sr * us[3][d] + sl * us[2][d]); ke += u[d] * u[d]; }
    double Hh = w * ((ps[2] + qs[2][nd + 1]) / sl + (ps[3] + qs[3][nd + 1]) / sr);
    a = sqrt(gm1 * (Hh - 0.5 * ke));
  } else {
    rho = 0.5 * (qs[2][0] + qs[3][0]);
    for (int d = 0; d < nd; d++) u[d] = 0.5 * (us[2][d] + us[3][d]);
    a = 0.5 * (as[2] + as[3]);
  }
  double L[5][5], Rm[5][5], kk[3] = {0, 0, 0};
  if (met) {
    double n2 = 0.0;
    for (int d = 0; d < nd; d++) { kk[d] = 0.5 * (met[2][d] + met[3][d]); n2 += kk[d] * kk[d]; }
    const double inv = pow(n2, -0.5);
    for (int d = 0; d < nd; d++) kk[d] *= inv;
  }
  eigensystem(nd, dir, met ? kk : NULL, c->gama, rho, u, a, L, Rm);
  /* characteristic flux / solution over the 6 stencil points and max |lambda| */
  double CF[5][6], CS[5][6], lam[5] = {0, 0, 0, 0, 0};
  for (int p = 0; p < 6; p++) {
    double F[5], ud = us[p][dir], pr = ps[p], am = as[p];
    const double *qv = qs[p];
    if (met) {
      double n2 = 0.0;
      ud = 0.0;
      for (int d = 0; d < nd; d++) { ud += met[p][d] * us[p][d]; n2 += met[p][d] * met[p][d]; }
      am = sqrt(n2) * as[p];
      F[0] = met[p][3] * (qv[0] * ud);
      for (int d = 0; d < nd; d++) F[1 + d] = met[p][3] * (qv[1 + d] * ud + met[p][d] * pr);
      F[nd + 1] = met[p][3] * ((pr + qv[nd + 1]) * ud);
    } else {
      F[0] = qv[1 + dir];
      for (int d = 0; d < nd; d++) F[1 + d] = qv[1 + d] * ud + (d == dir ? pr : 0.0);
      F[nd + 1] = (pr + qv[nd + 1]) * ud;
    }
    for (int jj = 0; jj < nv; jj++) {
      double cf = 0.0, cs = 0.0;
      for (int m = 0; m < nv; m++) { cf += L[jj][m] * F[m]; cs += L[jj][m] * qv[m]; }
      CF[jj][p] = cf; CS[jj][p] = cs;
      double l = fabs(eigenvalue(nd, jj, ud, am));
      if (l > lam[jj]) lam[jj] = l;
    }
  }
  double rec[5];
  for (int jj = 0; jj < nv; jj++) {
    double fp[6], fm[6];
    for (int p = 0; p < 6; p++) { fp[p] = 0.5 * (CF[jj][p] + lam[jj] * CS[jj][p]); fm[p] = 0.5 * (CF[jj][p] - lam[jj] * CS[jj][p]); }
    if (c->conv == OSBO_CONV_TENO) rec[jj] = c->order == 6 ? osbo_recon_teno6(fp, fm, c->eps, teno_ct) : osbo_recon_teno5(fp, fm, c->eps, teno_ct);
    else rec[jj] = osbo_recon_weno5(fp, fm, c->weno_z);
  }
  for (int m = 0; m < nv; m++) {
    double f = 0.0;
    for (int jj = 0; jj < nv; jj++) f += Rm[m][jj] * rec[jj];
    flux[m] = f;
  }
}

/* test hook: flux of one interface from 6 conservative states q6[p][m] (constituent relations applied here) */
void osbo_interface_flux(const osbo_cfg *c, int dir, const double *q6, double *flux) {
  const int nd = c->ndim, nv = nd + 2;
  double qs[6][5], us[6][3], ps[6], as[6];
  for (int p = 0; p < 6; p++) {
    double ke = 0.0;
    for (int m = 0; m < nv; m++) qs[p][m] = q6[p * nv + m];
    for (int d = 0; d < nd; d++) { us[p][d] = qs[p][1 + d] / qs[p][0]; ke += 0.5 * qs[p][0] * us[p][d] * us[p][d]; }
    ps[p] = (c->gama - 1.0) * (qs[p][nd + 1] - ke);
    as[p] = sqrt(c->gama * ps[p] / qs[p][0]);
  }
  interface_flux(c, nd, dir, qs, us, ps, as, c->teno_ct, flux, NULL);
}

static void llf_flux(const osbo_cfg *c, const grid_t *g, int dir, double *const *q, const prim_t *P, double **wk) {
  const int nd = g->ndim, nv = g->nv;
  int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  for (int d = 0; d < nd; d++) { lo[d] = 0; hi[d] = g->np[d]; }
  lo[dir] = -1;
  const long sd = g->s[dir];
  for (int k = lo[2]; k < hi[2]; k++) for (int j = lo[1]; j < hi[1]; j++) for (int i = lo[0]; i < hi[0]; i++) {
    const long x = gidx(g, i, j, k);
    double qs[6][5], us[6][3], ps[6], as[6], fl[5];
    for (int p = 0; p < 6; p++) {
      const long xp = x + (p - 2) * sd;
      for (int m = 0; m < nv; m++) qs[p][m] = q[m][xp];
      for (int d = 0; d < nd; d++) us[p][d] = P->u[d][xp];
      ps[p] = P->p[xp]; as[p] = P->a[xp];
    }
    double ct = c->teno_ct;
    if (c->teno_adaptive) {          /* sensor value of the interface's left point; halo points hold 0 */
      ct = adaptive_ct(c, c->theta[x]);
      if (dir == 0 && c->teno_store) c->teno_store[x] = ct;
    }
    double met[6][4];
    if (c->curv_detJ)
      for (int p = 0; p < 6; p++) {
        const long xp = x + (p - 2) * sd;
        for (int d = 0; d < nd; d++) met[p][d] = c->curv_D[dir][d][xp];
        met[p][3] = c->curv_detJ[xp];
      }
    interface_flux(c, nd, dir, qs, us, ps, as, ct, fl, c->curv_detJ ? met : NULL);
    for (int m = 0; m < nv; m++) wk[m][x] = fl[m];
  }
}

/* central 4th-order first / second derivative at x along stride s (scheme.py:81-85 weights;
 * opensblifunctions.py:485-521) */
static inline double d1c(const double *f, long x, long s, double inv) {
  return (1.0 / 12.0) * inv * (f[x - 2 * s] - 8.0 * f[x - s] + 8.0 * f[x + s] - f[x + 2 * s]);
}
static inline double d2c(const double *f, long x, long s, double inv2) {
  return (1.0 / 12.0) * inv2 * (-f[x - 2 * s] + 16.0 * f[x - s] - 30.0 * f[x] + 16.0 * f[x + s] - f[x + 2 * s]);
}

/* Central derivative with the one-sided boundary closure selected by the grid index along the derivative
 * direction (opensblifunctions.py:523-534 modify_boundary_formula; reduced_access_scheme.py:36-83,
 * Carpenter_scheme.py:38-102).  Closure tables hold, for the rows idx = 0..nr-1 next to side 0, the weights of the
 * boundary-absolute points 0..np-1; side 1 mirrors them (sign -1 for first derivatives). */
static double d1g(const osbo_cfg *c, const double *f, long x, long s, double inv, int dir, int idx, int n) {
  if (c->closure[dir][0] && idx < c->c_nr1) {
    double r = 0.0; const long x0 = x - idx * s;
    for (int p = 0; p < c->c_np1; p++) r += c->c_d1[idx * c->c_np1 + p] * f[x0 + p * s];
    return inv * r;
  }
  if (c->closure[dir][1] && n - 1 - idx < c->c_nr1) {
    const int row = n - 1 - idx; double r = 0.0; const long x0 = x + row * s;
    for (int p = 0; p < c->c_np1; p++) r -= c->c_d1[row * c->c_np1 + p] * f[x0 - p * s];
    return inv * r;
  }
  return d1c(f, x, s, inv);
}
static double d2g(const osbo_cfg *c, const double *f, long x, long s, double inv2, int dir, int idx, int n) {
  if (c->closure[dir][0] && idx < c->c_nr2) {
    double r = 0.0; const long x0 = x - idx * s;
    for (int p = 0; p < c->c_np2; p++) r += c->c_d2[idx * c->c_np2 + p] * f[x0 + p * s];
    return inv2 * r;
  }
  if (c->closure[dir][1] && n - 1 - idx < c->c_nr2) {
    const int row = n - 1 - idx; double r = 0.0; const long x0 = x + row * s;
    for (int p = 0; p < c->c_np2; p++) r += c->c_d2[row * c->c_np2 + p] * f[x0 - p * s];
    return inv2 * r;
  }
  return d2c(f, x, s, inv2);
}

/* ---------------------------------------------------------------------------------------------
 * General viscous terms: variable viscosity, stretched (diagonal-metric) grid, one-sided closures.
 * With  d_j f = D_jj delta_j f ,  d_jj f = D_jj^2 delta_jj f + D_jj SD_jjj delta_j f ,
 *       d_ij f = D_ii D_jj delta_out(delta_in f)  (in = lower, out = higher direction; metric.py:72-135,
 *       opensblifunctions.py:540-549) and S_ij = d_j u_i + d_i u_j - 2/3 delta_ij div u  the reference expands
 * (app strings, e.g. katzer_SBLI.py:10-14; StoreSome.py:71-161):
 *   momentum_i += 1/Re [ sum_j d_j mu S_ij + mu ( sum_j d_jj u_i + 1/3 sum_j d_ij u_j ) ]
 *   energy     += kq [ sum_j d_j mu d_j T + mu sum_j d_jj T ] + sum_i u_i (momentum_i term) + mu/Re sum_ij S_ij d_j u_i
 * ------------------------------------------------------------------------------------------- */
static void viscous_general(const osbo_cfg *c, const grid_t *g, const prim_t *P, double *const *R,
                            double *dv[5][3], const double *inv, const double *inv2) {
  const int nd = g->ndim;
  /* stored xi-derivatives of u_v (v < nd), T (v = nd): over ranges widened by +-2 in the other directions */
  for (int v = 0; v < nd + 1; v++) {
    const double *f = v < nd ? P->u[v] : P->T;
    for (int d = 0; d < nd; d++) {
      int l[3] = {0, 0, 0}, hh[3] = {1, 1, 1};
      for (int e = 0; e < nd; e++) { l[e] = -2; hh[e] = g->np[e] + 2; }
      l[d] = 0; hh[d] = g->np[d];
      for (int k = l[2]; k < hh[2]; k++) for (int j = l[1]; j < hh[1]; j++) for (int i = l[0]; i < hh[0]; i++) {
        long x = gidx(g, i, j, k);
        int id[3] = {i, j, k};
        dv[v][d][x] = d1g(c, f, x, g->s[d], inv[d], d, id[d], g->np[d]);
      }
    }
  }
  const double iRe = 1.0 / c->Re;
  const double kq = iRe * (1.0 / (c->gama - 1.0)) * pow(c->Minf, -2) * (1.0 / c->Pr);
  for (int k = 0; k < g->np[2]; k++) for (int j = 0; j < g->np[1]; j++) for (int i = 0; i < g->np[0]; i++) {
    long x = gidx(g, i, j, k);
    int id[3] = {i, j, k};
    double Dm[3], SDm[3], dmu[3], du[3][3], dT[3];
    for (int d = 0; d < nd; d++) {
      Dm[d] = c->D[d] ? c->D[d][x] : 1.0;
      SDm[d] = c->SD[d] ? c->SD[d][x] : 0.0;
      dmu[d] = c->visc_law == OSBO_MU_CONSTANT ? 0.0 : Dm[d] * d1g(c, P->mu, x, g->s[d], inv[d], d, id[d], g->np[d]);
      dT[d] = Dm[d] * dv[nd][d][x];
      for (int a = 0; a < nd; a++) du[a][d] = Dm[d] * dv[a][d][x];
    }
    double div = 0.0;
    for (int a = 0; a < nd; a++) div += du[a][a];
    const double mu = P->mu[x];
    double vis[3] = {0, 0, 0}, e = 0.0;
    for (int a = 0; a < nd; a++) {
      double s1 = 0.0, s2 = 0.0;
      for (int b = 0; b < nd; b++) {
        double Sab = du[a][b] + du[b][a] - (a == b ? (2.0 / 3.0) * div : 0.0);
        s1 += dmu[b] * Sab;
        double lap = Dm[b] * Dm[b] * d2g(c, P->u[a], x, g->s[b], inv2[b], b, id[b], g->np[b]) + Dm[b] * SDm[b] * dv[a][b][x];
        if (b == a) s2 += (4.0 / 3.0) * lap;
        else {
          s2 += lap;
          int in = a < b ? a : b, out = a < b ? b : a;
          s2 += (1.0 / 3.0) * Dm[a] * Dm[b] * d1g(c, dv[b][in], x, g->s[out], inv[out], out, id[out], g->np[out]);
        }
        e += iRe * mu * Sab * du[a][b];
      }
      vis[a] = iRe * (s1 + mu * s2);
      R[1 + a][x] += vis[a] - c->force[a];
      e += vis[a] * P->u[a][x] - c->force[a] * P->u[a][x];
    }
    double hT = 0.0;
    for (int d = 0; d < nd; d++)
      hT += dmu[d] * dT[d] + mu * (Dm[d] * Dm[d] * d2g(c, P->T, x, g->s[d], inv2[d], d, id[d], g->np[d]) + Dm[d] * SDm[d] * dv[nd][d][x]);
    R[nd + 1][x] += kq * hT + e;
  }
}

/* ---------------------------------------------------------------------------------------------
 * General Central(4) convective terms: one-sided closures next to walls (every first derivative of the convective
 * loops switches formula by grid index, opensblifunctions.py:523-534), diagonal metrics D_dd (metric.py:137-147) and the
 * two splittings the shipped apps use:
 *  form 0 (Blaisdell, Skew(), parsing.py:75-111):
 *     q_m:  -1/2 [ d(q_m u_j) + u_j d(q_m) + q_m d(u_j) ] ,  momentum_i: - d_i p ,  energy: - d_j(p u_j)
 *  form 1 (Feiereisen, compressible_TCF_Central/turbulent_channel.py:12-20):
 *     mass: - d_j(rho u_j)
 *     momentum_i: -1/2 [ d_j(rhou_i u_j) + rhou_j d_j(u_i) + u_i d_j(rhou_j) ] - d_i p
 *     energy:     -1/2 [ d_j(rhoE u_j) + rhou_j d_j(rhoE/rho) + (rhoE/rho) d_j(rhou_j) ] - d_j(p u_j)
 * ------------------------------------------------------------------------------------------- */
static int d1_stencil(const osbo_cfg *c, int dir, int idx, int n, double *w, int *off) {
  if (c->closure[dir][0] && idx < c->c_nr1) {
    for (int p = 0; p < c->c_np1; p++) { w[p] = c->c_d1[idx * c->c_np1 + p]; off[p] = p - idx; }
    return c->c_np1;
  }
  if (c->closure[dir][1] && n - 1 - idx < c->c_nr1) {
    const int row = n - 1 - idx;
    for (int p = 0; p < c->c_np1; p++) { w[p] = -c->c_d1[row * c->c_np1 + p]; off[p] = row - p; }
    return c->c_np1;
  }
  w[0] = 1.0 / 12.0; w[1] = -8.0 / 12.0; w[2] = 8.0 / 12.0; w[3] = -1.0 / 12.0;
  off[0] = -2; off[1] = -1; off[2] = 1; off[3] = 2;
  return 4;
}

static void central_general(const osbo_cfg *c, const grid_t *g, double *const *q, const prim_t *P, double *const *R, const double *inv) {
  const int nd = g->ndim, nv = g->nv;
  for (int k = 0; k < g->np[2]; k++) for (int j = 0; j < g->np[1]; j++) for (int i = 0; i < g->np[0]; i++) {
    const long x = gidx(g, i, j, k);
    const int id[3] = {i, j, k};
    double r[5] = {0, 0, 0, 0, 0};
    for (int d = 0; d < nd; d++) {
      double w[8]; int off[8];
      const int cnt = d1_stencil(c, d, id[d], g->np[d], w, off);
      const double sc = inv[d] * (c->D[d] ? c->D[d][x] : 1.0);
      double dqu[5] = {0, 0, 0, 0, 0}, dq[5] = {0, 0, 0, 0, 0}, du[3] = {0, 0, 0}, dp = 0.0, dpu = 0.0, dh = 0.0;
      for (int p = 0; p < cnt; p++) {
        const long xs = x + off[p] * g->s[d];
        const double ud = P->u[d][xs];
        for (int m = 0; m < nv; m++) { dqu[m] += w[p] * (q[m][xs] * ud); dq[m] += w[p] * q[m][xs]; }
        for (int a = 0; a < nd; a++) du[a] += w[p] * P->u[a][xs];
        dp += w[p] * P->p[xs]; dpu += w[p] * (P->p[xs] * ud);
        dh += w[p] * (q[nd + 1][xs] / q[0][xs]);
      }
      for (int m = 0; m < nv; m++) { dqu[m] *= sc; dq[m] *= sc; }
      for (int a = 0; a < nd; a++) du[a] *= sc;
      dp *= sc; dpu *= sc; dh *= sc;
      if (c->central_form == 0) {
        for (int m = 0; m < nv; m++) r[m] -= 0.5 * (dqu[m] + P->u[d][x] * dq[m] + q[m][x] * du[d]);
      } else {
        r[0] -= dq[1 + d];
        for (int a = 0; a < nd; a++) r[1 + a] -= 0.5 * (dqu[1 + a] + q[1 + d][x] * du[a] + P->u[a][x] * dq[1 + d]);
        r[nd + 1] -= 0.5 * (dqu[nd + 1] + q[1 + d][x] * dh + (q[nd + 1][x] / q[0][x]) * dq[1 + d]);
      }
      r[1 + d] -= dp;
      r[nd + 1] -= dpu;
    }
    for (int m = 0; m < nv; m++) R[m][x] = r[m];
  }
}

/* ---------------------------------------------------------------------------------------------
 * Spatial residual = what the stage's "spatial kernels" leave in Residual_m
 * ------------------------------------------------------------------------------------------- */
static int g_src_iter = -1;   /* iteration number seen by the mass source; -1: use c->src_iter0 */
void osbo_residual(const osbo_cfg *c, double *const *q, double *const *R) {
  grid_t g; grid_init(c, &g);
  const int nd = g.ndim, nv = g.nv;
  double *buf = (double *)calloc((size_t)g.n * (7 + 15 + 15), sizeof(double));
  prim_t P; double *w = buf;
  for (int d = 0; d < 3; d++) { P.u[d] = w; w += g.n; }
  P.p = w; w += g.n; P.a = w; w += g.n; P.T = w; w += g.n; P.mu = w; w += g.n;
  double *wk[3][5]; for (int d = 0; d < 3; d++) for (int m = 0; m < 5; m++) { wk[d][m] = w; w += g.n; }
  double *dv[5][3]; for (int v = 0; v < 5; v++) for (int d = 0; d < 3; d++) { dv[v][d] = w; w += g.n; }
  double inv[3], inv2[3];
  for (int d = 0; d < nd; d++) { inv[d] = 1.0 / c->delta[d]; inv2[d] = pow(c->delta[d], -2); }

  constituent(c, &g, q, &P);
  int general = c->visc_law != OSBO_MU_CONSTANT || c->force[0] != 0.0 || c->force[1] != 0.0 || c->force[2] != 0.0;
  for (int d = 0; d < nd; d++) if (c->D[d] || c->closure[d][0] || c->closure[d][1]) general = 1;

  if (c->teno_adaptive) {
    /* modified Ducros sensor (shock_sensors.py:12-49), evaluated on the interior only */
    for (int k = 0; k < g.np[2]; k++) for (int j = 0; j < g.np[1]; j++) for (int i = 0; i < g.np[0]; i++) {
      long x = gidx(&g, i, j, k);
      int id[3] = {i, j, k};
      double du[3][3];
      for (int a = 0; a < nd; a++) for (int b = 0; b < nd; b++)
        du[a][b] = d1g(c, P.u[a], x, g.s[b], inv[b], b, id[b], g.np[b]) * (c->D[b] ? c->D[b][x] : 1.0);
      double div = 0.0, vort = 0.0;
      for (int a = 0; a < nd; a++) div += du[a][a];
      if (nd == 2) vort = (du[1][0] - du[0][1]) * (du[1][0] - du[0][1]);
      else if (nd == 3) vort = (du[2][1] - du[1][2]) * (du[2][1] - du[1][2]) + (du[0][2] - du[2][0]) * (du[0][2] - du[2][0]) + (du[1][0] - du[0][1]) * (du[1][0] - du[0][1]);
      c->theta[x] = (0.5 - 0.5 * tanh(2.5 + 250.0 * div)) * div * div / (c->sensor_eps + div * div + vort);
    }
  }

  if (c->conv != OSBO_CONV_CENTRAL) {
    for (int d = 0; d < nd; d++) llf_flux(c, &g, d, q, &P, wk[d]);
    /* "Residual" kernel: shock_capturing.py:21-34; {Weno,Teno}Derivative opensblifunctions.py:575-593,658-676 */
    for (int k = 0; k < g.np[2]; k++) for (int j = 0; j < g.np[1]; j++) for (int i = 0; i < g.np[0]; i++) {
      long x = gidx(&g, i, j, k);
      for (int m = 0; m < nv; m++) {
        double r = 0.0;
        for (int d = 0; d < nd; d++) r -= inv[d] * (wk[d][m][x] - wk[d][m][x - g.s[d]]) * (c->D[d] ? c->D[d][x] : 1.0);
        R[m][x] = c->curv_detJ ? r / c->curv_detJ[x] : r;
      }
    }
  } else if (general || c->central_form != 0) {
    central_general(c, &g, q, &P, R, inv);
  } else {
    /* Central(4) skew-symmetric (Blaisdell) convective terms: parsing.py:75-111, scheme.py:187-271.
     *  mass:     -1/2 [ d(rho u_j)/dx_j + u_j d(rho)/dx_j + rho du_j/dx_j ]
     *  momentum: -1/2 [ d(rhou_i u_j)/dx_j + u_j d(rhou_i)/dx_j + rhou_i du_j/dx_j ] - dp/dx_i
     *  energy:   -1/2 [ d(rhoE u_j)/dx_j + u_j d(rhoE)/dx_j + rhoE du_j/dx_j ] - d(p u_j)/dx_j   */
    for (int k = 0; k < g.np[2]; k++) for (int j = 0; j < g.np[1]; j++) for (int i = 0; i < g.np[0]; i++) {
      long x = gidx(&g, i, j, k);
      double div = 0.0;
      for (int d = 0; d < nd; d++) div += d1c(P.u[d], x, g.s[d], inv[d]);
      for (int m = 0; m < nv; m++) {
        double cons = 0.0, adv = 0.0;
        for (int d = 0; d < nd; d++) {
          long s = g.s[d];
          /* derivative of the product q_m * u_d ("Convective terms group d" work arrays) */
          double fm2 = q[m][x - 2 * s] * P.u[d][x - 2 * s], fm1 = q[m][x - s] * P.u[d][x - s];
          double fp1 = q[m][x + s] * P.u[d][x + s], fp2 = q[m][x + 2 * s] * P.u[d][x + 2 * s];
          cons += (1.0 / 12.0) * inv[d] * (fm2 - 8.0 * fm1 + 8.0 * fp1 - fp2);
          adv += P.u[d][x] * d1c(q[m], x, s, inv[d]);
        }
        double r = -0.5 * (cons + adv + q[m][x] * div);
        if (m >= 1 && m <= nd) r -= d1c(P.p, x, g.s[m - 1], inv[m - 1]);
        if (m == nd + 1)
          for (int d = 0; d < nd; d++) {
            long s = g.s[d];
            r -= (1.0 / 12.0) * inv[d] * (P.p[x - 2 * s] * P.u[d][x - 2 * s] - 8.0 * P.p[x - s] * P.u[d][x - s]
                                          + 8.0 * P.p[x + s] * P.u[d][x + s] - P.p[x + 2 * s] * P.u[d][x + 2 * s]);
          }
        R[m][x] = r;
      }
    }
  }

  if (c->src_amp) {
    const double fac = sin(c->src_rate * (g_src_iter < 0 ? c->src_iter0 : g_src_iter));
    for (int k = 0; k < g.np[2]; k++) for (int j = 0; j < g.np[1]; j++) for (int i = 0; i < g.np[0]; i++) {
      long x = gidx(&g, i, j, k);
      R[0][x] += c->src_amp[x] * fac;
    }
  }
  if (c->viscous && general) viscous_general(c, &g, &P, R, dv, inv, inv2);
  if (c->viscous && !general) {
    /* Stored first derivatives d(u_v)/dx_d, d(T)/dx_d over ranges widened by +-2 in the other directions
     * (StoreSome.py:85-134 "Derivative evaluation"; scheme.py:256-267 "Viscous CD"). */
    int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    for (int d = 0; d < nd; d++) { lo[d] = -2; hi[d] = g.np[d] + 2; }
    for (int v = 0; v < nd + 1; v++) {
      const double *f = v < nd ? P.u[v] : P.T;
      for (int d = 0; d < nd; d++) {
        int l[3], hh[3];
        for (int e = 0; e < 3; e++) { l[e] = lo[e]; hh[e] = hi[e]; }
        l[d] = 0; hh[d] = g.np[d];
        for (int k = l[2]; k < hh[2]; k++) for (int j = l[1]; j < hh[1]; j++) for (int i = l[0]; i < hh[0]; i++) {
          long x = gidx(&g, i, j, k);
          dv[v][d][x] = d1c(f, x, g.s[d], inv[d]);
        }
      }
    }
    /* "Viscous terms": tau_ij = (1/Re)(du_i/dx_j + du_j/dx_i - 2/3 delta_ij div u),
     *  q_j = 1/((gama-1) Minf^2 Pr Re) dT/dx_j  (app strings, taylor_green_vortex.py:17-18);
     *  momentum_i += d tau_ij/dx_j ; energy += d q_j/dx_j + d(u_i tau_ij)/dx_j.
     *  Homogeneous second derivatives use the 5-point second-derivative formula; mixed ones are the
     *  derivative (outer, higher direction) of the stored derivative (inner, lower direction). */
    const double iRe = 1.0 / c->Re;
    const double kq = iRe * (1.0 / (c->gama - 1.0)) * pow(c->Minf, -2) * (1.0 / c->Pr);
    for (int k = 0; k < g.np[2]; k++) for (int j = 0; j < g.np[1]; j++) for (int i = 0; i < g.np[0]; i++) {
      long x = gidx(&g, i, j, k);
      double vis[3] = {0, 0, 0};
      for (int a = 0; a < nd; a++) {
        double s = 0.0;
        for (int b = 0; b < nd; b++) {
          if (b == a) s += (4.0 / 3.0) * d2c(P.u[a], x, g.s[a], inv2[a]);
          else {
            s += d2c(P.u[a], x, g.s[b], inv2[b]);
            int in = a < b ? a : b, out = a < b ? b : a; /* d/dx_out ( d u_b / dx_in ) */
            s += (1.0 / 3.0) * d1c(dv[b][in], x, g.s[out], inv[out]);
          }
        }
        vis[a] = iRe * s;
        R[1 + a][x] += vis[a];
      }
      double lapT = 0.0;
      for (int d = 0; d < nd; d++) lapT += d2c(P.T, x, g.s[d], inv2[d]);
      double e = kq * lapT, div = 0.0;
      for (int a = 0; a < nd; a++) div += dv[a][a][x];
      for (int a = 0; a < nd; a++) {
        for (int b = a + 1; b < nd; b++) { double sab = dv[a][b][x] + dv[b][a][x]; e += iRe * sab * sab; }
        e += iRe * (2.0 * dv[a][a][x] - (2.0 / 3.0) * div) * dv[a][a][x];
        e += vis[a] * P.u[a][x];
      }
      R[nd + 1][x] += e;
    }
  }
  free(buf);
}

/* ---------------------------------------------------------------------------------------------
 * Time loop: algorithm.py:440-474.
 *   iter: BCs ; [SBLI: save q_old = q] ; stage: CR, spatial, RK update, BCs.
 * RK SBLI (rk_sbli.py:102-133): q = q_old + dt*rknew[s]*R ; q_old += dt*rkold[s]*R.
 * RK LS  (rk_LS.py:139-166):    tmp = dt*R + A[s]*tmp ; q += B[s]*tmp.
 * ------------------------------------------------------------------------------------------- */
/* One stage of the loop for a decomposed run: stage < 0 = iteration start (BCs + rk_sbli save),
 * otherwise residual + RK update + BCs of that stage.  Faces marked OSBO_BC_EXCHANGE are left to the caller. */
int osbo_stage(const osbo_cfg *c, double *const *q, double *const *rk_reg, int stage) {
  grid_t g; grid_init(c, &g);
  const int nv = g.nv;
  if (stage < 0) {
    osbo_apply_bcs(c, q);
    if (c->rk == OSBO_RK_SBLI)
      for (int m = 0; m < nv; m++)
        for (int k = 0; k < g.np[2]; k++) for (int j = 0; j < g.np[1]; j++) for (int i = 0; i < g.np[0]; i++) {
          long x = gidx(&g, i, j, k); rk_reg[m][x] = q[m][x];
        }
    return 0;
  }
  double *Rbuf = (double *)calloc((size_t)g.n * nv, sizeof(double));
  if (!Rbuf) return 1;
  double *R[5]; for (int m = 0; m < nv; m++) R[m] = Rbuf + (size_t)m * g.n;
  osbo_residual(c, q, R);
  for (int m = 0; m < nv; m++)
    for (int k = 0; k < g.np[2]; k++) for (int j = 0; j < g.np[1]; j++) for (int i = 0; i < g.np[0]; i++) {
      long x = gidx(&g, i, j, k);
      if (c->rk == OSBO_RK_SBLI) {
        q[m][x] = c->dt * c->rk_b[stage] * R[m][x] + rk_reg[m][x];
        rk_reg[m][x] = c->dt * c->rk_a[stage] * R[m][x] + rk_reg[m][x];
      } else {
        rk_reg[m][x] = c->dt * R[m][x] + c->rk_a[stage] * rk_reg[m][x];
        q[m][x] = c->rk_b[stage] * rk_reg[m][x] + q[m][x];
      }
    }
  osbo_apply_bcs(c, q);
  free(Rbuf);
  return 0;
}

int osbo_advance(const osbo_cfg *c, double *const *q, double *const *rk_reg, int nsteps) {
  grid_t g; grid_init(c, &g);
  const int nv = g.nv;
  double *Rbuf = (double *)calloc((size_t)g.n * nv, sizeof(double));
  if (!Rbuf) return 1;
  double *R[5]; for (int m = 0; m < nv; m++) R[m] = Rbuf + (size_t)m * g.n;
  for (int it = 0; it < nsteps; it++) {
    g_src_iter = c->src_iter0 + it;
    osbo_apply_bcs(c, q);
    if (c->rk == OSBO_RK_SBLI)
      for (int m = 0; m < nv; m++)
        for (int k = 0; k < g.np[2]; k++) for (int j = 0; j < g.np[1]; j++) for (int i = 0; i < g.np[0]; i++) {
          long x = gidx(&g, i, j, k); rk_reg[m][x] = q[m][x];
        }
    for (int s = 0; s < c->nstages; s++) {
      osbo_residual(c, q, R);
      for (int m = 0; m < nv; m++)
        for (int k = 0; k < g.np[2]; k++) for (int j = 0; j < g.np[1]; j++) for (int i = 0; i < g.np[0]; i++) {
          long x = gidx(&g, i, j, k);
          if (c->rk == OSBO_RK_SBLI) {
            q[m][x] = c->dt * c->rk_b[s] * R[m][x] + rk_reg[m][x];
            rk_reg[m][x] = c->dt * c->rk_a[s] * R[m][x] + rk_reg[m][x];
          } else {
            rk_reg[m][x] = c->dt * R[m][x] + c->rk_a[s] * rk_reg[m][x];
            q[m][x] = c->rk_b[s] * rk_reg[m][x] + q[m][x];
          }
        }
      osbo_apply_bcs(c, q);
    }
  }
  g_src_iter = -1;
  free(Rbuf);
  return 0;
}
