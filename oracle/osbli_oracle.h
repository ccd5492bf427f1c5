/* osbli_oracle.h -- CPU restatement (plain C99) of the OpenSBLI per-timestep solver hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (opensbli_b200) never does.
 *
 * It restates, loop by loop and un-fused, what the reference's generated OPS-C program executes
 * for the canonical compressible Euler / Navier-Stokes system (ideal gas, conservative variables
 * rho, rhou_i, rhoE):
 *   program order ............ opensbli/code_generation/algorithm/algorithm.py:384-477
 *   constituent relations .... app scripts (strings); opensbli/equation_types/opensbliequations.py:297-321
 *   characteristic LLF flux .. opensbli/schemes/spatial/shock_capturing.py:357-536
 *   Roe / simple average ..... opensbli/schemes/spatial/averaging.py:31-114
 *   eigensystems ............. opensbli/physical_models/euler_eigensystem.py:57-135
 *   WENO-JS / WENO-Z ......... opensbli/schemes/spatial/weno.py:35-465
 *   TENO5 / TENO6 ............ opensbli/schemes/spatial/teno.py:57-465
 *   central-4 conv + viscous . opensbli/schemes/spatial/scheme.py:63-387, optimisations/StoreSome.py:71-161
 *   RK3 / low-storage RK ..... opensbli/schemes/temporal/rk_sbli.py:58-133, rk_LS.py:70-166
 *   periodic / Dirichlet ..... opensbli/core/boundary_conditions/periodic.py:42-56, dirichlet.py:28-41
 *
 * Parity pinning: the reference has NO tests or golden vectors of its own;
 * this restatement is pinned against (i) outputs of the reference itself run here -- its generated
 * C compiled against oracle/ops_seq.h into oracle/_ref/<config>/ref_seq -- and (ii) the fixtures in
 * tests/golden/ minted from those executables by tests/golden/make_golden.py.
 *
 * Data layout: one padded array per variable, x fastest, halo `halo` (=5, opsc.py:707-711) on both
 * sides of every active dimension:  idx(i,j,k) = (i+h) + px*((j+h) + py*(k+h)).
 */
#ifndef OSBLI_ORACLE_H
#define OSBLI_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

enum { OSBO_CONV_CENTRAL = 0, OSBO_CONV_WENO = 1, OSBO_CONV_TENO = 2 };
enum { OSBO_AVG_SIMPLE = 0, OSBO_AVG_ROE = 1 };
enum { OSBO_RK_SBLI = 0, OSBO_RK_LS = 1 };
enum { OSBO_BC_PERIODIC = 0, OSBO_BC_DIRICHLET = 1, OSBO_BC_EXCHANGE = 2 /* halo filled by the caller (decomposed run) */,
       OSBO_BC_ISOTHERMAL_WALL = 3, OSBO_BC_EXTRAPOLATION = 4, OSBO_BC_INLET_PRESSURE_EXTRAPOLATE = 5,
       OSBO_BC_SYMMETRY = 6, OSBO_BC_DIRICHLET_FIELD = 7 /* imposed state varies along the face */,
       OSBO_BC_ADIABATIC_WALL = 8, OSBO_BC_ZERO_GRADIENT_OUTLET = 9, OSBO_BC_PRESSURE_OUTLET = 10, OSBO_BC_INVISCID_WALL = 11,
       OSBO_BC_SPLIT = 14 /* SplitBC (bc_core.py:200-217): several boundary classes on one face, each over its own part of the plane */ };
enum { OSBO_MU_CONSTANT = 0, OSBO_MU_SUTHERLAND = 1, OSBO_MU_POWER = 2 };

typedef struct {
  int ndim;
  int np[3];
  int halo;        /* storage halo, 5 */
  int conv;        /* OSBO_CONV_* */
  int order;       /* central 4 | weno 5 | teno 5,6 */
  int weno_z;      /* 0 JS, 1 Z */
  int averaging;   /* OSBO_AVG_* */
  int viscous;     /* 0 Euler, 1 constant-viscosity Navier-Stokes */
  int rk;          /* OSBO_RK_* */
  int nstages;
  double rk_a[8];  /* SBLI: rkold ; LS: A */
  double rk_b[8];  /* SBLI: rknew ; LS: B */
  double gama, Minf, Re, Pr, dt, eps, teno_ct;
  double delta[3];
  int bc[3][2];
  double bc_q[3][2][5]; /* Dirichlet: conservative state imposed on boundary + halo points */
  /* ---- general path: stretched grids, variable viscosity, one-sided closures, wall/inflow/outflow BCs,
   *      adaptive TENO (BASELINE config 4, apps/katzer_SBLI) ---- */
  int visc_law;               /* OSBO_MU_*: mu = 1 | T^1.5 (1+S/Tr)/(T+S/Tr) | T^mu_exp */
  double SuthT, RefT, mu_exp;
  const double *D[3];         /* D_dd = d xi_d / d x_d (padded array) or NULL: direction not stretched (metric.py:137-147) */
  const double *SD[3];        /* SD_ddd (second-derivative metric, metric.py:149-212) or NULL */
  int closure[3][2];          /* 1: central derivatives switch to the one-sided closure near this face */
  int c_nr1, c_np1, c_nr2, c_np2;       /* closure tables: rows x points (boundary-absolute points 0..np-1) */
  double c_d1[4 * 6], c_d2[2 * 6];      /* ReducedAccess: 2x5, 2x5; Carpenter: 4x6, 2x5 */
  int teno_adaptive;          /* C_T from the Ducros sensor (teno.py:430-443) */
  double teno_a1, teno_a2, sensor_eps;
  double *theta;              /* sensor array (padded, written by the residual evaluation) or NULL */
  double *teno_store;         /* optional output of C_T of the direction-0 sweep ('TENO' dataset) or NULL */
  double Twall;
  int extrap_order[3][2];
  const double *bc_face[3][2];/* OSBO_BC_DIRICHLET_FIELD: [nv][padded tangential extent] */
  double force[3];            /* constant body force c_j: momentum_i -= c_i, energy -= c_j u_j (turbulent_channel.py:15-16) */
  int bc_free[3][2];          /* OSBO_BC_DIRICHLET_FIELD: bit m set = conserved variable m is NOT imposed on this face; bit 8 set = the
                               * imposed energy is the table value + 1/2 sum(free momentum^2)/rho (transitional_SBLI.py:134-139) */
  /* time-periodic mass source of apps/transitional_SBLI (transitional_SBLI.py:77-89): Residual_rho += src_amp(x) sin(src_rate * iter),
   * iter = src_iter0 + number of completed steps (the loop counter of algorithm.py:440-474) */
  const double *src_amp;      /* padded array or NULL */
  double src_rate;
  int src_iter0;
  /* fully curvilinear grid (strong-conservation form of apps/euler_wave_curvilinear/euler_wave.py:12-18): metric arrays
   * curv_D[i][j] = D_ij = d xi_i / d x_j and the Jacobian determinant detJ (padded, valid in the scheme halos); NULL: Cartesian
   * or diagonal metrics.  Eigensystems then carry the direction cosines k~ = D_i. / |D_i.| (euler_eigensystem.py:18-54). */
  const double *curv_D[3][3];
  const double *curv_detJ;
  double back_pressure;       /* OSBO_BC_PRESSURE_OUTLET: imposed outlet pressure (pressure_outlet.py:22-47) */
  /* OSBO_BC_SPLIT faces: parts in the order applied; each has a boundary kind and the evaluation range [lo, hi) per direction that
   * the reference takes from its run-time arrays split_range_<d><s><n> + split_halo_range_<d><s><n> (bc_core.py:110-127) */
  int split_n[3][2];
  int split_kind[3][2][8];
  int split_lo[3][2][8][3], split_hi[3][2][8][3];
  int split_order[3][2][8];   /* extrapolation order of a part */
  double split_q[3][2][8][5]; /* Dirichlet state of a part */
  int halo_m, halo_p;         /* depth of the boundary / periodic halos when not the scheme's own (0: default) */
  int central_form;           /* Central(4) convective split: 0 Blaisdell skew form (taylor_green_vortex.py:8-11, laminar_channel.py:7-9),
                               * 1 Feiereisen quadratic split (compressible_TCF_Central/turbulent_channel.py:12-20) */
} osbo_cfg;

/* number of doubles of one padded array */
long osbo_padded_size(const osbo_cfg *c);

/* Advance nsteps full RK steps.  q[m] (m < ndim+2) point to padded arrays; rk_reg[m] are the RK
 * register arrays (tempRK_* for LS, *_RKold for SBLI; zero-halo in the reference, padded here for
 * simplicity -- only interior is touched).  Returns 0 on success. */
int osbo_advance(const osbo_cfg *c, double *const *q, double *const *rk_reg, int nsteps);

/* stage < 0: iteration start (BCs, save); else one RK stage incl. its BCs (for decomposed runs) */
int osbo_stage(const osbo_cfg *c, double *const *q, double *const *rk_reg, int stage);
/* Pieces, exposed for unit tests. */
void osbo_apply_bcs(const osbo_cfg *c, double *const *q);
void osbo_residual(const osbo_cfg *c, double *const *q, double *const *R); /* CR + spatial kernels */
/* one-interface reconstructions on the 6-point window f[0..5] = f(i-2..i+3):
 * fp = 1/2(CF + lambda CS), fm = 1/2(CF - lambda CS); returns Recon (both sides summed). */
double osbo_recon_teno5(const double *fp, const double *fm, double eps, double ct);
double osbo_recon_teno6(const double *fp, const double *fm, double eps, double ct);
double osbo_recon_weno5(const double *fp, const double *fm, int z);
/* flux of one interface (direction dir) from the 6 conservative stencil states q6[p*nv+m], p <-> offsets -2..3 */
void osbo_interface_flux(const osbo_cfg *c, int dir, const double *q6, double *flux);

#ifdef __cplusplus
}
#endif
#endif
