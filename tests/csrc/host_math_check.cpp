// host_math_check.cpp -- compiles opensbli_b200/csrc/osb_math.cuh (the arithmetic every CUDA sweep kernel
// inlines) for the HOST, so the restructured eigen-projections / division-free TENO cut-off can be checked
// against the CPU oracle without a GPU.  Test infrastructure only; never loaded by the product.
#include "../../opensbli_b200/csrc/osb_math.cuh"
#include "../../opensbli_b200/csrc/osb_flux.cuh"
#include "../../opensbli_b200/csrc/osb_flux3.cuh"

using namespace osb;

template <int ND, int DIR, int RECON, int AVG>
static void run(const double *q6, double gama, double Minf, const SchemeParams &sp, double *flux) {
  Point<ND> pt[6];
  for (int p = 0; p < 6; p++) {
    const double *q = q6 + p * (ND + 2);
    pt[p].rho = q[0];
    double ke = 0.0;
    for (int d = 0; d < ND; d++) { pt[p].m[d] = q[1 + d]; pt[p].u[d] = q[1 + d] / q[0]; ke += 0.5 * q[0] * pt[p].u[d] * pt[p].u[d]; }
    pt[p].E = q[ND + 1];
    pt[p].pr = (gama - 1.0) * (q[ND + 1] - ke);
    pt[p].a = sqrt(gama * pt[p].pr / q[0]);
  }
  (void)Minf;
  interface_flux<ND, DIR, RECON, AVG>(pt, gama, sp, flux);
}

template <int ND, int DIR>
static int dispatch(int recon, int avg, const double *q6, double gama, const SchemeParams &sp, double *flux) {
#define CASE(R, A) if (recon == R && avg == A) { run<ND, DIR, R, A>(q6, gama, 0.0, sp, flux); return 0; }
  CASE(RECON_WENO5_JS, AVG_SIMPLE) CASE(RECON_WENO5_JS, AVG_ROE) CASE(RECON_WENO5_Z, AVG_SIMPLE) CASE(RECON_WENO5_Z, AVG_ROE)
  CASE(RECON_TENO5, AVG_SIMPLE) CASE(RECON_TENO5, AVG_ROE) CASE(RECON_TENO6, AVG_SIMPLE) CASE(RECON_TENO6, AVG_ROE)
#undef CASE
  return 1;
}

extern "C" int hostcheck_interface_flux(int nd, int dir, int recon, int avg, const double *q6, double gama,
                                        double eps, double ct, double *flux) {
  SchemeParams sp = make_scheme_params(eps, ct);
  if (nd == 1 && dir == 0) return dispatch<1, 0>(recon, avg, q6, gama, sp, flux);
  if (nd == 2 && dir == 0) return dispatch<2, 0>(recon, avg, q6, gama, sp, flux);
  if (nd == 2 && dir == 1) return dispatch<2, 1>(recon, avg, q6, gama, sp, flux);
  if (nd == 3 && dir == 0) return dispatch<3, 0>(recon, avg, q6, gama, sp, flux);
  if (nd == 3 && dir == 1) return dispatch<3, 1>(recon, avg, q6, gama, sp, flux);
  if (nd == 3 && dir == 2) return dispatch<3, 2>(recon, avg, q6, gama, sp, flux);
  return 1;
}

// v3: the pass-split form of the sweep kernels (osb_flux3.cuh): same staged window, intermediate split fluxes handed over
// through a column (here a plain array, stride 1)
template <int ND, int DIR, int RECON, int AVG>
static void run_split(const double *q6, double gama, const SchemeParams &sp, double *flux) {
  typedef SV<ND> V;
  double st[V::N * 6];
  for (int p = 0; p < 6; p++) {                   // what stage_values<ND, DIR> (osb_types.cuh) leaves in shared memory
    const double *q = q6 + p * (ND + 2);
    const double y = 1.0 / sqrt(q[0]), irho = y * y;
    double mu = 0.0;
    for (int d = 0; d < ND; d++) {
      st[(V::M0 + d) * 6 + p] = q[1 + d];
      const double u = q[1 + d] * irho;
      if (d == DIR) st[V::UD * 6 + p] = u;
      mu += q[1 + d] * u;
    }
    const double pr = (gama - 1.0) * (q[ND + 1] - 0.5 * mu);
    st[V::RHO * 6 + p] = q[0]; st[V::Y * 6 + p] = y; st[V::E * 6 + p] = q[ND + 1];
    st[V::P * 6 + p] = pr; st[V::A * 6 + p] = sqrt(gama * pr * irho);
  }
  double G[F3<RECON>::NCOL];
  interface_flux_split<ND, DIR, RECON, AVG>(st, WinAffine{1}, 6, G, 1, gama, sp, flux);
}

template <int ND, int DIR>
static int dispatch_split(int recon, int avg, const double *q6, double gama, const SchemeParams &sp, double *flux) {
#define CASE(R, A) if (recon == R && avg == A) { run_split<ND, DIR, R, A>(q6, gama, sp, flux); return 0; }
  CASE(RECON_WENO5_JS, AVG_SIMPLE) CASE(RECON_WENO5_JS, AVG_ROE) CASE(RECON_WENO5_Z, AVG_SIMPLE) CASE(RECON_WENO5_Z, AVG_ROE)
  CASE(RECON_TENO5, AVG_SIMPLE) CASE(RECON_TENO5, AVG_ROE) CASE(RECON_TENO6, AVG_SIMPLE) CASE(RECON_TENO6, AVG_ROE)
#undef CASE
  return 1;
}

extern "C" int hostcheck_interface_flux_split(int nd, int dir, int recon, int avg, const double *q6, double gama,
                                              double eps, double ct, double *flux) {
  SchemeParams sp = make_scheme_params(eps, ct);
  if (nd == 1 && dir == 0) return dispatch_split<1, 0>(recon, avg, q6, gama, sp, flux);
  if (nd == 2 && dir == 0) return dispatch_split<2, 0>(recon, avg, q6, gama, sp, flux);
  if (nd == 2 && dir == 1) return dispatch_split<2, 1>(recon, avg, q6, gama, sp, flux);
  if (nd == 3 && dir == 0) return dispatch_split<3, 0>(recon, avg, q6, gama, sp, flux);
  if (nd == 3 && dir == 1) return dispatch_split<3, 1>(recon, avg, q6, gama, sp, flux);
  if (nd == 3 && dir == 2) return dispatch_split<3, 2>(recon, avg, q6, gama, sp, flux);
  return 1;
}
