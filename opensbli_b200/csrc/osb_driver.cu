// osb_driver.cu -- context, plan parser, launch orchestration and the C ABI (include/osbli_b200.h).
//
// The step order reproduces TraditionalAlgorithmRK.generate_solution
// (opensbli/code_generation/algorithm/algorithm.py:440-474):
//   per iteration: BC kernels/exchanges (dir0 side0, dir0 side1, dir1 side0, ...) ; [rk_sbli: save]
//   per stage:     constituent relations ; spatial kernels ; RK update ; BC kernels/exchanges
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "../../include/osbli_b200.h"
#include "osb_kernels.cuh"
#include "osb_flux_api.h"
#include "osb_tma_host.h"

using namespace osb;

namespace {

thread_local std::string g_create_error;

enum ConvKind { CONV_CENTRAL = 0, CONV_WENO = 1, CONV_TENO = 2, CONV_GENERIC = 3 /* every loop of the step is a run-time compiled kernel */ };
enum RkKind { RK_SBLI = 0, RK_LS = 1 };
enum BcKind { BC_PERIODIC = 0, BC_DIRICHLET = 1, BC_EXCHANGE = 2 /* neighbour rank owns the halo */, BC_ISOTHERMAL_WALL = 3,
              BC_EXTRAPOLATION = 4, BC_INLET_PRESSURE = 5, BC_SYMMETRY = 6, BC_DIRICHLET_FIELD = 7, BC_ADIABATIC_WALL = 8,
              BC_ZERO_GRADIENT = 9, BC_PRESSURE_OUTLET = 10, BC_INVISCID_WALL = 11,
              BC_GENERIC = 12 /* run-time compiled kernel registered with when = 100 + 2 dir + side */,
              BC_SPLIT = 14 /* several boundary classes share the face, each over its own part of the plane (SplitBC, bc_core.py:200-217) */,
              BC_OPEN = 13 /* halo left as uploaded: the face of a window cut out of a larger block (guard planes absorb the error) */ };

struct BcPart {              // one part of a split face: boundary class + evaluation range [lo, hi) per direction (halo extension included)
  int kind = BC_SYMMETRY, order = 0;
  int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  double q[5] = {0, 0, 0, 0, 0};
};

struct BcSpec {
  std::vector<BcPart> parts; // BC_SPLIT
  int kind = BC_PERIODIC;
  double q[5] = {0, 0, 0, 0, 0};
  int order = 0;            // extrapolation order
  int free_mask = 0;        // dirichlet_field: variables left free (bit m) / kinetic energy of the free momenta added to the energy (bit 8)
  bool closure = false;     // central derivatives use the one-sided closure next to this face
};

struct Plan {
  int nd = 0;
  int np[3] = {1, 1, 1};
  double delta[3] = {1, 1, 1};
  int conv = CONV_TENO, order = 5;
  bool weno_z = false;
  int averaging = AVG_ROE;
  bool viscous = false;
  int rk = RK_LS;
  std::vector<double> rk_a, rk_b;
  std::map<std::string, double> consts;
  BcSpec bc[3][2];
  // general path
  int visc_law = 0;
  double mu_exp = 0.0;
  bool metric[3] = {false, false, false};
  bool teno_adaptive = false;
  bool forcing = false;
  int generic_stages = 0;     // CONV_GENERIC: number of RK stages; the step is user kernels with when = 200 (iteration start), 209 (every stage), 210 + s
  int halo_m = 0, halo_p = 0; // depth of the boundary / exchange halos when it is not the scheme's own (0: 2/2 central, 3/4 WENO / TENO)
  int central_form = 0;      // 0 Blaisdell skew form, 1 Feiereisen quadratic split
  bool curvilinear = false;  // full metric tensor D_ij + detJ (2-D strong-conservation form)
  bool mass_source = false;  // Residual_rho += BF_amp(x) sin(src_rate * iteration)
  double src_rate = 0.0;
  long long iteration0 = 0;
  Closures cl{};
};

struct Field {
  std::string name;
  double *dev = nullptr;
};

}  // namespace

struct osb_ctx {
  Plan plan;
  int device = 0;
  cudaStream_t stream = nullptr;
  GridDev grid{};
  FieldPtrs fp{};
  PhysConst pc{};
  SchemeParams sp{};
  std::vector<Field> fields;
  std::string error;
  long long launches = 0;
  // profiling
  bool profiling = false;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> prof_events;
  // peers (slab decomposition)
  cudaEvent_t timer0 = nullptr, timer1 = nullptr;
  GeneralPtrs gp{};
  AdaptiveCT ad{};
  CurvPtrs cp{};
  bool general = false;
  double *face_table[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  long long face_size[3] = {0, 0, 0};
  // neighbours' arrays: [side][set][m], set 0 = the buffers that held q at creation, set 1 = those that held Residual (the
  // two sets exchange roles every stage on the out-of-place paths; all ranks step in lockstep, so parities agree)
  double *peer_buf[2][2][5] = {{{nullptr}}};
  double *own_buf[2][5] = {{nullptr}};
  bool peer_open[2] = {false, false};
  long long peer_np[2] = {0, 0};                // neighbours' slab thickness along the slab axis
  // stream-ordered neighbour synchronisation (flag words written by the neighbours through peer pointers)
  unsigned long long *flags = nullptr;          // [0] low nbr read-done, [1] low nbr pushed, [2] high nbr read-done, [3] high nbr pushed, [7] error
  unsigned long long *peer_flags[2] = {nullptr, nullptr};
  unsigned long long epoch_sig[2] = {0, 0}, epoch_wait[2] = {0, 0};
  // adaptive TENO in a decomposed run: the sweep along the slab axis reads the shock sensor one plane below the slab
  double *peer_theta[2] = {nullptr, nullptr};   // neighbours' theta arrays
  unsigned long long epoch_theta = 0;           // flag slot [4]: "low neighbour has stored its top sensor plane"
  // one time step captured as a CUDA graph (launch-bound small grids): invalidated when a constant changes
  cudaGraphExec_t step_graph = nullptr;
  long long graph_launches = 0;
  bool use_graph = true;
  long long iteration = 0;                      // loop counter of algorithm.py:440-474 (argument of the mass source)
  struct UserKernel { cudaLibrary_t lib = nullptr; cudaKernel_t kern = nullptr; std::vector<std::string> fields; int range[6]; int when = 0; bool writes_state = false; bool uses_iter = false; };
  std::vector<UserKernel> user_kernels;
  unsigned long long *slow_count = nullptr;     // bench instrumentation (osb_slow_path_count)
  bool count_slow = false;
  double *diag_buf = nullptr;                   // partial sums of the diagnostics kernels
  bool prim_stale = false;                      // u, p, a, T arrays lag the state (stage kernels that derive them on the fly)
  int swap_parity = 0;                          // fused central path: q and Residual buffers exchange roles every stage
  // window pipeline over a host-resident state (osb_host_planes_*): copy streams ordered against the compute stream by events
  cudaStream_t s_up = nullptr, s_down = nullptr;
  cudaEvent_t ev_up = nullptr, ev_comp = nullptr, ev_down = nullptr;
  bool down_pending = false;
};

namespace {

#define OSB_CUDA(ctx, call)                                                                 \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      (ctx)->error = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

int fail(osb_ctx *ctx, const std::string &msg) {
  ctx->error = msg;
  return 1;
}

bool parse_plan(const std::string &text, Plan &P, std::string &err) {
  std::istringstream in(text);
  std::string line;
  bool header = false;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string key;
    if (!(ls >> key) || key[0] == '#') continue;
    if (key == "osbli_plan") { int v; ls >> v; if (v != 1) { err = "unsupported plan version"; return false; } header = true; }
    else if (key == "ndim") ls >> P.nd;
    else if (key == "np") { for (int d = 0; d < P.nd; d++) ls >> P.np[d]; }
    else if (key == "delta") { for (int d = 0; d < P.nd; d++) ls >> P.delta[d]; }
    else if (key == "conv") { std::string v; ls >> v; P.conv = v == "central" ? CONV_CENTRAL : v == "weno" ? CONV_WENO : v == "teno" ? CONV_TENO : v == "generic" ? CONV_GENERIC : -1; if (P.conv < 0) { err = "unknown conv scheme " + v; return false; } }
    else if (key == "order") ls >> P.order;
    else if (key == "weno_formulation") { std::string v; ls >> v; P.weno_z = (v == "Z"); }
    else if (key == "averaging") { std::string v; ls >> v; P.averaging = v == "roe" ? AVG_ROE : AVG_SIMPLE; }
    else if (key == "viscous") { int v; ls >> v; P.viscous = v != 0; }
    else if (key == "generic_stages") { ls >> P.generic_stages; if (P.generic_stages < 1 || P.generic_stages > 8) { err = "generic_stages must lie in 1..8"; return false; } }
    else if (key == "halos") { ls >> P.halo_m >> P.halo_p; if (P.halo_m < 2 || P.halo_p < 2 || P.halo_m > 5 || P.halo_p > 5) { err = "halos must lie in 2..5"; return false; } }
    else if (key == "rk") { std::string v; ls >> v; P.rk = v == "sbli" ? RK_SBLI : RK_LS; }
    else if (key == "rk_a") { double v; while (ls >> v) P.rk_a.push_back(v); }
    else if (key == "rk_b") { double v; while (ls >> v) P.rk_b.push_back(v); }
    else if (key == "const") { std::string n; double v; ls >> n >> v; P.consts[n] = v; }
    else if (key == "bc") {
      int d, s; std::string kind; ls >> d >> s >> kind;
      if (d < 0 || d > 2 || s < 0 || s > 1) { err = "bad bc line: " + line; return false; }
      BcSpec &b = P.bc[d][s];
      if (kind == "periodic") b.kind = BC_PERIODIC;
      else if (kind == "exchange") b.kind = BC_EXCHANGE;
      else if (kind == "open") b.kind = BC_OPEN;
      else if (kind == "split") b.kind = BC_SPLIT;
      else if (kind == "dirichlet") { b.kind = BC_DIRICHLET; for (int m = 0; m < P.nd + 2; m++) ls >> b.q[m]; }
      else if (kind == "dirichlet_field") b.kind = BC_DIRICHLET_FIELD;
      else if (kind == "isothermal_wall") b.kind = BC_ISOTHERMAL_WALL;
      else if (kind == "adiabatic_wall") b.kind = BC_ADIABATIC_WALL;
      else if (kind == "extrapolation") { b.kind = BC_EXTRAPOLATION; ls >> b.order; }
      else if (kind == "inlet_pressure_extrapolate") { b.kind = BC_INLET_PRESSURE; if (s != 0) { err = "inlet_pressure_extrapolate is defined for side 0 only"; return false; } }
      else if (kind == "symmetry") b.kind = BC_SYMMETRY;
      else if (kind == "zero_gradient_outlet") b.kind = BC_ZERO_GRADIENT;
      else if (kind == "generic") b.kind = BC_GENERIC;
      else if (kind == "inviscid_wall") b.kind = BC_INVISCID_WALL;
      else if (kind == "pressure_outlet") { b.kind = BC_PRESSURE_OUTLET; if (s != 1) { err = "pressure_outlet is defined for side 1 only"; return false; } }
      else { err = "unsupported boundary condition '" + kind + "'"; return false; }
      std::string tok;
      while (ls >> tok) {
        if (tok == "closure") b.closure = true;
        else if (tok == "free") { int m; if (!(ls >> m) || m < 0 || m > 4) { err = "bad bc line: " + line; return false; } b.free_mask |= 1 << m; }
        else if (tok == "ke_free") b.free_mask |= 256;
        else { err = "bad bc line: " + line; return false; }
      }
    }
    else if (key == "bc_part") {
      int d, s; std::string kind; ls >> d >> s >> kind;
      if (d < 0 || d >= P.nd || s < 0 || s > 1 || P.bc[d][s].kind != BC_SPLIT) { err = "bc_part without a 'split' face: " + line; return false; }
      BcPart part;
      static const std::map<std::string, int> kinds = {{"dirichlet", BC_DIRICHLET}, {"isothermal_wall", BC_ISOTHERMAL_WALL}, {"adiabatic_wall", BC_ADIABATIC_WALL},
        {"extrapolation", BC_EXTRAPOLATION}, {"inlet_pressure_extrapolate", BC_INLET_PRESSURE}, {"symmetry", BC_SYMMETRY}, {"zero_gradient_outlet", BC_ZERO_GRADIENT},
        {"pressure_outlet", BC_PRESSURE_OUTLET}, {"inviscid_wall", BC_INVISCID_WALL}};
      auto it = kinds.find(kind);
      if (it == kinds.end()) { err = "unsupported boundary condition '" + kind + "' in a split face"; return false; }
      part.kind = it->second;
      for (int e = 0; e < P.nd; e++) if (!(ls >> part.lo[e] >> part.hi[e]) || part.hi[e] <= part.lo[e]) { err = "bad bc_part range: " + line; return false; }
      if (part.hi[d] - part.lo[d] != 1) { err = "a split part covers one plane along its face normal: " + line; return false; }
      if (part.kind == BC_DIRICHLET) { for (int m = 0; m < P.nd + 2; m++) if (!(ls >> part.q[m])) { err = "bad bc_part line: " + line; return false; } }
      else ls >> part.order;
      if ((part.kind == BC_PRESSURE_OUTLET && s != 1) || (part.kind == BC_INLET_PRESSURE && s != 0)) { err = "boundary class on the wrong side: " + line; return false; }
      P.bc[d][s].parts.push_back(part);
    }
    else if (key == "viscosity") {
      std::string v; ls >> v;
      if (v == "constant") P.visc_law = 0; else if (v == "sutherland") P.visc_law = 1;
      else if (v == "power") { P.visc_law = 2; ls >> P.mu_exp; } else { err = "unknown viscosity law " + v; return false; }
    }
    else if (key == "metric") { int d2, on; ls >> d2 >> on; if (d2 < 0 || d2 > 2) { err = "bad metric line"; return false; } P.metric[d2] = on != 0; }
    else if (key == "teno_adaptive") { int v; ls >> v; P.teno_adaptive = v != 0; }
    else if (key == "forcing") { int v; ls >> v; P.forcing = v != 0; }
    else if (key == "curvilinear") { int v; ls >> v; P.curvilinear = v != 0; }
    else if (key == "mass_source") { ls >> P.src_rate >> P.iteration0; P.mass_source = true; }
    else if (key == "central_form") { std::string v; ls >> v; P.central_form = v == "blaisdell" ? 0 : v == "feiereisen" ? 1 : -1; if (P.central_form < 0) { err = "unknown central_form " + v; return false; } }
    else if (key == "closure_d1" || key == "closure_d2") {
      int nr, np; ls >> nr >> np;
      if (nr < 1 || np < 1 || nr > (key == "closure_d1" ? 4 : 2) || np > 6) { err = "closure table too large: " + line; return false; }
      double *dst = key == "closure_d1" ? P.cl.d1 : P.cl.d2;
      for (int i = 0; i < nr * np; i++) ls >> dst[i];
      if (key == "closure_d1") { P.cl.nr1 = nr; P.cl.np1 = np; } else { P.cl.nr2 = nr; P.cl.np2 = np; }
    } else { err = "unknown plan key '" + key + "'"; return false; }
    if (ls.fail() && !ls.eof()) { err = "malformed plan line: " + line; return false; }
  }
  if (!header) { err = "missing 'osbli_plan 1' header"; return false; }
  if (P.nd < 1 || P.nd > 3) { err = "ndim must be 1, 2 or 3"; return false; }
  for (int d = 0; d < P.nd; d++) if (P.np[d] < 6) { err = "np must be >= 6 in every direction"; return false; }
  if (P.rk_a.empty() || P.rk_a.size() != P.rk_b.size()) { err = "rk_a / rk_b missing or of different length"; return false; }
  if (P.conv == CONV_CENTRAL && P.order != 4) { err = "central scheme: only order 4 is implemented"; return false; }
  if (P.conv == CONV_WENO && P.order != 5) { err = "WENO: only order 5 (k=3) is implemented"; return false; }
  if (P.conv == CONV_TENO && P.order != 5 && P.order != 6) { err = "TENO: only orders 5 and 6 are implemented"; return false; }
  if (P.conv == CONV_GENERIC) {
    if (!P.generic_stages) { err = "conv generic needs generic_stages"; return false; }
    if (P.halo_m <= 0) { err = "conv generic needs halos"; return false; }
    P.consts.emplace("gama", 1.4); P.consts.emplace("dt", 0.0);       // unused: the printed kernels carry their own constants
  } else if (P.generic_stages) { err = "generic_stages needs conv generic"; return false; }
  for (const char *k : {"gama", "dt"}) if (!P.consts.count(k)) { err = std::string("missing constant ") + k; return false; }
  if (P.viscous) for (const char *k : {"Re", "Pr", "Minf"}) if (!P.consts.count(k)) { err = std::string("missing constant ") + k; return false; }
  bool any_closure = false;
  for (int d = 0; d < P.nd; d++) for (int s = 0; s < 2; s++) { P.cl.on[d][s] = P.bc[d][s].closure ? 1 : 0; any_closure |= P.bc[d][s].closure; }
  if (any_closure && (P.cl.nr1 == 0 || P.cl.nr2 == 0)) { err = "a face requests a derivative closure but closure_d1/closure_d2 tables are missing"; return false; }
  if (P.teno_adaptive) {
    if (P.conv != CONV_TENO) { err = "teno_adaptive needs a TENO scheme"; return false; }
    for (const char *k : {"teno_a1", "teno_a2"}) if (!P.consts.count(k)) { err = std::string("missing constant ") + k; return false; }
  }
  if (P.forcing && !P.viscous) { err = "body forcing is implemented together with the viscous terms only"; return false; }
  if (P.visc_law == 1) for (const char *k : {"SuthT", "RefT"}) if (!P.consts.count(k)) { err = std::string("missing constant ") + k; return false; }
  for (int d = 0; d < P.nd; d++) for (int s = 0; s < 2; s++) {
    if (P.bc[d][s].kind == BC_ISOTHERMAL_WALL && !P.consts.count("Twall")) { err = "missing constant Twall"; return false; }
    if (P.bc[d][s].kind == BC_PRESSURE_OUTLET && !P.consts.count("back_pressure")) { err = "missing constant back_pressure"; return false; }
    if (P.bc[d][s].kind == BC_SPLIT && P.bc[d][s].parts.empty()) { err = "split face without parts"; return false; }
    for (const BcPart &part : P.bc[d][s].parts) {
      if (part.kind == BC_ISOTHERMAL_WALL && !P.consts.count("Twall")) { err = "missing constant Twall"; return false; }
      if (part.kind == BC_PRESSURE_OUTLET && !P.consts.count("back_pressure")) { err = "missing constant back_pressure"; return false; }
      for (int e = 0; e < P.nd; e++) if (part.lo[e] < -5 || part.hi[e] > P.np[e] + 5) { err = "split part outside the padded block"; return false; }
    }
  }
  return true;
}

void scheme_halos(const Plan &P, int &hm, int &hp) {
  if (P.halo_m > 0) { hm = P.halo_m; hp = P.halo_p; return; }     // set by the plan: a block with further consumers of the halos (filters)
  if (P.conv == CONV_CENTRAL) { hm = 2; hp = 2; } else { hm = 3; hp = 4; }
}

void refresh_constants(osb_ctx *c) {
  const Plan &P = c->plan;
  auto get = [&](const char *k, double d) { auto it = P.consts.find(k); return it == P.consts.end() ? d : it->second; };
  c->pc.gama = get("gama", 1.4); c->pc.Minf = get("Minf", 1.0); c->pc.Re = get("Re", 1.0); c->pc.Pr = get("Pr", 1.0);
  if (P.visc_law == 0) c->pc.Re /= get("mu", 1.0);      // constant-viscosity apps may carry a constant `mu` in mu/Re (viscous_shock_tube.py:14-16)
  c->pc.dt = get("dt", 0.0);
  for (int d = 0; d < 3; d++) { c->pc.inv[d] = 1.0 / P.delta[d]; c->pc.inv2[d] = pow(P.delta[d], -2); }
  c->sp = make_scheme_params(get("eps", 1e-16), get("TENO_CT", 1e-6));
  c->sp.slow_count = c->count_slow ? c->slow_count : nullptr;
  c->ad = make_adaptive_ct(P.teno_adaptive, get("teno_a1", 0.0), get("teno_a2", 0.0));
  c->pc.visc_law = P.visc_law; c->pc.mu_exp = P.mu_exp;
  c->pc.SuthT = get("SuthT", 0.0); c->pc.RefT = get("RefT", 1.0); c->pc.Twall = get("Twall", 1.0);
  c->pc.sensor_eps = get("epsilon", 1e-12);
  c->pc.src_factor = P.mass_source ? sin(P.src_rate * (double)c->iteration) : 0.0;
  for (int d = 0; d < 3; d++) c->pc.force[d] = P.forcing ? get(("c" + std::to_string(d)).c_str(), 0.0) : 0.0;
}

Field *find_field(osb_ctx *c, const char *name) {
  std::string n(name);
  if (n.size() > 3 && n.compare(n.size() - 3, 3, "_B0") == 0) n.resize(n.size() - 3);
  for (auto &f : c->fields) if (f.name == n) return &f;
  return nullptr;
}

// ---- launch helpers ----------------------------------------------------------------------------
struct Launcher {
  osb_ctx *c;
  int fam;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  Launcher(osb_ctx *ctx, int family) : c(ctx), fam(family) {
    if (c->profiling) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, c->stream); }
  }
  ~Launcher() {
    c->launches++;
    if (c->profiling) { cudaEventRecord(e1, c->stream); c->prof_events.push_back({fam, {e0, e1}}); }
  }
};

dim3 grid3(int n0, int n1, int n2, dim3 b) { return dim3((n0 + b.x - 1) / b.x, (n1 + b.y - 1) / b.y, (n2 + b.z - 1) / b.z); }

template <int ND>
void launch_prim(osb_ctx *c) {
  int hm, hp; scheme_halos(c->plan, hm, hp);
  int lo[3] = {0, 0, 0}, n[3] = {1, 1, 1};
  for (int d = 0; d < ND; d++) { lo[d] = -hm; n[d] = c->grid.np[d] + hm + hp; }
  dim3 b(128, 2, 1);
  Launcher L(c, OSB_FAM_PRIM);
  k_prim<ND><<<grid3(n[0], n[1], n[2], b), b, 0, c->stream>>>(c->grid, c->fp, c->pc, c->gp.mu, lo[0], lo[1], lo[2], n[0], n[1], n[2]);
  c->prim_stale = false;
}

void neighbour_signal(osb_ctx *c, int kind);

// Sweep order: in 3-D the z sweep (the only one that reads the halos owned by neighbour ranks of a slab decomposition)
// goes first, so that the "read done" notification overlaps with the x and y sweeps.
// third-generation sweeps (osb_flux3.cuh, one translation unit per ndim x reconstruction)
template <int ND, int RECON, int AVG>
void launch_flux(osb_ctx *c) {
  FluxArgs a{c->grid, c->fp, c->pc, c->sp, c->ad, c->gp};
  if (ND >= 3) {
    { Launcher L(c, OSB_FAM_FLUX); flux3_sweep(ND, RECON, 2, AVG, false, a, c->stream); }
    neighbour_signal(c, 0);
  }
  { Launcher L(c, OSB_FAM_FLUX); flux3_sweep(ND, RECON, 0, AVG, ND >= 3, a, c->stream); }
  if (ND >= 2) { Launcher L(c, OSB_FAM_FLUX); flux3_sweep(ND, RECON, 1, AVG, true, a, c->stream); }
  if (ND < 3) neighbour_signal(c, 0);
}

template <int RECON, int AVG>
void launch_flux_curv2d(osb_ctx *c) {
  const GridDev &g = c->grid;
  dim3 bf(128, 1, 1), br(64, 4, 1);
  {
    Launcher L(c, OSB_FAM_FLUX);
    k_flux_curv2d<0, RECON, AVG><<<dim3((g.np[0] + 1 + 127) / 128, g.np[1], 1), bf, 0, c->stream>>>(g, c->fp, c->pc, c->sp, c->cp);
  }
  { Launcher L(c, OSB_FAM_FLUX); k_resid_curv2d<0, false><<<grid3(g.np[0], g.np[1], 1, br), br, 0, c->stream>>>(g, c->fp, c->pc, c->cp); }
  {
    Launcher L(c, OSB_FAM_FLUX);
    k_flux_curv2d<1, RECON, AVG><<<dim3((g.np[1] + 1 + 127) / 128, g.np[0], 1), bf, 0, c->stream>>>(g, c->fp, c->pc, c->sp, c->cp);
  }
  { Launcher L(c, OSB_FAM_FLUX); k_resid_curv2d<1, true><<<grid3(g.np[0], g.np[1], 1, br), br, 0, c->stream>>>(g, c->fp, c->pc, c->cp); }
}

template <int ND>
void launch_flux_scheme(osb_ctx *c) {
  if (ND == 2 && c->plan.curvilinear) {
    const Plan &Pc = c->plan;
    const bool roe_ = Pc.averaging == AVG_ROE;
    if (Pc.conv == CONV_TENO && Pc.order == 5) { roe_ ? launch_flux_curv2d<RECON_TENO5, AVG_ROE>(c) : launch_flux_curv2d<RECON_TENO5, AVG_SIMPLE>(c); }
    else if (Pc.conv == CONV_TENO) { roe_ ? launch_flux_curv2d<RECON_TENO6, AVG_ROE>(c) : launch_flux_curv2d<RECON_TENO6, AVG_SIMPLE>(c); }
    else if (Pc.weno_z) { roe_ ? launch_flux_curv2d<RECON_WENO5_Z, AVG_ROE>(c) : launch_flux_curv2d<RECON_WENO5_Z, AVG_SIMPLE>(c); }
    else { roe_ ? launch_flux_curv2d<RECON_WENO5_JS, AVG_ROE>(c) : launch_flux_curv2d<RECON_WENO5_JS, AVG_SIMPLE>(c); }
    neighbour_signal(c, 0);
    return;
  }
  const Plan &P = c->plan;
  const bool roe = P.averaging == AVG_ROE;
  if (P.conv == CONV_TENO && P.order == 5) { roe ? launch_flux<ND, RECON_TENO5, AVG_ROE>(c) : launch_flux<ND, RECON_TENO5, AVG_SIMPLE>(c); }
  else if (P.conv == CONV_TENO) { roe ? launch_flux<ND, RECON_TENO6, AVG_ROE>(c) : launch_flux<ND, RECON_TENO6, AVG_SIMPLE>(c); }
  else if (P.weno_z) { roe ? launch_flux<ND, RECON_WENO5_Z, AVG_ROE>(c) : launch_flux<ND, RECON_WENO5_Z, AVG_SIMPLE>(c); }
  else { roe ? launch_flux<ND, RECON_WENO5_JS, AVG_ROE>(c) : launch_flux<ND, RECON_WENO5_JS, AVG_SIMPLE>(c); }
}

bool has_exchange(const osb_ctx *c) {
  const int d = c->plan.nd - 1;
  return (c->plan.bc[d][0].kind == BC_EXCHANGE && c->peer_open[0]) || (c->plan.bc[d][1].kind == BC_EXCHANGE && c->peer_open[1]);
}

bool fused_push_enabled() {
  static const bool on = getenv("OSB_NO_FUSED_PUSH") == nullptr;
  return on;
}

// out_of_place: the launching kernel writes the new state into the Residual-role buffers, so the neighbours' copies go into
// THEIR Residual-role buffers (which become their q after the stage)
PeerPush peer_push(const osb_ctx *c, bool out_of_place = false) {
  PeerPush pp;
  const int set = out_of_place ? 1 - c->swap_parity : c->swap_parity;
  if (!fused_push_enabled()) return PeerPush{};
  const int d = c->plan.nd - 1;
  int hm, hp; scheme_halos(c->plan, hm, hp);
  pp.hm = hm; pp.hp = hp;
  pp.np_lo = (int)c->peer_np[0];
  for (int m = 0; m < 5; m++) {
    pp.lo[m] = (c->plan.bc[d][0].kind == BC_EXCHANGE && c->peer_open[0]) ? c->peer_buf[0][set][m] : nullptr;
    pp.hi[m] = (c->plan.bc[d][1].kind == BC_EXCHANGE && c->peer_open[1]) ? c->peer_buf[1][set][m] : nullptr;
  }
  return pp;
}

// kind 0: "I have finished reading my halos", kind 1: "my pushes into your halos have landed".
// Flag slots of a rank: [0] low neighbour read-done, [1] low neighbour pushed, [2] high neighbour read-done, [3] high pushed.
// A rank is the HIGH neighbour of its low neighbour (writes that rank's slots 2,3) and the LOW neighbour of its high one.
void neighbour_signal(osb_ctx *c, int kind) {
  if (!has_exchange(c)) return;
  const int d = c->plan.nd - 1;
  const bool lo = c->plan.bc[d][0].kind == BC_EXCHANGE && c->peer_open[0], hi = c->plan.bc[d][1].kind == BC_EXCHANGE && c->peer_open[1];
  const unsigned long long e = ++c->epoch_sig[kind];
  k_signal<<<1, 1, 0, c->stream>>>(lo ? c->peer_flags[0] + 2 + kind : nullptr, hi ? c->peer_flags[1] + 0 + kind : nullptr, e);
  c->launches++;
}
void neighbour_wait(osb_ctx *c, int kind) {
  if (!has_exchange(c)) return;
  const int d = c->plan.nd - 1;
  const bool lo = c->plan.bc[d][0].kind == BC_EXCHANGE && c->peer_open[0], hi = c->plan.bc[d][1].kind == BC_EXCHANGE && c->peer_open[1];
  const unsigned long long e = ++c->epoch_wait[kind];
  Launcher L(c, OSB_FAM_SYNC);
  k_wait<<<1, 1, 0, c->stream>>>(lo ? c->flags + 0 + kind : nullptr, hi ? c->flags + 2 + kind : nullptr, e, c->flags + 7);
}

int push_planes_memcpy(osb_ctx *c);

bool stage_tma_enabled() {
  static const bool on = getenv("OSB_NO_STAGE_TMA") == nullptr;
  return on;
}

template <int RK, bool FROMQ = false>
void launch_viscous_tiled(osb_ctx *c, double a, double b) {
  const GridDev &g = c->grid;
  if (FROMQ && RK != 0 && stage_tma_enabled()) {          // TMA plane pipeline when the layout allows it (even padded x-extent)
    TmaMaps5 maps;
    bool ok = ((g.h - 3) & 1) == 0;
    for (int m = 0; m < 5 && ok; m++) ok = tma_make_map(g, c->fp.q[m], VTM_BX, VT_HY, 1, maps.m[m]);
    if (ok) {
      auto kern = k_viscous3d_tma<(RK == 0 ? 1 : RK)>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vtm_smem_bytes());
      dim3 bl(VT_X, VT_Y, 1), gr((g.np[0] + VT_X - 1) / VT_X, (g.np[1] + VT_Y - 1) / VT_Y, (g.np[2] + g.zlen - 1) / g.zlen);
      Launcher L(c, OSB_FAM_VISCOUS);
      kern<<<gr, bl, vtm_smem_bytes(), c->stream>>>(g, c->fp, c->pc, a, b, peer_push(c, true), maps);
      return;
    }
  }
  auto kern = k_viscous3d_tiled<RK, FROMQ>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vt_smem_bytes());
  dim3 bl(VT_X, VT_Y, 1), gr((g.np[0] + VT_X - 1) / VT_X, (g.np[1] + VT_Y - 1) / VT_Y, (g.np[2] + g.zlen - 1) / g.zlen);
  Launcher L(c, OSB_FAM_VISCOUS);
  kern<<<gr, bl, vt_smem_bytes(), c->stream>>>(g, c->fp, c->pc, a, b, RK == 0 ? PeerPush{} : peer_push(c, FROMQ));
}

// Stage kernels that write the new state out of place into the Residual buffers: the buffers then exchange roles with
// the q buffers (the field table follows, so "rho" always names the current density).
void swap_q_and_residual(osb_ctx *c) {
  for (int m = 0; m < 5; m++) {
    if (!c->fp.q[m] || !c->fp.R[m]) continue;
    for (auto &f : c->fields) { if (f.dev == c->fp.q[m]) f.dev = c->fp.R[m]; else if (f.dev == c->fp.R[m]) f.dev = c->fp.q[m]; }
    std::swap(c->fp.q[m], c->fp.R[m]);
  }
  c->swap_parity ^= 1;
  c->prim_stale = true;
}

// Shock-capturing 3-D stages on one GPU: the viscous + RK kernel derives (u, T) from q itself, so the constituent-relation
// kernel is not launched at all (the flux sweeps stage their own primitives too).
bool viscous_from_q_ok(const osb_ctx *c) {
  static const bool on = getenv("OSB_NO_VISCOUS_FROM_Q") == nullptr;
  const Plan &P = c->plan;
  if (!on || P.nd != 3 || P.conv == CONV_CENTRAL || !P.viscous || c->general || P.teno_adaptive) return false;
  for (int s = 0; s < 2; s++) if (P.bc[2][s].kind == BC_EXCHANGE && !(c->peer_open[s] && fused_push_enabled())) return false;
  return true;
}

// phase A of a stage: everything that reads the halos of q (constituent relations, sensor, convective terms)
template <int ND>
void launch_phase_a(osb_ctx *c, int stage = -1) {
  const GridDev &g = c->grid;
  if (!(ND == 3 && stage >= 0 && viscous_from_q_ok(c))) launch_prim<ND>(c);
  if (c->plan.teno_adaptive) {
    dim3 b(64, 2, 2);
    Launcher L(c, OSB_FAM_PRIM);
    k_theta<(ND > 1 ? ND : 2)><<<grid3(g.np[0], g.np[1], g.np[2], b), b, 0, c->stream>>>(g, c->fp, c->pc, c->plan.cl, c->gp);
    if (has_exchange(c)) {
      // the interface between the slab's first point and the plane below it takes C_T from the sensor of that plane
      // (the left point of the interface): it belongs to the low neighbour, which stores it here
      const int d = ND - 1;
      const bool lo = c->plan.bc[d][0].kind == BC_EXCHANGE && c->peer_open[0], hi = c->plan.bc[d][1].kind == BC_EXCHANGE && c->peer_open[1];
      const unsigned long long e = ++c->epoch_theta;
      if (hi && c->peer_theta[1])
        cudaMemcpyAsync(c->peer_theta[1] + (long long)(g.h - 1) * g.s[d], c->gp.theta + (long long)(g.h + g.np[d] - 1) * g.s[d],
                        sizeof(double) * g.s[d], cudaMemcpyDeviceToDevice, c->stream);
      k_signal<<<1, 1, 0, c->stream>>>(nullptr, hi ? c->peer_flags[1] + 4 : nullptr, e);
      c->launches++;
      Launcher L2(c, OSB_FAM_SYNC);
      k_wait<<<1, 1, 0, c->stream>>>(lo ? c->flags + 4 : nullptr, nullptr, e, c->flags + 7);
    }
  }
  if (c->plan.conv == CONV_CENTRAL) {
    dim3 b(64, 2, 2);
    {
      Launcher L(c, OSB_FAM_CENTRAL);
      if (c->general || c->plan.central_form != 0) {
        if (c->plan.central_form == 0) k_central_general<ND, 0><<<grid3(g.np[0], g.np[1], g.np[2], b), b, 0, c->stream>>>(g, c->fp, c->pc, c->plan.cl, c->gp);
        else k_central_general<ND, 1><<<grid3(g.np[0], g.np[1], g.np[2], b), b, 0, c->stream>>>(g, c->fp, c->pc, c->plan.cl, c->gp);
      } else {
        k_central<ND><<<grid3(g.np[0], g.np[1], g.np[2], b), b, 0, c->stream>>>(g, c->fp, c->pc);
      }
    }
    neighbour_signal(c, 0);
  } else {
    launch_flux_scheme<ND>(c);
  }
}

// phase B: viscous terms (stencils on u, T only).  stage >= 0: fuse the RK update (and the halo push of a decomposed run)
// of that stage into the kernel where possible; returns true if the RK update was fused.
template <int ND>
// returns 0: Residual complete, RK update still to do; 1: RK update fused; 2: RK update and peer stores of the boundary planes fused
int launch_phase_b(osb_ctx *c, int stage) {
  const GridDev &g = c->grid;
  if (c->plan.viscous) {
    if (ND == 3 && !c->general) {
      if (stage < 0) launch_viscous_tiled<0>(c, 0.0, 0.0);
      else if (viscous_from_q_ok(c)) {
        if (c->plan.rk == RK_LS) launch_viscous_tiled<1, true>(c, c->plan.rk_a[stage], c->plan.rk_b[stage]);
        else launch_viscous_tiled<2, true>(c, c->plan.rk_a[stage], c->plan.rk_b[stage]);
        swap_q_and_residual(c);
      }
      else if (c->plan.rk == RK_LS) launch_viscous_tiled<1>(c, c->plan.rk_a[stage], c->plan.rk_b[stage]);
      else launch_viscous_tiled<2>(c, c->plan.rk_a[stage], c->plan.rk_b[stage]);
      return stage >= 0 ? 2 : 0;
    }
    dim3 b(64, 2, 2);
    Launcher L(c, OSB_FAM_VISCOUS);
    static const bool tiled_general = getenv("OSB_NO_TILED_GENERAL") == nullptr;
    if (c->general && ND == 3 && tiled_general) {
      cudaFuncSetAttribute(k_viscous3d_tiled_general<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vtg_smem_bytes());
      cudaFuncSetAttribute(k_viscous3d_tiled_general<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vtg_smem_bytes());
      cudaFuncSetAttribute(k_viscous3d_tiled_general<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vtg_smem_bytes());
      dim3 bl(VT_X, VT_Y, 1), gr((g.np[0] + VT_X - 1) / VT_X, (g.np[1] + VT_Y - 1) / VT_Y, (g.np[2] + g.zlen - 1) / g.zlen);
      const double ra = stage >= 0 ? c->plan.rk_a[stage] : 0.0, rb = stage >= 0 ? c->plan.rk_b[stage] : 0.0;
      if (stage < 0) k_viscous3d_tiled_general<0><<<gr, bl, vtg_smem_bytes(), c->stream>>>(g, c->fp, c->pc, c->plan.cl, c->gp, ra, rb);
      else if (c->plan.rk == RK_LS) k_viscous3d_tiled_general<1><<<gr, bl, vtg_smem_bytes(), c->stream>>>(g, c->fp, c->pc, c->plan.cl, c->gp, ra, rb);
      else k_viscous3d_tiled_general<2><<<gr, bl, vtg_smem_bytes(), c->stream>>>(g, c->fp, c->pc, c->plan.cl, c->gp, ra, rb);
      return stage >= 0 ? 1 : 0;
    }
    else if (c->general) k_viscous_general<ND><<<grid3(g.np[0], g.np[1], g.np[2], b), b, 0, c->stream>>>(g, c->fp, c->pc, c->plan.cl, c->gp);
    else k_viscous<ND><<<grid3(g.np[0], g.np[1], g.np[2], b), b, 0, c->stream>>>(g, c->fp, c->pc);
  }
  return 0;
}

void box_launch_cfg(const Box &b, int nv, unsigned &blocks) {
  const long long cnt = (long long)b.n[0] * b.n[1] * b.n[2] * nv;
  long long nb = (cnt + 255) / 256;
  if (nb > 148 * 16) nb = 148 * 16;
  if (nb < 1) nb = 1;
  blocks = (unsigned)nb;
}

int run_user_kernels(osb_ctx *c, int when);

// plane kernels: one thread per point of the boundary plane (or of a part of it), each filling its column of halo points
void launch_plane_bc(osb_ctx *c, int kind, int order, int free_mask, int d, int s, const PlaneSpec &ps) {
  const Plan &P = c->plan;
  const GridDev &g = c->grid;
  const int nv = P.nd + 2;
  const long long cnt = (long long)ps.n[0] * ps.n[1] * ps.n[2];
  const unsigned nb = (unsigned)((cnt + 127) / 128);
  Launcher L(c, OSB_FAM_BC);
  switch (kind) {
    case BC_DIRICHLET_FIELD: k_bc_dirichlet_field<<<nb, 128, 0, c->stream>>>(g, c->fp, nv, ps, c->face_table[d][s], c->face_size[d], free_mask); break;
    case BC_EXTRAPOLATION: k_bc_extrapolation<<<nb, 128, 0, c->stream>>>(g, c->fp, nv, ps, order); break;
    case BC_SYMMETRY: k_bc_symmetry<<<nb, 128, 0, c->stream>>>(g, c->fp, nv, ps); break;
    case BC_ZERO_GRADIENT: k_bc_zero_gradient<<<nb, 128, 0, c->stream>>>(g, c->fp, nv, ps); break;
    case BC_INVISCID_WALL: k_bc_inviscid_wall<<<nb, 128, 0, c->stream>>>(g, c->fp, nv, ps); break;
    case BC_PRESSURE_OUTLET: {
      const double bp = P.consts.at("back_pressure");
      if (P.nd == 1) k_bc_pressure_outlet<1><<<nb, 128, 0, c->stream>>>(g, c->fp, c->pc, ps, bp);
      else if (P.nd == 2) k_bc_pressure_outlet<2><<<nb, 128, 0, c->stream>>>(g, c->fp, c->pc, ps, bp);
      else k_bc_pressure_outlet<3><<<nb, 128, 0, c->stream>>>(g, c->fp, c->pc, ps, bp);
      break;
    }
    case BC_INLET_PRESSURE:
      if (P.nd == 1) k_bc_inlet_pressure<1><<<nb, 128, 0, c->stream>>>(g, c->fp, c->pc, ps);
      else if (P.nd == 2) k_bc_inlet_pressure<2><<<nb, 128, 0, c->stream>>>(g, c->fp, c->pc, ps);
      else k_bc_inlet_pressure<3><<<nb, 128, 0, c->stream>>>(g, c->fp, c->pc, ps);
      break;
    case BC_ISOTHERMAL_WALL:
      if (P.nd == 1) k_bc_isothermal_wall<1><<<nb, 128, 0, c->stream>>>(g, c->fp, c->pc, ps);
      else if (P.nd == 2) k_bc_isothermal_wall<2><<<nb, 128, 0, c->stream>>>(g, c->fp, c->pc, ps);
      else k_bc_isothermal_wall<3><<<nb, 128, 0, c->stream>>>(g, c->fp, c->pc, ps);
      break;
    case BC_ADIABATIC_WALL:
      if (P.nd == 1) k_bc_adiabatic_wall<1><<<nb, 128, 0, c->stream>>>(g, c->fp, c->pc, ps);
      else if (P.nd == 2) k_bc_adiabatic_wall<2><<<nb, 128, 0, c->stream>>>(g, c->fp, c->pc, ps);
      else k_bc_adiabatic_wall<3><<<nb, 128, 0, c->stream>>>(g, c->fp, c->pc, ps);
      break;
    default: break;
  }
}

int launch_bcs(osb_ctx *c) {
  const Plan &P = c->plan;
  const GridDev &g = c->grid;
  const int nv = P.nd + 2;
  int hm, hp; scheme_halos(P, hm, hp);
  for (int d = 0; d < P.nd; d++)
    for (int s = 0; s < 2; s++) {
      const BcSpec &b = P.bc[d][s];
      if (b.kind == BC_EXCHANGE) continue;   // filled by the neighbour rank (osb_halo_push)
      if (b.kind == BC_OPEN) continue;       // window face: nothing to fill
      if (b.kind == BC_GENERIC) {            // boundary class without a hand-written kernel: its run-time compiled kernel
        bool found = false;
        for (auto &k : c->user_kernels) found = found || k.when == 100 + 2 * d + s;
        if (!found) return fail(c, "face (" + std::to_string(d) + ", " + std::to_string(s) + ") has a 'generic' boundary condition but no kernel was registered for it");
        if (run_user_kernels(c, 100 + 2 * d + s)) return 1;
        continue;
      }
      Box full;
      for (int e = 0; e < 3; e++) { full.lo[e] = e < P.nd ? -hm : 0; full.n[e] = e < P.nd ? g.np[e] + hm + hp : 1; }
      unsigned blocks;
      if (b.kind == BC_PERIODIC) {
        // periodic.py:42-56: side 0 copies [0,hm) -> [np,np+hm); side 1 copies [np-hm, np-hm+hp) -> [-hm, -hm+hp)
        Box src = full, dst = full;
        if (s == 0) { src.lo[d] = 0; dst.lo[d] = g.np[d]; src.n[d] = dst.n[d] = hm; }
        else { src.lo[d] = g.np[d] - hm; dst.lo[d] = -hm; src.n[d] = dst.n[d] = hp; }
        box_launch_cfg(src, nv, blocks);
        Launcher L(c, OSB_FAM_BC);
        k_copy_box<<<blocks, 256, 0, c->stream>>>(g, c->fp, nv, src, dst);
      } else if (b.kind == BC_DIRICHLET) {
        // dirichlet.py:28-41: boundary plane + halo planes of that side
        Box dst = full;
        if (s == 0) { dst.lo[d] = -hm; dst.n[d] = hm + 1; } else { dst.lo[d] = g.np[d] - 1; dst.n[d] = hp + 1; }
        DirichletState st;
        for (int m = 0; m < 5; m++) st.q[m] = b.q[m];
        box_launch_cfg(dst, nv, blocks);
        Launcher L(c, OSB_FAM_BC);
        k_fill_box<<<blocks, 256, 0, c->stream>>>(g, c->fp, nv, dst, st);
      } else if (b.kind == BC_SPLIT) {
        // SplitBC (bc_core.py:200-217): the parts in the order given, each over its own range of the boundary plane
        for (const BcPart &part : b.parts) {
          if (part.kind == BC_DIRICHLET) {
            Box dst;
            for (int e = 0; e < 3; e++) { dst.lo[e] = e < P.nd ? part.lo[e] : 0; dst.n[e] = e < P.nd ? part.hi[e] - part.lo[e] : 1; }
            if (s == 0) { dst.lo[d] = -hm; dst.n[d] = hm + 1; } else { dst.lo[d] = g.np[d] - 1; dst.n[d] = hp + 1; }
            DirichletState st;
            for (int m = 0; m < 5; m++) st.q[m] = part.q[m];
            box_launch_cfg(dst, nv, blocks);
            Launcher L(c, OSB_FAM_BC);
            k_fill_box<<<blocks, 256, 0, c->stream>>>(g, c->fp, nv, dst, st);
            continue;
          }
          PlaneSpec ps;
          ps.dir = d; ps.side = s; ps.nh = s == 0 ? hm : hp;
          for (int e = 0; e < 3; e++) { ps.lo[e] = e < P.nd ? part.lo[e] : 0; ps.n[e] = e < P.nd ? part.hi[e] - part.lo[e] : 1; }
          ps.lo[d] = s == 0 ? 0 : g.np[d] - 1; ps.n[d] = 1;
          launch_plane_bc(c, part.kind, part.order, 0, d, s, ps);
        }
      } else {
        // plane kernels: boundary plane of (d, s), tangential range incl. the scheme halos
        PlaneSpec ps;
        ps.dir = d; ps.side = s; ps.nh = s == 0 ? hm : hp;
        for (int e = 0; e < 3; e++) { ps.lo[e] = full.lo[e]; ps.n[e] = full.n[e]; }
        ps.lo[d] = s == 0 ? 0 : g.np[d] - 1; ps.n[d] = 1;
        launch_plane_bc(c, b.kind, b.order, b.free_mask, d, s, ps);
      }
    }
  return 0;
}

template <int ND>
void launch_rk(osb_ctx *c, int stage) {
  const GridDev &g = c->grid;
  dim3 b(128, 2, 1);
  Launcher L(c, OSB_FAM_RK);
  if (c->plan.rk == RK_LS)
    k_rk_ls<ND><<<grid3(g.np[0], g.np[1], g.np[2], b), b, 0, c->stream>>>(g, c->fp, c->pc.dt, c->plan.rk_a[stage], c->plan.rk_b[stage]);
  else
    k_rk_sbli<ND><<<grid3(g.np[0], g.np[1], g.np[2], b), b, 0, c->stream>>>(g, c->fp, c->pc.dt, c->plan.rk_a[stage], c->plan.rk_b[stage]);
}

template <int ND>
void launch_save(osb_ctx *c) {
  const GridDev &g = c->grid;
  dim3 b(128, 2, 1);
  Launcher L(c, OSB_FAM_RK);
  k_rk_save<ND><<<grid3(g.np[0], g.np[1], g.np[2], b), b, 0, c->stream>>>(g, c->fp);
}

template <int ND>
void launch_residual(osb_ctx *c) {
  launch_phase_a<ND>(c);
  launch_phase_b<ND>(c, -1);
}

// Central(4) + viscous + RK in one out-of-place kernel (3-D, constant viscosity, uniform grid, single GPU): the new state is
// written into the Residual buffers, which then exchange roles with the q buffers (the field table follows).
bool fused_central_ok(const osb_ctx *c) {
  static const bool on = getenv("OSB_NO_FUSED_CENTRAL") == nullptr;
  const Plan &P = c->plan;
  if (!on || P.nd != 3 || P.conv != CONV_CENTRAL || !P.viscous || c->general || P.central_form != 0) return false;
  for (int s = 0; s < 2; s++) if (P.bc[2][s].kind == BC_EXCHANGE && !(c->peer_open[s] && fused_push_enabled())) return false;
  return true;
}

void launch_central_fused(osb_ctx *c, int stage) {
  const GridDev &g = c->grid;
  QPtrs qi, qo, rk;
  for (int m = 0; m < 5; m++) { qi.q[m] = c->fp.q[m]; qo.q[m] = c->fp.R[m]; rk.q[m] = c->fp.rk[m]; }
  const bool push = has_exchange(c);
  TmaMaps5 maps;
  bool tma = stage_tma_enabled() && ((g.h - 3) & 1) == 0;
  for (int m = 0; m < 5 && tma; m++) tma = tma_make_map(g, c->fp.q[m], VTM_BX, VT_HY, 1, maps.m[m]);
  if (!tma) memset(&maps, 0, sizeof(maps));
  auto launch = [&](auto kern, int first) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ct_smem_bytes(tma));
    dim3 bl(VT_X, VT_Y, 1), gr((g.np[0] + VT_X - 1) / VT_X, (g.np[1] + VT_Y - 1) / VT_Y, (g.np[2] + g.zlen - 1) / g.zlen);
    Launcher L(c, OSB_FAM_CENTRAL);
    kern<<<gr, bl, ct_smem_bytes(tma), c->stream>>>(g, qi, qo, rk, c->pc, c->plan.rk_a[stage], c->plan.rk_b[stage], first, push ? peer_push(c, true) : PeerPush{}, maps);
  };
  const int first = (c->plan.rk == RK_LS) ? 0 : (stage == 0);
  if (c->plan.rk == RK_LS) {
    if (push) { if (tma) launch(k_central3d_fused<1, true, true>, first); else launch(k_central3d_fused<1, true, false>, first); }
    else { if (tma) launch(k_central3d_fused<1, false, true>, first); else launch(k_central3d_fused<1, false, false>, first); }
  } else {
    if (push) { if (tma) launch(k_central3d_fused<2, true, true>, first); else launch(k_central3d_fused<2, true, false>, first); }
    else { if (tma) launch(k_central3d_fused<2, false, true>, first); else launch(k_central3d_fused<2, false, false>, first); }
  }
  swap_q_and_residual(c);
}

// ---- run-time compiled point-wise user kernels ---------------------------------------------------
struct UserFields { double *p[OSB_MAX_USER_FIELDS]; long long iter; };     // iter: the loop counter the reference hands to kernels as ops_arg_gbl

bool is_primitive(const osb_ctx *c, const double *dev) {
  bool prim = dev == c->fp.p || dev == c->fp.a || dev == c->fp.T;
  for (int d = 0; d < c->plan.nd; d++) prim = prim || dev == c->fp.u[d];
  return prim;
}
void launch_prim_nd(osb_ctx *c) {
  switch (c->plan.nd) { case 1: launch_prim<1>(c); break; case 2: launch_prim<2>(c); break; default: launch_prim<3>(c); }
}

int run_user_kernels(osb_ctx *c, int when) {
  const GridDev &g = c->grid;
  for (auto &k : c->user_kernels) {
    if (k.when != when && !(when >= 210 && when < 220 && k.when == 209)) continue;      // 209: every stage
    // u_i, p, a, T are not written by the stage kernels that derive them on the fly: a user kernel that reads one of them
    // sees the constituent relations of the current state (what the reference's loop reads after its last stage)
    if (c->prim_stale && c->plan.conv != CONV_GENERIC)
      for (auto &n : k.fields) { Field *f = find_field(c, n.c_str()); if (f && is_primitive(c, f->dev)) { launch_prim_nd(c); break; } }
    UserFields uf{};
    uf.iter = c->iteration;
    for (size_t i = 0; i < k.fields.size(); i++) {
      Field *f = find_field(c, k.fields[i].c_str());
      if (!f) return fail(c, "user kernel field vanished: " + k.fields[i]);
      uf.p[i] = f->dev;                      // looked up per launch: the q / Residual buffers exchange roles
    }
    int lo[3] = {0, 0, 0}, n[3] = {1, 1, 1};
    for (int d = 0; d < g.nd; d++) { lo[d] = k.range[2 * d]; n[d] = k.range[2 * d + 1] - k.range[2 * d]; }
    long long off = g.off, s1 = g.nd > 1 ? g.s[1] : 0, s2 = g.nd > 2 ? g.s[2] : 0;
    void *args[] = {&off, &n[0], &n[1], &n[2], &lo[0], &lo[1], &lo[2], &s1, &s2, &uf};
    dim3 bl(128, 1, 1), gr((n[0] + 127) / 128, n[1], n[2]);
    Launcher L(c, OSB_FAM_USER);
    OSB_CUDA(c, cudaLaunchKernel((const void *)k.kern, gr, bl, args, 0, c->stream));
    if (k.writes_state && c->plan.conv != CONV_GENERIC) c->prim_stale = true;
  }
  return 0;
}

// One stage of the loop (s < 0: iteration start).  In a decomposed run the neighbour exchange is part of the stage and is
// ordered on the stream by flag words, so a whole run can be enqueued without host synchronisation:
//   phase A (reads halos) -> "read done" handshake -> phase B + RK (+ fused peer push) -> "pushed" handshake -> local BCs
// CONV_GENERIC: a stage is the list of run-time compiled kernels the back end printed for it, boundary kernels and periodic copies
// included, in program order (algorithm.py:440-474)
int generic_stage(osb_ctx *c, int s) {
  const int last = c->plan.generic_stages - 1;
  if (s < 0) return run_user_kernels(c, 200);
  if (run_user_kernels(c, 210 + s)) return 1;
  if (s == last) { if (run_user_kernels(c, 0)) return 1; c->iteration++; }
  OSB_CUDA(c, cudaGetLastError());
  return 0;
}

template <int ND>
int stage_nd(osb_ctx *c, int s) {
  if (c->plan.conv == CONV_GENERIC) return generic_stage(c, s);
  const bool ex = has_exchange(c);
  if (c->plan.mass_source && s <= 0) c->pc.src_factor = sin(c->plan.src_rate * (double)c->iteration);
  struct StepCounter { osb_ctx *c; int s; ~StepCounter() { if (s == (int)c->plan.rk_a.size() - 1) c->iteration++; } } counter{c, s};
  if (ND == 3 && fused_central_ok(c)) {
    if (s >= 0) {
      launch_central_fused(c, s);                       // the kernel also stores the boundary planes into the neighbours' buffers
      if (ex) { neighbour_signal(c, 1); neighbour_wait(c, 1); }
      if (launch_bcs(c)) return 1;
    } else {
      // planes first, rank-local BCs after the neighbours' planes have landed: the BC kernels also rewrite the x/y-halo
      // parts of the received planes, which must neither race with nor precede the neighbour's stores
      if (ex) { if (push_planes_memcpy(c)) return 1; neighbour_signal(c, 1); neighbour_wait(c, 1); }
      if (launch_bcs(c)) return 1;
    }
    if (s == (int)c->plan.rk_a.size() - 1 && run_user_kernels(c, 0)) return 1;
    OSB_CUDA(c, cudaGetLastError());
    return 0;
  }
  if (s < 0) {
    if (ex) { if (push_planes_memcpy(c)) return 1; neighbour_signal(c, 1); neighbour_wait(c, 1); }
    if (launch_bcs(c)) return 1;
    if (c->plan.rk == RK_SBLI) launch_save<ND>(c);
  } else {
    launch_phase_a<ND>(c, s);     // sends the "read done" notification as soon as the halo-reading kernels are enqueued
    neighbour_wait(c, 0);
    const int fused = launch_phase_b<ND>(c, s);
    if (!fused) launch_rk<ND>(c, s);
    if (ex && fused == 2 && fused_push_enabled()) {            // planes were pushed by the kernel: interior part; the receiver's own BCs complete the halos
      neighbour_signal(c, 1); neighbour_wait(c, 1);
      if (launch_bcs(c)) return 1;
    } else {
      if (ex) { if (push_planes_memcpy(c)) return 1; neighbour_signal(c, 1); neighbour_wait(c, 1); }
      if (launch_bcs(c)) return 1;
    }
  }
  if (s == (int)c->plan.rk_a.size() - 1 && run_user_kernels(c, 0)) return 1;
  OSB_CUDA(c, cudaGetLastError());
  return 0;
}

template <int ND>
int step_nd(osb_ctx *c, int nsteps) {
  const int nstages = (int)c->plan.rk_a.size();
  for (int it = 0; it < nsteps; it++) {
    if (stage_nd<ND>(c, -1)) return 1;
    for (int s = 0; s < nstages; s++) if (stage_nd<ND>(c, s)) return 1;
  }
  return 0;
}
int do_stage(osb_ctx *c, int s) {
  if (s >= (int)c->plan.rk_a.size()) return fail(c, "stage index out of range");
  switch (c->plan.nd) {
    case 1: return stage_nd<1>(c, s);
    case 2: return stage_nd<2>(c, s);
    default: return stage_nd<3>(c, s);
  }
}

int step_dispatch(osb_ctx *c, int nsteps) {
  switch (c->plan.nd) {
    case 1: return step_nd<1>(c, nsteps);
    case 2: return step_nd<2>(c, nsteps);
    default: return step_nd<3>(c, nsteps);
  }
}

void drop_graph(osb_ctx *c) {
  if (c->step_graph) { cudaGraphExecDestroy(c->step_graph); c->step_graph = nullptr; }
}

// Small grids are launch-bound (a Sod step is ~20 kernels of a few microseconds): one time step is captured once as a
// CUDA graph and replayed.  Not used while profiling (events between launches) nor in decomposed runs (the handshake
// kernels carry a new epoch number every stage).
int do_step(osb_ctx *c, int nsteps) {
  const long long pts = (long long)c->grid.np[0] * c->grid.np[1] * c->grid.np[2];
  bool iter_kernels = false;
  for (auto &k : c->user_kernels) iter_kernels = iter_kernels || k.uses_iter;
  if (!c->use_graph || c->profiling || c->plan.mass_source || iter_kernels || has_exchange(c) || pts > (1LL << 22) || nsteps < 2) return step_dispatch(c, nsteps);
  // the fused central path exchanges buffer roles every stage: the captured unit must bring them back (two steps if the
  // number of stages is odd) and may only be replayed from the parity it was captured at (0)
  const int unit = ((fused_central_ok(c) || viscous_from_q_ok(c)) && (c->plan.rk_a.size() % 2)) ? 2 : 1;
  if (c->swap_parity) { if (step_dispatch(c, 1)) return 1; nsteps--; }
  if (nsteps < unit) return step_dispatch(c, nsteps);
  if (!c->step_graph) {
    cudaGraph_t graph = nullptr;
    const long long l0 = c->launches, it0 = c->iteration;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return step_dispatch(c, nsteps); }
    const int rc = step_dispatch(c, unit);
    const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    c->graph_launches = c->launches - l0;
    c->launches = l0;
    c->iteration = it0;              // nothing has run yet: the capture only recorded the launches
    if (rc || e != cudaSuccess || !graph) { cudaGetLastError(); if (graph) cudaGraphDestroy(graph); c->use_graph = false; return step_dispatch(c, nsteps); }
    if (cudaGraphInstantiate(&c->step_graph, graph, 0) != cudaSuccess) { cudaGetLastError(); cudaGraphDestroy(graph); c->use_graph = false; return step_dispatch(c, nsteps); }
    cudaGraphDestroy(graph);
  }
  for (int it = 0; it + unit <= nsteps; it += unit) {
    OSB_CUDA(c, cudaGraphLaunch(c->step_graph, c->stream));
    c->launches += c->graph_launches;
    // host-side effects of the replayed unit: the loop counter advances, and on the paths whose stage kernels derive the
    // primitives on the fly the u/p/a/T arrays now lag the state again
    c->iteration += unit;
    if (fused_central_ok(c) || viscous_from_q_ok(c)) c->prim_stale = true;
  }
  return step_dispatch(c, nsteps % unit);
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char *osb_last_error(const osb_ctx *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int osb_create(const char *plan_text, int device, osb_ctx **out) {
  if (!plan_text || !out) { g_create_error = "null argument"; return 1; }
  *out = nullptr;
  osb_ctx *c = new osb_ctx();
  std::string err;
  if (!parse_plan(plan_text, c->plan, err)) { g_create_error = "plan: " + err; delete c; return 1; }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(ce) + "): the B200 back end has no CPU fallback";
    delete c; return 2;
  }
  if (device >= 0) { ce = cudaSetDevice(device); if (ce != cudaSuccess) { g_create_error = cudaGetErrorString(ce); delete c; return 2; } }
  cudaGetDevice(&c->device);
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { g_create_error = "cudaStreamCreate failed"; delete c; return 2; }
  const Plan &P = c->plan;
  GridDev &g = c->grid;
  g.nd = P.nd; g.h = 5;
  long long n = 1; g.off = 0;
  for (int d = 0; d < 3; d++) {
    g.np[d] = d < P.nd ? P.np[d] : 1;
    g.pd[d] = d < P.nd ? P.np[d] + 2 * g.h : 1;
    g.s[d] = n;
    if (d < P.nd) g.off += g.h * n;
    n *= g.pd[d];
  }
  g.n = n;
  // z-marching kernels: the longest column chunk that still gives every SM a few blocks
  {
    const long long columns = (long long)((g.np[0] + VT_X - 1) / VT_X) * ((g.np[1] + VT_Y - 1) / VT_Y);
    g.zlen = 8;
    for (int z : {32, 16}) if (columns * ((g.np[2] + z - 1) / z) >= 148 * 4) { g.zlen = z; break; }
  }
  refresh_constants(c);
  // fields (names = reference dataset names without the _B0 suffix)
  const int nv = P.nd + 2;
  std::vector<std::string> qn = {"rho"};
  for (int d = 0; d < P.nd; d++) qn.push_back("rhou" + std::to_string(d));
  qn.push_back("rhoE");
  auto add = [&](const std::string &name, double **slot) -> bool {
    Field f; f.name = name;
    if (cudaMalloc(&f.dev, sizeof(double) * g.n) != cudaSuccess) return false;
    cudaMemsetAsync(f.dev, 0, sizeof(double) * g.n, c->stream);   // OPS semantics: zero-initialised dats
    *slot = f.dev;
    c->fields.push_back(f);
    return true;
  };
  bool ok = true;
  for (int m = 0; m < nv && ok; m++) ok = add(qn[m], &c->fp.q[m]);
  for (int d = 0; d < P.nd && ok; d++) ok = add("u" + std::to_string(d), &c->fp.u[d]);
  ok = ok && add("p", &c->fp.p) && add("a", &c->fp.a) && add("T", &c->fp.T);
  for (int m = 0; m < nv && ok; m++) ok = add("Residual" + std::to_string(m), &c->fp.R[m]);
  for (int m = 0; m < nv && ok; m++) ok = add(P.rk == RK_LS ? "tempRK_" + qn[m] : qn[m] + "_RKold", &c->fp.rk[m]);
  for (int m = 0; m < nv; m++) { c->own_buf[0][m] = c->fp.q[m]; c->own_buf[1][m] = c->fp.R[m]; }
  // general path: metric fields (uploaded by the caller), viscosity, sensor
  c->general = P.visc_law != 0 || P.forcing || P.mass_source;
  c->iteration = P.iteration0;
  if (P.mass_source && !P.viscous) { g_create_error = "mass_source is implemented on the viscous general path only"; osb_destroy(c); return 1; }
  for (int d = 0; d < P.nd; d++) {
    if (P.metric[d] || P.bc[d][0].closure || P.bc[d][1].closure) c->general = true;
    if (P.metric[d] && ok) {
      double *pD = nullptr, *pS = nullptr;
      const std::string dd = std::to_string(d);
      ok = add("D" + dd + dd, &pD) && add("SD" + dd + dd + dd, &pS);
      c->gp.D[d] = pD; c->gp.SD[d] = pS;
    }
  }
  if (ok && (c->general || P.visc_law != 0)) ok = add("mu", &c->gp.mu);
  if (P.curvilinear) {
    if (P.nd != 2 || P.viscous || P.conv == CONV_CENTRAL || P.teno_adaptive) {
      g_create_error = "curvilinear grids are implemented for 2-D inviscid shock-capturing schemes"; osb_destroy(c); return 1;
    }
    for (int i = 0; i < 2 && ok; i++)
      for (int j = 0; j < 2 && ok; j++) { double *pd = nullptr; ok = add("D" + std::to_string(i) + std::to_string(j), &pd); c->cp.D[i][j] = pd; }
    if (ok) { double *pd = nullptr; ok = add("detJ", &pd); c->cp.detJ = pd; }
    for (int m = 0; m < 4 && ok; m++) ok = add("wk_flux" + std::to_string(m), &c->cp.wk[m]);
  }
  if (ok && P.mass_source) { double *ps = nullptr; ok = add("BF_amp", &ps); c->gp.src = ps; }
  if (ok && P.teno_adaptive) ok = add("theta", &c->gp.theta) && add("TENO", &c->gp.teno_store);
  if (P.teno_adaptive && P.nd < 2) { g_create_error = "adaptive TENO needs at least 2 dimensions (vorticity)"; osb_destroy(c); return 1; }
  for (int d = 0; d < P.nd && ok; d++) {
    long long ts = 1;
    for (int e = 0; e < P.nd; e++) if (e != d) ts *= g.pd[e];
    c->face_size[d] = ts;
    for (int s = 0; s < 2 && ok; s++)
      if (P.bc[d][s].kind == BC_DIRICHLET_FIELD) {
        ok = cudaMalloc(&c->face_table[d][s], sizeof(double) * ts * nv) == cudaSuccess;
        if (ok) cudaMemsetAsync(c->face_table[d][s], 0, sizeof(double) * ts * nv, c->stream);
      }
  }
  if (!ok) { g_create_error = "cudaMalloc failed (out of device memory)"; osb_destroy(c); return 3; }
  if (cudaMalloc(&c->flags, 8 * sizeof(unsigned long long)) != cudaSuccess) { g_create_error = "cudaMalloc failed"; osb_destroy(c); return 3; }
  cudaMemsetAsync(c->flags, 0, 8 * sizeof(unsigned long long), c->stream);
  cudaStreamSynchronize(c->stream);
  *out = c;
  return 0;
}

int osb_destroy(osb_ctx *c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  for (int s = 0; s < 2; s++)
    if (c->peer_open[s]) {
      for (int t = 0; t < 2; t++) for (int m = 0; m < 5; m++) if (c->peer_buf[s][t][m]) cudaIpcCloseMemHandle(c->peer_buf[s][t][m]);
      if (c->peer_flags[s]) cudaIpcCloseMemHandle(c->peer_flags[s]);
      if (c->peer_theta[s]) cudaIpcCloseMemHandle(c->peer_theta[s]);
    }
  drop_graph(c);
  for (auto &k : c->user_kernels) if (k.lib) cudaLibraryUnload(k.lib);
  if (c->flags) cudaFree(c->flags);
  if (c->diag_buf) cudaFree(c->diag_buf);
  if (c->slow_count) cudaFree(c->slow_count);
  for (auto &f : c->fields) cudaFree(f.dev);
  for (int d = 0; d < 3; d++) for (int s = 0; s < 2; s++) if (c->face_table[d][s]) cudaFree(c->face_table[d][s]);
  if (c->timer0) { cudaEventDestroy(c->timer0); cudaEventDestroy(c->timer1); }
  if (c->s_up) cudaStreamDestroy(c->s_up);
  if (c->s_down) cudaStreamDestroy(c->s_down);
  for (cudaEvent_t e : {c->ev_up, c->ev_comp, c->ev_down}) if (e) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int osb_set_const_f64(osb_ctx *c, const char *name, double v) {
  if (!c || !name) return 1;
  c->plan.consts[name] = v;
  refresh_constants(c);
  drop_graph(c);          // constants are baked into the captured kernel arguments
  return 0;
}
int osb_create_field(osb_ctx *c, const char *name) {
  if (!c || !name || !*name) return 1;
  if (find_field(c, name)) return 0;
  cudaSetDevice(c->device);
  Field f; f.name = name;
  if (f.name.size() > 3 && f.name.compare(f.name.size() - 3, 3, "_B0") == 0) f.name.resize(f.name.size() - 3);
  OSB_CUDA(c, cudaMalloc(&f.dev, sizeof(double) * c->grid.n));
  OSB_CUDA(c, cudaMemsetAsync(f.dev, 0, sizeof(double) * c->grid.n, c->stream));
  c->fields.push_back(f);
  return 0;
}
int osb_add_user_kernel(osb_ctx *c, const char *source, const char *entry, const char *fields, const int range[6], int when) {
  if (!c || !source || !entry || !fields || !range) return 1;
  cudaSetDevice(c->device);
  osb_ctx::UserKernel k;
  k.when = when;
  k.uses_iter = strstr(source, "f.iter") != nullptr;           // reads the loop counter: its launches cannot be replayed from a captured graph
  for (int i = 0; i < 6; i++) k.range[i] = range[i];
  std::vector<bool> written;
  {
    std::stringstream ss(fields);
    std::string n;
    while (std::getline(ss, n, ',')) if (!n.empty()) {
      const bool w = n[0] == '+';
      k.fields.push_back(w ? n.substr(1) : n);
      written.push_back(w);
    }
  }
  // a kernel that writes a conserved array (filters: SFD, the WENO filter) leaves the primitive arrays behind the state
  for (size_t i = 0; i < k.fields.size(); i++)
    if (written[i]) { Field *f = find_field(c, k.fields[i].c_str()); for (int m = 0; f && m < c->plan.nd + 2; m++) k.writes_state = k.writes_state || f->dev == c->fp.q[m] || f->dev == c->fp.R[m]; }
  if (k.fields.size() > OSB_MAX_USER_FIELDS) return fail(c, "user kernel uses too many arrays");
  for (size_t i = 0; i < k.fields.size(); i++)
    if (!find_field(c, k.fields[i].c_str())) {
      // a dataset the kernel writes (e.g. a running mean) is created zero-initialised, as OPS declares datasets; one it only
      // reads must exist: reading zeros in place of a dataset nobody uploaded would be a silent wrong answer
      if (!written[i]) return fail(c, "user kernel reads dataset '" + k.fields[i] + "' which does not exist on the device: create it with osb_create_field and upload it first");
      if (osb_create_field(c, k.fields[i].c_str())) return 1;
    }
  // NVRTC is loaded on demand: the library itself must not depend on it (it also loads on hosts without a driver)
  typedef int (*create_t)(void **, const char *, const char *, int, const char *const *, const char *const *);
  typedef int (*compile_t)(void *, int, const char *const *);
  typedef int (*size_t_fn)(void *, size_t *);
  typedef int (*get_t)(void *, char *);
  typedef int (*destroy_t)(void **);
  void *h = dlopen("libnvrtc.so.12", RTLD_NOW);
  if (!h) h = dlopen("libnvrtc.so", RTLD_NOW);
  if (!h) h = dlopen("/usr/local/cuda/lib64/libnvrtc.so", RTLD_NOW);
  if (!h) return fail(c, std::string("osb_add_user_kernel: cannot load NVRTC: ") + dlerror());
  auto create = (create_t)dlsym(h, "nvrtcCreateProgram");
  auto compile = (compile_t)dlsym(h, "nvrtcCompileProgram");
  auto log_size = (size_t_fn)dlsym(h, "nvrtcGetProgramLogSize");
  auto get_log = (get_t)dlsym(h, "nvrtcGetProgramLog");
  auto bin_size = (size_t_fn)dlsym(h, "nvrtcGetCUBINSize");
  auto get_bin = (get_t)dlsym(h, "nvrtcGetCUBIN");
  auto destroy = (destroy_t)dlsym(h, "nvrtcDestroyProgram");
  if (!create || !compile || !log_size || !get_log || !bin_size || !get_bin || !destroy) return fail(c, "osb_add_user_kernel: NVRTC symbols missing");
  void *prog = nullptr;
  if (create(&prog, source, "osb_user_kernel.cu", 0, nullptr, nullptr)) return fail(c, "nvrtcCreateProgram failed");
  const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false"};
  const int rc = compile(prog, 3, opts);
  if (rc) {
    size_t ls = 0; log_size(prog, &ls);
    std::string log(ls, ' ');
    if (ls) get_log(prog, &log[0]);
    destroy(&prog);
    return fail(c, "user kernel does not compile:\n" + log);
  }
  size_t bs = 0; bin_size(prog, &bs);
  std::vector<char> bin(bs);
  get_bin(prog, bin.data());
  destroy(&prog);
  OSB_CUDA(c, cudaLibraryLoadData(&k.lib, bin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  OSB_CUDA(c, cudaLibraryGetKernel(&k.kern, k.lib, entry));
  c->user_kernels.push_back(k);
  drop_graph(c);
  return 0;
}
int osb_run_user_kernels(osb_ctx *c, int when) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  return run_user_kernels(c, when);
}

int osb_set_iteration(osb_ctx *c, long long iteration) {
  if (!c || iteration < 0) return 1;
  c->iteration = iteration;
  refresh_constants(c);
  return 0;
}
long long osb_get_iteration(const osb_ctx *c) { return c ? c->iteration : -1; }
int osb_get_const_f64(const osb_ctx *c, const char *name, double *v) {
  if (!c || !name || !v) return 1;
  auto it = c->plan.consts.find(name);
  if (it == c->plan.consts.end()) return 1;
  *v = it->second;
  return 0;
}

int osb_num_fields(const osb_ctx *c) { return c ? (int)c->fields.size() : 0; }
const char *osb_field_name(const osb_ctx *c, int i) { return (c && i >= 0 && i < (int)c->fields.size()) ? c->fields[i].name.c_str() : nullptr; }

int osb_field_info(const osb_ctx *c, const char *name, int *dims, int *halo_m, int *halo_p) {
  if (!c || !name) return 1;
  if (!find_field(const_cast<osb_ctx *>(c), name)) return fail(const_cast<osb_ctx *>(c), std::string("unknown field ") + name);
  for (int d = 0; d < 3; d++) {
    if (dims) dims[d] = c->grid.np[d];
    if (halo_m) halo_m[d] = d < c->grid.nd ? -c->grid.h : 0;
    if (halo_p) halo_p[d] = d < c->grid.nd ? c->grid.h : 0;
  }
  return 0;
}

int osb_upload(osb_ctx *c, const char *name, const double *host) {
  if (!c || !name || !host) return 1;
  Field *f = find_field(c, name);
  if (!f) return fail(c, std::string("unknown field ") + name);
  OSB_CUDA(c, cudaMemcpyAsync(f->dev, host, sizeof(double) * c->grid.n, cudaMemcpyHostToDevice, c->stream));
  OSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}
// u_i, p, a, T are not kept up to date by the stage kernels that derive them on the fly: evaluate the constituent relations
// of the current state before handing such an array out
static void refresh_primitives_for(osb_ctx *c, const Field *f) {
  if (!c->prim_stale || c->plan.conv == CONV_GENERIC || !is_primitive(c, f->dev)) return;
  cudaSetDevice(c->device);
  launch_prim_nd(c);
}

int osb_download(osb_ctx *c, const char *name, double *host) {
  if (!c || !name || !host) return 1;
  Field *f = find_field(c, name);
  if (!f) return fail(c, std::string("unknown field ") + name);
  refresh_primitives_for(c, f);
  OSB_CUDA(c, cudaMemcpyAsync(host, f->dev, sizeof(double) * c->grid.n, cudaMemcpyDeviceToHost, c->stream));
  OSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}
int osb_read_point(osb_ctx *c, const char *name, int i, int j, int k, double *value) {
  if (!c || !name || !value) return 1;
  Field *f = find_field(c, name);
  if (!f) return fail(c, std::string("unknown field ") + name);
  const GridDev &g = c->grid;
  const int id[3] = {i, j, k};
  for (int d = 0; d < g.nd; d++) if (id[d] < -g.h || id[d] >= g.np[d] + g.h) return fail(c, "osb_read_point: index outside the padded array");
  refresh_primitives_for(c, f);
  const long long x = g.off + i + (g.nd > 1 ? j * g.s[1] : 0) + (g.nd > 2 ? k * g.s[2] : 0);
  OSB_CUDA(c, cudaMemcpyAsync(value, f->dev + x, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  OSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}
int osb_upload_face(osb_ctx *c, int dir, int side, const double *table) {
  if (!c || !table || dir < 0 || dir >= c->plan.nd || side < 0 || side > 1) return 1;
  if (!c->face_table[dir][side]) return fail(c, "osb_upload_face: this face has no dirichlet_field boundary condition");
  const size_t bytes = sizeof(double) * c->face_size[dir] * (c->plan.nd + 2);
  OSB_CUDA(c, cudaMemcpyAsync(c->face_table[dir][side], table, bytes, cudaMemcpyHostToDevice, c->stream));
  OSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}
int osb_device_ptr(osb_ctx *c, const char *name, double **p) {
  if (!c || !name || !p) return 1;
  Field *f = find_field(c, name);
  if (!f) return fail(c, std::string("unknown field ") + name);
  refresh_primitives_for(c, f);
  *p = f->dev;
  return 0;
}

int osb_step(osb_ctx *c, int nsteps) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  return do_step(c, nsteps);
}
int osb_step_begin(osb_ctx *c) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  return do_stage(c, -1);
}
int osb_stage(osb_ctx *c, int stage) {
  if (!c || stage < 0) return 1;
  cudaSetDevice(c->device);
  return do_stage(c, stage);
}
int osb_sync(osb_ctx *c) {
  if (!c) return 1;
  OSB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->epoch_wait[0] || c->epoch_wait[1]) {
    unsigned long long err = 0;
    OSB_CUDA(c, cudaMemcpy(&err, c->flags + 7, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) return fail(c, "timed out waiting for a neighbour rank (halo exchange epoch " + std::to_string(err) + ")");
  }
  return 0;
}
int osb_apply_bcs(osb_ctx *c) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  if (c->plan.conv == CONV_GENERIC) return run_user_kernels(c, 200);     // generic path: the iteration-start list IS the boundary pass
  if (launch_bcs(c)) return 1;
  OSB_CUDA(c, cudaGetLastError());
  return 0;
}
int osb_residual(osb_ctx *c) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  if (c->plan.conv == CONV_GENERIC) return fail(c, "osb_residual: a program on the generic path has no separate residual evaluation (its stage lists include the RK update)");
  switch (c->plan.nd) { case 1: launch_residual<1>(c); break; case 2: launch_residual<2>(c); break; default: launch_residual<3>(c); }
  OSB_CUDA(c, cudaGetLastError());
  return 0;
}

int osb_step_timed(osb_ctx *c, int nsteps, double *ms) {
  if (!c || !ms) return 1;
  cudaSetDevice(c->device);
  cudaEvent_t e0, e1;
  OSB_CUDA(c, cudaEventCreate(&e0)); OSB_CUDA(c, cudaEventCreate(&e1));
  OSB_CUDA(c, cudaStreamSynchronize(c->stream));
  OSB_CUDA(c, cudaEventRecord(e0, c->stream));
  int rc = do_step(c, nsteps);
  OSB_CUDA(c, cudaEventRecord(e1, c->stream));
  OSB_CUDA(c, cudaEventSynchronize(e1));
  float t = 0; cudaEventElapsedTime(&t, e0, e1);
  *ms = t;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return rc;
}

int osb_timer_start(osb_ctx *c) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  if (!c->timer0) { OSB_CUDA(c, cudaEventCreate(&c->timer0)); OSB_CUDA(c, cudaEventCreate(&c->timer1)); }
  OSB_CUDA(c, cudaEventRecord(c->timer0, c->stream));
  return 0;
}
int osb_timer_stop(osb_ctx *c, double *ms) {
  if (!c || !ms || !c->timer0) return 1;
  cudaSetDevice(c->device);
  OSB_CUDA(c, cudaEventRecord(c->timer1, c->stream));
  OSB_CUDA(c, cudaEventSynchronize(c->timer1));
  float t = 0; OSB_CUDA(c, cudaEventElapsedTime(&t, c->timer0, c->timer1));
  *ms = t;
  return 0;
}

int osb_advance_host(osb_ctx *c, const double *const *q_in, double *const *q_out, int nsteps, double *ms) {
  if (!c || !q_in || !q_out) return 1;
  cudaSetDevice(c->device);
  const int nv = c->plan.nd + 2;
  const size_t bytes = sizeof(double) * c->grid.n;
  cudaEvent_t e0, e1;
  OSB_CUDA(c, cudaEventCreate(&e0)); OSB_CUDA(c, cudaEventCreate(&e1));
  OSB_CUDA(c, cudaStreamSynchronize(c->stream));
  OSB_CUDA(c, cudaEventRecord(e0, c->stream));
  for (int m = 0; m < nv; m++) OSB_CUDA(c, cudaMemcpyAsync(c->fp.q[m], q_in[m], bytes, cudaMemcpyHostToDevice, c->stream));
  int rc = do_step(c, nsteps);
  for (int m = 0; m < nv; m++) OSB_CUDA(c, cudaMemcpyAsync(q_out[m], c->fp.q[m], bytes, cudaMemcpyDeviceToHost, c->stream));
  OSB_CUDA(c, cudaEventRecord(e1, c->stream));
  OSB_CUDA(c, cudaEventSynchronize(e1));
  float t = 0; cudaEventElapsedTime(&t, e0, e1);
  if (ms) *ms = t;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return rc;
}

// ---- window pipeline over a host-resident state ------------------------------------------------------
// A block that lives in host memory is advanced window by window (planes along the slowest axis, 'open' faces, guard planes
// on both sides): while one context sweeps its window, the next one receives its planes and the previous one returns its
// result, so the three engines (H2D, SMs, D2H) work at the same time.  Each context orders its own three streams by events.
static int pipeline_init(osb_ctx *c) {
  if (c->s_up) return 0;
  OSB_CUDA(c, cudaStreamCreateWithFlags(&c->s_up, cudaStreamNonBlocking));
  OSB_CUDA(c, cudaStreamCreateWithFlags(&c->s_down, cudaStreamNonBlocking));
  OSB_CUDA(c, cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming));
  OSB_CUDA(c, cudaEventCreateWithFlags(&c->ev_comp, cudaEventDisableTiming));
  OSB_CUDA(c, cudaEventCreateWithFlags(&c->ev_down, cudaEventDisableTiming));
  return 0;
}
static int plane_range_ok(osb_ctx *c, int plane0, int nplanes) {
  const int d = c->plan.nd - 1;
  if (plane0 < 0 || nplanes < 1 || plane0 + nplanes > c->grid.pd[d]) return fail(c, "plane range outside the padded array");
  return 0;
}
int osb_host_planes_upload(osb_ctx *c, const double *const *src, int plane0, int nplanes) {
  if (!c || !src) return 1;
  cudaSetDevice(c->device);
  if (pipeline_init(c) || plane_range_ok(c, plane0, nplanes)) return 1;
  const int d = c->plan.nd - 1, nv = c->plan.nd + 2;
  // the planes being overwritten may still be on their way back to the host from the previous window of this context
  if (c->down_pending) OSB_CUDA(c, cudaStreamWaitEvent(c->s_up, c->ev_down, 0));
  const size_t plane = sizeof(double) * c->grid.s[d];
  for (int m = 0; m < nv; m++)
    OSB_CUDA(c, cudaMemcpyAsync(c->fp.q[m] + (long long)plane0 * c->grid.s[d], src[m], plane * nplanes, cudaMemcpyHostToDevice, c->s_up));
  OSB_CUDA(c, cudaEventRecord(c->ev_up, c->s_up));
  return 0;
}
// The window is about to be advanced as a fresh block: the RK registers of the low-storage scheme enter the first stage multiplied
// by A[0] = 0, which does not clear what an earlier window left in the guard planes if that is not finite (0 * NaN)
static int window_reset_registers(osb_ctx *c) {
  const int nv = c->plan.nd + 2;
  for (int m = 0; m < nv; m++) OSB_CUDA(c, cudaMemsetAsync(c->fp.rk[m], 0, sizeof(double) * c->grid.n, c->stream));
  return 0;
}
int osb_host_planes_ready(osb_ctx *c) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  if (pipeline_init(c)) return 1;
  OSB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_up, 0));     // work enqueued on the compute stream from here on sees the uploads
  return window_reset_registers(c);
}
int osb_host_planes_download(osb_ctx *c, double *const *dst, int plane0, int nplanes) {
  if (!c || !dst) return 1;
  cudaSetDevice(c->device);
  if (pipeline_init(c) || plane_range_ok(c, plane0, nplanes)) return 1;
  const int d = c->plan.nd - 1, nv = c->plan.nd + 2;
  OSB_CUDA(c, cudaEventRecord(c->ev_comp, c->stream));
  OSB_CUDA(c, cudaStreamWaitEvent(c->s_down, c->ev_comp, 0));
  const size_t plane = sizeof(double) * c->grid.s[d];
  for (int m = 0; m < nv; m++)
    OSB_CUDA(c, cudaMemcpyAsync(dst[m], c->fp.q[m] + (long long)plane0 * c->grid.s[d], plane * nplanes, cudaMemcpyDeviceToHost, c->s_down));
  OSB_CUDA(c, cudaEventRecord(c->ev_down, c->s_down));
  c->down_pending = true;
  return 0;
}
// Staging copy of the whole block on the device: the host arrays cross PCIe once, in plane order, and every window takes its
// planes (guard and halo planes included, which neighbouring windows share) from here with device-to-device copies.
struct osb_staging {
  int device = 0, nv = 0, nplanes = 0;
  long long plane = 0;                       // doubles per plane
  double *buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  double *peer[2][5] = {{nullptr}};          // neighbour ranks' staging copies (slab-decomposed blocks), CUDA IPC
  cudaStream_t s_up = nullptr;
  std::string error;
};
int osb_staging_create(int device, int nv, long long plane_doubles, int nplanes, osb_staging **out) {
  if (!out || nv < 1 || nv > 5 || plane_doubles < 1 || nplanes < 1) return 1;
  *out = nullptr;
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return 2;
  osb_staging *st = new osb_staging();
  cudaGetDevice(&st->device);
  st->nv = nv; st->nplanes = nplanes; st->plane = plane_doubles;
  bool ok = cudaStreamCreateWithFlags(&st->s_up, cudaStreamNonBlocking) == cudaSuccess;
  for (int m = 0; m < nv && ok; m++) ok = cudaMalloc(&st->buf[m], sizeof(double) * plane_doubles * nplanes) == cudaSuccess;
  if (!ok) { cudaGetLastError(); osb_staging_destroy(st); return 3; }
  *out = st;
  return 0;
}
int osb_staging_destroy(osb_staging *st) {
  if (!st) return 0;
  cudaSetDevice(st->device);
  for (int side = 0; side < 2; side++) for (int m = 0; m < 5; m++) if (st->peer[side][m]) cudaIpcCloseMemHandle(st->peer[side][m]);
  for (int m = 0; m < 5; m++) if (st->buf[m]) cudaFree(st->buf[m]);
  if (st->s_up) cudaStreamDestroy(st->s_up);
  delete st;
  return 0;
}
const char *osb_staging_last_error(const osb_staging *st) { return st ? st->error.c_str() : "null stage"; }
#define OSB_STAGE_CUDA(st, call)                                                                                     \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { (st)->error = std::string(#call) + ": " + cudaGetErrorString(e_); return 1; } } while (0)
int osb_staging_upload(osb_staging *st, const double *const *src, int plane0, int nplanes) {
  if (!st || !src) return 1;
  cudaSetDevice(st->device);
  if (plane0 < 0 || nplanes < 1 || plane0 + nplanes > st->nplanes) { st->error = "plane range outside the staged block"; return 1; }
  for (int m = 0; m < st->nv; m++)
    OSB_STAGE_CUDA(st, cudaMemcpyAsync(st->buf[m] + plane0 * st->plane, src[m], sizeof(double) * st->plane * nplanes, cudaMemcpyHostToDevice, st->s_up));
  return 0;
}
int osb_staging_feed(osb_staging *st, osb_ctx *c, int stage_plane0, int plane0, int nplanes) {
  if (!st || !c) return 1;
  cudaSetDevice(c->device);
  if (pipeline_init(c) || plane_range_ok(c, plane0, nplanes)) return 1;
  const int d = c->plan.nd - 1, nv = c->plan.nd + 2;
  if (nv != st->nv || c->grid.s[d] != st->plane) return fail(c, "osb_staging_feed: the staged planes do not have this context's layout");
  if (stage_plane0 < 0 || stage_plane0 + nplanes > st->nplanes) return fail(c, "osb_staging_feed: plane range outside the staged block");
  // everything enqueued on the stage's upload stream so far must have landed; the planes being overwritten may still be on
  // their way to the host from the previous window of this context
  cudaEvent_t e;
  OSB_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  OSB_CUDA(c, cudaEventRecord(e, st->s_up));
  OSB_CUDA(c, cudaStreamWaitEvent(c->stream, e, 0));
  OSB_CUDA(c, cudaEventDestroy(e));            // released by the runtime once the wait has been satisfied
  if (c->down_pending) OSB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_down, 0));
  for (int m = 0; m < nv; m++)
    OSB_CUDA(c, cudaMemcpyAsync(c->fp.q[m] + (long long)plane0 * st->plane, st->buf[m] + (long long)stage_plane0 * st->plane,
                                sizeof(double) * st->plane * nplanes, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}
int osb_staging_fed(osb_ctx *c) {                // all planes of the window are in place: it starts as a fresh block
  if (!c) return 1;
  cudaSetDevice(c->device);
  return window_reset_registers(c);
}
// slab-decomposed blocks: the planes a rank's first and last windows need from its neighbours are pulled out of the neighbours'
// staging copies over NVLink (the caller orders the ranks: pull after every rank's boundary planes have landed, re-upload after
// every rank's pulls have)
int osb_staging_ipc_export(osb_staging *st, void *handles, int *nbytes) {
  if (!st || !handles || !nbytes) return 1;
  cudaSetDevice(st->device);
  for (int m = 0; m < st->nv; m++) {
    cudaIpcMemHandle_t h;
    OSB_STAGE_CUDA(st, cudaIpcGetMemHandle(&h, st->buf[m]));
    memcpy((char *)handles + m * sizeof(h), &h, sizeof(h));
  }
  *nbytes = st->nv * (int)sizeof(cudaIpcMemHandle_t);
  return 0;
}
int osb_staging_ipc_import(osb_staging *st, int side, const void *handles, int nbytes) {
  if (!st || !handles || side < 0 || side > 1) return 1;
  cudaSetDevice(st->device);
  if (nbytes != st->nv * (int)sizeof(cudaIpcMemHandle_t)) { st->error = "osb_staging_ipc_import: wrong handle size"; return 1; }
  for (int m = 0; m < st->nv; m++) {
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + m * sizeof(h), sizeof(h));
    void *p = nullptr;
    OSB_STAGE_CUDA(st, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    st->peer[side][m] = (double *)p;
  }
  return 0;
}
int osb_staging_pull(osb_staging *st, int side, int src_plane0, int dst_plane0, int nplanes) {
  if (!st || side < 0 || side > 1) return 1;
  cudaSetDevice(st->device);
  if (!st->peer[side][0]) { st->error = "osb_staging_pull: no neighbour imported on this side"; return 1; }
  if (src_plane0 < 0 || dst_plane0 < 0 || nplanes < 1 || dst_plane0 + nplanes > st->nplanes) { st->error = "osb_staging_pull: plane range outside the staged block"; return 1; }
  for (int m = 0; m < st->nv; m++)
    OSB_STAGE_CUDA(st, cudaMemcpyAsync(st->buf[m] + dst_plane0 * st->plane, st->peer[side][m] + src_plane0 * st->plane,
                                       sizeof(double) * st->plane * nplanes, cudaMemcpyDefault, st->s_up));
  return 0;
}
int osb_staging_sync(osb_staging *st) {
  if (!st) return 1;
  cudaSetDevice(st->device);
  OSB_STAGE_CUDA(st, cudaStreamSynchronize(st->s_up));
  return 0;
}

int osb_host_planes_sync(osb_ctx *c) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  if (c->s_up) OSB_CUDA(c, cudaStreamSynchronize(c->s_up));
  OSB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->s_down) OSB_CUDA(c, cudaStreamSynchronize(c->s_down));
  c->down_pending = false;
  return 0;
}

// ---- in-loop diagnostics ---------------------------------------------------------------------------
static int run_diag(osb_ctx *c, const double *field, double *out) {
  cudaSetDevice(c->device);
  const int nblocks = 148 * 4;
  if (!c->diag_buf) OSB_CUDA(c, cudaMalloc(&c->diag_buf, sizeof(double) * DIAG_N * (nblocks + 1)));
  double *partial = c->diag_buf + DIAG_N, *res = c->diag_buf;
  {
    Launcher L(c, OSB_FAM_USER);
    switch (c->plan.nd) {
      case 1: k_diag_partial<1><<<nblocks, 256, 0, c->stream>>>(c->grid, c->fp, c->pc, c->gp, field, partial); break;
      case 2: k_diag_partial<2><<<nblocks, 256, 0, c->stream>>>(c->grid, c->fp, c->pc, c->gp, field, partial); break;
      default: k_diag_partial<3><<<nblocks, 256, 0, c->stream>>>(c->grid, c->fp, c->pc, c->gp, field, partial); break;
    }
  }
  { Launcher L(c, OSB_FAM_USER); k_diag_final<<<1, 32, 0, c->stream>>>(partial, nblocks, res); }
  OSB_CUDA(c, cudaMemcpyAsync(out, res, sizeof(double) * DIAG_N, cudaMemcpyDeviceToHost, c->stream));
  OSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}
int osb_nan_check(osb_ctx *c, const char *name, long long *n_bad) {
  if (!c || !name || !n_bad) return 1;
  Field *f = find_field(c, name);
  if (!f) return fail(c, std::string("unknown field ") + name);
  refresh_primitives_for(c, f);
  double out[DIAG_N];
  if (run_diag(c, f->dev, out)) return 1;
  *n_bad = (long long)out[5];
  return 0;
}
int osb_diagnostics(osb_ctx *c, double *sums) {
  if (!c || !sums) return 1;
  return run_diag(c, nullptr, sums);
}

// Instrumentation for the bench: count, over the next steps, the TENO5 characteristic waves that leave the all-pass shortcut
// (osb_math.cuh teno5_front) for the full cut-off path.  enable != 0 arms the counter (and clears it); the count is returned.
int osb_slow_path_count(osb_ctx *c, int enable, long long *count) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  if (!c->slow_count) OSB_CUDA(c, cudaMalloc(&c->slow_count, sizeof(unsigned long long)));
  OSB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (count) {
    unsigned long long v = 0;
    OSB_CUDA(c, cudaMemcpy(&v, c->slow_count, sizeof(v), cudaMemcpyDeviceToHost));
    *count = (long long)v;
  }
  OSB_CUDA(c, cudaMemset(c->slow_count, 0, sizeof(unsigned long long)));
  c->sp.slow_count = enable ? c->slow_count : nullptr;
  c->count_slow = enable != 0;
  drop_graph(c);
  return 0;
}

int osb_launch_count(const osb_ctx *c, long long *n) { if (!c || !n) return 1; *n = c->launches; return 0; }

int osb_profile_step(osb_ctx *c, double *fam_ms, long long *fam_n) {
  if (!c || !fam_ms) return 1;
  cudaSetDevice(c->device);
  OSB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->profiling = true; c->prof_events.clear();
  int rc = do_step(c, 1);
  c->profiling = false;
  OSB_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < OSB_NFAM; i++) { fam_ms[i] = 0.0; if (fam_n) fam_n[i] = 0; }
  const bool list = getenv("OSB_PROFILE_LIST") != nullptr;       // per-launch times in launch order (kernel experiments)
  int idx = 0;
  for (auto &pe : c->prof_events) {
    float t = 0; cudaEventElapsedTime(&t, pe.second.first, pe.second.second);
    if (list && t > 0.2f) fprintf(stderr, "osb_profile launch %d family %d %.3f ms\n", idx, pe.first, t);
    idx++;
    fam_ms[pe.first] += t; if (fam_n) fam_n[pe.first]++;
    cudaEventDestroy(pe.second.first); cudaEventDestroy(pe.second.second);
  }
  c->prof_events.clear();
  return rc;
}

// ---- slab decomposition: CUDA IPC peer access -----------------------------------------------------
int osb_ipc_export(osb_ctx *c, void *handles, int *nbytes) {
  if (!c || !handles || !nbytes) return 1;
  cudaSetDevice(c->device);
  const int nv = c->plan.nd + 2;
  for (int t = 0; t < 2; t++)
    for (int m = 0; m < nv; m++) {
      cudaIpcMemHandle_t h;
      OSB_CUDA(c, cudaIpcGetMemHandle(&h, c->own_buf[t][m]));
      memcpy((char *)handles + (t * nv + m) * sizeof(h), &h, sizeof(h));
    }
  {
    cudaIpcMemHandle_t h;
    OSB_CUDA(c, cudaIpcGetMemHandle(&h, c->flags));
    memcpy((char *)handles + 2 * nv * sizeof(h), &h, sizeof(h));
  }
  *nbytes = (2 * nv + 1) * (int)sizeof(cudaIpcMemHandle_t);
  if (c->gp.theta) {       // adaptive TENO: the sensor array too
    cudaIpcMemHandle_t h;
    OSB_CUDA(c, cudaIpcGetMemHandle(&h, c->gp.theta));
    memcpy((char *)handles + *nbytes, &h, sizeof(h));
    *nbytes += (int)sizeof(h);
  }
  {                        // slab thickness (neighbours may differ by one plane when the block does not divide evenly)
    const long long npd = c->grid.np[c->plan.nd - 1];
    memcpy((char *)handles + *nbytes, &npd, sizeof(npd));
    *nbytes += (int)sizeof(npd);
  }
  return 0;
}
int osb_ipc_import(osb_ctx *c, int side, const void *handles, int nbytes) {
  if (!c || !handles || side < 0 || side > 1) return 1;
  cudaSetDevice(c->device);
  const int nv = c->plan.nd + 2;
  const int nh = 2 * nv + 1 + (c->gp.theta ? 1 : 0);
  if (nbytes != nh * (int)sizeof(cudaIpcMemHandle_t) + (int)sizeof(long long)) return fail(c, "osb_ipc_import: wrong handle size");
  memcpy(&c->peer_np[side], (const char *)handles + nh * sizeof(cudaIpcMemHandle_t), sizeof(long long));
  if (c->gp.theta) {
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + (2 * nv + 1) * sizeof(h), sizeof(h));
    void *p = nullptr;
    OSB_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_theta[side] = (double *)p;
  }
  for (int t = 0; t < 2; t++)
    for (int m = 0; m < nv; m++) {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char *)handles + (t * nv + m) * sizeof(h), sizeof(h));
      void *p = nullptr;
      OSB_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      c->peer_buf[side][t][m] = (double *)p;
    }
  {
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + 2 * nv * sizeof(h), sizeof(h));
    void *p = nullptr;
    OSB_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_flags[side] = (unsigned long long *)p;
  }
  c->peer_open[side] = true;
  drop_graph(c);
  return 0;
}
int osb_halo_push(osb_ctx *c) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  return push_planes_memcpy(c);
}
}  // extern "C"
namespace {
int push_planes_memcpy(osb_ctx *c) {
  const Plan &P = c->plan;
  const GridDev &g = c->grid;
  const int d = P.nd - 1;   // slabs along the slowest axis
  const int nv = P.nd + 2;
  int hm, hp; scheme_halos(P, hm, hp);
  // planes along the slowest axis are contiguous: [plane index] * s[d], each s[d] doubles
  const size_t plane = sizeof(double) * g.s[d];
  for (int m = 0; m < nv; m++) {
    if (P.bc[d][1].kind == BC_EXCHANGE && c->peer_open[1]) {
      // my top hm planes [np-hm, np) -> high neighbour's low halo [-hm, 0)
      OSB_CUDA(c, cudaMemcpyAsync(c->peer_buf[1][c->swap_parity][m] + (g.h - hm) * g.s[d], c->fp.q[m] + (g.h + g.np[d] - hm) * g.s[d], plane * hm, cudaMemcpyDeviceToDevice, c->stream));
    }
    if (P.bc[d][0].kind == BC_EXCHANGE && c->peer_open[0]) {
      // my bottom hp planes [0, hp) -> low neighbour's high halo [np_lo, np_lo+hp) (the neighbour's own thickness)
      OSB_CUDA(c, cudaMemcpyAsync(c->peer_buf[0][c->swap_parity][m] + (g.h + c->peer_np[0]) * g.s[d], c->fp.q[m] + g.h * g.s[d], plane * hp, cudaMemcpyDeviceToDevice, c->stream));
    }
  }
  return 0;
}
}  // namespace
extern "C" {

int osb_measure_fp64_peak(int device, double *tflops) {
  if (!tflops) return 1;
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return 2;
  cudaDeviceProp prop; int dev = 0; cudaGetDevice(&dev);
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 2;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  double *out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return 3;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_dfma_peak<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  *tflops = best;
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

}  // extern "C"
