/* osbli_b200.h -- C ABI of the B200-native execution back end for OpenSBLI.
 *
 * This is the drop-in boundary: everything the reference obtains today by
 * emitting OPS-C text (opensbli/code_generation/opsc.py:250-282) and linking the external OPS
 * runtime is reached through these entry points instead.  The caller is the Python back-end class
 * `opensbli_b200.B200(alg)` (same call shape as `OPSC(alg)`, opsc.py:253) via ctypes; signatures
 * carry only plain pointers and sizes.  Single-threaded host control, one context per process/GPU.
 * Every function returns 0 on success and a non-zero code otherwise (never throws or aborts);
 * the message is available from osb_last_error().  Device memory is owned by the context; host
 * buffers are owned by the caller.
 *
 * Field arrays cross the boundary in the reference's own layout (opsc.py:693-722 ops_decl_dat,
 * halo rule opsc.py:707-711): one array per variable, x fastest, padded by 5 halo points on both
 * sides of every active dimension.  Field names are the reference's dataset names without the
 * block suffix: rho, rhou0.., rhoE, u0.., p, a, T, Residual0.., tempRK_rho.. / rho_RKold..
 */
#ifndef OSBLI_B200_H
#define OSBLI_B200_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct osb_ctx osb_ctx;

/* Create a solver context from a plan (text, "osbli_plan 1" format written by
 * opensbli_b200.plan.to_text).  Replaces: OPSC.__init__ code emission + ops_init / ops_decl_block /
 * ops_decl_dat / ops_decl_stencil / ops_decl_halo / ops_partition (opsc.py:435-593, 693-722).
 * device < 0 selects the current CUDA device. */
int osb_create(const char *plan_text, int device, osb_ctx **ctx);
int osb_destroy(osb_ctx *ctx);
/* ctx may be NULL (error of a failed osb_create). */
const char *osb_last_error(const osb_ctx *ctx);

/* Runtime constants; replaces ops_decl_const + the `name=Input;` substitution
 * (opsc.py:625-654, utilities/helperfunctions.py:130-149). */
int osb_set_const_f64(osb_ctx *ctx, const char *name, double value);
int osb_get_const_f64(const osb_ctx *ctx, const char *name, double *value);

/* Iteration number seen by time-dependent source terms: the loop counter `iter` of the generated time loop
 * (algorithm.py:440-474; e.g. the forcing sin(omega dt iter) of apps/transitional_SBLI/transitional_SBLI.py:77-89).
 * Starts at the plan's iteration0 (restart_iteration_no), advances by one per completed step. */
int osb_set_iteration(osb_ctx *ctx, long long iteration);
long long osb_get_iteration(const osb_ctx *ctx);

/* Field access; replaces ops_decl_dat(_hdf5) initial values and ops_fetch_dat_hdf5_file
 * (opsc.py:693-722, core/io_hdf5.py:99-127).  dims/halo_m/halo_p have 3 entries. */
int osb_num_fields(const osb_ctx *ctx);
const char *osb_field_name(const osb_ctx *ctx, int index);
int osb_field_info(const osb_ctx *ctx, const char *name, int *dims, int *halo_m, int *halo_p);
int osb_upload(osb_ctx *ctx, const char *name, const double *host_padded);
int osb_download(osb_ctx *ctx, const char *name, double *host_padded);
/* One value of a dataset at grid index (i, j, k) (unused indices 0): what the reference's SimulationMonitor reads with a
 * one-point reduction loop (simulation_monitors.py:17-160).  Synchronises the stream. */
int osb_read_point(osb_ctx *ctx, const char *name, int i, int j, int k, double *value);
int osb_device_ptr(osb_ctx *ctx, const char *name, double **device_ptr);
/* State imposed by a `dirichlet_field` boundary (equations of a DirichletBC that depend on the position along the face,
 * dirichlet.py:28-41): table[m][t], m < ndim+2, t = padded tangential index (the field index with dimension dir removed). */
int osb_upload_face(osb_ctx *ctx, int dir, int side, const double *table);

/* The time loop body, in the reference's program order (algorithm.py:440-474):
 *   per iteration: BCs ; [save] ; per stage: constituent relations, spatial kernels, RK update, BCs.
 * osb_step enqueues nsteps iterations on the context's stream and returns without synchronising. */
int osb_step(osb_ctx *ctx, int nsteps);
/* osb_sync also surfaces a neighbour time-out of a decomposed run: a device-side wait gives up after ~10 s, later waits
 * return at once, and the kernels already enqueued keep stepping on halos that were not refreshed -- the state is invalid from
 * the reported exchange epoch on. */
int osb_sync(osb_ctx *ctx);
/* Pieces of the loop, for parity tests: boundary conditions on q; residual of the current q
 * (constituent relations + all spatial kernels) left in Residual0.. */
int osb_apply_bcs(osb_ctx *ctx);
int osb_residual(osb_ctx *ctx);

/* Stage-level control for a decomposed run (one context per rank): osb_step_begin = the iteration-start
 * BCs (+ rk_sbli save); osb_stage(s) = constituent relations, spatial kernels, RK update and the rank-local
 * BCs of stage s.  Between stages the caller exchanges halos with osb_halo_push. */
int osb_step_begin(osb_ctx *ctx);
int osb_stage(osb_ctx *ctx, int stage);

/* Same as osb_step, timed on the device with CUDA events on the context's stream (includes a sync). */
int osb_step_timed(osb_ctx *ctx, int nsteps, double *elapsed_ms);
/* Device-side stopwatch on the context's stream (CUDA events): start records an event, stop records a second
 * one, synchronises on it and returns the elapsed device time; host-driven gaps between launches are included. */
int osb_timer_start(osb_ctx *ctx);
int osb_timer_stop(osb_ctx *ctx, double *elapsed_ms);
/* End-to-end: copy the conserved arrays from host (pinned or pageable, reference layout), advance
 * nsteps, copy them back.  q_in/q_out hold ndim+2 pointers.  elapsed_ms covers H2D + steps + D2H. */
int osb_advance_host(osb_ctx *ctx, const double *const *q_in, double *const *q_out, int nsteps, double *elapsed_ms);
/* Window pipeline over a host-resident block (opensbli_b200/hostpipe.py drives it; the reference reads and writes whole
 * blocks through ops_fetch / HDF5, io_hdf5.py:99-127, so there is no counterpart to cite): the context holds one WINDOW of
 * the block -- a run of planes along the slowest axis whose cut faces carry the 'open' boundary type -- and owns an upload
 * and a download stream next to its compute stream.
 *   osb_host_planes_upload    copy nplanes planes of each conserved array (src[m] points at the first plane in PINNED host
 *                             memory) into padded local planes [plane0, plane0 + nplanes); asynchronous; waits for a pending
 *                             download of this context first
 *   osb_host_planes_ready     make the compute stream wait for the uploads enqueued so far
 *   osb_host_planes_download  copy padded local planes [plane0, ...) of the state as it is after the work enqueued on the
 *                             compute stream so far; asynchronous
 *   osb_host_planes_sync      wait for all three streams */
int osb_host_planes_upload(osb_ctx *ctx, const double *const *src, int plane0, int nplanes);
int osb_host_planes_ready(osb_ctx *ctx);
int osb_host_planes_download(osb_ctx *ctx, double *const *dst, int plane0, int nplanes);
int osb_host_planes_sync(osb_ctx *ctx);
/* Staging copy of the whole block on the device, so that the host arrays cross PCIe once although neighbouring windows
 * share their guard and halo planes: osb_staging_upload copies planes host -> stage on the stage's own stream (pinned memory,
 * asynchronous); osb_staging_feed makes the window's compute stream wait for the uploads enqueued so far (and for a pending
 * download of that context) and copies stage planes [stage_plane0, +nplanes) into the window's padded local planes
 * [plane0, +nplanes) device to device; osb_staging_fed is called once per window after its last feed. */
typedef struct osb_staging osb_staging;
int osb_staging_create(int device, int nv, long long plane_doubles, int nplanes, osb_staging **out);
int osb_staging_destroy(osb_staging *stage);
const char *osb_staging_last_error(const osb_staging *stage);
int osb_staging_upload(osb_staging *stage, const double *const *src, int plane0, int nplanes);
int osb_staging_feed(osb_staging *stage, osb_ctx *ctx, int stage_plane0, int plane0, int nplanes);
int osb_staging_fed(osb_ctx *ctx);
int osb_staging_sync(osb_staging *stage);
/* Slab-decomposed blocks (one rank per GPU): a rank's staging copy carries guard + halo planes below and above its slab, which
 * it pulls out of the neighbours' staging copies over NVLink (CUDA IPC, same exchange of handles as osb_ipc_export / _import);
 * the caller orders the ranks (pull after every rank's boundary planes have landed).  With them in place the rank's windows
 * need no per-stage halo exchange at all. */
int osb_staging_ipc_export(osb_staging *stage, void *handles, int *nbytes);
int osb_staging_ipc_import(osb_staging *stage, int side, const void *handles, int nbytes);
int osb_staging_pull(osb_staging *stage, int side, int src_plane0, int dst_plane0, int nplanes);

/* Instrumentation: number of kernels launched by this context so far; per-family device time of
 * one profiled step (events around each launch).  Families: see OSB_FAM_*. */
enum { OSB_FAM_PRIM = 0, OSB_FAM_FLUX = 1, OSB_FAM_CENTRAL = 2, OSB_FAM_VISCOUS = 3, OSB_FAM_RK = 4, OSB_FAM_BC = 5,
       OSB_FAM_SYNC = 6 /* waiting for neighbour ranks */, OSB_FAM_USER = 7 /* run-time compiled user kernels */, OSB_NFAM = 8 };
int osb_launch_count(const osb_ctx *ctx, long long *count);
/* Bench instrumentation: number of TENO5 characteristic waves (5 per interface per sweep) that needed the full cut-off
 * evaluation instead of the all-stencils-pass shortcut since the counter was armed.  enable != 0 arms and clears it. */
int osb_slow_path_count(osb_ctx *ctx, int enable, long long *count);
int osb_profile_step(osb_ctx *ctx, double *family_ms /* [OSB_NFAM] */, long long *family_launches /* [OSB_NFAM] */);

/* In-loop diagnostics (device reductions, deterministic summation order; synchronise the stream):
 * osb_nan_check counts the non-finite values of a dataset over the interior points -- what `ops_NaNcheck(rho_B0)` tests in the
 * reference's time loop (core/diagnostics/simulation_monitors.py:112-113,180-181; utilities/helperfunctions.py:172-190,
 * print_iteration_ops(NaN_check=...)).
 * osb_diagnostics returns interior sums of this context's block: [0] sum rho, [1] sum 1/2 rho |u|^2 (kinetic energy),
 * [2] sum 1/2 rho |curl u|^2 (enstrophy; 4th-order central differences, metric-scaled on stretched grids), [3] sum rhoE,
 * [4] max Mach number, [5] number of points with non-finite rho or rhoE.  Ranks of a decomposed run add / max their values.
 * (The reference computes the Taylor-Green kinetic energy and enstrophy offline from its dumps.) */
enum { OSB_NDIAG = 6 };
int osb_nan_check(osb_ctx *ctx, const char *name, long long *n_nonfinite);
int osb_diagnostics(osb_ctx *ctx, double *sums /* [OSB_NDIAG] */);

/* Point-wise user kernels (the reference's `User kernel` loops, e.g. the statistics accumulation of
 * apps/channel_flow/*: stats.py, opsc.py kernel emission): app-specific arithmetic outside the solver's hot loops, given as
 * CUDA C source of one `extern "C" __global__` entry with the signature
 *     entry(long long off, int n0, int n1, int n2, int lo0, int lo1, int lo2, long long s1, long long s2, UserFields f)
 * where `struct UserFields { double *p[OSB_MAX_USER_FIELDS]; long long iter; }` holds the arrays named in `fields` (comma separated)
 * and the loop counter of the time loop (what the reference hands to a kernel as ops_arg_gbl `iter`; a kernel that reads `f.iter`
 * keeps the step out of the CUDA-graph replay).  A name
 * prefixed with '+' is written by the kernel: if it does not exist yet it is created zero-initialised, as OPS declares
 * datasets (opsc.py:693-722).  A name the kernel only reads must exist (solver field, or osb_create_field + osb_upload):
 * otherwise the call fails -- reading zeros in place of a dataset nobody provided would be a silent wrong answer.
 * u_i, p, a, T read by a user kernel are the constituent relations of the state at the time of the launch.
 * Compiled once with NVRTC for sm_100a.
 * when = 0: launched at the end of every iteration of osb_step (after the last stage's boundary conditions);
 * when = 1: launched by osb_run_user_kernels(ctx, 1) (loops after the time loop);
 * when = 100 + 2 dir + side: the boundary kernel of that face, for plans whose `bc dir side generic` line hands the face to a
 *   run-time compiled kernel (boundary classes without a hand-written kernel: the reference's kernel equations printed with
 *   relative offsets, bc_core.py:104-198); launched wherever the face's boundary condition is applied, in the reference's order;
 * generic path (plan `conv generic`, `generic_stages N`: a program outside the hand-written kernels -- every loop of its time
 * step is such a kernel, boundary kernels and periodic copies included, launched in registration order = program order,
 * algorithm.py:440-474):  when = 200: at the start of every iteration;  when = 209: in every RK stage;  when = 210 + s: in
 * stage s only (a loop that reads the stage counter, rkA[stage], is compiled once per stage);  when = 0: after the last stage. */
enum { OSB_MAX_USER_FIELDS = 96 };
/* Declare an additional dataset (zero-initialised; no-op if it exists): what ops_decl_dat does for a dataset that only user
 * kernels touch, e.g. a coordinate array x0 evaluated by the cold path and uploaded with osb_upload (opsc.py:693-722). */
int osb_create_field(osb_ctx *ctx, const char *name);
int osb_add_user_kernel(osb_ctx *ctx, const char *cuda_source, const char *entry, const char *fields, const int range[6], int when);
int osb_run_user_kernels(osb_ctx *ctx, int when);

/* Multi-GPU (slab decomposition along the slowest axis): direct peer access to a neighbour's
 * arrays through CUDA IPC.  See INTEGRATION.md. */
int osb_ipc_export(osb_ctx *ctx, void *handles /* room for (2 nq + 2) * 64 + 8 bytes: q buffers, Residual buffers, flag words, shock sensor (adaptive TENO), slab thickness */, int *nbytes);
int osb_ipc_import(osb_ctx *ctx, int side /* 0 = low neighbour, 1 = high neighbour */, const void *handles, int nbytes);
/* push this rank's boundary planes of q into the neighbours' halo planes (peer stores over NVLink) */
int osb_halo_push(osb_ctx *ctx);

/* FP64 pipe micro-benchmark (dependent-free DFMA chains); used by bench.py for the roofline peak. */
int osb_measure_fp64_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
