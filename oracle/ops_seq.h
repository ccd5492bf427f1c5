// ops_seq.h -- header-only, sequential / OpenMP stand-in for the 20-odd OPS symbols that
// OpenSBLI's OPSC back end emits (reference call sites: opensbli/core/kernel.py:232-274,
// opensbli/code_generation/opsc.py:445-468,531-593,693-722, opensbli/core/io_hdf5.py:99-127,
// opensbli/code_generation/algorithm/algorithm.py:301-327).
//
// TEST INFRASTRUCTURE ONLY (oracle).  The real OPS library is third-party, un-vendored and
// unpinned (OP-DSL/OPS); this stand-in executes the reference's *generated C* unmodified:
//   * ops_par_loop  = nested loops, x fastest, in program order (OPS "seq" semantics);
//                     with -DOPS_OMP the outer indices are OpenMP-parallel ("OPS openmp").
//   * ops_dat       = one zero-initialised padded array per variable, element (i,j,k) at
//                     (i-d_m0) + pdim0*((j-d_m1) + pdim1*(k-d_m2)).
//   * ops_halo_transfer = gather to a temporary, then scatter (the TENO side-1 periodic copy
//                     overlaps interior plane 0, so buffering matters).
//   * ops_fetch_dat_hdf5_file = raw binary dump  [int32 hdr[10]; double data[]].
// Extra (not OPS): ops_env_int / ops_env_double let the harness override simulation
// parameters (grid size, niter, dt ...) at run time through environment variables.
#ifndef OSBLI_OPS_SEQ_H
#define OSBLI_OPS_SEQ_H
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <cmath>
#include <vector>
#include <string>
#include <chrono>
#include <utility>
#ifdef OPS_OMP
#include <omp.h>
#endif

#define OPS_READ 0
#define OPS_WRITE 1
#define OPS_RW 2
#define OPS_INC 3
#define OPS_MIN 4
#define OPS_MAX 5

#define OPS_MAX_ARGS 192
static int ops_xdim[OPS_MAX_ARGS];
static int ops_ydim[OPS_MAX_ARGS];
#include "ops_seq_acc.h"   // generated: OPS_ACC0..OPS_ACC191 for the arity chosen by OPS_1D/2D/3D

struct ops_dat_core {
  int ndim;
  int size[3], d_m[3], d_p[3], pdim[3];
  std::vector<double> data;
  std::string name;
  inline double *at(int i, int j, int k) {
    return data.data() + (size_t)(i - d_m[0]) + (size_t)pdim[0] * ((size_t)(j - d_m[1]) + (size_t)pdim[1] * (size_t)(k - d_m[2]));
  }
};
typedef ops_dat_core *ops_dat;
static std::vector<ops_dat> ops_all_dats;   // registry, for the OSBLI_DUMP_ALL debugging dump
struct ops_block_core { int ndim; std::string name; };
typedef ops_block_core *ops_block;
typedef int ops_stencil;
struct ops_halo_core { ops_dat from, to; int iter[3], from_base[3], to_base[3]; };
typedef ops_halo_core *ops_halo;
struct ops_halo_group_core { std::vector<ops_halo> halos; };
typedef ops_halo_group_core *ops_halo_group;
struct ops_reduction_core { double value; std::string name; };
typedef ops_reduction_core *ops_reduction;

struct ops_arg { int kind; ops_dat dat; void *gbl; int acc; };   // kind 0 dat, 1 gbl, 2 idx, 3 reduce

static inline int ops_env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return s ? atoi(s) : dflt;
}
static inline double ops_env_double(const char *name, double dflt) {
  const char *s = getenv(name);
  return s ? strtod(s, NULL) : dflt;
}

static inline void ops_init(int, char **, int) {}
static inline void ops_partition(const char *) {}
static inline void ops_decl_const(const char *, int, const char *, void *) {}
static inline void ops_timing_output(FILE *) {}
static inline void ops_printf(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); }
static inline void ops_fprintf(FILE *f, const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(f, fmt, ap); va_end(ap); }
static inline void ops_timers(double *cpu, double *et) {
  using namespace std::chrono;
  *et = duration<double>(steady_clock::now().time_since_epoch()).count();
  *cpu = *et;
}
static inline ops_block ops_decl_block(int ndim, const char *name) { return new ops_block_core{ndim, name}; }
static inline ops_stencil ops_decl_stencil(int, int, int *, const char *) { return 0; }

static inline ops_dat ops_decl_dat(ops_block b, int, int *size, int *, int *d_m, int *d_p, double *, const char *, const char *name) {
  ops_dat d = new ops_dat_core();
  d->ndim = b->ndim; d->name = name;
  size_t n = 1;
  for (int i = 0; i < 3; i++) {
    d->size[i] = i < b->ndim ? size[i] : 1;
    d->d_m[i] = i < b->ndim ? d_m[i] : 0;
    d->d_p[i] = i < b->ndim ? d_p[i] : 0;
    d->pdim[i] = d->size[i] - d->d_m[i] + d->d_p[i];
    n *= (size_t)d->pdim[i];
  }
  d->data.assign(n, 0.0);
  ops_all_dats.push_back(d);
  return d;
}

static inline ops_arg ops_arg_dat(ops_dat d, int, ops_stencil, const char *, int acc) { return ops_arg{0, d, NULL, acc}; }
template <class T> static inline ops_arg ops_arg_gbl(T *p, int, const char *, int acc) { return ops_arg{1, NULL, (void *)p, acc}; }
static inline ops_arg ops_arg_idx() { return ops_arg{2, NULL, NULL, 0}; }
static inline ops_reduction ops_decl_reduction_handle(int, const char *, const char *name) { return new ops_reduction_core{0.0, name}; }
static inline ops_arg ops_arg_reduce(ops_reduction r, int, const char *, int acc) { return ops_arg{3, NULL, (void *)&r->value, acc}; }
template <class T> static inline void ops_reduction_result(ops_reduction r, T *out) { *out = (T)r->value; r->value = 0.0; }

struct ops_ptr {
  void *p;
  template <class T> operator T *() const { return (T *)p; }
};

template <class K, size_t... I>
static inline void ops_call_(K kernel, ops_arg *a, int i, int j, int k, int *idx, std::index_sequence<I...>) {
  kernel(ops_ptr{a[I].kind == 0 ? (void *)a[I].dat->at(i, j, k) : (a[I].kind == 2 ? (void *)idx : a[I].gbl)}...);
}

template <class K, class... A>
static inline void ops_par_loop(K kernel, const char *, ops_block, int ndim, int *range, A... args) {
  constexpr int N = sizeof...(A);
  ops_arg a[N] = {args...};
  bool has_reduce = false;
  for (int n = 0; n < N; n++) {
    if (a[n].kind == 0) { ops_xdim[n] = a[n].dat->pdim[0]; ops_ydim[n] = a[n].dat->pdim[1]; }
    if (a[n].kind == 3) has_reduce = true;
  }
  int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  for (int d = 0; d < ndim; d++) { lo[d] = range[2 * d]; hi[d] = range[2 * d + 1]; }
  (void)has_reduce;
#ifdef OPS_OMP
  if (!has_reduce) {
    if (ndim == 1) {
#pragma omp parallel for schedule(static)
      for (int i = lo[0]; i < hi[0]; i++) { int idx[3] = {i, 0, 0}; ops_call_(kernel, a, i, 0, 0, idx, std::make_index_sequence<N>{}); }
    } else {
#pragma omp parallel for collapse(2) schedule(static)
      for (int k = lo[2]; k < hi[2]; k++)
        for (int j = lo[1]; j < hi[1]; j++)
          for (int i = lo[0]; i < hi[0]; i++) { int idx[3] = {i, j, k}; ops_call_(kernel, a, i, j, k, idx, std::make_index_sequence<N>{}); }
    }
    return;
  }
#endif
  for (int k = lo[2]; k < hi[2]; k++)
    for (int j = lo[1]; j < hi[1]; j++)
      for (int i = lo[0]; i < hi[0]; i++) { int idx[3] = {i, j, k}; ops_call_(kernel, a, i, j, k, idx, std::make_index_sequence<N>{}); }
}

static inline ops_halo ops_decl_halo(ops_dat from, ops_dat to, int *iter, int *from_base, int *to_base, int *, int *) {
  ops_halo h = new ops_halo_core();
  h->from = from; h->to = to;
  for (int i = 0; i < 3; i++) {
    h->iter[i] = i < from->ndim ? iter[i] : 1;
    h->from_base[i] = i < from->ndim ? from_base[i] : 0;
    h->to_base[i] = i < from->ndim ? to_base[i] : 0;
  }
  return h;
}
static inline ops_halo_group ops_decl_halo_group(int n, ops_halo *h) {
  ops_halo_group g = new ops_halo_group_core();
  g->halos.assign(h, h + n);
  return g;
}
static inline void ops_halo_transfer(ops_halo_group g) {
  for (ops_halo h : g->halos) {
    std::vector<double> tmp((size_t)h->iter[0] * h->iter[1] * h->iter[2]);
    size_t n = 0;
    for (int k = 0; k < h->iter[2]; k++)
      for (int j = 0; j < h->iter[1]; j++)
        for (int i = 0; i < h->iter[0]; i++)
          tmp[n++] = *h->from->at(h->from_base[0] + i, h->from_base[1] + j, h->from_base[2] + k);
    n = 0;
    for (int k = 0; k < h->iter[2]; k++)
      for (int j = 0; j < h->iter[1]; j++)
        for (int i = 0; i < h->iter[0]; i++)
          *h->to->at(h->to_base[0] + i, h->to_base[1] + j, h->to_base[2] + k) = tmp[n++];
  }
}

static inline void ops_NaNcheck(ops_dat d) {
  for (double v : d->data)
    if (v != v) { printf("NaN detected in %s\n", d->name.c_str()); exit(2); }
}

static inline void ops_fetch_block_hdf5_file(ops_block, const char *) {}
// raw dump:  $OSBLI_OUT/<file>.<dat>.bin = int32 {ndim, size[3], d_m[3], d_p[3]} + padded doubles
static inline void ops_fetch_dat_hdf5_file(ops_dat d, const char *fname) {
  const char *out = getenv("OSBLI_OUT");
  if (!out) return;   // timing runs: no dump
  std::string path = std::string(out) + "/" + fname + "." + d->name + ".bin";
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) { perror(path.c_str()); exit(3); }
  int hdr[10] = {d->ndim, d->size[0], d->size[1], d->size[2], d->d_m[0], d->d_m[1], d->d_m[2], d->d_p[0], d->d_p[1], d->d_p[2]};
  fwrite(hdr, sizeof(int), 10, f);
  fwrite(d->data.data(), sizeof(double), d->data.size(), f);
  fclose(f);
}
// OSBLI_DUMP_ALL=1: dump every declared dat (work arrays, residuals, primitives) at exit.
static inline void ops_exit() {
  if (getenv("OSBLI_DUMP_ALL") && getenv("OSBLI_OUT"))
    for (ops_dat d : ops_all_dats) ops_fetch_dat_hdf5_file(d, "all");
}
#endif
