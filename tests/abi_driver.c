/*
 * abi_driver.c -- a plain-C caller of libcmdg.so: no Python, no torch, no C++.
 *
 * Demonstrates the drop-in boundary of include/cmdg.h the way a foreign host (Julia's ccall) uses it:
 * dlopen the shared library, resolve every entry point by name, hand over device pointers to arrays in the
 * reference's own layouts, and call create / bind_grid / bind_state / tendency / lsrk_steps / destroy.
 *
 * Input: a directory written by tests/test_gpu_parity.py::test_plain_c_abi_driver (the oracle dumps the
 * raw arrays of a small dry-baroclinic-wave case plus the tendency and the state after `nsteps` LSRK54
 * steps that it computed itself):
 *     meta.txt                     key value lines: sizes, descriptor fields, dt, nsteps
 *     vgeo.bin sgeo.bin D.bin      Float64, reference layout (vgeo Np x 25 x nelem, sgeo 5 x Nfp x 6 x nelem, D Julia order)
 *     vmapM.bin vmapP.bin elemtobndy.bin   Int64, 1-based
 *     Q.bin aux.bin                Float64  Np x nstate x nelem / Np x naux x nelem
 *     expect_tendency.bin expect_state.bin   Float64, real elements
 * Output (stdout): `ABI_DRIVER tendency_rel_l2=<x> state_rel_l2=<y> launches=<n>`; exit code 0 iff both
 * relative L2 differences are <= 1e-12.
 *
 * Build (done by __graft_entry__.build()):
 *   gcc -O2 -std=c99 -Iinclude -I/usr/local/cuda/include tests/abi_driver.c -o tests/abi_driver \
 *       -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64 -lcudart -ldl -lm
 */
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cmdg.h"

#define DIE(...) do { fprintf(stderr, "abi_driver: " __VA_ARGS__); fprintf(stderr, "\n"); exit(2); } while (0)
#define CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) DIE("%s: %s", #x, cudaGetErrorString(e_)); } while (0)

/* entry points, resolved with dlsym: the typedefs restate the prototypes of include/cmdg.h */
typedef int (*create_fn)(const cmdg_desc *, cmdg_handle *);
typedef int (*destroy_fn)(cmdg_handle);
typedef const char *(*last_error_fn)(cmdg_handle);
typedef int (*bind_grid_fn)(cmdg_handle, const void *, const void *, const int64_t *, const int64_t *,
                            const int64_t *, const void *, const int64_t *, int64_t, const int64_t *, int64_t,
                            const int64_t *, int64_t, const int64_t *, int64_t, const int32_t *,
                            const int64_t *, const int64_t *, int32_t);
typedef int (*bind_state_fn)(cmdg_handle, void *, void *);
typedef int (*tendency_fn)(cmdg_handle, void *, void *, double, double, double, cmdg_stream);
typedef int (*lsrk_steps_fn)(cmdg_handle, void *, void *, double, double, int32_t, const double *,
                             const double *, const double *, int64_t, cmdg_stream);
typedef int (*lsrk_steps_host_fn)(cmdg_handle, void *, double, double, int32_t, const double *, const double *,
                                  const double *, int64_t);
typedef int (*sync_fn)(cmdg_handle);
typedef int64_t (*launches_fn)(cmdg_handle);
typedef int (*version_fn)(void);

static char g_dir[4096];

static void *read_file(const char *name, size_t *bytes) {
  char path[4608];
  snprintf(path, sizeof path, "%s/%s", g_dir, name);
  FILE *f = fopen(path, "rb");
  if (!f) DIE("cannot open %s", path);
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  void *p = malloc(n > 0 ? (size_t)n : 1);
  if (fread(p, 1, (size_t)n, f) != (size_t)n) DIE("short read of %s", path);
  fclose(f);
  *bytes = (size_t)n;
  return p;
}

static void *to_device(const void *host, size_t bytes) {
  void *d = NULL;
  CUDA(cudaMalloc(&d, bytes ? bytes : 8));
  CUDA(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
  return d;
}

/* meta.txt: "key value" per line */
static char g_keys[128][64];
static double g_vals[128];
static int g_nmeta = 0;
static void read_meta(void) {
  char path[4608];
  snprintf(path, sizeof path, "%s/meta.txt", g_dir);
  FILE *f = fopen(path, "r");
  if (!f) DIE("cannot open %s", path);
  while (g_nmeta < 128 && fscanf(f, "%63s %lf", g_keys[g_nmeta], &g_vals[g_nmeta]) == 2) g_nmeta++;
  fclose(f);
}
static double meta(const char *key) {
  for (int i = 0; i < g_nmeta; ++i)
    if (!strcmp(g_keys[i], key)) return g_vals[i];
  DIE("meta.txt lacks %s", key);
  return 0;
}

static double rel_l2(const double *a, const double *b, size_t n) {
  double num = 0, den = 0;
  for (size_t i = 0; i < n; ++i) {
    num += (a[i] - b[i]) * (a[i] - b[i]);
    den += b[i] * b[i];
  }
  return sqrt(num / (den > 0 ? den : 1e-300));
}

int main(int argc, char **argv) {
  if (argc < 3) DIE("usage: abi_driver <path/to/libcmdg.so> <case directory>");
  snprintf(g_dir, sizeof g_dir, "%s", argv[2]);
  void *lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) DIE("dlopen: %s", dlerror());
#define SYM(type, name) type name = (type)dlsym(lib, "cmdg_" #name); if (!name) DIE("missing symbol cmdg_%s", #name)
  SYM(version_fn, version);
  SYM(create_fn, create);
  SYM(destroy_fn, destroy);
  SYM(last_error_fn, last_error);
  SYM(bind_grid_fn, bind_grid);
  SYM(bind_state_fn, bind_state);
  SYM(tendency_fn, tendency);
  SYM(lsrk_steps_fn, lsrk_steps);
  SYM(lsrk_steps_host_fn, lsrk_steps_host);
  SYM(sync_fn, sync);
  SYM(launches_fn, kernel_launches);
#undef SYM
  if (version() != CMDG_VERSION) DIE("library version %d, header %d", version(), CMDG_VERSION);
#define CHECK(h, call) do { int rc_ = (call); if (rc_) DIE("%s -> %d: %s", #call, rc_, last_error(h)); } while (0)

  read_meta();
  const int64_t nelem = (int64_t)meta("nelem"), nreal = (int64_t)meta("nrealelem");
  const int Np = 125, Nfp = 25, nstate = (int)meta("nstate"), naux = (int)meta("naux");

  cmdg_desc d;
  memset(&d, 0, sizeof d);
  d.struct_bytes = (int32_t)sizeof d;
  d.float_bytes = CMDG_F64;
  d.dim = 3;
  d.N = 4;
  d.nelem = nelem;
  d.nrealelem = nreal;
  d.nvertelem = (int32_t)meta("nvertelem");
  d.model = CMDG_MODEL_ATMOS_DRY;
  d.nf_first = (int32_t)meta("nf_first");
  d.nf_second = d.nf_gradient = CMDG_NF_CENTRAL;
  d.orientation = (int32_t)meta("orientation");
  d.ref_state = (int32_t)meta("ref_state");
  d.subtract_off = (int32_t)meta("subtract_off");
  d.turbulence = (int32_t)meta("turbulence");
  d.turb_param = meta("turb_param");
  d.sources = (int32_t)meta("sources");
  d.diffusion_direction = (int32_t)meta("diffusion_direction");
  d.skip_zero_viscosity = (int32_t)meta("skip_zero_viscosity");
  d.write_aux_diagnostics = 1;
  d.nbc = (int32_t)meta("nbc");
  for (int i = 0; i < d.nbc; ++i) d.bc_kind[i] = CMDG_BC_FREESLIP;
  d.nstate = nstate;
  d.naux = naux;
  d.ngrad = (int32_t)meta("ngrad");
  d.ngradflux = (int32_t)meta("ngradflux");
  d.R_d = meta("R_d"); d.cp_d = meta("cp_d"); d.cv_d = meta("cv_d"); d.T_0 = meta("T_0");
  d.MSLP = meta("MSLP"); d.grav = meta("grav"); d.Omega = meta("Omega"); d.inv_Pr_turb = meta("inv_Pr_turb");
  d.day = meta("day");

  size_t b;
  void *vgeo = read_file("vgeo.bin", &b);
  if (b != (size_t)nelem * 25 * Np * 8) DIE("vgeo.bin has %zu bytes", b);
  void *d_vgeo = to_device(vgeo, b);
  void *sgeo = read_file("sgeo.bin", &b);
  void *d_sgeo = to_device(sgeo, b);
  void *vM = read_file("vmapM.bin", &b);
  if (b != (size_t)nelem * 6 * Nfp * 8) DIE("vmapM.bin has %zu bytes", b);
  void *d_vM = to_device(vM, b);
  void *vP = read_file("vmapP.bin", &b);
  void *d_vP = to_device(vP, b);
  void *bnd = read_file("elemtobndy.bin", &b);
  void *d_bnd = to_device(bnd, b);
  void *D = read_file("D.bin", &b);
  void *d_D = to_device(D, b);
  size_t qbytes, abytes;
  double *Q = (double *)read_file("Q.bin", &qbytes);
  if (qbytes != (size_t)nelem * nstate * Np * 8) DIE("Q.bin has %zu bytes", qbytes);
  void *d_Q = to_device(Q, qbytes);
  void *aux = read_file("aux.bin", &abytes);
  void *d_aux = to_device(aux, abytes);
  void *d_gf = NULL, *d_dQ = NULL;
  CUDA(cudaMalloc(&d_gf, (size_t)nelem * (d.ngradflux > 0 ? d.ngradflux : 1) * Np * 8));
  CUDA(cudaMemset(d_gf, 0, (size_t)nelem * (d.ngradflux > 0 ? d.ngradflux : 1) * Np * 8));
  CUDA(cudaMalloc(&d_dQ, qbytes));
  /* beta = 0 must not read the old tendency: prefill with NaN bytes */
  CUDA(cudaMemset(d_dQ, 0xff, qbytes));
  /* single rank: every real element is interior (1-based ids) */
  int64_t *inter = (int64_t *)malloc(sizeof(int64_t) * (size_t)nreal);
  for (int64_t e = 0; e < nreal; ++e) inter[e] = e + 1;
  void *d_inter = to_device(inter, sizeof(int64_t) * (size_t)nreal);

  cmdg_handle h = NULL;
  CHECK(NULL, create(&d, &h));
  int32_t no_rank = 0;
  int64_t no_range[2] = {0, 0};
  CHECK(h, bind_grid(h, d_vgeo, d_sgeo, (const int64_t *)d_vM, (const int64_t *)d_vP, (const int64_t *)d_bnd, d_D,
                     (const int64_t *)d_inter, nreal, (const int64_t *)d_inter, 0, (const int64_t *)d_inter, 0,
                     (const int64_t *)d_inter, 0, &no_rank, no_range, no_range, 0));
  CHECK(h, bind_state(h, d_aux, d_gf));

  /* (dg::DGModel)(tendency, Q, nothing, t, 1, 0) */
  CHECK(h, tendency(h, d_dQ, d_Q, 0.0, 1.0, 0.0, NULL));
  CHECK(h, sync(h));
  const size_t nrealvals = (size_t)nreal * nstate * Np;
  double *got = (double *)malloc(nrealvals * 8);
  CUDA(cudaMemcpy(got, d_dQ, nrealvals * 8, cudaMemcpyDeviceToHost));
  double *expect = (double *)read_file("expect_tendency.bin", &b);
  if (b != nrealvals * 8) DIE("expect_tendency.bin has %zu bytes", b);
  const double r_t = rel_l2(got, expect, nrealvals);

  /* dostep! x nsteps with the LSRK54CarpenterKennedy tableau (LowStorageRungeKuttaMethod.jl:293-327) */
  const double rka[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                         -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
  const double rkb[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                         1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                         2277821191437.0 / 14882151754819.0};
  const double rkc[5] = {0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
                         2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0};
  CUDA(cudaMemset(d_dQ, 0, qbytes));
  CHECK(h, lsrk_steps(h, d_Q, d_dQ, 0.0, meta("dt"), 5, rka, rkb, rkc, (int64_t)meta("nsteps"), NULL));
  CHECK(h, sync(h));
  CUDA(cudaMemcpy(got, d_Q, nrealvals * 8, cudaMemcpyDeviceToHost));
  free(expect);
  expect = (double *)read_file("expect_state.bin", &b);
  const double r_s = rel_l2(got, expect, nrealvals);
  /* the same steps through a HOST buffer (cmdg_lsrk_steps_host), here plain malloc'ed (pageable) memory still holding
     the initial state: one call per step as a host-side time loop does; must reproduce the resident result bit for bit */
  for (int64_t i = 0; i < (int64_t)meta("nsteps"); ++i)
    CHECK(h, lsrk_steps_host(h, Q, (double)i * meta("dt"), meta("dt"), 5, rka, rkb, rkc, 1));
  double hmax = 0.0;
  for (size_t i = 0; i < nrealvals; ++i) {
    const double dd = fabs(Q[i] - got[i]);
    if (!(dd <= hmax)) hmax = dd;      /* NaN-propagating maximum */
  }
  const long long nl = (long long)kernel_launches(h);
  CHECK(h, destroy(h));
  printf("ABI_DRIVER tendency_rel_l2=%.3e state_rel_l2=%.3e host_path_max_abs_diff=%.3e launches=%lld\n", r_t, r_s,
         hmax, nl);
  return (r_t <= 1e-12 && r_s <= 1e-12 && hmax == 0.0 && nl > 0) ? 0 : 1;
}
