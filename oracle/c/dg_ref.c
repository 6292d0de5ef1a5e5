/*
 * dg_ref.c -- plain C (C99 + OpenMP) restatement of the reference's DG tendency + LSRK path
 * for the dry AtmosModel, with the reference's own kernel schedule and array layouts.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py): it is the CPU twin of
 * the NumPy oracle, validated against it in tests/test_oracle_c.py, and is what bench.py
 * times as `cpu_baseline` / `--impl reference` ("restated reference CPU path").  The product
 * never links or calls it.
 *
 * Schedule per tendency evaluation (src/Numerics/DGMethods/DGModel.jl:85-427; with
 * second_order = 0 the nu = 0 gradient pass is skipped exactly as the GPU arm's
 * skip_zero_viscosity does):
 *   kernel_nodal_update_auxiliary_state!   DGModel_kernels.jl:1769-1825, moisture.jl:58-69
 *   volume_gradients! (H [+ V] launch)     DGModel_kernels.jl:934-1328      } second_order only
 *   dgsem_interface_gradients!             DGModel_kernels.jl:1365-1651     } (Held-Suarez config)
 *   volume_tendency! (horizontal launch)   DGModel_kernels.jl:64-309
 *   volume_tendency! (vertical launch)     DGModel_kernels.jl:312-548 (+ sources)
 *   dgsem_interface_tendency!              DGModel_kernels.jl:588-901 (faces 1..4, then 5..6)
 *   update!                                LowStorageRungeKuttaMethod.jl:146-158
 * Arrays: Q/dQ [nelem][5][Np], aux [nelem][A][Np], vgeo [nelem][25][Np],
 * sgeo [nelem][6][Nfp][5], vmapM/vmapP [nelem][6][Nfp] (1-based Int64), elemtobndy [nelem][6].
 */
#ifdef DGREF_LONG_DOUBLE
/* "truth" build: the same schedule evaluated in x87 extended precision (64-bit mantissa) on arrays of
 * long double, used to measure how far the Float64 evaluations (this twin, the NumPy oracle, the CUDA
 * kernels) are from the exactly rounded result when the tendency is a small residual of large terms */
#include <tgmath.h>
typedef long double real;
#define REAL_PI 3.14159265358979323846264338327950288L
#else
#include <math.h>
typedef double real;
#define REAL_PI 3.14159265358979323846
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NQ 5
#define NP 125
#define NFP 25
#define NS 5

typedef struct {
  double R_d, cp_d, cv_d, T_0, MSLP, grav, Omega;
  int32_t naux;
  int32_t a_Phi, a_gradPhi, a_ref_rho, a_ref_p, a_theta_v, a_T; /* -1 when absent */
  int32_t subtract_off, gravity, coriolis;
  int32_t nf_first; /* 0 Rusanov, 1 Central */
  int32_t bc_kind[6]; /* 1 free slip, 2 no slip */
  /* second-order path (gradient pass + viscous fluxes) and the Held-Suarez sources */
  int32_t second_order;         /* 0: first-order schedule only */
  int32_t turbulence;           /* 0 ConstantKinematic, 1 ConstantDynamic, 2 SmagorinskyLilly */
  int32_t with_divergence, horizontal_diffusion;
  int32_t a_Delta, ngradflux;   /* aux column of the Smagorinsky length scale; 9 or 10 */
  int32_t held_suarez, sponge;
  double turb_param, inv_Pr_turb, day;
  double sponge_z_max, sponge_z_sponge, sponge_alpha_max, sponge_gamma, sponge_u[3];
} ref_params;

#define NG 5   /* gradient variables: u[3], h_tot, theta_v (Smagorinsky only) */
#define NGF 10 /* gradient flux: grad h_tot[3], S11 S21 S31 S22 S32 S33, N2 (Smagorinsky only) */

static inline void thermo(const ref_params *P, const real *q, real Phi, real *T, real *p) {
  real rinv = 1.0 / q[0];
  real ke = rinv * (q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) / 2;
  real e_int = rinv * (q[4] - ke - q[0] * Phi);
  *T = P->T_0 + e_int / P->cv_d;
  *p = P->R_d * q[0] * (*T);
}

static inline void flux_first_order(const ref_params *P, const real *q, real Phi, real pref,
                                    real F[3][NS]) {
  real T, p;
  thermo(P, q, Phi, &T, &p);
  real u[3] = {q[1] / q[0], q[2] / q[0], q[3] / q[0]};
  real pp = P->subtract_off ? p - pref : p;
  for (int d = 0; d < 3; ++d) {
    F[d][0] = q[1 + d];
    for (int c = 0; c < 3; ++c) F[d][1 + c] = q[1 + d] * u[c];
    F[d][1 + d] += pp;
    F[d][4] = u[d] * q[4] + u[d] * p;
  }
}

static inline real wavespeed(const ref_params *P, const real *n, const real *q, real Phi) {
  real rinv = 1.0 / q[0];
  real un = fabs(n[0] * rinv * q[1] + n[1] * rinv * q[2] + n[2] * rinv * q[3]);
  real T, p;
  thermo(P, q, Phi, &T, &p);
  return un + sqrt(P->cp_d / P->cv_d * P->R_d * T);
}

/* compute_gradient_argument! (AtmosModel.jl:622-673, energy.jl, TurbulenceClosures.jl:446-454) */
static inline void gradient_argument(const ref_params *P, const real *q, real Phi, real theta_v,
                                     real *G) {
  real rinv = 1.0 / q[0];
  for (int d = 0; d < 3; ++d) G[d] = rinv * q[1 + d];
  real T, p;
  thermo(P, q, Phi, &T, &p);
  G[3] = q[4] * (1.0 / q[0]) + P->R_d * T;
  G[4] = P->turbulence == 2 ? theta_v : 0.0;
}

/* compute_gradient_flux! (AtmosModel.jl:675-744, TurbulenceClosures.jl:351-362,456-470): linear in gradG */
static inline void gradient_flux(const ref_params *P, real gG[3][NG], const real *gPhi,
                                 real theta_v, real *GF) {
  for (int d = 0; d < 3; ++d) GF[d] = gG[d][3];
  GF[3] = gG[0][0];
  GF[4] = (gG[1][0] + gG[0][1]) / 2;
  GF[5] = (gG[2][0] + gG[0][2]) / 2;
  GF[6] = gG[1][1];
  GF[7] = (gG[2][1] + gG[1][2]) / 2;
  GF[8] = gG[2][2];
  GF[9] = 0.0;
  if (P->turbulence == 2)
    GF[9] = (gG[0][4] * gPhi[0] + gG[1][4] * gPhi[1] + gG[2][4] * gPhi[2]) / theta_v;
}

/* flux_second_order! (tendencies_momentum.jl:36-43, tendencies_energy.jl:27-59) with
 * turbulence_tensors (TurbulenceClosures.jl:364-404 constant, :472-499 Smagorinsky-Lilly) */
static inline void flux_second_order(const ref_params *P, const real *q, const real *GF,
                                     const real *gPhi, real Delta, real F[3][NS]) {
  const real *S6 = GF + 3;
  real Sm[3][3] = {{S6[0], S6[1], S6[2]}, {S6[1], S6[3], S6[4]}, {S6[2], S6[4], S6[5]}};
  real nu[3], tau[3][3];
  if (P->turbulence == 2) {
    real norm2 = S6[0] * S6[0] + 2 * S6[1] * S6[1] + 2 * S6[2] * S6[2] + S6[3] * S6[3] +
                   2 * S6[4] * S6[4] + S6[5] * S6[5];
    real normS = sqrt(2 * norm2);
    real k[3] = {gPhi[0] / P->grav, gPhi[1] / P->grav, gPhi[2] / P->grav};
    double nS64 = (double)normS;
    real eps = (real)(nextafter(nS64, (double)INFINITY) - nS64); /* eps(Float64(normS)) */
    real Ri = GF[9] / (normS * normS + eps);
    real fb = 1.0 - Ri * P->inv_Pr_turb;
    fb = fb < 0.0 ? 0.0 : (fb > 1.0 ? 1.0 : fb);
    real f_b2 = sqrt(fb);
    real Cd = P->turb_param * Delta;
    real nu0 = normS * (Cd * Cd) + 1e-5;
    real dotnuk = nu0 * k[0] + nu0 * k[1] + nu0 * k[2];
    for (int i = 0; i < 3; ++i) {
      real nu_v = k[i] * dotnuk;
      nu[i] = (nu0 - nu_v) + nu_v * f_b2;
    }
  } else {
    real n0 = P->turbulence == 0 ? P->turb_param : P->turb_param / q[0];
    nu[0] = nu[1] = nu[2] = n0;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) tau[i][j] = (-2 * nu[i]) * Sm[i][j];
  if (P->turbulence != 2 && P->with_divergence) {
    real tr = S6[0] + S6[3] + S6[5];
    for (int i = 0; i < 3; ++i) tau[i][i] += (2 * nu[i] / 3) * tr;
  }
  for (int i = 0; i < 3; ++i) {
    F[i][0] = 0.0;
    for (int j = 0; j < 3; ++j) F[i][1 + j] = tau[i][j] * q[0];
    real visc = tau[i][0] * q[1] + tau[i][1] * q[2] + tau[i][2] * q[3];
    real D_t = nu[i] * P->inv_Pr_turb;
    F[i][4] = visc + ((-D_t) * GF[i]) * q[0];
  }
}

/* HeldSuarezForcing (experiments/AtmosGCM/heldsuarez.jl:112-172) and RayleighSponge
 * (tendencies_momentum.jl:104-137), added to src[1..4] in the reference's tuple order */
static inline void extended_sources(const ref_params *P, const real *q, const real *ax, int n_unused,
                                    real *src) {
  (void)n_unused;
  real Phi = ax[P->a_Phi * NP];
  if (P->held_suarez) {
    real k_a = 1.0 / (40 * P->day), k_f = 1.0 / P->day, k_s = 1.0 / (4 * P->day);
    real x0 = ax[0], x1 = ax[NP], x2 = ax[2 * NP];
    real phi = asin(x2 / sqrt(x0 * x0 + x1 * x1 + x2 * x2));
    real T, p;
    thermo(P, q, Phi, &T, &p);
    real sigma = p / P->MSLP;
    real exner = pow(sigma, P->R_d / P->cp_d);
    real dsig = (sigma - 0.7) / (1 - 0.7);
    real hf = dsig > 0 ? dsig : 0.0;
    real sp = sin(phi), cp = cos(phi);
    real T_eq = (315.0 - 60.0 * (sp * sp) - 10.0 * log(sigma) * (cp * cp)) * exner;
    T_eq = T_eq > 200.0 ? T_eq : 200.0;
    real k_T = k_a + (k_s - k_a) * hf * (cp * cp * cp * cp);
    real k_v = k_f * hf;
    real nh[3], nd = 0;
    for (int d = 0; d < 3; ++d) {
      nh[d] = ax[(P->a_gradPhi + d) * NP] / P->grav;
      nd += nh[d] * q[1 + d];
    }
    for (int d = 0; d < 3; ++d) src[1 + d] += -k_v * (q[1 + d] - nh[d] * nd);
    src[4] = -k_T * q[0] * P->cv_d * (T - T_eq);
  }
  if (P->sponge) {
    real z = Phi / P->grav;
    if (z >= P->sponge_z_sponge) {
      real r = (z - P->sponge_z_sponge) / (P->sponge_z_max - P->sponge_z_sponge);
      real beta = P->sponge_alpha_max * pow(sin(REAL_PI * (r / 2)), P->sponge_gamma);
      for (int d = 0; d < 3; ++d) src[1 + d] += -beta * (q[1 + d] - q[0] * P->sponge_u[d]);
    }
  }
}

int ref_real_bytes(void) { return (int)sizeof(real); }

int ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Launchers such as torchrun export OMP_NUM_THREADS=1; the CPU arm of bench.py asks for every core it
 * may run on instead. */
void ref_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void ref_update_aux(const ref_params *P, const real *Q, real *aux, int64_t e0, int64_t e1) {
#pragma omp parallel for schedule(static)
  for (int64_t e = e0; e < e1; ++e)
    for (int n = 0; n < NP; ++n) {
      real q[NS];
      for (int s = 0; s < NS; ++s) q[s] = Q[(e * NS + s) * NP + n];
      real Phi = P->a_Phi >= 0 ? aux[(e * P->naux + P->a_Phi) * NP + n] : 0.0;
      real T, p;
      thermo(P, q, Phi, &T, &p);
      aux[(e * P->naux + P->a_theta_v) * NP + n] = T / pow(p / P->MSLP, P->R_d / P->cp_d);
      aux[(e * P->naux + P->a_T) * NP + n] = T;
    }
}

/* direction: 0 = horizontal launch (xi1, xi2; no source), 1 = vertical launch (xi3; + source) */
void ref_volume_tendency(const ref_params *P, int direction, real *dQ, const real *Q,
                         const real *aux, const real *gf, const real *vgeo, const real *D,
                         int64_t nreal, real alpha, real beta) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nreal; ++e) {
    real Ft[3][NS][NP];
    real src[NS][NP];
    const real *vg = vgeo + e * 25 * NP;
    for (int n = 0; n < NP; ++n) {
      real q[NS], F[3][NS];
      for (int s = 0; s < NS; ++s) q[s] = Q[(e * NS + s) * NP + n];
      const real *ax = aux + e * P->naux * NP + n;
      real Phi = P->a_Phi >= 0 ? ax[P->a_Phi * NP] : 0.0;
      real pref = P->a_ref_p >= 0 ? ax[P->a_ref_p * NP] : 0.0;
      flux_first_order(P, q, Phi, pref, F);
      if (P->second_order) {
        real GF[NGF], F2[3][NS], gPhi[3] = {0, 0, 0};
        for (int c = 0; c < P->ngradflux; ++c) GF[c] = gf[(e * P->ngradflux + c) * NP + n];
        if (P->a_gradPhi >= 0)
          for (int d = 0; d < 3; ++d) gPhi[d] = ax[(P->a_gradPhi + d) * NP];
        flux_second_order(P, q, GF, gPhi, P->a_Delta >= 0 ? ax[P->a_Delta * NP] : 0.0, F2);
        for (int d = 0; d < 3; ++d)
          for (int s = 0; s < NS; ++s) F[d][s] += F2[d][s];
      }
      real M = vg[9 * NP + n];
      for (int m = (direction ? 2 : 0); m < (direction ? 3 : 2); ++m) {
        real a = vg[(0 + m) * NP + n], b = vg[(3 + m) * NP + n], c = vg[(6 + m) * NP + n];
        for (int s = 0; s < NS; ++s) Ft[m][s][n] = M * (a * F[0][s] + b * F[1][s] + c * F[2][s]);
      }
      if (direction) {
        for (int s = 0; s < NS; ++s) src[s][n] = 0.0;
        if (P->gravity) {
          real rr = q[0];
          if (P->subtract_off) rr -= ax[P->a_ref_rho * NP];
          for (int d = 0; d < 3; ++d) src[1 + d][n] = -rr * ax[(P->a_gradPhi + d) * NP];
        }
        if (P->coriolis) {
          real w = 2 * P->Omega;
          src[1][n] += w * q[2];
          src[2][n] += -(w * q[1]);
        }
        if (P->held_suarez || P->sponge) {
          real sn[NS];
          for (int s = 0; s < NS; ++s) sn[s] = src[s][n];
          extended_sources(P, q, ax, n, sn);
          for (int s = 0; s < NS; ++s) src[s][n] = sn[s];
        }
      }
    }
    for (int k = 0; k < NQ; ++k)
      for (int j = 0; j < NQ; ++j)
        for (int i = 0; i < NQ; ++i) {
          int ijk = i + NQ * (j + NQ * k);
          real MI = vg[10 * NP + ijk];
          for (int s = 0; s < NS; ++s) {
            real lt = 0.0;
            if (!direction) {
              for (int n = 0; n < NQ; ++n) {
                lt += MI * D[n * NQ + i] * Ft[0][s][n + NQ * (j + NQ * k)];
                lt += MI * D[n * NQ + j] * Ft[1][s][i + NQ * (n + NQ * k)];
              }
            } else {
              for (int n = 0; n < NQ; ++n) {
                lt += MI * D[n * NQ + k] * Ft[2][s][i + NQ * (j + NQ * n)];
                if (n == k) lt += src[s][ijk];
              }
            }
            real *t = &dQ[(e * NS + s) * NP + ijk];
            *t = beta != 0.0 ? alpha * lt + beta * (*t) : alpha * lt;
          }
        }
  }
}

/* faces f0..f1-1 of the listed elements (1-based ids), as one launch of the reference kernel */
void ref_interface_tendency(const ref_params *P, real *dQ, const real *Q, const real *aux,
                            const real *gf, const real *sgeo, const int64_t *vmapM, const int64_t *vmapP,
                            const int64_t *elemtobndy, const int64_t *elems, int64_t nelems,
                            int f0, int f1, real alpha) {
#pragma omp parallel for schedule(static)
  for (int64_t ei = 0; ei < nelems; ++ei) {
    int64_t e = elems[ei] - 1;
    for (int f = f0; f < f1; ++f)
      for (int n = 0; n < NFP; ++n) {
        const real *sg = sgeo + ((e * 6 + f) * NFP + n) * 5;
        real nrm[3] = {sg[0], sg[1], sg[2]};
        real sM = sg[3], vMI = sg[4];
        int64_t idm = vmapM[(e * 6 + f) * NFP + n] - 1, idp = vmapP[(e * 6 + f) * NFP + n] - 1;
        int64_t bctag = elemtobndy[e * 6 + f];
        if (bctag != 0) idp = idm;
        int64_t em = idm / NP, vm = idm % NP, ep = idp / NP, vp = idp % NP;
        real qm[NS], qp[NS];
        for (int s = 0; s < NS; ++s) {
          qm[s] = Q[(em * NS + s) * NP + vm];
          qp[s] = Q[(ep * NS + s) * NP + vp];
        }
        const real *am = aux + em * P->naux * NP + vm, *ap = aux + ep * P->naux * NP + vp;
        real Phim = P->a_Phi >= 0 ? am[P->a_Phi * NP] : 0.0, Phip = P->a_Phi >= 0 ? ap[P->a_Phi * NP] : 0.0;
        real prm = P->a_ref_p >= 0 ? am[P->a_ref_p * NP] : 0.0, prp = P->a_ref_p >= 0 ? ap[P->a_ref_p * NP] : 0.0;
        if (bctag != 0) {
          real run = qm[1] * nrm[0] + qm[2] * nrm[1] + qm[3] * nrm[2];
          if (P->bc_kind[bctag - 1] == 1)
            for (int d = 0; d < 3; ++d) qp[1 + d] = qm[1 + d] - 2 * run * nrm[d];
          else
            for (int d = 0; d < 3; ++d) qp[1 + d] = -qm[1 + d];
        }
        real qp2[NS];   /* the un-modified + state (second-order flux) */
        for (int s = 0; s < NS; ++s) qp2[s] = Q[(ep * NS + s) * NP + vp];
        real Fm[3][NS], Fp[3][NS], fl[NS];
        flux_first_order(P, qm, Phim, prm, Fm);
        flux_first_order(P, qp, Phip, prp, Fp);
        for (int s = 0; s < NS; ++s)
          fl[s] = (Fm[0][s] + Fp[0][s]) * (nrm[0] / 2) + (Fm[1][s] + Fp[1][s]) * (nrm[1] / 2) +
                  (Fm[2][s] + Fp[2][s]) * (nrm[2] / 2);
        if (P->nf_first == 0) {
          real lam = fmax(wavespeed(P, nrm, qm, Phim), wavespeed(P, nrm, qp, Phip));
          for (int s = 0; s < NS; ++s) fl[s] += (lam * (qm[s] - qp[s])) / 2;
        }
        if (P->second_order && bctag == 0) {
          /* CentralNumericalFluxSecondOrder (NumericalFluxes.jl:668-715); AtmosBC walls: no diffusive flux */
          real Gm[NGF], Gp[NGF], F2m[3][NS], F2p[3][NS], gpm[3] = {0, 0, 0}, gpp[3] = {0, 0, 0};
          for (int c = 0; c < P->ngradflux; ++c) {
            Gm[c] = gf[(em * P->ngradflux + c) * NP + vm];
            Gp[c] = gf[(ep * P->ngradflux + c) * NP + vp];
          }
          if (P->a_gradPhi >= 0)
            for (int d = 0; d < 3; ++d) {
              gpm[d] = am[(P->a_gradPhi + d) * NP];
              gpp[d] = ap[(P->a_gradPhi + d) * NP];
            }
          flux_second_order(P, qm, Gm, gpm, P->a_Delta >= 0 ? am[P->a_Delta * NP] : 0.0, F2m);
          flux_second_order(P, qp2, Gp, gpp, P->a_Delta >= 0 ? ap[P->a_Delta * NP] : 0.0, F2p);
          for (int s = 0; s < NS; ++s)
            fl[s] += (F2m[0][s] + F2p[0][s]) * (nrm[0] / 2) + (F2m[1][s] + F2p[1][s]) * (nrm[1] / 2) +
                     (F2m[2][s] + F2p[2][s]) * (nrm[2] / 2);
        }
        for (int s = 0; s < NS; ++s) dQ[(em * NS + s) * NP + vm] -= alpha * vMI * sM * fl[s];
      }
  }
}

/* volume_gradients! H launch (+ V launch unless diffusion_direction = HorizontalDirection):
 * GF = gf(xi_x D G) on real elements */
void ref_volume_gradients(const ref_params *P, const real *Q, const real *aux, real *gf,
                          const real *vgeo, const real *D, int64_t nreal) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nreal; ++e) {
    real G[NG][NP];
    const real *vg = vgeo + e * 25 * NP;
    for (int n = 0; n < NP; ++n) {
      real q[NS], g[NG];
      for (int s = 0; s < NS; ++s) q[s] = Q[(e * NS + s) * NP + n];
      const real *ax = aux + e * P->naux * NP + n;
      gradient_argument(P, q, P->a_Phi >= 0 ? ax[P->a_Phi * NP] : 0.0, ax[P->a_theta_v * NP], g);
      for (int c = 0; c < NG; ++c) G[c][n] = g[c];
    }
    for (int k = 0; k < NQ; ++k)
      for (int j = 0; j < NQ; ++j)
        for (int i = 0; i < NQ; ++i) {
          int ijk = i + NQ * (j + NQ * k);
          real G1[NG], G2[NG], G3[NG];
          for (int c = 0; c < NG; ++c) {
            G1[c] = G2[c] = G3[c] = 0.0;
            for (int n = 0; n < NQ; ++n) {
              G1[c] += D[i * NQ + n] * G[c][n + NQ * (j + NQ * k)];
              G2[c] += D[j * NQ + n] * G[c][i + NQ * (n + NQ * k)];
              G3[c] += D[k * NQ + n] * G[c][i + NQ * (j + NQ * n)];
            }
          }
          const real *ax = aux + e * P->naux * NP + ijk;
          real gPhi[3] = {0, 0, 0};
          if (P->a_gradPhi >= 0)
            for (int d = 0; d < 3; ++d) gPhi[d] = ax[(P->a_gradPhi + d) * NP];
          real th = ax[P->a_theta_v * NP];
          real gH[3][NG], gV[3][NG], GFh[NGF], GFv[NGF];
          for (int d = 0; d < 3; ++d)
            for (int c = 0; c < NG; ++c) {
              gH[d][c] = vg[(3 * d + 0) * NP + ijk] * G1[c] + vg[(3 * d + 1) * NP + ijk] * G2[c];
              gV[d][c] = vg[(3 * d + 2) * NP + ijk] * G3[c];
            }
          gradient_flux(P, gH, gPhi, th, GFh);
          if (!P->horizontal_diffusion) {
            gradient_flux(P, gV, gPhi, th, GFv);
            for (int c = 0; c < NGF; ++c) GFh[c] = GFh[c] + GFv[c];
          }
          for (int c = 0; c < P->ngradflux; ++c) gf[(e * P->ngradflux + c) * NP + ijk] = GFh[c];
        }
  }
}

/* dgsem_interface_gradients! with CentralNumericalFluxGradient (NumericalFluxes.jl:65-123) */
void ref_interface_gradients(const ref_params *P, const real *Q, const real *aux, real *gf,
                             const real *sgeo, const int64_t *vmapM, const int64_t *vmapP,
                             const int64_t *elemtobndy, const int64_t *elems, int64_t nelems) {
  const int nfaces = P->horizontal_diffusion ? 4 : 6;
#pragma omp parallel for schedule(static)
  for (int64_t ei = 0; ei < nelems; ++ei) {
    int64_t e = elems[ei] - 1;
    for (int f = 0; f < nfaces; ++f)
      for (int n = 0; n < NFP; ++n) {
        const real *sg = sgeo + ((e * 6 + f) * NFP + n) * 5;
        real nrm[3] = {sg[0], sg[1], sg[2]};
        real sM = sg[3], vMI = sg[4];
        int64_t idm = vmapM[(e * 6 + f) * NFP + n] - 1, idp = vmapP[(e * 6 + f) * NFP + n] - 1;
        int64_t bctag = elemtobndy[e * 6 + f];
        if (bctag != 0) idp = idm;
        int64_t em = idm / NP, vm = idm % NP, ep = idp / NP, vp = idp % NP;
        real qm[NS], qp[NS];
        for (int s = 0; s < NS; ++s) {
          qm[s] = Q[(em * NS + s) * NP + vm];
          qp[s] = Q[(ep * NS + s) * NP + vp];
        }
        const real *am = aux + em * P->naux * NP + vm, *ap = aux + ep * P->naux * NP + vp;
        real Phim = P->a_Phi >= 0 ? am[P->a_Phi * NP] : 0.0, Phip = P->a_Phi >= 0 ? ap[P->a_Phi * NP] : 0.0;
        real thm = am[P->a_theta_v * NP], thp = ap[P->a_theta_v * NP];
        real Gm[NG], Gs[NG];
        gradient_argument(P, qm, Phim, thm, Gm);
        if (bctag == 0) {
          real Gp[NG];
          gradient_argument(P, qp, Phip, thp, Gp);
          for (int c = 0; c < NG; ++c) Gs[c] = (Gp[c] + Gm[c]) / 2;
        } else {
          /* boundary_state!(::CentralNumericalFluxGradient) (bc_momentum.jl:34-43, 71-80), then the
           * nodal auxiliary update of the ghost state */
          real run = qm[1] * nrm[0] + qm[2] * nrm[1] + qm[3] * nrm[2];
          for (int s = 0; s < NS; ++s) qp[s] = qm[s];
          if (P->bc_kind[bctag - 1] == 1)
            for (int d = 0; d < 3; ++d) qp[1 + d] = qm[1 + d] - run * nrm[d];
          else
            for (int d = 0; d < 3; ++d) qp[1 + d] = 0 * qm[1 + d];
          real T, p;
          thermo(P, qp, Phim, &T, &p);
          gradient_argument(P, qp, Phim, T / pow(p / P->MSLP, P->R_d / P->cp_d), Gs);
        }
        real nGs[3][NG], nGm[3][NG], gfs[NGF], gfm[NGF], gPhi[3] = {0, 0, 0};
        for (int d = 0; d < 3; ++d)
          for (int c = 0; c < NG; ++c) {
            nGs[d][c] = nrm[d] * Gs[c];
            nGm[d][c] = nrm[d] * Gm[c];
          }
        if (P->a_gradPhi >= 0)
          for (int d = 0; d < 3; ++d) gPhi[d] = am[(P->a_gradPhi + d) * NP];
        gradient_flux(P, nGs, gPhi, thm, gfs);
        gradient_flux(P, nGm, gPhi, thm, gfm);
        for (int c = 0; c < P->ngradflux; ++c)
          gf[(em * P->ngradflux + c) * NP + vm] += vMI * sM * (gfs[c] - gfm[c]);
      }
  }
}

void ref_lsrk_update(real *dQ, real *Q, real rka, real rkb, real dt, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    Q[i] += rkb * dt * dQ[i];
    dQ[i] *= rka;
  }
}

/* One tendency evaluation, single rank (all real elements are "interior"). */
void ref_tendency(const ref_params *P, real *dQ, const real *Q, real *aux, real *gf,
                  const real *vgeo, const real *sgeo, const int64_t *vmapM,
                  const int64_t *vmapP, const int64_t *elemtobndy, const real *D,
                  const int64_t *elems, int64_t nreal, real alpha, real beta) {
  ref_update_aux(P, Q, aux, 0, nreal);
  if (P->second_order) {
    ref_volume_gradients(P, Q, aux, gf, vgeo, D, nreal);
    ref_interface_gradients(P, Q, aux, gf, sgeo, vmapM, vmapP, elemtobndy, elems, nreal);
  }
  ref_volume_tendency(P, 0, dQ, Q, aux, gf, vgeo, D, nreal, alpha, beta);
  ref_volume_tendency(P, 1, dQ, Q, aux, gf, vgeo, D, nreal, alpha, 1.0);
  ref_interface_tendency(P, dQ, Q, aux, gf, sgeo, vmapM, vmapP, elemtobndy, elems, nreal, 0, 4, alpha);
  ref_interface_tendency(P, dQ, Q, aux, gf, sgeo, vmapM, vmapP, elemtobndy, elems, nreal, 4, 6, alpha);
}

/* dostep! x nsteps (LowStorageRungeKuttaMethod.jl:102-144), single rank */
void ref_lsrk_steps(const ref_params *P, real *Q, real *dQ, real *aux, real *gf, const real *vgeo,
                    const real *sgeo, const int64_t *vmapM, const int64_t *vmapP,
                    const int64_t *elemtobndy, const real *D, const int64_t *elems,
                    int64_t nreal, real dt, int nstage, const real *rka, const real *rkb,
                    int64_t nsteps) {
  for (int64_t st = 0; st < nsteps; ++st)
    for (int s = 0; s < nstage; ++s) {
      ref_tendency(P, dQ, Q, aux, gf, vgeo, sgeo, vmapM, vmapP, elemtobndy, D, elems, nreal, 1.0, 1.0);
      ref_lsrk_update(dQ, Q, rka[(s + 1) % nstage], rkb[s], dt, nreal * NS * NP);
    }
}
