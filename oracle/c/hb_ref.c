/*
 * hb_ref.c -- plain C (C99 + OpenMP) restatement of the reference's DG tendency + LSRK path for the
 * ocean HydrostaticBoussinesqModel, with the reference's own kernel schedule and array layouts.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py): the CPU twin of oracle/ocean.py +
 * oracle/dgmodel.py for BASELINE.json configs[4], validated against them in tests/test_oracle_c.py and
 * timed by bench.py as the `cpu_baseline` of the ocean workload.  The product never links or calls it.
 *
 * Schedule per tendency evaluation on one rank (DGModel.jl:85-427 with the model hooks of
 * src/Ocean/HydrostaticBoussinesq/hydrostatic_boussinesq_model.jl:637-712):
 *   update_auxiliary_state!: vertical cutoff filter on u, vertical exponential filter on theta
 *                            (Filters.jl:651-792, in place on Q)
 *   volume_gradients! H then V (DGModel_kernels.jl:934-1328; gradient_flux applied to each launch's part)
 *   dgsem_interface_gradients! faces 1..6 (:1365-1651, CentralNumericalFluxGradient + gradient boundary states)
 *   update_auxiliary_state_gradient!: w = -div_h u, upward stack integrals of (w, -alphaT theta) with
 *                            Imat * JcV, pkin <- pkin(top) - pkin, wz0 <- w(top)  (:1903-2104)
 *   volume_tendency! H then V (+ sources)     (:64-548)
 *   dgsem_interface_tendency! faces 1..6      (:588-901; Rusanov with update_penalty!, central second-order
 *                            flux, flux-based ocean boundary conditions bc_velocity.jl / bc_temperature.jl)
 *   update!                                   (LowStorageRungeKuttaMethod.jl:146-158)
 * Arrays: Q/dQ [nelem][4][Np] (u1, u2, eta, theta), aux [nelem][8][Np] (y, w, pkin, wz0, ud[2], dGu[2]),
 * gf [nelem][10][Np] (div_h u, nu grad u (3 x 2), kappa grad theta[3]), vgeo [nelem][25][Np],
 * sgeo [nelem][6][Nfp][5], vmapM/vmapP [nelem][6][Nfp] (1-based Int64), elemtobndy [nelem][6];
 * D, Imat, Fc, Fe row-major [Nq][Nq]; elements of a stack are consecutive, bottom to top.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double real;
#define NQ 5
#define NP 125
#define NFP 25
#define HS 4
#define HA 8
#define HG 5
#define HGF 10
#define REAL_PI 3.14159265358979323846

enum { VEL_NOSLIP = 1, VEL_FREESLIP = 2, VEL_PENETRABLE_FREESLIP = 3, VEL_KINEMATIC_STRESS = 4 };
enum { TEMP_INSULATING = 1, TEMP_FLUX = 2 };

typedef struct {
  double grav, rho0, ch, cz, alphaT, nuh, nuz, kappah, kappaz, kappac, f0, beta;
  double Ly, tau0, lambda_r, thetaE; /* OceanGyre: wind stress, temperature relaxation */
  int32_t bc_vel[6], bc_temp[6];     /* per boundary tag 1..6 */
  int32_t nf_first;                  /* 0 Rusanov, 1 Central */
  int32_t nvert;                     /* elements per stack */
} hb_params;

void hb_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n > 0 ? n : 1);
#else
  (void)n;
#endif
}

/* gradient_flux (hydrostatic_boussinesq_model.jl:248-298): gG[d][s], s = u1, u2, ud1, ud2, theta */
static inline void gradient_flux(const hb_params *P, real gG[3][HG], real *GF) {
  const real nu[3] = {P->nuh, P->nuh, P->nuz};
  GF[0] = gG[0][0] + gG[1][1];
  for (int c = 0; c < 2; ++c)
    for (int d = 0; d < 3; ++d) GF[1 + 3 * c + d] = -nu[d] * gG[d][c];
  const real kz = gG[2][4] < 0 ? P->kappac : P->kappaz;
  GF[7] = -P->kappah * gG[0][4];
  GF[8] = -P->kappah * gG[1][4];
  GF[9] = -kz * gG[2][4];
}

/* flux_first_order (:427-507), uncoupled: F[d][s] */
static inline void flux_first_order(const hb_params *P, const real *q, real w, real pkin, real F[3][HS]) {
  for (int d = 0; d < 3; ++d)
    for (int s = 0; s < HS; ++s) F[d][s] = 0.0;
  const real pr = P->grav * q[2], pk = P->grav * pkin;
  F[0][0] = F[0][0] + pr;
  F[1][1] = F[1][1] + pr;
  F[0][0] = F[0][0] + pk;
  F[1][1] = F[1][1] + pk;
  const real v[3] = {q[0], q[1], w};
  for (int d = 0; d < 3; ++d) F[d][3] = v[d] * q[3];
}

/* flux_second_order (:527-540) */
static inline void flux_second_order(const real *GF, real F[3][HS]) {
  for (int d = 0; d < 3; ++d) {
    F[d][0] = GF[1 + d];
    F[d][1] = GF[4 + d];
    F[d][2] = 0.0;
    F[d][3] = GF[7 + d];
  }
}

/* boundary_state! of the velocity for the first-order (fac 2) / gradient (fac 1) fluxes (bc_velocity.jl) */
static inline void boundary_velocity(int kind, int gradient, const real *n, real *u1, real *u2, real *w) {
  if (kind == VEL_NOSLIP) {
    if (gradient) {
      *u1 = 0.0; *u2 = 0.0; *w = 0.0;
    } else {
      *u1 = -*u1; *u2 = -*u2; *w = -*w;
    }
  } else if (kind == VEL_FREESLIP) {
    const real v[3] = {*u1, *u2, *w};
    const real vn = n[0] * v[0] + n[1] * v[1] + n[2] * v[2];
    const real fac = gradient ? 1.0 : 2.0;
    *u1 = v[0] - fac * vn * n[0];
    *u2 = v[1] - fac * vn * n[1];
    *w = v[2] - fac * vn * n[2];
  } /* penetrable / kinematic stress: transmissive */
}

/* kernel_apply_filter!, VerticalDirection (Filters.jl:651-792): u1, u2 with Fc, theta with Fe */
void hb_ref_filter(real *Q, const real *Fc, const real *Fe, int64_t nreal) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nreal; ++e) {
    const int states[3] = {0, 1, 3};
    for (int si = 0; si < 3; ++si) {
      const real *F = si < 2 ? Fc : Fe;
      real *q = Q + (e * HS + states[si]) * NP;
      real out[NP];
      for (int k = 0; k < NQ; ++k)
        for (int ij = 0; ij < NQ * NQ; ++ij) {
          real a = 0.0;
          for (int n = 0; n < NQ; ++n) a = a + F[k * NQ + n] * q[ij + NQ * NQ * n];
          out[ij + NQ * NQ * k] = a;
        }
      memcpy(q, out, sizeof out);
    }
  }
}

void hb_ref_volume_gradients(const hb_params *P, const real *Q, real *gf, const real *vgeo, const real *D,
                             int64_t nreal) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nreal; ++e) {
    const real *vg = vgeo + e * 25 * NP;
    const int sQ[HG] = {0, 1, -1, -1, 3};   /* gradient argument: u1, u2, ud (zero, uncoupled), theta */
    for (int k = 0; k < NQ; ++k)
      for (int j = 0; j < NQ; ++j)
        for (int i = 0; i < NQ; ++i) {
          const int ijk = i + NQ * (j + NQ * k);
          real G1[HG], G2[HG], G3[HG];
          for (int s = 0; s < HG; ++s) {
            G1[s] = G2[s] = G3[s] = 0.0;
            if (sQ[s] < 0) continue;
            const real *g = Q + (e * HS + sQ[s]) * NP;
            for (int n = 0; n < NQ; ++n) {
              G1[s] = G1[s] + D[i * NQ + n] * g[n + NQ * (j + NQ * k)];
              G2[s] = G2[s] + D[j * NQ + n] * g[i + NQ * (n + NQ * k)];
            }
            for (int n = 0; n < NQ; ++n) G3[s] = G3[s] + D[k * NQ + n] * g[i + NQ * (j + NQ * n)];
          }
          real gH[3][HG], gV[3][HG], a[HGF], b[HGF];
          for (int d = 0; d < 3; ++d)
            for (int s = 0; s < HG; ++s) {
              gH[d][s] = vg[(3 * d + 0) * NP + ijk] * G1[s] + vg[(3 * d + 1) * NP + ijk] * G2[s];
              gV[d][s] = vg[(3 * d + 2) * NP + ijk] * G3[s];
            }
          gradient_flux(P, gH, a);
          gradient_flux(P, gV, b);
          for (int c = 0; c < HGF; ++c) gf[(e * HGF + c) * NP + ijk] = a[c] + b[c];
        }
  }
}

void hb_ref_interface_gradients(const hb_params *P, const real *Q, const real *aux, real *gf, const real *sgeo,
                                const int64_t *vmapM, const int64_t *vmapP, const int64_t *elemtobndy,
                                int64_t nreal) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nreal; ++e)
    for (int f = 0; f < 6; ++f)
      for (int n = 0; n < NFP; ++n) {
        const real *sg = sgeo + ((e * 6 + f) * NFP + n) * 5;
        const real nrm[3] = {sg[0], sg[1], sg[2]};
        const real sM = sg[3], vMI = sg[4];
        const int64_t idm = vmapM[(e * 6 + f) * NFP + n] - 1;
        int64_t idp = vmapP[(e * 6 + f) * NFP + n] - 1;
        const int64_t tag = elemtobndy[e * 6 + f];
        if (tag != 0) idp = idm;
        const int64_t em = idm / NP, vm = idm % NP, ep = idp / NP, vp = idp % NP;
        real Gm[HG] = {Q[(em * HS + 0) * NP + vm], Q[(em * HS + 1) * NP + vm], 0.0, 0.0, Q[(em * HS + 3) * NP + vm]};
        real Gp[HG] = {Q[(ep * HS + 0) * NP + vp], Q[(ep * HS + 1) * NP + vp], 0.0, 0.0, Q[(ep * HS + 3) * NP + vp]};
        real Gs[HG];
        for (int s = 0; s < HG; ++s) Gs[s] = (Gp[s] + Gm[s]) / 2;
        if (tag != 0) {
          real u1 = Gm[0], u2 = Gm[1], w = aux[(em * HA + 1) * NP + vm];
          boundary_velocity(P->bc_vel[tag - 1], 1, nrm, &u1, &u2, &w);
          Gs[0] = u1; Gs[1] = u2; Gs[2] = 0.0; Gs[3] = 0.0; Gs[4] = Gm[4];
        }
        real nGs[3][HG], nGm[3][HG], a[HGF], b[HGF];
        for (int d = 0; d < 3; ++d)
          for (int s = 0; s < HG; ++s) {
            nGs[d][s] = nrm[d] * Gs[s];
            nGm[d][s] = nrm[d] * Gm[s];
          }
        gradient_flux(P, nGs, a);
        gradient_flux(P, nGm, b);
        for (int c = 0; c < HGF; ++c) {
          real *t = &gf[(em * HGF + c) * NP + vm];
          *t = *t + vMI * sM * (a[c] - b[c]);
        }
      }
}

/* update_auxiliary_state_gradient! on the real elements (whole stacks) */
void hb_ref_column(const hb_params *P, const real *Q, real *aux, const real *gf, const real *vgeo,
                   const real *Imat, int64_t nreal) {
  const int nv = P->nvert;
  const int64_t nstack = nreal / nv;
#pragma omp parallel for schedule(static)
  for (int64_t st = 0; st < nstack; ++st) {
    const int64_t e0 = st * nv;
    for (int ij = 0; ij < NQ * NQ; ++ij) {
      real cw = 0.0, cp = 0.0;
      for (int ev = 0; ev < nv; ++ev) {
        const int64_t e = e0 + ev;
        real kw[NQ], kp[NQ], lw[NQ], lp[NQ];
        for (int k = 0; k < NQ; ++k) {
          const int nn = ij + NQ * NQ * k;
          const real jc = vgeo[(e * 25 + 15) * NP + nn];
          kw[k] = (-gf[(e * HGF + 0) * NP + nn]) * jc;
          kp[k] = (-P->alphaT * Q[(e * HS + 3) * NP + nn]) * jc;
        }
        for (int k = 0; k < NQ; ++k) {
          lw[k] = cw;
          lp[k] = cp;
          for (int n = 0; n < NQ; ++n) {
            lw[k] = lw[k] + Imat[k * NQ + n] * kw[n];
            lp[k] = lp[k] + Imat[k * NQ + n] * kp[n];
          }
          aux[(e * HA + 1) * NP + ij + NQ * NQ * k] = lw[k];
          aux[(e * HA + 2) * NP + ij + NQ * NQ * k] = lp[k];
        }
        cw = lw[NQ - 1];
        cp = lp[NQ - 1];
      }
      for (int ev = 0; ev < nv; ++ev) {
        const int64_t e = e0 + ev;
        for (int k = 0; k < NQ; ++k) {
          const int nn = ij + NQ * NQ * k;
          aux[(e * HA + 2) * NP + nn] = cp - aux[(e * HA + 2) * NP + nn];
          aux[(e * HA + 3) * NP + nn] = cw;
        }
      }
    }
  }
}

void hb_ref_volume_tendency(const hb_params *P, real *dQ, const real *Q, const real *aux, const real *gf,
                            const real *vgeo, const real *D, int64_t nreal, real alpha, real beta) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nreal; ++e) {
    real Ft[3][HS][NP], src[HS][NP];
    const real *vg = vgeo + e * 25 * NP;
    for (int n = 0; n < NP; ++n) {
      real q[HS], GF[HGF], F[3][HS], F2[3][HS];
      for (int s = 0; s < HS; ++s) q[s] = Q[(e * HS + s) * NP + n];
      for (int c = 0; c < HGF; ++c) GF[c] = gf[(e * HGF + c) * NP + n];
      const real *ax = aux + e * HA * NP + n;
      flux_first_order(P, q, ax[1 * NP], ax[2 * NP], F);
      flux_second_order(GF, F2);
      for (int d = 0; d < 3; ++d)
        for (int s = 0; s < HS; ++s) F[d][s] = F[d][s] + F2[d][s];
      const real M = vg[9 * NP + n];
      for (int m = 0; m < 3; ++m) {
        const real a = vg[(0 + m) * NP + n], b = vg[(3 + m) * NP + n], c = vg[(6 + m) * NP + n];
        for (int s = 0; s < HS; ++s) Ft[m][s][n] = M * (a * F[0][s] + b * F[1][s] + c * F[2][s]);
      }
      /* source! (:559-595): Coriolis on the f-plane / beta-plane, eta driven by w(z = 0) */
      const real fcor = P->f0 + P->beta * ax[0];
      src[0][n] = -(-fcor * q[1]);
      src[1][n] = -(fcor * q[0]);
      src[2][n] = ax[3 * NP];
      src[3][n] = 0.0;
    }
    for (int k = 0; k < NQ; ++k)
      for (int j = 0; j < NQ; ++j)
        for (int i = 0; i < NQ; ++i) {
          const int ijk = i + NQ * (j + NQ * k);
          const real MI = vg[10 * NP + ijk];
          for (int s = 0; s < HS; ++s) {
            real ltH = 0.0, ltV = 0.0;
            for (int n = 0; n < NQ; ++n) {
              ltH = ltH + (MI * D[n * NQ + i]) * Ft[0][s][n + NQ * (j + NQ * k)];
              ltH = ltH + (MI * D[n * NQ + j]) * Ft[1][s][i + NQ * (n + NQ * k)];
            }
            for (int n = 0; n < NQ; ++n) {
              ltV = ltV + (MI * D[n * NQ + k]) * Ft[2][s][i + NQ * (j + NQ * n)];
              if (n == k) ltV = ltV + src[s][ijk];
            }
            real *t = &dQ[(e * HS + s) * NP + ijk];
            real d = beta != 0.0 ? alpha * ltH + beta * (*t) : alpha * ltH;   /* horizontal launch */
            *t = alpha * ltV + 1 * d;                                          /* vertical launch */
          }
        }
  }
}

void hb_ref_interface_tendency(const hb_params *P, real *dQ, const real *Q, const real *aux, const real *gf,
                               const real *sgeo, const int64_t *vmapM, const int64_t *vmapP,
                               const int64_t *elemtobndy, int64_t nreal, real alpha) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nreal; ++e)
    for (int f = 0; f < 6; ++f)
      for (int n = 0; n < NFP; ++n) {
        const real *sg = sgeo + ((e * 6 + f) * NFP + n) * 5;
        const real nrm[3] = {sg[0], sg[1], sg[2]};
        const real sM = sg[3], vMI = sg[4];
        const int64_t idm = vmapM[(e * 6 + f) * NFP + n] - 1;
        int64_t idp = vmapP[(e * 6 + f) * NFP + n] - 1;
        const int64_t tag = elemtobndy[e * 6 + f];
        if (tag != 0) idp = idm;
        const int64_t em = idm / NP, vm = idm % NP, ep = idp / NP, vp = idp % NP;
        real qm[HS], qp[HS], gm[HGF], gp[HGF];
        for (int s = 0; s < HS; ++s) {
          qm[s] = Q[(em * HS + s) * NP + vm];
          qp[s] = Q[(ep * HS + s) * NP + vp];
        }
        for (int c = 0; c < HGF; ++c) {
          gm[c] = gf[(em * HGF + c) * NP + vm];
          gp[c] = gf[(ep * HGF + c) * NP + vp];
        }
        const real ym = aux[(em * HA + 0) * NP + vm];
        const real wm = aux[(em * HA + 1) * NP + vm], pkm = aux[(em * HA + 2) * NP + vm];
        real wp = aux[(ep * HA + 1) * NP + vp];
        const real pkp = aux[(ep * HA + 2) * NP + vp];
        if (tag != 0) boundary_velocity(P->bc_vel[tag - 1], 0, nrm, &qp[0], &qp[1], &wp);
        /* first-order numerical flux (NumericalFluxes.jl:219-340) */
        real Fm[3][HS], Fp[3][HS], fl[HS];
        flux_first_order(P, qm, wm, pkm, Fm);
        flux_first_order(P, qp, wp, pkp, Fp);
        for (int s = 0; s < HS; ++s) {
          const real F0 = Fm[0][s] + Fp[0][s], F1 = Fm[1][s] + Fp[1][s], F2 = Fm[2][s] + Fp[2][s];
          fl[s] = F0 * (nrm[0] / 2) + F1 * (nrm[1] / 2) + F2 * (nrm[2] / 2);
        }
        if (P->nf_first == 0) {
          const real lam = fabs(P->ch * nrm[0] + P->ch * nrm[1] + P->cz * nrm[2]);   /* same on both sides (:602) */
          for (int s = 0; s < HS; ++s) {
            real pen = lam * (qm[s] - qp[s]);
            if (s == 2) pen = -0.0 * pen;                                            /* update_penalty! (:609-624) */
            fl[s] = fl[s] + pen / 2;
          }
        }
        /* second-order numerical flux (:668-715) or the flux-based boundary conditions (:872-967) */
        real F2[3][HS], fl2[HS];
        if (tag == 0) {
          real A[3][HS], B[3][HS];
          flux_second_order(gm, A);
          flux_second_order(gp, B);
          for (int s = 0; s < HS; ++s) {
            const real a0 = A[0][s] + B[0][s], a1 = A[1][s] + B[1][s], a2 = A[2][s] + B[2][s];
            fl2[s] = a0 * (nrm[0] / 2) + a1 * (nrm[1] / 2) + a2 * (nrm[2] / 2);
          }
        } else {
          real GFp[HGF];
          for (int c = 0; c < HGF; ++c) GFp[c] = gm[c];
          const int vel = P->bc_vel[tag - 1], temp = P->bc_temp[tag - 1];
          if (vel == VEL_FREESLIP || vel == VEL_PENETRABLE_FREESLIP) {
            for (int c = 1; c < 7; ++c) GFp[c] = 0.0;
          } else if (vel == VEL_KINEMATIC_STRESS) {
            const real st[2] = {(P->tau0 / P->rho0) * cos(ym * REAL_PI / P->Ly), 0 * ym};
            for (int c = 0; c < 2; ++c)
              for (int d = 0; d < 3; ++d) GFp[1 + 3 * c + d] = nrm[d] * st[c];
          }
          if (temp == TEMP_INSULATING) {
            for (int c = 7; c < 10; ++c) GFp[c] = 0.0;
          } else if (temp == TEMP_FLUX) {
            const real theta_r = P->thetaE * (1 - ym / P->Ly);
            const real sf = P->lambda_r * (qm[3] - theta_r);
            for (int d = 0; d < 3; ++d) GFp[7 + d] = nrm[d] * sf;
          }
          flux_second_order(GFp, F2);
          for (int s = 0; s < HS; ++s) fl2[s] = F2[0][s] * nrm[0] + F2[1][s] * nrm[1] + F2[2][s] * nrm[2];
        }
        for (int s = 0; s < HS; ++s) {
          real *t = &dQ[(em * HS + s) * NP + vm];
          *t = *t - alpha * vMI * sM * (fl[s] + fl2[s]);
        }
      }
}

/* One tendency evaluation on one rank; Q is filtered in place as the reference does. */
void hb_ref_tendency(const hb_params *P, real *dQ, real *Q, real *aux, real *gf, const real *vgeo,
                     const real *sgeo, const int64_t *vmapM, const int64_t *vmapP, const int64_t *elemtobndy,
                     const real *D, const real *Imat, const real *Fc, const real *Fe, int64_t nreal,
                     real alpha, real beta) {
  hb_ref_filter(Q, Fc, Fe, nreal);
  hb_ref_volume_gradients(P, Q, gf, vgeo, D, nreal);
  hb_ref_interface_gradients(P, Q, aux, gf, sgeo, vmapM, vmapP, elemtobndy, nreal);
  hb_ref_column(P, Q, aux, gf, vgeo, Imat, nreal);
  hb_ref_volume_tendency(P, dQ, Q, aux, gf, vgeo, D, nreal, alpha, beta);
  hb_ref_interface_tendency(P, dQ, Q, aux, gf, sgeo, vmapM, vmapP, elemtobndy, nreal, alpha);
}

/* dostep! x nsteps (LowStorageRungeKuttaMethod.jl:102-144), one rank */
void hb_ref_lsrk_steps(const hb_params *P, real *Q, real *dQ, real *aux, real *gf, const real *vgeo,
                       const real *sgeo, const int64_t *vmapM, const int64_t *vmapP, const int64_t *elemtobndy,
                       const real *D, const real *Imat, const real *Fc, const real *Fe, int64_t nreal, real dt,
                       int nstage, const real *rka, const real *rkb, int64_t nsteps) {
  const int64_t n = nreal * HS * NP;
  for (int64_t st = 0; st < nsteps; ++st)
    for (int s = 0; s < nstage; ++s) {
      hb_ref_tendency(P, dQ, Q, aux, gf, vgeo, sgeo, vmapM, vmapP, elemtobndy, D, Imat, Fc, Fe, nreal, 1.0, 1.0);
      const real a = rka[(s + 1) % nstage], b = rkb[s];
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) {
        Q[i] += b * dt * dQ[i];
        dQ[i] *= a;
      }
    }
}
