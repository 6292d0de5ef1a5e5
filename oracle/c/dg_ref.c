/*
 * dg_ref.c -- plain C (C99 + OpenMP) restatement of the reference's DG tendency + LSRK path
 * for the dry AtmosModel, with the reference's own kernel schedule and array layouts.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py): it is the CPU twin of
 * the NumPy oracle, validated against it in tests/test_oracle_c.py, and is what bench.py
 * times as `cpu_baseline` / `--impl reference` ("restated reference CPU path").  The product
 * never links or calls it.
 *
 * Schedule per tendency evaluation (src/Numerics/DGMethods/DGModel.jl:85-427 for a
 * first-order model; the nu = 0 gradient pass is skipped exactly as the GPU arm's
 * skip_zero_viscosity does):
 *   kernel_nodal_update_auxiliary_state!   DGModel_kernels.jl:1769-1825, moisture.jl:58-69
 *   volume_tendency! (horizontal launch)   DGModel_kernels.jl:64-309
 *   volume_tendency! (vertical launch)     DGModel_kernels.jl:312-548 (+ sources)
 *   dgsem_interface_tendency!              DGModel_kernels.jl:588-901 (faces 1..4, then 5..6)
 *   update!                                LowStorageRungeKuttaMethod.jl:146-158
 * Arrays: Q/dQ [nelem][5][Np], aux [nelem][A][Np], vgeo [nelem][25][Np],
 * sgeo [nelem][6][Nfp][5], vmapM/vmapP [nelem][6][Nfp] (1-based Int64), elemtobndy [nelem][6].
 */
#include <math.h>
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
} ref_params;

static inline void thermo(const ref_params *P, const double *q, double Phi, double *T, double *p) {
  double rinv = 1.0 / q[0];
  double ke = rinv * (q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) / 2;
  double e_int = rinv * (q[4] - ke - q[0] * Phi);
  *T = P->T_0 + e_int / P->cv_d;
  *p = P->R_d * q[0] * (*T);
}

static inline void flux_first_order(const ref_params *P, const double *q, double Phi, double pref,
                                    double F[3][NS]) {
  double T, p;
  thermo(P, q, Phi, &T, &p);
  double u[3] = {q[1] / q[0], q[2] / q[0], q[3] / q[0]};
  double pp = P->subtract_off ? p - pref : p;
  for (int d = 0; d < 3; ++d) {
    F[d][0] = q[1 + d];
    for (int c = 0; c < 3; ++c) F[d][1 + c] = q[1 + d] * u[c];
    F[d][1 + d] += pp;
    F[d][4] = u[d] * q[4] + u[d] * p;
  }
}

static inline double wavespeed(const ref_params *P, const double *n, const double *q, double Phi) {
  double rinv = 1.0 / q[0];
  double un = fabs(n[0] * rinv * q[1] + n[1] * rinv * q[2] + n[2] * rinv * q[3]);
  double T, p;
  thermo(P, q, Phi, &T, &p);
  return un + sqrt(P->cp_d / P->cv_d * P->R_d * T);
}

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

void ref_update_aux(const ref_params *P, const double *Q, double *aux, int64_t e0, int64_t e1) {
#pragma omp parallel for schedule(static)
  for (int64_t e = e0; e < e1; ++e)
    for (int n = 0; n < NP; ++n) {
      double q[NS];
      for (int s = 0; s < NS; ++s) q[s] = Q[(e * NS + s) * NP + n];
      double Phi = P->a_Phi >= 0 ? aux[(e * P->naux + P->a_Phi) * NP + n] : 0.0;
      double T, p;
      thermo(P, q, Phi, &T, &p);
      aux[(e * P->naux + P->a_theta_v) * NP + n] = T / pow(p / P->MSLP, P->R_d / P->cp_d);
      aux[(e * P->naux + P->a_T) * NP + n] = T;
    }
}

/* direction: 0 = horizontal launch (xi1, xi2; no source), 1 = vertical launch (xi3; + source) */
void ref_volume_tendency(const ref_params *P, int direction, double *dQ, const double *Q,
                         const double *aux, const double *vgeo, const double *D, int64_t nreal,
                         double alpha, double beta) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nreal; ++e) {
    double Ft[3][NS][NP];
    double src[NS][NP];
    const double *vg = vgeo + e * 25 * NP;
    for (int n = 0; n < NP; ++n) {
      double q[NS], F[3][NS];
      for (int s = 0; s < NS; ++s) q[s] = Q[(e * NS + s) * NP + n];
      const double *ax = aux + e * P->naux * NP + n;
      double Phi = P->a_Phi >= 0 ? ax[P->a_Phi * NP] : 0.0;
      double pref = P->a_ref_p >= 0 ? ax[P->a_ref_p * NP] : 0.0;
      flux_first_order(P, q, Phi, pref, F);
      double M = vg[9 * NP + n];
      for (int m = (direction ? 2 : 0); m < (direction ? 3 : 2); ++m) {
        double a = vg[(0 + m) * NP + n], b = vg[(3 + m) * NP + n], c = vg[(6 + m) * NP + n];
        for (int s = 0; s < NS; ++s) Ft[m][s][n] = M * (a * F[0][s] + b * F[1][s] + c * F[2][s]);
      }
      if (direction) {
        for (int s = 0; s < NS; ++s) src[s][n] = 0.0;
        if (P->gravity) {
          double rr = q[0];
          if (P->subtract_off) rr -= ax[P->a_ref_rho * NP];
          for (int d = 0; d < 3; ++d) src[1 + d][n] = -rr * ax[(P->a_gradPhi + d) * NP];
        }
        if (P->coriolis) {
          double w = 2 * P->Omega;
          src[1][n] += w * q[2];
          src[2][n] += -(w * q[1]);
        }
      }
    }
    for (int k = 0; k < NQ; ++k)
      for (int j = 0; j < NQ; ++j)
        for (int i = 0; i < NQ; ++i) {
          int ijk = i + NQ * (j + NQ * k);
          double MI = vg[10 * NP + ijk];
          for (int s = 0; s < NS; ++s) {
            double lt = 0.0;
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
            double *t = &dQ[(e * NS + s) * NP + ijk];
            *t = beta != 0.0 ? alpha * lt + beta * (*t) : alpha * lt;
          }
        }
  }
}

/* faces f0..f1-1 of the listed elements (1-based ids), as one launch of the reference kernel */
void ref_interface_tendency(const ref_params *P, double *dQ, const double *Q, const double *aux,
                            const double *sgeo, const int64_t *vmapM, const int64_t *vmapP,
                            const int64_t *elemtobndy, const int64_t *elems, int64_t nelems,
                            int f0, int f1, double alpha) {
#pragma omp parallel for schedule(static)
  for (int64_t ei = 0; ei < nelems; ++ei) {
    int64_t e = elems[ei] - 1;
    for (int f = f0; f < f1; ++f)
      for (int n = 0; n < NFP; ++n) {
        const double *sg = sgeo + ((e * 6 + f) * NFP + n) * 5;
        double nrm[3] = {sg[0], sg[1], sg[2]};
        double sM = sg[3], vMI = sg[4];
        int64_t idm = vmapM[(e * 6 + f) * NFP + n] - 1, idp = vmapP[(e * 6 + f) * NFP + n] - 1;
        int64_t bctag = elemtobndy[e * 6 + f];
        if (bctag != 0) idp = idm;
        int64_t em = idm / NP, vm = idm % NP, ep = idp / NP, vp = idp % NP;
        double qm[NS], qp[NS];
        for (int s = 0; s < NS; ++s) {
          qm[s] = Q[(em * NS + s) * NP + vm];
          qp[s] = Q[(ep * NS + s) * NP + vp];
        }
        const double *am = aux + em * P->naux * NP + vm, *ap = aux + ep * P->naux * NP + vp;
        double Phim = P->a_Phi >= 0 ? am[P->a_Phi * NP] : 0.0, Phip = P->a_Phi >= 0 ? ap[P->a_Phi * NP] : 0.0;
        double prm = P->a_ref_p >= 0 ? am[P->a_ref_p * NP] : 0.0, prp = P->a_ref_p >= 0 ? ap[P->a_ref_p * NP] : 0.0;
        if (bctag != 0) {
          double run = qm[1] * nrm[0] + qm[2] * nrm[1] + qm[3] * nrm[2];
          if (P->bc_kind[bctag - 1] == 1)
            for (int d = 0; d < 3; ++d) qp[1 + d] = qm[1 + d] - 2 * run * nrm[d];
          else
            for (int d = 0; d < 3; ++d) qp[1 + d] = -qm[1 + d];
        }
        double Fm[3][NS], Fp[3][NS], fl[NS];
        flux_first_order(P, qm, Phim, prm, Fm);
        flux_first_order(P, qp, Phip, prp, Fp);
        for (int s = 0; s < NS; ++s)
          fl[s] = (Fm[0][s] + Fp[0][s]) * (nrm[0] / 2) + (Fm[1][s] + Fp[1][s]) * (nrm[1] / 2) +
                  (Fm[2][s] + Fp[2][s]) * (nrm[2] / 2);
        if (P->nf_first == 0) {
          double lam = fmax(wavespeed(P, nrm, qm, Phim), wavespeed(P, nrm, qp, Phip));
          for (int s = 0; s < NS; ++s) fl[s] += (lam * (qm[s] - qp[s])) / 2;
        }
        for (int s = 0; s < NS; ++s) dQ[(em * NS + s) * NP + vm] -= alpha * vMI * sM * fl[s];
      }
  }
}

void ref_lsrk_update(double *dQ, double *Q, double rka, double rkb, double dt, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    Q[i] += rkb * dt * dQ[i];
    dQ[i] *= rka;
  }
}

/* One tendency evaluation, single rank (all real elements are "interior"). */
void ref_tendency(const ref_params *P, double *dQ, const double *Q, double *aux,
                  const double *vgeo, const double *sgeo, const int64_t *vmapM,
                  const int64_t *vmapP, const int64_t *elemtobndy, const double *D,
                  const int64_t *elems, int64_t nreal, double alpha, double beta) {
  ref_update_aux(P, Q, aux, 0, nreal);
  ref_volume_tendency(P, 0, dQ, Q, aux, vgeo, D, nreal, alpha, beta);
  ref_volume_tendency(P, 1, dQ, Q, aux, vgeo, D, nreal, alpha, 1.0);
  ref_interface_tendency(P, dQ, Q, aux, sgeo, vmapM, vmapP, elemtobndy, elems, nreal, 0, 4, alpha);
  ref_interface_tendency(P, dQ, Q, aux, sgeo, vmapM, vmapP, elemtobndy, elems, nreal, 4, 6, alpha);
}

/* dostep! x nsteps (LowStorageRungeKuttaMethod.jl:102-144), single rank */
void ref_lsrk_steps(const ref_params *P, double *Q, double *dQ, double *aux, const double *vgeo,
                    const double *sgeo, const int64_t *vmapM, const int64_t *vmapP,
                    const int64_t *elemtobndy, const double *D, const int64_t *elems,
                    int64_t nreal, double dt, int nstage, const double *rka, const double *rkb,
                    int64_t nsteps) {
  for (int64_t st = 0; st < nsteps; ++st)
    for (int s = 0; s < nstage; ++s) {
      ref_tendency(P, dQ, Q, aux, vgeo, sgeo, vmapM, vmapP, elemtobndy, D, elems, nreal, 1.0, 1.0);
      ref_lsrk_update(dQ, Q, rka[(s + 1) % nstage], rkb[s], dt, nreal * NS * NP);
    }
}
