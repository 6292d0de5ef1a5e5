// cmdg_tracers.cuh -- passive tracers of the dry AtmosModel (NTracers{N}, src/Atmos/Model/tracers.jl,
// tendencies_tracers.jl): N <= CMDG_MAX_TRACERS advected scalars rho*chi_i after rho*e in Q, with the
// turbulent diffusion  F2[d][i] = -rho D_t[d] delta_chi[i] d(chi_i)/dx_d,  D_t = nu / Pr_t.
//
// The tracers do not feed back on the flow, so the five dynamic states keep their fused kernels
// (dg_gradient_kernel / dg_tendency_kernel, which only learn the column stride of Q) and the tracer
// columns get two kernels of their own, one block per element, one thread per node:
//
//   tracer_gradient_kernel : chi = rho*chi / rho, strong-form gradient + central face term
//                            (volume_gradients! / dgsem_interface_gradients! restricted to the tracer
//                            variables) -> grad chi columns of state_gradient_flux and the diffusive
//                            flux F2chi per node (nu from the dynamics' gradient kernel)
//   tracer_tendency_kernel : advective + diffusive flux, weak divergence, central + Rusanov face flux
//                            with the tracers' own wave speed |u.n| (wavespeed_tracers!, tracers.jl:165-180),
//                            ImpermeableTracer walls (bc_tracer.jl), alpha/beta combination and the
//                            fused RK stage update of the tracer columns.
//
// These are straightforward kernels (per-node contractions from shared memory, neighbour traces read
// through L1/L2): the tracer columns are a side path of configs[0], not the roofline kernel.
#pragma once
#include "cmdg_kernels.cuh"

#ifndef CMDG_MAX_TRACERS
#define CMDG_MAX_TRACERS 4
#endif

namespace cmdg {

template <class R>
struct TracerArgs {
  const R *Q;          // [nelem][nstate][Np]
  R *dQ, *Qout;        // tracer columns 5.. are written (Qout NULL = no stage update)
  R *gradflux;         // grad chi columns written at ngf_dyn + d + 3 i (NULL = don't)
  const R *Nu;         // [nreal][3][Np] from dg_gradient_kernel (NULL: constant closures, computed here)
  R *F2chi;            // [nelem][3 nt][Np]  column d + 3 i (ghost part by the halo exchange)
  const R *vgeoP, *sgeoP;
  const int2 *conn;
  const int *elems;
  int nt;
  R delta[CMDG_MAX_TRACERS];
  R alpha, beta, rkb_dt;
  int visc;            // diffusive flux present
  int rusanov;         // 1: Rusanov penalty, 0: central only
};

template <class R, int NQ>
__global__ void __launch_bounds__(Dims<NQ>::BLOCK)
tracer_gradient_kernel(const TracerArgs<R> A, const AtmosParams<R> P) {
  constexpr int NP = Dims<NQ>::NP, NFP = Dims<NQ>::NFP, NFN = Dims<NQ>::NFN, BLOCK = Dims<NQ>::BLOCK;
  constexpr int NT = CMDG_MAX_TRACERS;
  __shared__ R sChi[NT][NP];
  __shared__ R sW[NT][NFN];
  const int tid = threadIdx.x;
  const int e = A.elems ? A.elems[blockIdx.x] : blockIdx.x;
  const size_t eoffQ = (size_t)e * P.nstate * NP;
  const int nfaces = P.horizontal_diffusion ? 4 : 6;
  R rho = 1;
  if (tid < NP) {
    rho = A.Q[eoffQ + tid];
    const R rinv = R(1) / rho;
#pragma unroll
    for (int i = 0; i < NT; ++i)
      sChi[i][tid] = (i < A.nt) ? A.Q[eoffQ + (size_t)(5 + i) * NP + tid] * rinv : R(0);
  }
  __syncthreads();
  // faces: w_i = vMI sM (chi* - chi-), chi* = (chi+ + chi-)/2; walls: chi+ = chi- (ImpermeableTracer)
  for (int it = tid; it < nfaces * NFP; it += BLOCK) {
    const int f = it / NFP, fn = it - f * NFP;
    const int2 c = A.conn[(size_t)e * 6 + f];
    const int vm = face_to_vol<NQ>(f, fn % NQ, fn / NQ);
    R n[3], sMvMI;
    load_sgeo<R>(A.sgeoP + ((size_t)e * NFN + it) * 4, n, sMvMI);
    if (((c.y >> 4) & 15) == 0) {
      int a = fn % NQ;
      const int b = fn / NQ;
      if (c.y & 8) a = NQ - 1 - a;
      const int vp = face_to_vol<NQ>(c.y & 7, a, b);
      const R *qp = A.Q + (size_t)c.x * P.nstate * NP + vp;
      const R rinvp = R(1) / qp[0];
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        const R chim = sChi[i][vm];
        const R chip = (i < A.nt) ? qp[(size_t)(5 + i) * NP] * rinvp : R(0);
        sW[i][it] = sMvMI * (R(0.5) * (chip + chim) - chim);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NT; ++i) sW[i][it] = R(0);
    }
  }
  // volume: strong-form gradient xi_x (D chi)
  R dchi[3][NT];
  const int i_ = tid % NQ, j_ = (tid / NQ) % NQ, k_ = tid / (NQ * NQ);
  if (tid < NP) {
    R g[9], MI;
    load_vgeo<R>(A.vgeoP, e, tid, NP, g, MI);
#pragma unroll
    for (int c = 0; c < 9; ++c) g[c] *= MI;   // packed copy holds M * xi_x
    R G1[NT], G2[NT], G3[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) G1[i] = G2[i] = G3[i] = R(0);
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
      const R d1 = const_D<R>(i_ * NQ + n), d2 = const_D<R>(j_ * NQ + n), d3 = const_D<R>(k_ * NQ + n);
      const int o1 = n + NQ * (j_ + NQ * k_), o2 = i_ + NQ * (n + NQ * k_), o3 = i_ + NQ * (j_ + NQ * n);
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        G1[i] += d1 * sChi[i][o1];
        G2[i] += d2 * sChi[i][o2];
        G3[i] += d3 * sChi[i][o3];
      }
    }
    const R vfac = P.horizontal_diffusion ? R(0) : R(1);
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int i = 0; i < NT; ++i) dchi[d][i] = g[d] * G1[i] + g[3 + d] * G2[i] + vfac * (g[6 + d] * G3[i]);
  }
  __syncthreads();
  if (tid >= NP) return;
  const int fit[3] = {(i_ == 0) ? j_ + NQ * k_ : ((i_ == NQ - 1) ? NFP + j_ + NQ * k_ : -1),
                      (j_ == 0) ? 2 * NFP + i_ + NQ * k_ : ((j_ == NQ - 1) ? 3 * NFP + i_ + NQ * k_ : -1),
                      P.horizontal_diffusion ? -1
                                             : ((k_ == 0) ? 4 * NFP + i_ + NQ * j_
                                                          : ((k_ == NQ - 1) ? 5 * NFP + i_ + NQ * j_ : -1))};
#pragma unroll
  for (int dir = 0; dir < 3; ++dir) {
    const int it = fit[dir];
    if (it < 0) continue;
    R n[3], sMvMI;
    load_sgeo<R>(A.sgeoP + ((size_t)e * NFN + it) * 4, n, sMvMI);
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const R w = sW[i][it];
#pragma unroll
      for (int d = 0; d < 3; ++d) dchi[d][i] += n[d] * w;
    }
  }
  if (A.gradflux) {
    const size_t eoffG = ((size_t)e * P.ngradflux + P.ngf_dyn) * NP + tid;
#pragma unroll
    for (int i = 0; i < NT; ++i)
      if (i < A.nt) {
#pragma unroll
        for (int d = 0; d < 3; ++d) A.gradflux[eoffG + (size_t)(d + 3 * i) * NP] = dchi[d][i];
      }
  }
  // diffusive flux  -rho D_t[d] delta_i d(chi_i)/dx_d,  D_t = nu * inv_Pr_turb
  R Dt[3];
  if (A.Nu) {
#pragma unroll
    for (int d = 0; d < 3; ++d) Dt[d] = A.Nu[((size_t)e * 3 + d) * NP + tid] * P.inv_Pr_turb;
  } else {
    const R nuc = (P.turbulence == TURB_CONST_KINEMATIC) ? P.turb_param : P.turb_param / rho;
    Dt[0] = Dt[1] = Dt[2] = nuc * P.inv_Pr_turb;
  }
  const size_t eoffF = (size_t)e * 3 * A.nt * NP + tid;
#pragma unroll
  for (int i = 0; i < NT; ++i)
    if (i < A.nt) {
#pragma unroll
      for (int d = 0; d < 3; ++d)
        A.F2chi[eoffF + (size_t)(d + 3 * i) * NP] = (((-Dt[d]) * A.delta[i]) * dchi[d][i]) * rho;
    }
}

template <class R, int NQ>
__global__ void __launch_bounds__(Dims<NQ>::BLOCK)
tracer_tendency_kernel(const TracerArgs<R> A, const AtmosParams<R> P) {
  constexpr int NP = Dims<NQ>::NP, NFP = Dims<NQ>::NFP, NFN = Dims<NQ>::NFN, BLOCK = Dims<NQ>::BLOCK;
  constexpr int NT = CMDG_MAX_TRACERS;
  __shared__ R sQ[4][NP];            // rho, rho*u of my nodes
  __shared__ R sX[NT][NP];           // rho*chi of my nodes
  __shared__ R sF[3][NT][NP];        // contravariant fluxes M xi_m . F
  __shared__ R sFace[NT][NFN];       // vMI sM F* at my face nodes
  const int tid = threadIdx.x;
  const int e = A.elems ? A.elems[blockIdx.x] : blockIdx.x;
  const size_t eoffQ = (size_t)e * P.nstate * NP;
  const size_t eoffF = (size_t)e * 3 * A.nt * NP;
  R x[NT], MI = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i) x[i] = R(0);
  if (tid < NP) {
    R q[4], g[9];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      q[s] = A.Q[eoffQ + (size_t)s * NP + tid];
      sQ[s][tid] = q[s];
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      if (i < A.nt) x[i] = A.Q[eoffQ + (size_t)(5 + i) * NP + tid];
      sX[i][tid] = x[i];
    }
    load_vgeo<R>(A.vgeoP, e, tid, NP, g, MI);
    const R rinv = R(1) / q[0];
    const R u[3] = {q[1] * rinv, q[2] * rinv, q[3] * rinv};
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      R F[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        F[d] = x[i] * u[d];                          // flux(::Tracers, ::Advect)
        if (A.visc && i < A.nt) F[d] += A.F2chi[eoffF + (size_t)(d + 3 * i) * NP + tid];
      }
#pragma unroll
      for (int m = 0; m < 3; ++m) sF[m][i][tid] = g[3 * m] * F[0] + g[3 * m + 1] * F[1] + g[3 * m + 2] * F[2];
    }
  }
  __syncthreads();
  // faces
  for (int it = tid; it < NFN; it += BLOCK) {
    const int f = it / NFP, fn = it - f * NFP;
    const int2 c = A.conn[(size_t)e * 6 + f];
    const int bctag = (c.y >> 4) & 15;
    const int vm = face_to_vol<NQ>(f, fn % NQ, fn / NQ);
    R n[3], sMvMI;
    load_sgeo<R>(A.sgeoP + ((size_t)e * NFN + it) * 4, n, sMvMI);
    const R rhom = sQ[0][vm];
    const R qm[3] = {sQ[1][vm], sQ[2][vm], sQ[3][vm]};
    const R unm = (R(1) / rhom) * (qm[0] * n[0] + qm[1] * n[1] + qm[2] * n[2]);
    R unp, xp[NT], f2n[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) f2n[i] = R(0);
    if (bctag == 0) {
      int a = fn % NQ;
      const int b = fn / NQ;
      if (c.y & 8) a = NQ - 1 - a;
      const int vp = face_to_vol<NQ>(c.y & 7, a, b);
      const R *qp = A.Q + (size_t)c.x * P.nstate * NP + vp;
      unp = (R(1) / qp[0]) * (qp[(size_t)NP] * n[0] + qp[(size_t)2 * NP] * n[1] + qp[(size_t)3 * NP] * n[2]);
#pragma unroll
      for (int i = 0; i < NT; ++i) xp[i] = (i < A.nt) ? qp[(size_t)(5 + i) * NP] : R(0);
      if (A.visc) {
        // CentralNumericalFluxSecondOrder: n . (F2- + F2+) / 2, both sides from the per-node array
        const R *fm = A.F2chi + eoffF + vm;
        const R *fp = A.F2chi + (size_t)c.x * 3 * A.nt * NP + vp;
#pragma unroll
        for (int i = 0; i < NT; ++i)
          if (i < A.nt) {
            R sm = R(0);
#pragma unroll
            for (int d = 0; d < 3; ++d)
              sm += (fm[(size_t)(d + 3 * i) * NP] + fp[(size_t)(d + 3 * i) * NP]) * (n[d] * R(0.5));
            f2n[i] = sm;
          }
      }
    } else {
      // wall: the momentum of the ghost state is reflected (free slip) or reversed (no slip), rho and the
      // tracers are copied (ImpermeableTracer, bc_tracer.jl:9-15); no diffusive flux through the wall
      R qpw[3] = {qm[0], qm[1], qm[2]};
      if (P.bc_kind[bctag - 1] == BC_FREESLIP) {
        const R run = 2 * (qm[0] * n[0] + qm[1] * n[1] + qm[2] * n[2]);
        qpw[0] -= run * n[0];
        qpw[1] -= run * n[1];
        qpw[2] -= run * n[2];
      } else {
        qpw[0] = -qm[0];
        qpw[1] = -qm[1];
        qpw[2] = -qm[2];
      }
      unp = (R(1) / rhom) * (qpw[0] * n[0] + qpw[1] * n[1] + qpw[2] * n[2]);
#pragma unroll
      for (int i = 0; i < NT; ++i) xp[i] = sX[i][vm];
    }
    const R lam = fmax(fabs(unm), fabs(unp));
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const R xm = sX[i][vm];
      R fl = R(0.5) * (xm * unm + xp[i] * unp);
      if (A.rusanov) fl += R(0.5) * (lam * (xm - xp[i]));
      fl += f2n[i];
      sFace[i][it] = sMvMI * fl;
    }
  }
  // volume: weak derivative D^T (M xi . F)
  R acc[NT];
  const int i_ = tid % NQ, j_ = (tid / NQ) % NQ, k_ = tid / (NQ * NQ);
  if (tid < NP) {
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i] = R(0);
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
      const R d1 = const_D<R>(n * NQ + i_), d2 = const_D<R>(n * NQ + j_), d3 = const_D<R>(n * NQ + k_);
      const int o1 = n + NQ * (j_ + NQ * k_), o2 = i_ + NQ * (n + NQ * k_), o3 = i_ + NQ * (j_ + NQ * n);
#pragma unroll
      for (int i = 0; i < NT; ++i) acc[i] += d1 * sF[0][i][o1] + d2 * sF[1][i][o2] + d3 * sF[2][i][o3];
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i] *= MI;
  }
  __syncthreads();
  if (tid >= NP) return;
  const int fit[3] = {(i_ == 0) ? j_ + NQ * k_ : ((i_ == NQ - 1) ? NFP + j_ + NQ * k_ : -1),
                      (j_ == 0) ? 2 * NFP + i_ + NQ * k_ : ((j_ == NQ - 1) ? 3 * NFP + i_ + NQ * k_ : -1),
                      (k_ == 0) ? 4 * NFP + i_ + NQ * j_ : ((k_ == NQ - 1) ? 5 * NFP + i_ + NQ * j_ : -1)};
#pragma unroll
  for (int dir = 0; dir < 3; ++dir) {
    const int it = fit[dir];
    if (it < 0) continue;
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i] -= sFace[i][it];
  }
#pragma unroll
  for (int i = 0; i < NT; ++i)
    if (i < A.nt) {
      const size_t o = eoffQ + (size_t)(5 + i) * NP + tid;
      const R old = (A.beta != R(0)) ? A.dQ[o] : R(0);
      const R d = A.alpha * acc[i] + A.beta * old;
      A.dQ[o] = d;
      if (A.Qout) A.Qout[o] = x[i] + A.rkb_dt * d;
    }
}

}  // namespace cmdg
