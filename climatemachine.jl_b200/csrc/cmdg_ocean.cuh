// cmdg_ocean.cuh -- sm_100a device code for the ocean HydrostaticBoussinesqModel (HBModel).
//
// One tendency evaluation of the reference (src/Ocean/HydrostaticBoussinesq/
// hydrostatic_boussinesq_model.jl with the generic DGModel schedule, DGModel.jl:85-427) is
//
//   update_auxiliary_state!           2 x kernel_apply_filter! (vertical cutoff on u, vertical
//                                     exponential on theta)                        -> hb_filter_kernel
//   volume_gradients! H, V + 4 x dgsem_interface_gradients!                         -> hb_gradient_kernel
//   update_auxiliary_state_gradient!  w = -div_h u; kernel_indefinite_stack_integral!;
//                                     kernel_reverse_indefinite_stack_integral!; wz0 copy
//                                                                                   -> hb_column_kernel
//   volume_tendency! H, V + 4 x dgsem_interface_tendency! (+ update! when fused)    -> hb_tendency_kernel
//
// i.e. 4 launches instead of 15.  State layouts (hydrostatic_boussinesq_model.jl:105-232):
//   Q   : u1, u2, eta, theta                  aux : y, w, pkin, wz0, ud1, ud2, dGu1, dGu2
//   GF  : div_h u, nu grad u (3x2 column major: d/dx_d of u_c at 1 + 3c + d), kappa grad theta (7..9)
#pragma once
#include "cmdg_kernels.cuh"

namespace cmdg {

enum { HB_VEL_NOSLIP = 1, HB_VEL_FREESLIP = 2, HB_VEL_PENETRABLE_FREESLIP = 3, HB_VEL_KINEMATIC_STRESS = 4 };
enum { HB_TEMP_INSULATING = 1, HB_TEMP_FLUX = 2 };
enum { HB_S = 4, HB_A = 8, HB_GF = 10 };

template <class R>
struct HBParams {
  R grav, rho0, ch, cz, alphaT, nuh, nuz, kappah, kappaz, kappac, f0, beta;
  R Ly, tau0, lambda_r, thetaE;  // OceanGyre-type surface forcing (ocean_gyre.jl)
  int bc_vel[6], bc_temp[6];
  int nvertelem;
};

template <class R>
struct HBArgs {
  R *Q;              // [nelem][4][Np] (filtered in place by hb_filter_kernel)
  R *aux;            // [nelem][8][Np]
  R *gradflux;       // [nelem][10][Np]
  R *dQ, *Qout;
  const R *vgeoP, *sgeoP;
  const int2 *conn;
  const int *elems;
  const R *D;        // row-major [Nq][Nq]
  R alpha, beta, rkb_dt;
};

// gradient-flux map of HBModel (compute_gradient_flux!, :248-298): NOT linear in grad theta
// (convective adjustment switches kappa on the sign of d theta / dz), so it is evaluated
// separately wherever the reference evaluates it.
template <class R>
__device__ __forceinline__ void hb_gradient_flux(const HBParams<R> &P, const R dG[3][3], R gf[10]) {
  gf[0] = dG[0][0] + dG[1][1];
  const R nu[3] = {P.nuh, P.nuh, P.nuz};
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int d = 0; d < 3; ++d) gf[1 + 3 * c + d] = -nu[d] * dG[d][c];
  const R kz = dG[2][2] < R(0) ? P.kappac : P.kappaz;
  gf[7] = -P.kappah * dG[0][2];
  gf[8] = -P.kappah * dG[1][2];
  gf[9] = -kz * dG[2][2];
}

// ---------------------------------------------------------------------------------------
// kernel_apply_filter! (Filters.jl:651-792), VerticalDirection, targets (u1,u2) with the cutoff
// matrix and theta with the exponential matrix, fused in one pass over Q.
// Fc / Fe: row-major [Nq][Nq].
// ---------------------------------------------------------------------------------------
template <class R, int NQ>
__global__ void __launch_bounds__(Dims<NQ>::BLOCK)
hb_filter_kernel(R *__restrict__ Q, const R *__restrict__ Fc, const R *__restrict__ Fe, int nreal,
                 const int *__restrict__ elems) {
  constexpr int NP = Dims<NQ>::NP;
  __shared__ R s[3][NP];
  __shared__ R sF[2][NQ * NQ];
  const int tid = threadIdx.x, e = elems ? elems[blockIdx.x] : blockIdx.x;   // launch list or identity
  if (tid < NQ * NQ) {
    sF[0][tid] = Fc[tid];
    sF[1][tid] = Fe[tid];
  }
  const size_t off = (size_t)e * HB_S * NP + tid;
  if (tid < NP) {
    s[0][tid] = Q[off];
    s[1][tid] = Q[off + NP];
    s[2][tid] = Q[off + (size_t)3 * NP];
  }
  __syncthreads();
  if (tid < NP) {
    const int ij = tid % (NQ * NQ), k = tid / (NQ * NQ);
    R a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
      a0 += sF[0][k * NQ + n] * s[0][ij + NQ * NQ * n];
      a1 += sF[0][k * NQ + n] * s[1][ij + NQ * NQ * n];
      a2 += sF[1][k * NQ + n] * s[2][ij + NQ * NQ * n];
    }
    Q[off] = a0;
    Q[off + NP] = a1;
    Q[off + (size_t)3 * NP] = a2;
  }
}

// boundary ghost velocity for the first-order (fac = 2) / gradient (fac = 1) fluxes
// (bc_velocity.jl); v = (u1, u2, w)
template <class R>
__device__ __forceinline__ void hb_boundary_velocity(int kind, bool gradient, const R n[3], R v[3]) {
  if (kind == HB_VEL_NOSLIP) {
    if (gradient) v[0] = v[1] = v[2] = R(0);
    else { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; }
  } else if (kind == HB_VEL_FREESLIP) {
    const R vn = (gradient ? R(1) : R(2)) * (n[0] * v[0] + n[1] * v[1] + n[2] * v[2]);
    v[0] -= vn * n[0];
    v[1] -= vn * n[1];
    v[2] -= vn * n[2];
  }  // penetrable: transmissive
}

// ---------------------------------------------------------------------------------------
// Gradient pass: G = (u1, u2, theta) [the ud columns are zero for the uncoupled model].
// ---------------------------------------------------------------------------------------
template <class R, int NQ>
__global__ void __launch_bounds__(Dims<NQ>::BLOCK, 4)
hb_gradient_kernel(const HBArgs<R> A, const HBParams<R> P) {
  constexpr int NP = Dims<NQ>::NP, NFP = Dims<NQ>::NFP, NFN = Dims<NQ>::NFN;
  constexpr int BLOCK = Dims<NQ>::BLOCK;
  __shared__ R sG[3][NP];
  __shared__ R sW[NP];
  __shared__ R sFace[10][NFN];
  __shared__ R sD[NQ * NQ];
  __shared__ int2 sConn[6];
  const int tid = threadIdx.x;
  const int e = A.elems ? A.elems[blockIdx.x] : blockIdx.x;
  const size_t eoffQ = (size_t)e * HB_S * NP, eoffA = (size_t)e * HB_A * NP;
  if (tid < 6) sConn[tid] = A.conn[(size_t)e * 6 + tid];
  if (tid < NQ * NQ) sD[tid] = A.D[tid];
  if (tid < NP) {
    sG[0][tid] = A.Q[eoffQ + tid];
    sG[1][tid] = A.Q[eoffQ + NP + tid];
    sG[2][tid] = A.Q[eoffQ + (size_t)3 * NP + tid];
    sW[tid] = A.aux[eoffA + (size_t)1 * NP + tid];
  }
  __syncthreads();
  for (int it = tid; it < NFN; it += BLOCK) {
    const int f = it / NFP, fn = it - f * NFP;
    const int2 c = sConn[f];
    const int bctag = (c.y >> 4) & 15;
    const int vm = face_to_vol<NQ>(f, fn % NQ, fn / NQ);
    R n[3], sMvMI;
    load_sgeo<R>(A.sgeoP + ((size_t)e * NFN + it) * 4, n, sMvMI);
    const R Gm[3] = {sG[0][vm], sG[1][vm], sG[2][vm]};
    R Gs[3];
    if (bctag == 0) {
      int a = fn % NQ;
      if (c.y & 8) a = NQ - 1 - a;
      const int vp = face_to_vol<NQ>(c.y & 7, a, fn / NQ);
      const size_t op = (size_t)c.x * HB_S * NP + vp;
      Gs[0] = R(0.5) * (A.Q[op] + Gm[0]);
      Gs[1] = R(0.5) * (A.Q[op + NP] + Gm[1]);
      Gs[2] = R(0.5) * (A.Q[op + (size_t)3 * NP] + Gm[2]);
    } else {
      R v[3] = {Gm[0], Gm[1], sW[vm]};
      hb_boundary_velocity<R>(P.bc_vel[bctag - 1], true, n, v);
      Gs[0] = v[0];
      Gs[1] = v[1];
      Gs[2] = Gm[2];
    }
    R dGs[3][3], dGm[3][3], gfs[10], gfm[10];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        dGs[d][s] = n[d] * Gs[s];
        dGm[d][s] = n[d] * Gm[s];
      }
    hb_gradient_flux<R>(P, dGs, gfs);
    hb_gradient_flux<R>(P, dGm, gfm);
#pragma unroll
    for (int s = 0; s < 10; ++s) sFace[s][it] = sMvMI * (gfs[s] - gfm[s]);
  }
  R gfv[10];
  const int i = tid % NQ, j = (tid / NQ) % NQ, k = tid / (NQ * NQ);
  if (tid < NP) {
    R g[9], MI;
    load_vgeo<R>(A.vgeoP, e, tid, NP, g, MI);
#pragma unroll
    for (int c = 0; c < 9; ++c) g[c] *= MI;
    R G1[3] = {0, 0, 0}, G2[3] = {0, 0, 0}, G3[3] = {0, 0, 0};
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
      const R d1 = sD[i * NQ + n], d2 = sD[j * NQ + n], d3 = sD[k * NQ + n];
      const int o1 = n + NQ * (j + NQ * k), o2 = i + NQ * (n + NQ * k), o3 = i + NQ * (j + NQ * n);
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        G1[s] += d1 * sG[s][o1];
        G2[s] += d2 * sG[s][o2];
        G3[s] += d3 * sG[s][o3];
      }
    }
    // horizontal launch (GF = gf(grad_H G)), then vertical launch (GF += gf(grad_V G))
    R dH[3][3], dV[3][3], gh[10], gv[10];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        dH[d][s] = g[d] * G1[s] + g[3 + d] * G2[s];
        dV[d][s] = g[6 + d] * G3[s];
      }
    hb_gradient_flux<R>(P, dH, gh);
    hb_gradient_flux<R>(P, dV, gv);
#pragma unroll
    for (int s = 0; s < 10; ++s) gfv[s] = gh[s] + gv[s];
  }
  __syncthreads();
  if (tid < NP) {
    if (i == 0) for (int s = 0; s < 10; ++s) gfv[s] += sFace[s][0 * NFP + j + NQ * k];
    if (i == NQ - 1) for (int s = 0; s < 10; ++s) gfv[s] += sFace[s][1 * NFP + j + NQ * k];
    if (j == 0) for (int s = 0; s < 10; ++s) gfv[s] += sFace[s][2 * NFP + i + NQ * k];
    if (j == NQ - 1) for (int s = 0; s < 10; ++s) gfv[s] += sFace[s][3 * NFP + i + NQ * k];
    if (k == 0) for (int s = 0; s < 10; ++s) gfv[s] += sFace[s][4 * NFP + i + NQ * j];
    if (k == NQ - 1) for (int s = 0; s < 10; ++s) gfv[s] += sFace[s][5 * NFP + i + NQ * j];
    const size_t eoffG = (size_t)e * HB_GF * NP + tid;
#pragma unroll
    for (int s = 0; s < 10; ++s) A.gradflux[eoffG + (size_t)s * NP] = gfv[s];
  }
}

// ---------------------------------------------------------------------------------------
// update_auxiliary_state_gradient! (:675-712): one block per horizontal stack, one thread per
// horizontal node (i,j); marches up the stack.  Imat: row-major [Nq][Nq]; JcV: private copy of
// vgeo column 16 ([nelem][Np]).  stack0/nstack select real (with wz0) or ghost stacks.
// ---------------------------------------------------------------------------------------
template <class R, int NQ>
__global__ void hb_column_kernel(R *__restrict__ aux, const R *__restrict__ Q,
                                 const R *__restrict__ gradflux, const R *__restrict__ JcV,
                                 const R *__restrict__ Imat, R alphaT, int nvert, int elem0,
                                 int set_wz0, const int *__restrict__ stack_first) {
  constexpr int NP = Dims<NQ>::NP, NQH = NQ * NQ;
  __shared__ R sI[NQ * NQ];
  const int ij = threadIdx.x;
  if (ij < NQ * NQ) sI[ij] = Imat[ij];
  __syncthreads();
  if (ij >= NQH) return;
  // stacks elem0, elem0 + nvert, ... or the listed stacks (first element of each)
  const int e0 = stack_first ? stack_first[blockIdx.x] : elem0 + blockIdx.x * nvert;
  R cw = 0, cp = 0;  // carried integrals (top value of the element below)
  for (int ev = 0; ev < nvert; ++ev) {
    const size_t e = (size_t)(e0 + ev);
    R kw[NQ], kp[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const int n = ij + NQH * k;
      const R jc = JcV[e * NP + n];
      kw[k] = -gradflux[(e * HB_GF + 0) * NP + n] * jc;      // A.w = -D.div_h u, then * JcV
      kp[k] = (-alphaT * Q[(e * HB_S + 3) * NP + n]) * jc;
    }
    R lw[NQ], lp[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      lw[k] = cw;
      lp[k] = cp;
#pragma unroll
      for (int n = 0; n < NQ; ++n) {
        lw[k] += sI[k * NQ + n] * kw[n];
        lp[k] += sI[k * NQ + n] * kp[n];
      }
      aux[(e * HB_A + 1) * NP + ij + NQH * k] = lw[k];
      aux[(e * HB_A + 2) * NP + ij + NQH * k] = lp[k];
    }
    cw = lw[NQ - 1];
    cp = lp[NQ - 1];
  }
  // reverse integral: pkin <- pkin(top) - pkin; wz0 <- w(top) over the whole column
  for (int ev = 0; ev < nvert; ++ev) {
    const size_t e = (size_t)(e0 + ev);
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const size_t o = (e * HB_A + 2) * NP + ij + NQH * k;
      aux[o] = cp - aux[o];
      if (set_wz0) aux[(e * HB_A + 3) * NP + ij + NQH * k] = cw;
    }
  }
}

// The same as a segmented scan: one block per stack, thread = (element slot, horizontal node).  Every
// element's local integral (carry 0) is independent; only the element-top values chain up the stack
// (kernel_indefinite_stack_integral!, DGModel_kernels.jl:1903-2007, carries them in a serial k-loop, one
// thread per (i, j) and a whole stack per thread: 2.7 warps per SM at 20 x 20 stacks and 50 dependent
// DRAM round trips -- 114 us, 21 % of an ocean stage).  Pass 1: tops of all elements in parallel -> shared
// memory; exclusive prefix over the stack by the first slot's threads; pass 2: local integrals again (the
// inputs are L1/L2 hits) + carry, reverse integral (:2009-2104) and wz0.  Dynamic shared memory:
// 2 * nvert * Nq^2 values.  Differs from the serial kernel in rounding order only.
template <class R, int NQ, int SLOTS>
__global__ void __launch_bounds__(SLOTS * NQ * NQ)
hb_column_scan_kernel(R *__restrict__ aux, const R *__restrict__ Q, const R *__restrict__ gradflux,
                      const R *__restrict__ JcV, const R *__restrict__ Imat, R alphaT, int nvert, int elem0,
                      int set_wz0, const int *__restrict__ stack_first) {
  constexpr int NP = Dims<NQ>::NP, NQH = NQ * NQ;
  extern __shared__ __align__(16) unsigned char hb_col_smem[];
  R *topw = reinterpret_cast<R *>(hb_col_smem);     // [nvert][NQH]
  R *topp = topw + (size_t)nvert * NQH;             // [nvert][NQH]
  __shared__ R sI[NQ * NQ];
  __shared__ R tot[2][NQH];
  const int tid = threadIdx.x;
  const int ij = tid % NQH, slot = tid / NQH;
  if (tid < NQ * NQ) sI[tid] = Imat[tid];
  const int e0 = stack_first ? stack_first[blockIdx.x] : elem0 + blockIdx.x * nvert;
  __syncthreads();
  // pass 1: element tops (row Nq-1 of Imat applied to the integrands)
  for (int ev = slot; ev < nvert; ev += SLOTS) {
    const size_t e = (size_t)(e0 + ev);
    R tw = 0, tp = 0;
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
      const int nn = ij + NQH * n;
      const R jc = JcV[e * NP + nn];
      tw += sI[(NQ - 1) * NQ + n] * (-gradflux[(e * HB_GF + 0) * NP + nn] * jc);
      tp += sI[(NQ - 1) * NQ + n] * ((-alphaT * Q[(e * HB_S + 3) * NP + nn]) * jc);
    }
    topw[ev * NQH + ij] = tw;
    topp[ev * NQH + ij] = tp;
  }
  __syncthreads();
  // exclusive prefix up the stack (carry of element ev = integral up to its bottom)
  if (slot == 0) {
    R cw = 0, cp = 0;
    for (int ev = 0; ev < nvert; ++ev) {
      const R tw = topw[ev * NQH + ij], tp = topp[ev * NQH + ij];
      topw[ev * NQH + ij] = cw;
      topp[ev * NQH + ij] = cp;
      cw += tw;
      cp += tp;
    }
    tot[0][ij] = cw;
    tot[1][ij] = cp;
  }
  __syncthreads();
  const R totw = tot[0][ij], totp = tot[1][ij];
  // pass 2: w = carry + local integral; pkin = pkin(top of stack) - (carry + local); wz0 = w(top of stack)
  for (int ev = slot; ev < nvert; ev += SLOTS) {
    const size_t e = (size_t)(e0 + ev);
    R kw[NQ], kp[NQ];
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
      const int nn = ij + NQH * n;
      const R jc = JcV[e * NP + nn];
      kw[n] = -gradflux[(e * HB_GF + 0) * NP + nn] * jc;
      kp[n] = (-alphaT * Q[(e * HB_S + 3) * NP + nn]) * jc;
    }
    const R cw = topw[ev * NQH + ij], cp = topp[ev * NQH + ij];
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      R lw = cw, lp = cp;
#pragma unroll
      for (int n = 0; n < NQ; ++n) {
        lw += sI[k * NQ + n] * kw[n];
        lp += sI[k * NQ + n] * kp[n];
      }
      const int nn = ij + NQH * k;
      aux[(e * HB_A + 1) * NP + nn] = lw;
      aux[(e * HB_A + 2) * NP + nn] = totp - lp;
      if (set_wz0) aux[(e * HB_A + 3) * NP + nn] = totw;
    }
  }
}

template <class R>
__global__ void extract_column_kernel(R *__restrict__ out, const R *__restrict__ vgeo, int Np,
                                      int nvgeo, int col, size_t nelem) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nelem * Np) return;
  const size_t e = idx / Np;
  out[idx] = vgeo[(e * nvgeo + col) * Np + (idx - e * Np)];
}

// second-order flux F2[d][s] from the gradient-flux columns (flux_second_order!, :527-540)
template <class R>
__device__ __forceinline__ void hb_flux2_normal(const R *gf, const R n[3], R fn[4]) {
  fn[0] = gf[1] * n[0] + gf[2] * n[1] + gf[3] * n[2];
  fn[1] = gf[4] * n[0] + gf[5] * n[1] + gf[6] * n[2];
  fn[2] = R(0);
  fn[3] = gf[7] * n[0] + gf[8] * n[1] + gf[9] * n[2];
}

// ---------------------------------------------------------------------------------------
// Fused tendency (+ RK stage update) for HBModel.
// ---------------------------------------------------------------------------------------
template <class R, int NQ, int NF1>
__global__ void __launch_bounds__(Dims<NQ>::BLOCK, 4)
hb_tendency_kernel(const HBArgs<R> A, const HBParams<R> P) {
  constexpr int NP = Dims<NQ>::NP, NFP = Dims<NQ>::NFP, NFN = Dims<NQ>::NFN;
  constexpr int BLOCK = Dims<NQ>::BLOCK;
  constexpr int NITEM = (NFN + BLOCK - 1) / BLOCK;
  __shared__ R sQ[HB_S][NP];
  __shared__ R sA[3][NP];            // y, w, pkin
  __shared__ R sGF[9][NP];           // gradient-flux columns 1..9
  __shared__ R sF[3][3][NP];         // contravariant fluxes of u1, u2, theta (eta has none)
  __shared__ R sFace[3][NFN];
  __shared__ R sD[NQ * NQ];
  const int tid = threadIdx.x;
  const int e = A.elems ? A.elems[blockIdx.x] : blockIdx.x;
  const size_t eoffQ = (size_t)e * HB_S * NP, eoffA = (size_t)e * HB_A * NP,
               eoffG = (size_t)e * HB_GF * NP;
  if (tid < NQ * NQ) sD[tid] = A.D[tid];
  int2 cn[NITEM];
#pragma unroll
  for (int r = 0; r < NITEM; ++r) {
    const int it = tid + r * BLOCK;
    cn[r] = (it < NFN) ? A.conn[(size_t)e * 6 + it / NFP] : make_int2(0, 0);
  }
  R q[HB_S] = {0, 0, 0, 0}, src[HB_S] = {0, 0, 0, 0}, MI = 0;
  if (tid < NP) {
#pragma unroll
    for (int s = 0; s < HB_S; ++s) q[s] = A.Q[eoffQ + (size_t)s * NP + tid];
    const R y = A.aux[eoffA + tid], w = A.aux[eoffA + NP + tid];
    const R pkin = A.aux[eoffA + (size_t)2 * NP + tid], wz0 = A.aux[eoffA + (size_t)3 * NP + tid];
    R gf[10];
#pragma unroll
    for (int s = 1; s < 10; ++s) gf[s] = A.gradflux[eoffG + (size_t)s * NP + tid];
    R g[9];
    load_vgeo<R>(A.vgeoP, e, tid, NP, g, MI);
#pragma unroll
    for (int s = 0; s < HB_S; ++s) sQ[s][tid] = q[s];
    sA[0][tid] = y;
    sA[1][tid] = w;
    sA[2][tid] = pkin;
#pragma unroll
    for (int s = 1; s < 10; ++s) sGF[s - 1][tid] = gf[s];
    // F1 + F2 (flux_first_order! :427-507, flux_second_order! :527-540)
    const R pr = P.grav * q[2] + P.grav * pkin;
    const R v[3] = {q[0], q[1], w};
    R F[3][3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      F[d][0] = gf[1 + d];
      F[d][1] = gf[4 + d];
      F[d][2] = v[d] * q[3] + gf[7 + d];
    }
    F[0][0] += pr;
    F[1][1] += pr;
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int s = 0; s < 3; ++s)
        sF[m][s][tid] = g[3 * m] * F[0][s] + g[3 * m + 1] * F[1][s] + g[3 * m + 2] * F[2][s];
    const R fc = P.f0 + P.beta * y;
    src[0] = fc * q[1];
    src[1] = -(fc * q[0]);
    src[2] = wz0;
  }
  __syncthreads();
  R dQold[HB_S] = {0, 0, 0, 0};
  if (tid < NP && A.beta != R(0)) {
#pragma unroll
    for (int s = 0; s < HB_S; ++s) dQold[s] = A.dQ[eoffQ + (size_t)s * NP + tid];
  }
  R acc[3] = {0, 0, 0};
  const int i = tid % NQ, j = (tid / NQ) % NQ, k = tid / (NQ * NQ);
  if (tid < NP) {
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
      const R d1 = sD[n * NQ + i], d2 = sD[n * NQ + j], d3 = sD[n * NQ + k];
      const int o1 = n + NQ * (j + NQ * k), o2 = i + NQ * (n + NQ * k), o3 = i + NQ * (j + NQ * n);
#pragma unroll
      for (int s = 0; s < 3; ++s) acc[s] += d1 * sF[0][s][o1] + d2 * sF[1][s][o2] + d3 * sF[2][s][o3];
    }
#pragma unroll
    for (int s = 0; s < 3; ++s) acc[s] *= MI;
  }
  // faces
#pragma unroll
  for (int r = 0; r < NITEM; ++r) {
    const int it = tid + r * BLOCK;
    if (it >= NFN) break;
    const int f = it / NFP, fn = it - f * NFP;
    const int2 c = cn[r];
    const int bctag = (c.y >> 4) & 15;
    const int vm = face_to_vol<NQ>(f, fn % NQ, fn / NQ);
    R n[3], sMvMI;
    load_sgeo<R>(A.sgeoP + ((size_t)e * NFN + it) * 4, n, sMvMI);
    const R um[2] = {sQ[0][vm], sQ[1][vm]}, etam = sQ[2][vm], thm = sQ[3][vm];
    const R ym = sA[0][vm], wm = sA[1][vm], pkm = sA[2][vm];
    R gfm[10], gfp[10];
#pragma unroll
    for (int s = 1; s < 10; ++s) gfm[s] = sGF[s - 1][vm];
    R up[2], etap = etam, thp = thm, wp = wm, pkp = pkm;
    R f2n[4];
    if (bctag == 0) {
      int a = fn % NQ;
      if (c.y & 8) a = NQ - 1 - a;
      const int vp = face_to_vol<NQ>(c.y & 7, a, fn / NQ);
      const size_t oq = (size_t)c.x * HB_S * NP + vp, oa = (size_t)c.x * HB_A * NP + vp,
                   og = (size_t)c.x * HB_GF * NP + vp;
      up[0] = A.Q[oq];
      up[1] = A.Q[oq + NP];
      etap = A.Q[oq + (size_t)2 * NP];
      thp = A.Q[oq + (size_t)3 * NP];
      wp = A.aux[oa + NP];
      pkp = A.aux[oa + (size_t)2 * NP];
#pragma unroll
      for (int s = 1; s < 10; ++s) gfp[s] = A.gradflux[og + (size_t)s * NP];
      R fm2[4], fp2[4];
      hb_flux2_normal<R>(gfm, n, fm2);
      hb_flux2_normal<R>(gfp, n, fp2);
#pragma unroll
      for (int s = 0; s < 4; ++s) f2n[s] = R(0.5) * (fm2[s] + fp2[s]);
    } else {
      // first-order ghost state (bc_velocity.jl / bc_temperature.jl)
      R v[3] = {um[0], um[1], wm};
      const int kv = P.bc_vel[bctag - 1], kt = P.bc_temp[bctag - 1];
      hb_boundary_velocity<R>(kv == HB_VEL_KINEMATIC_STRESS ? HB_VEL_PENETRABLE_FREESLIP : kv, false, n, v);
      up[0] = v[0];
      up[1] = v[1];
      wp = v[2];
      // second-order boundary flux = F2(boundary_state(GF-)) . n  (NumericalFluxes.jl:872-967)
#pragma unroll
      for (int s = 1; s < 10; ++s) gfp[s] = gfm[s];
      if (kv == HB_VEL_FREESLIP || kv == HB_VEL_PENETRABLE_FREESLIP) {
#pragma unroll
        for (int s = 1; s < 7; ++s) gfp[s] = R(0);
      } else if (kv == HB_VEL_KINEMATIC_STRESS) {
        const R st0 = (P.tau0 / P.rho0) * cos(ym * R(3.14159265358979323846) / P.Ly);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          gfp[1 + d] = n[d] * st0;
          gfp[4 + d] = R(0);
        }
      }
      if (kt == HB_TEMP_INSULATING) {
        gfp[7] = gfp[8] = gfp[9] = R(0);
      } else {
        const R sf = P.lambda_r * (thm - P.thetaE * (R(1) - ym / P.Ly));
#pragma unroll
        for (int d = 0; d < 3; ++d) gfp[7 + d] = n[d] * sf;
      }
      hb_flux2_normal<R>(gfp, n, f2n);
    }
    // first-order numerical flux: central + Rusanov penalty without the eta component
    const R prm = P.grav * etam + P.grav * pkm, prp = P.grav * etap + P.grav * pkp;
    const R vnm = um[0] * n[0] + um[1] * n[1] + wm * n[2];
    const R vnp = up[0] * n[0] + up[1] * n[1] + wp * n[2];
    R fl[3];
    fl[0] = R(0.5) * (prm + prp) * n[0];
    fl[1] = R(0.5) * (prm + prp) * n[1];
    fl[2] = R(0.5) * (vnm * thm + vnp * thp);
    if (NF1 == NF_RUSANOV) {
      const R lam = R(0.5) * fabs(P.ch * n[0] + P.ch * n[1] + P.cz * n[2]);
      fl[0] += lam * (um[0] - up[0]);
      fl[1] += lam * (um[1] - up[1]);
      fl[2] += lam * (thm - thp);
    }
    sFace[0][it] = sMvMI * (fl[0] + f2n[0]);
    sFace[1][it] = sMvMI * (fl[1] + f2n[1]);
    sFace[2][it] = sMvMI * (fl[2] + f2n[3]);
  }
  __syncthreads();
  if (tid < NP) {
    if (i == 0) for (int s = 0; s < 3; ++s) acc[s] -= sFace[s][0 * NFP + j + NQ * k];
    if (i == NQ - 1) for (int s = 0; s < 3; ++s) acc[s] -= sFace[s][1 * NFP + j + NQ * k];
    if (j == 0) for (int s = 0; s < 3; ++s) acc[s] -= sFace[s][2 * NFP + i + NQ * k];
    if (j == NQ - 1) for (int s = 0; s < 3; ++s) acc[s] -= sFace[s][3 * NFP + i + NQ * k];
    if (k == 0) for (int s = 0; s < 3; ++s) acc[s] -= sFace[s][4 * NFP + i + NQ * j];
    if (k == NQ - 1) for (int s = 0; s < 3; ++s) acc[s] -= sFace[s][5 * NFP + i + NQ * j];
    const R T[HB_S] = {acc[0] + src[0], acc[1] + src[1], src[2], acc[2]};
#pragma unroll
    for (int s = 0; s < HB_S; ++s) {
      const R d = A.alpha * T[s] + A.beta * dQold[s];
      A.dQ[eoffQ + (size_t)s * NP + tid] = d;
      if (A.Qout) A.Qout[eoffQ + (size_t)s * NP + tid] = q[s] + A.rkb_dt * d;
    }
  }
}

}  // namespace cmdg
