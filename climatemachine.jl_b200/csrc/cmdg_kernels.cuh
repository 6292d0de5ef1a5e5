// cmdg_kernels.cuh -- sm_100a device code of libcmdg.
//
// One thread block owns one element (Np = Nq^3 nodes, one thread per node).  A tendency
// evaluation is ONE kernel: volume term, all six face terms, sources, the alpha/beta
// combination with the old tendency and (optionally) the low-storage RK stage update are
// fused, so every array word is touched once per stage:
//
//   reference (src/Numerics/DGMethods/DGModel_kernels.jl)          here
//   volume_tendency! H  (:64-309)   \
//   volume_tendency! V  (:312-548)   |
//   dgsem_interface_tendency! x4     |--> dg_tendency_kernel
//       (:588-901)                   |
//   kernel_nodal_update_auxiliary_state! (:1769-1825)
//   update! (ODESolvers/LowStorageRungeKuttaMethod.jl:146-158)  /
//
// Memory-system design (the path is HBM-bound, ~1.3 flop/B in FP64):
//   * state, tendency and geometry are read with fully coalesced 8-byte loads, one
//     contiguous Np-run per (state, element) -- the Np x nstate x nelem layout of
//     MPIStateArray makes every column of an element a 1000-byte contiguous run;
//   * geometry comes from a private packed copy (10 of the 25 vgeo columns with the mass
//     matrix folded into the metric terms, 4 of 5 sgeo rows with vMI*sM folded), built
//     once in cmdg_bind_grid;
//   * the 25 x 6 Int64 vmap+ entries per element are replaced by one 8-byte descriptor
//     per face (neighbour element, neighbour face, orientation, boundary tag);
//   * contravariant fluxes are staged in shared memory and contracted with the Nq x Nq
//     derivative matrix held in shared memory; neighbour traces are gathered once per face
//     node into shared memory at kernel start so their latency overlaps the volume work.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cmdg {

enum { NF_RUSANOV = 0, NF_CENTRAL = 1, NF_ROE = 2 };
enum { TURB_CONST_KINEMATIC = 0, TURB_CONST_DYNAMIC = 1, TURB_SMAGORINSKY = 2 };
enum { SRC_GRAVITY = 1, SRC_CORIOLIS = 2, SRC_HELD_SUAREZ = 4, SRC_RAYLEIGH_SPONGE = 8 };
enum { BC_FREESLIP = 1, BC_NOSLIP = 2 };

// Uniform (per launch) physics parameters of the dry AtmosModel.
template <class R>
struct AtmosParams {
  R R_d, cp_d, cv_d, T_0, MSLP, grav, two_Omega, inv_Pr_turb;
  R gamma;       // cp_d / cv_d
  R inv_cv;      // 1 / cv_d
  R kappa;       // R_d / cp_d
  R turb_param;  // nu, rho*nu or C_smag
  int turbulence, with_divergence;
  int sources;
  int subtract_off;
  int horizontal_diffusion;
  int bc_kind[6];
  // auxiliary-state column ids (0-based), -1 when absent
  int a_Phi, a_gradPhi, a_ref_rho, a_ref_p, a_Delta, a_Delta_h, a_theta_v, a_T;
  R hyper_tau;   // DryBiharmonic time scale (a_Delta_h >= 0 when hyperdiffusion is on)
  int naux, ngradflux;
  // column strides of Q / dQ (5 + passive tracers) and the number of gradient-flux columns of the
  // dynamics (9 or 10; the tracers' grad chi columns follow them)
  int nstate, ngf_dyn;
  // HeldSuarezForcing / RayleighSponge (SRCX kernels only)
  R inv_day, sponge_z_max, sponge_z_sponge, sponge_alpha_max, sponge_gamma, sponge_u[3];
};

template <class R>
struct TendArgs {
  const R *Q;          // [nelem][5][Np]
  const R *aux;        // [nelem][naux][Np]
  const R *gradflux;   // [nelem][ngradflux][Np]
  R *dQ;               // [nelem][5][Np]
  R *Qout;             // fused RK update target (may alias nothing in Q); NULL = no update
  R *aux_out;          // write theta_v / air_T here (NULL = don't)
  const R *vgeoP;      // [nreal][5][Np] pairs: M*xi{m}x{d} (m-major) c = 0..8, MI c = 9; pair p = (c 2p, c 2p+1)  (80 B per node)
  const R *sgeoP;      // [nreal][6*Nfp][4]  n1,n2,n3, sM*vMI         (32 B per face node)
  const int2 *conn;    // [nreal][6]  x = neighbour element (0-based), y = meta
  const R *fauxP;      // [nreal][6*Nfp][2]  neighbour's geopotential and reference pressure at the matching node (16 B per face node)
  const int *elems;    // launch list (0-based element ids) or NULL for identity
  const R *D;          // [Nq][Nq] row-major copy of Julia's D (D[a][b] = D_julia[a+1,b+1])
  R alpha, beta;       // dQ = alpha*RHS + beta*dQ
  R rkb_dt;            // Qout = Q + rkb_dt * dQ
  R t;
  int pf_dist;         // L2 prefetch distance in launch-list entries (0 = off)
  // second-order path: diffusive flux evaluated once per node by the gradient kernel
  const R *F2;         // [nelem][12][Np]   F2[d][s], s = 1..4, at column 4*d + s-1 (ghosts by exchange)
  const R *Fn;         // [nreal][6*Nfp][4] n . F2 at the element's own face nodes
  int nreal;
  int nelem;           // real + ghost elements of Q (bulk-copy windows must stay inside the array)
  // tail prefetch: the last pfn_n[0] + pfn_n[1] blocks of this launch pull the inputs of the first
  // elements of the NEXT launch(es) into L2 (their first wave would otherwise start on cold DRAM misses:
  // the stage's output was written ~0.5 ms -- several L2 capacities -- earlier).  pfn_list[i] = launch
  // list of the next launch (NULL = identity), pfn_Q = the state it will read (this launch's Qout).
  const int *pfn_list[2];
  int pfn_n[2];
  const R *pfn_Q;
  // single-launch halo trigger: the first n_signal blocks of the launch (the exterior elements, listed first)
  // bump *ext_done when their outputs are globally visible; a stream memory operation on the halo stream
  // (cuStreamWaitValue32) waits for the count and then starts pack -> NCCL -> unpack while the same launch
  // goes on with the interior elements
  unsigned *ext_done;
  int n_signal;
};

// conn.y layout: bits 0-2 neighbour face (0..5), bit 3 flip of first face index,
// bits 4-7 boundary tag (0 = interior face)
__host__ __device__ inline int conn_meta(int nface, int flip, int bctag) {
  return nface | (flip << 3) | (bctag << 4);
}

template <int NQ>
__device__ __forceinline__ int face_to_vol(int f, int a, int b) {
  switch (f) {
    case 0: return NQ * (a + NQ * b);
    case 1: return (NQ - 1) + NQ * (a + NQ * b);
    case 2: return a + NQ * NQ * b;
    case 3: return a + NQ * ((NQ - 1) + NQ * b);
    case 4: return a + NQ * b;
    default: return a + NQ * (b + NQ * (NQ - 1));
  }
}

// Index tables for Nq = 5 (compile-time constants in global memory, read through L1 with lane-consecutive
// indices).  The face-item decomposition (it -> face, face node, (a, b), minus-side volume node), the
// neighbour-side volume node for every (neighbour face, flip) and a node's three face items were integer
// div / mod by 5 and 25 plus a six-way switch, recomputed in every phase: 12-16 % of the tendency kernel's
// instructions (ncu source counters, round 2).
constexpr int f2v_host(int NQ, int f, int a, int b) {
  return f == 0 ? NQ * (a + NQ * b)
       : f == 1 ? (NQ - 1) + NQ * (a + NQ * b)
       : f == 2 ? a + NQ * NQ * b
       : f == 3 ? a + NQ * ((NQ - 1) + NQ * b)
       : f == 4 ? a + NQ * b
                : a + NQ * (b + NQ * (NQ - 1));
}
struct IndexTables5 {
  unsigned item[150];          // vm | f << 8 | fn << 12
  unsigned char vp[16][25];    // [neighbour face | flip << 3][fn] -> neighbour volume node
  unsigned node[128];          // it1 | it2 << 8 | it3 << 16 (255 = the node is not on a face in that direction)
};
constexpr IndexTables5 make_index_tables5() {
  IndexTables5 t{};
  for (int it = 0; it < 150; ++it) {
    const int f = it / 25, fn = it % 25;
    t.item[it] = (unsigned)f2v_host(5, f, fn % 5, fn / 5) | ((unsigned)f << 8) | ((unsigned)fn << 12);
  }
  for (int m = 0; m < 16; ++m)
    for (int fn = 0; fn < 25; ++fn) {
      const int f = m & 7, a = (m & 8) ? 4 - fn % 5 : fn % 5;
      t.vp[m][fn] = (unsigned char)(f < 6 ? f2v_host(5, f, a, fn / 5) : 0);
    }
  for (int n = 0; n < 128; ++n) {
    const int i = n % 5, j = (n / 5) % 5, k = n / 25;
    const int it1 = n >= 125 ? 255 : (i == 0 ? j + 5 * k : (i == 4 ? 25 + j + 5 * k : 255));
    const int it2 = n >= 125 ? 255 : (j == 0 ? 50 + i + 5 * k : (j == 4 ? 75 + i + 5 * k : 255));
    const int it3 = n >= 125 ? 255 : (k == 0 ? 100 + i + 5 * j : (k == 4 ? 125 + i + 5 * j : 255));
    t.node[n] = (unsigned)it1 | ((unsigned)it2 << 8) | ((unsigned)it3 << 16);
  }
  return t;
}
__device__ const IndexTables5 d_tbl5 = make_index_tables5();

template <class R> struct Vec2;
template <> struct Vec2<double> { typedef double2 type; };
template <> struct Vec2<float> { typedef float2 type; };

// node geometry (10 words) / face-node geometry (4 words) as vector loads.
// vgeoP is stored as five pair columns per element, [e][5][Np] of (2 words): lane n of a warp reads pair
// c at ((e * 5 + c) * Np + n), i.e. one warp-level load covers 32 x 16 contiguous bytes = 4-5 cache lines.
// (Round 1 stored [e][Np][10]: the same five 16-byte loads per thread then had an 80-byte lane stride and
// touched 20 lines each -- 400 L1 tag requests per element, 39 % of all global requests of the tendency
// kernel, whose limiter is the L1 / LSU pipe.)
template <class R>
__device__ __forceinline__ void load_vgeo(const R *__restrict__ vgeoP, int e, int n, int Np, R g[9], R &MI) {
  typedef typename Vec2<R>::type V;
  const V *__restrict__ v = reinterpret_cast<const V *>(vgeoP) + (size_t)e * 5 * Np + n;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const V x = v[(size_t)c * Np];
    g[2 * c] = x.x;
    g[2 * c + 1] = x.y;
  }
  const V x = v[(size_t)4 * Np];
  g[8] = x.x;
  MI = x.y;
}
template <class R>
__device__ __forceinline__ void load_sgeo(const R *__restrict__ p, R n[3], R &sMvMI) {
  typedef typename Vec2<R>::type V;
  const V *__restrict__ v = reinterpret_cast<const V *>(p);
  const V a = v[0], b = v[1];
  n[0] = a.x;
  n[1] = a.y;
  n[2] = b.x;
  sMvMI = b.y;
}

template <class R> __device__ __forceinline__ R rsqrt_(R x);
template <> __device__ __forceinline__ double rsqrt_(double x) { return 1.0 / sqrt(x); }
template <> __device__ __forceinline__ float rsqrt_(float x) { return 1.0f / sqrtf(x); }
template <class R> __device__ __forceinline__ R fast_rsqrt_(R x);
template <> __device__ __forceinline__ double fast_rsqrt_(double x) { return rsqrt(x); }
template <> __device__ __forceinline__ float fast_rsqrt_(float x) { return rsqrtf(x); }
template <class R> __device__ __forceinline__ R sqrt_(R x);
template <> __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
template <class R> __device__ __forceinline__ R pow_(R x, R y);
template <> __device__ __forceinline__ double pow_(double x, double y) { return pow(x, y); }
template <> __device__ __forceinline__ float pow_(float x, float y) { return powf(x, y); }
template <class R> __device__ __forceinline__ R cbrt_(R x);
template <> __device__ __forceinline__ double cbrt_(double x) { return cbrt(x); }
template <> __device__ __forceinline__ float cbrt_(float x) { return cbrtf(x); }

// Dry-air thermodynamic state from the prognostic state
// (src/Atmos/Model/thermo_states.jl:67-77, moisture.jl:32-46; Thermodynamics.jl PhaseDry).
template <class R>
struct Thermo {
  R rinv, T, p;
};
template <class R>
__device__ __forceinline__ Thermo<R> thermo(const AtmosParams<R> &P, const R q[5], R Phi) {
  Thermo<R> th;
  th.rinv = R(1) / q[0];
  R ke = th.rinv * (q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) * R(0.5);
  R e_int = th.rinv * (q[4] - ke - q[0] * Phi);
  th.T = P.T_0 + e_int * P.inv_cv;
  th.p = P.R_d * q[0] * th.T;
  return th;
}

// Normal component of the first-order flux, n.F(q)  (tendencies_{mass,momentum,energy}.jl)
template <class R>
__device__ __forceinline__ void normal_flux(const R q[5], R rinv, R pflux, R p, const R n[3],
                                            R fn[5], R &un) {
  un = rinv * (q[1] * n[0] + q[2] * n[1] + q[3] * n[2]);
  fn[0] = q[0] * un;
  fn[1] = q[1] * un + pflux * n[0];
  fn[2] = q[2] * un + pflux * n[1];
  fn[3] = q[3] * un + pflux * n[2];
  fn[4] = (q[4] + p) * un;
}

// Second-order (diffusive) flux F2[d][s] of the dry AtmosModel
// (tendencies_momentum.jl:36-43, tendencies_energy.jl:27-59, TurbulenceClosures.jl:364-499).
// gf: grad h_tot[3], S11,S21,S31,S22,S32,S33, [N2]
// Diagonal of the turbulent viscosity tensor nu (turbulence_tensors, TurbulenceClosures.jl:368-404
// constant, :472-499 Smagorinsky-Lilly with the buoyancy correction).
template <class R>
__device__ __forceinline__ void turbulence_nu(const AtmosParams<R> &P, const R q[5], const R *gf,
                                              const R gradPhi[3], R Delta, R nu[3]) {
  const R S[3][3] = {{gf[3], gf[4], gf[5]}, {gf[4], gf[6], gf[7]}, {gf[5], gf[7], gf[8]}};
  if (P.turbulence == TURB_SMAGORINSKY) {
    R norm2 = S[0][0] * S[0][0] + 2 * S[1][0] * S[1][0] + 2 * S[2][0] * S[2][0] +
              S[1][1] * S[1][1] + 2 * S[2][1] * S[2][1] + S[2][2] * S[2][2];
    R normS = sqrt_<R>(2 * norm2);
    const R ig = R(1) / P.grav;   // uniform: one division per call instead of three
    R k[3] = {gradPhi[0] * ig, gradPhi[1] * ig, gradPhi[2] * ig};
    // eps(normS): spacing of floating point numbers at normS (Julia's eps(x)), from the exponent
    // bits for normal numbers, nextafter otherwise
    R epsn;
    if (sizeof(R) == 8) {
      const double x = fabs((double)normS);
      const long long eb = __double_as_longlong(x) & 0x7ff0000000000000LL;
      if (eb > (53LL << 52) && eb < 0x7ff0000000000000LL)
        epsn = (R)__longlong_as_double(eb - (52LL << 52));
      else
        epsn = (R)(x == 0.0 ? 4.9406564584124654e-324 : (nextafter(x, 1.0e308 * 10) - x));
    } else {
      const float x = fabsf((float)normS);
      const int eb = __float_as_int(x) & 0x7f800000;
      if (eb > (24 << 23) && eb < 0x7f800000)
        epsn = (R)__int_as_float(eb - (23 << 23));
      else
        epsn = (R)(x == 0.0f ? 1.4012984643e-45f : (nextafterf(x, 3.0e38f * 10) - x));
    }
    R Ri = gf[9] / (normS * normS + epsn);
    R fb = R(1) - Ri * P.inv_Pr_turb;
    fb = fb < R(0) ? R(0) : (fb > R(1) ? R(1) : fb);
    R f_b2 = sqrt_<R>(fb);
    R Cd = P.turb_param * Delta;
    R nu0 = normS * (Cd * Cd) + R(1e-5);
    R dotnuk = nu0 * k[0] + nu0 * k[1] + nu0 * k[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      R nu_v = k[i] * dotnuk;
      nu[i] = (nu0 - nu_v) + nu_v * f_b2;
    }
  } else {
    R nuc = (P.turbulence == TURB_CONST_KINEMATIC) ? P.turb_param : P.turb_param / q[0];
    nu[0] = nu[1] = nu[2] = nuc;
  }
}

template <class R>
__device__ __forceinline__ void flux_second_order(const AtmosParams<R> &P, const R q[5],
                                                  const R *gf, const R gradPhi[3], R Delta,
                                                  R F2[3][5]) {
  const R S[3][3] = {{gf[3], gf[4], gf[5]}, {gf[4], gf[6], gf[7]}, {gf[5], gf[7], gf[8]}};
  R nu[3];
  turbulence_nu<R>(P, q, gf, gradPhi, Delta, nu);
  R trS = S[0][0] + S[1][1] + S[2][2];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    R tau[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) tau[j] = (-2 * nu[i]) * S[i][j];
    if (P.turbulence != TURB_SMAGORINSKY && P.with_divergence) tau[i] += (2 * nu[i] / 3) * trS;
    F2[i][0] = R(0);
    F2[i][1] = tau[0] * q[0];
    F2[i][2] = tau[1] * q[0];
    F2[i][3] = tau[2] * q[0];
    R D_t = nu[i] * P.inv_Pr_turb;
    F2[i][4] = (tau[0] * q[1] + tau[1] * q[2] + tau[2] * q[3]) + ((-D_t) * gf[i]) * q[0];
  }
}

// Roe dissipation for the dry model (src/Atmos/Model/AtmosModel.jl:967-1062).
template <class R>
__device__ __forceinline__ void roe_dissipation(const AtmosParams<R> &P, const R n[3],
                                                const R qm[5], const Thermo<R> &tm,
                                                const R qp[5], const Thermo<R> &tp, R Phi,
                                                R diss[5]) {
  R um[3] = {qm[1] * tm.rinv, qm[2] * tm.rinv, qm[3] * tm.rinv};
  R up[3] = {qp[1] * tp.rinv, qp[2] * tp.rinv, qp[3] * tp.rinv};
  R hm = qm[4] * tm.rinv + P.R_d * tm.T;
  R hp = qp[4] * tp.rinv + P.R_d * tp.T;
  R c2m = P.gamma * P.R_d * tm.T, c2p = P.gamma * P.R_d * tp.T;
  R srm = sqrt_<R>(qm[0]), srp = sqrt_<R>(qp[0]);
  R iw = R(1) / (srm + srp);
  R rt = sqrt_<R>(qm[0] * qp[0]);
  R ut[3] = {(srm * um[0] + srp * up[0]) * iw, (srm * um[1] + srp * up[1]) * iw,
             (srm * um[2] + srp * up[2]) * iw};
  R ht = (srm * hm + srp * hp) * iw;
  R ct2 = (srm * c2m + srp * c2p) * iw;
  R ct = sqrt_<R>(ct2);
  R utn = ut[0] * n[0] + ut[1] * n[1] + ut[2] * n[2];
  R drho = qp[0] - qm[0];
  R dp = tp.p - tm.p;
  R du[3] = {up[0] - um[0], up[1] - um[1], up[2] - um[2]};
  R dun = du[0] * n[0] + du[1] * n[1] + du[2] * n[2];
  R w1 = fabs(utn - ct) * (dp - rt * ct * dun) / (2 * ct2);
  R w2 = fabs(utn + ct) * (dp + rt * ct * dun) / (2 * ct2);
  R w3 = fabs(utn) * (drho - dp / ct2);
  R w4 = fabs(utn) * rt;
  diss[0] = (w1 + w2 + w3) * R(0.5);
#pragma unroll
  for (int c = 0; c < 3; ++c)
    diss[1 + c] = (w1 * (ut[c] - ct * n[c]) + w2 * (ut[c] + ct * n[c]) + w3 * ut[c] +
                   w4 * (du[c] - dun * n[c])) * R(0.5);
  R utut = ut[0] * ut[0] + ut[1] * ut[1] + ut[2] * ut[2];
  R utdu = ut[0] * du[0] + ut[1] * du[1] + ut[2] * du[2];
  diss[4] = (w1 * (ht - ct * utn) + w2 * (ht + ct * utn) +
             w3 * (utut * R(0.5) + Phi - P.T_0 * P.cv_d) + w4 * (utdu - utn * dun)) * R(0.5);
}

// Derivative matrix in constant memory (row-major copy of Julia's D, at most 8 x 8): the
// contraction reads it through the constant cache (LDC), which keeps ~110 shared-memory
// wavefronts per element off the LSU data pipe -- the unit that bounds the tendency kernel.
__constant__ double c_D64[64];
__constant__ float c_D32[64];
template <class R> __device__ __forceinline__ R const_D(int idx);
template <> __device__ __forceinline__ double const_D<double>(int idx) { return c_D64[idx]; }
template <> __device__ __forceinline__ float const_D<float>(int idx) { return c_D32[idx]; }

template <int NQ>
struct Dims {
  static constexpr int NP = NQ * NQ * NQ;
  static constexpr int NFP = NQ * NQ;
  static constexpr int NFN = 6 * NFP;
#ifdef CMDG_BLOCK160
  // 5 warps for Nq = 5: warps 0-3 own the 125 nodes, warp 4 takes face items 128..149 so that
  // every warp handles exactly one face item
  static constexpr int BLOCK = (NQ == 5) ? 160 : ((NP + 31) / 32) * 32;
#else
  static constexpr int BLOCK = ((NP + 31) / 32) * 32;
#endif
};

// ---------------------------------------------------------------------------------------
// Fused DG tendency (+ optional RK stage update).  Template switches select the code that
// is compiled in: NF1 numerical flux, AUX = model has orientation/reference-state columns,
// VISC = second-order fluxes from the gradient-flux array are included.
// ---------------------------------------------------------------------------------------
// Pull [p, p+bytes) into L2, one 128-byte line per thread (no registers or shared memory are
// tied up, nothing waits on it).
template <int BLOCK>
__device__ __forceinline__ void prefetch_l2_range(const void *p, size_t bytes, int tid) {
  const uintptr_t a0 = (uintptr_t)p & ~(uintptr_t)127;
  const int lines = (int)(((uintptr_t)p + bytes - a0 + 127) >> 7);
  for (int l = tid; l < lines; l += BLOCK)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a0 + ((uintptr_t)l << 7)));
}

// Same with one instruction: bulk L2 prefetch of the 16-byte-aligned part of [p, p+bytes)
// (one thread issues it; the first/last <16 bytes share a 128-byte line with their neighbours).
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, size_t bytes) {
  const uintptr_t a0 = ((uintptr_t)p + 15) & ~(uintptr_t)15;
  const uintptr_t a1 = ((uintptr_t)p + bytes) & ~(uintptr_t)15;
  if (a1 > a0)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((unsigned)(a1 - a0)) : "memory");
}

template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(s), "l"(gmem), "n"(BYTES));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;\n" ::: "memory");
}

// Bulk (TMA engine) copy global -> shared with completion on an mbarrier: no registers, no LSU wavefronts.
// dst / src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void mbar_init(unsigned long long *mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(mbar)), "r"(count)
               : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *mbar) {
  const unsigned m = (unsigned)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"(m)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *mbar, unsigned parity) {
  const unsigned m = (unsigned)__cvta_generic_to_shared(mbar);
  unsigned done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(m), "r"(parity)
                 : "memory");
  } while (!done);
}

// Experiment, off by default (-DCMDG_FACE_AUX=1 builds it).  The plus-side geopotential and reference pressure of
// a face node are constants of the grid (aux columns the model never changes after init): instead of gathering
// them through L1 at every launch (2 of the 7 scattered 8-byte cp.async per face node plus their shared-memory
// round trip) they can be read from a packed per-face-node copy, one coalesced 16-byte load (built by
// pack_face_aux_kernel at the first launch after cmdg_bind_state).  Measured on one box (ncu, 61 440 elements,
// profiles/r2_ncu_metrics_dg_tendency_face_aux*.csv): L1 data-pipe wavefronts 108.5 M -> 101.8 M (shared 76.2 M ->
// 68.1 M), DRAM reads +6 %, 926 k -> 919 k cycles (-0.8 %); bench 68.6 -> 69.1 GDOF/s.  The Held-Suarez
// instantiation (VISC, 124 registers) got SLOWER: 0.828 -> 0.871 ms per launch.  The kernel is past the point
// where fewer LSU wavefronts buy time (L1 75 %, DRAM 69 %: latency at 5 blocks per SM), and a cached copy of
// caller-owned aux columns is a semantic liability for a drop-in, so the gathers stay.
#ifndef CMDG_FACE_AUX
#define CMDG_FACE_AUX 0
#endif
template <class R, int NQ, bool AUX, bool VISC>
struct TendSmem {
  static constexpr int NP = Dims<NQ>::NP, NFN = Dims<NQ>::NFN;
  // own state, [5][NP] starting at Qraw[0] or, when the element's first word is only 8-byte aligned in global
  // memory and the state comes in by a bulk copy of the enclosing 16-byte aligned window, at Qraw[1]
  alignas(16) R Qraw[5 * NP + 3];
  unsigned long long mbar;                 // completion barrier of that bulk copy
  // contravariant fluxes M xi_m . F.  F12 is read by (k-plane, state) lanes, F3 by (i-plane,
  // state) lanes; the state stride of F3 is padded to 8 mod 16 + 5 doubles so that both are free of
  // bank conflicts (k-plane lanes: 25 k + 125 s, i-plane lanes: i + 133 s)
  static constexpr int NP3 = (NQ == 5) ? 133 : NP;
  R F12[2][5][NP];
  R F3[5][NP3];
  R Qp[5][NFN];                            // neighbour traces; reused for the face results
  R P[NP], Rinv[NP];                       // own pressure, 1/rho
  R Phi[AUX ? NP : 1], Pref[AUX ? NP : 1]; // own geopotential, reference pressure
  R Ap[2][(AUX && !CMDG_FACE_AUX) ? NFN : 1];   // neighbour geopotential, reference pressure (gather variant)
  alignas(16) R Fnp[VISC ? NFN : 1][4];    // neighbour's n+ . F2+ at my face items (16-byte cp.async)
};

#ifndef CMDG_TEND_MINBLOCKS
#define CMDG_TEND_MINBLOCKS 5
#endif
template <class R> __device__ __forceinline__ R log_(R x);
template <> __device__ __forceinline__ double log_(double x) { return log(x); }
template <> __device__ __forceinline__ float log_(float x) { return logf(x); }
template <class R> __device__ __forceinline__ R exp_(R x);
template <> __device__ __forceinline__ double exp_(double x) { return exp(x); }
template <> __device__ __forceinline__ float exp_(float x) { return expf(x); }
template <class R> __device__ __forceinline__ R sinpi_(R x);
template <> __device__ __forceinline__ double sinpi_(double x) { return sinpi(x); }
template <> __device__ __forceinline__ float sinpi_(float x) { return sinpif(x); }

// HeldSuarezForcing (experiments/AtmosGCM/heldsuarez.jl:112-172) and RayleighSponge
// (src/Atmos/Model/tendencies_momentum.jl:104-137) added to src[1..4].  sin(lat) = x3/|x|
// (Orientations.jl:178-179) is used directly instead of sin(asin(.)).
template <class R>
__device__ __forceinline__ void extended_sources(const AtmosParams<R> &P, const R q[5],
                                                 const Thermo<R> &th, R Phi, const R gPhi[3],
                                                 const R x[3], R src[5]) {
  if (P.sources & SRC_HELD_SUAREZ) {
    const R k_a = P.inv_day / R(40), k_f = P.inv_day, k_s = P.inv_day / R(4);
    const R s2 = x[2] * x[2] / (x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);  // sin^2(lat)
    const R c2 = R(1) - s2;
    const R sigma = th.p / P.MSLP;
    const R lsig = log_<R>(sigma);
    const R exner = exp_<R>(P.kappa * lsig);   // sigma^kappa with the logarithm shared with T_eq
    const R sigma_b = R(7) / R(10);
    const R dsig = (sigma - sigma_b) / (R(1) - sigma_b);
    const R hf = dsig > R(0) ? dsig : R(0);
    R T_eq = (R(315) - R(60) * s2 - R(10) * lsig * c2) * exner;
    T_eq = T_eq > R(200) ? T_eq : R(200);
    const R k_T = k_a + (k_s - k_a) * hf * (c2 * c2);
    const R k_v = k_f * hf;
    const R ig = R(1) / P.grav;
    const R nh[3] = {gPhi[0] * ig, gPhi[1] * ig, gPhi[2] * ig};
    const R nd = nh[0] * q[1] + nh[1] * q[2] + nh[2] * q[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) src[1 + d] += -k_v * (q[1 + d] - nh[d] * nd);
    src[4] += -k_T * q[0] * P.cv_d * (th.T - T_eq);
  }
  if (P.sources & SRC_RAYLEIGH_SPONGE) {
    const R z = Phi / P.grav;
    if (z >= P.sponge_z_sponge) {
      const R r = (z - P.sponge_z_sponge) / (P.sponge_z_max - P.sponge_z_sponge);
      const R sn = sinpi_<R>(r / 2);
      // gamma = 2 in every driver of the reference: a product instead of pow (uniform branch)
      const R beta = P.sponge_alpha_max * (P.sponge_gamma == R(2) ? sn * sn : pow_<R>(sn, P.sponge_gamma));
#pragma unroll
      for (int d = 0; d < 3; ++d) src[1 + d] += -beta * (q[1 + d] - q[0] * P.sponge_u[d]);
    }
  }
}

// xi3 part of the weak derivative for one (i-plane, state): out3[j][c] = sum_n D[n][c] F3[i + Nq j + Nq^2 n],
// in place.  F3 = &S.F3[state][i]; every plane is read and rewritten by its own lane only.
template <class R, int NQ>
__device__ __forceinline__ void contract_xi3(R *F3) {
#pragma unroll
  for (int b = 0; b < NQ; ++b) {
    R f[NQ], o[NQ];
#pragma unroll
    for (int n = 0; n < NQ; ++n) f[n] = F3[NQ * b + NQ * NQ * n];
#pragma unroll
    for (int c = 0; c < NQ; ++c) {
      o[c] = R(0);
#pragma unroll
      for (int n = 0; n < NQ; ++n) o[c] += const_D<R>(n * NQ + c) * f[n];
    }
#pragma unroll
    for (int c = 0; c < NQ; ++c) F3[NQ * b + NQ * NQ * c] = o[c];
  }
}

template <class R, int NQ, int NF1, bool AUX, bool VISC, bool SRCX>
#ifndef CMDG_VISC_MINBLOCKS
#define CMDG_VISC_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(Dims<NQ>::BLOCK, (NQ <= 5 ? (VISC ? CMDG_VISC_MINBLOCKS : CMDG_TEND_MINBLOCKS) : 1))
dg_tendency_kernel(const TendArgs<R> A, const AtmosParams<R> P) {
  constexpr int NP = Dims<NQ>::NP, NFP = Dims<NQ>::NFP, NFN = Dims<NQ>::NFN;
  constexpr int BLOCK = Dims<NQ>::BLOCK;
  constexpr int NWARP = BLOCK / 32;
  constexpr int FSTRIDE = (NWARP - 1) * 32;              // face threads per block
  // Face-aligned items: when a face fits a warp and the faces divide evenly among the face warps, face
  // warp w owns faces w, w + (NWARP-1), ... and lane = face node.  Same number of rounds as packing the
  // 6 Nfp items densely (2 for Nq = 5), but a warp never straddles two faces, which is what made the
  // minus-side reads S.Q[s][vm] (stride 5 / 25 doubles) collide in shared-memory banks: with one face
  // per warp every half-warp access of the face phase is conflict-free except the xi2-faces' first one.
  // Measured (ncu, 61 440 elements, profiles/r2_ncu_full_dg_tendency_baroclinic_face_aligned.txt): excessive
  // shared wavefronts 8.5 M -> 3.8 M (14 % -> 6 % of all), but the second round then runs three 25-lane
  // warps instead of 54 dense lanes: +12 % warp instructions, 945 k -> 963 k cycles.  Net loss, so the
  // dense packing stays the default (-DCMDG_FACE_ALIGNED=1 selects this mapping).
#ifndef CMDG_FACE_ALIGNED
#define CMDG_FACE_ALIGNED 0
#endif
  constexpr bool FALIGN = CMDG_FACE_ALIGNED && (NFP <= 32) && (6 % (NWARP - 1) == 0);
  constexpr int NITEM = FALIGN ? 6 / (NWARP - 1) : (NFN + FSTRIDE - 1) / FSTRIDE;   // face items per face thread
  static_assert(NWARP >= 2 && NQ * 5 <= 32, "plane contraction: one warp holds Nq x 5 (plane, state) pairs");
#ifndef CMDG_SPLIT_XI3
#define CMDG_SPLIT_XI3 1
#endif
  // dense packing, three face warps: the third one has items in the first round only
  constexpr bool SPLIT_XI3 = CMDG_SPLIT_XI3 && !FALIGN && NWARP == 4 && NFN > 2 * 32 && NFN <= 2 * FSTRIDE - 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TendSmem<R, NQ, AUX, VISC> &S = *reinterpret_cast<TendSmem<R, NQ, AUX, VISC> *>(smem_raw);

  const int tid = threadIdx.x;
  const int e = A.elems ? A.elems[blockIdx.x] : blockIdx.x;
  const size_t eoffQ = (size_t)e * P.nstate * NP;
  const size_t eoffA = (size_t)e * P.naux * NP;
  const R *__restrict__ Qg = A.Q;
  const R *__restrict__ auxg = A.aux;

  // ---- own state by ONE bulk copy (TMA engine) straight into shared memory: the five state columns of an element
  // are 5 Np contiguous words, but only 8-byte aligned for every other element -- the copy takes the enclosing
  // 16-byte aligned window (one extra word in front or behind) and the shared-memory view starts at word 0 or 1.
  // Replaces five coalesced LDG + five STS per node thread by five LDS.  Measured (ncu, 61 440 elements, one box,
  // profiles/r2_ncu_metrics_dg_tendency_tma_state*.csv): UBLKCP + SYNCS in the SASS, parity green, L1 data-pipe
  // wavefronts 107.9 M -> 107.1 M, but 926 k -> 939 k cycles and 238.8 M -> 258.1 M warp instructions (the
  // mbarrier try_wait spin and the extra block barrier that publishes the mbarrier), bench 68.4 -> 68.2 GDOF/s:
  // a small net loss, so it is off by default (-DCMDG_TMA_Q=1 builds it).  Nothing else in this kernel can go
  // through the bulk-copy engine with a gain: geometry, aux and the old tendency are consumed in registers (LDS
  // would replace LDG one for one), and tensor-map copies of the face traces are impossible on the contractual
  // layout (Np = 125 words per column: strides of 40 / 200 / 1000 bytes, not multiples of 16).
#ifndef CMDG_TMA_Q
#define CMDG_TMA_Q 0
#endif
  constexpr unsigned QWIN = (unsigned)((5 * NP + 1 + (16 / sizeof(R) - 1)) / (16 / sizeof(R)) * 16);   // bytes
  const int qoff = (int)(eoffQ & (16 / sizeof(R) - 1));      // words in front of the element inside its window
  const bool bulk = CMDG_TMA_Q && (((uintptr_t)Qg & 15) == 0) &&
                    (eoffQ - qoff) * sizeof(R) + QWIN <= (size_t)A.nelem * P.nstate * NP * sizeof(R);
  R *Qs = S.Qraw + (bulk ? qoff : 0);
  if (CMDG_TMA_Q) {
    if (tid == 0) mbar_init(&S.mbar, 1);
    __syncthreads();
    if (bulk && tid == 0) bulk_load_g2s(S.Qraw, Qg + (eoffQ - qoff), QWIN, &S.mbar);
  }

  // ---- (a) face descriptors of my face items ----
  // Warp specialisation: after the node phase one warp (rotating with the block index, so
  // that the four schedulers share that work) contracts the fluxes with D plane by plane while
  // the other warps compute the numerical fluxes of the 6 * Nfp face items.  A face thread
  // gathers (cp.async) the neighbour traces of exactly the items it will later compute.
  const int warp = tid >> 5, lane = tid & 31;
  const int cw = blockIdx.x % NWARP;
  const int fw = warp < cw ? warp : warp - 1;             // face-warp index (unused by the contraction warp)
  const int ft = (warp == cw) ? -1 : (FALIGN ? (lane < NFP ? 0 : -1) : fw * 32 + lane);
  // item of round r: dense (ft + r * FSTRIDE) or face-aligned ((fw + r (NWARP-1)) Nfp + lane)
#define CMDG_ITEM(r) (FALIGN ? ((ft >= 0) ? (fw + (r) * (NWARP - 1)) * NFP + lane : NFN) : ft + (r) * FSTRIDE)
  int2 cn[NITEM];
  unsigned ti[NITEM];   // item info from the index table: vm | f << 8 | fn << 12
  static_assert(NQ == 5, "index tables are built for Nq = 5");
#pragma unroll
  for (int r = 0; r < NITEM; ++r) {
    const int it = CMDG_ITEM(r);
    ti[r] = (ft >= 0 && it < NFN) ? d_tbl5.item[it] : 0u;
    cn[r] = (ft >= 0 && it < NFN) ? A.conn[(size_t)e * 6 + ((ti[r] >> 8) & 15)] : make_int2(0, 0);
    // face geometry of my items is needed only in the face phase: park it in L1 now
    if (ft >= 0 && it < NFN) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(A.sgeoP + ((size_t)e * NFN + it) * 4));
      if (AUX && CMDG_FACE_AUX)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.fauxP + ((size_t)e * NFN + it) * 2));
    }
  }

  // ---- L2 prefetch for the block that will replace this one on the SM: the one-shot kernel
  // is latency-bound (two dependent DRAM round trips per block); with the own-element data
  // already in L2 those become L2 hits and DRAM stays busy during the compute phases ----
  if (A.pf_dist > 0 && blockIdx.x + A.pf_dist < gridDim.x && tid == BLOCK - 1) {
    const int bn = blockIdx.x + A.pf_dist;
    const int en = A.elems ? A.elems[bn] : bn;
    prefetch_l2_bulk(Qg + (size_t)en * P.nstate * NP, 5 * NP * sizeof(R));
    prefetch_l2_bulk(A.vgeoP + (size_t)en * NP * 10, NP * 10 * sizeof(R));
    prefetch_l2_bulk(A.sgeoP + (size_t)en * NFN * 4, NFN * 4 * sizeof(R));
    if (A.beta != R(0)) prefetch_l2_bulk(A.dQ + (size_t)en * P.nstate * NP, 5 * NP * sizeof(R));
    if (AUX && CMDG_FACE_AUX) prefetch_l2_bulk(A.fauxP + (size_t)en * NFN * 2, NFN * 2 * sizeof(R));
    if (AUX) {
      const int lo = P.a_Phi >= 0 ? P.a_Phi : P.a_ref_rho;
      const int hi = P.a_ref_p >= 0 ? P.a_ref_p + 1 : P.a_gradPhi + 3;
      if (lo >= 0 && hi > lo)
        prefetch_l2_bulk(auxg + ((size_t)en * P.naux + lo) * NP, (size_t)(hi - lo) * NP * sizeof(R));
      // HeldSuarezForcing reads the coordinates (latitude)
      if (SRCX) prefetch_l2_bulk(auxg + (size_t)en * P.naux * NP, (size_t)3 * NP * sizeof(R));
    }
    if (VISC) {
      prefetch_l2_bulk(A.F2 + (size_t)en * 12 * NP, (size_t)12 * NP * sizeof(R));
      if (en < A.nreal) prefetch_l2_bulk(A.Fn + (size_t)en * NFN * 4, (size_t)NFN * 4 * sizeof(R));
    }
    asm volatile("prefetch.global.L2 [%0];" ::"l"(A.conn + (size_t)en * 6));
  }
  // ---- tail prefetch for the next launch(es), by the last blocks of this one ----
  if (tid == BLOCK - 2 && A.pfn_n[0] + A.pfn_n[1] > 0) {
    const int back = (int)gridDim.x - 1 - (int)blockIdx.x;     // 0 for the last block
    if (back < A.pfn_n[0] + A.pfn_n[1]) {
      const int li = back < A.pfn_n[0] ? 0 : 1;
      const int j = li == 0 ? back : back - A.pfn_n[0];
      const int en = A.pfn_list[li] ? A.pfn_list[li][j] : j;
      prefetch_l2_bulk(A.pfn_Q + (size_t)en * P.nstate * NP, 5 * NP * sizeof(R));
      prefetch_l2_bulk(A.dQ + (size_t)en * P.nstate * NP, 5 * NP * sizeof(R));
      prefetch_l2_bulk(A.vgeoP + (size_t)en * NP * 10, NP * 10 * sizeof(R));
      prefetch_l2_bulk(A.sgeoP + (size_t)en * NFN * 4, NFN * 4 * sizeof(R));
      if (AUX && CMDG_FACE_AUX) prefetch_l2_bulk(A.fauxP + (size_t)en * NFN * 2, NFN * 2 * sizeof(R));
      if (AUX) {
        const int lo = P.a_Phi >= 0 ? P.a_Phi : P.a_ref_rho;
        const int hi = P.a_ref_p >= 0 ? P.a_ref_p + 1 : P.a_gradPhi + 3;
        if (lo >= 0 && hi > lo)
          prefetch_l2_bulk(auxg + ((size_t)en * P.naux + lo) * NP, (size_t)(hi - lo) * NP * sizeof(R));
      }
    }
  }

  // ---- (b) issue my node's loads ----
  R MI = 0;
  R dQold[5] = {0, 0, 0, 0, 0};
  R q[5] = {1, 0, 0, 0, 0};
  R g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  R Phi = 0, pref = 0, rref = 0;
  R gPhi[3] = {0, 0, 0};
  R xc[3] = {0, 0, 0};
  R f2[12];
  if (tid < NP) {
    if (!bulk) {
#pragma unroll
      for (int s = 0; s < 5; ++s) q[s] = Qg[eoffQ + (size_t)s * NP + tid];
    }
    if (AUX && SRCX) {
#pragma unroll
      for (int d = 0; d < 3; ++d) xc[d] = auxg[eoffA + (size_t)d * NP + tid];
    }
    if (AUX) {
      if (P.a_Phi >= 0) Phi = auxg[eoffA + (size_t)P.a_Phi * NP + tid];
      if (P.a_ref_p >= 0) pref = auxg[eoffA + (size_t)P.a_ref_p * NP + tid];
      if ((P.sources & SRC_GRAVITY) && P.subtract_off)
        rref = auxg[eoffA + (size_t)P.a_ref_rho * NP + tid];
      if (P.a_gradPhi >= 0 && (P.sources & SRC_GRAVITY)) {
#pragma unroll
        for (int d = 0; d < 3; ++d) gPhi[d] = auxg[eoffA + (size_t)(P.a_gradPhi + d) * NP + tid];
      }
    }
    if (VISC) {
      // diffusive flux of my node, evaluated by the gradient kernel (12 coalesced loads)
      const size_t eoffF = (size_t)e * 12 * NP + tid;
#pragma unroll
      for (int c = 0; c < 12; ++c) f2[c] = A.F2[eoffF + (size_t)c * NP];
    }
    load_vgeo<R>(A.vgeoP, e, tid, NP, g, MI);
  }

  // ---- (c) asynchronous gathers of the neighbour traces into shared memory ----
#pragma unroll
  for (int r = 0; r < NITEM; ++r) {
    const int it = CMDG_ITEM(r);
    if (ft >= 0 && it < NFN && ((cn[r].y >> 4) & 15) == 0) {
      const int fn = (int)(ti[r] >> 12);
      const int vp = d_tbl5.vp[cn[r].y & 15][fn];
      const size_t offp = (size_t)cn[r].x * P.nstate * NP + vp;
#pragma unroll
      for (int s = 0; s < 5; ++s) cp_async<sizeof(R)>(&S.Qp[s][it], Qg + offp + (size_t)s * NP);
      if (AUX && !CMDG_FACE_AUX) {
        const size_t offa = (size_t)cn[r].x * P.naux * NP + vp;
        constexpr bool G = AUX && !CMDG_FACE_AUX;
        if (P.a_Phi >= 0) cp_async<sizeof(R)>(&S.Ap[0][G ? it : 0], auxg + offa + (size_t)P.a_Phi * NP);
        if (P.a_ref_p >= 0)
          cp_async<sizeof(R)>(&S.Ap[1][G ? it : 0], auxg + offa + (size_t)P.a_ref_p * NP);
      }
      if (VISC && cn[r].x < A.nreal) {
        // the neighbour's own normal diffusive flux at the matching face node (32 contiguous bytes)
        const int a = (cn[r].y & 8) ? NQ - 1 - fn % NQ : fn % NQ;
        const R *fnp = A.Fn + ((size_t)cn[r].x * NFN + (cn[r].y & 7) * NFP + a + NQ * (fn / NQ)) * 4;
        cp_async<2 * sizeof(R)>(&S.Fnp[VISC ? it : 0][0], fnp);
        cp_async<2 * sizeof(R)>(&S.Fnp[VISC ? it : 0][2], fnp + 2);
      }
    }
  }

  // ---- (d) volume: fluxes at my node ----
  R src[5] = {0, 0, 0, 0, 0};
  if (bulk) mbar_wait(&S.mbar, 0);     // every thread that later reads the state observes the completion itself
  if (tid < NP) {
    if (bulk) {
#pragma unroll
      for (int s = 0; s < 5; ++s) q[s] = Qs[s * NP + tid];
    } else {
#pragma unroll
      for (int s = 0; s < 5; ++s) Qs[s * NP + tid] = q[s];
    }
    const Thermo<R> th = thermo<R>(P, q, Phi);
    const R pflux = (AUX && P.subtract_off) ? th.p - pref : th.p;
    S.P[tid] = th.p;
    S.Rinv[tid] = th.rinv;
    if (AUX) {
      S.Phi[tid] = Phi;
      S.Pref[tid] = pref;
    }
    if (A.aux_out) {
      // kernel_nodal_update_auxiliary_state! / DryModel (moisture.jl:58-69)
      // (p / MSLP)^kappa as exp(kappa log(.)) as in the gradient kernel: half the instructions of pow
      A.aux_out[eoffA + (size_t)P.a_theta_v * NP + tid] = th.T / exp_<R>(P.kappa * log_<R>(th.p / P.MSLP));
      A.aux_out[eoffA + (size_t)P.a_T * NP + tid] = th.T;
    }
    const R u[3] = {q[1] * th.rinv, q[2] * th.rinv, q[3] * th.rinv};
    R F[3][5];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      F[d][0] = q[1 + d];
      F[d][1] = q[1 + d] * u[0];
      F[d][2] = q[1 + d] * u[1];
      F[d][3] = q[1 + d] * u[2];
      F[d][1 + d] += pflux;
      F[d][4] = u[d] * (q[4] + th.p);
    }
    if (VISC) {
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int s = 1; s < 5; ++s) F[d][s] += f2[4 * d + s - 1];
    }
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int s = 0; s < 5; ++s)
      {
        const R v = g[3 * m] * F[0][s] + g[3 * m + 1] * F[1][s] + g[3 * m + 2] * F[2][s];
        if (m < 2) S.F12[m < 2 ? m : 0][s][tid] = v;
        else S.F3[s][tid] = v;
      }
    // sources (tendencies_momentum.jl:66-92); added in the vertical launch in the reference
    if (AUX && (P.sources & SRC_GRAVITY)) {
      const R rr = q[0] - rref;
      src[1] = -rr * gPhi[0];
      src[2] = -rr * gPhi[1];
      src[3] = -rr * gPhi[2];
    }
    if (P.sources & SRC_CORIOLIS) {
      src[1] += P.two_Omega * q[2];
      src[2] -= P.two_Omega * q[1];
    }
    if (AUX && SRCX) {
      if (!(P.sources & SRC_GRAVITY)) {
#pragma unroll
        for (int d = 0; d < 3; ++d) gPhi[d] = auxg[eoffA + (size_t)(P.a_gradPhi + d) * NP + tid];
      }
      extended_sources<R>(P, q, th, Phi, gPhi, xc, src);
    }
  }
  __syncthreads();

  // old tendency: needed only in the final combine, its latency hides behind the rest
  if (tid < NP && A.beta != R(0)) {
#pragma unroll
    for (int s = 0; s < 5; ++s) dQold[s] = A.dQ[eoffQ + (size_t)s * NP + tid];
  }

  // ---- volume: weak derivative  D^T (M xi . F), by the contraction warp ----
  // Lane (k, s) owns the k-plane of state s: the xi1- and xi2-contractions stay inside the plane
  // (each flux value is loaded once and D comes from the constant bank as an immediate
  // operand), only xi3 reads the other planes: 7 shared loads per output instead of 18.  The
  // tendency kernel is bound by the LSU data pipe, so this -- not the flops -- is what counts.
  // The result replaces the F[0] plane the lane has just consumed.
  if (warp == cw) {
    if (lane < NQ * 5) {
      // (1) lane = (k-plane pk, state ps): xi1 and xi2, result in place of the F12[0] plane
      // (2) the same lane as (i-plane pk, state ps): xi3, result in place of the F3 plane.
      // Each plane is read and rewritten by its own lane only: no synchronisation in between.
      const int pk = lane % NQ, ps = lane / NQ;
      R pa[NQ][NQ];
#pragma unroll
      for (int b = 0; b < NQ; ++b)
#pragma unroll
        for (int a = 0; a < NQ; ++a) pa[b][a] = R(0);
      R *F1 = &S.F12[0][ps][NQ * NQ * pk];
      const R *F2 = &S.F12[1][ps][NQ * NQ * pk];
#pragma unroll
      for (int b = 0; b < NQ; ++b) {
        R f[NQ];
#pragma unroll
        for (int n = 0; n < NQ; ++n) f[n] = F1[n + NQ * b];
#pragma unroll
        for (int a = 0; a < NQ; ++a)
#pragma unroll
          for (int n = 0; n < NQ; ++n) pa[b][a] += const_D<R>(n * NQ + a) * f[n];
      }
#pragma unroll
      for (int n = 0; n < NQ; ++n) {
        R f[NQ];
#pragma unroll
        for (int a = 0; a < NQ; ++a) f[a] = F2[a + NQ * n];
#pragma unroll
        for (int b = 0; b < NQ; ++b)
#pragma unroll
          for (int a = 0; a < NQ; ++a) pa[b][a] += const_D<R>(n * NQ + b) * f[a];
      }
#pragma unroll
      for (int b = 0; b < NQ; ++b)
#pragma unroll
        for (int a = 0; a < NQ; ++a) F1[a + NQ * b] = pa[b][a];
      if (!SPLIT_XI3) contract_xi3<R, NQ>(&S.F3[ps][pk]);
    }
  } else {
  // ---- faces: numerical flux at every face node of this element ----
  cp_async_wait_all();
#pragma unroll
  for (int r = 0; r < NITEM; ++r) {
    const int it = CMDG_ITEM(r);
    if (it >= NFN) break;
    const int fn = (int)(ti[r] >> 12);
    const int2 c = cn[r];
    const int bctag = (c.y >> 4) & 15;
    const int vm = (int)(ti[r] & 255u);
    R n[3], sMvMI;
    load_sgeo<R>(A.sgeoP + ((size_t)e * NFN + it) * 4, n, sMvMI);
    typename Vec2<R>::type fa;
    fa.x = fa.y = R(0);
    if (AUX && CMDG_FACE_AUX)
      fa = *reinterpret_cast<const typename Vec2<R>::type *>(A.fauxP + ((size_t)e * NFN + it) * 2);
    // second-order path: my own normal diffusive flux n . F2- (gradient kernel), loaded early
    typename Vec2<R>::type fnm0, fnm1;
    fnm0.x = fnm0.y = fnm1.x = fnm1.y = R(0);
    if (VISC && bctag == 0) {
      const typename Vec2<R>::type *pf =
          reinterpret_cast<const typename Vec2<R>::type *>(A.Fn + ((size_t)e * NFN + it) * 4);
      fnm0 = pf[0];
      fnm1 = pf[1];
    }
    R qm[5], qp[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) qm[s] = Qs[s * NP + vm];
    Thermo<R> tm;
    tm.rinv = S.Rinv[vm];
    tm.p = S.P[vm];
    tm.T = tm.p * tm.rinv / P.R_d;
    const R Phim = AUX ? S.Phi[AUX ? vm : 0] : R(0);
    const R prefm = AUX ? S.Pref[AUX ? vm : 0] : R(0);
    R Phip = Phim, prefp = prefm;
    if (bctag == 0) {
#pragma unroll
      for (int s = 0; s < 5; ++s) qp[s] = S.Qp[s][it];
      if (AUX && CMDG_FACE_AUX) {
        if (P.a_Phi >= 0) Phip = fa.x;
        if (P.a_ref_p >= 0) prefp = fa.y;
      } else if (AUX) {
        constexpr bool G = AUX && !CMDG_FACE_AUX;
        if (P.a_Phi >= 0) Phip = S.Ap[0][G ? it : 0];
        if (P.a_ref_p >= 0) prefp = S.Ap[1][G ? it : 0];
      }
    } else {
      // boundary_state! (src/Atmos/Model/bc_momentum.jl:24-33, 60-70)
#pragma unroll
      for (int s = 0; s < 5; ++s) qp[s] = qm[s];
      const int kind = P.bc_kind[bctag - 1];
      if (kind == BC_FREESLIP) {
        const R run = 2 * (qm[1] * n[0] + qm[2] * n[1] + qm[3] * n[2]);
        qp[1] -= run * n[0];
        qp[2] -= run * n[1];
        qp[3] -= run * n[2];
      } else {
        qp[1] = -qm[1];
        qp[2] = -qm[2];
        qp[3] = -qm[3];
      }
    }
    const Thermo<R> tp = thermo<R>(P, qp, Phip);
    R fm[5], fp[5], unm, unp;
    normal_flux<R>(qm, tm.rinv, (AUX && P.subtract_off) ? tm.p - prefm : tm.p, tm.p, n, fm, unm);
    normal_flux<R>(qp, tp.rinv, (AUX && P.subtract_off) ? tp.p - prefp : tp.p, tp.p, n, fp, unp);
    R fl[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) fl[s] = R(0.5) * (fm[s] + fp[s]);
    if (NF1 == NF_RUSANOV) {
      // sound speeds as x * rsqrt(x) (x > 0): ~40 % fewer instructions than the correctly rounded sqrt, which
      // was the single largest instruction consumer of the kernel (4.9 %); 1-2 ulp in a wave-speed bound
      const R c2m = P.gamma * tm.p * tm.rinv, c2p = P.gamma * P.R_d * tp.T;
      const R cm = c2m * fast_rsqrt_<R>(c2m);
      const R cp = c2p * fast_rsqrt_<R>(c2p);
      const R lam = R(0.5) * fmax(fabs(unm) + cm, fabs(unp) + cp);
#pragma unroll
      for (int s = 0; s < 5; ++s) fl[s] += lam * (qm[s] - qp[s]);
    } else if (NF1 == NF_ROE) {
      R diss[5];
      roe_dissipation<R>(P, n, qm, tm, qp, tp, Phim, diss);
#pragma unroll
      for (int s = 0; s < 5; ++s) fl[s] -= diss[s];
    }
    if (VISC && bctag == 0) {
      // CentralNumericalFluxSecondOrder (NumericalFluxes.jl:668-715): n . (F2- + F2+) / 2 with each
      // side's flux evaluated once per node by the gradient kernel; wall faces carry no diffusive
      // flux for FreeSlip/NoSlip + Insulating (bc_momentum.jl:44-49, bc_energy.jl:12-17)
      R fnp[4];
      if (c.x < A.nreal) {
        // the neighbour stored n+ . F2+ with its own (opposite) normal
#pragma unroll
        for (int s = 0; s < 4; ++s) fnp[s] = -S.Fnp[VISC ? it : 0][s];
      } else {
        // ghost neighbour: its F2 arrived by the halo exchange, contract with my normal
        const int vp = d_tbl5.vp[c.y & 15][fn];
        const R *pg = A.F2 + (size_t)c.x * 12 * NP + vp;
#pragma unroll
        for (int s = 0; s < 4; ++s)
          fnp[s] = n[0] * pg[(size_t)s * NP] + n[1] * pg[(size_t)(4 + s) * NP] + n[2] * pg[(size_t)(8 + s) * NP];
      }
      fl[1] += R(0.5) * (fnm0.x + fnp[0]);
      fl[2] += R(0.5) * (fnm0.y + fnp[1]);
      fl[3] += R(0.5) * (fnm1.x + fnp[2]);
      fl[4] += R(0.5) * (fnm1.y + fnp[3]);
    }
    // stash vMI*sM*F* in place of the neighbour trace (same thread wrote/reads this slot)
#pragma unroll
    for (int s = 0; s < 5; ++s) S.Qp[s][it] = sMvMI * fl[s];
  }
  // the last face warp has one round of face items where the others have two: it takes the xi3 contraction
  // (F3 planes, untouched by anybody else between the two barriers) off the contraction warp, whose dependent
  // shared-load -> DFMA chain was the longest path between the barriers (ncu: 13 % of the warp samples were
  // barrier stalls; warp instructions between the barriers 600 / 500 / 500 / 250 before, 400 / 500 / 500 / 450 now)
  if (SPLIT_XI3 && fw == NWARP - 2 && lane < NQ * 5) contract_xi3<R, NQ>(&S.F3[lane / NQ][lane % NQ]);
  }
  __syncthreads();

  // ---- combine: tendency[vid-] -= vMI sM F*  in face order 1..6, then alpha/beta, RK ----
  if (tid < NP) {
    R acc[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) acc[s] = MI * (S.F12[0][s][tid] + S.F3[s][tid]) + src[s];
    // a node lies on at most one face per direction: three predicated reads instead of six
    const unsigned nt = d_tbl5.node[tid];
    const int it1 = (int)(nt & 255u), it2 = (int)((nt >> 8) & 255u), it3 = (int)((nt >> 16) & 255u);
    if (it1 != 255) {
#pragma unroll
      for (int s = 0; s < 5; ++s) acc[s] -= S.Qp[s][it1];
    }
    if (it2 != 255) {
#pragma unroll
      for (int s = 0; s < 5; ++s) acc[s] -= S.Qp[s][it2];
    }
    if (it3 != 255) {
#pragma unroll
      for (int s = 0; s < 5; ++s) acc[s] -= S.Qp[s][it3];
    }
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      const R d = A.alpha * acc[s] + A.beta * dQold[s];
      A.dQ[eoffQ + (size_t)s * NP + tid] = d;
      if (A.Qout) A.Qout[eoffQ + (size_t)s * NP + tid] = q[s] + A.rkb_dt * d;
    }
  }
  if (A.ext_done != nullptr && (int)blockIdx.x < A.n_signal) {   // uniform per block
    __threadfence();          // my stores (dQ, Qout) are visible device-wide ...
    __syncthreads();          // ... for every thread of the block, before the block is counted
    if (tid == 0) atomicAdd(A.ext_done, 1u);
  }
#undef CMDG_ITEM
}

// Packed plus-side aux constants of every face node (see CMDG_FACE_AUX): one block per real element.
template <class R, int NQ>
__global__ void __launch_bounds__(Dims<NQ>::NFN <= 160 ? 160 : 1024)
pack_face_aux_kernel(const R *__restrict__ aux, const int2 *__restrict__ conn, int naux, int a_Phi, int a_ref_p,
                     R *__restrict__ fauxP) {
  constexpr int NP = Dims<NQ>::NP, NFN = Dims<NQ>::NFN;
  static_assert(NQ == 5, "index tables are built for Nq = 5");
  const int e = blockIdx.x;
  for (int it = threadIdx.x; it < NFN; it += blockDim.x) {
    const unsigned ti = d_tbl5.item[it];
    const int2 c = conn[(size_t)e * 6 + ((ti >> 8) & 15)];
    R phi = R(0), pref = R(0);
    if (((c.y >> 4) & 15) == 0) {
      const int vp = d_tbl5.vp[c.y & 15][ti >> 12];
      const size_t offa = (size_t)c.x * naux * NP + vp;
      if (a_Phi >= 0) phi = aux[offa + (size_t)a_Phi * NP];
      if (a_ref_p >= 0) pref = aux[offa + (size_t)a_ref_p * NP];
    }
    fauxP[((size_t)e * NFN + it) * 2] = phi;
    fauxP[((size_t)e * NFN + it) * 2 + 1] = pref;
  }
}

// ---------------------------------------------------------------------------------------
// courant(local_courant, dg, m, Q, dt, simtime, direction) (SpaceDiscretization.jl:307-365):
// kernel_min_neighbor_distance! (Grids.jl:1219-1336) + kernel_local_courant!
// (DGModel_kernels.jl:3028-3100) + maximum, fused: one block per element, the element maximum is
// folded into *result with an atomic max on the bit pattern (Courant numbers are >= 0).
// kind: 0 advective, 1 nondiffusive, 2 diffusive (src/Atmos/Model/courant.jl:12-86).
// vgeo is the caller's array in the reference layout (x1, x2, x3 = columns 12, 13, 14 of 25).
// ---------------------------------------------------------------------------------------
enum { COURANT_ADVECTIVE = 0, COURANT_NONDIFFUSIVE = 1, COURANT_DIFFUSIVE = 2 };
template <class R, int NQ>
__global__ void __launch_bounds__(Dims<NQ>::BLOCK)
courant_kernel(const R *__restrict__ Q, const R *__restrict__ aux, const R *__restrict__ gradflux,
               const R *__restrict__ vgeo, const AtmosParams<R> P, R dt, int kind, int direction,
               unsigned long long *result) {
  constexpr int NP = Dims<NQ>::NP, BLOCK = Dims<NQ>::BLOCK;
  __shared__ R sx[3][NP];
  __shared__ R red[BLOCK / 32];
  const int tid = threadIdx.x, e = blockIdx.x;
  if (tid < NP) {
#pragma unroll
    for (int d = 0; d < 3; ++d) sx[d][tid] = vgeo[((size_t)e * 25 + 12 + d) * NP + tid];
  }
  __syncthreads();
  R c = R(0);
  if (tid < NP) {
    const int idx[3] = {tid % NQ, (tid / NQ) % NQ, tid / (NQ * NQ)};
    const int stride[3] = {1, NQ, NQ * NQ};
    const bool use[3] = {direction != 2, direction != 2, direction != 1};
    R md = sizeof(R) == 8 ? (R)1.7976931348623157e308 : (R)3.4028234e38f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (!use[d]) continue;
#pragma unroll
      for (int sgn = -1; sgn <= 1; sgn += 2) {
        const int o = idx[d] + sgn;
        if (o < 0 || o >= NQ) continue;
        const int nb = tid + sgn * stride[d];
        const R d0 = sx[0][tid] - sx[0][nb], d1 = sx[1][tid] - sx[1][nb], d2 = sx[2][tid] - sx[2][nb];
        const R dist = sqrt_<R>(d0 * d0 + d1 * d1 + d2 * d2);
        md = dist < md ? dist : md;
      }
    }
    R q[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) q[s] = Q[((size_t)e * P.nstate + s) * NP + tid];
    const size_t oa = (size_t)e * P.naux * NP + tid;
    R k[3] = {0, 0, 0}, gP[3] = {0, 0, 0};
    if (P.a_gradPhi >= 0) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        gP[d] = aux[oa + (size_t)(P.a_gradPhi + d) * NP];
        k[d] = gP[d] / P.grav;
      }
    }
    if (kind == COURANT_DIFFUSIVE) {
      R gf[10];
#pragma unroll
      for (int s = 0; s < 10; ++s)
        gf[s] = (s < P.ngf_dyn) ? gradflux[((size_t)e * P.ngradflux + s) * NP + tid] : R(0);
      const R Delta = P.a_Delta >= 0 ? aux[oa + (size_t)P.a_Delta * NP] : R(0);
      R nu[3];
      turbulence_nu<R>(P, q, gf, gP, Delta, nu);
      R nrm;
      if (P.turbulence != TURB_SMAGORINSKY) {
        nrm = nu[0];
      } else {
        const R dot = nu[0] * k[0] + nu[1] * k[1] + nu[2] * k[2];
        if (direction == 2) nrm = dot;
        else if (direction == 1) {
          const R v0 = nu[0] - dot * k[0], v1 = nu[1] - dot * k[1], v2 = nu[2] - dot * k[2];
          nrm = sqrt_<R>(v0 * v0 + v1 * v1 + v2 * v2);
        } else nrm = sqrt_<R>(nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]);
      }
      c = dt * nrm / (md * md);
    } else {
      R normu;
      const R dot = q[1] * k[0] + q[2] * k[1] + q[3] * k[2];
      if (direction == 2) normu = fabs(dot) / q[0];
      else if (direction == 1) {
        const R v0 = (q[1] - dot * k[0]) / q[0], v1 = (q[2] - dot * k[1]) / q[0], v2 = (q[3] - dot * k[2]) / q[0];
        normu = sqrt_<R>(v0 * v0 + v1 * v1 + v2 * v2);
      } else {
        const R u0 = q[1] / q[0], u1 = q[2] / q[0], u2 = q[3] / q[0];
        normu = sqrt_<R>(u0 * u0 + u1 * u1 + u2 * u2);
      }
      R ss = R(0);
      if (kind == COURANT_NONDIFFUSIVE) {
        const R Phi = P.a_Phi >= 0 ? aux[oa + (size_t)P.a_Phi * NP] : R(0);
        const Thermo<R> th = thermo<R>(P, q, Phi);
        ss = sqrt_<R>(P.gamma * P.R_d * th.T);
      }
      c = dt * (normu + ss) / md;
    }
  }
  // block maximum (NaN-free inputs assumed; a NaN state shows up as NaN through fmax semantics elsewhere)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const R other = __shfl_xor_sync(0xffffffffu, c, o);
    c = other > c ? other : c;
  }
  if ((tid & 31) == 0) red[tid >> 5] = c;
  __syncthreads();
  if (tid == 0) {
    R m = red[0];
#pragma unroll
    for (int w = 1; w < BLOCK / 32; ++w) m = red[w] > m ? red[w] : m;
    atomicMax(result, (unsigned long long)__double_as_longlong((double)m));
  }
}

// ---------------------------------------------------------------------------------------
// Filters.apply! (src/Numerics/Mesh/Filters.jl:440-505, kernel_apply_filter! :651-792) as ONE pass
// over Q: the reference's horizontal launch (xi1, xi2 with filter_matrices[1]) and vertical launch
// (xi3 with filter_matrices[end]) are fused, the state is read and written once.  Targets:
// FilterIndices (bit mask of states) and AtmosFilterPerturbations (src/Atmos/Model/filters.jl:4-48:
// rho - rho_ref, rhoe - rhoe_ref, momentum as is).  Between the two launches the reference adds the
// reference state back and subtracts it again; that rounding is reproduced.
// Wh / Wv: row-major [Nq][Nq] (W[i][n] multiplies the value at n).
// ---------------------------------------------------------------------------------------
enum { FILTER_TARGET_INDICES = 0, FILTER_TARGET_ATMOS_PERTURBATIONS = 1 };
template <class R, int NQ, int MAXS>
__global__ void __launch_bounds__(Dims<NQ>::BLOCK)
filter_kernel(R *__restrict__ Q, const R *__restrict__ aux, const R *__restrict__ Wh,
              const R *__restrict__ Wv, int nstate, int naux, unsigned mask, int target,
              int a_ref_rho, int a_ref_rhoe, int do_h, int do_v, int julia_layout) {
  constexpr int NP = Dims<NQ>::NP;
  __shared__ R s[MAXS][NP];
  __shared__ R sW[2][NQ * NQ];
  const int tid = threadIdx.x, e = blockIdx.x;
  if (tid < NQ * NQ) {
    // row-major W[r][c] at r * Nq + c; the caller's Julia (column-major) matrix has it at r + Nq * c
    const int src = julia_layout ? (tid / NQ) + NQ * (tid % NQ) : tid;
    sW[0][tid] = Wh[src];
    sW[1][tid] = Wv[src];
  }
  const int i = tid % NQ, j = (tid / NQ) % NQ, k = tid / (NQ * NQ);
  R v[MAXS], ref[MAXS];
#pragma unroll
  for (int c = 0; c < MAXS; ++c) v[c] = ref[c] = R(0);
  if (tid < NP) {
    const size_t off = (size_t)e * nstate * NP + tid;
#pragma unroll
    for (int c = 0; c < MAXS; ++c)
      if (c < nstate && ((mask >> c) & 1u)) v[c] = Q[off + (size_t)c * NP];
    if (target == FILTER_TARGET_ATMOS_PERTURBATIONS) {
      const size_t oa = (size_t)e * naux * NP + tid;
      ref[0] = aux[oa + (size_t)a_ref_rho * NP];
      if (MAXS > 4) ref[MAXS > 4 ? 4 : 0] = aux[oa + (size_t)a_ref_rhoe * NP];
    }
#pragma unroll
    for (int c = 0; c < MAXS; ++c) {
      v[c] -= ref[c];
      s[c][tid] = v[c];
    }
  }
  __syncthreads();
  // three tensor-product sweeps; sweep d contracts along xi_{d+1}
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const bool on = d < 2 ? do_h : do_v;
    if (!on) continue;
    if (d == 2 && do_h) {
      // boundary between the reference's two launches: result(+ref) then argument(-ref)
      if (tid < NP) {
#pragma unroll
        for (int c = 0; c < MAXS; ++c) s[c][tid] = (v[c] + ref[c]) - ref[c];
      }
      __syncthreads();
    }
    const R *W = sW[d < 2 ? 0 : 1];
    const int me = d == 0 ? i : (d == 1 ? j : k);
    const int stride = d == 0 ? 1 : (d == 1 ? NQ : NQ * NQ);
    const int base = tid - me * stride;
    if (tid < NP) {
#pragma unroll
      for (int c = 0; c < MAXS; ++c) v[c] = R(0);
#pragma unroll
      for (int n = 0; n < NQ; ++n) {
        const R w = W[me * NQ + n];
#pragma unroll
        for (int c = 0; c < MAXS; ++c) v[c] += w * s[c][base + n * stride];
      }
    }
    __syncthreads();   // every read of s in this sweep is done
    if (d == 0) {
      if (tid < NP) {
#pragma unroll
        for (int c = 0; c < MAXS; ++c) s[c][tid] = v[c];
      }
      __syncthreads();
    }
  }
  if (tid < NP) {
    const size_t off = (size_t)e * nstate * NP + tid;
#pragma unroll
    for (int c = 0; c < MAXS; ++c)
      if (c < nstate && ((mask >> c) & 1u)) Q[off + (size_t)c * NP] = v[c] + ref[c];
  }
}

// ---------------------------------------------------------------------------------------
// Gradient pass: volume_gradients! H/V (DGModel_kernels.jl:934-1328) +
// dgsem_interface_gradients! (:1365-1651) with CentralNumericalFluxGradient
// (NumericalFluxes.jl:65-123), fused per element.  Writes the gradient-flux array.
//   G  = (u1,u2,u3,h_tot[,theta_v])          (AtmosModel.jl:622-673)
//   GF = (grad h_tot, sym(grad u)[, N2])      (AtmosModel.jl:675-744)
// ---------------------------------------------------------------------------------------
template <class R>
struct GradArgs {
  const R *Q, *aux;
  R *gradflux;
  const R *vgeoP, *sgeoP;
  const int2 *conn;
  const int *elems;
  const R *D;
  // outputs for the tendency kernel (private): diffusive flux per node and its normal component at
  // the element's own face nodes (NULL = not wanted)
  R *F2;   // [nelem][12][Np]
  R *Fn;   // [nreal][6*Nfp][4]
  // DryBiharmonic (HYPER kernels): gradient of (u_h, h_tot), column 3*s + d  [nelem][12][Np]
  R *Qhg;
  int pf_dist;   // L2 prefetch distance in launch-list entries (0 = off)
  // passive tracers: diagonal of the turbulent viscosity tensor per node, [nreal][3][Np] (NULL = not wanted)
  R *Nu;
};

template <class R>
__device__ __forceinline__ void gradient_argument(const AtmosParams<R> &P, const R q[5], R Phi,
                                                  R G[5]) {
  const Thermo<R> th = thermo<R>(P, q, Phi);
  G[0] = th.rinv * q[1];
  G[1] = th.rinv * q[2];
  G[2] = th.rinv * q[3];
  G[3] = q[4] * th.rinv + P.R_d * th.T;
  // aux.moisture.theta_v, refreshed from Q; only the Smagorinsky closure differentiates it.
  // (p / MSLP)^kappa as exp(kappa log(.)): half the instructions of pow, same value to 1e-16
  G[4] = (P.turbulence == TURB_SMAGORINSKY)
             ? th.T / exp_<R>(P.kappa * log_<R>(th.p / P.MSLP))
             : R(0);
}

// gf = linear map of the gradient  dG[d][g]  (TurbulenceClosures.jl:351-362,456-470)
template <class R>
__device__ __forceinline__ void gradient_flux(const AtmosParams<R> &P, const R dG[3][5],
                                              const R gradPhi[3], R theta_v, R gf[10]) {
  gf[0] = dG[0][3];
  gf[1] = dG[1][3];
  gf[2] = dG[2][3];
  gf[3] = dG[0][0];
  gf[4] = (dG[1][0] + dG[0][1]) * R(0.5);
  gf[5] = (dG[2][0] + dG[0][2]) * R(0.5);
  gf[6] = dG[1][1];
  gf[7] = (dG[2][1] + dG[1][2]) * R(0.5);
  gf[8] = dG[2][2];
  gf[9] = (P.turbulence == TURB_SMAGORINSKY)
              ? (dG[0][4] * gradPhi[0] + dG[1][4] * gradPhi[1] + dG[2][4] * gradPhi[2]) / theta_v
              : R(0);
}

#ifndef CMDG_GRAD_MINBLOCKS
#define CMDG_GRAD_MINBLOCKS 5
#endif
// HYPER (DryBiharmonic, TurbulenceClosures.jl:793-848): the gradient argument grows by
// u_h = (I - k k') u and h_tot (compute_gradient_argument!, :813-826); their horizontal gradients
// (volume + central face term, DGModel_kernels.jl:1081-1098, 1618-1628) go to Qhg.  grad h_tot is
// the first three gradient-flux columns, so only u_h costs extra contractions.  The diffusive flux is
// then assembled by hyper_flux_kernel, not here.
// Fn[e][it][0..3] = n . F2[.][1..4] at the <= 3 face nodes that coincide with volume node (i, j, k)
template <class R, int NQ>
__device__ __forceinline__ void write_normal_flux(const R *__restrict__ sgeoP, R *__restrict__ Fn, int e,
                                                  int i, int j, int k, const R F2[3][5]) {
  constexpr int NFP = Dims<NQ>::NFP, NFN = Dims<NQ>::NFN;
  typedef typename Vec2<R>::type V2;
  const int fit[3] = {(i == 0) ? j + NQ * k : ((i == NQ - 1) ? NFP + j + NQ * k : -1),
                      (j == 0) ? 2 * NFP + i + NQ * k : ((j == NQ - 1) ? 3 * NFP + i + NQ * k : -1),
                      (k == 0) ? 4 * NFP + i + NQ * j : ((k == NQ - 1) ? 5 * NFP + i + NQ * j : -1)};
#pragma unroll
  for (int dir = 0; dir < 3; ++dir) {
    const int it = fit[dir];
    if (it < 0) continue;
    R n[3], sMvMI;
    load_sgeo<R>(sgeoP + ((size_t)e * NFN + it) * 4, n, sMvMI);
    V2 o0, o1;
    o0.x = n[0] * F2[0][1] + n[1] * F2[1][1] + n[2] * F2[2][1];
    o0.y = n[0] * F2[0][2] + n[1] * F2[1][2] + n[2] * F2[2][2];
    o1.x = n[0] * F2[0][3] + n[1] * F2[1][3] + n[2] * F2[2][3];
    o1.y = n[0] * F2[0][4] + n[1] * F2[1][4] + n[2] * F2[2][4];
    V2 *po = reinterpret_cast<V2 *>(Fn + ((size_t)e * NFN + it) * 4);
    po[0] = o0;
    po[1] = o1;
  }
}

// The face term is linear in (G* - G-), so a face item only stores w = vMI sM (G* - G-) (5 values, in
// place of the neighbour trace it consumed); the node thread accumulates n (x) w of its <= 3 faces
// onto its volume gradient and applies the gradient-flux map ONCE -- half the shared-memory traffic
// of storing / re-reading 10 flux columns per face node (the kernel is LSU-bound like the tendency
// kernel).  n . F2 at the element's own face nodes is written by the node threads too (no staging).
// Round-2 experiment (profiles/r2_ncu_metrics_dg_gradient_*.csv, 61 440 elements, Held-Suarez): the plane-lane
// contraction + warp specialisation of dg_tendency_kernel was ported to this kernel (one rotating warp
// differentiates G plane by plane while three warps do the face items).  It cut the L1 data-pipe wavefronts
// (106.9 M -> 95.9 M) and the instructions (379.6 M -> 363.8 M) but not the time: 1.358 M -> 1.386 M cycles.
// Unlike the tendency kernel this one is not bound by the LSU pipe (L1 47-54 %, DRAM 44 %) but by latency at
// 4.7 warps per scheduler and ~6000 warp instructions per element (theta_v = T (MSLP / p)^kappa at 225 points,
// the Smagorinsky closure, n . F2 on up to three faces per node); serialising the contraction in one warp
// lengthens the block's critical path by as much as the saved shared-memory traffic shortens it.  The per-node
// contraction below therefore stays.
template <class R, int NQ, bool AUX, bool HYPER>
struct GradSmem {
  static constexpr int NP = Dims<NQ>::NP, NFN = Dims<NQ>::NFN;
  R G[5][NP];                         // gradient arguments at my nodes
  R Q[5][NP];                         // my state (wall boundary states only)
  R Phi[AUX ? NP : 1];
  R Qp[6][NFN];                       // neighbour traces (Q+, Phi+), gathered asynchronously; then w
  R Uh[HYPER ? 3 : 1][HYPER ? NP : 1];     // u_h at my nodes
  R K[HYPER ? 3 : 1][HYPER ? NP : 1];      // k = grad Phi / grav at my nodes
  R Kp[HYPER ? 3 : 1][HYPER ? NFN : 1];    // neighbour grad Phi; then w of u_h
};

template <class R, int NQ, bool AUX, bool HYPER>
__global__ void __launch_bounds__(Dims<NQ>::BLOCK, (NQ <= 5 ? (HYPER ? 4 : CMDG_GRAD_MINBLOCKS) : 1))
dg_gradient_kernel(const GradArgs<R> A, const AtmosParams<R> P) {
  constexpr int NP = Dims<NQ>::NP, NFP = Dims<NQ>::NFP, NFN = Dims<NQ>::NFN;
  constexpr int BLOCK = Dims<NQ>::BLOCK;
  typedef typename Vec2<R>::type V2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GradSmem<R, NQ, AUX, HYPER> &S = *reinterpret_cast<GradSmem<R, NQ, AUX, HYPER> *>(smem_raw);
  R(&sG)[5][NP] = S.G;
  R(&sQ)[5][NP] = S.Q;
  R(&sPhi)[AUX ? NP : 1] = S.Phi;
  R(&sQp)[6][NFN] = S.Qp;
  constexpr int NITEM = (NFN + BLOCK - 1) / BLOCK;
  const int tid = threadIdx.x;
  const int e = A.elems ? A.elems[blockIdx.x] : blockIdx.x;
  const size_t eoffQ = (size_t)e * P.nstate * NP;
  const size_t eoffA = (size_t)e * P.naux * NP;
  const int nfaces = P.horizontal_diffusion ? 4 : 6;
  const bool smag = P.turbulence == TURB_SMAGORINSKY;
  // face descriptors of my items, then cp.async gathers of the neighbour state: their latency
  // overlaps the node phase (the same thread consumes what it gathered: no barrier needed)
  int2 cn[NITEM];
#pragma unroll
  for (int r = 0; r < NITEM; ++r) {
    const int it = tid + r * BLOCK;
    cn[r] = (it < nfaces * NFP) ? A.conn[(size_t)e * 6 + it / NFP] : make_int2(0, 16);
    if (it < NFN) asm volatile("prefetch.global.L1 [%0];" ::"l"(A.sgeoP + ((size_t)e * NFN + it) * 4));
  }
  // L2 prefetch for the block that will replace this one on the SM (see dg_tendency_kernel)
  if (A.pf_dist > 0 && blockIdx.x + A.pf_dist < gridDim.x && tid == BLOCK - 1) {
    const int bn = blockIdx.x + A.pf_dist;
    const int en = A.elems ? A.elems[bn] : bn;
    prefetch_l2_bulk(A.Q + (size_t)en * P.nstate * NP, 5 * NP * sizeof(R));
    prefetch_l2_bulk(A.vgeoP + (size_t)en * NP * 10, NP * 10 * sizeof(R));
    prefetch_l2_bulk(A.sgeoP + (size_t)en * NFN * 4, (size_t)NFN * 4 * sizeof(R));
    if (AUX && P.a_Phi >= 0) {
      // Phi, and grad Phi (the three columns after it) when the closure / u_h projection needs it
      const int ncol = (HYPER || smag) ? 4 : 1;
      prefetch_l2_bulk(A.aux + ((size_t)en * P.naux + P.a_Phi) * NP, (size_t)ncol * NP * sizeof(R));
      if (P.a_Delta >= 0) prefetch_l2_bulk(A.aux + ((size_t)en * P.naux + P.a_Delta) * NP, NP * sizeof(R));
    }
    asm volatile("prefetch.global.L2 [%0];" ::"l"(A.conn + (size_t)en * 6));
  }
#pragma unroll
  for (int r = 0; r < NITEM; ++r) {
    const int it = tid + r * BLOCK;
    if (it < nfaces * NFP && ((cn[r].y >> 4) & 15) == 0) {
      const int fn = it % NFP;
      int a = fn % NQ;
      const int b = fn / NQ;
      if (cn[r].y & 8) a = NQ - 1 - a;
      const int vp = face_to_vol<NQ>(cn[r].y & 7, a, b);
      const size_t offp = (size_t)cn[r].x * P.nstate * NP + vp;
#pragma unroll
      for (int s = 0; s < 5; ++s) cp_async<sizeof(R)>(&sQp[s][it], A.Q + offp + (size_t)s * NP);
      if (AUX && P.a_Phi >= 0)
        cp_async<sizeof(R)>(&sQp[5][it], A.aux + (size_t)cn[r].x * P.naux * NP + (size_t)P.a_Phi * NP + vp);
      if (HYPER) {
        // the neighbour's own grad Phi (element-local derivative: differs from mine at truncation level)
#pragma unroll
        for (int d = 0; d < 3; ++d)
          cp_async<sizeof(R)>(&S.Kp[HYPER ? d : 0][HYPER ? it : 0],
                              A.aux + (size_t)cn[r].x * P.naux * NP + (size_t)(P.a_gradPhi + d) * NP + vp);
      }
    }
  }

  R q[5] = {1, 0, 0, 0, 0}, G[5], Phi = 0, gPhi[3] = {0, 0, 0}, Delta = 0;
  R g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, MI = 0;
  const R inv_grav = R(1) / P.grav;
  if (tid < NP) {
#pragma unroll
    for (int s = 0; s < 5; ++s) q[s] = A.Q[eoffQ + (size_t)s * NP + tid];
    if (AUX && P.a_Phi >= 0) Phi = A.aux[eoffA + (size_t)P.a_Phi * NP + tid];
    if (AUX && P.a_gradPhi >= 0 && (HYPER || smag)) {
#pragma unroll
      for (int d = 0; d < 3; ++d) gPhi[d] = A.aux[eoffA + (size_t)(P.a_gradPhi + d) * NP + tid];
    }
    if (!HYPER && AUX && P.a_Delta >= 0 && A.F2) Delta = A.aux[eoffA + (size_t)P.a_Delta * NP + tid];
    load_vgeo<R>(A.vgeoP, e, tid, NP, g, MI);
    gradient_argument<R>(P, q, Phi, G);
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      sG[s][tid] = G[s];
      sQ[s][tid] = q[s];
    }
    if (AUX) sPhi[tid] = Phi;
    if (HYPER) {
      const R k[3] = {gPhi[0] * inv_grav, gPhi[1] * inv_grav, gPhi[2] * inv_grav};
      const R ku = k[0] * G[0] + k[1] * G[1] + k[2] * G[2];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        S.K[HYPER ? d : 0][HYPER ? tid : 0] = k[d];
        S.Uh[HYPER ? d : 0][HYPER ? tid : 0] = G[d] - k[d] * ku;
      }
    }
  }
  __syncthreads();

  // faces: w = vMI sM (G* - G-),  G* = (G+ + G-)/2 or g(boundary_state(Q-))
  cp_async_wait_all();
#pragma unroll
  for (int r = 0; r < NITEM; ++r) {
    const int it = tid + r * BLOCK;
    if (it >= nfaces * NFP) break;
    const int f = it / NFP, fn = it - f * NFP;
    const int2 c = cn[r];
    const int bctag = (c.y >> 4) & 15;
    const int vm = face_to_vol<NQ>(f, fn % NQ, fn / NQ);
    const V2 *pg = reinterpret_cast<const V2 *>(A.sgeoP + ((size_t)e * NFN + it) * 4);
    const V2 g1 = pg[1];            // n3, sM vMI
    const R sMvMI = g1.y;
    R Gm[5], qp[5], Gs[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) Gm[s] = sG[s][vm];
    R uhm[3] = {0, 0, 0}, dUh[3] = {0, 0, 0};
    if (HYPER) {
#pragma unroll
      for (int d = 0; d < 3; ++d) uhm[d] = S.Uh[HYPER ? d : 0][HYPER ? vm : 0];
    }
    if (bctag == 0) {
#pragma unroll
      for (int s = 0; s < 5; ++s) qp[s] = sQp[s][it];
      R Phip = 0;
      if (AUX && P.a_Phi >= 0) Phip = sQp[5][it];
      gradient_argument<R>(P, qp, Phip, Gs);
      if (HYPER) {
        const R kp[3] = {S.Kp[0][HYPER ? it : 0] * inv_grav, S.Kp[HYPER ? 1 : 0][HYPER ? it : 0] * inv_grav,
                         S.Kp[HYPER ? 2 : 0][HYPER ? it : 0] * inv_grav};
        const R ku = kp[0] * Gs[0] + kp[1] * Gs[1] + kp[2] * Gs[2];
#pragma unroll
        for (int d = 0; d < 3; ++d) dUh[d] = R(0.5) * ((Gs[d] - kp[d] * ku) + uhm[d]) - uhm[d];
      }
#pragma unroll
      for (int s = 0; s < 5; ++s) Gs[s] = R(0.5) * (Gs[s] + Gm[s]) - Gm[s];
    } else {
      // gradient-flux boundary state (bc_momentum.jl:34-43, 71-80)
      const V2 g0 = pg[0];
      const R n[3] = {g0.x, g0.y, g1.x};
#pragma unroll
      for (int s = 0; s < 5; ++s) qp[s] = sQ[s][vm];
      const R Phim = AUX ? sPhi[AUX ? vm : 0] : R(0);
      const int kind = P.bc_kind[bctag - 1];
      if (kind == BC_FREESLIP) {
        const R run = qp[1] * n[0] + qp[2] * n[1] + qp[3] * n[2];
        qp[1] -= run * n[0];
        qp[2] -= run * n[1];
        qp[3] -= run * n[2];
      } else {
        qp[1] = qp[2] = qp[3] = R(0);
      }
      gradient_argument<R>(P, qp, Phim, Gs);
      if (HYPER) {
        // the boundary state keeps the minus-side auxiliary state: same k
        const R km[3] = {S.K[0][HYPER ? vm : 0], S.K[HYPER ? 1 : 0][HYPER ? vm : 0], S.K[HYPER ? 2 : 0][HYPER ? vm : 0]};
        const R ku = km[0] * Gs[0] + km[1] * Gs[1] + km[2] * Gs[2];
#pragma unroll
        for (int d = 0; d < 3; ++d) dUh[d] = (Gs[d] - km[d] * ku) - uhm[d];
      }
#pragma unroll
      for (int s = 0; s < 5; ++s) Gs[s] -= Gm[s];
    }
    // in place of the trace this thread gathered (nobody else touches the slot)
#pragma unroll
    for (int s = 0; s < 5; ++s)
      if (s < 4 || smag) sQp[s][it] = sMvMI * Gs[s];
    if (HYPER) {
#pragma unroll
      for (int d = 0; d < 3; ++d) S.Kp[HYPER ? d : 0][HYPER ? it : 0] = sMvMI * dUh[d];
    }
  }

  // volume: strong-form gradient  xi_x * (D G)
  R dG[3][5];
  R dH[HYPER ? 3 : 1][3];   // dH[d][c] = d u_h,c / d x_d
  const int i = tid % NQ, j = (tid / NQ) % NQ, k = tid / (NQ * NQ);
  if (tid < NP) {
#pragma unroll
    for (int c = 0; c < 9; ++c) g[c] *= MI;  // packed copy holds M*xi_x
    R G1[5] = {0, 0, 0, 0, 0}, G2[5] = {0, 0, 0, 0, 0}, G3[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
      const R d1 = const_D<R>(i * NQ + n), d2 = const_D<R>(j * NQ + n), d3 = const_D<R>(k * NQ + n);
      const int o1 = n + NQ * (j + NQ * k), o2 = i + NQ * (n + NQ * k), o3 = i + NQ * (j + NQ * n);
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        if (s == 4 && !smag) continue;
        G1[s] += d1 * sG[s][o1];
        G2[s] += d2 * sG[s][o2];
        if (!P.horizontal_diffusion) G3[s] += d3 * sG[s][o3];
      }
    }
    const R vfac = P.horizontal_diffusion ? R(0) : R(1);
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int s = 0; s < 5; ++s)
        dG[d][s] = g[d] * G1[s] + g[3 + d] * G2[s] + vfac * (g[6 + d] * G3[s]);
    if (HYPER) {
      // horizontal gradient of u_h (the diffusion direction is horizontal with DryBiharmonic)
      R H1[3] = {0, 0, 0}, H2[3] = {0, 0, 0};
#pragma unroll
      for (int n = 0; n < NQ; ++n) {
        const R d1 = const_D<R>(i * NQ + n), d2 = const_D<R>(j * NQ + n);
        const int o1 = n + NQ * (j + NQ * k), o2 = i + NQ * (n + NQ * k);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          H1[c] += d1 * S.Uh[HYPER ? c : 0][HYPER ? o1 : 0];
          H2[c] += d2 * S.Uh[HYPER ? c : 0][HYPER ? o2 : 0];
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 0; d < 3; ++d) dH[HYPER ? d : 0][c] = g[d] * H1[c] + g[3 + d] * H2[c];
    }
  }
  __syncthreads();
  if (tid >= NP) return;
  // my <= 3 faces: item index per direction (-1 = interior node in that direction)
  const int fit[3] = {(i == 0) ? j + NQ * k : ((i == NQ - 1) ? NFP + j + NQ * k : -1),
                      (j == 0) ? 2 * NFP + i + NQ * k : ((j == NQ - 1) ? 3 * NFP + i + NQ * k : -1),
                      P.horizontal_diffusion ? -1
                                             : ((k == 0) ? 4 * NFP + i + NQ * j
                                                         : ((k == NQ - 1) ? 5 * NFP + i + NQ * j : -1))};
#pragma unroll
  for (int dir = 0; dir < 3; ++dir) {
    const int it = fit[dir];
    if (it < 0) continue;
    R n[3], sMvMI;
    load_sgeo<R>(A.sgeoP + ((size_t)e * NFN + it) * 4, n, sMvMI);
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      if (s == 4 && !smag) continue;
      const R w = sQp[s][it];
#pragma unroll
      for (int d = 0; d < 3; ++d) dG[d][s] += n[d] * w;
    }
    if (HYPER) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const R w = S.Kp[HYPER ? c : 0][HYPER ? it : 0];
#pragma unroll
        for (int d = 0; d < 3; ++d) dH[HYPER ? d : 0][c] += n[d] * w;
      }
    }
  }
  R gfv[10];
  gradient_flux<R>(P, dG, gPhi, G[4], gfv);
  if (A.gradflux) {
    const size_t eoffG = (size_t)e * P.ngradflux * NP + tid;
#pragma unroll
    for (int s = 0; s < 10; ++s)
      if (s < P.ngf_dyn) A.gradflux[eoffG + (size_t)s * NP] = gfv[s];
  }
  if (HYPER) {
    const size_t eoffH = (size_t)e * 12 * NP + tid;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int d = 0; d < 3; ++d) A.Qhg[eoffH + (size_t)(3 * c + d) * NP] = dH[HYPER ? d : 0][c];
#pragma unroll
    for (int d = 0; d < 3; ++d) A.Qhg[eoffH + (size_t)(9 + d) * NP] = gfv[d];
    return;
  }
  if (!A.F2) return;
  // ---- second-order flux F2(Q, GF, aux) of this node, once (flux_second_order!, kernels.jl:84-105):
  // the tendency kernel then needs no closure evaluation, neither in the volume nor on faces.
  // F2 goes to global memory (volume term, ghost exchange) and, contracted with the face normals of
  // the <= 3 faces this node lies on, to Fn.
  R F2[3][5];
  flux_second_order<R>(P, q, gfv, gPhi, Delta, F2);
  const size_t eoffF = (size_t)e * 12 * NP + tid;
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int c = 1; c < 5; ++c) A.F2[eoffF + (size_t)(4 * d + c - 1) * NP] = F2[d][c];
  write_normal_flux<R, NQ>(A.sgeoP, A.Fn, e, i, j, k, F2);
  if (A.Nu) {
    // D_t = nu / Pr_t of this node for the tracer diffusion (tracer_gradient_kernel)
    R nu[3];
    turbulence_nu<R>(P, q, gfv, gPhi, Delta, nu);
#pragma unroll
    for (int d = 0; d < 3; ++d) A.Nu[((size_t)e * 3 + d) * NP + tid] = nu[d];
  }
}

// ---------------------------------------------------------------------------------------
// DryBiharmonic hyperdiffusion, horizontal direction (schedule DGModel.jl:226-310).  The reference
// runs volume_divergence_of_gradients! (DGModel_kernels.jl:2132-2228) + interface_divergence_of_
// gradients! (:2359-2490, CentralNumericalFluxDivergence, NumericalFluxes.jl:716-770), then
// volume_gradients_of_laplacians! (:2521-2680) + interface_gradients_of_laplacians! (:2860-3026,
// CentralNumericalFluxHigherOrder, NumericalFluxes.jl:772-835) and evaluates flux_second_order!
// of the hyperdiffusive state again in the volume and on both sides of every face.  Here: one
// kernel per pass (volume + the four horizontal faces fused per element), and the second one
// finishes with the TOTAL diffusive flux F2 = F2(viscous) + F2(hyper) per node plus its normal
// component at the element's own face nodes, which is all the tendency kernel consumes -- the
// hyperdiffusive state itself never goes to memory and its halo exchange is replaced by the F2 one.
// ---------------------------------------------------------------------------------------
template <class R>
struct HyperArgs {
  const R *Q, *aux, *gradflux;
  const R *vgeoP, *sgeoP;
  const int2 *conn;
  const int *elems;
  R *Qhg;   // [nelem][12][Np] gradients, column 3*s + d (s: u_h1..3, h_tot)
  R *Qhd;   // [nelem][4][Np]  horizontal Laplacians
  R *F2;    // [nelem][12][Np]
  R *Fn;    // [nreal][6*Nfp][4]
  int pf_dist;   // L2 prefetch distance in launch-list entries (0 = off)
};

// Qhd[s] = -MI D^T (M xi_h . grad G_s) + sum_{f<4} vMI sM (grad+ + grad-) . n / 2
template <class R, int NQ>
__global__ void __launch_bounds__(Dims<NQ>::BLOCK, (NQ <= 5 ? 4 : 1))
hyper_divergence_kernel(const HyperArgs<R> A) {
  constexpr int NP = Dims<NQ>::NP, NFP = Dims<NQ>::NFP, NFN = Dims<NQ>::NFN;
  constexpr int BLOCK = Dims<NQ>::BLOCK, NH = 4 * NFP;
  static_assert(NH <= BLOCK, "one horizontal face node per thread");
  __shared__ R sGr[12][NP];     // my gradients
  __shared__ R sS[2][4][NP];    // M xi_m . grad G_s, m = 1, 2
  __shared__ R sGp[12][NH];     // neighbour gradients at my horizontal face nodes
  __shared__ R sFace[4][NH];
  const int tid = threadIdx.x;
  const int e = A.elems ? A.elems[blockIdx.x] : blockIdx.x;
  int2 cn = make_int2(0, 16);
  if (A.pf_dist > 0 && blockIdx.x + A.pf_dist < gridDim.x && tid == BLOCK - 1) {
    const int bn = blockIdx.x + A.pf_dist;
    const int en = A.elems ? A.elems[bn] : bn;
    prefetch_l2_bulk(A.Qhg + (size_t)en * 12 * NP, 12 * NP * sizeof(R));
    prefetch_l2_bulk(A.vgeoP + (size_t)en * NP * 10, NP * 10 * sizeof(R));
    prefetch_l2_bulk(A.sgeoP + (size_t)en * NFN * 4, (size_t)NH * 4 * sizeof(R));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(A.conn + (size_t)en * 6));
  }
  if (tid < NH) {
    cn = A.conn[(size_t)e * 6 + tid / NFP];
    asm volatile("prefetch.global.L1 [%0];" ::"l"(A.sgeoP + ((size_t)e * NFN + tid) * 4));
    if (((cn.y >> 4) & 15) == 0) {
      const int fn = tid % NFP;
      int a = fn % NQ;
      const int b = fn / NQ;
      if (cn.y & 8) a = NQ - 1 - a;
      const int vp = face_to_vol<NQ>(cn.y & 7, a, b);
      const R *pg = A.Qhg + (size_t)cn.x * 12 * NP + vp;
#pragma unroll
      for (int c = 0; c < 12; ++c) cp_async<sizeof(R)>(&sGp[c][tid], pg + (size_t)c * NP);
    }
  }
  R MI = 0;
  if (tid < NP) {
    R g[9], gr[12];
    const size_t eo = (size_t)e * 12 * NP + tid;
#pragma unroll
    for (int c = 0; c < 12; ++c) gr[c] = A.Qhg[eo + (size_t)c * NP];
    load_vgeo<R>(A.vgeoP, e, tid, NP, g, MI);
#pragma unroll
    for (int c = 0; c < 12; ++c) sGr[c][tid] = gr[c];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int s = 0; s < 4; ++s)
        sS[m][s][tid] = g[3 * m] * gr[3 * s] + g[3 * m + 1] * gr[3 * s + 1] + g[3 * m + 2] * gr[3 * s + 2];
  }
  cp_async_wait_all();
  __syncthreads();
  if (tid < NH) {
    const int f = tid / NFP, fn = tid - f * NFP;
    const int vm = face_to_vol<NQ>(f, fn % NQ, fn / NQ);
    const bool wall = ((cn.y >> 4) & 15) != 0;
    R n[3], sMvMI;
    load_sgeo<R>(A.sgeoP + ((size_t)e * NFN + tid) * 4, n, sMvMI);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      R l = R(0);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const R gm = sGr[3 * s + d][vm];
        // walls: boundary_state! of the divergence flux is a no-op, grad+ = grad-
        const R gp = wall ? gm : sGp[3 * s + d][tid];
        l += (gp + gm) * (n[d] * R(0.5));
      }
      sFace[s][tid] = sMvMI * l;
    }
  }
  R div[4] = {0, 0, 0, 0};
  const int i = tid % NQ, j = (tid / NQ) % NQ, k = tid / (NQ * NQ);
  if (tid < NP) {
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
      const R d1 = const_D<R>(n * NQ + i), d2 = const_D<R>(n * NQ + j);
      const int o1 = n + NQ * (j + NQ * k), o2 = i + NQ * (n + NQ * k);
#pragma unroll
      for (int s = 0; s < 4; ++s) div[s] += d1 * sS[0][s][o1] + d2 * sS[1][s][o2];
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) div[s] *= -MI;
  }
  __syncthreads();
  if (tid < NP) {
    const int f1 = (i == 0) ? j + NQ * k : ((i == NQ - 1) ? NFP + j + NQ * k : -1);
    const int f2 = (j == 0) ? 2 * NFP + i + NQ * k : ((j == NQ - 1) ? 3 * NFP + i + NQ * k : -1);
    if (f1 >= 0) {
#pragma unroll
      for (int s = 0; s < 4; ++s) div[s] += sFace[s][f1];
    }
    if (f2 >= 0) {
#pragma unroll
      for (int s = 0; s < 4; ++s) div[s] += sFace[s][f2];
    }
    const size_t eo = (size_t)e * 4 * NP + tid;
#pragma unroll
    for (int s = 0; s < 4; ++s) A.Qhd[eo + (size_t)s * NP] = div[s];
  }
}

// H[3 s + d] = nu4 (xi_h,d D lap_s) + sum_{f<4} vMI sM nu4 n_d (lap+ - lap-) / 2
// (transform_post_gradient_laplacian!, TurbulenceClosures.jl:828-848: nu4 = (Delta_h / 2)^4 / 2 / tau),
// then F2 = flux_second_order(viscous) + HyperdiffViscousFlux / HyperdiffEnthalpyFlux
// (tendencies_momentum.jl:51-54, tendencies_energy.jl:40-48) and Fn = n . F2 on my faces.
template <class R, int NQ, bool AUX>
__global__ void __launch_bounds__(Dims<NQ>::BLOCK, (NQ <= 5 ? 4 : 1))
hyper_flux_kernel(const HyperArgs<R> A, const AtmosParams<R> P) {
  constexpr int NP = Dims<NQ>::NP, NFP = Dims<NQ>::NFP, NFN = Dims<NQ>::NFN;
  constexpr int BLOCK = Dims<NQ>::BLOCK, NH = 4 * NFP;
  typedef typename Vec2<R>::type V2;
  __shared__ R sL[4][NP];
  __shared__ R sNu[NP];
  __shared__ R sLp[4][NH];      // neighbour Laplacians; then w = vMI sM nu4 (lap+ - lap-) / 2
  const int tid = threadIdx.x;
  const int e = A.elems ? A.elems[blockIdx.x] : blockIdx.x;
  const size_t eoffA = (size_t)e * P.naux * NP;
  int2 cn = make_int2(0, 16);
  const bool viscous = P.turbulence == TURB_SMAGORINSKY || P.turb_param != R(0);
  if (A.pf_dist > 0 && blockIdx.x + A.pf_dist < gridDim.x && tid == BLOCK - 1) {
    const int bn = blockIdx.x + A.pf_dist;
    const int en = A.elems ? A.elems[bn] : bn;
    prefetch_l2_bulk(A.Qhd + (size_t)en * 4 * NP, 4 * NP * sizeof(R));
    prefetch_l2_bulk(A.Q + (size_t)en * P.nstate * NP, 5 * NP * sizeof(R));
    prefetch_l2_bulk(A.aux + ((size_t)en * P.naux + P.a_Delta_h) * NP, NP * sizeof(R));
    prefetch_l2_bulk(A.vgeoP + (size_t)en * NP * 10, NP * 10 * sizeof(R));
    prefetch_l2_bulk(A.sgeoP + (size_t)en * NFN * 4, (size_t)NFN * 4 * sizeof(R));
    if (viscous) prefetch_l2_bulk(A.gradflux + (size_t)en * P.ngradflux * NP, (size_t)P.ngradflux * NP * sizeof(R));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(A.conn + (size_t)en * 6));
  }
  for (int it = tid; it < NFN; it += BLOCK)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(A.sgeoP + ((size_t)e * NFN + it) * 4));
  if (tid < NH) {
    cn = A.conn[(size_t)e * 6 + tid / NFP];
    if (((cn.y >> 4) & 15) == 0) {
      const int fn = tid % NFP;
      int a = fn % NQ;
      const int b = fn / NQ;
      if (cn.y & 8) a = NQ - 1 - a;
      const int vp = face_to_vol<NQ>(cn.y & 7, a, b);
      const R *pl = A.Qhd + (size_t)cn.x * 4 * NP + vp;
#pragma unroll
      for (int s = 0; s < 4; ++s) cp_async<sizeof(R)>(&sLp[s][tid], pl + (size_t)s * NP);
    }
  }
  R q[5] = {1, 0, 0, 0, 0}, g[9], MI = 0, nu4 = 0;
  R gf[10], gPhi[3] = {0, 0, 0}, Delta = R(0);
  if (tid < NP) {
    const size_t eo = (size_t)e * 4 * NP + tid;
#pragma unroll
    for (int s = 0; s < 4; ++s) sL[s][tid] = A.Qhd[eo + (size_t)s * NP];
#pragma unroll
    for (int s = 0; s < 5; ++s) q[s] = A.Q[(size_t)e * P.nstate * NP + (size_t)s * NP + tid];
    const R hD = A.aux[eoffA + (size_t)P.a_Delta_h * NP + tid] * R(0.5);
    load_vgeo<R>(A.vgeoP, e, tid, NP, g, MI);
    if (viscous) {
      // issued early: consumed after the face terms
      const size_t eoffG = (size_t)e * P.ngradflux * NP + tid;
#pragma unroll
      for (int s = 0; s < 10; ++s) gf[s] = (s < P.ngf_dyn) ? A.gradflux[eoffG + (size_t)s * NP] : R(0);
      if (AUX && P.turbulence == TURB_SMAGORINSKY) {
#pragma unroll
        for (int d = 0; d < 3; ++d) gPhi[d] = A.aux[eoffA + (size_t)(P.a_gradPhi + d) * NP + tid];
        Delta = A.aux[eoffA + (size_t)P.a_Delta * NP + tid];
      }
    }
    const R hD2 = hD * hD;
    nu4 = hD2 * hD2 / R(2) / P.hyper_tau;
    sNu[tid] = nu4;
  }
  cp_async_wait_all();
  __syncthreads();
  if (tid < NH) {
    const int f = tid / NFP, fn = tid - f * NFP;
    const int vm = face_to_vol<NQ>(f, fn % NQ, fn / NQ);
    const bool wall = ((cn.y >> 4) & 15) != 0;
    const V2 g1 = reinterpret_cast<const V2 *>(A.sgeoP + ((size_t)e * NFN + tid) * 4)[1];
    const R w = g1.y * sNu[vm];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      // walls: lap+ = lap- (no boundary flux of the Laplacian)
      const R dl = wall ? R(0) : (sLp[s][tid] - sL[s][vm]) * R(0.5);
      sLp[s][tid] = w * dl;
    }
  }
  R H[12];
  const int i = tid % NQ, j = (tid / NQ) % NQ, k = tid / (NQ * NQ);
  if (tid < NP) {
    R L1[4] = {0, 0, 0, 0}, L2[4] = {0, 0, 0, 0};
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
      const R d1 = const_D<R>(i * NQ + n), d2 = const_D<R>(j * NQ + n);
      const int o1 = n + NQ * (j + NQ * k), o2 = i + NQ * (n + NQ * k);
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        L1[s] += d1 * sL[s][o1];
        L2[s] += d2 * sL[s][o2];
      }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int d = 0; d < 3; ++d) H[3 * s + d] = nu4 * ((MI * g[d]) * L1[s] + (MI * g[3 + d]) * L2[s]);
  }
  __syncthreads();
  if (tid >= NP) return;
  const int fit[2] = {(i == 0) ? j + NQ * k : ((i == NQ - 1) ? NFP + j + NQ * k : -1),
                      (j == 0) ? 2 * NFP + i + NQ * k : ((j == NQ - 1) ? 3 * NFP + i + NQ * k : -1)};
#pragma unroll
  for (int dir = 0; dir < 2; ++dir) {
    const int it = fit[dir];
    if (it < 0) continue;
    R n[3], sMvMI;
    load_sgeo<R>(A.sgeoP + ((size_t)e * NFN + it) * 4, n, sMvMI);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const R w = sLp[s][it];
#pragma unroll
      for (int d = 0; d < 3; ++d) H[3 * s + d] += n[d] * w;
    }
  }
  // viscous part (skipped when the closure is identically zero, as in the GCM drivers'
  // ConstantKinematicViscosity(0))
  R F2[3][5];
  if (viscous) {
    flux_second_order<R>(P, q, gf, gPhi, Delta, F2);
  } else {
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int c = 0; c < 5; ++c) F2[d][c] = R(0);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int c = 0; c < 3; ++c) F2[d][1 + c] += q[0] * H[d + 3 * c];
    F2[d][4] += (H[d] * q[1] + H[d + 3] * q[2] + H[d + 6] * q[3]) + H[9 + d] * q[0];
  }
  const size_t eoffF = (size_t)e * 12 * NP + tid;
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int c = 1; c < 5; ++c) A.F2[eoffF + (size_t)(4 * d + c - 1) * NP] = F2[d][c];
  write_normal_flux<R, NQ>(A.sgeoP, A.Fn, e, i, j, k, F2);
}

// ---------------------------------------------------------------------------------------
// update! of LowStorageRungeKutta2N (LowStorageRungeKuttaMethod.jl:146-158)
// ---------------------------------------------------------------------------------------
template <class R>
__global__ void lsrk_update_kernel(R *__restrict__ dQ, R *__restrict__ Q, R rka, R rkb, R dt,
                                   size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const R d = dQ[i];
    Q[i] += rkb * dt * d;
    dQ[i] = d * rka;
  }
}

template <class R>
__global__ void scale_kernel(R *__restrict__ x, R a, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= a;
}

// kernel_fillsendbuf! / kernel_transferrecvbuf! (src/Arrays/MPIStateArrays.jl:837-871)
// vmap holds 0-based linear node ids (converted once at bind time).
template <class R>
__global__ void pack_kernel(R *__restrict__ sendbuf, const R *__restrict__ buf,
                            const int64_t *__restrict__ vmap, int64_t nmap, int Np, int nvar) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nmap) return;
  const int64_t id = vmap[i];
  const int64_t e = id / Np, n = id - e * Np;
  for (int s = 0; s < nvar; ++s) sendbuf[i * nvar + s] = buf[(e * nvar + s) * Np + n];
}
template <class R>
__global__ void unpack_kernel(R *__restrict__ buf, const R *__restrict__ recvbuf,
                              const int64_t *__restrict__ vmap, int64_t nmap, int Np, int nvar) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nmap) return;
  const int64_t id = vmap[i];
  const int64_t e = id / Np, n = id - e * Np;
  for (int s = 0; s < nvar; ++s) buf[(e * nvar + s) * Np + n] = recvbuf[i * nvar + s];
}

// Packed private geometry (built once in cmdg_bind_grid).
// non-finite scan of realview(Q) (check_for_crashes, MPIStateArrays.jl:910-935)
template <class R>
__global__ void nonfinite_kernel(const R *__restrict__ x, size_t n, int *__restrict__ flag) {
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    bad |= !isfinite(x[i]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

//   vgeoP[e][c / 2][n][c % 2], c = 3*m + d : M * d(xi_{m+1})/d(x_{d+1});  c = 9 : MI
//   (reference vgeo columns, Grids.jl:76-92: xi{m}x{d} at 3*(d-1)+(m-1), M = 9, MI = 10)
template <class R>
__global__ void pack_vgeo_kernel(R *__restrict__ out, const R *__restrict__ vgeo, int Np,
                                 int nvgeo, size_t nreal) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nreal * Np) return;
  const size_t e = idx / Np;
  const int n = (int)(idx - e * Np);
  const R *v = vgeo + e * (size_t)nvgeo * Np + n;
  // pair-column layout: word c of node n at ((e * 5 + c / 2) * Np + n) * 2 + c % 2
  R *o = out + e * (size_t)Np * 10;
  const R M = v[(size_t)9 * Np];
  for (int m = 0; m < 3; ++m)
    for (int d = 0; d < 3; ++d) {
      const int c = 3 * m + d;
      o[((size_t)(c / 2) * Np + n) * 2 + c % 2] = M * v[(size_t)(3 * d + m) * Np];
    }
  o[((size_t)4 * Np + n) * 2 + 1] = v[(size_t)10 * Np];
}
//   sgeoP[e][f*Nfp + n][c], c = 0..2 unit normal, c = 3 : sM * vMI
//   (reference sgeo[c, n, f, e], Grids.jl:129-146)
template <class R>
__global__ void pack_sgeo_kernel(R *__restrict__ out, const R *__restrict__ sgeo, int Nfp,
                                 size_t nreal) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nreal * 6 * Nfp) return;
  const size_t ef = idx / Nfp;
  const int n = (int)(idx - ef * Nfp);
  const R *s = sgeo + (ef * Nfp + n) * 5;
  R *o = out + (ef * (size_t)Nfp + n) * 4;
  o[0] = s[0];
  o[1] = s[1];
  o[2] = s[2];
  o[3] = s[3] * s[4];
}

}  // namespace cmdg
