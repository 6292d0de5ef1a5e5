// cmdg.cu -- host side of libcmdg: C ABI (include/cmdg.h), handle lifecycle, packed private
// geometry, kernel dispatch, low-storage RK driver and the NCCL halo exchange.
//
// Reference call stack replaced (ClimateMachine.jl):
//   dostep!  (ODESolvers/LowStorageRungeKuttaMethod.jl:102-144)
//     -> (dg::DGModel)(tendency, Q, p, t, alpha, beta)   (DGMethods/DGModel.jl:85-427)
//          begin_ghost_exchange! / end_ghost_exchange!   (Arrays/MPIStateArrays.jl:411-483)
//          launch_volume_gradients!, launch_interface_gradients!,
//          launch_volume_tendency!, launch_interface_tendency!  (DGMethods/SpaceDiscretization.jl)
//     -> update!                                         (LowStorageRungeKuttaMethod.jl:146-158)
#include "../../include/cmdg.h"
#include "cmdg_kernels.cuh"
#include "cmdg_ocean.cuh"
#include "cmdg_tracers.cuh"

#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

using namespace cmdg;

namespace {

std::string g_create_error;

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string &err) {
    if (lib) return true;
    // torch ships its own libnccl.so.2; if it is already mapped, dlopen by soname returns it
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
      err = std::string("cannot load libnccl.so.2: ") + dlerror();
      return false;
    }
#define CMDG_SYM(field, name)                                   \
  field = reinterpret_cast<decltype(field)>(dlsym(lib, name)); \
  if (!field) {                                                 \
    err = std::string("libnccl misses symbol ") + name;         \
    return false;                                               \
  }
    CMDG_SYM(GetUniqueId, "ncclGetUniqueId");
    CMDG_SYM(CommInitRank, "ncclCommInitRank");
    CMDG_SYM(CommDestroy, "ncclCommDestroy");
    CMDG_SYM(GroupStart, "ncclGroupStart");
    CMDG_SYM(GroupEnd, "ncclGroupEnd");
    CMDG_SYM(Send, "ncclSend");
    CMDG_SYM(Recv, "ncclRecv");
    CMDG_SYM(AllReduce, "ncclAllReduce");
    CMDG_SYM(GetErrorString, "ncclGetErrorString");
#undef CMDG_SYM
    return true;
  }
};
NcclApi g_nccl;

// cuStreamWaitValue32 through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWaitValue32Fn stream_wait_value32() {
  static StreamWaitValue32Fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<StreamWaitValue32Fn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

}  // namespace

struct cmdg_handle_s {
  cmdg_desc d{};
  std::string err;
  int Nq = 0, Np = 0, Nfp = 0;
  int device = 0; // CUDA device current at cmdg_create (the handle's GPU)
  size_t fb = 8;  // bytes per float
  bool aux_model = false, visc = false;
  int tail_pf = 1036;     // tail prefetch for the next launch: one wave (148 SMs x 5 blocks) + pf_dist elements (CMDG_TAILPF)
  int pf_dist = 296;      // L2 prefetch distance of the one-shot tendency kernel: 2 blocks per SM ahead (CMDG_PF overrides)
  // caller-owned device arrays
  void *aux = nullptr, *gradflux = nullptr;
  // private device buffers
  void *vgeoP = nullptr, *sgeoP = nullptr, *Ddev = nullptr;
  // plus-side geopotential / reference pressure per face node (constants of the grid; refreshed from the bound
  // aux array by the first launch after cmdg_bind_state)
  void *fauxP = nullptr;
  bool faux_dirty = true;
  std::vector<double> Dhost;  // row-major D, mirrored into the constant-memory copy before launches
  int2 *conn = nullptr;
  int *interior = nullptr, *exterior = nullptr;
  int64_t ninterior = 0, nexterior = 0;
  int64_t *vmapsend0 = nullptr, *vmaprecv0 = nullptr;
  int64_t nvmapsend = 0, nvmaprecv = 0;
  std::vector<int> nabrtorank;
  std::vector<int64_t> sendrange, recvrange;  // 0-based [first, last) pairs
  void *sendbuf = nullptr, *recvbuf = nullptr;
  size_t commbuf_states = 0;
  // per-step filter (cmdg_set_step_filter): row-major device copies of the two matrices
  void *courant_dev = nullptr;
  int *crash_dev = nullptr;   // [2]: local flag, reduced flag (cmdg_check_for_crashes)
  void *stepWh = nullptr, *stepWv = nullptr, *tmpWh = nullptr, *tmpWv = nullptr;
  int step_filter_target = -1, step_filter_dir = 0;
  unsigned step_filter_mask = 0;
  // second-order path: diffusive flux per node (ghost part filled by the halo exchange) and its
  // normal component at every real element's own face nodes, written by the gradient kernel
  void *F2dev = nullptr, *FnDev = nullptr;
  // DryBiharmonic: states_higher_order (create_states.jl:20-27), private: grad of (u_h, h_tot)
  // (12 columns) and their horizontal Laplacians (4 columns)
  bool hyper = false;
  void *Qhg = nullptr, *Qhd = nullptr;
  // passive tracers: nu per node from the gradient kernel, tracer diffusive flux per node
  int ntracers = 0;
  void *NuDev = nullptr, *F2chi = nullptr;
  void *Qtmp = nullptr;              // ping-pong partner of Q in cmdg_lsrk_steps
  void *Qdev = nullptr, *dQdev = nullptr;  // device state of cmdg_lsrk_steps_host
  // Pipelined host path (cmdg_lsrk_steps_host; single rank, Euler path).  The upload is cut into element-range
  // chunks on a copy stream and the first stage runs chunk by chunk over the elements whose stencil (the
  // element and its face neighbours) has landed; the last stage runs range by range and every range goes back
  // to the host while the next one is computed.  Only the first upload chunk's latency, the stragglers of the
  // first stage and the last download chunk stay exposed.
  struct HostPipe {
    int nch = 0;
    std::vector<int64_t> first;       // upload / download chunk c = elements [first[c], first[c+1])
    std::vector<int64_t> ready_off;   // ready_list[ready_off[c] .. ready_off[c+1]) can run once chunk c has landed
    int *ready_list = nullptr;        // device
    int *identity = nullptr;          // device: 0 .. nreal-1 (range launches of the last stage)
    cudaStream_t copy = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_k;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  } hp;
  std::vector<int> nbr_host;         // [nreal][6] neighbour element of every face (-1: boundary face)
  bool grid_bound = false;
  // ocean (HBModel)
  bool is_hb = false, ocean_set = false;
  cmdg_ocean_desc od{};
  void *Fc = nullptr, *Fe = nullptr, *Imat = nullptr, *JcV = nullptr;  // row-major operators, vgeo col 16
  const void *vgeo_bound = nullptr;
  // communication
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  // fused stepper on > 1 rank: the exterior chain (exterior kernel -> pack -> NCCL -> unpack) runs on
  // this high-priority side stream concurrently with the interior kernel (CMDG_OVERLAP=0 turns it off).
  // Measured at 2 GPUs, LSRK54 steps/s: 331 (serial, normal-priority NCCL stream) -> 339 (high-priority
  // NCCL stream) -> 344 (+ this chain).
  cudaStream_t ext_stream = nullptr;
  cudaEvent_t ev_ext = nullptr, ev_int = nullptr, ev_extk = nullptr, ev_gextk = nullptr, ev_gint = nullptr;
  cudaEvent_t ev_hb_fext = nullptr, ev_hb_fint = nullptr;   // ocean two-chain schedule: filters of the two chains
  int *int_stacks = nullptr, *ext_stacks = nullptr;          // first element of every interior / exterior stack (HBModel)
  bool overlap_exterior = true;
  // single-launch schedule (CMDG_OVERLAP=2): launch list [exterior..., interior...], counter of finished exterior
  // blocks (cumulative over stages), its running target on the host
  int *all_list = nullptr;
  unsigned *ext_done = nullptr;
  unsigned ext_target = 0;
  int overlap_mode = 2;
  bool exchange_open = false;
  // bookkeeping
  int64_t launches = 0;
  bool timing = false;
  std::vector<cudaEvent_t> tev;
  size_t tev_used = 0;
  double last_ms = -1;
  int64_t last_nl = 0;
  // kernel class of every timed launch (CMDG_KCLASS_*) and the per-class totals of the last call
  std::vector<int> tev_class;
  double class_ms[CMDG_KCLASS_COUNT] = {0};
  int64_t class_n[CMDG_KCLASS_COUNT] = {0};
  // per-stage timeline of the last two steps of a cmdg_lsrk_steps call (diagnostic; CMDG_TIMELINE=<path>
  // with timing enabled): labelled marks recorded on the stream that does the work, dumped as CSV
  struct TlMark { cudaEvent_t ev; int label; long long info; int step, stage; };
  std::vector<TlMark> tl;
  size_t tl_used = 0;
  std::string tl_path;
  bool tl_on = false;
  int tl_step = 0, tl_stage = 0;
};
enum { TL_STEP_BEGIN = 0, TL_KERNEL_BEGIN, TL_KERNEL_END, TL_PACK_END, TL_NCCL_BEGIN, TL_NCCL_END,
       TL_UNPACK_BEGIN, TL_UNPACK_END, TL_COUNT };
static const char *const TL_NAMES[TL_COUNT] = {"step_begin", "kernel_begin", "kernel_end", "pack_end", "nccl_begin",
                                               "nccl_end", "unpack_begin", "unpack_end"};

namespace {

int fail(cmdg_handle h, int code, const std::string &msg) {
  if (h) h->err = msg;
  else g_create_error = msg;
  return code;
}
#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(h, CMDG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
#define NC(call)                                                                            \
  do {                                                                                      \
    ncclResult_t r_ = (call);                                                               \
    if (r_ != ncclSuccess)                                                                  \
      return fail(h, CMDG_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); \
  } while (0)

template <class R>
AtmosParams<R> make_params(const cmdg_handle_s *h) {
  const cmdg_desc &d = h->d;
  AtmosParams<R> P{};
  P.R_d = (R)d.R_d;
  P.cp_d = (R)d.cp_d;
  P.cv_d = (R)d.cv_d;
  P.T_0 = (R)d.T_0;
  P.MSLP = (R)d.MSLP;
  P.grav = (R)d.grav;
  P.two_Omega = (R)(2 * (R)d.Omega);
  P.inv_Pr_turb = (R)d.inv_Pr_turb;
  P.gamma = P.cp_d / P.cv_d;
  P.inv_cv = R(1) / P.cv_d;
  P.kappa = P.R_d / P.cp_d;
  P.turb_param = (R)d.turb_param;
  P.turbulence = d.turbulence;
  P.with_divergence = d.turb_with_divergence;
  P.sources = d.sources;
  P.subtract_off = (d.ref_state == CMDG_REF_HYDROSTATIC) && d.subtract_off;
  P.horizontal_diffusion = d.diffusion_direction == CMDG_DIR_HORIZONTAL;
  for (int i = 0; i < 6; ++i) P.bc_kind[i] = d.bc_kind[i];
  // auxiliary layout, vars_state(::AtmosModel, ::Auxiliary) (AtmosModel.jl:479-497)
  int c = 3;
  P.a_Phi = P.a_gradPhi = P.a_ref_rho = P.a_ref_p = P.a_Delta = P.a_Delta_h = -1;
  P.hyper_tau = (R)d.hyper_tau;
  if (d.orientation != CMDG_ORIENT_NONE) {
    P.a_Phi = c;
    P.a_gradPhi = c + 1;
    c += 4;
  }
  if (d.ref_state == CMDG_REF_HYDROSTATIC) {
    P.a_ref_rho = c;
    P.a_ref_p = c + 1;
    c += 7;
  }
  if (d.turbulence == CMDG_TURB_SMAGORINSKY) P.a_Delta = c++;
  if (d.hyperdiffusion == CMDG_HYPER_DRY_BIHARMONIC) P.a_Delta_h = c++;
  P.a_theta_v = c;
  P.a_T = c + 1;
  P.naux = c + 2 + d.ntracers;
  P.nstate = d.nstate;
  P.ngf_dyn = d.turbulence == CMDG_TURB_SMAGORINSKY ? 10 : 9;
  P.ngradflux = d.ngradflux;
  P.inv_day = d.day > 0 ? (R)(R(1) / (R)d.day) : R(0);
  P.sponge_z_max = (R)d.sponge_z_max;
  P.sponge_z_sponge = (R)d.sponge_z_sponge;
  P.sponge_alpha_max = (R)d.sponge_alpha_max;
  P.sponge_gamma = (R)d.sponge_gamma;
  for (int i = 0; i < 3; ++i) P.sponge_u[i] = (R)d.sponge_u_relax[i];
  return P;
}

int expected_naux(const cmdg_desc &d) {
  int c = 3;
  if (d.orientation != CMDG_ORIENT_NONE) c += 4;
  if (d.ref_state == CMDG_REF_HYDROSTATIC) c += 7;
  if (d.turbulence == CMDG_TURB_SMAGORINSKY) c += 1;
  if (d.hyperdiffusion == CMDG_HYPER_DRY_BIHARMONIC) c += 1;
  return c + 2 + d.ntracers;
}

cudaEvent_t timing_event(cmdg_handle h, int kclass = CMDG_KCLASS_TENDENCY) {
  if (h->tev_used == h->tev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->tev.push_back(e);
    h->tev_class.push_back(0);
  }
  h->tev_class[h->tev_used] = kclass;
  return h->tev[h->tev_used++];
}

void tl_mark(cmdg_handle h, cudaStream_t st, int label, long long info = 0) {
  if (!h->tl_on) return;
  if (h->tl_used == h->tl.size()) {
    cmdg_handle_s::TlMark m{};
    cudaEventCreate(&m.ev);
    h->tl.push_back(m);
  }
  cmdg_handle_s::TlMark &m = h->tl[h->tl_used++];
  m.label = label;
  m.info = info;
  m.step = h->tl_step;
  m.stage = h->tl_stage;
  cudaEventRecord(m.ev, st);
}

void tl_dump(cmdg_handle h) {
  if (h->tl_path.empty() || h->tl_used == 0) return;
  const std::string path = h->tl_path + ".rank" + std::to_string(h->rank) + ".csv";
  FILE *f = fopen(path.c_str(), "w");
  if (!f) return;
  fprintf(f, "step,stage,label,info,t_us\n");
  cudaEvent_t t0 = h->tl[0].ev;
  for (size_t i = 0; i < h->tl_used; ++i) {
    const cmdg_handle_s::TlMark &m = h->tl[i];
    if (m.label == TL_STEP_BEGIN) t0 = m.ev;
    float x = 0;
    cudaEventElapsedTime(&x, t0, m.ev);
    fprintf(f, "%d,%d,%s,%lld,%.3f\n", m.step, m.stage, TL_NAMES[m.label], m.info, 1e3 * (double)x);
  }
  fclose(f);
}

// The constant-memory copy of D is per process: re-upload when another handle used it last.
const cmdg_handle_s *g_constD_owner = nullptr;
int g_constD_device = -1;
template <class R>
int ensure_const_D(cmdg_handle h, cudaStream_t st) {
  if (g_constD_owner == h && g_constD_device == h->device) return 0;
  if (sizeof(R) == 8) {
    CU(cudaMemcpyToSymbolAsync(c_D64, h->Dhost.data(), 64 * sizeof(double), 0, cudaMemcpyHostToDevice, st));
  } else {
    static float tmp[64];
    for (int i = 0; i < 64; ++i) tmp[i] = (float)h->Dhost[i];
    CU(cudaMemcpyToSymbolAsync(c_D32, tmp, 64 * sizeof(float), 0, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
  }
  g_constD_owner = h;
  g_constD_device = h->device;
  return 0;
}

template <class R, int NQ, int NF1, bool AUX, bool VISC, bool SRCX = false>
int launch_tend_inst(cmdg_handle h, const TendArgs<R> &a, const AtmosParams<R> &P, int64_t n,
                     cudaStream_t st) {
  using SM = TendSmem<R, NQ, AUX, VISC>;
  if (int rc = ensure_const_D<R>(h, st)) return rc;
  auto kern = dg_tendency_kernel<R, NQ, NF1, AUX, VISC, SRCX>;
  // the attribute is per device (a process may hold handles on several GPUs): once per device and instantiation
  static unsigned long long attr_devices = 0;   // bit d: attribute set on device d for this instantiation
  if (h->device >= 64 || !((attr_devices >> h->device) & 1ull)) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM)));
    if (h->device < 64) attr_devices |= 1ull << h->device;
  }
  if (h->timing) cudaEventRecord(timing_event(h), st);
  tl_mark(h, st, TL_KERNEL_BEGIN, (long long)n);
  kern<<<(unsigned)n, Dims<NQ>::BLOCK, sizeof(SM), st>>>(a, P);
  tl_mark(h, st, TL_KERNEL_END, (long long)n);
  if (h->timing) cudaEventRecord(timing_event(h), st);
  CU(cudaGetLastError());
  h->launches++;
  return 0;
}

template <class R, int NQ, int NF1>
int launch_tend_nf(cmdg_handle h, const TendArgs<R> &a, const AtmosParams<R> &P, int64_t n,
                   cudaStream_t st) {
  if (h->aux_model) {
    if (P.sources & (SRC_HELD_SUAREZ | SRC_RAYLEIGH_SPONGE)) {
      if (h->visc) return launch_tend_inst<R, NQ, NF1, true, true, true>(h, a, P, n, st);
      return launch_tend_inst<R, NQ, NF1, true, false, true>(h, a, P, n, st);
    }
    if (h->visc) return launch_tend_inst<R, NQ, NF1, true, true>(h, a, P, n, st);
    return launch_tend_inst<R, NQ, NF1, true, false>(h, a, P, n, st);
  }
  if (h->visc) return launch_tend_inst<R, NQ, NF1, false, true>(h, a, P, n, st);
  return launch_tend_inst<R, NQ, NF1, false, false>(h, a, P, n, st);
}

template <class R>
int launch_tendency(cmdg_handle h, const TendArgs<R> &a, int64_t n, cudaStream_t st) {
  if (n <= 0) return 0;
  const AtmosParams<R> P = make_params<R>(h);
  switch (h->d.nf_first) {
    case CMDG_NF_RUSANOV: return launch_tend_nf<R, 5, NF_RUSANOV>(h, a, P, n, st);
    case CMDG_NF_CENTRAL: return launch_tend_nf<R, 5, NF_CENTRAL>(h, a, P, n, st);
    case CMDG_NF_ROE: return launch_tend_nf<R, 5, NF_ROE>(h, a, P, n, st);
  }
  return fail(h, CMDG_ERR_UNSUPPORTED, "unsupported first-order numerical flux");
}

template <class R, bool AUX, bool HYPER>
int launch_gradient_inst(cmdg_handle h, const GradArgs<R> &a, const AtmosParams<R> &P, int64_t n,
                         cudaStream_t st) {
  using SM = GradSmem<R, 5, AUX, HYPER>;
  auto kern = dg_gradient_kernel<R, 5, AUX, HYPER>;
  static unsigned long long attr_devices = 0;
  if (h->device >= 64 || !((attr_devices >> h->device) & 1ull)) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM)));
    if (h->device < 64) attr_devices |= 1ull << h->device;
  }
  if (h->timing) cudaEventRecord(timing_event(h, CMDG_KCLASS_GRADIENT), st);
  kern<<<(unsigned)n, Dims<5>::BLOCK, sizeof(SM), st>>>(a, P);
  if (h->timing) cudaEventRecord(timing_event(h, CMDG_KCLASS_GRADIENT), st);
  CU(cudaGetLastError());
  h->launches++;
  return 0;
}

template <class R>
int launch_gradient(cmdg_handle h, const GradArgs<R> &a, int64_t n, cudaStream_t st) {
  if (n <= 0) return 0;
  if (int rc = ensure_const_D<R>(h, st)) return rc;
  const AtmosParams<R> P = make_params<R>(h);
  if (h->hyper) return launch_gradient_inst<R, true, true>(h, a, P, n, st);
  if (h->aux_model) return launch_gradient_inst<R, true, false>(h, a, P, n, st);
  return launch_gradient_inst<R, false, false>(h, a, P, n, st);
}

template <class R> int exchange_begin_t(cmdg_handle h, void *array, int nstate, cudaStream_t st);
template <class R> int exchange_end_t(cmdg_handle h, void *array, int nstate, cudaStream_t st);

// DryBiharmonic passes 2 and 3 (hyper_divergence_kernel, hyper_flux_kernel)
template <class R>
HyperArgs<R> hyper_args(cmdg_handle h, const void *Q) {
  HyperArgs<R> a{};
  a.Q = (const R *)Q;
  a.aux = (const R *)h->aux;
  a.gradflux = (const R *)h->gradflux;
  a.vgeoP = (const R *)h->vgeoP;
  a.sgeoP = (const R *)h->sgeoP;
  a.conn = h->conn;
  a.Qhg = (R *)h->Qhg;
  a.Qhd = (R *)h->Qhd;
  a.F2 = (R *)h->F2dev;
  a.Fn = (R *)h->FnDev;
  a.pf_dist = h->pf_dist;
  return a;
}
template <class R>
int launch_hyper(cmdg_handle h, int pass, const HyperArgs<R> &a, int64_t n, cudaStream_t st) {
  if (n <= 0) return 0;
  if (int rc = ensure_const_D<R>(h, st)) return rc;
  const int kc = pass == 2 ? CMDG_KCLASS_HYPER_DIVERGENCE : CMDG_KCLASS_HYPER_FLUX;
  if (h->timing) cudaEventRecord(timing_event(h, kc), st);
  if (pass == 2) {
    hyper_divergence_kernel<R, 5><<<(unsigned)n, Dims<5>::BLOCK, 0, st>>>(a);
  } else {
    const AtmosParams<R> P = make_params<R>(h);
    hyper_flux_kernel<R, 5, true><<<(unsigned)n, Dims<5>::BLOCK, 0, st>>>(a, P);
  }
  if (h->timing) cudaEventRecord(timing_event(h, kc), st);
  CU(cudaGetLastError());
  h->launches++;
  return 0;
}

// Everything between the state exchange and the tendency kernel on the second-order path: gradient
// kernel (+ the two hyperdiffusion passes with their exchanges), leaving F2 / Fn ready and the F2
// halo in flight (`first`/`second` = the launch lists processed before / after each exchange starts;
// the reference order is interior-first, the fused stepper's exterior-first).
template <class R>
int second_order_passes(cmdg_handle h, GradArgs<R> ga, const void *Q, bool par, bool exterior_first,
                        bool q_exchange_open, cudaStream_t st) {
  int rc;
  const int64_t nreal = h->d.nrealelem;
  if (!par) {
    ga.elems = nullptr;
    if ((rc = launch_gradient<R>(h, ga, nreal, st))) return rc;
    if (h->hyper) {
      HyperArgs<R> ha = hyper_args<R>(h, Q);
      if ((rc = launch_hyper<R>(h, 2, ha, nreal, st))) return rc;
      if ((rc = launch_hyper<R>(h, 3, ha, nreal, st))) return rc;
    }
    return 0;
  }
  const int *ext = h->exterior, *inr = h->interior;
  const int64_t next = h->nexterior, ninr = h->ninterior;
  void *Qv = const_cast<void *>(Q);
  if (exterior_first) {
    // ghosts of Q are already in place
    ga.elems = ext;
    if ((rc = launch_gradient<R>(h, ga, next, st))) return rc;
    if (!h->hyper) {
      if ((rc = exchange_begin_t<R>(h, h->F2dev, 12, st))) return rc;
      ga.elems = inr;
      if ((rc = launch_gradient<R>(h, ga, ninr, st))) return rc;
      return exchange_end_t<R>(h, h->F2dev, 12, st);
    }
    if ((rc = exchange_begin_t<R>(h, h->Qhg, 12, st))) return rc;
    ga.elems = inr;
    if ((rc = launch_gradient<R>(h, ga, ninr, st))) return rc;
    if ((rc = exchange_end_t<R>(h, h->Qhg, 12, st))) return rc;
    HyperArgs<R> ha = hyper_args<R>(h, Q);
    ha.elems = ext;
    if ((rc = launch_hyper<R>(h, 2, ha, next, st))) return rc;
    if ((rc = exchange_begin_t<R>(h, h->Qhd, 4, st))) return rc;
    ha.elems = inr;
    if ((rc = launch_hyper<R>(h, 2, ha, ninr, st))) return rc;
    if ((rc = exchange_end_t<R>(h, h->Qhd, 4, st))) return rc;
    ha.elems = ext;
    if ((rc = launch_hyper<R>(h, 3, ha, next, st))) return rc;
    if ((rc = exchange_begin_t<R>(h, h->F2dev, 12, st))) return rc;
    ha.elems = inr;
    if ((rc = launch_hyper<R>(h, 3, ha, ninr, st))) return rc;
    return exchange_end_t<R>(h, h->F2dev, 12, st);
  }
  // reference order (DGModel.jl:125-310): interior while the Q halo is in flight, then exterior
  ga.elems = inr;
  if ((rc = launch_gradient<R>(h, ga, ninr, st))) return rc;
  if (q_exchange_open && (rc = exchange_end_t<R>(h, Qv, h->d.nstate, st))) return rc;
  ga.elems = ext;
  if ((rc = launch_gradient<R>(h, ga, next, st))) return rc;
  if (h->hyper) {
    HyperArgs<R> ha = hyper_args<R>(h, Q);
    if ((rc = exchange_begin_t<R>(h, h->Qhg, 12, st))) return rc;
    ha.elems = inr;
    if ((rc = launch_hyper<R>(h, 2, ha, ninr, st))) return rc;
    if ((rc = exchange_end_t<R>(h, h->Qhg, 12, st))) return rc;
    ha.elems = ext;
    if ((rc = launch_hyper<R>(h, 2, ha, next, st))) return rc;
    if ((rc = exchange_begin_t<R>(h, h->Qhd, 4, st))) return rc;
    ha.elems = inr;
    if ((rc = launch_hyper<R>(h, 3, ha, ninr, st))) return rc;
    if ((rc = exchange_end_t<R>(h, h->Qhd, 4, st))) return rc;
    ha.elems = ext;
    if ((rc = launch_hyper<R>(h, 3, ha, next, st))) return rc;
  }
  return exchange_begin_t<R>(h, h->F2dev, 12, st);   // ended by the caller after the interior tendency
}

// ------------------------------------------------------------------------------------
// halo exchange (MPIStateArrays.jl:411-514)
// ------------------------------------------------------------------------------------
template <class R>
int exchange_begin_t(cmdg_handle h, void *array, int nstate, cudaStream_t st) {
  if (!h->comm || h->nabrtorank.empty()) return 0;
  if (h->exchange_open) return fail(h, CMDG_ERR_INVALID,
                                    "The current ghost exchange must end before another begins.");
  if ((size_t)nstate > h->commbuf_states) {
    if (h->sendbuf) cudaFree(h->sendbuf);
    if (h->recvbuf) cudaFree(h->recvbuf);
    CU(cudaMalloc(&h->sendbuf, (size_t)h->nvmapsend * nstate * sizeof(R) + 16));
    CU(cudaMalloc(&h->recvbuf, (size_t)h->nvmaprecv * nstate * sizeof(R) + 16));
    h->commbuf_states = nstate;
  }
  // kernel_fillsendbuf! on the compute stream, then hand over to the comm stream
  if (h->nvmapsend > 0) {
    pack_kernel<R><<<(unsigned)((h->nvmapsend + 255) / 256), 256, 0, st>>>(
        (R *)h->sendbuf, (const R *)array, h->vmapsend0, h->nvmapsend, h->Np, nstate);
    CU(cudaGetLastError());
    h->launches++;
  }
  tl_mark(h, st, TL_PACK_END, (long long)h->nvmapsend * nstate);
  CU(cudaEventRecord(h->ev_ready, st));
  CU(cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0));
  tl_mark(h, h->comm_stream, TL_NCCL_BEGIN);
  NC(g_nccl.GroupStart());
  for (size_t n = 0; n < h->nabrtorank.size(); ++n) {
    const int64_t r0 = h->recvrange[2 * n], r1 = h->recvrange[2 * n + 1];
    const int64_t s0 = h->sendrange[2 * n], s1 = h->sendrange[2 * n + 1];
    NC(g_nccl.Recv((char *)h->recvbuf + (size_t)r0 * nstate * sizeof(R),
                   (size_t)(r1 - r0) * nstate * sizeof(R), ncclInt8, h->nabrtorank[n], h->comm,
                   h->comm_stream));
    NC(g_nccl.Send((const char *)h->sendbuf + (size_t)s0 * nstate * sizeof(R),
                   (size_t)(s1 - s0) * nstate * sizeof(R), ncclInt8, h->nabrtorank[n], h->comm,
                   h->comm_stream));
  }
  NC(g_nccl.GroupEnd());
  tl_mark(h, h->comm_stream, TL_NCCL_END);
  CU(cudaEventRecord(h->ev_done, h->comm_stream));
  h->exchange_open = true;
  return 0;
}

template <class R>
int exchange_end_t(cmdg_handle h, void *array, int nstate, cudaStream_t st) {
  if (!h->comm || h->nabrtorank.empty()) return 0;
  if (!h->exchange_open)
    return fail(h, CMDG_ERR_INVALID, "A ghost exchange must begin before it ends.");
  CU(cudaStreamWaitEvent(st, h->ev_done, 0));
  tl_mark(h, st, TL_UNPACK_BEGIN);
  if (h->nvmaprecv > 0) {
    unpack_kernel<R><<<(unsigned)((h->nvmaprecv + 255) / 256), 256, 0, st>>>(
        (R *)array, (const R *)h->recvbuf, h->vmaprecv0, h->nvmaprecv, h->Np, nstate);
    CU(cudaGetLastError());
    h->launches++;
  }
  tl_mark(h, st, TL_UNPACK_END);
  h->exchange_open = false;
  return 0;
}


// ------------------------------------------------------------------------------------
// HBModel: one evaluation in the reference's order (DGModel.jl:85-427 with the model hooks of
// hydrostatic_boussinesq_model.jl:637-712).  Qout != nullptr fuses the RK stage update.
// ------------------------------------------------------------------------------------
template <class R>
HBParams<R> make_hb_params(const cmdg_handle_s *h) {
  const cmdg_ocean_desc &o = h->od;
  HBParams<R> P{};
  P.grav = (R)o.grav; P.rho0 = (R)o.rho0; P.ch = (R)o.ch; P.cz = (R)o.cz; P.alphaT = (R)o.alphaT;
  P.nuh = (R)o.nuh; P.nuz = (R)o.nuz; P.kappah = (R)o.kappah; P.kappaz = (R)o.kappaz;
  P.kappac = (R)o.kappac; P.f0 = (R)o.f0; P.beta = (R)o.beta;
  P.Ly = (R)o.Ly; P.tau0 = (R)o.tau0; P.lambda_r = (R)o.lambda_r; P.thetaE = (R)o.thetaE;
  for (int i = 0; i < 6; ++i) { P.bc_vel[i] = o.bc_velocity[i]; P.bc_temp[i] = o.bc_temperature[i]; }
  P.nvertelem = h->d.nvertelem;
  return P;
}

template <class R>
int hb_launch_tend(cmdg_handle h, const HBArgs<R> &a, const HBParams<R> &P, int64_t n, cudaStream_t st) {
  if (n <= 0) return 0;
  if (h->timing) cudaEventRecord(timing_event(h), st);
  tl_mark(h, st, TL_KERNEL_BEGIN, 400000000LL + (long long)n);   // timeline info: kernel family * 1e8 + elements
  if (h->d.nf_first == CMDG_NF_RUSANOV)
    hb_tendency_kernel<R, 5, NF_RUSANOV><<<(unsigned)n, Dims<5>::BLOCK, 0, st>>>(a, P);
  else
    hb_tendency_kernel<R, 5, NF_CENTRAL><<<(unsigned)n, Dims<5>::BLOCK, 0, st>>>(a, P);
  tl_mark(h, st, TL_KERNEL_END, 400000000LL + (long long)n);
  if (h->timing) cudaEventRecord(timing_event(h), st);
  CU(cudaGetLastError());
  h->launches++;
  return 0;
}

template <class R>
int hb_eval_t(cmdg_handle h, void *dQ, void *Q, void *Qout, double alpha, double beta,
              double rkb_dt, cudaStream_t st, bool two_chains = false) {
  const bool par = h->comm && !h->nabrtorank.empty();
  const int64_t nreal = h->d.nrealelem, nghost = h->d.nelem - nreal;
  const int nv = h->d.nvertelem;
  const HBParams<R> P = make_hb_params<R>(h);
  HBArgs<R> a{};
  a.Q = (R *)Q;
  a.aux = (R *)h->aux;
  a.gradflux = (R *)h->gradflux;
  a.dQ = (R *)dQ;
  a.Qout = (R *)Qout;
  a.vgeoP = (const R *)h->vgeoP;
  a.sgeoP = (const R *)h->sgeoP;
  a.conn = h->conn;
  a.D = (const R *)h->Ddev;
  a.alpha = (R)alpha;
  a.beta = (R)beta;
  a.rkb_dt = (R)rkb_dt;
  int rc;
  // timeline info of the marks: kernel family * 1e8 + elements
  auto grad = [&](const int *elems, int64_t n, cudaStream_t s) -> int {
    if (n <= 0) return 0;
    HBArgs<R> g = a;
    g.elems = elems;
    if (h->timing) cudaEventRecord(timing_event(h, CMDG_KCLASS_GRADIENT), s);
    tl_mark(h, s, TL_KERNEL_BEGIN, 200000000LL + (long long)n);
    hb_gradient_kernel<R, 5><<<(unsigned)n, Dims<5>::BLOCK, 0, s>>>(g, P);
    tl_mark(h, s, TL_KERNEL_END, 200000000LL + (long long)n);
    if (h->timing) cudaEventRecord(timing_event(h, CMDG_KCLASS_GRADIENT), s);
    CU(cudaGetLastError());
    h->launches++;
    return 0;
  };
  // stacks [elem0, elem0 + nelems) or, with `stacks`, the listed ones (nelems = their element count)
  auto column = [&](int64_t elem0, int64_t nelems, int set_wz0, const int *stacks, cudaStream_t s) -> int {
    if (nelems <= 0) return 0;
    // segmented scan (one block per stack, 32 element slots x 25 horizontal nodes) unless the stack is too
    // tall for the shared-memory carries or CMDG_HB_SERIAL_COLUMN asks for the reference-like serial march
    if (h->timing) cudaEventRecord(timing_event(h, CMDG_KCLASS_HB_COLUMN), s);
    tl_mark(h, s, TL_KERNEL_BEGIN, 300000000LL + (long long)nelems);
    constexpr int SLOTS = 32;
    const size_t carry_bytes = (size_t)2 * nv * 25 * sizeof(R);
    static const bool serial = getenv("CMDG_HB_SERIAL_COLUMN") != nullptr;
    if (!serial && carry_bytes <= 40 * 1024) {
      hb_column_scan_kernel<R, 5, SLOTS><<<(unsigned)(nelems / nv), SLOTS * 25, carry_bytes, s>>>(
          (R *)h->aux, (const R *)Q, (const R *)h->gradflux, (const R *)h->JcV, (const R *)h->Imat,
          P.alphaT, nv, (int)elem0, set_wz0, stacks);
    } else {
      hb_column_kernel<R, 5><<<(unsigned)(nelems / nv), 32, 0, s>>>(
          (R *)h->aux, (const R *)Q, (const R *)h->gradflux, (const R *)h->JcV, (const R *)h->Imat,
          P.alphaT, nv, (int)elem0, set_wz0, stacks);
    }
    tl_mark(h, s, TL_KERNEL_END, 300000000LL + (long long)nelems);
    if (h->timing) cudaEventRecord(timing_event(h, CMDG_KCLASS_HB_COLUMN), s);
    CU(cudaGetLastError());
    h->launches++;
    return 0;
  };
  // update_auxiliary_state!: vertical filters, in place
  auto filter = [&](const int *elems, int64_t n, cudaStream_t s) -> int {
    if (n <= 0) return 0;
    if (h->timing) cudaEventRecord(timing_event(h, CMDG_KCLASS_HB_FILTER), s);
    tl_mark(h, s, TL_KERNEL_BEGIN, 100000000LL + (long long)n);
    hb_filter_kernel<R, 5><<<(unsigned)n, Dims<5>::BLOCK, 0, s>>>((R *)Q, (const R *)h->Fc, (const R *)h->Fe,
                                                                  (int)nreal, elems);
    tl_mark(h, s, TL_KERNEL_END, 100000000LL + (long long)n);
    if (h->timing) cudaEventRecord(timing_event(h, CMDG_KCLASS_HB_FILTER), s);
    CU(cudaGetLastError());
    h->launches++;
    return 0;
  };
  auto tend = [&](const int *elems, int64_t n, cudaStream_t s) -> int {
    HBArgs<R> t = a;
    t.elems = elems;
    return hb_launch_tend<R>(h, t, P, n, s);
  };
  if (par && two_chains) {
    // Two chains coupled kernel to kernel (fused stepper, > 1 rank).  Whole stacks are interior or exterior, the
    // vertical filters and the stack integrals never leave a stack, so every kernel family splits cleanly:
    //   side stream (high priority): filter_ext -> [Q halo] -> grad_ext -> [GF halo || column_ext] -> column_ghost -> tend_ext
    //   main stream:                 filter_int ------------> grad_int -> column_int ----------------------------> tend_int
    // grad_x needs both filters, tend_x needs both gradients + column integrals; the main stream never waits for
    // NCCL itself, only for exterior KERNELS.  Stage s + 1 is ordered behind stage s through the filters: filter_x(s+1)
    // follows tend_x(s) in stream order, and every kernel that overwrites data the other chain's tend(s) reads
    // (GF, aux.w / pkin / wz0, the ping-pong state) first waits for the other chain's filter(s+1).
    cudaStream_t xs = h->ext_stream;
    if ((rc = filter(h->exterior, h->nexterior, xs))) return rc;
    CU(cudaEventRecord(h->ev_hb_fext, xs));
    if ((rc = filter(h->interior, h->ninterior, st))) return rc;
    CU(cudaEventRecord(h->ev_hb_fint, st));
    if ((rc = exchange_begin_t<R>(h, Q, HB_S, xs))) return rc;
    if ((rc = exchange_end_t<R>(h, Q, HB_S, xs))) return rc;
    CU(cudaStreamWaitEvent(xs, h->ev_hb_fint, 0));
    if ((rc = grad(h->exterior, h->nexterior, xs))) return rc;
    if ((rc = exchange_begin_t<R>(h, h->gradflux, HB_GF, xs))) return rc;
    if ((rc = column(0, h->nexterior, 1, h->ext_stacks, xs))) return rc;
    CU(cudaEventRecord(h->ev_gextk, xs));
    CU(cudaStreamWaitEvent(st, h->ev_hb_fext, 0));
    if ((rc = grad(h->interior, h->ninterior, st))) return rc;
    if ((rc = column(0, h->ninterior, 1, h->int_stacks, st))) return rc;
    CU(cudaEventRecord(h->ev_gint, st));
    if ((rc = exchange_end_t<R>(h, h->gradflux, HB_GF, xs))) return rc;
    if ((rc = column(nreal, nghost, 0, nullptr, xs))) return rc;
    CU(cudaStreamWaitEvent(xs, h->ev_gint, 0));
    if ((rc = tend(h->exterior, h->nexterior, xs))) return rc;
    CU(cudaEventRecord(h->ev_ext, xs));
    CU(cudaStreamWaitEvent(st, h->ev_gextk, 0));
    return tend(h->interior, h->ninterior, st);
  }
  if ((rc = filter(nullptr, nreal, st))) return rc;
  if (!par) {
    if ((rc = grad(nullptr, nreal, st))) return rc;
    if ((rc = column(0, nreal, 1, nullptr, st))) return rc;
    return tend(nullptr, nreal, st);
  }
  if ((rc = exchange_begin_t<R>(h, Q, HB_S, st))) return rc;
  if ((rc = grad(h->interior, h->ninterior, st))) return rc;
  if ((rc = exchange_end_t<R>(h, Q, HB_S, st))) return rc;
  if ((rc = grad(h->exterior, h->nexterior, st))) return rc;
  if ((rc = exchange_begin_t<R>(h, h->gradflux, HB_GF, st))) return rc;
  if ((rc = column(0, nreal, 1, nullptr, st))) return rc;
  if ((rc = tend(h->interior, h->ninterior, st))) return rc;
  if ((rc = exchange_end_t<R>(h, h->gradflux, HB_GF, st))) return rc;
  if ((rc = column(nreal, nghost, 0, nullptr, st))) return rc;
  return tend(h->exterior, h->nexterior, st);
}

// (Re)build the packed plus-side aux constants on `st`, before the first launch that reads them.  Callers
// invoke it on the stream every other stream of the schedule is ordered after.
template <class R>
int ensure_face_aux(cmdg_handle h, cudaStream_t st) {
  if (!CMDG_FACE_AUX || !h->aux_model || h->is_hb || !h->faux_dirty) return 0;
  const int64_t nreal = h->d.nrealelem;
  if (!h->fauxP) CU(cudaMalloc(&h->fauxP, (size_t)nreal * 6 * h->Nfp * 2 * sizeof(R) + 16));
  const AtmosParams<R> P = make_params<R>(h);
  if (nreal > 0) {
    pack_face_aux_kernel<R, 5><<<(unsigned)nreal, 160, 0, st>>>((const R *)h->aux, h->conn, P.naux, P.a_Phi,
                                                                P.a_ref_p, (R *)h->fauxP);
    CU(cudaGetLastError());
    h->launches++;
  }
  h->faux_dirty = false;
  return 0;
}

template <class R>
TendArgs<R> base_args(cmdg_handle h) {
  TendArgs<R> a{};
  a.aux = (const R *)h->aux;
  a.gradflux = (const R *)h->gradflux;
  a.vgeoP = (const R *)h->vgeoP;
  a.sgeoP = (const R *)h->sgeoP;
  a.conn = h->conn;
  a.fauxP = (const R *)h->fauxP;
  a.D = (const R *)h->Ddev;
  a.aux_out = h->d.write_aux_diagnostics ? (R *)h->aux : nullptr;
  a.pf_dist = h->pf_dist;
  a.F2 = (const R *)h->F2dev;
  a.Fn = (const R *)h->FnDev;
  a.nreal = (int)h->d.nrealelem;
  a.nelem = (int)h->d.nelem;
  return a;
}

template <class R>
int eval_with_tracers(cmdg_handle h, void *dQ, const void *Q, void *Qout, double alpha, double beta,
                      double rkb_dt, double t, bool exchange_q, bool write_diag, cudaStream_t st);

// One full evaluation in the reference's order (DGModel.jl:85-427).  With `Qout` set the
// stage update is fused (cmdg_lsrk_steps); the exchange schedule is then exterior-first, so
// that the halo of the *next* stage's state is in flight while the interior elements of this
// stage are computed.
template <class R>
int tendency_t(cmdg_handle h, void *dQ, void *Q, double t, double alpha, double beta,
               cudaStream_t st) {
  if (h->is_hb) return hb_eval_t<R>(h, dQ, Q, nullptr, alpha, beta, 0.0, st);
  if (int rc0 = ensure_face_aux<R>(h, st)) return rc0;
  if (h->ntracers) return eval_with_tracers<R>(h, dQ, Q, nullptr, alpha, beta, 0.0, t, true, true, st);
  const bool par = h->comm && !h->nabrtorank.empty();
  const int64_t nreal = h->d.nrealelem;
  TendArgs<R> a = base_args<R>(h);
  a.Q = (const R *)Q;
  a.dQ = (R *)dQ;
  a.Qout = nullptr;
  a.alpha = (R)alpha;
  a.beta = (R)beta;
  a.t = (R)t;
  GradArgs<R> ga{(const R *)Q, (const R *)h->aux, (R *)h->gradflux, (const R *)h->vgeoP,
                 (const R *)h->sgeoP, h->conn, nullptr, (const R *)h->Ddev, (R *)h->F2dev, (R *)h->FnDev,
                 (R *)h->Qhg, h->pf_dist};
  int rc;
  if (!par) {
    if (h->visc && (rc = second_order_passes<R>(h, ga, Q, false, false, false, st))) return rc;
    a.elems = nullptr;
    return launch_tendency<R>(h, a, nreal, st);
  }
  if ((rc = exchange_begin_t<R>(h, Q, h->d.nstate, st))) return rc;
  if (h->visc && (rc = second_order_passes<R>(h, ga, Q, true, false, true, st))) return rc;
  a.elems = h->interior;
  if ((rc = launch_tendency<R>(h, a, h->ninterior, st))) return rc;
  if (h->visc) {
    if ((rc = exchange_end_t<R>(h, h->F2dev, 12, st))) return rc;
  } else {
    if ((rc = exchange_end_t<R>(h, Q, h->d.nstate, st))) return rc;
  }
  a.elems = h->exterior;
  return launch_tendency<R>(h, a, h->nexterior, st);
}

// One evaluation with passive tracers, serial halo schedule: [Q halo] -> gradient kernel (all real
// elements; also nu per node) -> tracer gradient kernel -> F2 and F2chi halos -> tendency kernel of the
// five dynamic states -> tracer tendency kernel.  `Qout` != NULL fuses the RK stage update of all
// 5 + N columns.  The tracers are a side path of configs[0]: no interior / exterior overlap here.
template <class R>
int eval_with_tracers(cmdg_handle h, void *dQ, const void *Q, void *Qout, double alpha, double beta,
                      double rkb_dt, double t, bool exchange_q, bool write_diag, cudaStream_t st) {
  const bool par = h->comm && !h->nabrtorank.empty();
  const int64_t nreal = h->d.nrealelem;
  int rc;
  if (nreal <= 0) return 0;
  if (par && exchange_q) {
    if ((rc = exchange_begin_t<R>(h, const_cast<void *>(Q), h->d.nstate, st))) return rc;
    if ((rc = exchange_end_t<R>(h, const_cast<void *>(Q), h->d.nstate, st))) return rc;
  }
  if ((rc = ensure_const_D<R>(h, st))) return rc;
  const AtmosParams<R> P = make_params<R>(h);
  TracerArgs<R> ta{};
  ta.Q = (const R *)Q;
  ta.dQ = (R *)dQ;
  ta.Qout = (R *)Qout;
  ta.gradflux = write_diag ? (R *)h->gradflux : nullptr;
  ta.Nu = h->d.turbulence == CMDG_TURB_SMAGORINSKY ? (const R *)h->NuDev : nullptr;
  ta.F2chi = (R *)h->F2chi;
  ta.vgeoP = (const R *)h->vgeoP;
  ta.sgeoP = (const R *)h->sgeoP;
  ta.conn = h->conn;
  ta.elems = nullptr;
  ta.nt = h->ntracers;
  for (int i = 0; i < CMDG_MAX_TRACERS; ++i) ta.delta[i] = i < h->ntracers ? (R)h->d.tracer_delta_chi[i] : R(0);
  ta.alpha = (R)alpha;
  ta.beta = (R)beta;
  ta.rkb_dt = (R)rkb_dt;
  ta.visc = h->visc ? 1 : 0;
  ta.rusanov = h->d.nf_first == CMDG_NF_RUSANOV ? 1 : 0;
  if (h->visc) {
    GradArgs<R> ga{(const R *)Q, (const R *)h->aux, write_diag ? (R *)h->gradflux : nullptr, (const R *)h->vgeoP,
                   (const R *)h->sgeoP, h->conn, nullptr, (const R *)h->Ddev, (R *)h->F2dev, (R *)h->FnDev,
                   (R *)h->Qhg, h->pf_dist};
    ga.Nu = h->d.turbulence == CMDG_TURB_SMAGORINSKY ? (R *)h->NuDev : nullptr;
    if ((rc = launch_gradient<R>(h, ga, nreal, st))) return rc;
    if (h->timing) cudaEventRecord(timing_event(h, CMDG_KCLASS_TRACER_GRADIENT), st);
    tracer_gradient_kernel<R, 5><<<(unsigned)nreal, Dims<5>::BLOCK, 0, st>>>(ta, P);
    if (h->timing) cudaEventRecord(timing_event(h, CMDG_KCLASS_TRACER_GRADIENT), st);
    CU(cudaGetLastError());
    h->launches++;
    if (par) {
      if ((rc = exchange_begin_t<R>(h, h->F2dev, 12, st))) return rc;
      if ((rc = exchange_end_t<R>(h, h->F2dev, 12, st))) return rc;
      if ((rc = exchange_begin_t<R>(h, h->F2chi, 3 * h->ntracers, st))) return rc;
      if ((rc = exchange_end_t<R>(h, h->F2chi, 3 * h->ntracers, st))) return rc;
    }
  }
  TendArgs<R> a = base_args<R>(h);
  a.Q = (const R *)Q;
  a.dQ = (R *)dQ;
  a.Qout = (R *)Qout;
  a.alpha = (R)alpha;
  a.beta = (R)beta;
  a.rkb_dt = (R)rkb_dt;
  a.t = (R)t;
  a.elems = nullptr;
  if (!write_diag) a.aux_out = nullptr;
  if ((rc = launch_tendency<R>(h, a, nreal, st))) return rc;
  if (h->timing) cudaEventRecord(timing_event(h, CMDG_KCLASS_TRACER_TENDENCY), st);
  tracer_tendency_kernel<R, 5><<<(unsigned)nreal, Dims<5>::BLOCK, 0, st>>>(ta, P);
  if (h->timing) cudaEventRecord(timing_event(h, CMDG_KCLASS_TRACER_TENDENCY), st);
  CU(cudaGetLastError());
  h->launches++;
  if (par && Qout) {
    if ((rc = exchange_begin_t<R>(h, Qout, h->d.nstate, st))) return rc;
    if ((rc = exchange_end_t<R>(h, Qout, h->d.nstate, st))) return rc;
  }
  return 0;
}

template <class R>
int lsrk_update_t(cmdg_handle h, void *dQ, void *Q, double rka, double rkb, double dt,
                  cudaStream_t st) {
  const size_t n = (size_t)h->d.nrealelem * h->d.nstate * h->Np;
  if (n == 0) return 0;
  const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
  lsrk_update_kernel<R><<<blocks, 256, 0, st>>>((R *)dQ, (R *)Q, (R)rka, (R)rkb, (R)dt, n);
  CU(cudaGetLastError());
  h->launches++;
  return 0;
}

// Filters.apply! on real elements; Wh / Wv are row-major device matrices
template <class R>
int filter_apply_t(cmdg_handle h, void *Q, int nstate, int target, unsigned mask, const void *Wh,
                   const void *Wv, int direction, cudaStream_t st, int julia_layout = 0) {
  const int64_t nreal = h->d.nrealelem;
  if (nreal <= 0) return 0;
  int a_rho = 0, a_rhoe = 0;
  if (target == CMDG_FILTER_ATMOS_PERTURBATIONS) {
    if (h->is_hb || h->d.ref_state != CMDG_REF_HYDROSTATIC || !h->aux || nstate != 5)
      return fail(h, CMDG_ERR_INVALID, "AtmosFilterPerturbations needs a dry AtmosModel with a HydrostaticState and bound state_auxiliary");
    const AtmosParams<R> P = make_params<R>(h);
    a_rho = P.a_ref_rho;
    a_rhoe = P.a_ref_rho + 3;   // ref_state: rho, p, T, rhoe, ... (ref_state.jl:36-47)
    mask = 0x1f;
  }
  const int do_h = direction != CMDG_DIR_VERTICAL, do_v = direction != CMDG_DIR_HORIZONTAL;
  if (nstate <= 5)
    filter_kernel<R, 5, 5><<<(unsigned)nreal, Dims<5>::BLOCK, 0, st>>>(
        (R *)Q, (const R *)h->aux, (const R *)Wh, (const R *)Wv, nstate, h->d.naux, mask, target, a_rho,
        a_rhoe, do_h, do_v, julia_layout);
  else
    return fail(h, CMDG_ERR_UNSUPPORTED, "filter: at most 5 states");
  CU(cudaGetLastError());
  h->launches++;
  return 0;
}

// Julia (column-major) Nq x Nq device matrix -> row-major device copy
template <class R>
int transpose_small(cmdg_handle h, const void *src_dev, int n, void **dst_dev) {
  std::vector<R> a(n * n), b(n * n);
  CU(cudaMemcpy(a.data(), src_dev, a.size() * sizeof(R), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) b[i * n + j] = a[i + n * j];
  if (!*dst_dev) CU(cudaMalloc(dst_dev, b.size() * sizeof(R)));
  CU(cudaMemcpy(*dst_dev, b.data(), b.size() * sizeof(R), cudaMemcpyHostToDevice));
  return 0;
}

template <class R>
int lsrk_steps_t(cmdg_handle h, void *Q, void *dQ, double t0, double dt, int nstage,
                 const double *rka, const double *rkb, const double *rkc, int64_t nsteps,
                 cudaStream_t st, void *host_Q = nullptr) {
  // host_Q != NULL: pipelined host path (cmdg_lsrk_steps_host has enqueued the chunked upload on h->hp.copy)
  const bool par = h->comm && !h->nabrtorank.empty();
  const int64_t nreal = h->d.nrealelem;
  const size_t bytes = (size_t)h->d.nelem * h->d.nstate * h->Np * sizeof(R);
  if (!h->Qtmp) CU(cudaMalloc(&h->Qtmp, bytes));
  if (int rc0 = ensure_face_aux<R>(h, st)) return rc0;
  R *cur = (R *)Q, *nxt = (R *)h->Qtmp;
  if (par && !h->is_hb) {
    // ghosts of the initial state
    int rc;
    if ((rc = exchange_begin_t<R>(h, cur, h->d.nstate, st))) return rc;
    if ((rc = exchange_end_t<R>(h, cur, h->d.nstate, st))) return rc;
  }
  h->tev_used = 0;
  // Euler path on > 1 rank: exterior chain on the side stream, interior kernel on the caller's stream.
  // Interior elements have no ghost neighbours, so the interior kernel of stage s only needs the real
  // elements of stage s-1 (exterior + interior kernels); the exterior kernel of stage s needs those and
  // the ghosts unpacked by its own stream.  Per stage: max(interior, exterior + pack + NCCL + unpack).
  const bool overlap = par && !h->is_hb && !h->visc && !h->ntracers && h->step_filter_target < 0 && h->overlap_exterior &&
                       h->ext_stream && h->nexterior > 0 && h->ninterior > 0;
  // Second-order path (no hyperdiffusion, no tracers): the same two chains with two halos per stage,
  //   side stream:  grad_ext(s) -> [F2 halo] -> tend_ext(s) -> [Q halo] -> grad_ext(s+1) ...
  //   main stream:  grad_int(s) ---------------> tend_int(s) -----------> grad_int(s+1) ...
  // coupled kernel to kernel: tend_int(s) needs grad_ext(s) (Fn of its real neighbours), tend_ext(s) needs
  // grad_int(s) and its own F2 unpack, grad_int(s+1) needs tend_ext(s), grad_ext(s+1) needs tend_int(s) and its
  // own Q unpack.  (The reference overlaps the gradient-flux exchange with volume_tendency!, DGModel.jl:195-223.)
  const bool overlap2 = par && !h->is_hb && h->visc && !h->hyper && !h->ntracers && h->step_filter_target < 0 &&
                        h->overlap_exterior && h->ext_stream && h->nexterior > 0 && h->ninterior > 0;
  // Euler path, single launch (CMDG_OVERLAP=2, default when the driver offers stream memory operations): ONE
  // kernel per stage over [exterior..., interior...]; its exterior blocks count themselves done, the side stream
  // waits for that count with cuStreamWaitValue32 and runs pack -> NCCL -> unpack while the same kernel goes on with
  // the interior elements; the next stage's kernel waits for the unpack.  No small exterior launch (which costs about
  // twice as much per element as the big one), no concurrent kernels competing for SM slots.
  const bool overlap1 = overlap && h->overlap_mode >= 2 && h->all_list && h->ext_done && stream_wait_value32();
  // HBModel: two chains as well (hb_eval_t)
  const bool overlap_hb = par && h->is_hb && h->overlap_exterior && h->ext_stream && h->int_stacks && h->ext_stacks &&
                          h->step_filter_target < 0 && (h->d.nelem - nreal) % std::max(1, h->d.nvertelem) == 0;
  cudaStream_t xs = h->ext_stream;
  bool ext_pending = false;
  if (overlap || overlap2 || overlap_hb) {
    if (int rc = ensure_const_D<R>(h, st)) return rc;
    CU(cudaEventRecord(h->ev_int, st));
    CU(cudaStreamWaitEvent(xs, h->ev_int, 0));
  }
  h->tl_used = 0;
  for (int64_t step = 0; step < nsteps; ++step) {
    const double time = t0 + (double)step * dt;
    h->tl_on = h->timing && !h->tl_path.empty() && step + 2 >= nsteps;
    h->tl_step = (int)step;
    h->tl_stage = 0;
    tl_mark(h, st, TL_STEP_BEGIN);
    for (int s = 0; s < nstage; ++s) {
      h->tl_stage = s;
      if (h->is_hb) {
        int rc = hb_eval_t<R>(h, dQ, cur, nxt, 1.0, rka[s], (double)((R)rkb[s] * (R)dt), st, overlap_hb);
        if (rc) return rc;
        if (overlap_hb) ext_pending = true;
        // the vertical filters act on the stage state itself: the ghost layer of the new state is
        // refreshed by the next evaluation's exchange
        R *tmp2 = cur;
        cur = nxt;
        nxt = tmp2;
        continue;
      }
      if (h->ntracers) {
        int rc = eval_with_tracers<R>(h, dQ, cur, nxt, 1.0, rka[s], (double)((R)rkb[s] * (R)dt),
                                      time + rkc[s] * dt, false, s == nstage - 1, st);
        if (rc) return rc;
        R *tmp2 = cur;
        cur = nxt;
        nxt = tmp2;
        continue;
      }
      TendArgs<R> a = base_args<R>(h);
      a.Q = cur;
      a.dQ = (R *)dQ;
      a.Qout = nxt;
      a.alpha = R(1);
      a.beta = (R)rka[s];
      a.rkb_dt = (R)((R)rkb[s] * (R)dt);
      a.t = (R)(time + rkc[s] * dt);
      // aux diagnostics are refreshed by the last stage only (they are read after steps)
      if (s != nstage - 1) a.aux_out = nullptr;
      GradArgs<R> ga{cur, (const R *)h->aux, (R *)h->gradflux, (const R *)h->vgeoP,
                     (const R *)h->sgeoP, h->conn, nullptr, (const R *)h->Ddev, (R *)h->F2dev, (R *)h->FnDev,
                     (R *)h->Qhg, h->pf_dist};
      // like the aux diagnostics, state_gradient_flux is only read after the step: the last stage
      // refreshes it (the tendency kernel consumes F2 / Fn, not GF); DryBiharmonic with a non-zero
      // viscous closure re-reads GF in hyper_flux_kernel, so it is kept then
      const bool gf_needed = h->hyper && (h->d.turbulence == CMDG_TURB_SMAGORINSKY || h->d.turb_param != 0.0);
      if (s != nstage - 1 && !gf_needed) ga.gradflux = nullptr;
      int rc;
      if (overlap2) {
        const int64_t next = h->nexterior, ninr = h->ninterior;
        ga.elems = h->exterior;
        if ((rc = launch_gradient<R>(h, ga, next, xs))) return rc;
        CU(cudaEventRecord(h->ev_gextk, xs));
        if ((rc = exchange_begin_t<R>(h, h->F2dev, 12, xs))) return rc;
        if ((rc = exchange_end_t<R>(h, h->F2dev, 12, xs))) return rc;
        ga.elems = h->interior;
        if ((rc = launch_gradient<R>(h, ga, ninr, st))) return rc;
        CU(cudaEventRecord(h->ev_gint, st));
        CU(cudaStreamWaitEvent(xs, h->ev_gint, 0));
        a.elems = h->exterior;
        if ((rc = launch_tendency<R>(h, a, next, xs))) return rc;
        CU(cudaEventRecord(h->ev_extk, xs));
        if ((rc = exchange_begin_t<R>(h, nxt, h->d.nstate, xs))) return rc;
        if ((rc = exchange_end_t<R>(h, nxt, h->d.nstate, xs))) return rc;
        CU(cudaEventRecord(h->ev_ext, xs));
        CU(cudaStreamWaitEvent(st, h->ev_gextk, 0));
        a.elems = h->interior;
        if ((rc = launch_tendency<R>(h, a, ninr, st))) return rc;
        CU(cudaEventRecord(h->ev_int, st));
        CU(cudaStreamWaitEvent(xs, h->ev_int, 0));
        CU(cudaStreamWaitEvent(st, h->ev_extk, 0));
        ext_pending = true;
        R *tmp = cur;
        cur = nxt;
        nxt = tmp;
        continue;
      }
      // tail prefetch (Euler path): the last blocks of this stage's (last) launch warm L2 for the first
      // wave(s) of the next stage's launch(es): CMDG_TAILPF = number of elements per next launch (0 = off)
      const int tailpf = (h->visc || s == nstage - 1 && step + 1 == nsteps) ? 0 : h->tail_pf;
      a.pfn_Q = nxt;
      if (!par) {
        if (h->visc && (rc = second_order_passes<R>(h, ga, cur, false, true, false, st))) return rc;
        // (diagnostic: CMDG_FORCE_LIST=1 runs a rank's launch list [exterior..., interior...] without a communicator)
        static const bool force_list = getenv("CMDG_FORCE_LIST") != nullptr;
        a.elems = (force_list && h->all_list) ? h->all_list : nullptr;
        a.pfn_list[0] = a.elems;
        a.pfn_n[0] = (int)std::min<int64_t>(nreal, tailpf);
        const bool pipe_first = host_Q && step == 0 && s == 0;
        const bool pipe_last = host_Q && step + 1 == nsteps && s == nstage - 1;
        if (pipe_first || pipe_last) {
          // first stage of the call: chunk c of the launch list needs upload chunks 0..c; last stage: range by
          // range, each range's new state goes back to the host while the next range is computed
          cmdg_handle_s::HostPipe &hp = h->hp;
          a.pfn_n[0] = 0;
          const size_t esz = (size_t)h->d.nstate * h->Np;
          for (int c = 0; c < hp.nch; ++c) {
            int64_t n;
            if (pipe_first) {
              CU(cudaStreamWaitEvent(st, hp.ev_up[c], 0));
              a.elems = hp.ready_list + hp.ready_off[c];
              n = hp.ready_off[c + 1] - hp.ready_off[c];
            } else {
              a.elems = hp.identity + hp.first[c];
              n = hp.first[c + 1] - hp.first[c];
            }
            if ((rc = launch_tendency<R>(h, a, n, st))) return rc;
            if (pipe_last) {
              CU(cudaEventRecord(hp.ev_k[c], st));
              CU(cudaStreamWaitEvent(hp.copy, hp.ev_k[c], 0));
              const size_t o = (size_t)hp.first[c] * esz;
              CU(cudaMemcpyAsync((R *)host_Q + o, nxt + o, (size_t)n * esz * sizeof(R), cudaMemcpyDeviceToHost,
                                 hp.copy));
            }
          }
          if (pipe_last) {
            CU(cudaEventRecord(hp.ev_end, hp.copy));
            CU(cudaStreamWaitEvent(st, hp.ev_end, 0));
          }
        } else if ((rc = launch_tendency<R>(h, a, nreal, st))) return rc;
      } else {
        if (overlap1) {
          a.elems = h->all_list;
          a.ext_done = h->ext_done;
          a.n_signal = (int)h->nexterior;
          a.pfn_list[0] = h->all_list;
          a.pfn_n[0] = (int)std::min<int64_t>(nreal, tailpf);
          if (ext_pending) CU(cudaStreamWaitEvent(st, h->ev_ext, 0));   // ghosts of this stage's input state
          if ((rc = launch_tendency<R>(h, a, nreal, st))) return rc;
          h->ext_target += (unsigned)h->nexterior;
          if (stream_wait_value32()((CUstream)xs, (CUdeviceptr)(uintptr_t)h->ext_done, h->ext_target,
                                    CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
            return fail(h, CMDG_ERR_CUDA, "cuStreamWaitValue32 failed");
          if ((rc = exchange_begin_t<R>(h, nxt, h->d.nstate, xs))) return rc;
          if ((rc = exchange_end_t<R>(h, nxt, h->d.nstate, xs))) return rc;
          CU(cudaEventRecord(h->ev_ext, xs));
          ext_pending = true;
          R *tmp = cur;
          cur = nxt;
          nxt = tmp;
          continue;
        }
        if (overlap) {
          a.elems = h->exterior;
          if ((rc = launch_tendency<R>(h, a, h->nexterior, xs))) return rc;
          // interior elements never read ghosts: the next stage's interior kernel only needs this exterior
          // KERNEL (the real exterior elements of the new state), not the halo that follows it
          CU(cudaEventRecord(h->ev_extk, xs));
          if ((rc = exchange_begin_t<R>(h, nxt, h->d.nstate, xs))) return rc;
          if ((rc = exchange_end_t<R>(h, nxt, h->d.nstate, xs))) return rc;
          CU(cudaEventRecord(h->ev_ext, xs));
          a.elems = h->interior;
          a.pfn_list[0] = h->exterior;
          a.pfn_n[0] = (int)std::min<int64_t>(h->nexterior, tailpf);
          a.pfn_list[1] = h->interior;
          a.pfn_n[1] = (int)std::min<int64_t>(h->ninterior, tailpf);
          if ((rc = launch_tendency<R>(h, a, h->ninterior, st))) return rc;
          CU(cudaEventRecord(h->ev_int, st));
          // next stage: the exterior kernel (side stream, after its own unpack) needs this interior kernel;
          // the interior kernel (main stream, back to back with this one) needs this exterior kernel
          CU(cudaStreamWaitEvent(xs, h->ev_int, 0));
          CU(cudaStreamWaitEvent(st, h->ev_extk, 0));
          ext_pending = true;
          R *tmp = cur;
          cur = nxt;
          nxt = tmp;
          continue;
        }
        if (h->visc && (rc = second_order_passes<R>(h, ga, cur, true, true, false, st))) return rc;
        // a per-step filter changes the new state after the last stage: its halo goes out after
        // the filter instead of overlapping the interior kernel
        const bool filt_now = h->step_filter_target >= 0 && s == nstage - 1;
        a.elems = h->exterior;
        if ((rc = launch_tendency<R>(h, a, h->nexterior, st))) return rc;
        if (!filt_now && (rc = exchange_begin_t<R>(h, nxt, h->d.nstate, st))) return rc;
        a.elems = h->interior;
        if ((rc = launch_tendency<R>(h, a, h->ninterior, st))) return rc;
        if (!filt_now && (rc = exchange_end_t<R>(h, nxt, h->d.nstate, st))) return rc;
      }
      R *tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
    if (h->step_filter_target >= 0 && !h->is_hb) {
      int rc = filter_apply_t<R>(h, cur, h->d.nstate, h->step_filter_target, h->step_filter_mask,
                                 h->stepWh, h->stepWv, h->step_filter_dir, st);
      if (rc) return rc;
      if (par) {
        if ((rc = exchange_begin_t<R>(h, cur, h->d.nstate, st))) return rc;
        if ((rc = exchange_end_t<R>(h, cur, h->d.nstate, st))) return rc;
      }
    }
  }
  h->tl_on = false;
  // the caller's stream sees the last halo (ghosts of the final state) as well
  if (ext_pending) CU(cudaStreamWaitEvent(st, h->ev_ext, 0));
  // (pipelined host path: the new state has already gone to the host range by range)
  if (cur != (R *)Q && !host_Q) CU(cudaMemcpyAsync(Q, cur, bytes, cudaMemcpyDeviceToDevice, st));
  // the reference leaves dQ scaled by RKA[1] after the last stage (:130-141)
  const size_t n = (size_t)nreal * h->d.nstate * h->Np;
  if (rka[0] == 0.0) {
    CU(cudaMemsetAsync(dQ, 0, n * sizeof(R), st));
  } else if (n) {
    scale_kernel<R><<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(
        (R *)dQ, (R)rka[0], n);
    h->launches++;
  }
  if (h->timing) {
    CU(cudaStreamSynchronize(st));
    double ms = 0;
    int64_t nl = 0;
    for (int c = 0; c < CMDG_KCLASS_COUNT; ++c) h->class_ms[c] = 0, h->class_n[c] = 0;
    for (size_t i = 0; i + 1 < h->tev_used; i += 2) {
      float x = 0;
      cudaEventElapsedTime(&x, h->tev[i], h->tev[i + 1]);
      const int c = h->tev_class[i];
      h->class_ms[c] += x;
      h->class_n[c]++;
      if (c == CMDG_KCLASS_TENDENCY) ms += x, nl++;
    }
    h->last_ms = ms;
    h->last_nl = nl;
    tl_dump(h);
  }
  return 0;
}

// Derive the per-face neighbour descriptor from vmap+ / elemtobndy and verify that every
// face follows the conforming tensor-product pattern of Grids.jl:559-637 (orientation 1 or 3).
int build_conn(cmdg_handle h, const int64_t *vmapM_dev, const int64_t *vmapP_dev,
               const int64_t *elemtobndy_dev) {
  const int NQ = h->Nq, NP = h->Np, NFP = h->Nfp;
  const int64_t nreal = h->d.nrealelem, nelem = h->d.nelem;
  std::vector<int64_t> vP((size_t)nreal * 6 * NFP), vM((size_t)nreal * 6 * NFP), bnd((size_t)nreal * 6);
  CU(cudaMemcpy(vP.data(), vmapP_dev, vP.size() * 8, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(vM.data(), vmapM_dev, vM.size() * 8, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(bnd.data(), elemtobndy_dev, bnd.size() * 8, cudaMemcpyDeviceToHost));
  auto f2v = [NQ](int f, int a, int b) {
    switch (f) {
      case 0: return NQ * (a + NQ * b);
      case 1: return (NQ - 1) + NQ * (a + NQ * b);
      case 2: return a + NQ * NQ * b;
      case 3: return a + NQ * ((NQ - 1) + NQ * b);
      case 4: return a + NQ * b;
      default: return a + NQ * (b + NQ * (NQ - 1));
    }
  };
  std::vector<int2> conn((size_t)nreal * 6);
  h->nbr_host.assign((size_t)nreal * 6, -1);
  for (int64_t e = 0; e < nreal; ++e)
    for (int f = 0; f < 6; ++f) {
      const int64_t *pm = &vM[((size_t)e * 6 + f) * NFP];
      const int64_t *pp = &vP[((size_t)e * 6 + f) * NFP];
      for (int n = 0; n < NFP; ++n)
        if (pm[n] != (int64_t)NP * e + f2v(f, n % NQ, n / NQ) + 1)
          return fail(h, CMDG_ERR_UNSUPPORTED, "vmap- does not follow the Grids.jl face-mask order");
      const int64_t tag = bnd[(size_t)e * 6 + f];
      if (tag != 0) {
        if (tag < 0 || tag > h->d.nbc)
          return fail(h, CMDG_ERR_INVALID, "elemtobndy tag outside 1..nbc (BoundsError(bcs, bctag))");
        conn[(size_t)e * 6 + f] = make_int2((int)e, conn_meta(f, 0, (int)tag));
        continue;
      }
      const int64_t ep = (pp[0] - 1) / NP;
      if (ep < 0 || ep >= nelem) return fail(h, CMDG_ERR_INVALID, "vmap+ points outside the state array");
      int found = -1;
      for (int fp = 0; fp < 6 && found < 0; ++fp)
        for (int flip = 0; flip < 2 && found < 0; ++flip) {
          bool ok = true;
          for (int n = 0; n < NFP && ok; ++n) {
            const int a = flip ? NQ - 1 - n % NQ : n % NQ;
            ok = pp[n] == (int64_t)NP * ep + f2v(fp, a, n / NQ) + 1;
          }
          if (ok) found = fp | (flip << 3);
        }
      if (found < 0)
        return fail(h, CMDG_ERR_UNSUPPORTED,
                    "vmap+ face is not a conforming tensor-product face (orientation 1 or 3)");
      conn[(size_t)e * 6 + f] = make_int2((int)ep, found);
      h->nbr_host[(size_t)e * 6 + f] = (int)ep;
    }
  CU(cudaMalloc(&h->conn, conn.size() * sizeof(int2) + 16));
  CU(cudaMemcpy(h->conn, conn.data(), conn.size() * sizeof(int2), cudaMemcpyHostToDevice));
  return 0;
}

int to_zero_based_list(cmdg_handle h, const int64_t *dev, int64_t n, int **out) {
  *out = nullptr;
  if (n <= 0) return 0;
  std::vector<int64_t> tmp(n);
  CU(cudaMemcpy(tmp.data(), dev, n * 8, cudaMemcpyDeviceToHost));
  std::vector<int> v(n);
  for (int64_t i = 0; i < n; ++i) v[i] = (int)(tmp[i] - 1);
  CU(cudaMalloc(out, n * sizeof(int)));
  CU(cudaMemcpy(*out, v.data(), n * sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

int to_zero_based_map(cmdg_handle h, const int64_t *dev, int64_t n, int64_t **out) {
  *out = nullptr;
  if (n <= 0) return 0;
  std::vector<int64_t> tmp(n);
  CU(cudaMemcpy(tmp.data(), dev, n * 8, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; ++i) tmp[i] -= 1;
  CU(cudaMalloc(out, n * 8));
  CU(cudaMemcpy(*out, tmp.data(), n * 8, cudaMemcpyHostToDevice));
  return 0;
}

template <class R>
int bind_grid_t(cmdg_handle h, const void *vgeo, const void *sgeo, const void *D) {
  const int64_t nreal = h->d.nrealelem;
  const int NP = h->Np, NFP = h->Nfp, NQ = h->Nq;
  CU(cudaMalloc(&h->vgeoP, (size_t)nreal * 10 * NP * sizeof(R) + 16));
  CU(cudaMalloc(&h->sgeoP, (size_t)nreal * 6 * 4 * NFP * sizeof(R) + 16));
  if (nreal > 0) {
    const size_t nv = (size_t)nreal * NP, ns = (size_t)nreal * 6 * NFP;
    pack_vgeo_kernel<R><<<(unsigned)((nv + 255) / 256), 256>>>((R *)h->vgeoP, (const R *)vgeo, NP, 25, nreal);
    pack_sgeo_kernel<R><<<(unsigned)((ns + 255) / 256), 256>>>((R *)h->sgeoP, (const R *)sgeo, NFP, nreal);
    CU(cudaGetLastError());
    h->launches += 2;
  }
  // D arrives in Julia (column-major) order; keep a row-major copy D[a][b] = D_julia[a, b]
  std::vector<R> Dj(NQ * NQ), Dr(NQ * NQ);
  CU(cudaMemcpy(Dj.data(), D, Dj.size() * sizeof(R), cudaMemcpyDeviceToHost));
  for (int a = 0; a < NQ; ++a)
    for (int b = 0; b < NQ; ++b) Dr[a * NQ + b] = Dj[a + NQ * b];
  CU(cudaMalloc(&h->Ddev, Dr.size() * sizeof(R)));
  CU(cudaMemcpy(h->Ddev, Dr.data(), Dr.size() * sizeof(R), cudaMemcpyHostToDevice));
  h->Dhost.assign(64, 0.0);
  for (size_t i = 0; i < Dr.size(); ++i) h->Dhost[i] = (double)Dr[i];
  CU(cudaDeviceSynchronize());
  return 0;
}

template <class R>
int upload_rowmajor(cmdg_handle h, void **dst, const void *src_dev, int n) {
  std::vector<R> a(n * n), b(n * n);
  CU(cudaMemcpy(a.data(), src_dev, a.size() * sizeof(R), cudaMemcpyDeviceToHost));
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) b[r * n + c] = a[r + n * c];
  if (!*dst) CU(cudaMalloc(dst, b.size() * sizeof(R)));
  CU(cudaMemcpy(*dst, b.data(), b.size() * sizeof(R), cudaMemcpyHostToDevice));
  return 0;
}
#define DISPATCH_FT(h, expr64, expr32) ((h)->d.float_bytes == CMDG_F64 ? (expr64) : (expr32))

}  // namespace

// ======================================================================================
// C ABI
// ======================================================================================
extern "C" {

int cmdg_version(void) { return CMDG_VERSION; }

const char *cmdg_last_error(cmdg_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int cmdg_create(const cmdg_desc *d, cmdg_handle *out) {
  cmdg_handle h = nullptr;
  if (!d || !out) return fail(nullptr, CMDG_ERR_INVALID, "null argument");
  *out = nullptr;
  if (d->struct_bytes != (int32_t)sizeof(cmdg_desc))
    return fail(nullptr, CMDG_ERR_INVALID, "cmdg_desc size mismatch (ABI)");
  if (d->float_bytes != CMDG_F64 && d->float_bytes != CMDG_F32)
    return fail(nullptr, CMDG_ERR_UNSUPPORTED, "float type must be Float64 or Float32");
  if (d->dim != 3) return fail(nullptr, CMDG_ERR_UNSUPPORTED, "only dim = 3 is supported");
  if (d->N != 4)
    return fail(nullptr, CMDG_ERR_UNSUPPORTED, "only polynomial order N = 4 is compiled in");
  if (d->model != CMDG_MODEL_ATMOS_DRY && d->model != CMDG_MODEL_HB)
    return fail(nullptr, CMDG_ERR_UNSUPPORTED,
                "unsupported balance law (only the dry AtmosModel and the ocean HBModel are compiled in)");
  if (d->model == CMDG_MODEL_HB) {
    if (d->nf_first != CMDG_NF_RUSANOV && d->nf_first != CMDG_NF_CENTRAL)
      return fail(nullptr, CMDG_ERR_UNSUPPORTED, "HBModel supports Rusanov / Central first-order fluxes");
    if (d->nf_second != CMDG_NF_CENTRAL || d->nf_gradient != CMDG_NF_CENTRAL)
      return fail(nullptr, CMDG_ERR_UNSUPPORTED, "second-order / gradient fluxes must be Central");
    if (d->nstate != 4 || d->naux != 8 || d->ngrad != 5 || d->ngradflux != 10)
      return fail(nullptr, CMDG_ERR_INVALID, "HBModel state sizes are 4 / 8 / 5 / 10");
    if (d->nvertelem <= 0 || d->nrealelem % d->nvertelem != 0 || d->nelem % d->nvertelem != 0)
      return fail(nullptr, CMDG_ERR_INVALID, "HBModel needs a stacked topology (nvertelem > 0)");
    if (d->diffusion_direction != CMDG_DIR_EVERY)
      return fail(nullptr, CMDG_ERR_UNSUPPORTED, "HBModel: diffusion_direction must be EveryDirection");
    if (d->nrealelem < 0 || d->nelem < d->nrealelem || d->nelem > 0x7fffffffLL)
      return fail(nullptr, CMDG_ERR_INVALID, "bad element counts");
    int ndev0 = 0;
    if (cudaGetDeviceCount(&ndev0) != cudaSuccess || ndev0 == 0)
      return fail(nullptr, CMDG_ERR_NODEVICE, "no CUDA device available (libcmdg has no CPU fallback)");
    h = new cmdg_handle_s();
    cudaGetDevice(&h->device);
    h->d = *d;
    h->Nq = d->N + 1;
    h->Np = h->Nq * h->Nq * h->Nq;
    h->Nfp = h->Nq * h->Nq;
    h->fb = d->float_bytes;
    h->is_hb = true;
    h->visc = true;
    if (const char *kv = getenv("CMDG_TIMELINE")) h->tl_path = kv;
    if (const char *kv = getenv("CMDG_OVERLAP")) h->overlap_exterior = atoi(kv) != 0;
    // NCCL stream and the exterior chain of the two-chain schedule (hb_eval_t) at high priority, as for the
    // atmosphere: their small kernels must get SM slots while the interior kernels have thousands of blocks queued
    int plo = 0, phi = 0;
    cudaDeviceGetStreamPriorityRange(&plo, &phi);
    const bool two = h->overlap_exterior;
    bool ok = cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, two ? phi : plo) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming) == cudaSuccess;
    if (ok && two)
      ok = cudaStreamCreateWithPriority(&h->ext_stream, cudaStreamNonBlocking, phi) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->ev_ext, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->ev_int, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->ev_gextk, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->ev_gint, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->ev_hb_fext, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->ev_hb_fint, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      delete h;
      return fail(nullptr, CMDG_ERR_CUDA, "cannot create stream/events");
    }
    *out = h;
    return CMDG_OK;
  }
  if (d->nf_first < CMDG_NF_RUSANOV || d->nf_first > CMDG_NF_ROE)
    return fail(nullptr, CMDG_ERR_UNSUPPORTED, "unsupported first-order numerical flux");
  if (d->nf_second != CMDG_NF_CENTRAL || d->nf_gradient != CMDG_NF_CENTRAL)
    return fail(nullptr, CMDG_ERR_UNSUPPORTED, "second-order / gradient fluxes must be Central");
  if (d->orientation < 0 || d->orientation > CMDG_ORIENT_SPHERICAL ||
      d->ref_state < 0 || d->ref_state > CMDG_REF_HYDROSTATIC ||
      d->turbulence < 0 || d->turbulence > CMDG_TURB_SMAGORINSKY)
    return fail(nullptr, CMDG_ERR_UNSUPPORTED, "unsupported orientation / reference state / turbulence model");
  if (d->sources & ~(CMDG_SRC_GRAVITY | CMDG_SRC_CORIOLIS | CMDG_SRC_HELD_SUAREZ | CMDG_SRC_RAYLEIGH_SPONGE))
    return fail(nullptr, CMDG_ERR_UNSUPPORTED, "unsupported source term");
  if ((d->sources & (CMDG_SRC_HELD_SUAREZ | CMDG_SRC_RAYLEIGH_SPONGE)) && d->orientation == CMDG_ORIENT_NONE)
    return fail(nullptr, CMDG_ERR_INVALID, "HeldSuarezForcing / RayleighSponge need an orientation");
  if ((d->sources & CMDG_SRC_HELD_SUAREZ) && !(d->day > 0))
    return fail(nullptr, CMDG_ERR_INVALID, "HeldSuarezForcing needs cmdg_desc.day > 0");
  if ((d->sources & CMDG_SRC_RAYLEIGH_SPONGE) && !(d->sponge_z_max > d->sponge_z_sponge))
    return fail(nullptr, CMDG_ERR_INVALID, "RayleighSponge needs z_max > z_sponge");
  if ((d->sources & CMDG_SRC_GRAVITY) && d->orientation == CMDG_ORIENT_NONE)
    return fail(nullptr, CMDG_ERR_INVALID, "Gravity needs an orientation");
  if (d->ref_state == CMDG_REF_HYDROSTATIC && d->orientation == CMDG_ORIENT_NONE)
    return fail(nullptr, CMDG_ERR_INVALID, "HydrostaticState needs an orientation");
  if (d->turbulence == CMDG_TURB_SMAGORINSKY && d->orientation == CMDG_ORIENT_NONE)
    return fail(nullptr, CMDG_ERR_UNSUPPORTED, "SmagorinskyLilly needs an orientation");
  if (d->nbc < 0 || d->nbc > 6) return fail(nullptr, CMDG_ERR_INVALID, "nbc must be 0..6");
  for (int i = 0; i < d->nbc; ++i)
    if (d->bc_kind[i] != CMDG_BC_FREESLIP && d->bc_kind[i] != CMDG_BC_NOSLIP)
      return fail(nullptr, CMDG_ERR_UNSUPPORTED, "unsupported boundary condition");
  const int nt = d->ntracers;
  if (nt < 0 || nt > CMDG_MAX_TRACERS)
    return fail(nullptr, CMDG_ERR_UNSUPPORTED, "NTracers: 0..4 passive tracers are supported");
  if (d->nstate != 5 + nt)
    return fail(nullptr, CMDG_ERR_UNSUPPORTED, "dry AtmosModel has 5 + ntracers prognostic states (moisture unsupported)");
  if (nt && d->nf_first == CMDG_NF_ROE)
    return fail(nullptr, CMDG_ERR_UNSUPPORTED, "NTracers: Rusanov / Central first-order fluxes only");
  if (nt && d->hyperdiffusion != CMDG_HYPER_NONE)
    return fail(nullptr, CMDG_ERR_UNSUPPORTED, "NTracers with hyperdiffusion is not supported");
  if (d->naux != expected_naux(*d))
    return fail(nullptr, CMDG_ERR_INVALID, "naux does not match the model's auxiliary state");
  const int gf = d->turbulence == CMDG_TURB_SMAGORINSKY ? 10 : 9;
  if (d->hyperdiffusion != CMDG_HYPER_NONE && d->hyperdiffusion != CMDG_HYPER_DRY_BIHARMONIC)
    return fail(nullptr, CMDG_ERR_UNSUPPORTED, "unsupported hyperdiffusion model");
  const int hyp = d->hyperdiffusion == CMDG_HYPER_DRY_BIHARMONIC;
  if (hyp && d->orientation == CMDG_ORIENT_NONE)
    return fail(nullptr, CMDG_ERR_INVALID, "DryBiharmonic needs an orientation");
  if (hyp && d->diffusion_direction != CMDG_DIR_HORIZONTAL)
    return fail(nullptr, CMDG_ERR_UNSUPPORTED,
                "DryBiharmonic: only diffusion_direction = HorizontalDirection() (the reference's 3-D EveryDirection kernel does not run)");
  if (hyp && !(d->hyper_tau > 0)) return fail(nullptr, CMDG_ERR_INVALID, "DryBiharmonic needs hyper_tau > 0");
  if (d->ngradflux != gf + 3 * nt || d->ngrad != (gf == 10 ? 5 : 4) + 4 * hyp + nt)
    return fail(nullptr, CMDG_ERR_INVALID, "ngrad/ngradflux do not match the model");
  if (d->nrealelem < 0 || d->nelem < d->nrealelem || d->nelem > 0x7fffffffLL)
    return fail(nullptr, CMDG_ERR_INVALID, "bad element counts");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, CMDG_ERR_NODEVICE, "no CUDA device available (libcmdg has no CPU fallback)");
  h = new cmdg_handle_s();
  cudaGetDevice(&h->device);
  h->d = *d;
  h->Nq = d->N + 1;
  h->Np = h->Nq * h->Nq * h->Nq;
  h->Nfp = h->Nq * h->Nq;
  h->fb = d->float_bytes;
  h->aux_model = d->orientation != CMDG_ORIENT_NONE || d->ref_state != CMDG_REF_NONE;
  const bool zero_visc = d->turbulence != CMDG_TURB_SMAGORINSKY && d->turb_param == 0.0;
  h->visc = !(d->skip_zero_viscosity && zero_visc) || hyp;
  h->hyper = hyp;
  h->ntracers = nt;
  if (const char *kv = getenv("CMDG_PF")) h->pf_dist = atoi(kv);
  if (const char *kv = getenv("CMDG_TAILPF")) h->tail_pf = atoi(kv);
  if (const char *kv = getenv("CMDG_TIMELINE")) h->tl_path = kv;
  if (const char *kv = getenv("CMDG_OVERLAP")) {
    h->overlap_exterior = atoi(kv) != 0;
    h->overlap_mode = atoi(kv);
  }
  // the NCCL send/recv kernel is launched while the interior kernel still has thousands of blocks
  // queued: on a stream of the same priority its CTAs would be dispatched after them, i.e. the halo
  // would start when the interior kernel ends.  High priority puts them in front (CMDG_COMM_PRIO=0: off).
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  const char *cp = getenv("CMDG_COMM_PRIO");
  cudaError_t e1 = cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking,
                                                (cp && atoi(cp) == 0) ? prio_lo : prio_hi);
  cudaError_t e2 = cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming);
  cudaError_t e3 = cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming);
  if (e1 == cudaSuccess) e1 = cudaStreamCreateWithPriority(&h->ext_stream, cudaStreamNonBlocking, prio_hi);
  if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&h->ev_ext, cudaEventDisableTiming);
  if (e3 == cudaSuccess) e3 = cudaEventCreateWithFlags(&h->ev_int, cudaEventDisableTiming);
  if (e3 == cudaSuccess) e3 = cudaEventCreateWithFlags(&h->ev_extk, cudaEventDisableTiming);
  if (e3 == cudaSuccess) e3 = cudaEventCreateWithFlags(&h->ev_gextk, cudaEventDisableTiming);
  if (e3 == cudaSuccess) e3 = cudaEventCreateWithFlags(&h->ev_gint, cudaEventDisableTiming);
  if (e3 == cudaSuccess) e3 = cudaEventCreateWithFlags(&h->ev_hb_fext, cudaEventDisableTiming);
  if (e3 == cudaSuccess) e3 = cudaEventCreateWithFlags(&h->ev_hb_fint, cudaEventDisableTiming);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
    delete h;
    return fail(nullptr, CMDG_ERR_CUDA, "cannot create stream/events");
  }
  *out = h;
  return CMDG_OK;
}

int cmdg_destroy(cmdg_handle h) {
  if (g_constD_owner == h) g_constD_owner = nullptr;
  if (!h) return CMDG_OK;
  cudaDeviceSynchronize();
  void *bufs[] = {h->vgeoP, h->sgeoP, h->fauxP, h->Ddev, h->conn, h->interior, h->exterior, h->vmapsend0,
                  h->vmaprecv0, h->sendbuf, h->recvbuf, h->Qtmp, h->Qdev, h->dQdev,
                  h->Fc, h->Fe, h->Imat, h->JcV, h->stepWh, h->stepWv, h->tmpWh, h->tmpWv, h->courant_dev, h->crash_dev,
                  h->all_list, h->ext_done, h->int_stacks, h->ext_stacks, h->F2dev, h->FnDev, h->Qhg, h->Qhd,
                  h->NuDev, h->F2chi};
  for (void *p : bufs)
    if (p) cudaFree(p);
  if (h->hp.ready_list) cudaFree(h->hp.ready_list);
  if (h->hp.identity) cudaFree(h->hp.identity);
  if (h->hp.copy) cudaStreamDestroy(h->hp.copy);
  for (cudaEvent_t e : h->hp.ev_up) cudaEventDestroy(e);
  for (cudaEvent_t e : h->hp.ev_k) cudaEventDestroy(e);
  if (h->hp.ev_begin) cudaEventDestroy(h->hp.ev_begin);
  if (h->hp.ev_end) cudaEventDestroy(h->hp.ev_end);
  for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
  for (auto &m : h->tl) cudaEventDestroy(m.ev);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  if (h->ext_stream) cudaStreamDestroy(h->ext_stream);
  if (h->ev_ext) cudaEventDestroy(h->ev_ext);
  if (h->ev_int) cudaEventDestroy(h->ev_int);
  if (h->ev_extk) cudaEventDestroy(h->ev_extk);
  if (h->ev_gextk) cudaEventDestroy(h->ev_gextk);
  if (h->ev_gint) cudaEventDestroy(h->ev_gint);
  if (h->ev_hb_fext) cudaEventDestroy(h->ev_hb_fext);
  if (h->ev_hb_fint) cudaEventDestroy(h->ev_hb_fint);
  if (h->ev_ready) cudaEventDestroy(h->ev_ready);
  if (h->ev_done) cudaEventDestroy(h->ev_done);
  delete h;
  return CMDG_OK;
}

int cmdg_bind_grid(cmdg_handle h, const void *vgeo, const void *sgeo, const int64_t *vmapM,
                   const int64_t *vmapP, const int64_t *elemtobndy, const void *D,
                   const int64_t *interiorelems, int64_t ninterior,
                   const int64_t *exteriorelems, int64_t nexterior, const int64_t *vmapsend,
                   int64_t nvmapsend, const int64_t *vmaprecv, int64_t nvmaprecv,
                   const int32_t *nabrtorank, const int64_t *nabrtovmapsend,
                   const int64_t *nabrtovmaprecv, int32_t nnabr) {
  if (!h) return CMDG_ERR_INVALID;
  if (h->grid_bound) return fail(h, CMDG_ERR_INVALID, "grid already bound");
  if (h->is_hb && !h->ocean_set)
    return fail(h, CMDG_ERR_INVALID, "call cmdg_set_ocean_model before cmdg_bind_grid");
  if (!vgeo || !sgeo || !vmapM || !vmapP || !elemtobndy || !D)
    return fail(h, CMDG_ERR_INVALID, "null grid array");
  if (ninterior + nexterior != h->d.nrealelem)
    return fail(h, CMDG_ERR_INVALID, "interiorelems + exteriorelems must cover the real elements");
  int rc = DISPATCH_FT(h, bind_grid_t<double>(h, vgeo, sgeo, D), bind_grid_t<float>(h, vgeo, sgeo, D));
  if (rc) return rc;
  if ((rc = build_conn(h, vmapM, vmapP, elemtobndy))) return rc;
  if ((rc = to_zero_based_list(h, interiorelems, ninterior, &h->interior))) return rc;
  if ((rc = to_zero_based_list(h, exteriorelems, nexterior, &h->exterior))) return rc;
  h->ninterior = ninterior;
  h->nexterior = nexterior;
  if (nexterior > 0 && ninterior > 0) {
    // [exterior..., interior...] for the single-launch schedule + the counter its exterior blocks bump
    CU(cudaMalloc((void **)&h->all_list, (size_t)(nexterior + ninterior) * sizeof(int)));
    CU(cudaMemcpy(h->all_list, h->exterior, (size_t)nexterior * sizeof(int), cudaMemcpyDeviceToDevice));
    CU(cudaMemcpy(h->all_list + nexterior, h->interior, (size_t)ninterior * sizeof(int), cudaMemcpyDeviceToDevice));
    CU(cudaMalloc((void **)&h->ext_done, sizeof(unsigned)));
    CU(cudaMemset(h->ext_done, 0, sizeof(unsigned)));
  }
  if ((rc = to_zero_based_map(h, vmapsend, nvmapsend, &h->vmapsend0))) return rc;
  if ((rc = to_zero_based_map(h, vmaprecv, nvmaprecv, &h->vmaprecv0))) return rc;
  h->nvmapsend = nvmapsend;
  h->nvmaprecv = nvmaprecv;
  h->nabrtorank.assign(nabrtorank, nabrtorank + (nnabr > 0 ? nnabr : 0));
  for (int n = 0; n < nnabr; ++n) {
    h->sendrange.push_back(nabrtovmapsend[2 * n] - 1);
    h->sendrange.push_back(nabrtovmapsend[2 * n + 1]);
    h->recvrange.push_back(nabrtovmaprecv[2 * n] - 1);
    h->recvrange.push_back(nabrtovmaprecv[2 * n + 1]);
  }
  if (h->is_hb && nexterior > 0 && ninterior > 0 && h->d.nvertelem > 0) {
    // two-chain ocean schedule: stacks are whole on a rank, so a stack is interior or exterior as a whole; list the
    // first element of each (if the lists do not split by stacks the serial schedule is used)
    const int nv = h->d.nvertelem;
    auto stacks_of = [&](const int *dev_list, int64_t n, int **out) -> int {
      std::vector<int> v((size_t)n), firsts;
      CU(cudaMemcpy(v.data(), dev_list, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
      std::vector<char> in((size_t)h->d.nrealelem, 0);
      for (int e : v) in[(size_t)e] = 1;
      for (int e : v)
        if (e % nv == 0) {
          bool whole = e + nv <= (int)h->d.nrealelem;
          for (int k = 0; whole && k < nv; ++k) whole = in[(size_t)e + k] != 0;
          if (whole) firsts.push_back(e);
        }
      if ((int64_t)firsts.size() * nv != n) return 0;   // not whole stacks: *out stays NULL
      CU(cudaMalloc((void **)out, firsts.size() * sizeof(int)));
      CU(cudaMemcpy(*out, firsts.data(), firsts.size() * sizeof(int), cudaMemcpyHostToDevice));
      return 0;
    };
    if ((rc = stacks_of(h->interior, ninterior, &h->int_stacks))) return rc;
    if ((rc = stacks_of(h->exterior, nexterior, &h->ext_stacks))) return rc;
  }
  if (h->is_hb) {
    const size_t n = (size_t)h->d.nelem * h->Np;
    CU(cudaMalloc(&h->JcV, n * h->fb + 16));
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (h->d.float_bytes == CMDG_F64)
      extract_column_kernel<double><<<nb, 256>>>((double *)h->JcV, (const double *)vgeo, h->Np, 25, 15, h->d.nelem);
    else
      extract_column_kernel<float><<<nb, 256>>>((float *)h->JcV, (const float *)vgeo, h->Np, 25, 15, h->d.nelem);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    h->launches++;
  }
  h->grid_bound = true;
  return CMDG_OK;
}

int cmdg_set_ocean_model(cmdg_handle h, const cmdg_ocean_desc *o) {
  if (!h || !o) return CMDG_ERR_INVALID;
  if (!h->is_hb) return fail(h, CMDG_ERR_INVALID, "handle was not created with CMDG_MODEL_HB");
  if (o->struct_bytes != (int32_t)sizeof(cmdg_ocean_desc))
    return fail(h, CMDG_ERR_INVALID, "cmdg_ocean_desc size mismatch (ABI)");
  if (o->nbc < 0 || o->nbc > 6) return fail(h, CMDG_ERR_INVALID, "nbc must be 0..6");
  for (int i = 0; i < o->nbc; ++i) {
    if (o->bc_velocity[i] < CMDG_OCEAN_VEL_NOSLIP || o->bc_velocity[i] > CMDG_OCEAN_VEL_PENETRABLE_KINEMATIC_STRESS)
      return fail(h, CMDG_ERR_UNSUPPORTED, "unsupported ocean velocity boundary condition");
    if (o->bc_temperature[i] != CMDG_OCEAN_TEMP_INSULATING && o->bc_temperature[i] != CMDG_OCEAN_TEMP_FLUX)
      return fail(h, CMDG_ERR_UNSUPPORTED, "unsupported ocean temperature boundary condition");
  }
  h->od = *o;
  h->d.nbc = o->nbc;
  h->ocean_set = true;
  return CMDG_OK;
}

int cmdg_bind_ocean_operators(cmdg_handle h, const void *Fc, const void *Fe, const void *Imat) {
  if (!h || !Fc || !Fe || !Imat) return CMDG_ERR_INVALID;
  if (!h->is_hb) return fail(h, CMDG_ERR_INVALID, "handle was not created with CMDG_MODEL_HB");
  int rc;
  if (h->d.float_bytes == CMDG_F64) {
    if ((rc = upload_rowmajor<double>(h, &h->Fc, Fc, h->Nq))) return rc;
    if ((rc = upload_rowmajor<double>(h, &h->Fe, Fe, h->Nq))) return rc;
    if ((rc = upload_rowmajor<double>(h, &h->Imat, Imat, h->Nq))) return rc;
  } else {
    if ((rc = upload_rowmajor<float>(h, &h->Fc, Fc, h->Nq))) return rc;
    if ((rc = upload_rowmajor<float>(h, &h->Fe, Fe, h->Nq))) return rc;
    if ((rc = upload_rowmajor<float>(h, &h->Imat, Imat, h->Nq))) return rc;
  }
  return CMDG_OK;
}

int cmdg_bind_state(cmdg_handle h, void *aux, void *gradflux) {
  if (!h) return CMDG_ERR_INVALID;
  if (!aux) return fail(h, CMDG_ERR_INVALID, "state_auxiliary is null");
  if (h->visc && !gradflux)
    return fail(h, CMDG_ERR_INVALID, "state_gradient_flux is required unless skip_zero_viscosity applies");
  h->aux = aux;
  h->gradflux = gradflux;
  h->faux_dirty = true;
  if (h->visc && !h->is_hb && !h->F2dev) {
    // private outputs of the gradient kernel (zeroed: ghost entries are read before the first exchange
    // only on single-rank runs that have no ghosts)
    const size_t f2 = (size_t)h->d.nelem * 12 * h->Np * h->fb, fn = (size_t)h->d.nrealelem * 6 * h->Nfp * 4 * h->fb;
    CU(cudaMalloc(&h->F2dev, f2 + 16));
    CU(cudaMalloc(&h->FnDev, fn + 16));
    CU(cudaMemset(h->F2dev, 0, f2));
    CU(cudaMemset(h->FnDev, 0, fn));
    if (h->hyper) {
      const size_t hd = (size_t)h->d.nelem * 4 * h->Np * h->fb;
      CU(cudaMalloc(&h->Qhg, f2 + 16));
      CU(cudaMalloc(&h->Qhd, hd + 16));
      CU(cudaMemset(h->Qhg, 0, f2));
      CU(cudaMemset(h->Qhd, 0, hd));
    }
    if (h->ntracers) {
      const size_t nu = (size_t)h->d.nrealelem * 3 * h->Np * h->fb;
      const size_t fc = (size_t)h->d.nelem * 3 * h->ntracers * h->Np * h->fb;
      CU(cudaMalloc(&h->NuDev, nu + 16));
      CU(cudaMalloc(&h->F2chi, fc + 16));
      CU(cudaMemset(h->NuDev, 0, nu));
      CU(cudaMemset(h->F2chi, 0, fc));
    }
  }
  return CMDG_OK;
}

static int check_ready(cmdg_handle h) {
  if (!h) return CMDG_ERR_INVALID;
  if (!h->grid_bound || !h->aux) return fail(h, CMDG_ERR_INVALID, "bind the grid and the state first");
  if (h->is_hb && (!h->Fc || !h->Imat))
    return fail(h, CMDG_ERR_INVALID, "HBModel: call cmdg_bind_ocean_operators first");
  return 0;
}

int cmdg_tendency(cmdg_handle h, void *dQ, void *Q, double t, double alpha, double beta,
                  cmdg_stream stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!dQ || !Q) return fail(h, CMDG_ERR_INVALID, "null state array");
  cudaStream_t st = (cudaStream_t)stream;
  h->tev_used = 0;   // per-launch timing events are only evaluated by cmdg_lsrk_steps: reuse the pool here
  return DISPATCH_FT(h, tendency_t<double>(h, dQ, Q, t, alpha, beta, st),
                     tendency_t<float>(h, dQ, Q, t, alpha, beta, st));
}

int cmdg_lsrk_update(cmdg_handle h, void *dQ, void *Q, double rka, double rkb, double dt,
                     cmdg_stream stream) {
  if (!h) return CMDG_ERR_INVALID;
  if (!dQ || !Q) return fail(h, CMDG_ERR_INVALID, "null state array");
  cudaStream_t st = (cudaStream_t)stream;
  return DISPATCH_FT(h, lsrk_update_t<double>(h, dQ, Q, rka, rkb, dt, st),
                     lsrk_update_t<float>(h, dQ, Q, rka, rkb, dt, st));
}

int cmdg_lsrk_steps(cmdg_handle h, void *Q, void *dQ, double t0, double dt, int32_t nstage,
                    const double *rka, const double *rkb, const double *rkc, int64_t nsteps,
                    cmdg_stream stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!dQ || !Q || !rka || !rkb || !rkc || nstage <= 0 || nsteps < 0)
    return fail(h, CMDG_ERR_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  return DISPATCH_FT(h, lsrk_steps_t<double>(h, Q, dQ, t0, dt, nstage, rka, rkb, rkc, nsteps, st),
                     lsrk_steps_t<float>(h, Q, dQ, t0, dt, nstage, rka, rkb, rkc, nsteps, st));
}

// Chunk tables of the pipelined host path (once per handle): upload chunk of every element, the chunk that
// completes its stencil (itself and its face neighbours), launch list ordered by that chunk.
static int build_host_pipe(cmdg_handle h) {
  cmdg_handle_s::HostPipe &hp = h->hp;
  if (hp.nch) return 0;
  const int64_t nreal = h->d.nrealelem;
  int nch = 8;
  if (const char *v = getenv("CMDG_HOST_CHUNKS")) nch = std::max(1, std::min(64, atoi(v)));
  nch = (int)std::min<int64_t>(nch, nreal / 8);
  if ((int64_t)h->nbr_host.size() != nreal * 6) return fail(h, CMDG_ERR_INVALID, "host pipeline: no connectivity");
  std::vector<int64_t> first(nch + 1);
  for (int c = 0; c <= nch; ++c) first[c] = nreal * c / nch;
  std::vector<int> chunk_of((size_t)nreal), ready((size_t)nreal);
  for (int c = 0; c < nch; ++c)
    for (int64_t e = first[c]; e < first[c + 1]; ++e) chunk_of[(size_t)e] = c;
  std::vector<int64_t> off(nch + 1, 0);
  for (int64_t e = 0; e < nreal; ++e) {
    int r = chunk_of[(size_t)e];
    for (int f = 0; f < 6; ++f) {
      const int nb = h->nbr_host[(size_t)e * 6 + f];
      if (nb >= 0 && nb < nreal) r = std::max(r, chunk_of[(size_t)nb]);
    }
    ready[(size_t)e] = r;
    off[r + 1]++;
  }
  for (int c = 0; c < nch; ++c) off[c + 1] += off[c];
  std::vector<int> list((size_t)nreal), ident((size_t)nreal);
  std::vector<int64_t> pos(off.begin(), off.end() - 1);
  for (int64_t e = 0; e < nreal; ++e) {     // stable: the memory order is kept inside a chunk
    list[(size_t)pos[ready[(size_t)e]]++] = (int)e;
    ident[(size_t)e] = (int)e;
  }
  CU(cudaMalloc(&hp.ready_list, (size_t)nreal * sizeof(int)));
  CU(cudaMalloc(&hp.identity, (size_t)nreal * sizeof(int)));
  CU(cudaMemcpy(hp.ready_list, list.data(), (size_t)nreal * sizeof(int), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(hp.identity, ident.data(), (size_t)nreal * sizeof(int), cudaMemcpyHostToDevice));
  CU(cudaStreamCreateWithFlags(&hp.copy, cudaStreamNonBlocking));
  hp.ev_up.resize(nch);
  hp.ev_k.resize(nch);
  for (int c = 0; c < nch; ++c) {
    CU(cudaEventCreateWithFlags(&hp.ev_up[c], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&hp.ev_k[c], cudaEventDisableTiming));
  }
  CU(cudaEventCreateWithFlags(&hp.ev_begin, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&hp.ev_end, cudaEventDisableTiming));
  hp.first = first;
  hp.ready_off = off;
  hp.nch = nch;
  return 0;
}

int cmdg_lsrk_steps_host(cmdg_handle h, void *Q_host, double t0, double dt, int32_t nstage,
                         const double *rka, const double *rkb, const double *rkc,
                         int64_t nsteps) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!Q_host) return fail(h, CMDG_ERR_INVALID, "null host state");
  const size_t all = (size_t)h->d.nelem * h->d.nstate * h->Np * h->fb;
  const size_t real = (size_t)h->d.nrealelem * h->d.nstate * h->Np * h->fb;
  if (!h->Qdev) {
    CU(cudaMalloc(&h->Qdev, all));
    CU(cudaMalloc(&h->dQdev, all));
    CU(cudaMemset(h->Qdev, 0, all));
    CU(cudaMemset(h->dQdev, 0, all));
  }
  // Single rank, Euler path: upload, first stage, last stage and download are pipelined chunk by chunk
  // (CMDG_HOST_PIPE=0 restores copy -> steps -> copy).  The other paths need the whole state (ghost exchange,
  // gradient pass, column integrals) before their first kernel and copy it in one piece.
  static const bool pipe_on = !(getenv("CMDG_HOST_PIPE") && atoi(getenv("CMDG_HOST_PIPE")) == 0);
  const bool par = h->comm && !h->nabrtorank.empty();
  const bool piped = pipe_on && !par && !h->is_hb && !h->ntracers && !h->visc && h->step_filter_target < 0 &&
                     nstage > 0 && nsteps > 0 && (int64_t)nstage * nsteps >= 2 && rka && rkb && rkc &&
                     h->d.nrealelem >= 16 && h->d.nelem == h->d.nrealelem;
  if (piped) {
    if ((rc = build_host_pipe(h))) return rc;
    cmdg_handle_s::HostPipe &hp = h->hp;
    const size_t esz = (size_t)h->d.nstate * h->Np * h->fb;
    CU(cudaEventRecord(hp.ev_begin, 0));              // earlier work on the state buffers is finished first
    CU(cudaStreamWaitEvent(hp.copy, hp.ev_begin, 0));
    for (int c = 0; c < hp.nch; ++c) {
      const size_t o = (size_t)hp.first[c] * esz;
      CU(cudaMemcpyAsync((char *)h->Qdev + o, (const char *)Q_host + o, (size_t)(hp.first[c + 1] - hp.first[c]) * esz,
                         cudaMemcpyHostToDevice, hp.copy));
      CU(cudaEventRecord(hp.ev_up[c], hp.copy));
    }
    rc = DISPATCH_FT(h, lsrk_steps_t<double>(h, h->Qdev, h->dQdev, t0, dt, nstage, rka, rkb, rkc, nsteps, 0, Q_host),
                     lsrk_steps_t<float>(h, h->Qdev, h->dQdev, t0, dt, nstage, rka, rkb, rkc, nsteps, 0, Q_host));
    if (rc) return rc;
    CU(cudaStreamSynchronize(0));
    return CMDG_OK;
  }
  CU(cudaMemcpyAsync(h->Qdev, Q_host, real, cudaMemcpyHostToDevice, 0));
  rc = cmdg_lsrk_steps(h, h->Qdev, h->dQdev, t0, dt, nstage, rka, rkb, rkc, nsteps, nullptr);
  if (rc) return rc;
  CU(cudaMemcpyAsync(Q_host, h->Qdev, real, cudaMemcpyDeviceToHost, 0));
  CU(cudaStreamSynchronize(0));
  return CMDG_OK;
}

int cmdg_filter_apply(cmdg_handle h, void *Q, int32_t nstate, int32_t target, uint32_t state_mask,
                      const void *filter_h, const void *filter_v, int32_t direction,
                      cmdg_stream stream) {
  if (!h || !Q || !filter_h || !filter_v) return fail(h, CMDG_ERR_INVALID, "cmdg_filter_apply: null argument");
  if (!h->grid_bound) return fail(h, CMDG_ERR_INVALID, "cmdg_filter_apply before cmdg_bind_grid");
  if (target != CMDG_FILTER_INDICES && target != CMDG_FILTER_ATMOS_PERTURBATIONS)
    return fail(h, CMDG_ERR_UNSUPPORTED, "unsupported filter target");
  if (direction < CMDG_DIR_EVERY || direction > CMDG_DIR_VERTICAL)
    return fail(h, CMDG_ERR_INVALID, "bad filter direction");
  if (nstate < 1 || nstate > 5) return fail(h, CMDG_ERR_UNSUPPORTED, "filter: 1..5 states");
  // the matrices are read in the caller's Julia layout by the kernel itself: no scratch copy, no host
  // round trip, fully asynchronous on the caller's stream
  if (h->fb == 8)
    return filter_apply_t<double>(h, Q, nstate, target, state_mask, filter_h, filter_v, direction, (cudaStream_t)stream, 1);
  return filter_apply_t<float>(h, Q, nstate, target, state_mask, filter_h, filter_v, direction, (cudaStream_t)stream, 1);
}

int cmdg_set_step_filter(cmdg_handle h, int32_t target, uint32_t state_mask, const void *filter_h,
                         const void *filter_v, int32_t direction) {
  if (!h) return fail(h, CMDG_ERR_INVALID, "null handle");
  if (target < 0) {
    h->step_filter_target = -1;
    return CMDG_OK;
  }
  if (h->is_hb) return fail(h, CMDG_ERR_UNSUPPORTED, "per-step filter: AtmosModel handles only");
  if (!h->grid_bound || !filter_h || !filter_v) return fail(h, CMDG_ERR_INVALID, "cmdg_set_step_filter: bind the grid first and pass both matrices");
  if (target != CMDG_FILTER_INDICES && target != CMDG_FILTER_ATMOS_PERTURBATIONS)
    return fail(h, CMDG_ERR_UNSUPPORTED, "unsupported filter target");
  if (target == CMDG_FILTER_ATMOS_PERTURBATIONS && h->d.ref_state != CMDG_REF_HYDROSTATIC)
    return fail(h, CMDG_ERR_INVALID, "AtmosFilterPerturbations needs a HydrostaticState reference state");
  if (direction < CMDG_DIR_EVERY || direction > CMDG_DIR_VERTICAL)
    return fail(h, CMDG_ERR_INVALID, "bad filter direction");
  int rc;
  if (h->fb == 8) {
    if ((rc = transpose_small<double>(h, filter_h, h->Nq, &h->stepWh))) return rc;
    if ((rc = transpose_small<double>(h, filter_v, h->Nq, &h->stepWv))) return rc;
  } else {
    if ((rc = transpose_small<float>(h, filter_h, h->Nq, &h->stepWh))) return rc;
    if ((rc = transpose_small<float>(h, filter_v, h->Nq, &h->stepWv))) return rc;
  }
  h->step_filter_target = target;
  h->step_filter_mask = state_mask;
  h->step_filter_dir = direction;
  return CMDG_OK;
}

int cmdg_courant(cmdg_handle h, const void *Q, const void *vgeo, double dt, int32_t kind,
                 int32_t direction, double *result_host, cmdg_stream stream) {
  if (!h || !Q || !vgeo || !result_host) return fail(h, CMDG_ERR_INVALID, "cmdg_courant: null argument");
  if (h->is_hb) return fail(h, CMDG_ERR_UNSUPPORTED, "cmdg_courant: AtmosModel handles only");
  if (int rc = check_ready(h)) return rc;
  if (kind < CMDG_COURANT_ADVECTIVE || kind > CMDG_COURANT_DIFFUSIVE)
    return fail(h, CMDG_ERR_INVALID, "bad courant kind");
  if (direction < CMDG_DIR_EVERY || direction > CMDG_DIR_VERTICAL)
    return fail(h, CMDG_ERR_INVALID, "bad direction");
  if (kind == CMDG_COURANT_DIFFUSIVE && !h->gradflux)
    return fail(h, CMDG_ERR_INVALID, "diffusive courant needs the bound state_gradient_flux");
  cudaStream_t st = (cudaStream_t)stream;
  if (!h->courant_dev) CU(cudaMalloc(&h->courant_dev, sizeof(unsigned long long)));
  CU(cudaMemsetAsync(h->courant_dev, 0, sizeof(unsigned long long), st));
  const int64_t nreal = h->d.nrealelem;
  if (nreal > 0) {
    if (h->fb == 8) {
      const AtmosParams<double> P = make_params<double>(h);
      courant_kernel<double, 5><<<(unsigned)nreal, Dims<5>::BLOCK, 0, st>>>(
          (const double *)Q, (const double *)h->aux, (const double *)h->gradflux, (const double *)vgeo, P,
          dt, kind, direction, (unsigned long long *)h->courant_dev);
    } else {
      const AtmosParams<float> P = make_params<float>(h);
      courant_kernel<float, 5><<<(unsigned)nreal, Dims<5>::BLOCK, 0, st>>>(
          (const float *)Q, (const float *)h->aux, (const float *)h->gradflux, (const float *)vgeo, P,
          (float)dt, kind, direction, (unsigned long long *)h->courant_dev);
    }
    CU(cudaGetLastError());
    h->launches++;
  }
  unsigned long long bits = 0;
  CU(cudaMemcpyAsync(&bits, h->courant_dev, sizeof(bits), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  double v;
  memcpy(&v, &bits, sizeof(v));
  // typemin for a rank without real elements (SpaceDiscretization.jl:359-361)
  *result_host = nreal > 0 ? v : -INFINITY;
  return CMDG_OK;
}

int cmdg_comm_unique_id(void *id128) {
  std::string err;
  if (!id128) return CMDG_ERR_INVALID;
  if (!g_nccl.load(err)) return fail(nullptr, CMDG_ERR_NCCL, err);
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return fail(nullptr, CMDG_ERR_NCCL, g_nccl.GetErrorString(r));
  memcpy(id128, &id, 128);
  return CMDG_OK;
}

int cmdg_comm_init(cmdg_handle h, const void *id128, int32_t rank, int32_t nranks) {
  if (!h || !id128) return CMDG_ERR_INVALID;
  std::string err;
  if (!g_nccl.load(err)) return fail(h, CMDG_ERR_NCCL, err);
  if (h->comm) return fail(h, CMDG_ERR_INVALID, "communicator already initialised");
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NC(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
  h->rank = rank;
  h->nranks = nranks;
  // halo buffers for the widest array the library itself exchanges (Q, F2 / Qhg: 12, HB gradient flux: 10,
  // tracer fluxes) or the harness may (state_auxiliary): no cudaFree / cudaMalloc in the middle of a step
  if (!h->nabrtorank.empty()) {
    size_t ns = std::max<size_t>((size_t)h->d.nstate, (size_t)h->d.naux);
    ns = std::max<size_t>(ns, 12);
    ns = std::max<size_t>(ns, (size_t)(3 * h->d.ntracers));
    CU(cudaMalloc(&h->sendbuf, (size_t)h->nvmapsend * ns * h->fb + 16));
    CU(cudaMalloc(&h->recvbuf, (size_t)h->nvmaprecv * ns * h->fb + 16));
    h->commbuf_states = ns;
  }
  return CMDG_OK;
}

int cmdg_exchange_begin(cmdg_handle h, void *array, int32_t nstate, cmdg_stream stream) {
  if (!h || !array) return CMDG_ERR_INVALID;
  if (!h->grid_bound) return fail(h, CMDG_ERR_INVALID, "bind the grid first");
  cudaStream_t st = (cudaStream_t)stream;
  return DISPATCH_FT(h, exchange_begin_t<double>(h, array, nstate, st),
                     exchange_begin_t<float>(h, array, nstate, st));
}

int cmdg_exchange_end(cmdg_handle h, void *array, int32_t nstate, cmdg_stream stream) {
  if (!h || !array) return CMDG_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  return DISPATCH_FT(h, exchange_end_t<double>(h, array, nstate, st),
                     exchange_end_t<float>(h, array, nstate, st));
}

int cmdg_check_for_crashes(cmdg_handle h, const void *Q, int32_t nstate, int32_t *local_bad_host,
                           int32_t *any_bad_host, cmdg_stream stream) {
  if (!h || !Q || !local_bad_host || !any_bad_host || nstate <= 0)
    return fail(h, CMDG_ERR_INVALID, "cmdg_check_for_crashes: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (!h->crash_dev) CU(cudaMalloc((void **)&h->crash_dev, 2 * sizeof(int)));
  CU(cudaMemsetAsync(h->crash_dev, 0, 2 * sizeof(int), st));
  const size_t n = (size_t)h->d.nrealelem * nstate * h->Np;
  if (n) {
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
    if (h->fb == 8) nonfinite_kernel<double><<<blocks, 256, 0, st>>>((const double *)Q, n, h->crash_dev);
    else nonfinite_kernel<float><<<blocks, 256, 0, st>>>((const float *)Q, n, h->crash_dev);
    CU(cudaGetLastError());
    h->launches++;
  }
  if (h->comm) {
    // one communicator, one stream at a time: the reduction goes on the halo stream, ordered after the scan
    CU(cudaEventRecord(h->ev_ready, st));
    CU(cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0));
    NC(g_nccl.AllReduce(h->crash_dev, h->crash_dev + 1, 1, ncclInt32, ncclMax, h->comm, h->comm_stream));
    CU(cudaEventRecord(h->ev_done, h->comm_stream));
    CU(cudaStreamWaitEvent(st, h->ev_done, 0));
  } else {
    CU(cudaMemcpyAsync(h->crash_dev + 1, h->crash_dev, sizeof(int), cudaMemcpyDeviceToDevice, st));
  }
  int out[2] = {0, 0};
  CU(cudaMemcpyAsync(out, h->crash_dev, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  *local_bad_host = out[0];
  *any_bad_host = out[1];
  return CMDG_OK;
}

int cmdg_sync(cmdg_handle h) {
  if (!h) return CMDG_ERR_INVALID;
  CU(cudaStreamSynchronize(h->comm_stream));
  CU(cudaDeviceSynchronize());
  return CMDG_OK;
}

int64_t cmdg_kernel_launches(cmdg_handle h) { return h ? h->launches : -1; }

int cmdg_set_timing(cmdg_handle h, int32_t enable) {
  if (!h) return CMDG_ERR_INVALID;
  h->timing = enable != 0;
  h->last_ms = -1;
  h->last_nl = 0;
  return CMDG_OK;
}

double cmdg_last_kernel_ms(cmdg_handle h, int64_t *nl) {
  if (!h) return -1;
  if (nl) *nl = h->last_nl;
  return h->last_ms;
}

double cmdg_kernel_class_ms(cmdg_handle h, int32_t kclass, int64_t *nl) {
  if (!h || kclass < 0 || kclass >= CMDG_KCLASS_COUNT || h->last_ms < 0) return -1;
  if (nl) *nl = h->class_n[kclass];
  return h->class_ms[kclass];
}

}  // extern "C"
