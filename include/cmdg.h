/*
 * cmdg.h -- C ABI of libcmdg, the B200-native (sm_100a) DG tendency + LSRK library.
 *
 * This is the drop-in boundary for ONE path of CliMA/ClimateMachine.jl: the DGModel
 * tendency evaluation and the low-storage Runge-Kutta stage update.  Every entry point
 * below names the reference interface it replaces (paths relative to the reference
 * repository root).  A Julia `B200DGModel <: SpaceDiscretization` binds these with
 * `ccall`; the Python harness in climatemachine.jl_b200/ binds the
 * same symbols with ctypes.
 *
 * Conventions
 *  - plain C: pointers + sizes, no C++/torch types; all functions return 0 on success,
 *    a negative cmdg_status otherwise; cmdg_last_error() gives the message.
 *  - every array pointer is a DEVICE pointer in the reference's own layout unless the
 *    name ends in _host.  Julia arrays are column major, so `Q.data[n, s, e]` is at
 *    offset ((e-1)*nstate + (s-1))*Np + (n-1); index arrays hold the reference's
 *    1-based Int64 values (vmap = Np*(e-1)+n, src/Numerics/Mesh/Grids.jl:623-630).
 *  - arrays stay owned by the caller; the library keeps private packed copies of the
 *    static geometry/connectivity and never frees or reallocates caller memory.
 *  - one handle per (process, GPU); a handle is not thread safe.
 *  - calls taking a stream are asynchronous with respect to the host.
 *  - unsupported model/flux/polynomial-order combinations are an error
 *    (CMDG_ERR_UNSUPPORTED): there is no CPU fallback.
 */
#ifndef CMDG_H
#define CMDG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMDG_VERSION 100 /* 0.1.0 */

typedef struct cmdg_handle_s *cmdg_handle;
typedef void *cmdg_stream; /* cudaStream_t */

typedef enum {
  CMDG_OK = 0,
  CMDG_ERR_INVALID = -1,     /* bad argument / call order              */
  CMDG_ERR_UNSUPPORTED = -2, /* model or option outside the supported set */
  CMDG_ERR_CUDA = -3,        /* CUDA runtime error                     */
  CMDG_ERR_NCCL = -4,        /* NCCL error or libnccl not loadable     */
  CMDG_ERR_NODEVICE = -5     /* no CUDA device                         */
} cmdg_status;

enum { CMDG_F32 = 4, CMDG_F64 = 8 };
/* balance laws (src/Atmos/Model/AtmosModel.jl, src/Ocean/HydrostaticBoussinesq) */
enum { CMDG_MODEL_ATMOS_DRY = 1, CMDG_MODEL_HB = 2 };
/* src/Numerics/DGMethods/NumericalFluxes.jl:219-340; src/Atmos/Model/AtmosModel.jl:967-1078 */
enum { CMDG_NF_RUSANOV = 0, CMDG_NF_CENTRAL = 1, CMDG_NF_ROE = 2 };
/* src/Common/Orientations/Orientations.jl */
enum { CMDG_ORIENT_NONE = 0, CMDG_ORIENT_FLAT = 1, CMDG_ORIENT_SPHERICAL = 2 };
/* src/Atmos/Model/ref_state.jl:22-64 */
enum { CMDG_REF_NONE = 0, CMDG_REF_HYDROSTATIC = 1 };
/* src/Common/TurbulenceClosures/TurbulenceClosures.jl:287-499 */
enum { CMDG_TURB_CONSTANT_KINEMATIC = 0, CMDG_TURB_CONSTANT_DYNAMIC = 1, CMDG_TURB_SMAGORINSKY = 2 };
/* src/Atmos/Model/tendencies_momentum.jl:62-92 (Gravity, Coriolis), :104-137 (RayleighSponge);
 * HeldSuarezForcing: experiments/AtmosGCM/heldsuarez.jl:112-172 (= tutorials/Atmos/heldsuarez.jl:45-118) */
enum { CMDG_SRC_GRAVITY = 1, CMDG_SRC_CORIOLIS = 2, CMDG_SRC_HELD_SUAREZ = 4, CMDG_SRC_RAYLEIGH_SPONGE = 8 };
/* src/Atmos/Model/bc_momentum.jl:1-80 with Insulating energy (bc_energy.jl:10-17) */
enum { CMDG_BC_FREESLIP = 1, CMDG_BC_NOSLIP = 2 };
/* hyperdiffusion model (src/Common/TurbulenceClosures/TurbulenceClosures.jl:793-848) */
enum { CMDG_HYPER_NONE = 0, CMDG_HYPER_DRY_BIHARMONIC = 1 };
/* DGModel.direction / diffusion_direction (src/Numerics/DGMethods/DGModel.jl:3-19) */
enum { CMDG_DIR_EVERY = 0, CMDG_DIR_HORIZONTAL = 1 };

/*
 * Everything the reference encodes in the types of
 * `DGModel(balance_law, grid, numerical_flux_first_order, numerical_flux_second_order,
 *          numerical_flux_gradient; direction, diffusion_direction)`
 * (src/Numerics/DGMethods/DGModel.jl:22-65) and of the AtmosModel it wraps.
 */
typedef struct {
  int32_t struct_bytes;   /* = sizeof(cmdg_desc), ABI check */
  int32_t float_bytes;    /* CMDG_F64 or CMDG_F32 (eltype(Q)) */
  int32_t dim;            /* 3 */
  int32_t N;              /* polynomial order (same in all directions) */
  int64_t nelem;          /* length(topology.elems): real + ghost */
  int64_t nrealelem;      /* length(topology.realelems) */
  int32_t nvertelem;      /* stack size, 0 if the topology is not stacked */
  int32_t model;          /* CMDG_MODEL_* */
  int32_t nf_first;       /* CMDG_NF_* */
  int32_t nf_second;      /* CMDG_NF_CENTRAL */
  int32_t nf_gradient;    /* CMDG_NF_CENTRAL */
  int32_t orientation;    /* CMDG_ORIENT_* */
  int32_t ref_state;      /* CMDG_REF_* */
  int32_t subtract_off;   /* HydrostaticState.subtract_off */
  int32_t turbulence;     /* CMDG_TURB_* */
  int32_t turb_with_divergence;
  double turb_param;      /* nu, rho*nu or C_smag */
  int32_t sources;        /* bit mask of CMDG_SRC_* */
  int32_t diffusion_direction; /* CMDG_DIR_* */
  int32_t skip_zero_viscosity; /* 1: skip the gradient pass when nu == 0 (results differ
                                  from the reference only in the sign of zeros; GF is
                                  then not refreshed) */
  int32_t write_aux_diagnostics; /* 1: keep aux.moisture.{theta_v,air_T} up to date as
                                    kernel_nodal_update_auxiliary_state! does */
  int32_t nbc;            /* number of boundary tags in use (<= 6) */
  int32_t bc_kind[6];     /* CMDG_BC_* for elemtobndy tag 1..nbc */
  /* state sizes, checked against the model (number_states(bl, ...)) */
  int32_t nstate, naux, ngrad, ngradflux;
  /* CLIMAParameters.Planet values */
  double R_d, cp_d, cv_d, T_0, MSLP, grav, Omega, inv_Pr_turb;
  double day;             /* CLIMAParameters.Planet.day (HeldSuarezForcing rates) */
  /* RayleighSponge{FT}(z_max, z_sponge, alpha_max, u_relaxation, gamma) */
  double sponge_z_max, sponge_z_sponge, sponge_alpha_max, sponge_gamma, sponge_u_relax[3];
  /* DryBiharmonic{FT}(tau_timescale): needs diffusion_direction = CMDG_DIR_HORIZONTAL (the
   * reference's 3-D EveryDirection kernel does not run, DGModel_kernels.jl:2640-2650) and adds
   * aux.hyperdiffusion.Delta after aux.turbulence.Delta; ngrad grows by 4 (u_h, h_tot) */
  int32_t hyperdiffusion; /* CMDG_HYPER_* */
  /* NTracers{N, FT}(delta_chi) (src/Atmos/Model/tracers.jl:113-131): N <= CMDG_MAX_TRACERS passive
   * tracers rho*chi after rho*e (nstate = 5 + N), aux.tracers.delta_chi after aux.moisture (naux += N),
   * chi in the gradient variables (ngrad += N), grad chi (3 x N, column-major) at the end of the
   * gradient flux (ngradflux += 3 N).  Rusanov / Central first-order fluxes; not with DryBiharmonic. */
  int32_t ntracers;
  double hyper_tau;
  double tracer_delta_chi[4];
} cmdg_desc;
#define CMDG_MAX_TRACERS 4

/*
 * Ocean HydrostaticBoussinesqModel (src/Ocean/HydrostaticBoussinesq/hydrostatic_boussinesq_model.jl:40-103)
 * with an AbstractSimpleBoxProblem of the OceanGyre family (src/Ocean/OceanProblems/ocean_gyre.jl):
 * model constants, the surface forcing used by the flux-based boundary conditions, and one
 * OceanBC (src/Ocean/OceanBC.jl) per boundary tag.  Used with cmdg_desc.model = CMDG_MODEL_HB
 * (nstate 4, naux 8, ngrad 5, ngradflux 10; the Atmos-only fields of cmdg_desc are ignored).
 */
enum { CMDG_OCEAN_VEL_NOSLIP = 1, CMDG_OCEAN_VEL_FREESLIP = 2, CMDG_OCEAN_VEL_PENETRABLE_FREESLIP = 3,
       CMDG_OCEAN_VEL_PENETRABLE_KINEMATIC_STRESS = 4 };
enum { CMDG_OCEAN_TEMP_INSULATING = 1, CMDG_OCEAN_TEMP_FLUX = 2 };
typedef struct {
  int32_t struct_bytes;
  int32_t nbc;
  int32_t bc_velocity[6];     /* CMDG_OCEAN_VEL_*  for elemtobndy tag 1..nbc */
  int32_t bc_temperature[6];  /* CMDG_OCEAN_TEMP_* */
  double grav, rho0, ch, cz, alphaT, nuh, nuz, kappah, kappaz, kappac, f0, beta;
  double Lx, Ly, H, tau0, lambda_r, thetaE;
} cmdg_ocean_desc;

/* library version (CMDG_VERSION) */
int cmdg_version(void);

/* message of the last error on this handle (or of the last failed cmdg_create if h == NULL) */
const char *cmdg_last_error(cmdg_handle h);

/*
 * Replaces the DGModel constructor, src/Numerics/DGMethods/DGModel.jl:22-65.
 * Validates the model against the supported set; selects the device kernels.
 */
int cmdg_create(const cmdg_desc *desc, cmdg_handle *out);

/* Destroys the handle and its private device buffers (never caller arrays). */
int cmdg_destroy(cmdg_handle h);

/*
 * Binds the grid arrays of `DiscontinuousSpectralElementGrid`
 * (src/Numerics/Mesh/Grids.jl:170-265): vgeo (Np x 25 x nelem), sgeo (5 x Nfp x 6 x nelem),
 * vmap-/vmap+ (Nfp x 6 x nelem, Int64), elemtobndy (6 x nelem, Int64), D (Nq x Nq),
 * topology.interiorelems / exteriorelems (src/Numerics/Mesh/Topologies.jl:251-252),
 * vmapsend / vmaprecv and the per-neighbour ranges nabrtovmapsend / nabrtovmaprecv
 * (Grids.jl:761-811; ranges given as 1-based inclusive [first,last] pairs) and
 * topology.nabrtorank.  Builds the packed private geometry (10 of 25 vgeo columns,
 * 4 of 5 sgeo rows) and the per-face neighbour table derived from vmap+.
 * nabr* arrays are HOST pointers (they are host arrays in the reference too).
 */
int cmdg_bind_grid(cmdg_handle h, const void *vgeo, const void *sgeo, const int64_t *vmapM,
                   const int64_t *vmapP, const int64_t *elemtobndy, const void *D,
                   const int64_t *interiorelems, int64_t ninterior,
                   const int64_t *exteriorelems, int64_t nexterior, const int64_t *vmapsend,
                   int64_t nvmapsend, const int64_t *vmaprecv, int64_t nvmaprecv,
                   const int32_t *nabrtorank_host, const int64_t *nabrtovmapsend_host,
                   const int64_t *nabrtovmaprecv_host, int32_t nnabr);

/*
 * HBModel only.  cmdg_set_ocean_model: the model/problem constants (call before the first
 * tendency).  cmdg_bind_ocean_operators: the two vertical filter matrices that
 * update_auxiliary_state! applies in every evaluation (hydrostatic_boussinesq_model.jl:637-663;
 * `modeldata.vert_filter.filter_matrices[end]`, `modeldata.exp_filter.filter_matrices[end]`,
 * Nq x Nq, Julia layout) and grid.Imat[end] used by the stack integrals
 * (src/Numerics/Mesh/Grids.jl:1184-1206), all device pointers.
 */
int cmdg_set_ocean_model(cmdg_handle h, const cmdg_ocean_desc *ocean);
int cmdg_bind_ocean_operators(cmdg_handle h, const void *vert_filter_matrix,
                              const void *exp_filter_matrix, const void *Imat);

/*
 * Binds dg.state_auxiliary.data (Np x naux x nelem) and dg.state_gradient_flux.data
 * (Np x ngradflux x nelem) -- fields of DGModel, src/Numerics/DGMethods/DGModel.jl:3-19.
 * gradflux may be NULL when ngradflux == 0 or skip_zero_viscosity is set.
 */
int cmdg_bind_state(cmdg_handle h, void *state_auxiliary, void *state_gradient_flux);

/*
 * Replaces `(dg::DGModel)(tendency, Q, param, t, alpha, beta)`,
 * src/Numerics/DGMethods/DGModel.jl:85-427 (and the 4-argument `increment` form,
 * SpaceDiscretization.jl:68-77, as alpha = 1, beta = increment):
 *     tendency = alpha * RHS(Q, t) + beta * tendency      on real elements,
 * including the halo exchange of Q (and of the gradient flux) when a communicator was
 * initialised, ordered as the reference orders it: begin exchange -> interior elements
 * -> end exchange -> exterior elements.  Q's ghost elements are updated in place.
 */
int cmdg_tendency(cmdg_handle h, void *tendency, void *Q, double t, double alpha, double beta,
                  cmdg_stream stream);

/*
 * Replaces the `update!` kernel of LowStorageRungeKutta2N,
 * src/Numerics/ODESolvers/LowStorageRungeKuttaMethod.jl:146-158, on realview(Q):
 *     Q += rkb * dt * dQ ;  dQ *= rka
 */
int cmdg_lsrk_update(cmdg_handle h, void *dQ, void *Q, double rka, double rkb, double dt,
                     cmdg_stream stream);

/*
 * Replaces `dostep!(Q, lsrk::LowStorageRungeKutta2N, p, time)`,
 * src/Numerics/ODESolvers/LowStorageRungeKuttaMethod.jl:102-144, for `nsteps`
 * consecutive steps of size dt starting at time t0, with the tendency evaluation and the
 * stage update fused into one kernel per stage (Q ping-pongs with a private buffer;
 * the result is left in Q, dQ holds the reference's scaled residual).  rka/rkb/rkc are
 * HOST arrays of length nstage (RKA, RKB, RKC of the tableau, :293-327 / :349-410).
 */
int cmdg_lsrk_steps(cmdg_handle h, void *Q, void *dQ, double t0, double dt, int32_t nstage,
                    const double *rka_host, const double *rkb_host, const double *rkc_host,
                    int64_t nsteps, cmdg_stream stream);

/*
 * Same as cmdg_lsrk_steps but through HOST buffers: copies realview(Q) (Np*nstate*nrealelem
 * values) from Q_host to the device, runs nsteps steps, copies the result back and
 * synchronises.  This is the call the end-to-end benchmark times.  Q_host should be pinned
 * (cudaHostAlloc / cudaHostRegister); pageable memory works but serialises the copies.  On one rank
 * the Euler path overlaps the upload with the first stage and the download with the last one
 * (chunked copies on a private stream, bit-identical results; CMDG_HOST_PIPE=0 turns it off).
 */
int cmdg_lsrk_steps_host(cmdg_handle h, void *Q_host, double t0, double dt, int32_t nstage,
                         const double *rka_host, const double *rkb_host,
                         const double *rkc_host, int64_t nsteps);

/*
 * Replaces `Filters.apply!(Q, target, grid, filter; state_auxiliary, direction)`,
 * src/Numerics/Mesh/Filters.jl:440-505 (kernel_apply_filter!, :651-792), on real elements, as one
 * pass over Q.  target: CMDG_FILTER_INDICES = FilterIndices(I...) with `state_mask` bit s-1 set for
 * every filtered state s; CMDG_FILTER_ATMOS_PERTURBATIONS = AtmosFilterPerturbations(atmos)
 * (src/Atmos/Model/filters.jl:4-48; needs a HydrostaticState reference state in the bound
 * state_auxiliary).  filter_h / filter_v: `filter.filter_matrices[1]` / `[end]`, Nq x Nq device
 * arrays in Julia layout.  direction: CMDG_DIR_EVERY, CMDG_DIR_HORIZONTAL or CMDG_DIR_VERTICAL.
 */
enum { CMDG_FILTER_INDICES = 0, CMDG_FILTER_ATMOS_PERTURBATIONS = 1 };
enum { CMDG_DIR_VERTICAL = 2 };
int cmdg_filter_apply(cmdg_handle h, void *Q, int32_t nstate, int32_t target, uint32_t state_mask,
                      const void *filter_h, const void *filter_v, int32_t direction,
                      cmdg_stream stream);
/*
 * Registers a filter that cmdg_lsrk_steps / cmdg_lsrk_steps_host apply to Q after every completed
 * step -- the per-step `cbfilter` callback of the GCM drivers
 * (experiments/TestCase/baroclinic_wave.jl:265-277, tutorials/Atmos/heldsuarez.jl:256-268:
 * GenericCallbacks.EveryXSimulationSteps(1) do Filters.apply!(Q, AtmosFilterPerturbations(model),
 * grid, ExponentialFilter(grid, 0, order); state_auxiliary) end).  target < 0 removes it.  The
 * matrices are copied.
 */
int cmdg_set_step_filter(cmdg_handle h, int32_t target, uint32_t state_mask, const void *filter_h,
                         const void *filter_v, int32_t direction);

/*
 * Replaces `courant(local_courant, dg, m, Q, dt, simtime, direction)`,
 * src/Numerics/DGMethods/SpaceDiscretization.jl:307-365, for the AtmosModel pointwise numbers of
 * src/Atmos/Model/courant.jl:12-86 (kind = CMDG_COURANT_*): node distances
 * (kernel_min_neighbor_distance!, src/Numerics/Mesh/Grids.jl:1219-1336), the pointwise number and
 * the maximum over the rank's real elements in one kernel.  Synchronous; *result_host receives the
 * rank-local maximum (the caller applies MPI.Allreduce(max) as the reference does).  `vgeo` is
 * grid.vgeo (Np x 25 x nelem, device).  CMDG_COURANT_DIFFUSIVE reads the bound
 * state_gradient_flux of the last tendency evaluation.
 */
enum { CMDG_COURANT_ADVECTIVE = 0, CMDG_COURANT_NONDIFFUSIVE = 1, CMDG_COURANT_DIFFUSIVE = 2 };
int cmdg_courant(cmdg_handle h, const void *Q, const void *vgeo, double dt, int32_t kind,
                 int32_t direction, double *result_host, cmdg_stream stream);

/*
 * Halo exchange of MPIStateArray face data, src/Arrays/MPIStateArrays.jl:411-514:
 * cmdg_comm_unique_id fills a 128-byte ncclUniqueId on one rank; the caller broadcasts it
 * (MPI in Julia, torch.distributed in the Python harness); cmdg_comm_init joins the
 * communicator.  cmdg_exchange_begin = begin_ghost_exchange! (pack kernel
 * kernel_fillsendbuf! + grouped ncclSend/ncclRecv on the library's communication stream),
 * cmdg_exchange_end = end_ghost_exchange! (kernel_transferrecvbuf!, ordered after the
 * receive; the given stream waits on it).  nstate = size(Q.data, 2).
 */
int cmdg_comm_unique_id(void *id128_host);
int cmdg_comm_init(cmdg_handle h, const void *id128_host, int32_t rank, int32_t nranks);
int cmdg_exchange_begin(cmdg_handle h, void *array, int32_t nstate, cmdg_stream stream);
int cmdg_exchange_end(cmdg_handle h, void *array, int32_t nstate, cmdg_stream stream);

/*
 * Replaces the crash check of the reference's waits (`checked_wait` with `check_for_crashes`,
 * src/Arrays/MPIStateArrays.jl:910-935, which raises `ErrorOnRemoteNode` on the ranks that did not fail
 * themselves): counts the non-finite values in realview(Q) (Np x nstate x nrealelem) on this rank and, when a
 * communicator was initialised, reduces the flag over all ranks (ncclAllReduce, max), so that EVERY rank
 * learns that some rank holds NaN/Inf and can stop collectively instead of hanging in the next halo exchange.
 * Synchronous.  *local_bad_host = 1 if this rank's state is not finite, *any_bad_host = 1 if any rank's is.
 */
int cmdg_check_for_crashes(cmdg_handle h, const void *Q, int32_t nstate, int32_t *local_bad_host,
                           int32_t *any_bad_host, cmdg_stream stream);

/* Blocks until all work queued by this handle has finished (checked_wait, DGModel.jl:426). */
int cmdg_sync(cmdg_handle h);

/* Number of kernels this handle has launched so far (evidence for benchmarks). */
int64_t cmdg_kernel_launches(cmdg_handle h);

/* Device time in ms of the tendency kernels launched by the last cmdg_lsrk_steps* call
 * (CUDA events on the launching stream); returns <0 if timing was not enabled. */
int cmdg_set_timing(cmdg_handle h, int32_t enable);
double cmdg_last_kernel_ms(cmdg_handle h, int64_t *nlaunches_out);
/* The same split by kernel class: tendency (dg_tendency_kernel / hb_tendency_kernel), gradient pass
 * (dg_gradient_kernel / hb_gradient_kernel), the two DryBiharmonic passes, the two passive-tracer kernels,
 * and the HBModel's vertical filter and stack-integral (column) kernels. */
enum { CMDG_KCLASS_TENDENCY = 0, CMDG_KCLASS_GRADIENT = 1, CMDG_KCLASS_HYPER_DIVERGENCE = 2,
       CMDG_KCLASS_HYPER_FLUX = 3, CMDG_KCLASS_TRACER_GRADIENT = 4, CMDG_KCLASS_TRACER_TENDENCY = 5,
       CMDG_KCLASS_HB_FILTER = 6, CMDG_KCLASS_HB_COLUMN = 7, CMDG_KCLASS_COUNT = 8 };
double cmdg_kernel_class_ms(cmdg_handle h, int32_t kclass, int64_t *nlaunches_out);

#ifdef __cplusplus
}
#endif
#endif /* CMDG_H */
