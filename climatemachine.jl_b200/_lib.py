"""ctypes binding of libcmdg (include/cmdg.h).  Fails loudly if the CUDA library is missing:
there is no CPU or PyTorch fallback for the DG tendency path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CMDG_LIB", os.path.join(_HERE, "libcmdg.so"))  # CMDG_LIB: dev override

CMDG_F32, CMDG_F64 = 4, 8
MODEL_ATMOS_DRY, MODEL_HB = 1, 2
NF_RUSANOV, NF_CENTRAL, NF_ROE = 0, 1, 2
ORIENT_NONE, ORIENT_FLAT, ORIENT_SPHERICAL = 0, 1, 2
REF_NONE, REF_HYDROSTATIC = 0, 1
TURB_CONSTANT_KINEMATIC, TURB_CONSTANT_DYNAMIC, TURB_SMAGORINSKY = 0, 1, 2
SRC_GRAVITY, SRC_CORIOLIS, SRC_HELD_SUAREZ, SRC_RAYLEIGH_SPONGE = 1, 2, 4, 8
BC_FREESLIP, BC_NOSLIP = 1, 2
DIR_EVERY, DIR_HORIZONTAL, DIR_VERTICAL = 0, 1, 2
HYPER_NONE, HYPER_DRY_BIHARMONIC = 0, 1
FILTER_INDICES, FILTER_ATMOS_PERTURBATIONS = 0, 1
COURANT_ADVECTIVE, COURANT_NONDIFFUSIVE, COURANT_DIFFUSIVE = 0, 1, 2

ERRORS = {-1: "CMDG_ERR_INVALID", -2: "CMDG_ERR_UNSUPPORTED", -3: "CMDG_ERR_CUDA",
          -4: "CMDG_ERR_NCCL", -5: "CMDG_ERR_NODEVICE"}


class cmdg_desc(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_int32), ("float_bytes", C.c_int32), ("dim", C.c_int32),
        ("N", C.c_int32), ("nelem", C.c_int64), ("nrealelem", C.c_int64),
        ("nvertelem", C.c_int32), ("model", C.c_int32), ("nf_first", C.c_int32),
        ("nf_second", C.c_int32), ("nf_gradient", C.c_int32), ("orientation", C.c_int32),
        ("ref_state", C.c_int32), ("subtract_off", C.c_int32), ("turbulence", C.c_int32),
        ("turb_with_divergence", C.c_int32), ("turb_param", C.c_double),
        ("sources", C.c_int32), ("diffusion_direction", C.c_int32),
        ("skip_zero_viscosity", C.c_int32), ("write_aux_diagnostics", C.c_int32),
        ("nbc", C.c_int32), ("bc_kind", C.c_int32 * 6),
        ("nstate", C.c_int32), ("naux", C.c_int32), ("ngrad", C.c_int32),
        ("ngradflux", C.c_int32),
        ("R_d", C.c_double), ("cp_d", C.c_double), ("cv_d", C.c_double), ("T_0", C.c_double),
        ("MSLP", C.c_double), ("grav", C.c_double), ("Omega", C.c_double),
        ("inv_Pr_turb", C.c_double), ("day", C.c_double),
        ("sponge_z_max", C.c_double), ("sponge_z_sponge", C.c_double),
        ("sponge_alpha_max", C.c_double), ("sponge_gamma", C.c_double),
        ("sponge_u_relax", C.c_double * 3),
        ("hyperdiffusion", C.c_int32), ("ntracers", C.c_int32), ("hyper_tau", C.c_double),
        ("tracer_delta_chi", C.c_double * 4),
    ]


class cmdg_ocean_desc(C.Structure):
    _fields_ = [("struct_bytes", C.c_int32), ("nbc", C.c_int32), ("bc_velocity", C.c_int32 * 6),
                ("bc_temperature", C.c_int32 * 6)] + \
               [(n, C.c_double) for n in ("grav", "rho0", "ch", "cz", "alphaT", "nuh", "nuz", "kappah",
                                          "kappaz", "kappac", "f0", "beta", "Lx", "Ly", "H", "tau0",
                                          "lambda_r", "thetaE")]


OCEAN_VEL_NOSLIP, OCEAN_VEL_FREESLIP, OCEAN_VEL_PENETRABLE_FREESLIP, OCEAN_VEL_PENETRABLE_KINEMATIC_STRESS = 1, 2, 3, 4
OCEAN_TEMP_INSULATING, OCEAN_TEMP_FLUX = 1, 2


class CmdgError(RuntimeError):
    pass


_lib = None

# every symbol include/cmdg.h declares
SYMBOLS = [
    "cmdg_version", "cmdg_last_error", "cmdg_create", "cmdg_destroy", "cmdg_bind_grid",
    "cmdg_bind_state", "cmdg_tendency", "cmdg_lsrk_update", "cmdg_lsrk_steps",
    "cmdg_lsrk_steps_host", "cmdg_comm_unique_id", "cmdg_comm_init", "cmdg_exchange_begin",
    "cmdg_exchange_end", "cmdg_sync", "cmdg_kernel_launches", "cmdg_set_timing",
    "cmdg_last_kernel_ms", "cmdg_kernel_class_ms", "cmdg_set_ocean_model", "cmdg_bind_ocean_operators",
    "cmdg_filter_apply", "cmdg_set_step_filter", "cmdg_courant", "cmdg_check_for_crashes",
]


def lib():
    """Load libcmdg.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CmdgError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'`.  The DG tendency path has no CPU/PyTorch fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.cmdg_version.restype = C.c_int
    L.cmdg_last_error.restype = C.c_char_p
    L.cmdg_last_error.argtypes = [vp]
    L.cmdg_create.argtypes = [C.POINTER(cmdg_desc), C.POINTER(vp)]
    L.cmdg_destroy.argtypes = [vp]
    L.cmdg_bind_grid.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, i64, vp, i64, vp, i64,
                                 C.POINTER(i32), C.POINTER(i64), C.POINTER(i64), i32]
    L.cmdg_bind_state.argtypes = [vp, vp, vp]
    L.cmdg_set_ocean_model.argtypes = [vp, C.POINTER(cmdg_ocean_desc)]
    L.cmdg_bind_ocean_operators.argtypes = [vp, vp, vp, vp]
    L.cmdg_tendency.argtypes = [vp, vp, vp, dbl, dbl, dbl, vp]
    L.cmdg_lsrk_update.argtypes = [vp, vp, vp, dbl, dbl, dbl, vp]
    L.cmdg_lsrk_steps.argtypes = [vp, vp, vp, dbl, dbl, i32, C.POINTER(dbl), C.POINTER(dbl),
                                  C.POINTER(dbl), i64, vp]
    L.cmdg_lsrk_steps_host.argtypes = [vp, vp, dbl, dbl, i32, C.POINTER(dbl), C.POINTER(dbl),
                                       C.POINTER(dbl), i64]
    L.cmdg_filter_apply.argtypes = [vp, vp, i32, i32, C.c_uint32, vp, vp, i32, vp]
    L.cmdg_set_step_filter.argtypes = [vp, i32, C.c_uint32, vp, vp, i32]
    L.cmdg_courant.argtypes = [vp, vp, vp, dbl, i32, i32, C.POINTER(dbl), vp]
    L.cmdg_check_for_crashes.argtypes = [vp, vp, i32, C.POINTER(i32), C.POINTER(i32), vp]
    L.cmdg_comm_unique_id.argtypes = [vp]
    L.cmdg_comm_init.argtypes = [vp, vp, i32, i32]
    L.cmdg_exchange_begin.argtypes = [vp, vp, i32, vp]
    L.cmdg_exchange_end.argtypes = [vp, vp, i32, vp]
    L.cmdg_sync.argtypes = [vp]
    L.cmdg_kernel_launches.argtypes = [vp]
    L.cmdg_kernel_launches.restype = i64
    L.cmdg_set_timing.argtypes = [vp, i32]
    L.cmdg_last_kernel_ms.argtypes = [vp, C.POINTER(i64)]
    L.cmdg_last_kernel_ms.restype = dbl
    L.cmdg_kernel_class_ms.argtypes = [vp, i32, C.POINTER(i64)]
    L.cmdg_kernel_class_ms.restype = dbl
    missing = [name for name in SYMBOLS if not hasattr(L, name)]
    if missing:
        raise CmdgError(f"{LIB_PATH} does not export {missing} (stale build?)")
    _lib = L
    return L


def check(rc, handle=None):
    if rc != 0:
        msg = lib().cmdg_last_error(handle)
        raise CmdgError(f"{ERRORS.get(rc, rc)}: {msg.decode() if msg else ''}")
