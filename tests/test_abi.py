"""CPU-side checks of the drop-in boundary: libcmdg.so loads, exports every symbol that
include/cmdg.h declares, validates descriptors, and never computes without a GPU."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as ge

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def P():
    return ge.build()


def test_header_symbols_all_exported(P):
    hdr = open(os.path.join(ROOT, "include", "cmdg.h")).read()
    # strip comments, then collect function declarations
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cmdg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = P._lib.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in cmdg.h but not exported"
    assert declared == set(P._lib.SYMBOLS)
    assert L.cmdg_version() == 100


def test_desc_struct_matches_header(P):
    # ABI guard: the library rejects a descriptor whose size differs from its own
    d = P._lib.cmdg_desc()
    d.struct_bytes = C.sizeof(P._lib.cmdg_desc) - 4
    h = C.c_void_p()
    rc = P._lib.lib().cmdg_create(C.byref(d), C.byref(h))
    assert rc == -1
    assert b"size mismatch" in P._lib.lib().cmdg_last_error(None)


def _desc(P, **kw):
    d = P._lib.cmdg_desc()
    d.struct_bytes = C.sizeof(P._lib.cmdg_desc)
    d.float_bytes, d.dim, d.N = 8, 3, 4
    d.nelem = d.nrealelem = 1
    d.model = P._lib.MODEL_ATMOS_DRY
    d.nf_second = d.nf_gradient = P._lib.NF_CENTRAL
    d.nstate, d.naux, d.ngrad, d.ngradflux = 5, 5, 4, 9
    for k, v in kw.items():
        setattr(d, k, v)
    return d


@pytest.mark.parametrize("kw,code,msg", [
    (dict(model=3), -2, b"unsupported balance law"),
    (dict(N=7), -2, b"polynomial order"),
    (dict(nf_first=5), -2, b"numerical flux"),
    (dict(nstate=9), -2, b"tracers"),
    (dict(sources=16), -2, b"source"),
    (dict(naux=16), -1, b"naux"),
    (dict(float_bytes=2), -2, b"Float64 or Float32"),
    (dict(ntracers=5, nstate=10), -2, b"NTracers"),
    (dict(ntracers=2, nstate=7, naux=7, ngrad=6, ngradflux=15, nf_first=2), -2, b"Rusanov / Central"),
    (dict(ntracers=2, nstate=7, naux=7, ngrad=6, ngradflux=14), -1, b"ngrad/ngradflux"),
])
def test_unsupported_models_raise_not_fallback(P, kw, code, msg):
    h = C.c_void_p()
    rc = P._lib.lib().cmdg_create(C.byref(_desc(P, **kw)), C.byref(h))
    assert rc == code
    assert msg in P._lib.lib().cmdg_last_error(None)
    assert not h.value


def test_no_device_is_an_error_not_a_fallback(P):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = P._lib.lib().cmdg_create(C.byref(_desc(P)), C.byref(h))
    assert rc == -5 and b"no CPU fallback" in P._lib.lib().cmdg_last_error(None)
    # a valid NTracers{2} descriptor passes validation and stops at the same place
    rc = P._lib.lib().cmdg_create(C.byref(_desc(P, ntracers=2, nstate=7, naux=7, ngrad=6, ngradflux=15)), C.byref(h))
    assert rc == -5


def test_host_mirror_rejects_unsupported_models(P):
    m = P.AtmosModel(tracers=object())
    with pytest.raises(P.UnsupportedModelError):
        m.validate()
    with pytest.raises(P.UnsupportedModelError):
        P.AtmosModel(tracers=P.NTracers((1, 2, 3, 4, 5))).validate()
    m = P.AtmosModel(orientation=P.FlatOrientation(), ref_state=P.HydrostaticState(P.DryAdiabaticProfile()),
                     turbulence=P.SmagorinskyLilly(), tracers=P.NTracers((1, 2, 3, 4)))
    m.validate()   # risingbubble.jl: S, A, G, GF = 9, 21, 9, 22
    assert [m.number_states(k) for k in ("Prognostic", "Auxiliary", "Gradient", "GradientFlux")] == [9, 21, 9, 22]
    with pytest.raises(P.UnsupportedModelError):
        P.AtmosModel(source=(object(),)).validate()
    assert P.AtmosModel().number_states("Auxiliary") == 5
    m = P.AtmosModel(orientation=P.SphericalOrientation(),
                     ref_state=P.HydrostaticState(P.DecayingTemperatureProfile()),
                     turbulence=P.SmagorinskyLilly())
    assert m.number_states("Auxiliary") == 17 and m.number_states("GradientFlux") == 10
