"""Mechanical consistency of the reference-side binding julia/B200DGModel.jl with the C ABI (Julia is not
installed in the build image, so the shim cannot be executed here): every `ccall` names a symbol that
include/cmdg.h declares and passes as many arguments as the C prototype takes; the two descriptor structs
list the same fields, in the same order and with the same widths, as the C structs (through their ctypes
mirror, whose layout tests/test_abi.py checks against the library); and the dispatch bug of the round-1
draft (a `dostep!` method on a type parameter that `LowStorageRungeKutta2N` does not have) is gone."""
import ctypes as C
import os
import re

import __graft_entry__ as ge

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "julia", "B200DGModel.jl"), encoding="utf-8").read()
HDR = open(os.path.join(ROOT, "include", "cmdg.h"), encoding="utf-8").read()


def _split_top(s):
    """Split on commas that are not nested in (), [] or {}."""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _c_prototypes():
    protos = {}
    text = re.sub(r"/\*.*?\*/", "", HDR, flags=re.S)
    for m in re.finditer(r"\b(?:int|int64_t|double|const char \*)\s*(cmdg_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(_split_top(args))
    return protos


def _ccalls():
    calls = []
    for m in re.finditer(r"ccall\(\(:(cmdg_\w+), libcmdg\),", SRC):
        # balanced scan of the ccall argument list
        i = SRC.index("(", m.start())
        depth, j = 0, i
        while True:
            if SRC[j] in "([{":
                depth += 1
            elif SRC[j] in ")]}":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        parts = _split_top(SRC[i + 1:j])
        # parts: (:sym, lib), rettype, (argtypes...), args...
        argtypes = _split_top(parts[2].strip()[1:-1])
        calls.append((m.group(1), len(argtypes), len(parts) - 3))
    return calls


def test_every_ccall_matches_a_declared_prototype():
    protos = _c_prototypes()
    pkg = ge.load_package()
    assert set(protos) == set(pkg._lib.SYMBOLS), set(protos) ^ set(pkg._lib.SYMBOLS)
    calls = _ccalls()
    assert len(calls) >= 14
    for name, ntypes, nargs in calls:
        assert name in protos, f"{name} is not declared in include/cmdg.h"
        assert ntypes == nargs, f"{name}: {ntypes} argument types but {nargs} arguments"
        assert ntypes == protos[name], f"{name}: ccall passes {ntypes} arguments, the C prototype takes {protos[name]}"
    used = {c[0] for c in calls}
    for must in ("cmdg_create", "cmdg_bind_grid", "cmdg_bind_state", "cmdg_tendency", "cmdg_lsrk_steps",
                 "cmdg_comm_unique_id", "cmdg_comm_init", "cmdg_filter_apply", "cmdg_set_step_filter",
                 "cmdg_courant", "cmdg_set_ocean_model", "cmdg_bind_ocean_operators", "cmdg_destroy",
                 "cmdg_exchange_begin", "cmdg_exchange_end"):
        assert must in used, must


_JL = {"Int32": (C.c_int32, 4), "Int64": (C.c_int64, 8), "Float64": (C.c_double, 8)}


def _julia_struct(name):
    body = re.search(r"struct %s\n(.*?)\nend" % name, SRC, flags=re.S).group(1)
    fields = []
    for line in body.splitlines():
        line = line.split("#")[0].strip()
        if not line:
            continue
        fname, ftype = [x.strip() for x in line.split("::")]
        m = re.match(r"NTuple\{(\d+), (\w+)\}", ftype)
        if m:
            fields.append((fname, _JL[m.group(2)][0] * int(m.group(1))))
        else:
            fields.append((fname, _JL[ftype][0]))
    return fields


def test_descriptor_structs_match_the_c_layout():
    pkg = ge.load_package()
    for jl, ct in (("CmdgDesc", pkg._lib.cmdg_desc), ("CmdgOceanDesc", pkg._lib.cmdg_ocean_desc)):
        got = _julia_struct(jl)
        want = list(ct._fields_)
        assert [f for f, _ in got] == [f for f, _ in want], jl
        for (f, a), (_, b) in zip(got, want):
            assert C.sizeof(a) == C.sizeof(b) and getattr(a, "_length_", 1) == getattr(b, "_length_", 1), (jl, f)
        # a Julia isbits struct uses C layout rules: same field sequence => same offsets and size
        class Mirror(C.Structure):
            _fields_ = got
        assert C.sizeof(Mirror) == C.sizeof(ct), jl


def test_fused_step_is_reachable_by_dispatch():
    """`LowStorageRungeKutta2N{T, RT, AT, Nstages}` (LowStorageRungeKuttaMethod.jl:26-42) has no
    right-hand-side type parameter: the shim must not dispatch on one.  It wraps the stock solver in
    `B200LSRK <: AbstractODESolver` from constructor methods that are more specific in the rhs argument."""
    assert "LowStorageRungeKutta2N{<:" not in SRC
    assert re.search(r"struct B200LSRK\{L <: LowStorageRungeKutta2N\} <: AbstractODESolver", SRC)
    assert re.search(r"^LSRK54CarpenterKennedy\(F::B200DGModel, Q::AT; dt = 0, t0 = 0\) where \{AT <: AbstractArray\}", SRC, flags=re.M)
    assert re.search(r"^LSRK144NiegemannDiehlBusch\(F::B200DGModel, Q::AT; dt = 0, t0 = 0\) where \{AT <: AbstractArray\}", SRC, flags=re.M)
    assert re.search(r"^function dostep!\(Q, s::B200LSRK, p, time, slow_δ = nothing", SRC, flags=re.M)
    # Filters.apply! keeps the reference's positional signature (Q, target, grid, filter; kwargs)
    sig = re.search(r"function Filters\.apply!\((.*?)\)\n", SRC, flags=re.S).group(1)
    pos = _split_top(sig.split(";")[0])
    assert len(pos) == 4 and "B200DGModel" not in sig, sig
    assert "HydrostaticBoussinesqModel" in SRC and "cmdg_set_ocean_model" in SRC
