"""Extracts the golden filter matrices the reference's own test holds
(/root/reference/test/Numerics/Mesh/filter.jl:15-75, hex-encoded Float64, computed there with the
nodal-dg Filter1D.m code) into tests/golden/filter_matrices.json.  Run in the build container only
(the GPU box has no /root/reference); the JSON is committed."""
import json
import re
import struct

src = open("/root/reference/test/Numerics/Mesh/filter.jl").read()
blocks = re.findall(r"W = \[(.*?)\]", src, flags=re.S)[:2]
out = {}
for name, blk, meta in zip(("exponential_N4_Nc0_s32", "exponential_N3_Nc1_s4"), blocks,
                           (dict(N=4, Nc=0, s=32), dict(N=3, Nc=1, s=4))):
    rows = [re.findall(r"0x([0-9a-f]{16})", ln) for ln in blk.strip().splitlines()]
    rows = [r for r in rows if r]
    W = [[struct.unpack(">d", bytes.fromhex(h))[0] for h in r] for r in rows]
    out[name] = dict(meta, hex=rows, W=W, source="test/Numerics/Mesh/filter.jl")
json.dump(out, open(__file__.replace("make_filter_golden.py", "filter_matrices.json"), "w"), indent=1)
print({k: (len(v["W"]), len(v["W"][0])) for k, v in out.items()})
