"""Tracer code paths besides the Smagorinsky one: constant viscosity (D_t computed in the tracer kernel)
and the inviscid path (skip_zero_viscosity, no tracer gradient kernel)."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import parity
for name, kw in (("const kinematic", dict(turbulence=("constant_kinematic", 75.0, False))),
                 ("const dynamic", dict(turbulence=("constant_dynamic", 50.0, True))),
                 ("inviscid skip", dict(turbulence=("constant_kinematic", 0.0, False), skip_zero_viscosity=True))):
    try:
        print(name, parity.risingbubble_case(nelem=(5, 1, 5), nsteps=1, tracers=(1.0, 3.0), **kw), flush=True)
    except Exception:
        traceback.print_exc()
