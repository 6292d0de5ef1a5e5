"""Prints the parity numbers of the DryBiharmonic cases without asserting (debug aid)."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import parity
for kind, turb in (("box", ("constant_kinematic", 75.0, False)), ("sphere", ("constant_kinematic", 0.0, False)),
                   ("sphere", ("smagorinsky", 0.21)), ("walled_box", ("smagorinsky", 0.21))):
    try:
        print(kind, turb, parity.hyperdiffusion_case(kind=kind, turbulence=turb, nsteps=2), flush=True)
    except Exception:
        traceback.print_exc()
