"""Prints Float32 parity numbers of the LES box and the rising bubble (debug aid)."""
import sys, os, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import parity
for name, fn in (("box smag f32", lambda: parity.box_case(nsteps=2, FT=np.float32)),
                 ("box ck75 f32", lambda: parity.box_case(nsteps=2, FT=np.float32, turbulence=("constant_kinematic", 75.0, False))),
                 ("bubble f64", lambda: parity.risingbubble_case(nsteps=2)),
                 ("bubble f32", lambda: parity.risingbubble_case(nsteps=2, FT=np.float32))):
    try:
        print(name, fn(), flush=True)
    except Exception:
        traceback.print_exc()
