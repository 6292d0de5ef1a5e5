"""Prints the parity numbers of the tracer cases without asserting (debug aid)."""
import sys, os, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import parity
for name, fn in (("bubble + 4 tracers", lambda: parity.risingbubble_case(nsteps=2, tracers=(1.0, 2.0, 3.0, 4.0))),
                 ("bubble + 2 tracers central", lambda: parity.risingbubble_case(nsteps=1, tracers=(0.5, 2.0), nf="central")),
)[:2]:
    try:
        print(name, fn(), flush=True)
    except Exception:
        traceback.print_exc()
