"""Quick device timing of the fused LSRK54 stage kernel on a periodic-box vortex (dev tool;
uses the oracle only to build inputs)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tests import parity
from oracle import grids as ogrids, dgmodel as odg

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 24
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
t0 = time.time()
model, gs, setup, dt = parity.vortex_setup((ne, ne, ne))
g = gs[0]
print("grid built", time.time() - t0, "s; nelem", g.nreal, flush=True)
P = parity.pkg()
odgm = odg.DGModel(model, [g], "rusanov", skip_zero_viscosity=True)
dg, dgrid = parity.make_device_dg(odgm, g, "rusanov", skip_zero_viscosity=True)
Q0 = setup(g.vgeo[:g.nreal, ogrids._x1], g.vgeo[:g.nreal, ogrids._x2], g.vgeo[:g.nreal, ogrids._x3], np.float64(0))
Q = P.MPIStateArray(dgrid, 5)
Q.data[:g.nreal] = torch.as_tensor(np.ascontiguousarray(np.moveaxis(Q0, 0, 1))).cuda()
sol = P.LSRK54CarpenterKennedy(dg, Q, dt=dt, t0=0.0)
sol.dostep(Q, 0.0, nsteps=3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sol.dostep(Q, 0.0, nsteps=nsteps)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
nodes = g.nreal * 125
per_stage = ms / (nsteps * 5)
print(f"ne={ne}^3 nodes={nodes/1e6:.2f}M  {ms/nsteps:.3f} ms/step  {per_stage*1e3:.1f} us/stage")
print(f"GDOF/s = {nodes*5/per_stage/1e6:.2f};  algorithmic GB/s (353.6 B/node) = {nodes*353.6/per_stage/1e6:.1f}")
print("norm", P.norm(Q))
