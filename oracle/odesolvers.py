"""Low-storage Runge-Kutta 2N time stepping (test infrastructure -- see oracle/__init__.py).

Restates ``src/Numerics/ODESolvers/LowStorageRungeKuttaMethod.jl``:
``dostep!`` (``:102-144``), ``update!`` (``:146-158``), LSRK54CarpenterKennedy
tableau (``:293-327``), LSRK144NiegemannDiehlBusch (``:349-410``), and the
``solve!`` / ``general_dostep!`` loop of ``ODESolvers.jl:49-158``.
"""
from fractions import Fraction as Fr

import numpy as np

LSRK54_RKA = (Fr(0), Fr(-567301805773, 1357537059087), Fr(-2404267990393, 2016746695238),
              Fr(-3550918686646, 2091501179385), Fr(-1275806237668, 842570457699))
LSRK54_RKB = (Fr(1432997174477, 9575080441755), Fr(5161836677717, 13612068292357),
              Fr(1720146321549, 2090206949498), Fr(3134564353537, 4481467310338),
              Fr(2277821191437, 14882151754819))
LSRK54_RKC = (Fr(0), Fr(1432997174477, 9575080441755), Fr(2526269341429, 6820363962896),
              Fr(2006345519317, 3224310063776), Fr(2802321613138, 2924317926251))

LSRK144_RKA = (0.0, -0.7188012108672410, -0.7785331173421570, -0.0053282796654044,
               -0.8552979934029281, -3.9564138245774565, -1.5780575380587385,
               -2.0837094552574054, -0.7483334182761610, -0.7032861106563359,
               0.0013917096117681, -0.0932075369637460, -0.9514200470875948,
               -7.1151571693922548)
LSRK144_RKB = (0.0367762454319673, 0.3136296607553959, 0.1531848691869027, 0.0030097086818182,
               0.3326293790646110, 0.2440251405350864, 0.3718879239592277, 0.6204126221582444,
               0.1524043173028741, 0.0760894927419266, 0.0077604214040978, 0.0024647284755382,
               0.0780348340049386, 5.5059777270269628)
LSRK144_RKC = (0.0, 0.0367762454319673, 0.1249685262725025, 0.2446177702277698,
               0.2476149531070420, 0.2969311120382472, 0.3978149645802642, 0.5270854589440328,
               0.6981269994175695, 0.8190890835352128, 0.8527059887098624, 0.8604711817462826,
               0.8627060376969976, 0.8734213127600976)


def _conv(FT, xs):
    FT = np.dtype(FT).type
    return tuple(FT(x.numerator / x.denominator) if isinstance(x, Fr) else FT(x) for x in xs)


class LowStorageRungeKutta2N:
    def __init__(self, rhs, RKA, RKB, RKC, Q, dt=0.0, t0=0.0):
        """``Q``: list of per-rank MPIStateArrays; ``rhs(dQ, Q, t, increment=...)``."""
        self.rhs = rhs
        FT = Q[0].data.dtype
        self.RKA, self.RKB, self.RKC = _conv(FT, RKA), _conv(FT, RKB), _conv(FT, RKC)
        self.dt = FT.type(dt)
        self.t = FT.type(t0)
        self.dQ = [q.similar() for q in Q]
        self.steps = 0

    def dostep(self, Q, time):
        dt = self.dt
        ns = len(self.RKA)
        for s in range(ns):
            self.rhs(self.dQ, Q, time + self.RKC[s] * dt, increment=True)
            rka = self.RKA[(s + 1) % ns]
            for q, dq in zip(Q, self.dQ):
                rq, rdq = q.realdata, dq.realdata
                rq += self.RKB[s] * dt * rdq
                rdq *= rka

    def general_dostep(self, Q, timeend, adjustfinalstep=True):
        time, dt = self.t, self.dt
        final = False
        if adjustfinalstep and time + dt > timeend:
            orig = dt
            self.dt = timeend - time
            final = True
        self.dostep(Q, time)
        if not final:
            self.t = time + dt
        else:
            self.dt = orig
            self.t = timeend
        return self.t


def LSRK54CarpenterKennedy(rhs, Q, dt=0.0, t0=0.0):
    return LowStorageRungeKutta2N(rhs, LSRK54_RKA, LSRK54_RKB, LSRK54_RKC, Q, dt, t0)


def LSRK144NiegemannDiehlBusch(rhs, Q, dt=0.0, t0=0.0):
    return LowStorageRungeKutta2N(rhs, LSRK144_RKA, LSRK144_RKB, LSRK144_RKC, Q, dt, t0)


def solve(Q, solver, timeend=np.inf, numberofsteps=0, adjustfinalstep=True, callback=None):
    assert np.isfinite(timeend) or numberofsteps > 0
    step = 0
    time = solver.t
    while time < timeend:
        step += 1
        solver.steps = step
        time = solver.general_dostep(Q, timeend, adjustfinalstep)
        if callback is not None:
            callback(step, time)
        if step == numberofsteps:
            break
    return solver.t
