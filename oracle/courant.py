"""Courant numbers and node distances (test infrastructure -- see oracle/__init__.py).

Restates
* ``kernel_min_neighbor_distance!`` / ``min_node_distance`` (``src/Numerics/Mesh/Grids.jl:440-486,
  1219-1336``): per node, the minimum physical distance to its +-1 neighbours along the selected
  reference directions;
* ``courant(local_courant, dg, m, Q, dt, simtime, direction)``
  (``src/Numerics/DGMethods/SpaceDiscretization.jl:307-365`` with ``kernel_local_courant!``,
  ``DGModel_kernels.jl:3028-3100``): the maximum of the pointwise number over the real elements
  (the reference then takes ``MPI.Allreduce(max)``);
* the AtmosModel pointwise numbers ``advective_courant``, ``nondiffusive_courant``,
  ``diffusive_courant`` (``src/Atmos/Model/courant.jl:12-86``).

Pinned on the analytic expectation of ``test/Numerics/Mesh/min_node_distance.jl:60-83``.
"""
import numpy as np

from . import grids as G
from . import atmos as A


def min_neighbor_distance(g, direction="every"):
    """(nreal, Np) array of ``kernel_min_neighbor_distance!``."""
    Nq = g.Nq
    nr = g.nreal
    x = np.stack([g.vgeo[:nr, c, :] for c in (G._x1, G._x2, G._x3)], axis=-1)
    x = x.reshape(nr, Nq[2], Nq[1], Nq[0], 3)
    md = np.full((nr, Nq[2], Nq[1], Nq[0]), np.finfo(g.FT).max, dtype=g.FT)
    use = {"every": (True, True, True), "horizontal": (True, True, False),
           "vertical": (False, False, True)}[direction]
    for d, ax in ((0, 3), (1, 2), (2, 1)):   # xi1 -> axis 3, xi2 -> axis 2, xi3 -> axis 1
        if not use[d] or x.shape[ax] < 2:
            continue
        diff = np.sqrt(((np.diff(x, axis=ax)) ** 2).sum(-1))    # distance to the next node
        lo = [slice(None)] * 4
        hi = [slice(None)] * 4
        lo[ax], hi[ax] = slice(0, -1), slice(1, None)
        md[tuple(lo)] = np.minimum(md[tuple(lo)], diff)          # neighbour at +1
        md[tuple(hi)] = np.minimum(md[tuple(hi)], diff)          # neighbour at -1
    return md.reshape(nr, g.Np)


def min_node_distance(g, direction="every"):
    return float(min_neighbor_distance(g, direction).min())


def _norm_u(Q, k, direction):
    u = Q[1:4] / Q[0]
    if direction == "vertical":
        return np.abs(Q[1] * k[0] + Q[2] * k[1] + Q[3] * k[2]) / Q[0]
    if direction == "horizontal":
        dot = Q[1] * k[0] + Q[2] * k[1] + Q[3] * k[2]
        v = (Q[1:4] - dot * k) / Q[0]
        return np.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2)
    return np.sqrt(u[0] ** 2 + u[1] ** 2 + u[2] ** 2)


def _norm_nu(nu, k, direction):
    if np.ndim(nu[0]) == 0 and all(np.all(n == nu[0]) for n in nu) and False:
        return nu[0]
    nu = np.stack(nu)
    if direction == "vertical":
        return nu[0] * k[0] + nu[1] * k[1] + nu[2] * k[2]
    if direction == "horizontal":
        dot = nu[0] * k[0] + nu[1] * k[1] + nu[2] * k[2]
        v = nu - dot * k
        return np.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2)
    return np.sqrt(nu[0] ** 2 + nu[1] ** 2 + nu[2] ** 2)


def pointwise_courant(model, g, Q, aux, GF, dt, kind="nondiffusive", direction="every"):
    """(nreal, Np) pointwise Courant numbers; Q/aux/GF are ``(nelem, ns, Np)`` arrays."""
    nr = g.nreal
    dx = min_neighbor_distance(g, direction)
    q = np.moveaxis(Q[:nr], 1, 0)
    a = np.moveaxis(aux[:nr], 1, 0)
    if model.a_gradΦ is not None:
        k = a[model.a_gradΦ] / model.ps.grav
    else:
        k = np.zeros((3,) + q.shape[1:], dtype=q.dtype)
    if kind == "advective":
        return dt * _norm_u(q, k, direction) / dx
    if kind == "nondiffusive":
        T, _ = model.thermo(q, a)
        return dt * (_norm_u(q, k, direction) + A.soundspeed_air(model.ps, T)) / dx
    if kind == "diffusive":
        gf = np.moveaxis(GF[:nr], 1, 0)
        D_t, _ = model.turbulence_tensors(q, gf, a)
        if model.turbulence[0] == "smagorinsky":
            nu = [d / model.ps.inv_Pr_turb for d in D_t]      # nu = D_t / inv_Pr_turb (:472-499)
            nrm = _norm_nu(nu, k, direction)
        else:
            nrm = D_t[0] / model.ps.inv_Pr_turb + 0 * q[0]     # scalar nu: norm_nu(nu::Real) = nu
        return dt * nrm / (dx * dx)
    raise ValueError(kind)


def courant(model, g, Q, aux, GF, dt, kind="nondiffusive", direction="every"):
    return float(pointwise_courant(model, g, Q, aux, GF, dt, kind, direction).max())
