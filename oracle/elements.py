"""1-D element operators (test infrastructure -- see oracle/__init__.py).

Restates ``src/Numerics/Mesh/Elements.jl`` and the quadrature it takes from the
un-vendored GaussQuadrature.jl 0.5.5 (``Manifest.toml:357-361``).

* ``lglpoints``  <- ``Elements.jl:11-14`` (``GaussQuadrature.legendre(T, N+1, both)``).
  GaussQuadrature.jl uses Golub-Welsch; we use Newton on ``(1-x^2) P_N'(x)`` in
  extended precision and round, which yields the same correctly rounded nodes
  (pinned by ``test/Numerics/Mesh/Elements.jl`` exactness tests restated in
  ``tests/test_oracle_mesh.py``).
* ``baryweights`` <- ``Elements.jl:34-47``
* ``spectralderivative`` <- ``Elements.jl:60-82``
* ``interpolationmatrix`` <- ``Elements.jl:93-116``
* ``indefinite_integral_interpolation_matrix`` <- ``Grids.jl:1184-1206``
"""
import numpy as np


def _legendre_and_derivs(N, x):
    """P_N(x), P_N'(x), P_N''(x) by the three-term recurrence (long double)."""
    x = np.asarray(x, dtype=np.longdouble)
    p0 = np.ones_like(x)
    if N == 0:
        return p0, np.zeros_like(x), np.zeros_like(x)
    p1 = x.copy()
    for n in range(1, N):
        p0, p1 = p1, ((2 * n + 1) * x * p1 - n * p0) / (n + 1)
    # p1 = P_N, p0 = P_{N-1}
    with np.errstate(divide="ignore", invalid="ignore"):
        dp = N * (x * p1 - p0) / (x * x - 1)
        # Legendre ODE: (1-x^2) P'' - 2x P' + N(N+1) P = 0
        d2p = (2 * x * dp - N * (N + 1) * p1) / (1 - x * x)
    return p1, dp, d2p


def lglpoints(FT, N):
    """(N+1)-point Legendre-Gauss-Lobatto nodes and weights on [-1, 1]."""
    assert N >= 1
    FT = np.dtype(FT).type
    if N == 1:
        return np.array([-1, 1], dtype=FT), np.array([1, 1], dtype=FT)
    LD = np.longdouble
    # Chebyshev-Gauss-Lobatto initial guess for the interior nodes
    k = np.arange(1, N, dtype=LD)
    x = -np.cos(np.pi * k / N)
    for _ in range(100):
        _, dp, d2p = _legendre_and_derivs(N, x)
        dx = dp / d2p
        x = x - dx
        if np.max(np.abs(dx)) < 4 * np.finfo(LD).eps:
            break
    x = np.concatenate(([LD(-1)], x, [LD(1)]))
    # symmetrise
    x = (x - x[::-1]) / 2
    pN, _, _ = _legendre_and_derivs(N, x)
    pN[0] = (-1) ** N
    pN[-1] = 1
    w = 2 / (N * (N + 1) * pN * pN)
    return x.astype(FT), w.astype(FT)


def glpoints(FT, N):
    """(N+1)-point Gauss-Legendre rule (``Elements.jl:22-24``)."""
    x, w = np.polynomial.legendre.leggauss(N + 1)
    FT = np.dtype(FT).type
    return x.astype(FT), w.astype(FT)


def baryweights(r):
    r = np.asarray(r)
    Np = len(r)
    wb = np.ones(Np, dtype=r.dtype)
    for j in range(Np):
        for i in range(Np):
            if i != j:
                wb[j] = wb[j] * (r[j] - r[i])
        wb[j] = r.dtype.type(1) / wb[j]
    return wb


def spectralderivative(r, wb=None):
    """D[j, k] = d l_k / d xi (xi_j); same loop order as ``Elements.jl:60-82``."""
    r = np.asarray(r)
    if wb is None:
        wb = baryweights(r)
    Np = len(r)
    T = r.dtype.type
    D = np.zeros((Np, Np), dtype=r.dtype)
    for k in range(Np):
        for j in range(Np):
            if k == j:
                for l in range(Np):
                    if l != k:
                        D[j, k] = D[j, k] + T(1) / (r[k] - r[l])
            else:
                D[j, k] = (wb[k] / wb[j]) / (r[j] - r[k])
    return D


def interpolationmatrix(rsrc, rdst, wbsrc=None):
    rsrc = np.asarray(rsrc)
    rdst = np.asarray(rdst)
    if wbsrc is None:
        wbsrc = baryweights(rsrc)
    I = np.zeros((len(rdst), len(rsrc)), dtype=rsrc.dtype)
    for k in range(len(rdst)):
        for j in range(len(rsrc)):
            with np.errstate(divide="ignore"):
                I[k, j] = wbsrc[j] / (rdst[k] - rsrc[j])
            if not np.isfinite(I[k, j]):
                I[k, :] = 0
                I[k, j] = 1
                break
        I[k, :] = I[k, :] / np.sum(I[k, :])
    return I


def indefinite_integral_interpolation_matrix(r, w):
    r = np.asarray(r)
    w = np.asarray(w)
    Nq = len(r)
    Iint = np.zeros((Nq, Nq), dtype=r.dtype)
    Iint[0, :] = w[0] if Nq == 1 else 0
    wbary = baryweights(r)
    for n in range(1, Nq):
        rdst = (1 - r) / 2 * r[0] + (1 + r) / 2 * r[n]
        In = interpolationmatrix(r, rdst, wbary)
        delta = (r[n] - r[0]) / 2
        Iint[n, :] = delta * (w @ In)
    return Iint
