"""HydrostaticBoussinesqModel (ocean) balance law (test infrastructure -- see oracle/__init__.py).

Restates ``src/Ocean/HydrostaticBoussinesq/hydrostatic_boussinesq_model.jl`` (Uncoupled,
NonLinearAdvectionTerm for tracers, no momentum advection by default), its boundary
conditions (``bc_velocity.jl``, ``bc_temperature.jl``, ``src/Ocean/OceanBC.jl``), the
``OceanGyre`` problem (``src/Ocean/OceanProblems/ocean_gyre.jl``), the vertical filters it
applies inside every tendency evaluation (``src/Numerics/Mesh/Filters.jl:114-229, 275-314,
651-792``) and the column integrals of ``update_auxiliary_state_gradient!``
(``DGModel_kernels.jl:1903-2104``).

State layouts (``hydrostatic_boussinesq_model.jl:105-232``): prognostic ``u[2], eta, theta``;
auxiliary ``y, w, pkin, wz0, ud[2], dGu[2]``; gradient ``u[2], ud[2], theta``; gradient flux
``div_h u, nu grad u (3x2, column major), kappa grad theta[3]``.

Pinned end to end by ``test/Ocean/refvals/test_ocean_gyre_refvals.jl`` (``short``: 5x5x5
elements, N = 4, dt = 120 s, 30 LSRK144 steps; min/max/mean/std of every field to 10-12
digits) in ``tests/test_oracle_ocean.py``.
"""
import numpy as np

from . import grids as G


# --- spectral filters ---------------------------------------------------------------
def _orthonormal_legendre_vandermonde(r):
    """V[i, n] = orthonormal Legendre polynomial of degree n at r[i]
    (GaussQuadrature.orthonormal_poly with legendre_coefs)."""
    N = len(r) - 1
    V = np.zeros((N + 1, N + 1))
    for n in range(N + 1):
        c = np.zeros(n + 1)
        c[n] = 1
        V[:, n] = np.polynomial.legendre.legval(r, c) * np.sqrt((2 * n + 1) / 2)
    return V


def spectral_filter_matrix(r, Nc, sigma):
    """``Filters.jl:114-131``: V diag(sigma((n-Nc)/(N-Nc)), n >= Nc) V^-1."""
    N = len(r) - 1
    V = _orthonormal_legendre_vandermonde(np.asarray(r, dtype=np.float64))
    S = np.ones(N + 1)
    for n in range(Nc, N + 1):
        S[n] = sigma((n - Nc) / (N - Nc))
    return (V * S[None, :]) @ np.linalg.inv(V)


def cutoff_filter_matrix(r, Nc):
    return spectral_filter_matrix(r, Nc, lambda eta: 0.0)


def exponential_filter_matrix(r, Nc, s, alpha=None):
    alpha = -np.log(np.finfo(np.float64).eps) if alpha is None else alpha
    return spectral_filter_matrix(r, Nc, lambda eta: np.exp(-alpha * eta ** s))


def apply_vertical_filter(g, data, states, F):
    """kernel_apply_filter! with direction = VerticalDirection on real elements:
    Q[i,j,k,s] <- sum_n F[k,n] Q[i,j,n,s] (accumulated n = 1..Nq)."""
    nr, Nq = g.nreal, g.Nq
    for s in states:
        q = data[:nr, s, :].reshape(nr, Nq[2], Nq[1], Nq[0])
        out = np.zeros_like(q)
        for n in range(Nq[2]):
            out = out + F[:, n][None, :, None, None] * q[:, n:n + 1]
        data[:nr, s, :] = out.reshape(nr, g.Np)


class OceanGyre:
    """``ocean_gyre.jl``: wind stress, temperature relaxation, initial stratification."""

    def __init__(self, Lx, Ly, H, tau0=1e-1, lambda_r=4 / 86400, thetaE=10.0):
        self.Lx, self.Ly, self.H = float(Lx), float(Ly), float(H)
        self.tau0, self.lambda_r, self.thetaE = tau0, lambda_r, thetaE

    def init_state(self, x, y, z):
        Q = np.zeros((4,) + y.shape)
        Q[3] = (5 + 4 * np.cos(y * np.pi / self.Ly)) * (1 + z / self.H)
        return Q

    def kinematic_stress(self, y, rho):
        return [(self.tau0 / rho) * np.cos(y * np.pi / self.Ly), 0 * y]

    def surface_flux(self, y, theta):
        theta_r = self.thetaE * (1 - y / self.Ly)
        return self.lambda_r * (theta - theta_r)


class SimpleBox:
    """``SimpleBox`` with ``Fixed`` rotation (``src/Ocean/OceanProblems/simple_box_problem.jl:96-236``): the
    analytic barotropic + baroclinic spin-down of a standing gravity wave, no wind stress, no surface heat flux,
    Coriolis parameter identically zero."""

    tau0, lambda_r, thetaE = 0.0, 0.0, 0.0

    def __init__(self, Lx, Ly, H):
        self.Lx, self.Ly, self.H = float(Lx), float(Ly), float(H)

    def init_state(self, x, y, z, t=0.0, model=None):
        from scipy.linalg import expm
        m = model
        kx, kz = 2 * np.pi / self.Lx, 2 * np.pi / self.H
        gH = m.grav * self.H
        M = np.array([[-m.nuh * kx ** 2, gH * kx], [-kx, 0.0]])
        A = expm(M * t) @ np.ones(2)
        U = A[0] * np.sin(kx * x)
        eta = A[1] * np.cos(kx * x)
        lam = m.nuh * kx ** 2 + m.nuz * kz ** 2
        u0 = np.exp(-lam * t) * np.cos(kz * z) * np.sin(kx * x)
        Q = np.zeros((4,) + y.shape)
        Q[0] = u0 + U / self.H
        Q[2] = eta
        return Q

    def kinematic_stress(self, y, rho):
        return [0 * y, 0 * y]

    def surface_flux(self, y, theta):
        return 0 * y


class HBModel:
    """Pointwise physics + the model's update_auxiliary_state hooks."""
    S, A, Gn, GF = 4, 8, 5, 10
    a_y, a_w, a_pkin, a_wz0 = 0, 1, 2, 3

    def __init__(self, problem, grav=9.81, rho0=1000.0, ch=None, cz=0.0, alphaT=2e-4, nuh=5e3,
                 nuz=5e-3, kappah=1e3, kappaz=1e-4, kappac=1e-1, f0=1e-4, beta=1e-11,
                 bcs=(("noslip", "insulating"), ("noslip", "insulating"),
                      ("kinematic_stress", "temperature_flux")),
                 vert_filter=None, exp_filter=None):
        self.problem = problem
        self.grav, self.rho0 = grav, rho0
        self.ch = np.sqrt(grav * problem.H) if ch is None else ch
        self.cz, self.alphaT = cz, alphaT
        self.nuh, self.nuz, self.kappah, self.kappaz, self.kappac = nuh, nuz, kappah, kappaz, kappac
        self.f0, self.beta = f0, beta
        self.bcs = bcs
        self.vert_filter, self.exp_filter = vert_filter, exp_filter
        self.G = self.Gn
        self.FT = np.float64

    def viscous(self):
        return True

    # -- first order ---------------------------------------------------------------
    def flux_first_order(self, Q, aux):
        u1, u2, eta, th = Q
        w, pkin = aux[self.a_w], aux[self.a_pkin]
        v = [u1, u2, w]
        F = np.zeros((3, 4) + Q.shape[1:])
        pr = self.grav * eta
        pk = self.grav * pkin
        F[0, 0] = F[0, 0] + pr
        F[1, 1] = F[1, 1] + pr
        F[0, 0] = F[0, 0] + pk
        F[1, 1] = F[1, 1] + pk
        for d in range(3):
            F[d, 3] = v[d] * th
        return F

    def flux_second_order(self, Q, GF, aux):
        F = np.zeros((3, 4) + Q.shape[1:])
        for d in range(3):
            F[d, 0] = GF[1 + d]
            F[d, 1] = GF[4 + d]
            F[d, 3] = GF[7 + d]
        return F

    def source(self, Q, aux):
        S = np.zeros_like(Q)
        f = self.f0 + self.beta * aux[self.a_y]
        S[2] = aux[self.a_wz0]
        S[0] = -(-f * Q[1])
        S[1] = -(f * Q[0])
        return S

    def wavespeed(self, n, Q, aux):
        return np.abs(self.ch * n[0] + self.ch * n[1] + self.cz * n[2])

    def update_penalty(self, penalty):
        penalty[2] = -0.0 * penalty[2]

    # -- gradients -------------------------------------------------------------------
    def gradient_argument(self, Q, aux):
        Gt = np.zeros((5,) + Q.shape[1:])
        Gt[0:2] = Q[0:2]
        Gt[4] = Q[3]
        return Gt

    def gradient_flux(self, gradG, Q, aux):
        GFv = np.zeros((10,) + Q.shape[1:])
        GFv[0] = gradG[0, 0] + gradG[1, 1]
        nu = [self.nuh, self.nuh, self.nuz]
        for c in range(2):
            for d in range(3):
                GFv[1 + 3 * c + d] = -nu[d] * gradG[d, c]
        dthz = gradG[2, 4]
        kz = np.where(dthz < 0, self.kappac, self.kappaz)
        GFv[7] = -self.kappah * gradG[0, 4]
        GFv[8] = -self.kappah * gradG[1, 4]
        GFv[9] = -kz * gradG[2, 4]
        return GFv

    # -- boundary conditions ---------------------------------------------------------
    def boundary_state(self, kind, bctag, n, Qm, auxm):
        Qp, auxp = Qm.copy(), auxm.copy()
        vel, temp = self.bcs[bctag - 1]
        if vel == "noslip":
            if kind == "first":
                Qp[0:2] = -Qm[0:2]
                auxp[self.a_w] = -auxm[self.a_w]
            else:
                Qp[0:2] = 0
                auxp[self.a_w] = 0
        elif vel == "freeslip":
            v = [Qm[0], Qm[1], auxm[self.a_w]]
            vn = n[0] * v[0] + n[1] * v[1] + n[2] * v[2]
            fac = 2 if kind == "first" else 1
            Qp[0] = v[0] - fac * vn * n[0]
            Qp[1] = v[1] - fac * vn * n[1]
            auxp[self.a_w] = v[2] - fac * vn * n[2]
        elif vel in ("kinematic_stress", "penetrable_freeslip"):
            pass  # Penetrable: transmissive ghost state
        else:
            raise ValueError(vel)
        # Insulating / TemperatureFlux: theta+ = theta-
        return Qp, auxp

    def boundary_flux_second_order(self, bctag, n, Qm, GFm, auxm):
        """normal_boundary_flux_second_order! default: boundary_state!(nf2, ...) on the + copy,
        then the full second-order flux of the + side (NumericalFluxes.jl:872-967)."""
        vel, temp = self.bcs[bctag - 1]
        GFp = GFm.copy()
        if vel == "noslip":
            pass  # nu grad u+ = nu grad u-
        elif vel in ("freeslip", "penetrable_freeslip"):
            GFp[1:7] = 0
        elif vel == "kinematic_stress":
            st = self.problem.kinematic_stress(auxm[self.a_y], self.rho0)
            for c in range(2):
                for d in range(3):
                    GFp[1 + 3 * c + d] = n[d] * st[c]
        if temp == "insulating":
            GFp[7:10] = 0
        elif temp == "temperature_flux":
            sf = self.problem.surface_flux(auxm[self.a_y], Qm[3])
            for d in range(3):
                GFp[7 + d] = n[d] * sf
        return self.flux_second_order(Qm, GFp, auxm)

    # -- model hooks of the DG schedule ---------------------------------------------
    def init_state_auxiliary(self, dg):
        for g, aux in zip(dg.grids, dg.state_auxiliary):
            aux.data[:g.nreal, self.a_y, :] = g.vgeo[:g.nreal, G._x2, :]

    def update_auxiliary_state(self, dg, Q, elems):
        """hydrostatic_boussinesq_model.jl:637-663: vertical cutoff filter on u, vertical
        exponential filter on theta (real elements only)."""
        if elems != "real":
            return
        for g, q in zip(dg.grids, Q):
            if self.vert_filter is not None:
                apply_vertical_filter(g, q.data, (0, 1), self.vert_filter)
            if self.exp_filter is not None:
                apply_vertical_filter(g, q.data, (3,), self.exp_filter)

    def update_auxiliary_state_gradient(self, dg, Q, elems):
        """:675-712: w = -div_h u; upward integrals of (w, -alphaT theta) with Imat * JcV;
        pkin <- pkin(top) - pkin; wz0 <- w at the top node of the column."""
        for g, q, aux, gf in zip(dg.grids, Q, dg.state_auxiliary, dg.state_gradient_flux):
            nv = g.topology.stacksize
            sl = slice(0, g.nreal) if elems == "real" else slice(g.nreal, g.nelem)
            a, qd, gd = aux.data[sl], q.data[sl], gf.data[sl]
            ne = a.shape[0]
            if ne == 0:
                continue
            Nq = g.Nq
            Nqh = Nq[0] * Nq[1]
            a[:, self.a_w, :] = -gd[:, 0, :]
            nh = ne // nv
            kern = np.stack([a[:, self.a_w, :], -self.alphaT * qd[:, 3, :]])
            out = indefinite_stack_integral(g, kern, sl)
            w = out[0].reshape(ne, g.Np)
            pk = out[1]
            top = pk[:, nv - 1, Nq[2] - 1, :]
            pk = top[:, None, None, :] - pk
            a[:, self.a_w, :] = w
            a[:, self.a_pkin, :] = pk.reshape(ne, g.Np)
            if elems == "real":
                wtop = out[0][:, nv - 1, Nq[2] - 1, :]
                wz0 = np.broadcast_to(wtop[:, None, None, :], (nh, nv, Nq[2], Nqh))
                a[:, self.a_wz0, :] = wz0.reshape(ne, g.Np)


def indefinite_stack_integral(g, kern, sl=None):
    """``kernel_indefinite_stack_integral!`` (DGModel_kernels.jl:1903-1990): upward integral of ``kern``
    (nfields, nelem_in_sl, Np) along every vertical stack, element by element with ``Imat`` (the
    indefinite-integral interpolation matrix of the vertical LGL points) times ``JcV``, carrying the value
    at the top node of an element into the next.  Returns (nfields, nstacks, nvertelem, Nq3, Nq1*Nq2)."""
    sl = slice(0, g.nreal) if sl is None else sl
    nv = g.topology.stacksize
    Nq = g.Nq
    Nqh = Nq[0] * Nq[1]
    nf, ne = kern.shape[0], kern.shape[1]
    nh = ne // nv
    Imat = g.Imat[2]
    JcV = g.vgeo[sl, G._JcV, :].reshape(nh, nv, Nq[2], Nqh)
    kern = kern.reshape(nf, nh, nv, Nq[2], Nqh) * JcV[None]
    out = np.zeros_like(kern)
    carry = np.zeros((nf, nh, Nqh), dtype=kern.dtype)
    for ev in range(nv):
        li = np.repeat(carry[:, :, None, :], Nq[2], axis=2)
        for n in range(Nq[2]):
            li = li + Imat[:, n][None, None, :, None] * kern[:, :, ev, n:n + 1, :]
        out[:, :, ev] = li
        carry = li[:, :, Nq[2] - 1, :]
    return out


def statecheck(arr, ivar):
    """min, max, mean, std (n-1) of one field over real elements (StateCheck.jl:231-282)."""
    v = arr.realdata[:, ivar, :].ravel()
    mean = v.mean()
    return v.min(), v.max(), mean, np.sqrt(np.sum((v - mean) ** 2) / (v.size - 1))
