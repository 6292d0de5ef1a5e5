"""DGModel tendency evaluation (test infrastructure -- see oracle/__init__.py).

Restates, with the reference's launch schedule and accumulation order:

* ``(dg::DGModel)(tendency, Q, _, t, alpha, beta)``      <- ``DGModel.jl:85-427``
* ``launch_volume_tendency!`` (H then V, sources in V)   <- ``SpaceDiscretization.jl:1090-1203``
* ``volume_tendency!`` H / V kernels                     <- ``DGModel_kernels.jl:64-309, 312-548``
* ``launch_interface_tendency!`` / ``dgsem_interface_tendency!``
                                                         <- ``SpaceDiscretization.jl:1212-1368``, ``DGModel_kernels.jl:588-901``
* ``launch_volume_gradients!`` / ``volume_gradients!``   <- ``SpaceDiscretization.jl:502-585``, ``DGModel_kernels.jl:934-1328``
* ``dgsem_interface_gradients!``                         <- ``DGModel_kernels.jl:1365-1651``
* nodal auxiliary update                                 <- ``DGModel_kernels.jl:1769-1825``, ``SpaceDiscretization.jl:157-212``
* hyperdiffusion passes (DryBiharmonic, horizontal diffusion direction): gradients stored by the
  gradient kernels (``DGModel_kernels.jl:1081-1098,1618-1628``), ``volume_divergence_of_gradients!``
  (``:2132-2228``), ``interface_divergence_of_gradients!`` (``:2359-2490``),
  ``volume_gradients_of_laplacians!`` (``:2521-2680``), ``interface_gradients_of_laplacians!``
  (``:2860-3026``), ``CentralNumericalFluxDivergence`` / ``CentralNumericalFluxHigherOrder``
  (``NumericalFluxes.jl:716-835``), schedule ``DGModel.jl:226-310``
* numerical fluxes Rusanov / Central / Roe, boundary-flux plumbing,
  central gradient and second-order fluxes               <- ``NumericalFluxes.jl:65-123,163-340,668-715,872-918``
* auxiliary initialisation (coord, Phi, grad Phi by the element-local strong
  derivative, hydrostatic reference state with the discrete pressure-gradient
  balance)                                               <- ``AtmosModel.jl:880-925``, ``Orientations.jl``,
                                                            ``DGModel_kernels.jl:3097-3232``, ``ref_state.jl:70-264``

The operator acts on the arrays of *all* emulated ranks at once (lists indexed
by rank), performing the ghost exchanges where the reference does.
"""
import numpy as np

from . import grids as G
from .mpistatearrays import MPIStateArray, ghost_exchange
from . import atmos as A


def _sv(a):
    """(nelem, S, Np) -> state-major view (S, nelem, Np)."""
    return np.moveaxis(a, 1, 0)


class DGModel:
    def __init__(self, balance_law, grids, numerical_flux_first_order="rusanov",
                 numerical_flux_second_order="central", numerical_flux_gradient="central",
                 direction="every", diffusion_direction="every", skip_zero_viscosity=False,
                 init_aux=True):
        self.bl = balance_law
        self.grids = grids if isinstance(grids, (list, tuple)) else [grids]
        self.nf1 = numerical_flux_first_order
        assert numerical_flux_second_order == "central" and numerical_flux_gradient == "central"
        assert direction == "every"
        self.diffusion_direction = diffusion_direction
        self.skip_zero_viscosity = skip_zero_viscosity
        bl = balance_law
        self.state_auxiliary = [MPIStateArray.from_grid(g, bl.A) for g in self.grids]
        self.state_gradient_flux = [MPIStateArray.from_grid(g, bl.GF) for g in self.grids]
        self.nhyper = getattr(bl, "nhyper", 0)
        if self.nhyper:
            # states_higher_order (create_states.jl:20-27): max(3 * ngradlap, nhyper) and ngradlap columns
            assert diffusion_direction == "horizontal", \
                "hyperdiffusion: only HorizontalDirection (the reference's 3-D EveryDirection kernel is broken)"
            self.Qhypervisc_grad = [MPIStateArray.from_grid(g, max(3 * bl.ngradlap, bl.nhyper)) for g in self.grids]
            self.Qhypervisc_div = [MPIStateArray.from_grid(g, bl.ngradlap) for g in self.grids]
        if init_aux:
            self.init_state_auxiliary()

    # ------------------------------------------------------------------
    # auxiliary state
    # ------------------------------------------------------------------
    def local_gradient(self, g, f):
        """Element-local strong-form gradient of nodal field ``f`` (nelem, Np) on real
        elements: ``dgsem_auxiliary_field_gradient!`` H launch then V launch (+=)."""
        nr = g.nreal
        Nq = g.Nq
        fr = f[:nr].reshape(nr, Nq[2], Nq[1], Nq[0])
        D1, D2, D3 = g.D
        G1 = np.zeros_like(fr)
        G2 = np.zeros_like(fr)
        for n in range(Nq[0]):
            G1 = G1 + D1[:, n] * fr[:, :, :, n:n + 1]
            G2 = G2 + D2[:, n][:, None] * fr[:, :, n:n + 1, :]
        G3 = np.zeros_like(fr)
        for k in range(Nq[2]):  # Gxi3[n] += D[n, k] * f[k]
            G3 = G3 + D3[:, k][:, None, None] * fr[:, k:k + 1, :, :]
        vg = g.vgeo[:nr]
        G1, G2, G3 = [x.reshape(nr, g.Np) for x in (G1, G2, G3)]
        out = []
        for d in range(3):
            h = vg[:, G._xi1x1 + 3 * d] * G1
            h = h + vg[:, G._xi2x1 + 3 * d] * G2
            out.append(h + vg[:, G._xi3x1 + 3 * d] * G3)
        return out

    def init_state_auxiliary(self):
        bl = self.bl
        if hasattr(bl, "init_state_auxiliary"):
            bl.init_state_auxiliary(self)
            ghost_exchange(self.state_auxiliary)
            return
        ps = bl.ps
        # Orientation (Orientations.jl init_aux!)
        for g, aux in zip(self.grids, self.state_auxiliary):
            a = _sv(aux.data)
            nr = g.nreal
            x = [g.vgeo[:nr, G._x1], g.vgeo[:nr, G._x2], g.vgeo[:nr, G._x3]]
            if bl.orientation == "spherical":
                a[bl.a_Φ][:nr] = ps.grav * (np.sqrt(x[0] ** 2 + x[1] ** 2 + x[2] ** 2) - ps.planet_radius)
            elif bl.orientation == "flat":
                a[bl.a_Φ][:nr] = ps.grav * x[2]
        if bl.orientation != "none":
            ghost_exchange(self.state_auxiliary)
            for g, aux in zip(self.grids, self.state_auxiliary):
                a = _sv(aux.data)
                gr = self.local_gradient(g, a[bl.a_Φ])
                for d in range(3):
                    a[bl.a_gradΦ][d][:g.nreal] = gr[d]
        # Reference state (ref_state.jl:204-264)
        if bl.ref_state is not None:
            rs = bl.ref_state
            for g, aux in zip(self.grids, self.state_auxiliary):
                a = _sv(aux.data)
                nr = g.nreal
                z = a[bl.a_Φ][:nr] / ps.grav
                Tv, p = A.reference_profile(ps, rs, z, bl.FT)
                a[bl.a_ref["p"]][:nr] = p
                a[bl.a_ref["ρ"]][:nr] = p / (Tv * ps.R_d)
            ghost_exchange(self.state_auxiliary)
            gradp = self.reference_pressure_gradient()
            for g, aux, gp in zip(self.grids, self.state_auxiliary, gradp):
                a = _sv(aux.data)
                nr = g.nreal
                k = a[bl.a_gradΦ][:, :nr] / ps.grav
                gΦ = a[bl.a_gradΦ][:, :nr]
                gpr = _sv(gp.data)[:, :nr]
                num = -(k[0] * gpr[0] + k[1] * gpr[1] + k[2] * gpr[2])
                den = k[0] * gΦ[0] + k[1] * gΦ[1] + k[2] * gΦ[2]
                a[bl.a_ref["ρ"]][:nr] = num / den
            ghost_exchange(self.state_auxiliary)
            for g, aux in zip(self.grids, self.state_auxiliary):
                a = _sv(aux.data)
                nr = g.nreal
                ρ = a[bl.a_ref["ρ"]][:nr]
                p = a[bl.a_ref["p"]][:nr]
                T = p / (ρ * ps.R_d)  # PhaseDry_ρp -> air_temperature
                a[bl.a_ref["T"]][:nr] = T
                e_pot = a[bl.a_Φ][:nr]
                a[bl.a_ref["ρe"]][:nr] = ρ * A.total_energy(ps, bl.FT(0), e_pot, T)
        # atmos_nodal_init_state_auxiliary!: coord, turbulence Delta
        for g, aux in zip(self.grids, self.state_auxiliary):
            a = _sv(aux.data)
            nr = g.nreal
            a[0][:nr] = g.vgeo[:nr, G._x1]
            a[1][:nr] = g.vgeo[:nr, G._x2]
            a[2][:nr] = g.vgeo[:nr, G._x3]
            if bl.a_Δ is not None:
                vg = g.vgeo[:nr]
                det = (vg[:, 0] * (vg[:, 4] * vg[:, 8] - vg[:, 7] * vg[:, 5])
                       - vg[:, 3] * (vg[:, 1] * vg[:, 8] - vg[:, 7] * vg[:, 2])
                       + vg[:, 6] * (vg[:, 1] * vg[:, 5] - vg[:, 4] * vg[:, 2]))
                a[bl.a_Δ][:nr] = 2 / (np.cbrt(det) * max(1, *g.N))
            if getattr(bl, "a_Δh", None) is not None:
                a[bl.a_Δh][:nr] = lengthscale_horizontal(g)
            if getattr(bl, "a_δχ", None) is not None:
                for i, δ in enumerate(bl.tracers):      # atmos_init_aux!(::NTracers) (tracers.jl:133-140)
                    a[bl.a_δχ][i][:nr] = δ
        ghost_exchange(self.state_auxiliary)

    def reference_pressure_gradient(self):
        """``grad p_ref`` by a DGModel over ``PressureGradientModel`` with central
        fluxes (``ref_state.jl:204-264``): flux = -p I, boundary faces see p+ = p-."""
        outs = []
        bl = self.bl
        for g, aux in zip(self.grids, self.state_auxiliary):
            outs.append(MPIStateArray.from_grid(g, 3))
        for g, aux, out in zip(self.grids, self.state_auxiliary, outs):
            p = aux.data[:, bl.a_ref["p"], :]  # (nelem, Np), ghosts valid at face nodes
            nr = g.nreal

            def flux(pv):
                F = np.zeros((3, 3) + pv.shape, dtype=pv.dtype)
                for d in range(3):
                    F[d, d] = -pv
                return F

            T = self._volume_weak_divergence(g, flux(p[:nr]), None)
            dq = _sv(out.data)
            dq[:, :nr] = T
            pflat = p.reshape(-1)
            for f in range(6):
                e = np.arange(nr)
                idm = g.vmapM[:nr, f] - 1
                idp = g.vmapP[:nr, f] - 1
                bnd = g.elemtobndy[:nr, f] != 0
                idp = np.where(bnd[:, None], idm, idp)
                pm, pp = pflat[idm], pflat[idp]
                n = [g.sgeo[:nr, f, :, c] for c in range(3)]
                sM, vMI = g.sgeo[:nr, f, :, G._sM], g.sgeo[:nr, f, :, G._vMI]
                vid = idm - g.Np * e[:, None]
                for s in range(3):
                    fl = ((-pm) + (-pp)) * (n[s] / 2)
                    cur = out.data[e[:, None], s, vid]
                    out.data[e[:, None], s, vid] = cur - vMI * sM * fl
        return outs

    def update_auxiliary_state(self, Q, elems="real"):
        bl = self.bl
        if hasattr(bl, "update_auxiliary_state"):
            bl.update_auxiliary_state(self, Q, elems)
            return
        for g, q, aux in zip(self.grids, Q, self.state_auxiliary):
            sl = slice(0, g.nreal) if elems == "real" else slice(g.nreal, g.nelem)
            qs = _sv(q.data[sl])
            a = _sv(aux.data[sl])
            with np.errstate(all="ignore"):
                bl.nodal_update_aux(qs, a)

    # ------------------------------------------------------------------
    # volume kernels
    # ------------------------------------------------------------------
    def _volume_weak_divergence(self, g, F, source):
        """H launch then V launch of ``volume_tendency!`` on real elements.

        ``F``: (3, S, nreal, Np) physical fluxes; ``source``: (S, nreal, Np) or None.
        Returns the local tendency (S, nreal, Np) (alpha = 1, beta = 0), accumulated in
        the reference's order (H: n-loop alternating xi1/xi2; V: k-loop, source
        after the k == n term)."""
        nr = F.shape[2]
        S = F.shape[1]
        Nq = g.Nq
        vg = g.vgeo[:nr]
        M, MI = vg[:, G._M], vg[:, G._MI]
        Ft = []
        for m in range(3):
            a, b, c = vg[:, G._xi1x1 + m], vg[:, G._xi1x2 + m], vg[:, G._xi1x3 + m]
            Ft.append(M * (a * F[0] + b * F[1] + c * F[2]))
        shp = (S, nr, Nq[2], Nq[1], Nq[0])
        F1, F2, F3 = [x.reshape(shp) for x in Ft]
        MIr = MI.reshape(nr, Nq[2], Nq[1], Nq[0])
        D1, D2, D3 = g.D
        ltH = np.zeros(shp, dtype=F.dtype)
        for n in range(Nq[0]):
            ltH = ltH + (MIr * D1[n, :]) * F1[..., n:n + 1]
            ltH = ltH + (MIr * D2[n, :][:, None]) * F2[:, :, :, n:n + 1, :]
        ltV = np.zeros(shp, dtype=F.dtype)
        src = None if source is None else source.reshape(shp)
        for k in range(Nq[2]):
            ltV = ltV + (MIr * D3[k, :][:, None, None]) * F3[:, :, k:k + 1]
            if src is not None:
                ltV[:, :, k] = ltV[:, :, k] + src[:, :, k]
        self._last_H, self._last_V = ltH.reshape(S, nr, g.Np), ltV.reshape(S, nr, g.Np)
        return self._last_H + self._last_V

    def volume_tendency(self, tendency, Q, t, α, β):
        bl = self.bl
        for g, dq, q, aux, gf in zip(self.grids, tendency, Q, self.state_auxiliary,
                                     self.state_gradient_flux):
            nr = g.nreal
            qs, a = _sv(q.data[:nr]), _sv(aux.data[:nr])
            F = bl.flux_first_order(qs, a)
            if self._second():
                F = F + bl.flux_second_order(qs, _sv(gf.data[:nr]), a)
                if self.nhyper:
                    hg = self.Qhypervisc_grad[self.grids.index(g)]
                    F = F + bl.flux_hyperdiffusive(qs, _sv(hg.data[:nr]))
            src = bl.source(qs, a)
            self._volume_weak_divergence(g, F, src)
            d = _sv(dq.data[:nr])
            # H launch: dQ = alpha*T_H + beta*dQ ; V launch: dQ = alpha*T_V + 1*dQ
            if β != 0:
                d[...] = α * self._last_H + β * d
            else:
                d[...] = α * self._last_H
            d[...] = α * self._last_V + 1 * d

    def _skip2(self):
        return self.skip_zero_viscosity and not self.bl.viscous()

    def _second(self):
        """The gradient / second-order machinery runs when the law has gradient-flux or hyperdiffusive
        variables (DGModel.jl:118-123: ``num_state_gradient_flux > 0 || nhyperviscstate > 0``)."""
        return (self.bl.GF > 0 or self.nhyper > 0) and not self._skip2()

    # ------------------------------------------------------------------
    # numerical fluxes
    # ------------------------------------------------------------------
    def numerical_flux_first_order(self, n, Qm, am, Qp, ap):
        bl = self.bl
        Fm = bl.flux_first_order(Qm, am)
        Fp = bl.flux_first_order(Qp, ap)
        Fs = Fm + Fp
        fl = Fs[0] * (n[0] / 2) + Fs[1] * (n[1] / 2) + Fs[2] * (n[2] / 2)
        if self.nf1 == "central":
            return fl
        if self.nf1 == "rusanov":
            λ = np.maximum(bl.wavespeed(n, Qm, am), bl.wavespeed(n, Qp, ap))
            penalty = λ * (Qm - Qp)
            if hasattr(bl, "update_penalty"):
                bl.update_penalty(penalty)
            return fl + penalty / 2
        if self.nf1 == "roe":
            return fl - bl.roe_dissipation(n, Qm, am, Qp, ap)
        raise ValueError(self.nf1)

    # ------------------------------------------------------------------
    # interface kernels
    # ------------------------------------------------------------------
    def _face_data(self, g, elems, f, arr):
        """Minus/plus gathers of ``arr`` (nelem, S, Np) on face ``f`` of ``elems`` (0-based)."""
        idm = g.vmapM[elems, f] - 1
        idp = g.vmapP[elems, f] - 1
        bnd = g.elemtobndy[elems, f]
        idp = np.where((bnd != 0)[:, None], idm, idp)
        em, vm = np.divmod(idm, g.Np)
        ep, vp = np.divmod(idp, g.Np)
        return em, vm, ep, vp, bnd

    def interface_tendency(self, tendency, Q, t, α, which):
        bl = self.bl
        for g, dq, q, aux, gf in zip(self.grids, tendency, Q, self.state_auxiliary,
                                     self.state_gradient_flux):
            elems = (g.interiorelems if which == "interior" else g.exteriorelems) - 1
            if len(elems) == 0:
                continue
            second = self._second()
            for f in range(6):
                em, vm, ep, vp, bnd = self._face_data(g, elems, f, q.data)
                n = np.stack([g.sgeo[elems, f, :, c] for c in range(3)])
                sM, vMI = g.sgeo[elems, f, :, G._sM], g.sgeo[elems, f, :, G._vMI]
                Qm = np.moveaxis(q.data[em, :, vm], -1, 0)      # (S, ne, Nfp)
                am = np.moveaxis(aux.data[em, :, vm], -1, 0)
                Qp = np.moveaxis(q.data[ep, :, vp], -1, 0).copy()
                ap = np.moveaxis(aux.data[ep, :, vp], -1, 0).copy()
                isb = bnd != 0
                for tag in np.unique(bnd[isb]):
                    m = bnd == tag
                    Qb, ab = bl.boundary_state("first", int(tag), n[:, m], Qm[:, m], am[:, m])
                    Qp[:, m], ap[:, m] = Qb, ab
                fl = self.numerical_flux_first_order(n, Qm, am, Qp, ap)
                if second:
                    gm = np.moveaxis(gf.data[em, :, vm], -1, 0)
                    gp = np.moveaxis(gf.data[ep, :, vp], -1, 0)
                    F2 = bl.flux_second_order(Qm, gm, am)
                    # second-order flux on the + side uses the *un-modified* + state
                    Qp2 = np.moveaxis(q.data[ep, :, vp], -1, 0)
                    ap2 = np.moveaxis(aux.data[ep, :, vp], -1, 0)
                    F2 = F2 + bl.flux_second_order(Qp2, gp, ap2)
                    if self.nhyper:
                        hg = self.Qhypervisc_grad[self.grids.index(g)]
                        F2 = F2 + bl.flux_hyperdiffusive(Qm, np.moveaxis(hg.data[em, :, vm], -1, 0))
                        F2 = F2 + bl.flux_hyperdiffusive(Qp2, np.moveaxis(hg.data[ep, :, vp], -1, 0))
                    fl2 = F2[0] * (n[0] / 2) + F2[1] * (n[1] / 2) + F2[2] * (n[2] / 2)
                    # AtmosBC FreeSlip/NoSlip + Insulating: no diffusive boundary flux; models with
                    # flux-based BCs (normal_boundary_flux_second_order!, NumericalFluxes.jl:872-967)
                    # provide the full plus-side flux
                    fl2[:, isb] = 0
                    if hasattr(bl, "boundary_flux_second_order"):
                        for tag in np.unique(bnd[isb]):
                            m = bnd == tag
                            Fb = bl.boundary_flux_second_order(int(tag), n[:, m], Qm[:, m], gm[:, m], am[:, m])
                            fl2[:, m] = Fb[0] * n[0][m] + Fb[1] * n[1][m] + Fb[2] * n[2][m]
                    fl = fl + fl2
                for s in range(bl.S):
                    cur = dq.data[em, s, vm]
                    dq.data[em, s, vm] = cur - α * vMI * sM * fl[s]

    # ------------------------------------------------------------------
    # gradient pass
    # ------------------------------------------------------------------
    def volume_gradients(self, Q, t):
        bl = self.bl
        for g, q, aux, gf in zip(self.grids, Q, self.state_auxiliary, self.state_gradient_flux):
            nr = g.nreal
            Nq = g.Nq
            qs, a = _sv(q.data[:nr]), _sv(aux.data[:nr])
            Gt = bl.gradient_argument(qs, a)
            shp = (bl.G, nr, Nq[2], Nq[1], Nq[0])
            Gr = Gt.reshape(shp)
            D1, D2, D3 = g.D
            G1 = np.zeros(shp, dtype=q.data.dtype)
            G2 = np.zeros(shp, dtype=q.data.dtype)
            for n in range(Nq[0]):
                G1 = G1 + D1[:, n] * Gr[..., n:n + 1]
                G2 = G2 + D2[:, n][:, None] * Gr[:, :, :, n:n + 1, :]
            G3 = np.zeros(shp, dtype=q.data.dtype)
            for k in range(Nq[2]):
                G3 = G3 + D3[:, k][:, None, None] * Gr[:, :, k:k + 1]
            G1, G2, G3 = [x.reshape(bl.G, nr, g.Np) for x in (G1, G2, G3)]
            vg = g.vgeo[:nr]
            gradH = np.stack([vg[:, G._xi1x1 + 3 * d] * G1 + vg[:, G._xi2x1 + 3 * d] * G2
                              for d in range(3)])
            gfs = _sv(gf.data[:nr])
            gfs[...] = bl.gradient_flux(gradH, qs, a)
            if self.diffusion_direction == "every":
                gradV = np.stack([vg[:, G._xi3x1 + 3 * d] * G3 for d in range(3)])
                gfs[...] = gfs + bl.gradient_flux(gradV, qs, a)
            if self.nhyper:
                hg = _sv(self.Qhypervisc_grad[self.grids.index(g)].data[:nr])
                for s_ in range(bl.ngradlap):
                    for d in range(3):
                        hg[3 * s_ + d] = gradH[d, bl.hyper_G + s_]

    def interface_gradients(self, Q, t, which):
        bl = self.bl
        faces = range(6) if self.diffusion_direction == "every" else range(4)
        for g, q, aux, gf in zip(self.grids, Q, self.state_auxiliary, self.state_gradient_flux):
            elems = (g.interiorelems if which == "interior" else g.exteriorelems) - 1
            if len(elems) == 0:
                continue
            for f in faces:
                em, vm, ep, vp, bnd = self._face_data(g, elems, f, q.data)
                n = np.stack([g.sgeo[elems, f, :, c] for c in range(3)])
                sM, vMI = g.sgeo[elems, f, :, G._sM], g.sgeo[elems, f, :, G._vMI]
                Qm = np.moveaxis(q.data[em, :, vm], -1, 0)
                am = np.moveaxis(aux.data[em, :, vm], -1, 0)
                Qp = np.moveaxis(q.data[ep, :, vp], -1, 0).copy()
                ap = np.moveaxis(aux.data[ep, :, vp], -1, 0).copy()
                Gm = bl.gradient_argument(Qm, am)
                Gp = bl.gradient_argument(Qp, ap)
                Gstar = (Gp + Gm) / 2
                isb = bnd != 0
                for tag in np.unique(bnd[isb]):
                    m = bnd == tag
                    Qb, ab = bl.boundary_state("gradient", int(tag), n[:, m], Qm[:, m], am[:, m])
                    Gstar[:, m] = bl.gradient_argument(Qb, ab)
                nGstar = np.stack([n[d] * Gstar for d in range(3)])
                nGm = np.stack([n[d] * Gm for d in range(3)])
                gfstar = bl.gradient_flux(nGstar, Qm, am)
                gfm = bl.gradient_flux(nGm, Qm, am)
                for s in range(bl.GF):
                    cur = gf.data[em, s, vm]
                    gf.data[em, s, vm] = cur + vMI * sM * (gfstar[s] - gfm[s])
                if self.nhyper:
                    hg = self.Qhypervisc_grad[self.grids.index(g)]
                    for s_ in range(bl.ngradlap):
                        j = bl.hyper_G + s_
                        for d in range(3):
                            cur = hg.data[em, 3 * s_ + d, vm]
                            hg.data[em, 3 * s_ + d, vm] = cur + vMI * sM * (nGstar[d, j] - nGm[d, j])

    # ------------------------------------------------------------------
    # hyperdiffusion passes (HorizontalDirection)
    # ------------------------------------------------------------------
    def volume_divergence_of_gradients(self):
        """Qhypervisc_div[s] = -MI D^T (M xi_x . grad G_s), xi1 and xi2 only."""
        bl = self.bl
        for g, hg, hd in zip(self.grids, self.Qhypervisc_grad, self.Qhypervisc_div):
            nr, Nq = g.nreal, g.Nq
            vg = g.vgeo[:nr]
            M, MI = vg[:, G._M], vg[:, G._MI]
            gr = _sv(hg.data[:nr])
            shp = (nr, Nq[2], Nq[1], Nq[0])
            MIr = MI.reshape(shp)
            D1, D2, _ = g.D
            for s_ in range(bl.ngradlap):
                G1, G2, G3 = gr[3 * s_], gr[3 * s_ + 1], gr[3 * s_ + 2]
                s1 = (M * (vg[:, G._xi1x1] * G1 + vg[:, G._xi1x2] * G2 + vg[:, G._xi1x3] * G3)).reshape(shp)
                s2 = (M * (vg[:, G._xi2x1] * G1 + vg[:, G._xi2x2] * G2 + vg[:, G._xi2x3] * G3)).reshape(shp)
                div = np.zeros(shp, dtype=hg.data.dtype)
                for n in range(Nq[0]):
                    div = div - MIr * D1[n, :] * s1[..., n:n + 1]
                    div = div - MIr * D2[n, :][:, None] * s2[:, :, n:n + 1, :]
                hd.data[:nr, s_] = div.reshape(nr, g.Np)

    def interface_divergence_of_gradients(self, which):
        """+= vMI sM (grad+ + grad-)' n / 2; walls: grad+ = grad- (boundary_state! is a no-op)."""
        bl = self.bl
        for g, hg, hd in zip(self.grids, self.Qhypervisc_grad, self.Qhypervisc_div):
            elems = (g.interiorelems if which == "interior" else g.exteriorelems) - 1
            if len(elems) == 0:
                continue
            for f in range(4):
                em, vm, ep, vp, bnd = self._face_data(g, elems, f, hg.data)
                n = np.stack([g.sgeo[elems, f, :, c] for c in range(3)])
                sM, vMI = g.sgeo[elems, f, :, G._sM], g.sgeo[elems, f, :, G._vMI]
                gm = np.moveaxis(hg.data[em, :, vm], -1, 0)
                gp = np.moveaxis(hg.data[ep, :, vp], -1, 0)
                for s_ in range(bl.ngradlap):
                    ldiv = 0
                    for d in range(3):
                        ldiv = ldiv + (gp[3 * s_ + d] + gm[3 * s_ + d]) * (n[d] / 2)
                    cur = hd.data[em, s_, vm]
                    hd.data[em, s_, vm] = cur + vMI * sM * ldiv

    def volume_gradients_of_laplacians(self, Q):
        """Qhypervisc_grad = transform(xi_x D lap), xi1 and xi2 only (strong form)."""
        bl = self.bl
        for g, q, aux, hg, hd in zip(self.grids, Q, self.state_auxiliary, self.Qhypervisc_grad,
                                     self.Qhypervisc_div):
            nr, Nq = g.nreal, g.Nq
            vg = g.vgeo[:nr]
            shp = (bl.ngradlap, nr, Nq[2], Nq[1], Nq[0])
            lap = _sv(hd.data[:nr]).reshape(shp)
            D1, D2, _ = g.D
            L1 = np.zeros(shp, dtype=hd.data.dtype)
            L2 = np.zeros(shp, dtype=hd.data.dtype)
            for n in range(Nq[0]):
                L1 = L1 + D1[:, n] * lap[..., n:n + 1]
                L2 = L2 + D2[:, n][:, None] * lap[:, :, :, n:n + 1, :]
            L1, L2 = L1.reshape(bl.ngradlap, nr, g.Np), L2.reshape(bl.ngradlap, nr, g.Np)
            gl = np.stack([vg[:, G._xi1x1 + 3 * d] * L1 + vg[:, G._xi2x1 + 3 * d] * L2 for d in range(3)])
            H = bl.transform_post_gradient_laplacian(gl, _sv(q.data[:nr]), _sv(aux.data[:nr]))
            _sv(hg.data[:nr])[...] = H

    def interface_gradients_of_laplacians(self, Q, which):
        """+= vMI sM transform(n (lap+ - lap-) / 2) with the minus-side state; walls: lap+ = lap-."""
        bl = self.bl
        for g, q, aux, hg, hd in zip(self.grids, Q, self.state_auxiliary, self.Qhypervisc_grad,
                                     self.Qhypervisc_div):
            elems = (g.interiorelems if which == "interior" else g.exteriorelems) - 1
            if len(elems) == 0:
                continue
            for f in range(4):
                em, vm, ep, vp, bnd = self._face_data(g, elems, f, hd.data)
                n = np.stack([g.sgeo[elems, f, :, c] for c in range(3)])
                sM, vMI = g.sgeo[elems, f, :, G._sM], g.sgeo[elems, f, :, G._vMI]
                lm = np.moveaxis(hd.data[em, :, vm], -1, 0)
                lp = np.moveaxis(hd.data[ep, :, vp], -1, 0)
                Qm = np.moveaxis(q.data[em, :, vm], -1, 0)
                am = np.moveaxis(aux.data[em, :, vm], -1, 0)
                Gn = np.stack([n[d] * (lp - lm) / 2 for d in range(3)])
                H = bl.transform_post_gradient_laplacian(Gn, Qm, am)
                for s_ in range(bl.nhyper):
                    cur = hg.data[em, s_, vm]
                    hg.data[em, s_, vm] = cur + vMI * sM * H[s_]

    # ------------------------------------------------------------------
    # the tendency functor
    # ------------------------------------------------------------------
    def __call__(self, tendency, Q, t=0.0, α=1, β=0, increment=None):
        if increment is not None:
            α, β = 1, (1 if increment else 0)
        single = not isinstance(Q, (list, tuple))
        if single:
            tendency, Q = [tendency], [Q]
        bl = self.bl
        bl.t = t                                # time-dependent sources / boundary states (InitStateBC)
        self.update_auxiliary_state(Q, "real")
        ghost_exchange(Q)                       # begin_ghost_exchange!(Q)
        second = self._second()
        if second:
            self.volume_gradients(Q, t)
            self.interface_gradients(Q, t, "interior")
            self.update_auxiliary_state(Q, "ghost")   # after end_ghost_exchange!(Q)
            self.interface_gradients(Q, t, "exterior")
            ghost_exchange(self.state_gradient_flux)
            if self.nhyper:
                ghost_exchange(self.Qhypervisc_grad)
            if hasattr(bl, "update_auxiliary_state_gradient"):
                bl.update_auxiliary_state_gradient(self, Q, "real")
            if self.nhyper:
                self.volume_divergence_of_gradients()
                self.interface_divergence_of_gradients("interior")
                self.interface_divergence_of_gradients("exterior")
                ghost_exchange(self.Qhypervisc_div)
                self.volume_gradients_of_laplacians(Q)
                self.interface_gradients_of_laplacians(Q, "interior")
                self.interface_gradients_of_laplacians(Q, "exterior")
                ghost_exchange(self.Qhypervisc_grad)
        self.volume_tendency(tendency, Q, t, α, β)
        self.interface_tendency(tendency, Q, t, α, "interior")
        if second and hasattr(bl, "update_auxiliary_state_gradient"):
            bl.update_auxiliary_state_gradient(self, Q, "ghost")
        if not second:
            self.update_auxiliary_state(Q, "ghost")
        self.interface_tendency(tendency, Q, t, α, "exterior")


def lengthscale_horizontal(g):
    """``lengthscale_horizontal`` (src/Numerics/Mesh/Geometry.jl:129-152) at every node of the real
    elements: mean of |J e1| and |J e2| times 2 / N, J = (d xi / d x)^-1."""
    nr = g.nreal
    vg = g.vgeo[:nr]
    invJ = np.zeros((nr, g.Np, 3, 3), dtype=g.vgeo.dtype)
    for i in range(3):
        for j in range(3):
            invJ[..., i, j] = vg[:, G._xi1x1 + 3 * j + i]      # d xi_{i+1} / d x_{j+1} (Grids.jl:76-92)
    e = np.zeros((nr, g.Np, 3, 2), dtype=g.vgeo.dtype)
    e[..., 0, 0] = 1
    e[..., 1, 1] = 1
    sol = np.linalg.solve(invJ, e)
    Δ1 = np.sqrt((sol[..., 0] ** 2).sum(-1)) * 2 / g.N[0]
    Δ2 = np.sqrt((sol[..., 1] ** 2).sum(-1)) * 2 / g.N[1]
    return (Δ1 + Δ2) / 2


def init_ode_state(dg, init_fn, t=0.0):
    """``init_ode_state`` (``SpaceDiscretization.jl``/``DGModel.jl``): nodal ICs on real
    elements followed by a ghost exchange.  ``init_fn(x1, x2, x3, aux, t)`` -> (S, nreal, Np)."""
    Qs = []
    for g, aux in zip(dg.grids, dg.state_auxiliary):
        q = MPIStateArray.from_grid(g, dg.bl.S)
        nr = g.nreal
        a = _sv(aux.data[:nr])
        vals = init_fn(g.vgeo[:nr, G._x1], g.vgeo[:nr, G._x2], g.vgeo[:nr, G._x3], a, t)
        _sv(q.data[:nr])[...] = vals
        Qs.append(q)
    ghost_exchange(Qs)
    return Qs
