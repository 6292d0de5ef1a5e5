"""Known-answer checks of the oracle's HeldSuarezForcing / RayleighSponge restatement
(experiments/AtmosGCM/heldsuarez.jl:112-172, src/Atmos/Model/tendencies_momentum.jl:104-137).
The reference holds no golden vector for these sources ("parity unpinned"); the values below are
the closed-form Held & Suarez (1994) numbers the reference code encodes."""
import numpy as np

from oracle import atmos


def _model(sources):
    return atmos.DryAtmosModel(np.float64, orientation="spherical",
                               ref_state=dict(T_surf=290.0, T_min=220.0, H_t=8e3, subtract_off=True),
                               turbulence=("smagorinsky", 0.21), sources=sources,
                               bcs=("freeslip", "freeslip"))


def _state(m, x, z, T, p, u=(0.0, 0.0, 0.0)):
    ps = m.ps
    aux = np.zeros((m.A, 1))
    r = np.linalg.norm(x)
    xh = np.asarray(x) / r
    aux[0:3, 0] = xh * (ps.planet_radius + z)
    aux[m.a_Φ, 0] = ps.grav * z
    aux[m.a_gradΦ, 0] = ps.grav * xh
    ρ = p / (ps.R_d * T)
    e = 0.5 * sum(v * v for v in u) + ps.grav * z + ps.cv_d * (T - ps.T_0)
    Q = np.array([[ρ], [ρ * u[0]], [ρ * u[1]], [ρ * u[2]], [ρ * e]])
    return Q, aux


def test_held_suarez_coefficients_known_answers():
    m = _model(("held_suarez",))
    day = 86400.0
    # equator, surface pressure: T_eq = 315 K, k_T = k_s = 1/(4 day), k_v = k_f = 1/day
    Q, aux = _state(m, (1.0, 0.0, 0.0), 0.0, 300.0, 1.01325e5)
    k_v, k_T, T_eq = m.held_suarez_coefficients(Q, aux)
    assert np.allclose([k_v[0], k_T[0], T_eq[0]], [1 / day, 1 / (4 * day), 315.0], rtol=1e-13)
    # pole, surface pressure: T_eq = 315 - 60 = 255 K, k_T = k_a = 1/(40 day)
    Q, aux = _state(m, (0.0, 0.0, 1.0), 0.0, 250.0, 1.01325e5)
    k_v, k_T, T_eq = m.held_suarez_coefficients(Q, aux)
    assert np.allclose([k_T[0], T_eq[0]], [1 / (40 * day), 255.0], rtol=1e-12)
    # sigma = 0.5 at 45 degrees: no boundary layer, T_eq = (315 - 30 + 10 ln2 / 2) 0.5^(2/7)
    Q, aux = _state(m, (1.0, 0.0, 1.0), 5e3, 250.0, 0.5 * 1.01325e5)
    k_v, k_T, T_eq = m.held_suarez_coefficients(Q, aux)
    assert k_v[0] == 0 and np.isclose(k_T[0], 1 / (40 * day), rtol=1e-13)
    assert np.isclose(T_eq[0], (315 - 30 + 10 * np.log(2) * 0.5) * 0.5 ** (2 / 7), rtol=1e-13)
    # stratosphere floor
    Q, aux = _state(m, (1.0, 0.0, 0.0), 25e3, 210.0, 0.02 * 1.01325e5)
    assert m.held_suarez_coefficients(Q, aux)[2][0] == 200.0


def test_held_suarez_and_sponge_sources():
    day = 86400.0
    sp = ("rayleigh_sponge", 30e3, 12e3, 1 / 900, (0.0, 0.0, 0.0), 2.0)
    m = _model(("held_suarez", sp))
    # surface equator, wind (radial 1, tangential 3, 4): friction removes tangential momentum only
    Q, aux = _state(m, (1.0, 0.0, 0.0), 0.0, 300.0, 1.01325e5, u=(1.0, 3.0, 4.0))
    S = m.source(Q, aux)
    ρ = Q[0, 0]
    assert np.allclose(S[1:4, 0], [0.0, -ρ * 3 / day, -ρ * 4 / day], rtol=1e-12, atol=1e-18)
    assert np.isclose(S[4, 0], -(1 / (4 * day)) * ρ * m.ps.cv_d * (300.0 - 315.0), rtol=1e-12)
    # mid-sponge (z = 21 km, r = 1/2): beta = alpha sin(pi/4)^2 = alpha / 2
    Q, aux = _state(m, (0.0, 1.0, 0.0), 21e3, 220.0, 0.05 * 1.01325e5, u=(2.0, 0.0, -1.0))
    S = m.source(Q, aux)
    assert np.allclose(S[1:4, 0], -(0.5 / 900) * Q[1:4, 0], rtol=1e-12)
    # below the sponge and above the boundary layer: nothing acts on momentum
    Q, aux = _state(m, (0.0, 1.0, 0.0), 8e3, 240.0, 0.35 * 1.01325e5, u=(2.0, 0.0, -1.0))
    assert np.all(m.source(Q, aux)[1:4] == 0)
