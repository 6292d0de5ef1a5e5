"""Pins the oracle's spectral filters on the reference's own tests
(test/Numerics/Mesh/filter.jl): golden ExponentialFilter matrices (:15-75, committed as
tests/golden/filter_matrices.json) and the analytic cutoff-filter application test (:161-246)."""
import json
import os

import numpy as np
import pytest

from oracle import filters, grids, topologies as tp, elements

HERE = os.path.dirname(os.path.abspath(__file__))


def test_exponential_filter_matrices_match_reference_golden():
    gold = json.load(open(os.path.join(HERE, "golden", "filter_matrices.json")))
    for name, g in gold.items():
        r, _ = elements.lglpoints(np.float64, g["N"])
        W = filters.exponential_filter_matrix(r, g["Nc"], g["s"])
        Wg = np.array(g["W"])
        assert Wg.shape == (g["N"] + 1, g["N"] + 1)
        assert np.allclose(W, Wg, rtol=1e-8, atol=1e-9), name   # the reference's own gate is `≈`
        assert np.max(np.abs(W - Wg)) < 2e-9, name


# Legendre polynomials and the low/high split of filter.jl:135-152
def l1(r): return r
def l2(r): return (3 * r ** 2 - 1) / 2
def l3(r): return (5 * r ** 3 - 3 * r) / 2
def low(x, y, z): return 1 + 4 * l1(x) * l1(y) + 5 * l1(z) + 6 * l1(z) * l1(x)
def high(x, y, z): return l2(x) * l3(y) + l3(x) + l2(y) + l3(z) * l1(y)
FILTERED = {"every": high,
            "vertical": lambda x, y, z: l3(z) * l1(y),
            "horizontal": lambda x, y, z: l2(x) * l3(y) + l3(x) + l2(y)}


@pytest.mark.parametrize("N", [3, 4])
@pytest.mark.parametrize("direction", ["every", "horizontal", "vertical"])
def test_cutoff_filter_application_analytic(N, direction):
    """filter.jl:161-246 (dim = 3): states 1 and 3 are filtered with CutoffFilter(grid, 2), states
    2 and 4 left alone; the filtered states lose exactly the `filtered(direction)` part."""
    br = tuple(np.linspace(-1.0, 1.0, 2) for _ in range(3))
    topo = tp.BrickTopology(1, br, periodicity=(True, True, True))[0]
    g = grids.Grid(topo, N)
    x, y, z = (g.vgeo[:, c, :] for c in (grids._x1, grids._x2, grids._x3))
    full = low(x, y, z) + high(x, y, z)
    Q = np.stack([full, full, full, full], axis=1)
    W = filters.cutoff_filter_matrix(g.xi[0], 2)
    filters.apply(Q, filters.FilterIndices(0, 2), g, W, direction=direction)
    P = np.stack([full - FILTERED[direction](x, y, z), full, full - FILTERED[direction](x, y, z), full], axis=1)
    assert np.allclose(Q, P, rtol=1e-12, atol=1e-12)


def test_atmos_perturbation_target_preserves_reference_state():
    """AtmosFilterPerturbations: a state equal to the reference state is a fixed point, and the
    cell averages of the perturbations are preserved (V diag(1, ...) V^-1 keeps mode 0)."""
    from oracle import atmos, dgmodel as odg
    from tests import parity
    model, gs = parity.gcm_setup(ne=2, nvert=2)
    g = gs[0]
    dgm = odg.DGModel(model, [g], "rusanov")
    aux = dgm.state_auxiliary[0].data
    Q = np.zeros((g.nelem, 5, g.Np))
    Q[:, 0] = aux[:, model.a_ref["ρ"]]
    Q[:, 4] = aux[:, model.a_ref["ρe"]]
    Q0 = Q.copy()
    W = filters.exponential_filter_matrix(g.xi[0], 0, 10)
    filters.apply(Q, filters.AtmosFilterPerturbations(model), g, W, state_auxiliary=aux)
    assert np.allclose(Q, Q0, rtol=1e-14, atol=0)
    rng = np.random.default_rng(0)
    Q = Q0 + 1e-3 * Q0 * rng.standard_normal(Q0.shape)
    Q[:, 1:4] = rng.standard_normal((g.nelem, 3, g.Np))
    Q1 = Q.copy()
    filters.apply(Q1, filters.AtmosFilterPerturbations(model), g, W, state_auxiliary=aux)
    assert not np.allclose(Q1, Q)
    wq = np.kron(np.kron(g.w[2], g.w[1]), g.w[0])      # reference-element quadrature weights
    assert np.allclose(((Q1 - Q)[:g.nreal] * wq).sum(-1), 0, atol=1e-10 * np.abs(Q).max())
