"""Spectral filters applied to a state array (test infrastructure -- see oracle/__init__.py).

Restates ``src/Numerics/Mesh/Filters.jl``:

* filter matrices ``spectral_filter_matrix`` (``:114-131``), ``ExponentialFilter`` (``:172-229``),
  ``CutoffFilter`` (``:275-314``) -- the matrices themselves live in ``oracle/ocean.py`` (the ocean
  model uses them inside its tendency) and are re-exported here;
* ``Filters.apply!`` / ``apply_async!`` (``:440-505``): ``EveryDirection`` = one *horizontal*
  kernel launch (xi1 then xi2 with ``filter_matrices[1]``) followed by one *vertical* launch
  (xi3 with ``filter_matrices[end]``); each launch evaluates ``compute_filter_argument!`` /
  ``compute_filter_result!`` of the target;
* ``kernel_apply_filter!`` (``:651-792``): ``out[i] = sum_n W[i, n] in[n]`` accumulated n = 1..Nq;
* targets ``FilterIndices`` (``:60-98``) and ``AtmosFilterPerturbations``
  (``src/Atmos/Model/filters.jl:4-48``: filter ``rho - rho_ref`` and ``rhoe - rhoe_ref``, momentum as is).

GaussQuadrature.jl 0.5.5 (un-vendored, ``Manifest.toml:357-361``) supplies the orthonormal Legendre
Vandermonde; the restatement is pinned on the reference's golden filter matrices
(``test/Numerics/Mesh/filter.jl:15-75``, tests/test_oracle_filters.py) and its analytic
application test (``:161-246``).
"""
import numpy as np

from .ocean import spectral_filter_matrix, cutoff_filter_matrix, exponential_filter_matrix  # noqa: F401


class FilterIndices:
    """``FilterIndices(I...)`` with 0-based state indices."""

    def __init__(self, *I):
        self.I = tuple(int(i) for i in I)

    def argument(self, Q, aux):
        return np.stack([Q[i] for i in self.I])

    def result(self, Q, F, aux):
        out = Q.copy()
        for n, i in enumerate(self.I):
            out[i] = F[n]
        return out


class AtmosFilterPerturbations:
    """Dry branch of ``src/Atmos/Model/filters.jl:4-48``; ``model`` is an oracle DryAtmosModel."""

    def __init__(self, model):
        self.iρ, self.iρe = model.a_ref["ρ"], model.a_ref["ρe"]

    def argument(self, Q, aux):
        F = Q.copy()
        F[0] = F[0] - aux[self.iρ]
        F[4] = F[4] - aux[self.iρe]
        return F

    def result(self, Q, F, aux):
        out = F.copy()
        out[0] = out[0] + aux[self.iρ]
        out[4] = out[4] + aux[self.iρe]
        return out


def _contract(F, W, axis):
    """out[.., i, ..] = sum_n W[i, n] F[.., n, ..] along ``axis``, accumulated n = 0..Nq-1."""
    Fm = np.moveaxis(F, axis, -1)
    out = np.zeros_like(Fm)
    for n in range(W.shape[1]):
        out = out + W[:, n] * Fm[..., n:n + 1]
    return np.moveaxis(out, -1, axis)


def _launch(g, data, aux, target, W, dirs):
    """One ``kernel_apply_filter!`` launch over the real elements; ``dirs``: reference directions."""
    nr, Nq = g.nreal, g.Nq
    S = data.shape[1]
    Q = np.moveaxis(data[:nr], 1, 0).reshape(S, nr, Nq[2], Nq[1], Nq[0])
    A = None
    if aux is not None:
        A = np.moveaxis(aux[:nr], 1, 0).reshape(aux.shape[1], nr, Nq[2], Nq[1], Nq[0])
    F = target.argument(Q, A)
    for d in dirs:                       # xi1 -> last axis, xi2 -> axis -2, xi3 -> axis -3
        F = _contract(F, W, -1 - d)
    out = target.result(Q, F, A)
    data[:nr] = np.moveaxis(out.reshape(S, nr, g.Np), 0, 1)


def apply(Q, target, g, filter_matrices, state_auxiliary=None, direction="every"):
    """``Filters.apply!(Q, target, grid, filter; state_auxiliary, direction)``.

    ``Q``/``state_auxiliary``: arrays ``(nelem, nstate, Np)`` (MPIStateArray.data);
    ``filter_matrices``: one matrix or a (horizontal, vertical) pair."""
    if isinstance(filter_matrices, np.ndarray):
        filter_matrices = (filter_matrices, filter_matrices)
    Wh, Wv = filter_matrices[0], filter_matrices[-1]
    if direction in ("every", "horizontal"):
        _launch(g, Q, state_auxiliary, target, Wh, (0, 1))
    if direction in ("every", "vertical"):
        _launch(g, Q, state_auxiliary, target, Wv, (2,))
    return Q
