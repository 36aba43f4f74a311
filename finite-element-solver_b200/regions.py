"""Geometric regions and field evaluation (ref fem/regions.py), vectorised.

A region is a callable over an (n, dim) coordinate array returning a boolean mask, as in the
reference.  `evaluate_field` turns a constant or a callable of position into an (n, c) array;
callables are first tried on the whole coordinate array at once (`f(X.T)` with X.T[k] the k-th
coordinate column, which is what `lambda p: [g(p[0], p[1])]` computes elementwise) and only
fall back to the reference's per-point loop if that does not produce the right shape.
"""
import numpy as np


# the reference's coordinate tolerance (ref fem/regions.py:26): it only absorbs round-off of linspace / midpoint
# arithmetic -- 0.4 * 3.0 is not the coordinate 1.2 of a mesh vertex without it
DEFAULT_ATOL = 1e-9


def everywhere():
    return lambda X: np.ones(len(X), dtype=bool)


def on_plane(axis, value, atol=DEFAULT_ATOL):
    """Points whose `axis` coordinate equals `value` within atol, inclusive (ref fem/regions.py:35-38)."""
    return lambda X: np.abs(np.asarray(X)[:, axis] - value) <= atol


def in_box(lower, upper, atol=DEFAULT_ATOL):
    """Points inside an axis-aligned box, inclusive within atol; None leaves a side unbounded
    (ref fem/regions.py:41-58)."""
    def region(X):
        X = np.asarray(X)
        m = np.ones(len(X), dtype=bool)
        for k, lo in enumerate(lower):
            if lo is not None:
                m &= X[:, k] >= lo - atol
        for k, hi in enumerate(upper):
            if hi is not None:
                m &= X[:, k] <= hi + atol
        return m
    return region


def intersect(*regions):
    def region(X):
        m = np.ones(len(X), dtype=bool)
        for r in regions:
            m &= np.asarray(r(X), dtype=bool)
        return m
    region.mesh_bound = any(getattr(r, 'mesh_bound', False) for r in regions)
    return region


def union(*regions):
    def region(X):
        m = np.zeros(len(X), dtype=bool)
        for r in regions:
            m |= np.asarray(r(X), dtype=bool)
        return m
    region.mesh_bound = any(getattr(r, 'mesh_bound', False) for r in regions)
    return region


def at_indices(indices):
    idx = np.asarray(indices, dtype=int)

    def region(X):
        m = np.zeros(len(X), dtype=bool)
        m[idx] = True
        return m
    region.mesh_bound = True
    return region


def is_mesh_bound(region):
    return bool(getattr(region, 'mesh_bound', False))


def coerce_components(value, points, n_components):
    """(n, c) array from a constant / callable; None components become NaN."""
    n = len(points)
    if value is None:
        return np.zeros((n, n_components))
    if callable(value):
        out = None
        if n:
            try:
                trial = value(np.asarray(points).T)
                cols = [np.broadcast_to(np.nan if c is None else np.asarray(c, dtype=float), (n,))
                        for c in (trial if isinstance(trial, (list, tuple)) else [trial])]
                cand = np.stack(cols, axis=1)
                probe = np.atleast_1d(np.array([np.nan if c is None else float(c)
                                                for c in np.atleast_1d(np.asarray(value(points[0]), dtype=object))]))
                if cand.shape == (n, len(probe)) and np.allclose(cand[0], probe, equal_nan=True):
                    out = cand
            except Exception:
                out = None
        if out is None:
            out = np.array([[np.nan if c is None else float(c)
                             for c in np.atleast_1d(np.asarray(value(p), dtype=object))] for p in points]
                           ).reshape(n, -1)
        return out
    comps = np.atleast_1d(np.asarray(value, dtype=object))
    row = np.array([np.nan if c is None else float(c) for c in comps])
    return np.tile(row, (n, 1))


def evaluate_field(value, points, n_components):
    values = coerce_components(value, points, n_components)
    if values.shape != (len(points), n_components):
        raise ValueError(f'field must give {n_components} component(s) per point, got shape {values.shape} '
                         f'for {len(points)} point(s)')
    if np.any(np.isnan(values)):
        raise ValueError('field component is None (or NaN); every component must be a real number here')
    return values
