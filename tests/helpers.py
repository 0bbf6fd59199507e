"""Shared helpers for the parity tests."""
import numpy as np


def assert_grid_close(got, want, rtol=1e-4, atol_frac=1e-6, what=""):
    """north_star tolerance: float vote maps within 1e-4 relative.  atomics make the float32
    summation order arbitrary, so tiny values next to large cancellations get an
    absolute floor of atol_frac * max|want|."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    atol = atol_frac * max(1.0, float(np.abs(want).max()) if want.size else 1.0)
    err = np.abs(got - want)
    bad = err > atol + rtol * np.abs(want)
    assert not bad.any(), "%s: %d/%d elements off, max abs err %.3e (max |want| %.3e)" % (
        what, int(bad.sum()), bad.size, float(err.max()), float(np.abs(want).max()))


def small_scene(n, grid, num_rots, seed, **kw):
    from canonicalvoting_b200 import synthetic
    return synthetic.make_scene(n, grid, num_rots, seed=seed, **kw)
