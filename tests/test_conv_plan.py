"""CPU tests of the host-side planner of the persistent tensor-core convolution (csrc/sparse_conv_persist.cu `ps_plan` /
`ps_unit` through cvb200_sc_conv_plan): whatever the shape, the work units must cover every (row tile, channel block,
k-block) exactly once, pieces of a split tile must be contiguous and share the tile's scratch slot, and the shared-memory /
tensor-memory budgets of a CTA must hold.  No GPU needed."""
import ctypes

import numpy as np
import pytest

SHAPES = [  # (n_out, cin, cout, k3): the layers of MinkUNet34C on the C2 / C5 scenes + edge cases
    (50000, 96, 96, 27), (50000, 128, 96, 27), (50000, 128, 96, 1), (50000, 96, 64, 1), (50000, 32, 32, 8), (50000, 160, 32, 1),
    (17001, 32, 32, 27), (17001, 128, 96, 27), (17001, 96, 96, 27), (17001, 32, 32, 8), (4090, 32, 64, 27), (4090, 64, 64, 27),
    (4090, 192, 128, 27), (940, 64, 128, 27), (940, 128, 128, 27), (940, 384, 256, 27), (196, 128, 256, 27), (196, 256, 256, 27),
    (200000, 96, 96, 27), (72000, 128, 96, 27), (1, 32, 16, 1), (127, 64, 16, 27), (128, 64, 48, 27), (129, 512, 384, 27),
    (18944, 96, 96, 27), (18945, 96, 96, 27), (37888, 32, 1024, 1), (300, 32, 144, 125),
]


def plan(lib, n_out, cin, cout, k3):
    L = lib
    p = (ctypes.c_int32 * 12)()
    assert L.cvb200_sc_conv_plan(n_out, cin, cout, k3, p, None, 0) == 0
    keys = ("n_tiles", "n_splits", "n_whole", "ks", "n_units", "total_kb", "cblocks", "stages", "nc", "acc_stride", "tmem_cols", "smem_bytes")
    P = dict(zip(keys, p))
    u = (ctypes.c_int32 * (6 * P["n_units"]))()
    assert L.cvb200_sc_conv_plan(n_out, cin, cout, k3, p, u, P["n_units"]) == 0
    return P, np.ctypeslib.as_array(u).reshape(-1, 6).copy()


@pytest.mark.parametrize("n_out,cin,cout,k3", SHAPES)
def test_units_cover_the_convolution_exactly_once(lib_built, n_out, cin, cout, k3):
    from canonicalvoting_b200 import _lib
    P, U = plan(_lib.load(), n_out, cin, cout, k3)
    m_tiles = -(-n_out // 128)
    assert P["nc"] * P["n_splits"] == cout and P["nc"] % 16 == 0 and 16 <= P["nc"] <= 128
    assert P["n_tiles"] == m_tiles * P["n_splits"] and P["cblocks"] == cin // 32 and P["total_kb"] == k3 * (cin // 32)
    assert P["n_units"] == len(U) == P["n_whole"] + (P["n_tiles"] - P["n_whole"]) * P["ks"]
    # budgets of one CTA: 227 KiB of shared memory, 512 tensor-memory columns, two accumulators side by side
    assert P["smem_bytes"] <= 227 * 1024 and P["stages"] >= 2
    assert P["tmem_cols"] <= 512 and P["tmem_cols"] & (P["tmem_cols"] - 1) == 0 and P["acc_stride"] >= P["nc"] and 2 * P["acc_stride"] <= P["tmem_cols"]
    cover = np.zeros((m_tiles, P["n_splits"], P["total_kb"]), dtype=np.int32)
    pieces_of = {}
    for row0, n0, kb0, kb1, pieces, split_tile in U:
        assert row0 % 128 == 0 and 0 <= row0 < n_out and n0 % P["nc"] == 0 and 0 <= n0 < cout
        assert 0 <= kb0 <= kb1 <= P["total_kb"]
        cover[row0 // 128, n0 // P["nc"], kb0:kb1] += 1
        if pieces > 1:
            assert pieces == P["ks"] and 0 <= split_tile < 2 * 148            # scratch slots of one launch
            pieces_of.setdefault(split_tile, []).append((row0, n0, kb0, kb1))
        else:
            assert (kb0, kb1) == (0, P["total_kb"])
    assert (cover == 1).all()
    for slot, lst in pieces_of.items():                                        # one scratch slot = one tile, all its pieces
        assert len(lst) == P["ks"] and len({(r, n) for r, n, _, _ in lst}) == 1
        assert all(kb1 > kb0 for _, _, kb0, kb1 in lst), "an empty piece would never arrive at the tile's counter ... but must"
    if P["ks"] > 1:                                                             # a split is only planned for a partial wave
        assert P["n_whole"] % 148 == 0 and P["n_tiles"] - P["n_whole"] < 148


def test_plan_rejects_bad_shapes(lib_built):
    from canonicalvoting_b200 import _lib
    L = _lib.load()
    p = (ctypes.c_int32 * 12)()
    for bad in [(0, 32, 32, 27), (100, 33, 32, 27), (100, 32, 8, 27), (100, 32, 40, 27), (100, 32, 32, 0)]:
        assert L.cvb200_sc_conv_plan(*bad, p, None, 0) < 0
        assert b"sc_conv_plan" in L.cvb200_last_error()
