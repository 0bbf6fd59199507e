"""CPU checks of the EXPERIMENTAL bf16 convolution's design (csrc/sparse_conv_bf16.cu; not yet run on a GPU): the
flattened (offset, channel) contraction axis, the packed weight layout and the gather warp's two-offset id logic are
emulated in numpy exactly as the kernel indexes them and compared with the convolution oracle; the planner is checked like
the TF32 one (tests/test_conv_plan.py)."""
import ctypes

import numpy as np
import pytest
import torch

from canonicalvoting_b200.sparse import bf16 as B
from oracle import sparse_oracle as SO
from tests.test_conv_plan import SHAPES


def emulate(x, kernel, table):
    """Tile assembly of sc_conv_bf16_kernel for one 'row tile' = all rows: per k-block j, A[row, 8c:8c+8] is chunk c gathered
    through the ids of offset k_lo or k_lo + 1 (what the lane holds), B = the TMA box of the packed weights (zero beyond K)."""
    k3, cin, cout = kernel.shape
    n_out = table.shape[0]
    ktot = k3 * cin
    wp = B.pack_weights(torch.from_numpy(kernel)).float().numpy()
    total_kb = -(-ktot // B.KB)
    acc = np.zeros((n_out, cout), np.float64)
    for j in range(total_kb):
        k_lo = (B.KB * j) // cin
        single = (B.KB * j + B.KB - 1) // cin == k_lo
        ids_lo = table[:, k_lo] if k_lo < k3 else np.full(n_out, -1)
        ids_hi = table[:, k_lo + 1] if k_lo + 1 < k3 else np.full(n_out, -1)
        A = np.zeros((n_out, B.KB), np.float64)
        for c in range(8):
            flat = B.KB * j + 8 * c
            k_mine, ch = B.chunk_source(j, c, cin)
            assert k_mine in (k_lo, k_lo + 1) and (not single or k_mine == k_lo)     # a lane never needs a third offset
            assert ch % 8 == 0 and ch + 8 <= cin                                      # a chunk never straddles an offset
            ids = ids_lo if (single or k_mine == k_lo) else ids_hi
            ok = (ids >= 0) & (flat < ktot)
            A[ok, 8 * c:8 * c + 8] = x[ids[ok], ch:ch + 8]
        Bt = np.zeros((cout, B.KB), np.float64)
        w = min(B.KB, ktot - B.KB * j)
        Bt[:, :w] = wp[:, B.KB * j:B.KB * j + w]                                     # TMA zero-fills the tail
        acc += A @ Bt.T
    return acc


@pytest.mark.parametrize("cin,cout,K", [(32, 32, 3), (96, 96, 3), (128, 96, 3), (64, 16, 3), (32, 48, 2), (160, 32, 1), (96, 64, 5)])
def test_flattened_contraction_equals_the_convolution(cin, cout, K):
    g = torch.Generator().manual_seed(cin + cout + K)
    G, n = 12, 400
    lin = torch.randperm(G ** 3, generator=g)[:n]
    coords = torch.stack([torch.zeros_like(lin), lin // (G * G), (lin // G) % G, lin % G], 1).int()
    # inputs and weights rounded to bf16 first: the emulation and the oracle then see identical operands
    x = torch.randn(n, cin, generator=g).to(torch.bfloat16).float()
    kernel = (torch.randn(K ** 3, cin, cout, generator=g) * 0.1).to(torch.bfloat16).float()
    if K % 2 == 1:
        index = {tuple(c): i for i, c in enumerate(coords.tolist())}
        table = np.full((n, K ** 3), -1, np.int64)
        for k in range(K ** 3):                                       # offset order of the oracle: x fastest, centred
            d = (k % K - K // 2, (k // K) % K - K // 2, k // (K * K) - K // 2)
            for o, c in enumerate(coords.tolist()):
                table[o, k] = index.get((c[0], c[1] + d[0], c[2] + d[1], c[3] + d[2]), -1)
        want = SO.conv_same(coords, x.double(), kernel.double(), K, 1, None).numpy()
    else:
        table = torch.randint(-1, n, (n, K ** 3), generator=g).numpy()
        want = np.zeros((n, cout))
        for k in range(K ** 3):
            ok = table[:, k] >= 0
            want[ok] += x.double().numpy()[table[ok, k]] @ kernel.double().numpy()[k]
    got = emulate(x.numpy().astype(np.float64), kernel.numpy(), table)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("n_out,cin,cout,k3", SHAPES)
def test_bf16_plan_covers_every_k_block_once(lib_built, n_out, cin, cout, k3):
    from canonicalvoting_b200 import _lib
    L = _lib.load()
    p = (ctypes.c_int32 * 12)()
    assert L.cvb200_sc_conv_plan_bf16(n_out, cin, cout, k3, p, None, 0) == 0
    n_tiles, n_splits, n_whole, ks, n_units, total_kb, _, stages, nc, acc_stride, tmem_cols, smem = list(p)
    assert total_kb == -(-(k3 * cin) // 64) and nc * n_splits == cout and smem <= 227 * 1024 and stages >= 2
    assert tmem_cols <= 512 and 2 * acc_stride <= tmem_cols and acc_stride >= nc
    u = (ctypes.c_int32 * (6 * n_units))()
    assert L.cvb200_sc_conv_plan_bf16(n_out, cin, cout, k3, p, u, n_units) == 0
    U = np.ctypeslib.as_array(u).reshape(-1, 6)
    cover = np.zeros((-(-n_out // 128), n_splits, total_kb), np.int32)
    for row0, n0, kb0, kb1, pieces, split_tile in U:
        assert kb1 > kb0
        cover[row0 // 128, n0 // nc, kb0:kb1] += 1
    assert (cover == 1).all()


@pytest.mark.parametrize("cin,k3", [(32, 27), (96, 27), (128, 27), (64, 8), (160, 1), (96, 125), (32, 1)])
def test_kernel_index_arithmetic_equals_the_emulation(lib_built, cin, k3):
    """bf_chunk -- the function the CUDA gather warps call -- evaluated on the host for every (k-block, chunk) against the python
    formulation the emulation above uses."""
    from canonicalvoting_b200 import _lib
    L = _lib.load()
    ktot = k3 * cin
    out = (ctypes.c_int32 * 5)()
    for it in range(-(-ktot // B.KB)):
        k_lo = (B.KB * it) // cin
        single = (B.KB * it + B.KB - 1) // cin == k_lo
        for c in range(8):
            assert L.cvb200_sc_conv_bf16_chunk(it, c, cin, k3, out) == 0
            k_mine, ch = B.chunk_source(it, c, cin)
            assert list(out) == [k_lo, int(single), int(B.KB * it + 8 * c < ktot), int(k_mine != k_lo), ch]
            assert out[2] == 0 or k_mine < k3
