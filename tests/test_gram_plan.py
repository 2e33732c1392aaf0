"""Host-side schedule of the work-list Gram kernel (lobpcg_b200/csrc/gram_wl.cu), checked without a GPU through
lb2_gram_wl_plan_check: every needed 8x8 output block of every tile is owned by exactly one warp (masked diagonal
and ragged tiles included), the (tile, row-range) items partition [0, n) exactly for each tile, each CTA owns a
contiguous run of items, and the cost-weighted load is balanced."""
import ctypes as C

import numpy as np
import pytest

from lobpcg_b200 import api

SOLVER_SHAPES = [  # (ma, mb, upper, n): Gram shapes of the BASELINE configs C1..C5
    (60, 60, 1, 10_000), (384, 384, 1, 2_097_152), (600, 600, 1, 4_096_000), (900, 900, 1, 4_096_000),
    (900, 900, 1, 512_000), (600, 300, 0, 4_096_000), (300, 300, 1, 4_096_000), (256, 128, 0, 2_097_152),
    (896, 896, 1, 4_096_000), (904, 904, 1, 4_096_000), (1, 1, 1, 4096), (7, 260, 0, 8193), (1700, 1700, 1, 1_000_000),
]


def _check(ma, mb, upper, n, ncta, bk):
    st = (C.c_double * 4)()
    rc = api.lib().lb2_gram_wl_plan_check(ma, mb, upper, n, ncta, bk, st)
    assert rc == 0, f"plan check failed with code {rc} for {(ma, mb, upper, n, ncta, bk)}"
    return list(st)


@pytest.mark.parametrize("shape", SOLVER_SHAPES)
@pytest.mark.parametrize("bk", [16, 32])
def test_schedule_is_an_exact_cover(shape, bk):
    ma, mb, upper, n = shape
    tiles = ((ma + 127) // 128) * ((mb + 127) // 128)
    ncta = min(148, max(1, tiles * n // 4096))
    items, balance, waste, ntiles = _check(ma, mb, upper, n, ncta, bk)
    assert items <= ncta + ntiles
    if n >= 500_000:
        assert balance < 1.02           # busiest CTA within 2 % of the mean
    if n >= 500_000 and min(ma, mb) >= 256:
        assert waste < 1.07             # issued DMMA blocks / needed blocks (the first kernel: 1.2 - 1.45)


def test_random_shapes():
    rng = np.random.default_rng(11)
    for _ in range(300):
        upper = int(rng.integers(0, 2))
        ma = int(rng.integers(1, 1400))
        mb = ma if upper else int(rng.integers(1, 1400))
        n = int(rng.integers(4096, 5_000_000))
        ncta = int(rng.integers(1, 149))
        _check(ma, mb, upper, n, ncta, int(rng.choice([16, 32])))


def _sharing(ma, n, phase, window=64, samples=200, upper=1):
    share = C.c_double(0)
    assert api.lib().lb2_gram_wl_plan_sharing(ma, ma, upper, n, 148, 32, phase, window, samples, C.byref(share)) == 0
    return share.value


@pytest.mark.parametrize("m,n", [(896, 4_096_000), (600, 4_096_000), (384, 2_097_152), (896, 512_000)])
def test_phase_aligned_walk_lets_tiles_share_panels(m, n):
    """The cyclic, phase-aligned walk of the pieces (WlItem::c_start) puts tiles that share a panel on the same rows at
    the same time.  Model: all CTAs advance at their tiles' cost rate.  Measured on B200 at m=896, n=4.096 M: DRAM read
    205.4 -> 92.8 GB per launch (ratio 0.45), L2 hit rate 20 -> 50 % (profiles/ncu_traffic_r01.json)."""
    seq, aligned = _sharing(m, n, 0), _sharing(m, n, 1)
    assert seq > 0.9                      # walked from their first row, no two pieces are ever on the same rows
    assert aligned < 0.7 * seq      # fewer panels = fewer tiles per panel = less to share (m=384: 0.62, m=896: 0.38)
    if m == 896 and n == 4_096_000:
        assert aligned < 0.42


# ---- column-block schedule of the cached-Gram pass (lb2_<p>_gram_cols, gram_wl.cu: plan_schedule_cols) ------------------
COLS_SHAPES = [  # (m, nw, nprod, tri_c0, n): [X P W]^H [W | AW] of the BASELINE configs, full and soft-locked / first-pass widths
    (900, 300, 2, 600, 4_096_000), (900, 300, 2, 600, 512_000), (600, 200, 2, 400, 4_096_000), (384, 128, 2, 256, 2_097_152),
    (600, 300, 2, 300, 4_096_000), (750, 225, 2, 525, 4_096_000), (900, 300, 1, 600, 4_096_000), (600, 300, 1, -1, 4_096_000),
    (60, 20, 2, 40, 10_000), (240, 80, 2, 160, 32_768), (1, 1, 1, 0, 4096), (300, 300, 2, 0, 1_000_000),
]


def _check_cols(m, nw, nprod, tri_c0, n, ncta, bk):
    st = (C.c_double * 4)()
    rc = api.lib().lb2_gram_wl_cols_plan_check(m, nw, nprod, tri_c0, n, ncta, bk, st)
    assert rc == 0, f"column-block plan check failed with code {rc} for {(m, nw, nprod, tri_c0, n, ncta, bk)}"
    return list(st)


@pytest.mark.parametrize("shape", COLS_SHAPES)
@pytest.mark.parametrize("bk", [16, 32])
def test_column_block_schedule_is_an_exact_cover(shape, bk):
    m, nw, nprod, tri_c0, n = shape
    tiles = ((m + 127) // 128) * ((nw + 127) // 128) * nprod
    ncta = min(148, max(1, tiles * n // 4096))
    items, balance, ntiles, area = _check_cols(m, nw, nprod, tri_c0, n, ncta, bk)
    assert items <= ncta + ntiles
    if n >= 500_000:
        assert balance < 1.02
    if tri_c0 < 0:
        assert area == 1.0
    else:
        assert area <= 1.0
    if (m, nw, tri_c0) == (900, 300, 600):
        assert area < 0.9        # the tiles below the diagonal of W^H W / W^H A W are not computed


def test_column_block_random_shapes():
    rng = np.random.default_rng(12)
    for _ in range(300):
        nw = int(rng.integers(1, 500))
        nxp = int(rng.integers(0, 1000))
        tri = int(rng.choice([-1, nxp]))
        n = int(rng.integers(4096, 5_000_000))
        _check_cols(nxp + nw, nw, int(rng.integers(1, 3)), tri, n, int(rng.integers(1, 149)), int(rng.choice([16, 32])))
