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
