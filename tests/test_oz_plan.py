"""Host-side checks of the schedules of the int8 tensor path (lobpcg_b200/csrc/gram_i8.cu) — no GPU needed.
lb2_oz_plan_check builds the tile list / 2 x 2 super-tiles of the column-block products [X P W]^H [W | A W] and the work
schedule of each kernel variant and verifies: every output entry outside the tiles strictly below the diagonal of the Hermitian
block is written by exactly one tile, the items of every (tile or super-tile, level group) partition the 128-row chunks, every
worker's items are contiguous, lock-step cohorts give a CTA one item."""
import ctypes as C

import numpy as np
import pytest

from lobpcg_b200 import api

SHAPES = [  # (m, nw, nprod, tri_c0, n)
    (900, 300, 2, 600, 4_096_000), (900, 300, 2, 600, 512_000), (600, 200, 2, 400, 4_096_000), (384, 128, 2, 256, 2_097_152),
    (750, 225, 2, 525, 4_096_000), (900, 300, 1, 600, 4_096_000), (600, 300, 1, -1, 1_000_003), (240, 80, 2, 160, 262_144),
    (300, 300, 2, 0, 1_000_000), (131, 7, 2, 124, 300_000), (128, 128, 1, -1, 4096),
]


def check(m, nw, nprod, tri, n, workers, mode):
    st = (C.c_double * 4)()
    rc = api.lib().lb2_oz_plan_check(m, nw, nprod, tri, n, workers, mode, st)
    assert rc == 0, f"int8 schedule check failed with code {rc} for {(m, nw, nprod, tri, n, workers, mode)}"
    return list(st)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_int8_schedules_cover_the_outputs_exactly(shape, mode):
    m, nw, nprod, tri, n = shape
    workers = 33 if mode == 2 else 148
    items, balance, ntiles, used = check(m, nw, nprod, tri, n, workers, mode)
    assert used <= workers and items >= 1
    if mode != 1 and n >= 1_000_000:
        assert balance < 1.05          # equal-cost cut: the busiest worker is within 5 % of the mean (cost model)
    if (m, nw, nprod, tri) == (900, 300, 2, 600):
        assert ntiles == 42            # 48 tiles of the two 900 x 300 products minus the 6 below the diagonal of W^H W / W^H A W


def test_int8_schedules_random_shapes():
    rng = np.random.default_rng(31)
    for _ in range(200):
        nw = int(rng.integers(1, 420))
        nxp = int(rng.integers(0, 900))
        tri = int(rng.choice([-1, nxp]))
        n = int(rng.integers(4096, 5_000_000))
        mode = int(rng.integers(0, 3))
        check(nxp + nw, nw, int(rng.integers(1, 3)), tri, n, int(rng.integers(1, 149)) if mode != 2 else int(rng.integers(1, 38)), mode)
