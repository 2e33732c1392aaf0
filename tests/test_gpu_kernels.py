"""GPU parity tests of the individual kernels, through the C ABI (lobpcg_b200.api -> liblobpcg_b200.so),
against the CPU oracle (oracle/numpy_oracle.py) and the reference's golden vectors.

Tolerances: the kernels are floating point; sums are re-associated (split-n partial tiles, DMMA fragment
order), so parity is `rtol` relative to the magnitude of the result: 1e-12 (d,z), 2e-5 (s,c).  Generators
and integer index work (fill_uniform) are bit-exact."""
import json
import os
from pathlib import Path

import numpy as np
import pytest

from lobpcg_b200 import api
from lobpcg_b200 import problems as pr
from oracle import numpy_oracle as no

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
KA = json.loads((GOLD / "known_answers.json").read_text())
REF = np.load(GOLD / "reference_runs.npz")
DTYPES = [np.float64, np.float32, np.complex128, np.complex64]


def rtol(dt):
    return 1e-12 if np.dtype(dt) in (np.dtype(np.float64), np.dtype(np.complex128)) else 2e-5


def rand(rng, shape, dt):
    a = rng.standard_normal(shape)
    if np.dtype(dt).kind == "c":
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a.astype(dt))


def close(got, ref, tol):
    scale = max(float(np.abs(ref).max()), 1e-30) if ref.size else 1.0
    err = float(np.abs(got - ref).max()) / scale if ref.size else 0.0
    assert err < tol, f"rel err {err:.3e} >= {tol:.1e}"


# ------------------------------------------------------------------------------------------------ Gram
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(1, 1, 1), (37, 3, 5), (4099, 20, 20), (30001, 64, 33), (20000, 150, 129),
                                   (8193, 260, 7)])
def test_gram_matches_oracle(ctx, dt, shape):
    n, ma, mb = shape
    rng = np.random.default_rng(n + ma)
    A, B = rand(rng, (n, ma), dt), rand(rng, (n, mb), dt)
    dA, dB = api.DeviceArray.from_numpy(ctx, A), api.DeviceArray.from_numpy(ctx, B)
    close(api.gram(ctx, dA, dB).numpy(ctx), no.gram_cross(A, B), rtol(dt))
    G = api.gram(ctx, dA, dA, upper=True).numpy(ctx)
    full = A.conj().T @ A
    close(np.triu(G), no.gram_self(A), rtol(dt))       # reference semantics: upper triangle (syrk/herk)
    close(G, full, rtol(dt))                            # ours additionally mirrors => full Hermitian matrix


@pytest.mark.parametrize("shape", [(9000, 900), (20000, 300), (12345, 517), (8192, 128), (50000, 129), (6000, 1030),
                                   (4500, 256), (7001, 384), (30000, 136), (10001, 263)])
@pytest.mark.parametrize("opts", [{}, {"gram_bk": 16}, {"gram_strip_max": -1}, {"gram_wl": 0}, {"gram_strip_fma": -1}])
def test_gram_hermitian_worklist_kernel(ctx, shape, opts):
    """f64 Hermitian products S^H S and S^H (H S) (H = real diagonal, so the product is symmetric) through the
    work-list kernel: masked diagonal tiles, ragged last tile column (strip path and in-list path), odd n."""
    n, m = shape
    rng = np.random.default_rng(n + m)
    S = rand(rng, (n, m), np.float64)
    HS = np.asfortranarray(rng.uniform(0.5, 1.5, n)[:, None] * S)
    dS, dHS = api.DeviceArray.from_numpy(ctx, S), api.DeviceArray.from_numpy(ctx, HS)
    for k, v in opts.items():
        ctx.set_option(k, v)
    try:
        G1 = api.gram(ctx, dS, dS, upper=True).numpy(ctx)
        G2 = api.gram(ctx, dS, dHS, upper=True).numpy(ctx)
    finally:
        for k in opts:
            ctx.set_option(k, -1 if k == "gram_wl" else 0)
    close(G1, S.T @ S, 1e-12)
    close(G2, S.T @ HS, 1e-12)
    assert np.array_equal(G1, G1.T) and np.array_equal(G2, G2.T)     # mirrored exactly


def test_gram_worklist_rectangular_when_forced(ctx):
    rng = np.random.default_rng(3)
    A, B = rand(rng, (10000, 600), np.float64), rand(rng, (10000, 300), np.float64)
    dA, dB = api.DeviceArray.from_numpy(ctx, A), api.DeviceArray.from_numpy(ctx, B)
    ctx.set_option("gram_wl", 1)
    try:
        G = api.gram(ctx, dA, dB).numpy(ctx)
    finally:
        ctx.set_option("gram_wl", -1)
    close(G, A.T @ B, 1e-12)


def test_gram_worklist_is_deterministic(ctx):
    rng = np.random.default_rng(4)
    S = rand(rng, (30000, 420), np.float64)
    dS = api.DeviceArray.from_numpy(ctx, S)
    G1 = api.gram(ctx, dS, dS, upper=True).numpy(ctx)
    G2 = api.gram(ctx, dS, dS, upper=True).numpy(ctx)
    assert np.array_equal(G1, G2)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_gram_padded_leading_dimension_and_odd_alignment(ctx, dt):
    """ld > n, odd ld and an odd column offset force the non-vectorised cp.async path."""
    rng = np.random.default_rng(5)
    n, ma = 5001, 70
    A = rand(rng, (n, ma), dt)
    dA = api.DeviceArray.from_numpy(ctx, A, ld=n + 6)
    close(api.gram(ctx, dA, dA, upper=True).numpy(ctx), A.conj().T @ A, rtol(dt))
    dB = api.DeviceArray.from_numpy(ctx, A, ld=n + 1)
    close(api.gram(ctx, dB, dA).numpy(ctx), A.conj().T @ A, rtol(dt))


def test_gram_known_answers_from_reference_tests(ctx):
    k = KA["gram_self_d_general"]                      # reference tests/test_gram.c:156-174
    U = np.asfortranarray(np.array(k["U"], dtype=np.float64))
    G = api.gram(ctx, api.DeviceArray.from_numpy(ctx, U), api.DeviceArray.from_numpy(ctx, U), upper=True).numpy(ctx)
    assert abs(G[0, 0] - 2.0) < 1e-12 and abs(G[0, 1] - 1.0) < 1e-12 and abs(G[1, 1] - 2.0) < 1e-12
    k = KA["gram_self_z_with_B"]                       # tests/test_gram.c:203-227: G = U^H B U
    U = np.asfortranarray(np.array(k["U_re"], dtype=np.complex128) + 1j * np.array(k["U_im"]))
    B = api.diag_op(np.array(k["Bdiag"]), np.complex128)
    dU = api.DeviceArray.from_numpy(ctx, U)
    G = api.gram(ctx, dU, B.apply(ctx, dU)).numpy(ctx)
    assert abs(G[0, 0] - 4.0) < 1e-12 and abs(G[0, 1]) < 1e-12 and abs(G[1, 1] - 6.0) < 1e-12


def test_gram_is_deterministic(ctx):
    rng = np.random.default_rng(1)
    A = rand(rng, (50000, 96), np.float64)
    dA = api.DeviceArray.from_numpy(ctx, A)
    G1 = api.gram(ctx, dA, dA, upper=True).numpy(ctx)
    G2 = api.gram(ctx, dA, dA, upper=True).numpy(ctx)
    assert np.array_equal(G1, G2)


def test_gram_simt_and_dmma_paths_agree(ctx):
    rng = np.random.default_rng(2)
    A, B = rand(rng, (12345, 130), np.float64), rand(rng, (12345, 90), np.float64)
    dA, dB = api.DeviceArray.from_numpy(ctx, A), api.DeviceArray.from_numpy(ctx, B)
    G1 = api.gram(ctx, dA, dB).numpy(ctx)
    for tile in (64, 128):
        ctx.set_option("gram_tile", tile)
        close(api.gram(ctx, dA, dB).numpy(ctx), G1, 1e-13)
    ctx.set_option("gram_tile", 0)
    ctx.set_option("force_simt", 1)
    G2 = api.gram(ctx, dA, dB).numpy(ctx)
    ctx.set_option("force_simt", 0)
    close(G2, G1, 1e-13)


def test_gram_float_has_no_accumulation_bias(ctx):
    """The tensor core truncates when adding into its fp32 accumulator (relative bias ~2^-25 per MMA, -4.8e-4 on a Gram
    diagonal after 4 M rows); the tcgen05 kernel therefore accumulates short groups in TMEM and the long sum in fp64."""
    n = 1 << 20
    rng = np.random.default_rng(1)
    A = rand(rng, (n, 64), np.float32)
    dA = api.DeviceArray.from_numpy(ctx, A)
    G = api.gram(ctx, dA, dA, upper=True).numpy(ctx).astype(np.float64)
    ref = A.astype(np.float64).T @ A.astype(np.float64)
    d = (np.diag(G) - np.diag(ref)) / np.diag(ref)
    assert np.abs(d).max() < 5e-6 and abs(d.mean()) < 5e-6


@pytest.mark.parametrize("shape", [(4096, 128, 128), (5000, 20, 20), (20000, 300, 300), (33333, 129, 70), (8192, 260, 260),
                                   (7777, 64, 200)])
def test_gram_float_tcgen05_matches_fp64_reference(ctx, shape):
    """float Gram on tcgen05 (kind::tf32, 3xTF32 split, accumulator in TMEM; csrc/gram_tc5.cu) vs an fp64 product of the
    same float inputs: fp32-level accuracy (2e-5 of the largest entry), symmetric and rectangular, ragged tiles."""
    n, ma, mb = shape
    rng = np.random.default_rng(n + ma)
    A = rand(rng, (n, ma), np.float32)
    B = A if ma == mb else rand(rng, (n, mb), np.float32)
    dA = api.DeviceArray.from_numpy(ctx, A, ld=(n + 3) // 4 * 4)
    dB = dA if ma == mb else api.DeviceArray.from_numpy(ctx, B, ld=(n + 3) // 4 * 4)
    ctx.set_option("gram_tc5", 1)
    try:
        G = api.gram(ctx, dA, dB, upper=(ma == mb)).numpy(ctx)
        G2 = api.gram(ctx, dA, dB, upper=False).numpy(ctx)
    finally:
        ctx.set_option("gram_tc5", -1)
    ref = A.astype(np.float64).T @ B.astype(np.float64)
    close(G, ref, 2e-5)
    close(G2, ref, 2e-5)
    if ma == mb:
        assert np.array_equal(G, G.T)


# ------------------------------------------------------------------------------------------------ tall NN
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(1, 1, 1), (130, 5, 3), (4097, 60, 40), (30001, 33, 129), (9000, 300, 70),
                                   (5000, 64, 300), (5001, 37, 90), (5003, 20, 16), (4999, 50, 30), (5000, 33, 110)])
@pytest.mark.parametrize("ab", [(1.0, 0.0), (-1.0, 1.0)])
def test_tall_nn_matches_oracle(ctx, dt, shape, ab):
    n, kd, nb = shape
    rng = np.random.default_rng(n + kd + nb)
    S, Cm, O0 = rand(rng, (n, kd), dt), rand(rng, (kd, nb), dt), rand(rng, (n, nb), dt)
    dO = api.DeviceArray.from_numpy(ctx, O0)
    api.tall_nn(ctx, api.DeviceArray.from_numpy(ctx, S), api.DeviceArray.from_numpy(ctx, Cm), dO, alpha=ab[0], beta=ab[1])
    close(dO.numpy(ctx), ab[0] * (S @ Cm) + ab[1] * O0, rtol(dt) * 10)


@pytest.mark.parametrize("shape", [(4096, 32, 16), (5000, 60, 40), (20000, 300, 200), (33333, 129, 70), (8192, 900, 300),
                                   (7777, 600, 129), (4100, 7, 5)])
@pytest.mark.parametrize("ab", [(1.0, 0.0), (-1.0, 1.0)])
def test_tall_nn_float_tcgen05_matches_fp64_reference(ctx, shape, ab):
    """float projection on tcgen05 (csrc/nn_tc5.cu: MN-major A, run-time UMMA N, 3xTF32, fp64 drain) vs an fp64 product."""
    n, kd, nb = shape
    rng = np.random.default_rng(n + kd)
    S, Cm, O0 = rand(rng, (n, kd), np.float32), rand(rng, (kd, nb), np.float32), rand(rng, (n, nb), np.float32)
    ldn = (n + 3) // 4 * 4
    dO = api.DeviceArray.from_numpy(ctx, O0, ld=ldn)
    ctx.set_option("gram_tc5", 1)
    try:
        api.tall_nn(ctx, api.DeviceArray.from_numpy(ctx, S, ld=ldn), api.DeviceArray.from_numpy(ctx, Cm), dO, alpha=ab[0], beta=ab[1])
    finally:
        ctx.set_option("gram_tc5", -1)
    ref = ab[0] * (S.astype(np.float64) @ Cm.astype(np.float64)) + ab[1] * O0
    close(dO.numpy(ctx), ref, 2e-6)


def test_tall_nn_beta_zero_ignores_nan_output(ctx):
    rng = np.random.default_rng(3)
    S, Cm = rand(rng, (3000, 40), np.float64), rand(rng, (40, 24), np.float64)
    dO = api.DeviceArray.from_numpy(ctx, np.full((3000, 24), np.nan))
    api.tall_nn(ctx, api.DeviceArray.from_numpy(ctx, S), api.DeviceArray.from_numpy(ctx, Cm), dO, 1.0, 0.0)
    close(dO.numpy(ctx), S @ Cm, 1e-12)


# ------------------------------------------------------------------------------------------------ residual
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(1, 1), (777, 3), (100003, 17)])
def test_residual_and_norms_match_oracle(ctx, dt, shape):
    n, nc = shape
    rng = np.random.default_rng(n)
    AX, BX = rand(rng, (n, nc), dt), rand(rng, (n, nc), dt)
    lam = rng.standard_normal(nc).astype(api.REAL[api.PREFIX[np.dtype(dt)]])
    W, ss = api.residual(ctx, api.DeviceArray.from_numpy(ctx, AX), api.DeviceArray.from_numpy(ctx, BX),
                         api.DeviceArray.from_numpy(ctx, lam))
    Wr = no.get_residual(BX, AX, lam)      # oracle: W = AX - X diag(lam) with X := BX
    close(W.numpy(ctx), Wr, rtol(dt))
    close(ss.numpy(ctx), np.linalg.norm(Wr, axis=0) ** 2, rtol(dt) * 10)
    _, ss2 = api.residual(ctx, api.DeviceArray.from_numpy(ctx, AX), api.DeviceArray.from_numpy(ctx, BX),
                          api.DeviceArray.from_numpy(ctx, lam), write=False)
    assert np.array_equal(ss.numpy(ctx), ss2.numpy(ctx))
    close(api.col_sumsq(ctx, api.DeviceArray.from_numpy(ctx, AX)).numpy(ctx), np.linalg.norm(AX, axis=0) ** 2, rtol(dt) * 10)


def test_residual_known_answer_from_reference_tests(ctx):
    k = KA["residual_noneigvec_real"]                  # tests/test_residual.c:277-300: R = [-1, 0, 3]
    x = np.array(k["x"])[:, None]
    ax = np.array(k["Adiag"])[:, None] * x
    W, ss = api.residual(ctx, api.DeviceArray.from_numpy(ctx, ax), api.DeviceArray.from_numpy(ctx, x),
                         api.DeviceArray.from_numpy(ctx, np.array([k["lambda"]])))
    assert np.allclose(W.numpy(ctx)[:, 0], k["R"], atol=1e-15)
    assert abs(np.sqrt(ss.numpy(ctx)[0]) / (3.0 + 2.0) - np.sqrt(10.0) / 5.0) < 1e-15   # test_residual.c:450-529


# ------------------------------------------------------------------------------------------------ generator
@pytest.mark.parametrize("dt", DTYPES)
def test_fill_uniform_is_bit_exact_with_host_generator(ctx, dt):
    n, k = 1237, 5
    X = api.fill_uniform(ctx, n, k, dt, seed=7).numpy(ctx)
    assert np.array_equal(X, pr.initial_block(n, k, 7, dt))
    # row-partitioned fill reproduces the same global block (multi-GPU invariance)
    lo = api.fill_uniform(ctx, 600, k, dt, seed=7, n_global=n, row0=0).numpy(ctx)
    hi = api.fill_uniform(ctx, n - 600, k, dt, seed=7, n_global=n, row0=600).numpy(ctx)
    assert np.array_equal(np.vstack([lo, hi]), X)


# ------------------------------------------------------------------------------------------------ SpMM
GRIDS = [(1,), (5,), (300,), (33, 7), (100, 100), (7, 5, 3), (32, 8, 16), (45, 37, 29), (64, 64, 64)]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("grid", GRIDS)
def test_stencil_and_csr_match_oracle(ctx, dt, grid):
    n = int(np.prod(grid))
    rng = np.random.default_rng(n)
    for nc in (1, 6):
        X = rand(rng, (n, nc), dt)
        ref = no.op_stencil(grid)(X)
        dX = api.DeviceArray.from_numpy(ctx, X)
        close(api.stencil_op(grid, dt).apply(ctx, dX).numpy(ctx), ref, rtol(dt))
        rp, c, v = pr.laplacian_csr(grid)
        # CSR input that is a Dirichlet stencil is recognised and routed to the stencil kernel (capi.cu:
        # detect_stencil); LB2_CSR_NO_STENCIL_DETECT forces the general CSR kernel — both must match the oracle
        close(api.csr_op(rp, c, v.astype(dt)).apply(ctx, dX).numpy(ctx), ref, rtol(dt))
        os.environ["LB2_CSR_NO_STENCIL_DETECT"] = "1"
        try:
            close(api.csr_op(rp, c, v.astype(dt)).apply(ctx, dX).numpy(ctx), ref, rtol(dt))
        finally:
            del os.environ["LB2_CSR_NO_STENCIL_DETECT"]


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_stencil_with_potential_and_diag_operator(ctx, dt):
    grid = (21, 17, 13)
    n = int(np.prod(grid))
    rng = np.random.default_rng(9)
    X = rand(rng, (n, 5), dt)
    pot = pr.harmonic_potential(grid, 0.3)
    dX = api.DeviceArray.from_numpy(ctx, X)
    close(api.stencil_op(grid, dt, potential=pot).apply(ctx, dX).numpy(ctx), no.op_stencil(grid, potential=pot)(X), 1e-13)
    d = pr.mass_diagonal(n)
    close(api.diag_op(d, dt).apply(ctx, dX).numpy(ctx), d[:, None] * X, 1e-15)


def test_csr_general_sparsity_and_wide_blocks(ctx):
    import scipy.sparse as sp
    rng = np.random.default_rng(4)
    n = 5000
    M = sp.random(n, n, density=0.002, random_state=3, format="csr") + sp.eye(n, format="csr") * 3.0
    M = M.tocsr(); M.sort_indices()
    X = rand(rng, (n, 45), np.float64)
    op = api.csr_op(M.indptr, M.indices, M.data)
    dX = api.DeviceArray.from_numpy(ctx, X)
    for cols in (0, 4, 8, 16, 32):
        ctx.set_option("spmm_cols", cols)
        close(op.apply(ctx, dX).numpy(ctx), M @ X, 1e-13)
    ctx.set_option("spmm_cols", 0)


def test_bdg_operator_matches_dense_blocks(ctx):
    grid = (6, 5, 4)
    m = int(np.prod(grid))
    K = no.op_stencil(grid)
    shift, d = 0.5, 0.5 * np.exp(0.7j)
    rng = np.random.default_rng(6)
    X = rand(rng, (2 * m, 3), np.complex128)
    u, v = X[:m], X[m:]
    ref = np.vstack([K(u) + shift * u + d * v, np.conj(d) * u + K(v) + shift * v])
    got = api.bdg_op(grid, np.complex128, shift, d).apply(ctx, api.DeviceArray.from_numpy(ctx, X)).numpy(ctx)
    close(got, ref, 1e-13)


@pytest.mark.parametrize("dt", [np.complex128, np.float64])
def test_bdg_slabs_with_neighbour_planes_reproduce_the_global_operator(ctx, dt):
    """lb2_op_bdg_slab: the row-partitioned BdG operator, emulated on one GPU.  Both fields are split by the
    same z-slabs, a rank holds [u slab ; v slab], the block coupling is local and each field reads its own boundary planes
    from the neighbours' blocks."""
    g, world, nc = (8, 6, 12), 3, 4
    m = int(np.prod(g))
    plane, gzl = g[0] * g[1], g[2] // world
    ml = plane * gzl
    shift, d = 0.5, (0.5 * np.exp(0.7j) if np.dtype(dt).kind == "c" else 0.35)
    rng = np.random.default_rng(8)
    X = rand(rng, (2 * m, nc), dt)
    ref = api.bdg_op(g, dt, shift, d).apply(ctx, api.DeviceArray.from_numpy(ctx, X)).numpy(ctx)
    local = [np.asfortranarray(np.vstack([X[r * ml:(r + 1) * ml], X[m + r * ml:m + (r + 1) * ml]])) for r in range(world)]
    blocks = [api.DeviceArray.from_numpy(ctx, b) for b in local]
    item = np.dtype(dt).itemsize
    for r in range(world):
        op = api.bdg_slab_op(g, r * gzl, gzl, dt, shift, d)
        lo = blocks[r - 1].ptr + (gzl - 1) * plane * item if r > 0 else None      # last plane of the u slab below
        hi = blocks[r + 1].ptr if r + 1 < world else None                         # first plane of the u slab above
        api.set_halo(op, lo, hi, 2 * ml)
        Y = op.apply(ctx, blocks[r]).numpy(ctx)
        close(Y[:ml], ref[r * ml:(r + 1) * ml], 1e-13)
        close(Y[ml:], ref[m + r * ml:m + (r + 1) * ml], 1e-13)


def test_stencil_linearity_at_full_size(ctx):
    """Size-independent property at the BASELINE size (160^3): A(aX + bY) = a AX + b AY and symmetry
    <X, A Y> = <A X, Y>, checked through Gram kernels so that nothing leaves the device but 2x2 matrices."""
    g = (160, 160, 160)
    n = g[0] ** 3
    A = api.stencil_op(g, np.float64)
    X = api.fill_uniform(ctx, n, 2, np.float64, seed=1)
    AX = A.apply(ctx, X)
    G1 = api.gram(ctx, X, AX).numpy(ctx)          # X^T A X must be symmetric
    assert abs(G1[0, 1] - G1[1, 0]) < 1e-10 * abs(G1).max()
    Cm = api.DeviceArray.from_numpy(ctx, np.array([[2.0], [-3.0]]))
    Z = api.DeviceArray((n, 1), np.float64)
    api.tall_nn(ctx, X, Cm, Z)                    # Z = 2 x0 - 3 x1
    AZ = A.apply(ctx, Z)
    AZ2 = api.DeviceArray((n, 1), np.float64)
    api.tall_nn(ctx, AX, Cm, AZ2)
    diff = api.DeviceArray((n, 1), np.float64)
    one = api.DeviceArray.from_numpy(ctx, np.array([1.0]))
    W, ss = api.residual(ctx, AZ, AZ2, one)       # AZ - 1*AZ2
    assert np.sqrt(ss.numpy(ctx)[0]) < 1e-12 * np.sqrt(api.col_sumsq(ctx, AZ).numpy(ctx)[0])
    # analytic check: constant-one vector -> interior rows give 0, faces give boundary counts
    ones = api.DeviceArray.from_numpy(ctx, np.ones((n, 1)))
    s = api.col_sumsq(ctx, A.apply(ctx, ones)).numpy(ctx)[0]
    gi = g[0]
    expected = 6 * (gi - 2) ** 2 * 1 + 12 * (gi - 2) * 4 + 8 * 9
    assert abs(s - expected) < 1e-9 * expected


@pytest.mark.parametrize("dt", [np.float64, np.complex128, np.float32])
def test_stencil_z_slabs_with_halo_planes_reproduce_the_global_operator(ctx, dt):
    """Row-partitioned form used on several GPUs, emulated on one: each z-slab is applied by its own launch and
    reads the neighbouring slab's boundary plane through the halo pointers (on an NVSwitch box those are CUDA-IPC
    mappings of the neighbour rank's memory)."""
    g, world, nc = (20, 12, 16), 4, 5
    n = int(np.prod(g))
    plane = g[0] * g[1]
    rng = np.random.default_rng(12)
    X = rand(rng, (n, nc), dt)
    ref = no.op_stencil(g)(X)
    dX = api.DeviceArray.from_numpy(ctx, X)
    item = np.dtype(dt).itemsize
    out = []
    for r in range(world):
        gzl = g[2] // world
        row0, nl = r * gzl * plane, gzl * plane
        lo = dX.ptr + (row0 - plane) * item if r > 0 else None
        hi = dX.ptr + (row0 + nl) * item if r + 1 < world else None
        Y = api.stencil_halo_apply(ctx, (g[0], g[1], gzl), dX.rows(row0, nl), lo, hi, n)
        out.append(Y.numpy(ctx))
    close(np.vstack(out), ref, rtol(dt))


@pytest.mark.parametrize("dt", DTYPES)
def test_csr_row_blocks_with_neighbour_blocks_reproduce_the_global_operator(ctx, dt):
    """lb2_op_csr_slab: the row-partitioned CSR operator, emulated on one GPU — every row block is its own
    operator whose kernel reads rows of the neighbouring blocks in place (on an NVSwitch box: the neighbours' arenas).
    Random banded matrix (couplings up to one block away), ragged rows, an empty row, all column-group widths."""
    import scipy.sparse as sp
    from lobpcg_b200 import dist
    rng = np.random.default_rng(21)
    world, nb, nc = 4, 700, 21
    n = world * nb
    rows, cols = [], []
    for i in range(n):
        if i == 1234:
            continue                                     # empty row
        w = int(rng.integers(1, 12))
        cand = np.arange(max(0, (i // nb - 1) * nb), min(n, (i // nb + 2) * nb))    # own block and both neighbours
        rows += [i] * w
        cols += list(rng.choice(cand, size=w, replace=False))
    vals = rand(rng, (len(rows),), dt)
    M = sp.csr_matrix((vals, (rows, cols)), shape=(n, n)); M.sort_indices()
    X = rand(rng, (n, nc), dt)
    ref = M @ X
    blocks = [api.DeviceArray.from_numpy(ctx, np.asfortranarray(X[r * nb:(r + 1) * nb])) for r in range(world)]
    for spmm_cols in (0, 4, 8, 32):
        ctx.set_option("spmm_cols", spmm_cols)
        out = []
        for r in range(world):
            rp, c, v = dist.csr_row_block(M.indptr, M.indices, M.data, r * nb, nb)
            op = api.csr_slab_op(n, r * nb, rp, c, v)
            api.set_halo(op, blocks[r - 1].ptr if r > 0 else None, blocks[r + 1].ptr if r + 1 < world else None, nb)
            out.append(op.apply(ctx, blocks[r]).numpy(ctx))
        close(np.vstack(out), ref, 10 * rtol(dt))
    ctx.set_option("spmm_cols", 0)
    # a coupling beyond the neighbouring blocks is refused at construction
    far = sp.csr_matrix(([1.0], ([0], [n - 1])), shape=(n, n)) + sp.eye(n, format="csr")
    far = far.tocsr().astype(dt)
    rp, c, v = dist.csr_row_block(far.indptr, far.indices, far.data, 0, nb)
    with pytest.raises(api.LobpcgB200Error):
        api.csr_slab_op(n, 0, rp, c, v)


def test_csr_stencil_detection_handles_variable_diagonal_and_rejects_near_misses(ctx):
    """detect_stencil (capi.cu): a CSR matrix that IS a Dirichlet stencil with an arbitrary diagonal takes the
    stencil kernel; one changed off-diagonal value or a missing entry must fall back to the general kernel — the
    result is checked against the oracle either way."""
    import scipy.sparse as sp
    g = (12, 9, 7)
    n = int(np.prod(g))
    rng = np.random.default_rng(21)
    X = rand(rng, (n, 5), np.float64)
    dX = api.DeviceArray.from_numpy(ctx, X)
    pot = rng.standard_normal(n)
    rp, c, v = pr.laplacian_csr(g, potential=pot)
    M = sp.csr_matrix((v, c, rp), shape=(n, n))
    close(api.csr_op(rp, c, v).apply(ctx, dX).numpy(ctx), M @ X, 1e-13)
    v2 = v.copy(); v2[5] = -1.5                      # one off-diagonal differs
    M2 = sp.csr_matrix((v2, c, rp), shape=(n, n))
    close(api.csr_op(rp, c, v2).apply(ctx, dX).numpy(ctx), M2 @ X, 1e-13)
    M3 = M.tolil(); M3[40, 41] = 0.0; M3 = M3.tocsr(); M3.eliminate_zeros(); M3.sort_indices()   # missing neighbour
    close(api.csr_op(M3.indptr, M3.indices, M3.data).apply(ctx, dX).numpy(ctx), M3 @ X, 1e-13)
    M4 = (M + sp.eye(n, k=3, format="csr") * 0.25).tocsr(); M4.sort_indices()                    # extra band
    close(api.csr_op(M4.indptr, M4.indices, M4.data).apply(ctx, dX).numpy(ctx), M4 @ X, 1e-13)


# ------------------------------------------------------------------------------------------------ host <-> device
def test_pipelined_host_copy_round_trip_is_exact(ctx):
    """Blocks of 128 MB and more go through the pinned-ring copy (csrc/hostcopy.cu): odd sizes, both directions."""
    rng = np.random.default_rng(9)
    for nbytes in (128 * 2 ** 20 + 8, 333 * 2 ** 20 + 24):
        a = rng.integers(0, 2 ** 62, size=nbytes // 8, dtype=np.int64).view(np.float64)
        d = api.DeviceArray((a.size,), np.float64)
        api._ck(api.lib().lb2_memcpy_h2d(ctx.h, d.ptr, a.ctypes.data, a.nbytes), "h2d")
        b = np.zeros_like(a)
        api._ck(api.lib().lb2_memcpy_d2h(ctx.h, b.ctypes.data, d.ptr, b.nbytes), "d2h")
        assert np.array_equal(a.view(np.int64), b.view(np.int64))
        ss = api.col_sumsq(ctx, api.DeviceArray.from_numpy(ctx, np.ones((1000, 1))))     # small path still works
        assert abs(ss.numpy(ctx)[0] - 1000.0) < 1e-9
        d.free()


# ------------------------------------------------------------------------------------------------ preconditioner
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("degree", [0, 1, 2, 7])
@pytest.mark.parametrize("inner", ["stencil", "stencil+potential", "csr", "stencil-unfused"])
def test_chebyshev_operator_matches_oracle(ctx, dt, degree, inner, monkeypatch):
    """T = p(A) applied as a block operator (lb2_op_apply on an lb2_op_chebyshev handle) vs the numpy restatement:
    the step fused into the stencil kernel's epilogue, and SpMM + update kernel for other inner operators."""
    g = (10, 8, 7)
    n = 10 * 8 * 7
    rng = np.random.default_rng(degree)
    X = rand(rng, (n, 5), dt)
    pot = pr.harmonic_potential(g, 0.4) if inner == "stencil+potential" else None
    if inner == "csr":
        monkeypatch.setenv("LB2_CSR_NO_STENCIL_DETECT", "1")
        rp, c, v = pr.laplacian_csr(g, dtype=api.REAL[api.PREFIX[np.dtype(dt)]])
        A = api.csr_op(rp, c, v.astype(dt))
    else:
        A = api.stencil_op(g, dt, potential=pot)
    if inner == "stencil-unfused":
        monkeypatch.setenv("LB2_NO_CHEB_FUSE", "1")
    T = api.chebyshev_op(A, degree, 0.25, 13.0)
    Y = T.apply(ctx, api.DeviceArray.from_numpy(ctx, X)).numpy(ctx)
    ref = no.op_chebyshev(no.op_stencil(g, dt, potential=pot), degree, 0.25, 13.0)(
        X.astype(np.complex128 if np.dtype(dt).kind == "c" else np.float64))
    close(Y, ref, 1e-12 if rtol(dt) < 1e-6 else 1e-4)


# ------------------------------------------------------------------------------------------------ column-block Gram
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [  # (n, mxp, nw): S = [XP | W], products S^H W and S^H (A W)
    (5000, 40, 20), (9000, 600, 300), (20000, 256, 128), (8193, 300, 130), (4099, 0, 150), (30001, 525, 225),
    (2000, 24, 12), (12345, 263, 7)])
def test_gram_cols_matches_numpy(ctx, dt, shape):
    """lb2_<p>_gram_cols: both column-block products of the cached-Gram pass in one call.  Entries strictly below the
    diagonal of the Hermitian W block may be left untouched (checked: they are either exact or still zero); everything
    else must match numpy.  f64 with n >= 4096 runs the work-list kernel (aligned W-block tiles, balanced diagonal tiles,
    ragged edges), the other types one rectangular product each."""
    n, mxp, nw = shape
    rng = np.random.default_rng(n + mxp + nw)
    S = rand(rng, (n, mxp + nw), dt)
    W = np.asfortranarray(S[:, mxp:])
    h = rng.uniform(0.5, 1.5, n).astype(S.real.dtype)
    AW = np.asfortranarray(h[:, None] * W)                      # Hermitian "operator": W^H A W is Hermitian
    dS, dW, dAW = (api.DeviceArray.from_numpy(ctx, a) for a in (S, W, AW))
    for tri in (mxp, -1):
        G0, G1 = api.gram_cols(ctx, dS, dW, dAW, tri_c0=tri)
        for got, ref in ((G0.numpy(ctx), S.conj().T @ W), (G1.numpy(ctx), S.conj().T @ AW)):
            scale = float(np.abs(ref).max())
            err = np.abs(got - ref) / scale
            i, j = np.meshgrid(np.arange(mxp + nw), np.arange(nw), indexing="ij")
            below = (i - mxp > j) if tri >= 0 else np.zeros_like(i, dtype=bool)
            assert err[~below].max() < rtol(dt)
            ok_below = (err < rtol(dt)) | (got == 0)
            assert ok_below[below].all()
    G0, _ = api.gram_cols(ctx, dS, dAW, None, tri_c0=mxp)        # single product (ortho branch)
    got, ref = G0.numpy(ctx), S.conj().T @ AW
    i, j = np.meshgrid(np.arange(mxp + nw), np.arange(nw), indexing="ij")
    keep = ~(i - mxp > j)
    assert (np.abs(got - ref)[keep] / float(np.abs(ref).max())).max() < rtol(dt)


def test_gram_cols_full_size_linearity(ctx):
    """C5 shape (n = 4.096 M rows would need 29 GB of host memory for numpy; use a size-independent property instead):
    device-generated S, W = S[:, mxp:], AW = 2 W  =>  G1 = 2 G0 exactly, and the upper part of the W block of G0 equals
    the Hermitian work-list Gram of W."""
    n, mxp, nw = 1_000_000, 600, 300
    S = api.fill_uniform(ctx, n, mxp + nw, np.float64, 5)
    W = api.DeviceArray((n, nw), np.float64)
    sel = np.zeros((mxp + nw, nw), order="F"); sel[mxp:, :] = np.eye(nw)
    api.tall_nn(ctx, S, api.DeviceArray.from_numpy(ctx, sel), W)
    AW = api.DeviceArray((n, nw), np.float64)
    api.tall_nn(ctx, S, api.DeviceArray.from_numpy(ctx, 2.0 * sel), AW)
    G0, G1 = api.gram_cols(ctx, S, W, AW, tri_c0=mxp)
    g0, g1 = G0.numpy(ctx), G1.numpy(ctx)
    i, j = np.meshgrid(np.arange(mxp + nw), np.arange(nw), indexing="ij")
    keep = ~(i - mxp > j)
    assert np.max(np.abs(2.0 * g0 - g1)[keep]) / np.abs(g1).max() < 1e-13     # (the two products split their rows differently)
    Gww = api.gram(ctx, W, W, upper=True).numpy(ctx)
    up = np.triu_indices(nw)
    blk = g0[mxp:, :]
    assert np.max(np.abs(blk[up] - Gww[up])) / np.abs(Gww).max() < 1e-13


# ------------------------------------------------------------------------------------------------ int8 (Ozaki) f64 Gram
def _wide_columns(rng, n, m):
    """columns whose magnitudes span 14 decades and whose entries span several decades inside a column"""
    X = rng.standard_normal((n, m)) * np.exp(rng.uniform(-3, 3, (n, 1)))
    return np.asfortranarray(X * np.exp(rng.uniform(-16, 16, m))[None, :])


@pytest.mark.parametrize("shape", [(4096, 128, 128, True), (20000, 150, 150, True), (40003, 150, 100, False), (8200, 33, 7, False),
                                   (12000, 257, 257, True), (30001, 5, 300, False)])
def test_gram_i8_matches_extended_precision(ctx, shape):
    """gram_i8.cu (option gram_i8): f64 Gram through tcgen05.mma kind::i8 on a 7-slice int8 split with exact integer
    accumulation.  Error measured against a long-double product, relative to |a_i| |b_j| (columns of very different
    magnitude): must be at the level of the f64 DMMA path (a few 1e-16), ragged tiles, Hermitian mirror, n not a multiple of
    the 128-row chunk."""
    n, ma, mb, upper = shape
    rng = np.random.default_rng(n + ma)
    A = _wide_columns(rng, n, ma)
    B = A if upper else _wide_columns(rng, n, mb)
    ref = (A.astype(np.longdouble).T @ B.astype(np.longdouble)).astype(np.float64)
    scale = np.sqrt(np.outer((A * A).sum(0), (B * B).sum(0)))
    dA = api.DeviceArray.from_numpy(ctx, A)
    dB = dA if upper else api.DeviceArray.from_numpy(ctx, B)
    ctx.set_option("gram_i8", 1)
    try:
        G = api.gram(ctx, dA, dB, upper=upper).numpy(ctx)
    finally:
        ctx.set_option("gram_i8", 0)
    err = np.abs(G - ref) / scale
    assert err.max() < 2e-15, f"max error {err.max():.2e} relative to |a||b|"
    if upper:
        assert np.array_equal(G, G.T)


@pytest.mark.parametrize("shape", [(9000, 600, 300), (8193, 300, 130), (4099, 0, 150), (30001, 525, 225), (12345, 263, 7)])
def test_gram_cols_i8_matches_dmma_path(ctx, shape):
    """Column-block products of the cached-Gram pass through the int8 path: W0 given as a column view of S (the solver's
    B = I case: the slices of S serve both operands) and as a separate block; both products; must agree with numpy and with
    the DMMA work-list kernel on every entry that is not strictly below the diagonal of the Hermitian block."""
    n, mxp, nw = shape
    rng = np.random.default_rng(n + nw)
    S = _wide_columns(rng, n, mxp + nw)
    h = rng.uniform(0.5, 1.5, n)
    AW = np.asfortranarray(h[:, None] * S[:, mxp:])
    dS, dAW = api.DeviceArray.from_numpy(ctx, S), api.DeviceArray.from_numpy(ctx, AW)
    dWsep = api.DeviceArray.from_numpy(ctx, np.asfortranarray(S[:, mxp:]))
    ref0, ref1 = S.T @ S[:, mxp:], S.T @ AW          # f64 BLAS references: tolerance 2e-14 of |a| |b| (their own rounding)
    nrm = np.sqrt((S * S).sum(0))
    sc0 = np.outer(nrm, nrm[mxp:]); sc1 = np.outer(nrm, np.sqrt((AW * AW).sum(0)))
    i, j = np.meshgrid(np.arange(mxp + nw), np.arange(nw), indexing="ij")
    keep = ~(i - mxp > j)
    ctx.set_option("gram_i8", 1)
    try:
        # one tile per CTA in lock-step cohorts (default), with the equal-cost cut, and the 4-CTA cluster kernel (multicast tiles)
        for lockstep, cluster in ((1, 0), (0, 0), (1, 1)):
            ctx.set_option("oz_lockstep", lockstep); ctx.set_option("oz_cluster", cluster)
            for dW in (dS.cols(mxp, nw), dWsep):
                G0, G1 = api.gram_cols(ctx, dS, dW, dAW, tri_c0=mxp)
                assert (np.abs(G0.numpy(ctx) - ref0) / sc0)[keep].max() < 2e-14
                assert (np.abs(G1.numpy(ctx) - ref1) / sc1)[keep].max() < 2e-14
            G0, _ = api.gram_cols(ctx, dS, dAW, None, tri_c0=mxp)
            assert (np.abs(G0.numpy(ctx) - ref1) / sc1)[keep].max() < 2e-14
    finally:
        ctx.set_option("gram_i8", 0); ctx.set_option("oz_lockstep", 1); ctx.set_option("oz_cluster", 0)


@pytest.mark.parametrize("shape", [(4096, 128, 64), (20000, 150, 70), (9000, 900, 300), (8200, 33, 7), (12289, 384, 129), (30001, 600, 44)])
def test_tall_nn_i8_matches_extended_precision(ctx, shape):
    """Projection Out = S C on the int8 tensor path (gram_i8.cu: oz_nn_kernel; S read as an MN-major operand from the slices,
    the column exponents of S folded into the slices of C): error against a long-double product, normwise per output column, at
    f64 level; with slices left by a preceding column-block Gram of the same S (the solver's order) and without."""
    n, kd, nb = shape
    rng = np.random.default_rng(n + kd)
    S = _wide_columns(rng, n, kd)
    Cm = np.asfortranarray(rng.standard_normal((kd, nb)) * np.exp(rng.uniform(-6, 6, (kd, 1))) / np.sqrt((S * S).sum(0))[:, None])
    ref = (S.astype(np.longdouble) @ Cm.astype(np.longdouble)).astype(np.float64)
    # fixed point per COLUMN: the slices carry an entry of S to 2^-55 of its column's largest entry and an entry of
    # C' = diag(colmax S) C to 2^-55 of the largest entry of its column, so the error bound of Out[r, j] is absolute per output
    # column, sum_c colmax_c |C[c, j]| * O(sqrt(kd) 2^-53) — normwise like a backward-stable GEMM, not elementwise
    cmax = np.abs(S).max(0)
    scale = (cmax[:, None] * np.abs(Cm)).sum(0)[None, :]
    dS, dC = api.DeviceArray.from_numpy(ctx, S), api.DeviceArray.from_numpy(ctx, Cm)
    ctx.set_option("gram_i8", 1)
    try:
        for warm in (False, True):
            if warm and kd > nb:                      # leave the slices of S behind, as the Gram of a pass does
                api.gram_cols(ctx, dS, dS.cols(kd - nb, nb), None, tri_c0=kd - nb)
            Out = api.DeviceArray((n, nb), np.float64)
            api.tall_nn(ctx, dS, dC, Out)
            err = np.abs(Out.numpy(ctx) - ref) / scale
            assert err.max() < 1e-14, f"max error {err.max():.2e} relative to the normwise scale (warm={warm})"
            assert np.median(err) < 2e-15
    finally:
        ctx.set_option("gram_i8", 0)


def test_tall_nn_i8_update_on_cached_slices(ctx):
    """U <- U - V (V^H U) the way ortho_drop does it on the int8 path: the rectangular Gram V^H U leaves the slices of V behind,
    the solver vouches that V is unchanged (oz_reuse) and the update Out = alpha V C + beta Out reads them (no second split)."""
    rng = np.random.default_rng(9)
    n, nv, nu = 30000, 200, 70
    V = np.linalg.qr(rng.standard_normal((n, nv)))[0]
    U = rng.standard_normal((n, nu))
    dV, dU = api.DeviceArray.from_numpy(ctx, np.asfortranarray(V)), api.DeviceArray.from_numpy(ctx, np.asfortranarray(U))
    ctx.set_option("gram_i8", 1)
    try:
        Cm = api.gram(ctx, dV, dU)                                   # V^H U, slices of V stay behind
        ctx.set_option("oz_reuse", 1)
        l0 = ctx.launches
        api.tall_nn(ctx, dV, Cm, dU, alpha=-1.0, beta=1.0)
        launches = ctx.launches - l0
        ctx.set_option("oz_reuse", 0)
        got = dU.numpy(ctx)
        ref = U - V @ (V.T @ U)
        assert np.abs(got - ref).max() < 1e-13 * np.abs(U).max()
        assert np.abs(V.T @ got).max() < 1e-12
        assert launches == 2                                          # slices of the small matrix + the projection kernel: no split of V
    finally:
        ctx.set_option("gram_i8", 0); ctx.set_option("oz_reuse", 0)


def test_gram_cols_i8_exponent_hints(ctx):
    """The solver's column-block Gram carries the column exponents from call to call (the split then collects the maxima itself
    instead of a separate pass).  A column that grew past the guard bit or shrank by more than 4 bits must trigger a fresh
    split: results stay at the accuracy of the unhinted path whatever the previous call saw."""
    rng = np.random.default_rng(77)
    n, mxp, nw = 20000, 140, 70
    S = _wide_columns(rng, n, mxp + nw)
    i, j = np.meshgrid(np.arange(mxp + nw), np.arange(nw), indexing="ij")
    keep = ~(i - mxp > j)

    def run(Sm):
        AW = np.asfortranarray(1.5 * Sm[:, mxp:])
        dS, dAW = api.DeviceArray.from_numpy(ctx, np.asfortranarray(Sm)), api.DeviceArray.from_numpy(ctx, AW)
        G0, G1 = api.gram_cols(ctx, dS, dS.cols(mxp, nw), dAW, tri_c0=mxp)
        nrm = np.sqrt((Sm * Sm).sum(0))
        for got, ref, sc in ((G0.numpy(ctx), Sm.T @ Sm[:, mxp:], np.outer(nrm, nrm[mxp:])), (G1.numpy(ctx), Sm.T @ AW, 1.5 * np.outer(nrm, nrm[mxp:]))):
            assert (np.abs(got - ref) / sc)[keep].max() < 2e-14
        return G0.numpy(ctx)

    ctx.set_option("gram_i8", 1)
    try:
        a = run(S)                      # fresh maxima
        b = run(S)                      # hinted (same data: one guard bit coarser digits, same result to rounding)
        assert np.abs(a - b)[keep].max() <= 1e-15 * np.abs(a).max()
        S2 = S.copy(); S2[:, 5] *= 2.0 ** 9; S2[:, mxp + 3] *= 2.0 ** -30; S2[:, 17] *= 1.7
        run(S2)                         # column 5 overflows its hint, column mxp + 3 lost 30 bits: both must be split afresh
        run(S2)
        ctx.set_option("oz_hints", 0)
        c = run(S2)
        ctx.set_option("oz_hints", 1)
        d = run(S2)
        assert np.abs(c - d)[keep].max() <= 1e-15 * np.abs(c).max()
    finally:
        ctx.set_option("gram_i8", 0); ctx.set_option("oz_hints", 1)


def test_int8_path_propagates_non_finite_input(ctx):
    """A NaN or Inf anywhere in an operand cannot be represented by the integer slices: the int8 path flags it while it looks
    for the column maxima and returns NaN outputs (a reference BLAS would propagate it to the affected entries; here the whole
    result is poisoned, which the solver turns into its usual "factorisation failed" exit) instead of silently dropping it."""
    rng = np.random.default_rng(3)
    n, m = 9000, 40
    S = rng.standard_normal((n, m))
    S[1234, 7] = np.nan
    dS = api.DeviceArray.from_numpy(ctx, np.asfortranarray(S))
    Cm = api.DeviceArray.from_numpy(ctx, np.asfortranarray(rng.standard_normal((m, 9))))
    ctx.set_option("gram_i8", 1)
    try:
        assert np.isnan(api.gram(ctx, dS, dS, upper=True).numpy(ctx)).all()
        Out = api.DeviceArray((n, 9), np.float64)
        api.tall_nn(ctx, dS, Cm, Out)
        assert np.isnan(Out.numpy(ctx)).all()
        S[1234, 7] = 1.0                                   # and a clean block afterwards is clean again
        dS2 = api.DeviceArray.from_numpy(ctx, np.asfortranarray(S))
        assert np.isfinite(api.gram(ctx, dS2, dS2, upper=True).numpy(ctx)).all()
    finally:
        ctx.set_option("gram_i8", 0)


# ------------------------------------------------------------------------------------------------ windowed CSR kernel
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", [(3000, 40, 9, 5), (70001, 200, 33, 12), (5000, 256, 7, 40), (1100, 3, 4, 3)])
def test_csr_window_kernel_matches_scipy(ctx, dt, case):
    """csr_win_kernel (banded matrices: X window in shared memory, far couplings gathered): random band of half-width H
    plus far entries, ragged row blocks (n not a multiple of 512), column-group tails, all four types; the plain kernel
    (context option csr_window = 0) must give the same result."""
    import scipy.sparse as sp
    n, H, nc, far = case
    rng = np.random.default_rng(n + H)
    rows, cols = [], []
    for d in (0, 1, -1, H, -H, H // 2):
        i = np.arange(max(0, -d), min(n, n - d))
        rows.append(i); cols.append(i + d)
    rows.append(rng.integers(0, n, far * n // 10)); cols.append(rng.integers(0, n, far * n // 10))   # far / random couplings
    r, c = np.concatenate(rows), np.concatenate(cols)
    vals = rng.standard_normal(len(r)) + (1j * rng.standard_normal(len(r)) if np.dtype(dt).kind == "c" else 0)
    M = sp.csr_matrix((vals.astype(dt), (r, c)), shape=(n, n))
    M.sum_duplicates(); M.sort_indices()
    X = rand(rng, (n, nc), dt)
    ref = np.asarray(M @ X)
    dX = api.DeviceArray.from_numpy(ctx, X)
    op = api.csr_op(M.indptr, M.indices, M.data)
    plain = op.apply(ctx, dX).numpy(ctx)
    close(plain, ref, rtol(dt) * 10)
    ctx.set_option("csr_window", 1)          # opt-in: measured slower than the plain kernel on B200
    try:
        got = op.apply(ctx, dX).numpy(ctx)
    finally:
        ctx.set_option("csr_window", -1)
    close(got, ref, rtol(dt) * 10)


# ------------------------------------------------------------------------------------------------ staged CSR kernel
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", [(3001, 7, 9, "band"), (20000, 30, 17, "rand"), (6000, 90, 5, "rand"), (6000, 3, 6, "long"),
                                  (1500, 5, 3, "empty")])
def test_csr_kernels_match_scipy(ctx, dt, case):
    """General CSR kernels: csr_kernel in its chunked 1-D launch order (default) and on the 2-D grid, and csr_staged_kernel
    (opt-in: the (col, val) range of a row block staged in shared memory, 1 / 4 / 16 lanes per row).  Banded and random
    sparsity, mean row lengths that select every lanes-per-row variant, rows longer than the staging buffer (several
    passes), empty rows and an empty tail, ragged row blocks, ragged last chunk, column-group tails, all four types."""
    import scipy.sparse as sp
    n, per_row, nc, kind = case
    rng = np.random.default_rng(n + per_row)
    if kind == "band":
        rows, cols = [], []
        for d in (0, 1, -1, 40, -40, 777, -777):
            i = np.arange(max(0, -d), min(n, n - d))
            rows.append(i); cols.append(i + d)
        r, c = np.concatenate(rows), np.concatenate(cols)
    else:
        r = rng.integers(0, n, per_row * n); c = rng.integers(0, n, per_row * n)
        if kind == "long":        # three rows with more entries than the staging buffer holds (4096), next to short ones
            for row in (0, 350, n - 1):
                r = np.concatenate([r, np.full(n, row)]); c = np.concatenate([c, np.arange(n)])
            r = np.concatenate([r] + [np.full(5000, 351)]); c = np.concatenate([c, rng.integers(0, n, 5000)])
        if kind == "empty":
            keep = (r % 7 != 3) & (r < n - 300)      # empty rows inside and an empty tail of 300 rows
            r, c = r[keep], c[keep]
    vals = rng.standard_normal(len(r)) + (1j * rng.standard_normal(len(r)) if np.dtype(dt).kind == "c" else 0)
    M = sp.csr_matrix((vals.astype(dt), (r, c)), shape=(n, n))
    M.sum_duplicates(); M.sort_indices()
    X = rand(rng, (n, nc), dt)
    ref = np.asarray(M @ X)
    dX = api.DeviceArray.from_numpy(ctx, X)
    op = api.csr_op(M.indptr, M.indices, M.data)
    tol = rtol(dt) * (400 if kind == "long" else 10)
    close(op.apply(ctx, dX).numpy(ctx), ref, tol)          # default: plain kernel, chunked launch order
    try:
        for order in (0, 1, 3):                             # 2-D grid, and chunks of 1 / 3 row blocks (ragged last chunk)
            ctx.set_option("csr_order", order)
            close(op.apply(ctx, dX).numpy(ctx), ref, tol)
        ctx.set_option("csr_order", 512)
        for pipe in (0, 2):                                 # plain loop and two couplings per step (default 1: next pair ahead)
            ctx.set_option("csr_pipe", pipe)
            close(op.apply(ctx, dX).numpy(ctx), ref, tol)
        ctx.set_option("csr_pipe", 1)
        ctx.set_option("csr_staged", 1)
        for lpr in (1, 4, 16):
            ctx.set_option("csr_lpr", lpr)
            close(op.apply(ctx, dX).numpy(ctx), ref, tol)
        ctx.set_option("csr_lpr", 0)
        ctx.set_option("spmm_cols", 8)
        close(op.apply(ctx, dX).numpy(ctx), ref, tol)
    finally:
        ctx.set_option("csr_lpr", 0); ctx.set_option("spmm_cols", 0); ctx.set_option("csr_staged", 0); ctx.set_option("csr_order", 512); ctx.set_option("csr_pipe", 1)


# ------------------------------------------------------------------------------------------------ TMA-fed f32 Gram
@pytest.mark.parametrize("shape", [(4099, 20, 20), (30001, 64, 33), (20000, 150, 129), (65536, 256, 256), (100000, 300, 100),
                                   (8193, 260, 7)])
def test_gram_float_tma_variant_matches_numpy(ctx, shape):
    """gram_tc5_tma_kernel (the default f32 Gram; context option gram_tma = 0 selects the cp.async-fed kernel): operand panels brought in by cp.async.bulk.tensor (SWIZZLE_128B
    K-major tiles, mbarrier complete_tx) instead of per-thread cp.async; same 3xTF32 arithmetic, so the same tolerance as the
    default kernel, and the two must agree closely with each other."""
    n, ma, mb = shape
    rng = np.random.default_rng(n + ma)
    A, B = rand(rng, (n, ma), np.float32), rand(rng, (n, mb), np.float32)
    dA, dB = api.DeviceArray.from_numpy(ctx, A), api.DeviceArray.from_numpy(ctx, B)
    ref_ab = A.astype(np.float64).T @ B.astype(np.float64)
    ref_aa = A.astype(np.float64).T @ A.astype(np.float64)
    ctx.set_option("gram_tma", 0)            # the cp.async-fed kernel of round 1
    try:
        base_ab = api.gram(ctx, dA, dB).numpy(ctx)
    finally:
        ctx.set_option("gram_tma", -1)       # default: TMA-fed
    got_ab = api.gram(ctx, dA, dB).numpy(ctx)
    got_aa = api.gram(ctx, dA, dA, upper=True).numpy(ctx)
    close(got_ab, ref_ab, 2e-5)
    close(got_aa, ref_aa, 2e-5)
    close(got_ab, base_ab.astype(np.float64), 2e-6)
    assert np.array_equal(got_aa, got_aa.T)
