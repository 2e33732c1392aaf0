"""Host-side logic of the row-partitioned multi-GPU path, on CPU: partition arithmetic, and the id / handle
exchange over torch.distributed with the gloo backend at world_size 2."""
import os
import socket

import numpy as np
import pytest

from lobpcg_b200 import dist
from lobpcg_b200 import problems as pr


def test_slab_partition_arithmetic():
    parts = [dist.SlabPartition(160, 160, 160, 8, r) for r in range(8)]
    assert sum(p.n_local for p in parts) == 160 ** 3
    assert [p.z0 for p in parts] == list(range(0, 160, 20))
    assert parts[0].lo is None and parts[0].hi == 1 and parts[7].hi is None and parts[7].lo == 6
    assert parts[3].rows() == slice(3 * 20 * 25600, 4 * 20 * 25600)
    # 25600 x 300 x 8 B = 61 MB per direction at k'=300
    assert parts[3].halo_bytes_per_apply(300) == 2 * 25600 * 300 * 8
    assert parts[0].halo_bytes_per_apply(300) == 25600 * 300 * 8
    with pytest.raises(ValueError):
        dist.SlabPartition(10, 10, 10, 3, 0)


def test_partitioned_rows_reassemble_the_global_block():
    """Every rank generating its slab of X0 with the global counter gives the single-GPU block (the device
    generator uses the same counters: tests/test_gpu_kernels.py::test_fill_uniform...)."""
    g, k, world = (6, 5, 8), 3, 4
    n = int(np.prod(g))
    X = pr.initial_block(n, k, 11)
    pieces = []
    for r in range(world):
        p = dist.SlabPartition(*g, world, r)
        cols = [pr.splitmix_uniform(11, p.n_local, np.float64, start=j * n + p.row0) for j in range(k)]
        pieces.append(np.stack(cols, axis=1))
    assert np.array_equal(np.vstack(pieces), X)


def test_csr_row_blocks_with_relative_columns_reassemble_the_global_product():
    """dist.csr_row_block + the addressing rule of the partitioned CSR kernel (csrc/spmm.cu, HALO): a column index
    relative to the block's first row that is negative / >= n_local is a row of the lower / upper neighbour's block."""
    import scipy.sparse as sp
    g, world = (6, 5, 8), 4
    rp, c, v = pr.laplacian_csr(g, potential=pr.harmonic_potential(g, 0.4))
    n = len(rp) - 1
    M = sp.csr_matrix((v, c, rp), shape=(n, n))
    X = pr.initial_block(n, 3, 2)
    out = []
    for r in range(world):
        part = dist.SlabPartition(*g, world, r)
        nl = part.n_local
        rpl, cg, vl = dist.csr_row_block(rp, c, v, part.row0, nl)
        assert rpl[0] == 0 and rpl[-1] == len(cg) == len(vl)
        rel = cg.astype(np.int64) - part.row0
        assert rel.min() >= (-nl if r > 0 else 0) and rel.max() < (2 * nl if r + 1 < world else nl)
        blocks = {-1: X[part.row0 - nl:part.row0] if r > 0 else None, 0: X[part.rows()],
                  1: X[part.row0 + nl:part.row0 + 2 * nl] if r + 1 < world else None}
        Y = np.zeros((nl, 3))
        for i in range(nl):
            for q in range(rpl[i], rpl[i + 1]):
                side = -1 if rel[q] < 0 else (1 if rel[q] >= nl else 0)
                Y[i] += vl[q] * blocks[side][rel[q] - side * nl]
        out.append(Y)
    assert np.allclose(np.vstack(out), M @ X, atol=1e-13)


def test_bdg_local_rows_cover_both_fields_once():
    g, world = (4, 3, 8), 4
    m = int(np.prod(g))
    rows = np.concatenate([dist.bdg_local_rows(dist.SlabPartition(*g, world, r)) for r in range(world)])
    assert np.array_equal(np.sort(rows), np.arange(2 * m))
    p1 = dist.SlabPartition(*g, world, 1)
    assert dist.bdg_local_rows(p1)[0] == p1.row0 and dist.bdg_local_rows(p1)[p1.n_local] == m + p1.row0


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as td
    td.init_process_group(backend="gloo", rank=rank, world_size=world)
    try:
        uid = dist.exchange_bytes(bytes([7] * 128) if rank == 0 else b"", src=0)
        handles = dist.exchange_bytes(bytes([rank]) * 64)
        part = dist.SlabPartition(4, 4, 8, world, rank)
        # partial Gram sums + all-reduce == global Gram (what csrc/comm.cu does with NCCL)
        import torch
        X = pr.initial_block(part.n_global, 3, 5)
        G = torch.from_numpy(X[part.rows()].T @ X[part.rows()])
        td.all_reduce(G)
        q.put((rank, uid == bytes([7] * 128), [h[0] for h in handles], np.allclose(G.numpy(), X.T @ X), part.lo, part.hi))
    finally:
        td.destroy_process_group()


def test_gloo_world2_exchange_and_partial_gram_allreduce():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == (0, True, [0, 1], True, None, 1)
    assert res[1] == (1, True, [0, 1], True, 0, None)
