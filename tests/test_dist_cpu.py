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
