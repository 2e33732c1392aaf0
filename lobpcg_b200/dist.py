"""Row-partitioned multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous only.

The reference is single-address-space; the partition is new.  Every tall block (X, P, W, AX,
work blocks) is split into contiguous z-slabs of the grid, one per rank:

  * Gram partial sums and column norms -> NCCL all-reduce inside the C library (csrc/comm.cu); the unique id
    is created by rank 0 and broadcast here;
  * stencil halo planes -> each rank exports its solver arena through CUDA IPC, the two z-neighbours map it,
    and the stencil kernel reads the neighbour's boundary plane directly over NVLink (no exchange pass).

Host logic (partition arithmetic, id/handle exchange) is backend-agnostic and is covered on CPU with gloo
(tests/test_dist_cpu.py); anything touching CUDA needs a GPU box.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np


@dataclass
class SlabPartition:
    """z-slab partition of a gx*gy*gz grid over `world` ranks (equal slabs: gz % world == 0)."""
    gx: int
    gy: int
    gz: int
    world: int
    rank: int

    def __post_init__(self):
        if self.world < 1 or not (0 <= self.rank < self.world):
            raise ValueError("bad rank/world")
        if self.gz % self.world != 0:
            raise ValueError(f"gz={self.gz} must be divisible by the number of ranks ({self.world}): the peer-halo "
                             "addressing assumes identical arena layouts on every rank")

    @property
    def gz_local(self) -> int:
        return self.gz // self.world

    @property
    def z0(self) -> int:
        return self.rank * self.gz_local

    @property
    def plane(self) -> int:
        return self.gx * self.gy

    @property
    def n_global(self) -> int:
        return self.plane * self.gz

    @property
    def n_local(self) -> int:
        return self.plane * self.gz_local

    @property
    def row0(self) -> int:
        return self.plane * self.z0

    @property
    def lo(self):
        return self.rank - 1 if self.rank > 0 else None

    @property
    def hi(self):
        return self.rank + 1 if self.rank + 1 < self.world else None

    def rows(self) -> slice:
        return slice(self.row0, self.row0 + self.n_local)

    def halo_bytes_per_apply(self, ncols: int, itemsize: int = 8) -> int:
        """bytes this rank reads from its neighbours per block apply"""
        nb = (self.lo is not None) + (self.hi is not None)
        return nb * self.plane * ncols * itemsize


def nccl_library_path() -> str:
    """The NCCL that torch bundles (torch loads the same SONAME, so the loader hands back that copy)."""
    try:
        import nvidia.nccl as m
        p = Path(m.__path__[0]) / "lib" / "libnccl.so.2"
        if p.exists():
            return str(p)
    except Exception:
        pass
    return "libnccl.so.2"


def exchange_bytes(payload: bytes, src: int | None = None):
    """broadcast (src given) or all-gather (src None) of a small byte string through torch.distributed;
    works with gloo (CPU) and nccl backends."""
    import torch
    import torch.distributed as td
    if src is not None:
        box = [payload if td.get_rank() == src else None]
        td.broadcast_object_list(box, src=src)
        return box[0]
    out = [None] * td.get_world_size()
    td.all_gather_object(out, payload)
    return out


def init_process_group(ctx, backend: str = "nccl"):
    """torch.distributed rendezvous (env:// with MASTER_ADDR/PORT from torchrun) + NCCL communicator inside
    the C library, attached to `ctx`."""
    import torch.distributed as td
    from . import api
    if not td.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group(backend=backend)
    rank, world = td.get_rank(), td.get_world_size()
    path = nccl_library_path().encode()
    buf = C.create_string_buffer(128)
    if rank == 0:
        api._ck(api.lib().lb2_comm_unique_id(buf, path), "lb2_comm_unique_id")
    uid = exchange_bytes(buf.raw, src=0)
    api._ck(api.lib().lb2_ctx_attach_comm(ctx.h, rank, world, uid, path), "lb2_ctx_attach_comm")
    return rank, world


def partitioned_stencil(ctx, part: SlabPartition, dtype, k: int, potential=None):
    from . import api
    pot = None if potential is None else np.asarray(potential)[part.rows()]
    return api.stencil_slab_op((part.gx, part.gy, part.gz), part.z0, part.gz_local, dtype, potential_local=pot)


def partitioned_bdg(part: SlabPartition, dtype, shift, d):
    """This rank's part of the BdG operator A = [[K+shift, d], [conj d, K+shift]] (config C4): both fields are split by
    the same z-slabs, so the local rows are [u slab ; v slab] and the block coupling stays rank-local."""
    from . import api
    return api.bdg_slab_op((part.gx, part.gy, part.gz), part.z0, part.gz_local, dtype, shift, d)


def bdg_local_rows(part: SlabPartition) -> np.ndarray:
    """Global row indices of this rank's local rows of a BdG block vector (u slab, then v slab)."""
    m = part.n_global
    return np.concatenate([np.arange(part.row0, part.row0 + part.n_local),
                           m + np.arange(part.row0, part.row0 + part.n_local)])


def csr_row_block(rowptr, col, val, row0: int, n_local: int):
    """Rows [row0, row0 + n_local) of a CSR matrix: (rowptr_local starting at 0, global column indices, values)."""
    rowptr = np.asarray(rowptr)
    p0, p1 = int(rowptr[row0]), int(rowptr[row0 + n_local])
    return (rowptr[row0:row0 + n_local + 1] - p0).astype(np.int64), np.asarray(col)[p0:p1], np.asarray(val)[p0:p1]


def partitioned_csr(part, rowptr, col, val):
    """This rank's row block of a global CSR matrix as a device operator (``part``: anything with n_global, row0,
    n_local — e.g. SlabPartition).  Couplings may reach at most into the two neighbouring row blocks."""
    from . import api
    rp, c, v = csr_row_block(rowptr, col, val, part.row0, part.n_local)
    return api.csr_slab_op(part.n_global, part.row0, rp, c, v)


def global_spec_hi(op) -> float:
    """Maximum over ranks of the operators' local spectrum bounds (the window of a polynomial preconditioner must be
    the same on every rank)."""
    import torch
    import torch.distributed as td
    from . import api
    t = torch.tensor([api.spec_hi(op)], dtype=torch.float64)
    if td.get_backend() == "nccl":
        t = t.cuda()
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def attach(solver, part: SlabPartition):
    """Allocate the solver arena, exchange its CUDA-IPC handle with the z-neighbours and register the mapped
    peer bases.  Collective over all ranks; call before solver.init()."""
    from . import api
    solver.prepare()
    ptr, nbytes = solver.arena()
    h = C.create_string_buffer(64)
    api._ck(api.lib().lb2_ipc_get_handle(ptr, h), "lb2_ipc_get_handle")
    handles = exchange_bytes(h.raw)
    peers = {}
    for nb in (part.lo, part.hi):
        if nb is not None:
            p = api.lib().lb2_ipc_open_handle(handles[nb])
            if not p:
                raise api.LobpcgB200Error(f"cannot map the arena of rank {nb} (CUDA IPC / NVLink P2P unavailable?)")
            peers[nb] = p
    solver.set_peers(peers.get(part.lo), peers.get(part.hi))
    solver._peer_maps = peers
    return peers


def shutdown(ctx):
    import torch.distributed as td
    from . import api
    api.lib().lb2_ctx_detach_comm(ctx.h)
    if td.is_initialized():
        td.barrier()
        td.destroy_process_group()
