"""Multi-GPU drivers: one process per GPU, ``torch.distributed`` for the plumbing.

The reference has no distributed code at all; its closest notion is splitting the
independent 1-D lines of ``general_nd`` over a thread pool (pocketfft_hdronly.h:2778-2800,
3024-3046).  Two partitionings exist here:

* batch sharding  — independent rows/images are split contiguously over the ranks; every rank
  runs the ordinary single-GPU call on its shard.  No collective on the data path.
* slab fft2       — one 2-D complex transform split by row slabs.  Row FFTs locally, one
  all-to-all that turns row slabs into column slabs, column FFTs locally.  The result is left in
  column-slab layout unless ``restore=True`` (a second all-to-all).

The arithmetic is delegated to an *engine* object so that the host logic (partitioning, packing,
exchange) is testable on CPU with the gloo backend; the product engine is ``CudaEngine`` which
calls the C ABI.  Nothing here falls back to a CPU transform.
"""
from __future__ import annotations

import ctypes as C

from . import _lib


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous split: rank g gets rows [g*B/G, (g+1)*B/G)."""
    lo = n_rows * rank // world
    hi = n_rows * (rank + 1) // world
    return lo, hi


class CudaEngine:
    """Engine backed by libimpulse_fft_b200.so on the current CUDA device."""

    def c2c_axis(self, x, out, axis: int, forward: bool, fct: float = 1.0):
        from .desc import DataDesc, FFTDesc
        FFTDesc.init(axes=[axis], forward=forward, scalingFactor=fct).apply(DataDesc.init(out), DataDesc.init(x))
        return out

    def pack_blocks(self, x, out, nblocks: int):
        """x: [R, C] row slab -> out: [nblocks, R, C/nblocks] contiguous per-peer blocks."""
        import torch
        r, c = x.shape
        cb = c // nblocks
        code = _lib.F64 if x.dtype == torch.complex128 else _lib.F32
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.lib().impulse_fft_copy2d(code, x.data_ptr(), out.data_ptr(), r, cb, c, cb, nblocks, cb, r * cb,
                                                 C.c_void_p(stream)))
        return out

    def unpack_blocks(self, x, out, nblocks: int):
        """x: [nblocks, R, Cb] -> out: [R, nblocks*Cb] (inverse of pack_blocks)."""
        import torch
        nb, r, cb = x.shape
        code = _lib.F64 if x.dtype == torch.complex128 else _lib.F32
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.lib().impulse_fft_copy2d(code, x.data_ptr(), out.data_ptr(), r, cb, cb, nb * cb, nb, r * cb, cb,
                                                 C.c_void_p(stream)))
        return out

    def empty(self, shape, like):
        import torch
        return torch.empty(shape, dtype=like.dtype, device=like.device)


def fft_rows_sharded(x_local, forward: bool = True, fct: float = 1.0, engine=None):
    """Batch-sharded 1-D transform along the last axis: each rank transforms the rows it holds.
    No communication; provided for symmetry and for the scaling benchmark."""
    engine = engine or CudaEngine()
    out = engine.empty(tuple(x_local.shape), x_local)
    return engine.c2c_axis(x_local, out, x_local.ndim - 1, forward, fct)


def fft2_slab(x_local, forward: bool = True, fct: float = 1.0, group=None, engine=None, restore: bool = False):
    """2-D complex transform of an [R, C] array held as row slabs [R/P, C] on P ranks.

    Returns the column slab [R, C/P] of the result (rank p holds columns [p*C/P, (p+1)*C/P)), or
    with ``restore=True`` the row slab [R/P, C].  R and C must be divisible by P.

      1. rows:     FFT along axis 1 of the local slab                          (local, HBM-bound)
      2. pack:     [R/P, C] -> [P, R/P, C/P] contiguous per-peer blocks        (local copy)
      3. exchange: all_to_all_single — block q goes to rank q                  (NVLink, (P-1)/P of the slab)
      4. columns:  the received [P, R/P, C/P] IS the column slab [R, C/P];
                   FFT along axis 0                                            (local, strided -> two launches)
    """
    import torch.distributed as dist
    engine = engine or CudaEngine()
    world = dist.get_world_size(group)
    rl, c = x_local.shape
    if c % world:
        raise ValueError(f"columns ({c}) must be divisible by the number of ranks ({world})")
    cb = c // world
    rows = engine.c2c_axis(x_local, engine.empty((rl, c), x_local), 1, forward, fct)
    if world == 1:
        return engine.c2c_axis(rows, rows, 0, forward, 1.0)
    send = engine.pack_blocks(rows, engine.empty((world, rl, cb), x_local), world)
    recv = engine.empty((world, rl, cb), x_local)
    dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
    cols = recv.view(world * rl, cb)
    engine.c2c_axis(cols, cols, 0, forward, 1.0)
    if not restore:
        return cols
    # back to row slabs: rank q needs rows [q*rl, (q+1)*rl) of every column slab
    back = engine.empty((world, rl, cb), x_local)
    dist.all_to_all_single(back.view(-1), cols.view(-1), group=group)
    return engine.unpack_blocks(back, engine.empty((rl, c), x_local), world)
