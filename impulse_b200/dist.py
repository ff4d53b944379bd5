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
import os

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


class SlabFFT2P2P:
    """Slab-decomposed 2-D complex transform with the exchange FUSED into the column pass: every rank
    publishes its row-FFT output through CUDA IPC, and the column kernels of each rank load their column
    block straight out of the peers' memory over NVLink/NVSwitch (``impulse_fft_cols_from_parts``).  No
    pack, no all-to-all staging buffer; the only collective left is a one-element all-reduce used as a
    stream-ordered barrier ("all row slabs are complete").

    Two row buffers alternate between calls, so one barrier per transform suffices: a rank overwrites the
    buffer its peers read two calls ago, and every peer has passed the barrier of the previous call since.

    The shared buffers are allocated and mapped by the library itself (``impulse_fft_ipc_*``): the peers'
    memory has to be mapped into THIS device's address space for kernels to load from it, which the IPC
    tensors of torch.multiprocessing (opened on the owner's device) do not give.  ``torch.distributed``
    carries the 64-byte handles and the barrier.

    ``pull_chunks = J > 0`` selects the PIPELINED exchange instead: the rank's column block is cut into J chunks; a
    small copy kernel with deep memory-level parallelism (``impulse_fft_gather_parts``, side stream) gathers chunk j+1
    out of all row slabs into a local staging array while the column transform of chunk j runs on local memory
    (main stream) — NVLink and the FFT arithmetic overlap chunk by chunk instead of the FFT kernel waiting on every
    remote load.  ``IMPULSE_FFT_SLAB_PULL`` sets the default (0 = fused peer loads).
    """

    def __init__(self, rows_local: int, cols: int, dtype, group=None, pull_chunks: int | None = None, copy_ctas: int = 0):
        import torch
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if cols % self.world:
            raise ValueError(f"columns ({cols}) must be divisible by the number of ranks ({self.world})")
        if self.world > 8:
            raise ValueError("at most 8 ranks (one NVSwitch box)")
        self.rl, self.c, self.cb = int(rows_local), int(cols), int(cols) // self.world
        self.dtype = dtype
        self.esz = 16 if dtype == torch.complex128 else 8
        self.code = _lib.F64 if dtype == torch.complex128 else _lib.F32
        L = self.L = _lib.lib()
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf_bytes = self.rl * self.c * self.esz
        self.base = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        _lib.check(L.impulse_fft_ipc_alloc(2 * self.buf_bytes, C.byref(self.base), handle))
        gathered = [None] * self.world
        dist.all_gather_object(gathered, bytes(handle), group=group)
        self.peer_base = []   # rank q's two buffers as mapped into this device's address space
        self._opened = []
        for q in range(self.world):
            if q == self.rank:
                self.peer_base.append(self.base.value)
            else:
                p = C.c_void_p()
                hb = (C.c_ubyte * 64).from_buffer_copy(gathered[q])
                _lib.check(L.impulse_fft_ipc_open(hb, C.byref(p)))
                self._opened.append(p)
                self.peer_base.append(p.value)
        self.flag = torch.zeros(1, device=dev)
        self.step = 0
        if pull_chunks is None:
            pull_chunks = int(os.environ.get("IMPULSE_FFT_SLAB_PULL", "0"))
        per16 = 16 // self.esz
        while pull_chunks > 1 and (self.cb % pull_chunks or (self.cb // pull_chunks) % max(per16, 16)):
            pull_chunks -= 1
        if pull_chunks > 0 and (self.cb % per16 or self.c % per16):
            pull_chunks = 0
        self.pull_chunks = max(0, int(pull_chunks))
        self.copy_ctas = int(copy_ctas) or int(os.environ.get("IMPULSE_FFT_SLAB_COPY_CTAS", "0"))
        if self.pull_chunks:
            self.stage = torch.empty((self.world * self.rl, self.cb), dtype=dtype, device=dev)   # the gathered column block
            self.copy_stream = torch.cuda.Stream(device=dev)
            self.ev_rows = torch.cuda.Event()
            self.ev_chunk = [torch.cuda.Event() for _ in range(self.pull_chunks)]
            self.ev_done = torch.cuda.Event()
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def close(self):
        import torch
        import torch.distributed as dist
        if self.base is None:
            return
        torch.cuda.synchronize()
        dist.barrier(group=self.group)      # nobody is still reading
        for p in self._opened:
            self.L.impulse_fft_ipc_close(p)
        self._opened = []
        dist.barrier(group=self.group)      # every mapping is gone before the owner frees
        self.L.impulse_fft_ipc_free(self.base)
        self.base = None

    def __call__(self, x_local, forward: bool = True, fct: float = 1.0, out=None):
        import torch
        import torch.distributed as dist
        L = self.L
        k = self.step & 1
        self.step += 1
        stream = C.c_void_p(torch.cuda.current_stream(x_local.device).cuda_stream)
        # 1. row FFTs of the local slab straight into the published buffer
        shape = (C.c_size_t * 2)(self.rl, self.c)
        sin = (C.c_ssize_t * 2)(x_local.stride(0) * self.esz, x_local.stride(1) * self.esz)
        sout = (C.c_ssize_t * 2)(self.c * self.esz, self.esz)
        axes = (C.c_size_t * 1)(1)
        rows_ptr = self.base.value + k * self.buf_bytes
        _lib.check(L.impulse_fft_c2c(self.code, 2, shape, sin, sout, 1, axes, int(forward), x_local.data_ptr(), rows_ptr,
                                     float(fct), 0, stream))
        # 2. stream-ordered barrier: completes once every rank's rows are written
        dist.all_reduce(self.flag, group=self.group)
        # 3. column transforms, loading the column block from all row slabs (local + peers over NVLink)
        if out is None:
            out = torch.empty((self.world * self.rl, self.cb), dtype=self.dtype, device=x_local.device)
        parts = (C.c_void_p * self.world)(*[self.peer_base[q] + k * self.buf_bytes for q in range(self.world)])
        if not self.pull_chunks:
            _lib.check(L.impulse_fft_cols_from_parts(self.code, self.world, parts, self.rl, self.c, self.rank * self.cb, self.cb,
                                                     out.data_ptr(), self.cb, int(forward), 1.0, stream))
            return out
        # pipelined exchange: gather chunk j (copy stream) || column transform of chunk j-1 on local memory (main stream)
        main = torch.cuda.current_stream(x_local.device)
        self.ev_rows.record(main)                       # rows written everywhere (the all-reduce above has completed here)
        self.copy_stream.wait_event(self.ev_rows)
        self.copy_stream.wait_event(self.ev_done)       # the previous call's column transforms have read the staging array
        cstream = C.c_void_p(self.copy_stream.cuda_stream)
        J, cw, R = self.pull_chunks, self.cb // self.pull_chunks, self.world * self.rl
        cshape = (C.c_size_t * 2)(R, cw)
        cst = (C.c_ssize_t * 2)(self.cb * self.esz, self.esz)
        caxes = (C.c_size_t * 1)(0)
        for j in range(J):
            _lib.check(L.impulse_fft_gather_parts(self.code, self.world, parts, self.rl, self.c, self.rank * self.cb + j * cw, cw,
                                                  self.stage.data_ptr() + j * cw * self.esz, self.cb, self.copy_ctas, cstream))
            self.ev_chunk[j].record(self.copy_stream)
        for j in range(J):
            main.wait_event(self.ev_chunk[j])
            off = j * cw * self.esz
            _lib.check(L.impulse_fft_c2c(self.code, 2, cshape, cst, cst, 1, caxes, int(forward), self.stage.data_ptr() + off,
                                         out.data_ptr() + off, 1.0, 0, stream))
        self.ev_done.record(main)
        return out


class DistPlan:
    """The single-process multi-GPU driver of the C ABI (``impulse_fft_dist_create / execute / execute_parts /
    destroy``, include/impulse_fft_b200.h): what a Nim / C host with several B200s calls — no torch.distributed, no
    second process.  ``mode`` is "batch" (dimension 0 split over the devices, no communication) or "slab" (one 2-D
    complex transform, row slabs, exchange fused into the column kernels over peer memory).  Host numpy arrays in and
    out for ``__call__``; per-device CUDA tensors for ``run_parts``."""

    MODES = {"batch": 0, "slab": 1}

    def __init__(self, mode: str, kind: int, dtype_code: int, shape, stride_in, stride_out, axes, forward: bool, devices,
                 layout: int = _lib.HERMITIAN):
        L = self.L = _lib.lib()
        d = _lib.Desc()
        d.kind, d.dtype, d.real_layout, d.forward = kind, dtype_code, layout, int(forward)
        d.ndim, d.naxes = len(shape), len(axes)
        for i, (s, a, b) in enumerate(zip(shape, stride_in, stride_out)):
            d.shape[i], d.stride_in[i], d.stride_out[i] = s, a, b
        for i, a in enumerate(axes):
            d.axes[i] = a
        self.devices = [int(x) for x in devices]
        devs = (C.c_int * len(self.devices))(*self.devices)
        self.h = C.c_void_p()
        _lib.check(L.impulse_fft_dist_create(C.byref(self.h), self.MODES[mode], C.byref(d), len(self.devices), devs))

    def shard(self, index: int):
        lo, hi = C.c_size_t(), C.c_size_t()
        _lib.check(self.L.impulse_fft_dist_shard(self.h, index, C.byref(lo), C.byref(hi)))
        return int(lo.value), int(hi.value)

    def __call__(self, a_in, a_out, fct: float = 1.0):
        _lib.check(self.L.impulse_fft_dist_execute(self.h, a_in.ctypes.data, a_out.ctypes.data, float(fct)))
        return a_out

    def run_parts(self, ins, outs, fct: float = 1.0):
        n = len(self.devices)
        pi = (C.c_void_p * n)(*[t.data_ptr() if t is not None else None for t in ins])
        po = (C.c_void_p * n)(*[t.data_ptr() if t is not None else None for t in outs])
        _lib.check(self.L.impulse_fft_dist_execute_parts(self.h, pi, po, float(fct)))
        return outs

    def close(self):
        if self.h is not None and self.h.value:
            self.L.impulse_fft_dist_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 — interpreter shutdown
            pass
