"""FFT-based 2-D filtering (BASELINE config 5): per image  r2c -> pointwise multiply -> c2r.

The reference has no FFT filter (its `benchmarks/image_filters` is a spatial 3x3 filter,
SURVEY fact 6); this composes the path's own entry points the way a user of the C++ backend would:
`pocketfft::r2c` over the last two axes (pocketfft_hdronly.h:3334-3349), a complex multiply with
the kernel's spectrum, and `pocketfft::c2r` with fct = 1/(H*W) (pocketfft_hdronly.h:3366-3390).
Convolution is CIRCULAR at the image size (primary figure); the kernel is
zero-padded to the image size with its centre wrapped to (0, 0).
"""
from __future__ import annotations

import ctypes as C

from . import _buffers as B
from . import _lib
from .desc import DataDesc, FFTDesc


class FFTFilter2D:
    """Circular 2-D convolution of a batch of real images [B, H, W] with one small real kernel."""

    def __init__(self, kernel, height: int, width: int):
        import torch
        if not (B.is_torch(kernel) and kernel.is_cuda):
            raise TypeError("kernel must be a CUDA tensor")
        kh, kw = kernel.shape
        if kh > height or kw > width:
            raise ValueError("kernel larger than the image")
        self.h, self.w = int(height), int(width)
        self.rdtype = kernel.dtype
        self.cdtype = torch.complex64 if kernel.dtype == torch.float32 else torch.complex128
        self.code = _lib.F32 if kernel.dtype == torch.float32 else _lib.F64
        # zero-pad, centre at (0,0): element (i,j) of the kernel goes to ((i-kh//2) mod H, (j-kw//2) mod W)
        pad = torch.zeros((self.h, self.w), dtype=kernel.dtype, device=kernel.device)
        ii = (torch.arange(kh, device=kernel.device) - kh // 2) % self.h
        jj = (torch.arange(kw, device=kernel.device) - kw // 2) % self.w
        pad[ii[:, None], jj[None, :]] = kernel
        # half spectra live in rows padded to a multiple of four bins: 32-byte aligned runs for the column kernel's
        # 4-line tiles and 16-byte aligned rows for the staged c2r rows; the multiplier is padded alike, because it is
        # indexed by element offset
        wc = self.w // 2 + 1
        self.pitch = (wc + 3) // 4 * 4
        self._spectrum_buf = torch.zeros((self.h, self.pitch), dtype=self.cdtype, device=kernel.device)
        self.spectrum = self._spectrum_buf[:, :wc]
        FFTDesc.init(axes=[0, 1], forward=True).apply(DataDesc.init(self.spectrum), DataDesc.init(pad))
        self._spec_buf = None

    def apply(self, images, out=None):
        """images: [B, H, W] real CUDA tensor (same dtype as the kernel). Returns the filtered batch."""
        import torch
        if images.ndim != 3 or images.shape[1] != self.h or images.shape[2] != self.w:
            raise ValueError("images must have shape [B, H, W]")
        if images.dtype != self.rdtype:
            raise TypeError("image dtype differs from the kernel dtype")
        b = images.shape[0]
        wc = self.w // 2 + 1
        if self._spec_buf is None or self._spec_buf.shape[0] != b:
            self._spec_buf = torch.empty((b, self.h, self.pitch), dtype=self.cdtype, device=images.device)
        spec = self._spec_buf[:, :, :wc]
        if out is None:
            out = torch.empty_like(images)
        # rows: real -> half spectrum; columns: FFT -> x filter spectrum -> inverse FFT as one axis convolution
        # (one pass with the whole column in shared memory up to 4096 rows, three passes beyond; the 2-D spectrum is
        # never materialised); rows: half spectrum -> real
        FFTDesc.init(axes=[2], forward=True).apply(DataDesc.init(spec), DataDesc.init(images))
        stream = C.c_void_p(torch.cuda.current_stream(images.device).cuda_stream)
        esz = spec.element_size()
        shape = (C.c_size_t * 3)(b, self.h, wc)
        st = (C.c_ssize_t * 3)(self.h * self.pitch * esz, self.pitch * esz, esz)
        _lib.check(_lib.lib().impulse_fft_convolve_axis(self.code, 3, shape, st, st, 1, spec.data_ptr(), spec.data_ptr(), 1.0,
                                                        self._spectrum_buf.data_ptr(), self.h * self.pitch, stream))
        FFTDesc.init(axes=[2], forward=False, scalingFactor=1.0 / (self.h * self.w)).apply(
            DataDesc.init(out), DataDesc.init(spec))
        return out
