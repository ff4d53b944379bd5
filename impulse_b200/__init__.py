"""impulse_b200 — B200-native FFT engine behind the FFT API of SciNim/impulse.

Python host layer over the C ABI (libimpulse_fft_b200.so).  Two API mirrors:
  * the C backend:   fft, ifft, rfft, rfft_packed, fft_inplace, unpackFFT, symmetrize, Normalize kinds
  * the C++ backend: DataDesc, FFTDesc, apply
plus the multi-GPU drivers in impulse_b200.dist.
"""
from ._lib import FFTError, bind_host_to_device, last_kernel, launch_count  # noqa: F401
from .desc import (DataDesc, DCTDesc, FFTDesc, apply, r2r_fftpack, r2r_genuine_hartley,  # noqa: F401
                   r2r_separable_hartley)
from .fft import (fft, fft_inplace, ifft, initNormalize, isOdd, nkBackward, nkCustom, nkForward,  # noqa: F401
                  nkOrtho, rfft, rfft_packed, symmetrize, symmTargetSize, unpackFFT)

__all__ = ["fft", "ifft", "rfft", "rfft_packed", "fft_inplace", "unpackFFT", "symmetrize", "symmTargetSize",
           "initNormalize", "isOdd", "nkBackward", "nkOrtho", "nkForward", "nkCustom",
           "DataDesc", "FFTDesc", "DCTDesc", "apply", "r2r_fftpack", "r2r_separable_hartley", "r2r_genuine_hartley", "FFTError", "launch_count", "last_kernel", "bind_host_to_device"]
