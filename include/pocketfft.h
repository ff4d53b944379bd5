/* pocketfft.h — the ten C symbols of the reference's C backend, exported by
 * libimpulse_fft_b200.so so that impulse/fft/c_pocketfft/pocketfft.nim:71-81 can bind the GPU
 * engine without source changes beyond dropping `{.compile: "pocketfft.c".}`.
 *
 * Interface replaced: /root/reference/impulse/fft/c_pocketfft/pocketfft.h:18-32 (declarations)
 * implemented there by pocketfft.c:2060-2190.  Semantics kept:
 *   - in place on `c`; unnormalised; the result is multiplied by `fct`;
 *   - forward uses exp(-2 pi i jk/N) (pocketfft.c:946-951);
 *   - real transforms use the FFTPACK halfcomplex packing (pocketfft.nim:228-238);
 *   - make_*_plan returns NULL for length 0 or on failure (pocketfft.c:2068-2070);
 *   - execute returns 0 on success and -1 on failure (pocketfft.c:878,1707,1952);
 *   - plans are immutable and may be shared between threads (c_pocketfft/README.md:32-36).
 * Difference: `c` may also be a device pointer; host pointers are staged through the GPU.
 * The plan structs are opaque (the Nim mirror of their fields, pocketfft.nim:21-69, is never
 * dereferenced).
 */
#ifndef IMPULSE_B200_POCKETFFT_H
#define IMPULSE_B200_POCKETFFT_H

#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

struct cfft_plan_i;
typedef struct cfft_plan_i *cfft_plan;
cfft_plan make_cfft_plan(size_t length);
void destroy_cfft_plan(cfft_plan plan);
int cfft_backward(cfft_plan plan, double c[], double fct);
int cfft_forward(cfft_plan plan, double c[], double fct);
size_t cfft_length(cfft_plan plan);

struct rfft_plan_i;
typedef struct rfft_plan_i *rfft_plan;
rfft_plan make_rfft_plan(size_t length);
void destroy_rfft_plan(rfft_plan plan);
int rfft_backward(rfft_plan plan, double c[], double fct);
int rfft_forward(rfft_plan plan, double c[], double fct);
size_t rfft_length(rfft_plan plan);

#ifdef __cplusplus
}
#endif
#endif
