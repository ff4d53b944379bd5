// Shared host/device description of one batched line-FFT launch ("LineJob").
//
// A LineJob is a small phase program interpreted by every CTA of the generic engine
// (fft_device.cuh): LOAD a tile of `C` lines into shared memory, run in-place radix
// passes (and the Bluestein / real pre- and post-steps) there, STORE the tile.
// The host planner (planner.cpp) builds it; the kernel receives it by value as a
// __grid_constant__ parameter.
//
// Replaces, for the GPU, the reference's plan structs cfftp_plan_i / rfftp_plan_i /
// fftblue_plan_i (c_pocketfft/pocketfft.c:267-279, 1060-1072, 1880-1887) and the
// per-axis line iteration of general_nd (cpp_pocketfft/pocketfft_hdronly.h:3011-3050).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define IMP_HD __host__ __device__ __forceinline__
#else
#define IMP_HD inline
#endif

namespace impulse {

constexpr int kMaxPhases = 40;
constexpr int kMaxBatchDims = 3;
constexpr int kSmemHeaderBytes = 2048;  // per-line header: in offset, out offset, four-step index (3 x int64 x 64 lines)
constexpr int kMaxLinesPerCta = 64;
constexpr int kMaxThreads = 256;
constexpr int kMaxThreadsBig = 512;  // tiles that leave room for one CTA per SM only
constexpr int kMinCtasPerSm = 3;
constexpr uint32_t kMaxGenericRadix = 31;  // odd prime radices up to this are direct; above -> Bluestein

enum PhaseOp : uint8_t {
  OP_PASS_DIF = 1,      // in-place decimation-in-frequency radix pass (natural in -> digit-reversed out)
  OP_PASS_DIT = 2,      // transposed pass (digit-reversed in -> natural out)
  OP_BLUE_PRE = 3,      // s[n] = s[n]*conj(bk[n]) (n<L), 0 (L<=n<n_fft)        pocketfft.c:1954-1968
  OP_BLUE_MUL = 4,      // s[p] = conj(s[p]*bkf[p])                              pocketfft.c:1973-1987
  OP_BLUE_POST = 5,     // s[k] = conj(s[k]*bk[k]) (k<L)                         pocketfft.c:1993-2005
  OP_C2R_PRE_EVEN = 6,  // half-spectrum X[0..M] -> packed complex Z[0..M-1]
  OP_MUL_CONJ_BK = 7,   // s[k] = s[k]*conj(bk[k]) (k<L): last step of the multi-launch Bluestein
};

enum LoadMode : uint8_t {
  LD_C = 0,          // complex line, n_seq elements
  LD_R_PAIRS = 1,    // even-N real line viewed as N/2 complex (x[2m], x[2m+1])
  LD_R_ZEROIM = 2,   // real line as complex with zero imaginary part
  LD_HERM_EVEN = 3,  // c2r even N: X[0..N/2] verbatim (imag of bins 0 and N/2 dropped)
  LD_HERM_FULL = 4,  // c2r odd N: X[0..(N-1)/2] Hermitian-extended to N complex
  LD_HC_EVEN = 5,    // FFTPACK halfcomplex reals -> X[0..N/2]   (pocketfft.nim:228-238)
  LD_HC_FULL = 6,    // FFTPACK halfcomplex reals -> Hermitian-extended N complex
  // real-to-real transforms (DCT/DST I-IV, pocketfft_hdronly.h:2424-2648) embedded in a complex FFT of M points
  LD_X_ZPAD = 7,     // u[n] = x[n] (n<N), 0 above                          DCT-II, DST-II   (M = 2N)
  LD_X_TW = 8,       // u[n] = f_n x[n] W_4N^n (n<N), 0 above               DCT-III, DCT-IV, DST-IV
  LD_X_TW_SHIFT = 9, // u[m] = f_m x[m-1] W_4N^m (1<=m<=N), 0 elsewhere     DST-III
  LD_X_SYM = 10,     // even extension, M = 2(N-1)                          DCT-I
  LD_X_ASYM = 11,    // odd extension,  M = 2(N+1)                          DST-I
};

enum StoreMode : uint8_t {
  ST_C = 0,          // complex line, N elements
  ST_HERM_HALF = 1,  // bins 0..N/2 of a length-N complex result (odd-N r2c)
  ST_R2C_EVEN = 2,   // even-N r2c: bins 0..N/2 from the N/2-point complex FFT + post-twiddle
  ST_R_PAIRS = 3,    // c2r even: z[m] -> x[2m], x[2m+1]
  ST_R_REALPART = 4, // c2r odd: real parts
  ST_HC_EVEN = 5,    // even-N r2c written as FFTPACK halfcomplex reals
  ST_HC_FULL = 6,    // odd-N r2c written as FFTPACK halfcomplex reals
  ST_R2C_EVEN_SYM = 7,   // even-N r2c, all N bins (Hermitian half + conjugate mirror): fused `symmetrize`
  ST_HERM_SYM = 8,       // odd-N r2c, all N bins                                   (pocketfft.nim:160-171)
  ST_X = 9,              // real-to-real: y[k] = s * (Re | -Im)(W_8N^(mul*k+add) * F[k+shift])
  ST_HARTLEY_EVEN = 10,  // even-N real line -> y[k] = Re X[k] + Im X[k], y[N-k] = Re X[k] - Im X[k]
  ST_HARTLEY_FULL = 11,  // odd-N, same combination                       (pocketfft_hdronly.h:3066-3095)
};

enum JobFlags : uint32_t {
  F_CONJ_IN = 1u << 0,      // conjugate values as read from global (c2r with forward=true)
  F_CONJ_SEQ = 1u << 1,     // conjugate the sequence placed in smem (backward via conj trick)
  F_CONJ_OUT = 1u << 2,     // conjugate the FFT result as read from smem (backward via conj trick)
  F_CONJ_RESULT = 1u << 3,  // conjugate the stored spectrum (r2c with forward=false)
  F_IN_LINES_FAST = 1u << 4,   // global reads: consecutive threads walk adjacent lines (strided axis)
  F_OUT_LINES_FAST = 1u << 5,  // global writes: same
  F_VEC_IN = 1u << 6,       // LD_R_PAIRS may use one 2-element vector load
  F_VEC_OUT = 1u << 7,      // ST_R_PAIRS may use one 2-element vector store
  F_NEG_EVEN_IN = 1u << 8,  // negate real input elements 2, 4, 6, ... (ExecR2R, pocketfft_hdronly.h:3134-3136)
  F_NEG_EVEN_OUT = 1u << 9, // negate real output elements 2, 4, 6, ...       (pocketfft_hdronly.h:3138-3140)
};

// specialised register-resident kernels (fast_kernels.cu); 0 = generic phase interpreter
enum FastId : uint32_t {
  FAST_NONE = 0,
  FAST2_1024_F64 = 1,
  FAST2_512_F64 = 2,
  FAST2_256_F64 = 3,
  FAST2_1024_F32 = 4,
  FAST3_2048_F64 = 5,
  FAST3_4096_F64 = 6,
  FAST3_8192_F64 = 7,
  FAST3_2048_F32 = 8,
  FAST3_4096_F32 = 9,
  FAST3_500_F64 = 10,
  FAST3_1944_F64 = 11,
  FAST3_1000_F64 = 12,
  COL2_64_F64 = 13,
  COL2_128_F64 = 14,
  COL2_256_F64 = 15,
  COL2_64_F32 = 16,
  COL2_128_F32 = 17,
  COL2_256_F32 = 18,
  COL2_32_F64 = 19,
  COL2_512_F64 = 20,
  COL2_32_F32 = 21,
  COL2_512_F32 = 22,
  FASTBLUE_2048_F64 = 23,
  FASTBLUE_4096_F64 = 24,
  FASTBLUE_8192_F64 = 25,
  COLCONV_32_F64 = 26,   // fused middle pass of a convolution along a strided axis (colconv2_kernel)
  COLCONV_64_F64 = 27,
  COLCONV_128_F64 = 28,
  COLCONV_32_F32 = 29,
  COLCONV_64_F32 = 30,
  COLCONV_128_F32 = 31,
  FAST2_16_F64 = 32,     // short contiguous rows, several rows per warp (fast2p_kernel / fast2_kernel)
  FAST2_32_F64 = 33,
  FAST2_64_F64 = 34,
  FAST2_128_F64 = 35,
  FAST2_16_F32 = 36,
  FAST2_32_F32 = 37,
  FAST2_64_F32 = 38,
  FAST2_128_F32 = 39,
  FAST2_256_F32 = 40,
  FAST2_512_F32 = 41,
  FAST3R_256_F64 = 42,   // three-pass kernels instantiated for real rows only (r2c / c2r of 512, 1024, 2048 points)
  FAST3R_512_F64 = 43,
  FAST3R_1024_F64 = 44,
  FAST3R_256_F32 = 45,
  FAST3R_512_F32 = 46,
  FAST3R_1024_F32 = 47,
  FAST3_8192_F32 = 48,
  FAST2R_16_F64 = 49,    // short real rows (r2c / c2r of 32 ... 256 points), several rows per warp (fast2r_kernel)
  FAST2R_32_F64 = 50,
  FAST2R_64_F64 = 51,
  FAST2R_128_F64 = 52,
  FAST2R_16_F32 = 53,
  FAST2R_32_F32 = 54,
  FAST2R_64_F32 = 55,
  FAST2R_128_F32 = 56,
  FAST3R_500_F64 = 57,   // r2c-only shapes whose pass 3 pairs bin k with bin N-k in registers (10*10*5, 18*18*6)
  FAST3R_1944_F64 = 58,
  FAST3C_2048_F64 = 59,  // c2r-only shapes whose pass 1 pairs point n with point N-n in registers (8*16*16, 8*8*16)
  FAST3C_2048_F32 = 60,
  FAST3C_1024_F64 = 61,
  FAST3C_1024_F32 = 62,
  FAST3_500_F32 = 63,    // float32 instances of the 500 / 1000 / 1944-point shapes
  FAST3_1944_F32 = 64,
  FAST3_1000_F32 = 65,
  FAST3R_500_F32 = 66,
  FAST3R_1944_F32 = 67,
  FAST3P_512_F64 = 68,   // 8*8*8 with 16 points per thread: r2c AND c2r of 1024 points paired in registers
  FAST3P_512_F32 = 69,
  FASTBLUE_2048_F32 = 70,  // fused Bluestein in float32 (selected with IMPULSE_FFT_BLUE_F32=1 until measured)
  FASTBLUE_4096_F32 = 71,
  FASTBLUE_8192_F32 = 72,
  FAST3_1536_F64 = 73,     // more mixed-radix shapes (IMPULSE_FFT_MORE_SHAPES=0 switches them off): 8*24*8, 10*20*10, 10*20*20
  FAST3_2000_F64 = 74,
  FAST3_4000_F64 = 75,
  FAST3_2187_F64 = 76,     // 3^7 = 27*9*9, 3000 = 10*30*10, 3^8 = 27*27*9 (complex rows; round 2)
  FAST3_3000_F64 = 77,
  FAST3_6561_F64 = 78,
  FAST2_100_F64 = 79,      // short non-power-of-two rows on the two-pass warp kernel: 10*10, 27*9, 25*25
  FAST2_243_F64 = 80,
  FAST2_625_F64 = 81,
  FAST2_100_F32 = 82,
  FAST2_243_F32 = 83,
  FAST2_625_F32 = 84,
  FAST3_1536_F32 = 85,     // float32 complex rows of the round-2 shapes
  FAST3_2000_F32 = 86,
  FAST3_4000_F32 = 87,
  FAST3_2187_F32 = 88,
  FAST3_3000_F32 = 89,
  FAST3_6561_F32 = 90,
  COLCONVW_512_F32 = 91,   // whole-axis convolution along a strided axis, the axis resident in shared memory (colconvw_kernel)
  COLCONVW_1024_F32 = 92,
  COLCONVW_2048_F32 = 93,
  COLCONVW_4096_F32 = 94,
  COLCONVW_512_F64 = 95,
  COLCONVW_1024_F64 = 96,
  COLCONVW_2048_F64 = 97,
  COLCONVW_4096_F64 = 98,
  COLW_1024_F32 = 99,      // plain c2c along a strided axis, the axis resident in shared memory (colconvw_kernel, CW_FWD / CW_BWD)
  COLW_2048_F32 = 100,
  COLW_1024_F64 = 101,
  COLW_2048_F64 = 102,
  FAST2_8_F64 = 103,       // tiny rows on the two-pass warp kernels: 8 = 4*2 (16 rows per warp), 4 = 2*2
  FAST2_4_F64 = 104,
  FAST2_8_F32 = 105,
  FAST2_4_F32 = 106,
  FAST2R_8_F64 = 107,      // real rows of 16 / 8 points
  FAST2R_4_F64 = 108,
  FAST2R_8_F32 = 109,
  FAST2R_4_F32 = 110,
  // more short non-power-of-two complex rows on the two-pass warp kernel (R1 points per thread x R2 threads per row)
  FAST2_50_F64 = 111,
  FAST2_72_F64 = 112,
  FAST2_81_F64 = 113,
  FAST2_96_F64 = 114,
  FAST2_192_F64 = 115,
  FAST2_200_F64 = 116,
  FAST2_400_F64 = 117,
  FAST2_576_F64 = 118,
  FAST2_729_F64 = 119,
  FAST2_900_F64 = 120,
  FAST2_50_F32 = 121,
  FAST2_72_F32 = 122,
  FAST2_81_F32 = 123,
  FAST2_96_F32 = 124,
  FAST2_192_F32 = 125,
  FAST2_200_F32 = 126,
  FAST2_400_F32 = 127,
  FAST2_576_F32 = 128,
  FAST2_729_F32 = 129,
  FAST2_900_F32 = 130,
};

struct Phase {
  uint8_t op;
  uint8_t radix;
  uint16_t pad;
  uint32_t l1;
  uint32_t ido;
};

struct LineJob {
  // lengths
  uint32_t n_fft;    // complex FFT length run in shared memory (n2 for Bluestein)
  uint32_t n_seq;    // logical complex sequence length L (N, or N/2 for even real)
  uint32_t n_real;   // real length N (real transforms) or L
  uint32_t n_load;   // element slots iterated by LOAD per line
  uint32_t n_store;  // element slots iterated by STORE per line
  // tile geometry
  uint32_t log_c;      // lines per CTA = 1 << log_c
  uint32_t pitch;      // shared-memory elements per line
  uint32_t swz_mask;   // 0, 7 (16-byte elements) or 15 (8-byte elements)
  uint32_t flags;
  uint8_t load_mode, store_mode, dtype /*0=f32,1=f64*/, nphases;
  uint32_t fast_id;    // FastId
  uint64_t n_lines;
  // line index -> global offset (elements of the respective side's element type)
  uint64_t bdim[kMaxBatchDims];
  int64_t bs_in[kMaxBatchDims], bs_out[kMaxBatchDims];
  int64_t es_in, es_out;
  const void *in;
  void *out;
  // tables (device pointers owned by the plan cache)
  const void *tw;          // exp(-2*pi*i*m/n_fft), m < n_fft
  const uint32_t *perm;    // position of frequency k after the DIF passes; null = identity
  const void *tw_r;        // exp(-2*pi*i*k/N), k <= N/2 (even real transforms)
  const void *bk;          // exp(+i*pi*n^2/L), n < L
  const void *bkf;         // FFT_{n_fft}(wrapped bk)/n_fft at digit-reversed positions
  // four-step (long lines split N = N1*N2 over two launches): the first launch multiplies output
  // element k1 of the line with index n2 along batch dim tw4_dim by W_N^(k1*n2) = hi[m>>shift]*lo[m&mask]
  const void *tw4_hi, *tw4_lo;
  uint32_t tw4_n;          // N of the split transform; 0 = no store twiddle
  uint32_t tw4_shift;
  uint32_t tw4_dim;
  uint32_t zero_pad_from;  // ST_C: elements e >= this are stored as zero (0 = off; Bluestein staging)
  const void *f3_tw1, *f3_tw2;  // twiddle tables of the three-pass register kernels ([k1][i1], [k2][i2])
  const void *fb_bf, *fb_corr;  // fused Bluestein (fastblue_kernel): FFT(b)/M natural order, alias corrections
  uint32_t fb_d;                // its deficiency d = max(0, 2L-1-M)
  // real-to-real (DCT/DST): W_8N^m table (m < 2N+2), load factors (first / other / last input element),
  // store recipe
  const void *x_tw;
  double x_f0, x_f, x_fl, x_s, x_s0, x_sn;
  uint32_t x_shift, x_wadd;   // F index offset; twiddle index 2k + x_wadd (x_wadd = 0xffffffff: no twiddle)
  uint32_t x_im;              // 0: real part, 1: minus imaginary part
  // segmented input (LD_C): element e of a line is read from seg_base[e / seg_len] at (e % seg_len)*es_in —
  // the line is distributed over several allocations (peer GPUs' row slabs in the multi-GPU 2-D transform)
  uint32_t seg_len;            // 0 = off
  const void *seg_base[8];
  // caller-supplied multiplier fused into the store (FFT -> pointwise multiply): output element at element
  // offset o of the output array is multiplied by umul[o % umul_mod]  (umul_mod = 0: off)
  const void *umul;
  uint64_t umul_mod;
  uint32_t col_in_rows;    // column kernels: every line is a contiguous row on the input side (staged through smem)
  const void *mul_tab;     // ST_C: multiply output element e by mul_tab[line_index + mul_stride*e] (null = off)
  uint32_t mul_stride;
  double fct;
  Phase ph[kMaxPhases];
};

// Elementwise passes around a complex transform whose line does not fit in one CTA's shared memory: the
// real <-> complex conversions of long real transforms and the chirp multiplications of long Bluestein
// transforms.  One thread per element of an N-D iteration space; `work` is the complex work array, `user`
// the caller-side array (real or complex elements depending on the mode / layout).
enum AuxMode : int {
  AUX_R2C_POST_EVEN = 1,  // work Z[0..M) (M = N/2 point FFT of the packed line) -> bins 0..M in the output layout
  AUX_C2R_PRE_EVEN = 2,   // half spectrum -> Z[0..M) whose backward FFT is the packed real line
  AUX_R2C_PRE_ODD = 3,    // real line -> complex line with zero imaginary part
  AUX_R2C_POST_ODD = 4,   // bins 0..(N-1)/2 of the N-point FFT -> output layout
  AUX_C2R_PRE_ODD = 5,    // half spectrum -> Hermitian-extended N-point line
  AUX_C2R_POST_ODD = 6,   // real parts
  AUX_BLUE_PRE = 7,       // w[n] = x[n] * conj(b[n]) (n < L), 0 (L <= n < n2)      pocketfft.c:1954-1968
  AUX_BLUE_POST = 8,      // y[k] = w[k] * conj(b[k]) * fct (k < L)                 pocketfft.c:1993-2005
  // long even real lines whose real side cannot be addressed as packed complex pairs (strided axis, odd row
  // stride, negated elements): gather / scatter between the strided real line and the complex work line
  AUX_R2C_PACK_EVEN = 9,    // Z[m] = (x[2m], x[2m+1]), m < M
  AUX_C2R_UNPACK_EVEN = 10, // x[2m] = Re Z[m] * fct, x[2m+1] = Im Z[m] * fct
  // DCT / DST lines whose embedding (2N, 2N-2, 2N+2 complex points) does not fit one CTA: the LOAD / STORE modes of the
  // line kernel (LD_X_*, ST_X) as elementwise passes around a complex transform of any length
  AUX_X_EMBED = 11,         // u[e] (e < M) of the embedding selected by x_load, from the real input line
  AUX_X_EXTRACT = 12,       // real output e (e < N) from the transformed work line
};
constexpr int kMaxAuxDims = 8;
struct AuxJob {
  int mode, dtype, layout, ndim;   // layout: RealLayout of the user side (real modes)
  uint32_t flags;                  // F_CONJ_IN (c2r forward / backward Bluestein input), F_CONJ_RESULT
  uint32_t N;                      // real length (real modes) or L (Bluestein modes)
  uint32_t M;                      // length of the complex work line
  uint32_t shape[kMaxAuxDims];     // iteration space; the LAST dimension runs along the transform axis
  int64_t s_user[kMaxAuxDims];     // strides in elements of the respective side
  int64_t s_work[kMaxAuxDims];
  const void *in;
  void *out;
  const void *tab;                 // real modes: W_N^k (k <= N/2); Bluestein: b[n]
  double fct;
  uint64_t total;
  // DCT / DST passes: the same parameters as the LineJob fields of these names
  int x_load;                      // LD_X_* mode of the embedding
  const void *x_tw;
  double x_f0, x_f, x_fl, x_s, x_s0, x_sn;
  uint32_t x_shift, x_wadd, x_im;
};

// Genuine (non-separable) Hartley transform, last step (pocketfft_hdronly.h:3432-3444): the contiguous
// half spectrum F of an N-D r2c is folded into the real output, out[k] = Re F[k] + Im F[k] and
// out[-k mod shape] = Re F[k] - Im F[k].  One element of F per thread.
constexpr int kMaxCombineDims = 8;
struct CombineJob {
  const void *in;   // complex, C-contiguous, shape hshape
  void *out;        // real, strides so (elements)
  int dtype;        // 0 = f32, 1 = f64
  int ndim;
  uint32_t half_axis;                 // the axis whose extent is halved in F (axes.back())
  uint32_t hshape[kMaxCombineDims];   // iteration space
  uint32_t full[kMaxCombineDims];     // output shape
  uint8_t rev[kMaxCombineDims];       // transformed axes: the mirrored element has index (full - k) % full
  int64_t so[kMaxCombineDims];
  uint64_t total;
};

// shared-memory position of logical index a within a line
IMP_HD uint32_t swz(uint32_t a, uint32_t mask) {
  return a ^ (((a >> 3) ^ (a >> 6) ^ (a >> 9) ^ (a >> 12)) & mask);
}
IMP_HD uint32_t swz16(uint32_t a, uint32_t mask) {
  return a ^ (((a >> 4) ^ (a >> 8) ^ (a >> 12)) & mask);
}

}  // namespace impulse
