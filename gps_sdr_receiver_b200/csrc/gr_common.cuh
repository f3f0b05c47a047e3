// Shared definitions for the B200 GPS hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GR_N 2048            // samples per 1-ms C/A code period (gpsglob.py:119-121)
#define GR_FS 2048000.0f     // sample rate, float32 like the reference time base
#define GR_MAX_PRN 37
#define GR_FFT_THREADS 128   // one FFT-2048 = 128 threads x 16 points

#if defined(__CUDACC__)
#define GR_HD __host__ __device__ __forceinline__
#else
#define GR_HD inline
#endif

struct cf { float x, y; };   // complex float (kept POD so it lives in registers)

GR_HD cf cmul(cf a, cf b) { return cf{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
GR_HD cf cmul_conj(cf a, cf b) { return cf{a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y}; }  // a * conj(b)
GR_HD cf cadd(cf a, cf b) { return cf{a.x + b.x, a.y + b.y}; }
GR_HD cf csub(cf a, cf b) { return cf{a.x - b.x, a.y - b.y}; }

// Device-resident constant tables (built once by gr_init, gr_tables.cu).
struct GrTables {
    const float*  code;      // [GR_MAX_PRN+1][2048] resampled C/A code, float32 (exact, see gr_tables.cu)
    const float2* conjspec;  // [GR_MAX_PRN+1][2048] conj(fft(code)), complex64, natural order
    const float2* tw1;       // [128][16]  W_2048^(t*k)    forward (e^{-i...})
    const float2* tw2;       // [8][16]    W_128^(n3*k)    forward
    const int8_t* chips;     // [GR_MAX_PRN+1][1024] the 1023 Gold-code chips (+1/-1), padded
};

#if defined(__CUDACC__)
// ---- the reference's NCO, sample by sample ------------------------------------------------------------------------
// The reference rotates sample n by exp(-i fl32(phase + fl32(w32 * t_n))), t_n = fl32((n + 1) / fs) (gpsrecv.py:32-33,
// 232-235; gpslib.py:1053-1054, 1343-1346).  At |w t| ~ 1e3 .. 1e4 rad the float32 rounding of that argument (6e-5 .. 1e-3
// rad) is the largest term in the distance between the reference and the mathematically exact rotation, so the forms
// of the kernels that have to stay within 1e-4 of the reference evaluate the same argument per sample.
// fl32(k / fs) for an integer-valued float k <= 2^23 in two instructions: 1 / fs as a two-term float32 sum y_hi + y_lo
// (accurate to 2^-49), t = fma(k, y_hi, fl32(k * y_lo)).  Equal to the correctly rounded IEEE division for every
// k = 1 .. 2^23 (checked exhaustively on the host, tests/test_capi_host.py); the generic division is ~10 instructions
// with a slow-path branch, and the reference-exact forms of the kernels run one per sample.
__device__ __forceinline__ float tsec_of(float k) {
    return __fmaf_rn(k, __uint_as_float(889393775u) /* fl32(1 / fs) = 4.882812731921149e-07 */,
                     __fmul_rn(k, __uint_as_float(2832262496u) /* fl32(1 / fs - y_hi) = -2.31921144615288e-14 */));
}
// exp(-i arg), |arg| < 2^22 turns: 2-constant Cody-Waite reduction (6.28125 = 201/32: k * C1 is exact) + MUFU sin / cos,
// |err| < 5e-7; the rounding to a whole number of turns by the 1.5 * 2^23 trick (two full-rate additions, no FRND)
__device__ __forceinline__ cf nco_fast2(float arg) {
    const float k = __fadd_rn(fmaf(arg, 0.15915494309189535f, 12582912.0f), -12582912.0f);
    float r = fmaf(k, -6.28125f, arg);
    r = fmaf(k, -1.9353071795864769e-3f, r);
    return cf{__cosf(r), -__sinf(r)};
}
// the reference's factor for sample n of a wipe-off with carried phase `phase32`; fn1 = (float)(n + 1)
__device__ __forceinline__ cf nco_exact(float w32, float phase32, float fn1) {
    return nco_fast2(__fadd_rn(phase32, __fmul_rn(w32, tsec_of(fn1))));
}
#endif
