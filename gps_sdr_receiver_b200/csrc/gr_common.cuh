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
