// CTA-level complex FFT-2048, second generation ("w" = warp-local second exchange).
//
// Same factorisation and butterflies as gr_fft2048.cuh (2048 = 16 x 16 x 8, decimation in
// frequency, FP32, 128 threads x 16 points), so every output is bit-identical to it; what
// changes is how the two transposes use shared memory, because ncu shows the inverse-FFT
// acquisition kernel bound by the shared-memory data pipe and by barrier stalls
// (profiles/acq_r01_v2_ncu_summary.md):
//
//   exchange 1 (all 128 threads): buf1[k1>>1][t][k1&1] -- the writer stores PAIRS of outputs as
//     one 128-bit word (8 STS.128 instead of 16 STS.64), the reader's 64-bit loads cover 128
//     contiguous bytes per half-warp.  No padding, no bank conflicts.  The buffer is double
//     buffered by the caller, so ONE block barrier per transform is enough.
//   exchange 2 (the 8 lanes that share k1): stays inside a warp.  Lane (k1loc = l>>3, n3 = l&7)
//     writes its 16 stage-2 outputs as 8 x 128-bit; lane (k1loc = l&3, k2 = (l>>2) + 8h) reads the
//     eight n3 values of its two radix-8 groups.  Additive, conflict-free layout
//     E = 148*k1loc + 18*(k2>>1) + 2*n3 + (k2&1)  (8-byte units); only __syncwarp() around it.
//
// Register layout:  input  v[j] = x[t + 128 j]
//                   output v[j] = X[fftw_out_base(t) + 128 j]   (a fixed permutation of the
//                   128 residues, see fftw_out_base) -- callers that only reduce over all lags
//                   (acquisition) just use that index for the arg-max.
#pragma once
#include "gr_fft2048.cuh"

#define GR_W_BUF1_BYTES (8 * 128 * 16)                 // one exchange-1 buffer: 16 KiB
#define GR_W_A 148                                     // exchange-2 strides, 8-byte units
#define GR_W_B 18
#define GR_W_WARP_UNITS (4 * GR_W_A)                   // 592 x 8 B per warp
#define GR_W_BUF2_BYTES (4 * GR_W_WARP_UNITS * 8)      // 18 944 B for the 4 warps
#define GR_W_SMEM_BYTES (2 * GR_W_BUF1_BYTES + GR_W_BUF2_BYTES)   // 51 712 B per CTA

// residue (mod 128) of the output indices held by thread t
GR_HD int fftw_out_base(int t) {
    const int w = t >> 5, l = t & 31;
    return 4 * w + (l & 3) + 16 * (l >> 2);
}

#if defined(__CUDACC__)
// stage-1 twiddle, exchange 1 write (caller applies the twiddles before) ... split into pieces so
// that callers can interleave their own loads (TMEM-resident twiddles).
__device__ __forceinline__ void fftw_ex1_write(float4* buf1, int t, const cf* v) {
#pragma unroll
    for (int m = 0; m < 8; ++m) buf1[m * 128 + t] = make_float4(v[2 * m].x, v[2 * m].y, v[2 * m + 1].x, v[2 * m + 1].y);
}
__device__ __forceinline__ void fftw_ex1_read(const float4* buf1, int t, cf* v) {
    const int k1 = t >> 3, n3 = t & 7;
    const cf* b = reinterpret_cast<const cf*>(buf1) + 2 * ((k1 >> 1) * 128 + n3) + (k1 & 1);
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = b[16 * n2];
}
// buf2w: this warp's exchange-2 region (GR_W_WARP_UNITS x 8 B, 16-byte aligned)
__device__ __forceinline__ void fftw_ex2_write(cf* buf2w, int l, const cf* v) {
    float4* p = reinterpret_cast<float4*>(buf2w + GR_W_A * (l >> 3) + 2 * (l & 7));
#pragma unroll
    for (int m = 0; m < 8; ++m) p[(GR_W_B / 2) * m] = make_float4(v[2 * m].x, v[2 * m].y, v[2 * m + 1].x, v[2 * m + 1].y);
}
__device__ __forceinline__ void fftw_ex2_read_stage3(const cf* buf2w, int l, cf* v) {
    // groups (k1loc = l & 3, k2 = (l >> 2) + 8 h): m = k2 >> 1 = (l >> 3) + 4 h, parity = (l >> 2) & 1
    const cf* p = buf2w + GR_W_A * (l & 3) + GR_W_B * (l >> 3) + ((l >> 2) & 1);
    cf a[8], b[8];
#pragma unroll
    for (int n3 = 0; n3 < 8; ++n3) {
        a[n3] = p[2 * n3];
        b[n3] = p[2 * n3 + 4 * GR_W_B];
    }
    dft8(a);
    dft8(b);
#pragma unroll
    for (int k3 = 0; k3 < 8; ++k3) {
        v[2 * k3] = a[k3];
        v[2 * k3 + 1] = b[k3];
    }
}
#endif
