// CTA-level complex FFT of length 2048 for 128 threads x 16 points, FP32.
//
// Factorisation 2048 = 16 x 16 x 8 (decimation in frequency), two shared-memory
// transposes.  Register layout is the same on input and output:
//     thread t holds element  t + 128*j  in v[j],  j = 0..15
// so that global loads/stores are coalesced and the transform can be chained
// (forward -> spectrum multiply -> inverse) without any reordering pass.
//
//   n = n1*128 + n2*8 + n3          k = k1 + 16*k2 + 256*k3
//   stage 1: thread (n2,n3): radix-16 over n1, twiddle W_2048^((n2*8+n3)*k1)
//   stage 2: thread (k1,n3): radix-16 over n2, twiddle W_128^(n3*k2)
//   stage 3: thread g=(k2,k1), two groups per thread: radix-8 over n3
//
// Only the forward transform (e^{-i 2 pi nk/N}) is implemented.  The inverse is
// obtained with the swap identity  ifft(y) = swap(fft(swap(y)))  (swap = exchange
// real and imaginary parts), which costs nothing when the caller produces the
// swapped operand directly and only needs |.|^2 of the result.  One code path
// keeps the unrolled butterflies inside the instruction cache.
//
// Shared memory: two padded buffers so that one __syncthreads per transpose is
// enough ("FFT-2048"):  buf1[16][136] and buf2[8][258] cf.
// Both paddings make every 64-bit access conflict-free per half-warp.
#pragma once
#include "gr_common.cuh"

#define GR_B1_STRIDE 136
#define GR_B2_STRIDE 258
#define GR_B1_ELEMS (16 * GR_B1_STRIDE)                 // 2176
#define GR_B2_ELEMS (8 * GR_B2_STRIDE)                  // 2064
#define GR_FFT_SMEM_BYTES ((GR_B1_ELEMS + GR_B2_ELEMS) * 8)   // 33920

#define GR_C1 0.92387953251128674f   // cos(pi/8)
#define GR_S1 0.38268343236508977f   // sin(pi/8)
#define GR_R2 0.70710678118654752f   // 1/sqrt(2)

// y = x * (-i)
GR_HD cf mul_mi(cf a) { return cf{a.y, -a.x}; }

// forward radix-4, in place: (a0,a1,a2,a3) -> DFT4
GR_HD void bf4(cf& a0, cf& a1, cf& a2, cf& a3) {
    cf s02 = cadd(a0, a2), d02 = csub(a0, a2);
    cf s13 = cadd(a1, a3), d13 = mul_mi(csub(a1, a3));   // -i (a1 - a3)
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = cadd(d02, d13);
    a3 = csub(d02, d13);
}

// forward DFT-16 in place: a[k] <- sum_j a[j] W16^(jk)
GR_HD void dft16(cf* a) {
    // step 1: four radix-4 over j1 for each j0 (elements j0, j0+4, j0+8, j0+12)
    bf4(a[0], a[4], a[8], a[12]);
    bf4(a[1], a[5], a[9], a[13]);
    bf4(a[2], a[6], a[10], a[14]);
    bf4(a[3], a[7], a[11], a[15]);
    // now a[j0 + 4*k1] = T[j0][k1]; step 2: multiply by W16^(j0*k1)
    // j0 = 1: k1 = 1,2,3 -> W1, W2, W3
    a[5]  = cmul(a[5],  cf{GR_C1, -GR_S1});
    a[9]  = cf{(a[9].x + a[9].y) * GR_R2, (a[9].y - a[9].x) * GR_R2};          // W2 = (1-i)/sqrt2
    a[13] = cmul(a[13], cf{GR_S1, -GR_C1});
    // j0 = 2: W2, W4, W6
    a[6]  = cf{(a[6].x + a[6].y) * GR_R2, (a[6].y - a[6].x) * GR_R2};
    a[10] = mul_mi(a[10]);                                                      // W4 = -i
    a[14] = cf{(a[14].y - a[14].x) * GR_R2, -(a[14].x + a[14].y) * GR_R2};      // W6 = (-1-i)/sqrt2
    // j0 = 3: W3, W6, W9
    a[7]  = cmul(a[7],  cf{GR_S1, -GR_C1});
    a[11] = cf{(a[11].y - a[11].x) * GR_R2, -(a[11].x + a[11].y) * GR_R2};
    a[15] = cmul(a[15], cf{-GR_C1, GR_S1});
    // step 3: radix-4 over j0 for each k1 (elements 4*k1 + j0); result k = k1 + 4*k0 sits at 4*k1 + k0
    bf4(a[0], a[1], a[2], a[3]);
    bf4(a[4], a[5], a[6], a[7]);
    bf4(a[8], a[9], a[10], a[11]);
    bf4(a[12], a[13], a[14], a[15]);
    // transpose 4x4 so that a[k] is at index k: position 4*k1 + k0 -> k1 + 4*k0
    cf t;
    t = a[1];  a[1]  = a[4];  a[4]  = t;
    t = a[2];  a[2]  = a[8];  a[8]  = t;
    t = a[3];  a[3]  = a[12]; a[12] = t;
    t = a[6];  a[6]  = a[9];  a[9]  = t;
    t = a[7];  a[7]  = a[13]; a[13] = t;
    t = a[11]; a[11] = a[14]; a[14] = t;
}

// forward DFT-8 in place on a[0..7] (stride 1): a[k] <- sum_j a[j] W8^(jk)
GR_HD void dft8(cf* a) {
    // j = j0 + 2*j1 (j0 in 0..1, j1 in 0..3); k = k1 + 4*k0
    bf4(a[0], a[2], a[4], a[6]);        // j0 = 0 -> T[0][k1] at a[2*k1]
    bf4(a[1], a[3], a[5], a[7]);        // j0 = 1 -> T[1][k1] at a[2*k1+1]
    // twiddle W8^(j0*k1), j0 = 1
    a[3] = cf{(a[3].x + a[3].y) * GR_R2, (a[3].y - a[3].x) * GR_R2};            // W8^1
    a[5] = mul_mi(a[5]);                                                         // W8^2
    a[7] = cf{(a[7].y - a[7].x) * GR_R2, -(a[7].x + a[7].y) * GR_R2};           // W8^3
    // radix-2 over j0: A[k1] = T0 + T1, A[k1+4] = T0 - T1
    cf r[8];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        r[k1]     = cadd(a[2 * k1], a[2 * k1 + 1]);
        r[k1 + 4] = csub(a[2 * k1], a[2 * k1 + 1]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = r[k];
}

// ---- the three stages, split so that a host emulation can run them per "thread" ----

// tw1: this thread's 16 stage-1 twiddles W_2048^(t*k1) (forward); tw2: W_128^(n3*k2)
template <int TW1_STRIDE = 1>
GR_HD void fft_stage1(cf* v, const cf* tw1) {
    dft16(v);
#pragma unroll
    for (int k = 1; k < 16; ++k) v[k] = cmul(v[k], tw1[k * TW1_STRIDE]);
}
GR_HD void fft_ex1_write(cf* buf1, int t, const cf* v) {
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) buf1[k1 * GR_B1_STRIDE + t] = v[k1];
}
GR_HD void fft_ex1_read(const cf* buf1, int t, cf* v) {
    const int k1 = t >> 3, n3 = t & 7;
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = buf1[k1 * GR_B1_STRIDE + n2 * 8 + n3];
}
GR_HD void fft_stage2(cf* v, const cf* tw2) {
    dft16(v);
#pragma unroll
    for (int k = 1; k < 16; ++k) v[k] = cmul(v[k], tw2[k]);
}
GR_HD void fft_ex2_write(cf* buf2, int t, const cf* v) {
    const int k1 = t >> 3, n3 = t & 7;
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) buf2[n3 * GR_B2_STRIDE + k2 * 16 + k1] = v[k2];
}
// after this v[2*n3 + h] holds B'[g = t + 128 h][n3]; stage 3 leaves X[t + 128*(2*k3+h)] in v[2*k3+h]
GR_HD void fft_ex2_read_stage3(const cf* buf2, int t, cf* v) {
    cf a[8], b[8];
#pragma unroll
    for (int n3 = 0; n3 < 8; ++n3) {
        a[n3] = buf2[n3 * GR_B2_STRIDE + t];
        b[n3] = buf2[n3 * GR_B2_STRIDE + t + 128];
    }
    dft8(a);
    dft8(b);
#pragma unroll
    for (int k3 = 0; k3 < 8; ++k3) {
        v[2 * k3]     = a[k3];
        v[2 * k3 + 1] = b[k3];
    }
}

#if defined(__CUDACC__)
// Forward FFT-2048 across the 128 threads of a CTA (or of a 128-thread group
// using its own buffers and `bar_id` as named barrier).  smem = buf1 | buf2.
// TW1_STRIDE > 1: stage-1 twiddles read from a shared-memory table laid out [k][thread].
// kOneBuf: both transposes go through ONE buffer of GR_ONEBUF_BYTES (two more barriers per transform: before the first
// store, because the previous transform's last loads read the same memory, and between the first transpose's loads and
// the second one's stores).  Same arithmetic, same results; for callers that are short of shared memory.
#define GR_ONEBUF_BYTES (GR_B1_ELEMS * 8)              // 17408 >= GR_B2_ELEMS * 8
template <bool kWholeCta, int TW1_STRIDE = 1, bool kOneBuf = false>
__device__ __forceinline__ void fft2048(cf* v, cf* smem, const cf* tw1, const cf* tw2, int t, int bar_id = 1) {
    cf* buf1 = smem;
    cf* buf2 = kOneBuf ? smem : smem + GR_B1_ELEMS;
    auto sync = [&]() { if (kWholeCta) __syncthreads(); else asm volatile("bar.sync %0, 128;" ::"r"(bar_id)); };
    fft_stage1<TW1_STRIDE>(v, tw1);
    if (kOneBuf) sync();
    fft_ex1_write(buf1, t, v);
    sync();
    fft_ex1_read(buf1, t, v);
    fft_stage2(v, tw2);
    if (kOneBuf) sync();
    fft_ex2_write(buf2, t, v);
    sync();
    fft_ex2_read_stage3(buf2, t, v);
}
#endif
