// Complex arithmetic on the packed FP32 instructions of sm_100a (FADD2 / FMUL2 / FFMA2, PTX add / mul / fma .f32x2).
//
// A complex number is ONE 64-bit register pair (re = low half, im = high half), so a complex add is one instruction
// instead of two and a complex multiply two instead of four: ptxas folds the re<->im swap, a per-half sign and a
// scalar broadcast into operand modifiers (`R4.F32x2.LO_HI.NP`, `UR6.F32`), which is what the helpers below rely on
// (checked with cuobjdump: no MOV / FMUL(-1) is emitted for cpk_swap / cpk_mul_mi / the constant pairs).
// Same butterflies as gr_fft2048.cuh (dft16 = 4 x radix-4, twiddles, 4 x radix-4; dft8 = 2 x radix-4, twiddles,
// 4 x radix-2); results differ from the scalar forms only where multiply + add became one fused multiply-add.
#pragma once
#include "gr_fft2048.cuh"

#if defined(__CUDACC__)
typedef unsigned long long cpk;

__device__ __forceinline__ cpk cpk_make(float re, float im) {
    cpk r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(re), "f"(im));
    return r;
}
__device__ __forceinline__ void cpk_split(cpk a, float& re, float& im) { asm("mov.b64 {%0, %1}, %2;" : "=f"(re), "=f"(im) : "l"(a)); }
__device__ __forceinline__ float cpk_re(cpk a) { float x, y; cpk_split(a, x, y); return x; }
__device__ __forceinline__ float cpk_im(cpk a) { float x, y; cpk_split(a, x, y); return y; }
__device__ __forceinline__ cpk cpk_add(cpk a, cpk b) { cpk r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ cpk cpk_sub(cpk a, cpk b) { cpk r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ cpk cpk_mul(cpk a, cpk b) { cpk r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ cpk cpk_fma(cpk a, cpk b, cpk c) { cpk r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ cpk cpk_swap(cpk a) { float x, y; cpk_split(a, x, y); return cpk_make(y, x); }
// a * (-i) = (im, -re)
__device__ __forceinline__ cpk cpk_mul_mi(cpk a) { float x, y; cpk_split(a, x, y); return cpk_make(y, -x); }
__device__ __forceinline__ cpk cpk_scale(cpk a, float k) { return cpk_mul(a, cpk_make(k, k)); }
// a * (wr + i wi) = a * (wr, wr) + swap(a) * (-wi, wi)
__device__ __forceinline__ cpk cpk_cmul(cpk a, float wr, float wi) { return cpk_fma(cpk_swap(a), cpk_make(-wi, wi), cpk_mul(a, cpk_make(wr, wr))); }
// a + b * (-i) and a - b * (-i): (a.re + b.im, a.im - b.re), (a.re - b.im, a.im + b.re)
__device__ __forceinline__ cpk cpk_add_mi(cpk a, cpk b) { return cpk_add(a, cpk_mul_mi(b)); }
__device__ __forceinline__ cpk cpk_sub_mi(cpk a, cpk b) { return cpk_sub(a, cpk_mul_mi(b)); }

__device__ __forceinline__ cpk cpk_neg(cpk a) { float x, y; cpk_split(a, x, y); return cpk_make(-x, -y); }
// acc + a * (wr + i wi): two FFMA2
__device__ __forceinline__ cpk cpk_mac(cpk acc, cpk a, float wr, float wi) {
    return cpk_fma(cpk_swap(a), cpk_make(-wi, wi), cpk_fma(a, cpk_make(wr, wr), acc));
}
// 2 a - b: one FFMA2.  With s = p + q already formed, p - q = 2 p - s costs one instruction instead of the two of a
// second multiply-accumulate chain (q is a product here).
__device__ __forceinline__ cpk cpk_2amb(cpk a, cpk b) { return cpk_fma(a, cpk_make(2.f, 2.f), cpk_neg(b)); }

__device__ __forceinline__ void cpk_bf4(cpk& a0, cpk& a1, cpk& a2, cpk& a3) {
    const cpk s02 = cpk_add(a0, a2), d02 = cpk_sub(a0, a2);
    const cpk s13 = cpk_add(a1, a3), t = cpk_sub(a1, a3);
    a0 = cpk_add(s02, s13);
    a2 = cpk_sub(s02, s13);
    a1 = cpk_add_mi(d02, t);
    a3 = cpk_sub_mi(d02, t);
}
// Radix-4 butterfly of (a0, w1 a1, w2 a2, w3 a3): the twiddle products are folded into the first layer of
// additions (14 packed instructions instead of 6 + 8).
__device__ __forceinline__ void cpk_bf4_w123(cpk& a0, cpk& a1, cpk& a2, cpk& a3, float w1r, float w1i, float w2r, float w2i,
                                             float w3r, float w3i) {
    const cpk s02 = cpk_mac(a0, a2, w2r, w2i), d02 = cpk_2amb(a0, s02);
    const cpk p1 = cpk_cmul(a1, w1r, w1i);
    const cpk s13 = cpk_mac(p1, a3, w3r, w3i), t = cpk_2amb(p1, s13);
    a0 = cpk_add(s02, s13);
    a2 = cpk_sub(s02, s13);
    a1 = cpk_add_mi(d02, t);
    a3 = cpk_sub_mi(d02, t);
}
// ... of (w0 a0, w1 a1, w2 a2, w3 a3); w = {w0r, w0i, w1r, w1i, w2r, w2i, w3r, w3i}
__device__ __forceinline__ void cpk_bf4_w0123(cpk& a0, cpk& a1, cpk& a2, cpk& a3, const float* w) {
    a0 = cpk_cmul(a0, w[0], w[1]);
    cpk_bf4_w123(a0, a1, a2, a3, w[2], w[3], w[4], w[5], w[6], w[7]);
}
// ... of (a0, w a1, -i a2, w' a3): the W16 row (1, W2, W4, W6)
__device__ __forceinline__ void cpk_bf4_w1_mi_w3(cpk& a0, cpk& a1, cpk& a2, cpk& a3, float w1r, float w1i, float w3r, float w3i) {
    const cpk s02 = cpk_add_mi(a0, a2), d02 = cpk_sub_mi(a0, a2);
    const cpk p1 = cpk_cmul(a1, w1r, w1i);
    const cpk s13 = cpk_mac(p1, a3, w3r, w3i), t = cpk_2amb(p1, s13);
    a0 = cpk_add(s02, s13);
    a2 = cpk_sub(s02, s13);
    a1 = cpk_add_mi(d02, t);
    a3 = cpk_sub_mi(d02, t);
}

// second half of the radix-16 butterfly: the W16^(j0 k1) twiddles folded into the four radix-4 over j0, then the
// 4 x 4 transposition that leaves a[k] at index k
__device__ __forceinline__ void cpk_dft16_out(cpk* a) {
    cpk_bf4(a[0], a[1], a[2], a[3]);
    cpk_bf4_w123(a[4], a[5], a[6], a[7], GR_C1, -GR_S1, GR_R2, -GR_R2, GR_S1, -GR_C1);
    cpk_bf4_w1_mi_w3(a[8], a[9], a[10], a[11], GR_R2, -GR_R2, -GR_R2, -GR_R2);
    cpk_bf4_w123(a[12], a[13], a[14], a[15], GR_S1, -GR_C1, -GR_R2, -GR_R2, -GR_C1, GR_S1);
    cpk t;
    t = a[1];  a[1]  = a[4];  a[4]  = t;
    t = a[2];  a[2]  = a[8];  a[8]  = t;
    t = a[3];  a[3]  = a[12]; a[12] = t;
    t = a[6];  a[6]  = a[9];  a[9]  = t;
    t = a[7];  a[7]  = a[13]; a[13] = t;
    t = a[11]; a[11] = a[14]; a[14] = t;
}
// first half for j0 = J0, J0 + 1 (inputs j0 + 4 m), every input multiplied by its own twiddle on the way in:
// w[8 (j0 - J0) + 2 m], w[.. + 1] = twiddle of input j0 + 4 m
template <int J0>
__device__ __forceinline__ void cpk_dft16_in_tw(cpk* a, const float* w) {
    cpk_bf4_w0123(a[J0], a[J0 + 4], a[J0 + 8], a[J0 + 12], w);
    cpk_bf4_w0123(a[J0 + 1], a[J0 + 5], a[J0 + 9], a[J0 + 13], w + 8);
}

// forward DFT-16 in place: a[k] <- sum_j a[j] W16^(jk)
__device__ __forceinline__ void cpk_dft16(cpk* a) {
    cpk_bf4(a[0], a[4], a[8], a[12]);
    cpk_bf4(a[1], a[5], a[9], a[13]);
    cpk_bf4(a[2], a[6], a[10], a[14]);
    cpk_bf4(a[3], a[7], a[11], a[15]);
    cpk_dft16_out(a);
}

// second half of the radix-8 butterfly (W8 twiddles folded into the radix-2 layer)
__device__ __forceinline__ void cpk_dft8_out(cpk* a) {
    cpk r[8];
    r[0] = cpk_add(a[0], a[1]);
    r[4] = cpk_sub(a[0], a[1]);
    r[1] = cpk_mac(a[2], a[3], GR_R2, -GR_R2);
    r[5] = cpk_2amb(a[2], r[1]);
    r[2] = cpk_add_mi(a[4], a[5]);
    r[6] = cpk_sub_mi(a[4], a[5]);
    r[3] = cpk_mac(a[6], a[7], -GR_R2, -GR_R2);
    r[7] = cpk_2amb(a[6], r[3]);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = r[k];
}
// forward DFT-8 in place
__device__ __forceinline__ void cpk_dft8(cpk* a) {
    cpk_bf4(a[0], a[2], a[4], a[6]);
    cpk_bf4(a[1], a[3], a[5], a[7]);
    cpk_dft8_out(a);
}
// forward DFT-8 of (a0, w1 a1, ..., w7 a7); w[2 (n - 1)], w[2 (n - 1) + 1] = twiddle of input n = 1..7
__device__ __forceinline__ void cpk_dft8_tw(cpk* a, const float* w) {
    cpk_bf4_w123(a[0], a[2], a[4], a[6], w[2], w[3], w[6], w[7], w[10], w[11]);
    a[1] = cpk_cmul(a[1], w[0], w[1]);
    cpk_bf4_w123(a[1], a[3], a[5], a[7], w[4], w[5], w[8], w[9], w[12], w[13]);
    cpk_dft8_out(a);
}
// ---- the reference's NCO for TWO consecutive samples on the packed instructions -------------------------------------
// fn = (n + 1, n + 2) as floats.  Lane by lane the same IEEE operations in the same order as nco_exact (gr_common.cuh):
// bit-identical factors, half the issue slots for the nine argument instructions (the two MUFU per sample stay scalar).
__device__ __forceinline__ cpk cpk_bc(float v) { return cpk_make(v, v); }
// exp(-i arg) for the two halves of `arg` (nco_fast2 twice)
__device__ __forceinline__ void nco_fast2_pair(cpk arg, cf& e0, cf& e1) {
    const cpk k = cpk_add(cpk_fma(arg, cpk_bc(0.15915494309189535f), cpk_bc(12582912.0f)), cpk_bc(-12582912.0f));
    cpk r = cpk_fma(k, cpk_bc(-6.28125f), arg);
    r = cpk_fma(k, cpk_bc(-1.9353071795864769e-3f), r);
    float r0, r1;
    cpk_split(r, r0, r1);
    e0 = cf{__cosf(r0), -__sinf(r0)};
    e1 = cf{__cosf(r1), -__sinf(r1)};
}
// (tsec_of(k0), tsec_of(k1)) for fn = (k0, k1)
__device__ __forceinline__ cpk tsec_of2(cpk fn) {
    return cpk_fma(fn, cpk_bc(__uint_as_float(889393775u)), cpk_mul(fn, cpk_bc(__uint_as_float(2832262496u))));
}
// fl32(phase + fl32(w t)) needs the product ROUNDED before the addition.  ptxas contracts `mul.rn.f32x2` + `add.rn.f32x2`
// into one FFMA2 (seen in the SASS; it even rewrites fma(p, 1, phase) into that), which silently drops the rounding --
// the scalar `add.rn.f32` is never contracted, so the two additions are scalar.
__device__ __forceinline__ void nco_exact2(float w32, float phase32, cpk fn, cf& e0, cf& e1) {
    float p0, p1;
    cpk_split(cpk_mul(cpk_bc(w32), tsec_of2(fn)), p0, p1);
    nco_fast2_pair(cpk_make(__fadd_rn(phase32, p0), __fadd_rn(phase32, p1)), e0, e1);
}
#endif
