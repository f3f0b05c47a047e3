// CTA-level complex FFT-2048, third generation ("t" = second exchange through tensor memory).
//
// Same factorisation and butterflies as gr_fft2048.cuh / gr_fft2048w.cuh (bit-identical results).
// The first transpose (all 128 threads) stays in shared memory (gr_fft2048w.cuh exchange 1: 128-bit
// stores, double buffered, one block barrier).  The second transpose only involves the 8 lanes of a
// warp that share k1, and tools/ubench shows the shared-memory data pipe (128 B/clk/SM, shared with
// warp shuffles and with L1 fills) is what bounds the kernel, while TMEM moves ~850 B/clk/SM.  TMEM
// is lane-private for the 32x32b shape, but the 16x256b shape scatters a thread's registers over
// other lanes (the MMA accumulator fragment layout, measured by tools/ubench/tmem_probe.cu):
//
//     16x256b.x4 store, thread T, register 16 b + 4 i + 2 r + q  ->  lane 16 b + 8 r + T/4, column 8 i + 2 (T%4) + q
//     32x32b.x32 load,  thread L, register c                     <-  lane L, column c
//
// One store/load round therefore turns the lane bits (T1, T0) into register bits and the register
// bits (b, r) into lane bits.  Two rounds (with the SAME register order, so ptxas needs no moves)
// take  lane (T4 T3 T2 T1 T0), complex register (K3 K2 K1 K0)  to
//       lane (K2 T0 K3 K0 T4), complex register (K1 T1 T3 T2).
// Stage 2 runs with  n3 = (T >> 1) & 7,  k1loc = 2 (T >> 4) + (T & 1)  and stores its output k2 =
// (e3 e2 e1 e0) in register K = (e2 e1 e3 e0); after the two rounds lane L'' = (e1 T0 e2 e0 T4) holds,
// for h = e3 in {0, 1}, the eight n3 inputs of the radix-8 group (k1loc, k2) in registers
//       8 h + 4 (n3 & 1) + 2 (n3 >> 2) + ((n3 >> 1) & 1).
// Output registers: v[j] = X[ fftt_out_base(t) + 128 j ].
#pragma once
#include "gr_fft2048w.cuh"
#include "gr_cpk.cuh"

// residue (mod 128) of the output indices held by thread t:  k1 + 16 (k2 & 7)
GR_HD int fftt_out_base(int t) {
    const int w = t >> 5, L = t & 31;
    const int k1loc = 2 * (L & 1) + ((L >> 3) & 1);
    const int k2lo = 4 * ((L >> 2) & 1) + 2 * ((L >> 4) & 1) + ((L >> 1) & 1);
    return 4 * w + k1loc + 16 * k2lo;
}

#if defined(__CUDACC__)
// exchange-1 read for the stage-2 thread mapping of this file
__device__ __forceinline__ void fftt_ex1_read(const float4* buf1, int t, cf* v) {
    // k1 = 4 w + 2 T4 + T0, n3 = (T >> 1) & 7: unit = (2 w + T4) * 256 + 16 n2 + (T & 15)
    const int w = t >> 5, L = t & 31;
    const cf* b = reinterpret_cast<const cf*>(buf1) + (2 * w + (L >> 4)) * 256 + (L & 15);
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = b[16 * n2];
}
// stage-2 twiddle index of this thread: W_128^(n3 k2)
__device__ __forceinline__ int fftt_n3(int t) { return (t >> 1) & 7; }

__device__ __forceinline__ void tm_st_16x256b_x4(uint32_t taddr, const float* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
        :: "r"(taddr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),
           "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]));
}
__device__ __forceinline__ void tm_ld_32x32b_x32(uint32_t taddr, float* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]),
          "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]), "=f"(r[17]), "=f"(r[18]),
          "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]),
          "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
        : "r"(taddr));
}

// One store/load round on this warp's 32 lanes x 32 columns at `taddr` (lane field = warp's base lane).
__device__ __forceinline__ void tm_ld_32x32b_x16x2(uint32_t taddr, float* r) {      // two 16-register blocks, one wait
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];\n"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%33];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]),
          "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]), "=f"(r[17]), "=f"(r[18]),
          "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]),
          "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
        : "r"(taddr), "r"(taddr + 16));
}
__device__ __forceinline__ void tm_round(uint32_t taddr, float* r) {
    tm_st_16x256b_x4(taddr, r);
    tm_st_16x256b_x4(taddr + (16u << 16), r + 16);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#ifdef GR_TMEM_X32      // build switch of an experiment (profiles/acq_r02_x32_ab.log): one 32-register fetch instead of two 16-register ones is
                        // 0.7 % faster in the 4-CTA form and 7 % slower in the quad form (32 consecutive registers: spills at its 128-register cap)
    tm_ld_32x32b_x32(taddr, r);
#else
    tm_ld_32x32b_x16x2(taddr, r);
#endif
}

// Exchange 2 + stage 3: v[k2] (stage-2 outputs, twiddled) -> v[2 k3 + h] = X[base + 128 (2 k3 + h)]
__device__ __forceinline__ void fftt_ex2_stage3(uint32_t taddr, cf* v) {
    float r[32];
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
        const int K = 8 * ((k2 >> 2) & 1) + 4 * ((k2 >> 1) & 1) + 2 * ((k2 >> 3) & 1) + (k2 & 1);   // (e2 e1 e3 e0)
        r[2 * K] = v[k2].x;
        r[2 * K + 1] = v[k2].y;
    }
    tm_round(taddr, r);
    tm_round(taddr, r);
    cf a[8], b[8];
#pragma unroll
    for (int n3 = 0; n3 < 8; ++n3) {
        const int K = 4 * (n3 & 1) + 2 * (n3 >> 2) + ((n3 >> 1) & 1);
        a[n3] = cf{r[2 * K], r[2 * K + 1]};
        b[n3] = cf{r[2 * (8 + K)], r[2 * (8 + K) + 1]};
    }
    dft8(a);
    dft8(b);
#pragma unroll
    for (int k3 = 0; k3 < 8; ++k3) {
        v[2 * k3] = a[k3];
        v[2 * k3 + 1] = b[k3];
    }
}

// ---- the same transform on packed complex registers (gr_cpk.cuh: FADD2 / FMUL2 / FFMA2) ----
__device__ __forceinline__ void fftt_ex1_write_pk(float4* buf1, int t, const cpk* v) {
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        float4 q;
        cpk_split(v[2 * m], q.x, q.y);
        cpk_split(v[2 * m + 1], q.z, q.w);
        buf1[m * 128 + t] = q;
    }
}
__device__ __forceinline__ void fftt_ex1_read_pk(const float4* buf1, int t, cpk* v) {
    const int w = t >> 5, L = t & 31;
    const cpk* b = reinterpret_cast<const cpk*>(buf1) + (2 * w + (L >> 4)) * 256 + (L & 15);
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = b[16 * n2];
}
// Exchange 2 + stage 3 with the stage-2 twiddles W_128^(n3 k2) applied on the input side of the two radix-8 groups
// (tw_addr: 2 x 16 TMEM columns, 7 twiddles per group, fetched behind the second round's loads).
__device__ __forceinline__ void fftt_ex2_stage3_pk(uint32_t taddr, uint32_t tw_addr, cpk* v) {
    float r[32];
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
        const int K = 8 * ((k2 >> 2) & 1) + 4 * ((k2 >> 1) & 1) + 2 * ((k2 >> 3) & 1) + (k2 & 1);   // (e2 e1 e3 e0)
        cpk_split(v[k2], r[2 * K], r[2 * K + 1]);
    }
    tm_round(taddr, r);
    tm_st_16x256b_x4(taddr, r);
    tm_st_16x256b_x4(taddr + (16u << 16), r + 16);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    float wa[16], wb[16];
    asm volatile(
#ifdef GR_TMEM_X32
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%64];\n"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%66];\n"
#else
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%64];\n"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%65];\n"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47}, [%66];\n"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%67];\n"
#endif
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]),
          "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]), "=f"(r[17]), "=f"(r[18]),
          "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]),
          "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31]),
          "=f"(wa[0]), "=f"(wa[1]), "=f"(wa[2]), "=f"(wa[3]), "=f"(wa[4]), "=f"(wa[5]), "=f"(wa[6]), "=f"(wa[7]), "=f"(wa[8]),
          "=f"(wa[9]), "=f"(wa[10]), "=f"(wa[11]), "=f"(wa[12]), "=f"(wa[13]), "=f"(wa[14]), "=f"(wa[15]),
          "=f"(wb[0]), "=f"(wb[1]), "=f"(wb[2]), "=f"(wb[3]), "=f"(wb[4]), "=f"(wb[5]), "=f"(wb[6]), "=f"(wb[7]), "=f"(wb[8]),
          "=f"(wb[9]), "=f"(wb[10]), "=f"(wb[11]), "=f"(wb[12]), "=f"(wb[13]), "=f"(wb[14]), "=f"(wb[15])
        : "r"(taddr), "r"(taddr + 16), "r"(tw_addr), "r"(tw_addr + 16));
    cpk a[8], b[8];
#pragma unroll
    for (int n3 = 0; n3 < 8; ++n3) {
        const int K = 4 * (n3 & 1) + 2 * (n3 >> 2) + ((n3 >> 1) & 1);
        a[n3] = cpk_make(r[2 * K], r[2 * K + 1]);
        b[n3] = cpk_make(r[2 * (8 + K)], r[2 * (8 + K) + 1]);
    }
    cpk_dft8_tw(a, wa);
    cpk_dft8_tw(b, wb);
#pragma unroll
    for (int k3 = 0; k3 < 8; ++k3) {
        v[2 * k3] = a[k3];
        v[2 * k3 + 1] = b[k3];
    }
}
#endif
