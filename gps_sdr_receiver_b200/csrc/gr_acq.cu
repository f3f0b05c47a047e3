// Acquisition: raw I/Q -> Doppler wipe-off -> coherent fold -> FFT-2048 -> x conj(code spectrum)
// -> inverse FFT -> |.|^2 non-coherent accumulation -> argmax / mean / std / second peak.
// Only gr_acq_cell tuples (32 bytes per PRN and Doppler bin) leave the GPU's memory system.
//
// Replaces, for a whole PRN x Doppler grid and many recordings per launch,
//   gpsrecv.demodDoppler   src/gpsrecv.py:232-235   (wipe-off, t = (n+1)/fs float32, phase 0)
//   gpsrecv.sweepAllSats   src/gpsrecv.py:241-274   (sum of 1-ms FFTs / avg, x conj spectrum, ifft, abs)
//   gpsrecv.findCodePhase  src/gpsrecv.py:217-227   (argmax, mean, population std, z)
// and the generalisation BASELINE.json names (tcoh coherent x nnoncoh non-coherent).
//
// Two grid kernels:
//   acq_fwd_kernel  CTA = (recording, non-coherent interval, chunk of BASE bins): wipe-off, fold of the tcoh
//                   1-ms blocks in the time domain (sum of FFTs = FFT of the sum), ONE forward FFT per base bin.  Doppler
//                   bins 1 kHz (= fs / 2048) apart have spectra that are circular shifts of each other, so only one
//                   spectrum per 1-kHz class is computed (gr_acq_plan_create); it goes to a scratch array (L2) in natural
//                   order, twice (plain and shifted by one bin: any rotation becomes 16-byte aligned bulk copies).
//   acq_inv_kernel  persistent, 4 CTAs / SM; work item = (recording, Doppler bin, group of 4 PRNs).  Per PRN and
//                   interval: the bin's rotated spectrum by TMA into shared memory, x conj code spectrum (held in tensor
//                   memory), FFT-2048 on the packed FP32 instructions with its second transpose through tensor memory,
//                   |.|^2 accumulated on chip; per PRN the 2048 lags are reduced to one cell.  99 % of the time of a search.
// acq_best_kernel then picks the best bin per (recording, PRN).
#include <stdio.h>
#include <string.h>
#include <vector>

#include "gr_fft2048t.cuh"      // includes gr_fft2048w.cuh and gr_fft2048.cuh
#include "gr_internal.h"

#define GR_ACQ_HOST_CHUNKS 4
struct gr_acq_plan {
    int nprn, nbins, tcoh, nnoncoh, mode, in_format;
    int nbase;             // distinct forward spectra per (recording, interval): bins 1 kHz apart share one, see gr_acq_plan_create
    int exact_nco;         // GPSB200_ACQ_EXACT_NCO: the reference's float32 phase argument for EVERY sample of every bin
    int32_t* d_prns;
    float* d_w32;          // fl32(2*pi*f) per BASE bin (python-float product rounded once, gpsrecv.py:233)
    int32_t* d_bin_base;   // [nbins] base spectrum of a bin
    int32_t* d_bin_shift;  // [nbins] its circular shift in FFT bins, 0..2047
    std::vector<int32_t> h_prns, h_bin_code;   // host copies: small plans travel in the kernel parameters (constant bank)
    // staging for the host entry point
    void* d_in;  size_t in_bytes;
    gr_acq_cell* d_out; size_t out_bytes;
    gr_acq_cell* d_cells; size_t cells_bytes;     // scratch grid of gr_acq_search_*
    gr_acq_best* d_best; size_t best_bytes;
    float2* d_spec; size_t spec_bytes;            // forward spectra of one sub-batch of recordings
    cudaStream_t stream;
    cudaStream_t s_in;                            // host entry points: copy-in stream and per-chunk events (created on first use)
    cudaEvent_t ev_in[GR_ACQ_HOST_CHUNKS];
    bool pipe_ready;
    int last_launches;
    int force_quad;                               // -1: the form of the inverse kernel is chosen per call; 0 / 1: GPSB200_ACQ_QUAD at creation
    int last_inv_form;                            // GR_ACQ_INV_4CTA / GR_ACQ_INV_QUAD of the last run's inverse launches
    // The scratch above (d_spec, d_cells) belongs to the plan, so two *_dev calls of one plan must not overlap: a call on
    // another stream than the previous one first makes its stream wait for the event recorded behind the previous call.
    cudaEvent_t ev_last;
    cudaStream_t last_stream;
    bool has_last;
};

#define GR_ACQ_MAX_PRN_C 64
#define GR_ACQ_MAX_BIN_C 1024
struct AcqArgs {
    const void* samples;
    long long rec_stride;      // samples
    const int32_t* prns;
    const float* w32;          // per base bin
    const int32_t* bin_base;   // per bin: base spectrum, circular shift (FFT bins)
    const int32_t* bin_shift;
    int nrec, nprn, nbins, nbase, ngroups, tcoh, nnoncoh, mode;
    int nchunks, bins_per_chunk;   // forward kernel: base bins per CTA
    int quad_units;                // quad inverse kernel: it takes units 0 .. quad_units-1 of the nrec * nbins (recording, bin) units
    int work0;                     // 4-CTA inverse kernel: it takes work items work0 .. nrec * nbins * ngroups - 1
    int exact_nco;                 // per-sample float32 phase arguments also for tcoh > 1 (see acq_fwd_kernel)
    int need_e1;                   // some bin has an odd shift: the forward kernel writes the second (one-bin shifted) copy too
    float scale;               // 1 / (tcoh * 2048)
    // PRN list and (base << 16 | shift) per bin in the kernel parameters when they fit: the inverse kernel looks them up
    // at job boundaries, where a global load would be an exposed L2 round trip (in_params = 0: use the arrays above)
    int in_params;
    int32_t prn_c[GR_ACQ_MAX_PRN_C];
    int32_t bin_c[GR_ACQ_MAX_BIN_C];
    gr_acq_cell* out;
    float2* spec;              // scratch: forward spectra [nrec][nbase][nnoncoh][2][2048]: natural order, and shifted by one bin
    GrTables tab;
};

// reference sample conversion, gpsrecv.py:168-173: complex64 / 127.5 - (1+1j).
// numpy divides complex64 by the real scalar as  re * fl32(1/127.5)  (Smith's algorithm
// with zero imaginary divisor), then subtracts 1 -- two separately rounded float32 ops.
template <int IN_FMT>
__device__ __forceinline__ cf load_sample(const void* base, long long n) {
    if (IN_FMT == GR_IN_U8IQ) {
        // byte -> integer-valued float by splicing it into the mantissa of 2^23 (PRMT + FADD, exact for 0..255): keeps
        // the quarter-rate I2F unit out of the sample loop; then the reference's two float32 roundings
        const unsigned v = __ldg(reinterpret_cast<const unsigned short*>(base) + n);
        const float bi = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7540)) - 8388608.0f;
        const float bq = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7541)) - 8388608.0f;
        const float scl = 1.0f / 127.5f;
        return cf{__fsub_rn(__fmul_rn(bi, scl), 1.0f), __fsub_rn(__fmul_rn(bq, scl), 1.0f)};
    } else {
        const float2 v = __ldg(reinterpret_cast<const float2*>(base) + n);
        return cf{v.x, v.y};
    }
}

// t[n] = (n+1)/fs in float32 (gpsrecv.py:32-33), arg = fl32(w32 * t[n]) (+ phase 0)
__device__ __forceinline__ float nco_arg(float w32, long long n) {
    const float tsec = __fdiv_rn((float)(n + 1), GR_FS);
    return __fmul_rn(w32, tsec);
}
// ---- kernel 1: forward spectra ------------------------------------------------------------------
// CTA = (recording, non-coherent interval, chunk of base bins); it loops over its bins with the FFT
// twiddles (and, for tcoh = 1, the 16 samples per thread) held in registers: wipe-off, time-domain fold of
// the tcoh blocks, ONE forward FFT per base bin; the spectrum goes to the plan's scratch in natural order, twice
// (E0[m] = X[m], E1[m] = X[m + 1]), which is what the inverse kernel's rotated TMA fetch expects.
//
// NCO: the phase argument is the reference's float32 one, arg = fl32(w32 * fl32((n+1)/fs))
// (gpsrecv.py:32-33, 232-235); sin/cos of it come from a 2-constant Cody-Waite reduction + MUFU
// (|err| < 5e-7 absolute, i.e. below the float32 rounding of the argument itself, 3e-5 rad at 10 kHz x 10 ms).
__device__ __forceinline__ cf nco_fast(float arg) {            // exp(-i arg)
    const float k = rintf(arg * 0.15915494309189535f);
    float r = fmaf(k, -6.28125f, arg);                         // 6.28125 = 201/32: k * C1 is exact
    r = fmaf(k, -1.9353071795864769e-3f, r);                   // 2 pi - 6.28125
    return cf{__cosf(r), -__sinf(r)};
}

template <int IN_FMT, bool kOneBlock>
__global__ void __launch_bounds__(GR_FFT_THREADS) acq_fwd_kernel(const AcqArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf* smem = reinterpret_cast<cf*>(smem_raw);
    const int t = threadIdx.x;
    int id = blockIdx.x;
    const int chunk = id % a.nchunks; id /= a.nchunks;
    const int k = id % a.nnoncoh;
    const int rec = id / a.nnoncoh;
    const int bin0 = chunk * a.bins_per_chunk;
    const int bin1 = min(a.nbase, bin0 + a.bins_per_chunk);

    cf tw1[16], tw2[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float2 u = a.tab.tw1[t * 16 + i];
        const float2 v = a.tab.tw2[(t & 7) * 16 + i];
        tw1[i] = cf{u.x, u.y};
        tw2[i] = cf{v.x, v.y};
    }
    const long long rec_off = (long long)rec * a.rec_stride;
    const void* src = (IN_FMT == GR_IN_U8IQ)
                          ? (const void*)(reinterpret_cast<const uchar2*>(a.samples) + rec_off)
                          : (const void*)(reinterpret_cast<const float2*>(a.samples) + rec_off);
    const long long base0 = (long long)k * a.tcoh * GR_N + t;
    cf s0[kOneBlock ? 16 : 1];
    float ts0[kOneBlock ? 16 : 1];                                 // fl32((n+1)/fs) of this thread's samples: bin-independent
    if (kOneBlock) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            s0[kOneBlock ? j : 0] = load_sample<IN_FMT>(src, base0 + 128 * j);
            ts0[kOneBlock ? j : 0] = __fdiv_rn((float)(base0 + 128 * j + 1), GR_FS);
        }
    }
    cf* Rtab = smem + GR_B1_ELEMS + GR_B2_ELEMS;                    // [tcoh] block rotations of the current bin (multi-block form)
    const int bin_step = (!kOneBlock && a.exact_nco) ? 2 : 1;
    for (int bin = bin0; bin < bin1; bin += bin_step) {
        const float w32 = a.w32[bin];
        cf X[16];
        cf Y[(kOneBlock ? 1 : 16)];                                 // second bin of the pair (exact form only)
        if (kOneBlock) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const cf e = nco_fast(__fmul_rn(w32, ts0[kOneBlock ? j : 0]));   // exp(-i arg), arg as in nco_arg
                X[j].x = s0[kOneBlock ? j : 0].x * e.x - s0[kOneBlock ? j : 0].y * e.y;
                X[j].y = s0[kOneBlock ? j : 0].y * e.x + s0[kOneBlock ? j : 0].x * e.y;
            }
        } else {
            // tcoh > 1: sample n = n0 + 2048 i of block i is rotated by e(n0) * R_i, R_i = exp(-i w 2048 i / fs): 16 + tcoh
            // sin/cos per thread and bin instead of 16 tcoh, and the fold is ONE complex FMA per sample.  (The float32
            // rounding of the reference's per-sample phase argument, up to 3e-5 rad, is then not reproduced sample by
            // sample; the effect on the correlation is below 1e-6 relative.)
            if (a.exact_nco) {
                // reference-exact form: every sample of every block gets the reference's own float32 argument
                // fl32(w32 * fl32((n + 1) / fs)), tcoh x more sin / cos.  For searches whose |w t| is so large (10 kHz x 200 ms
                // = 1.2e4 rad, ulp 1e-3 rad) that the reference's rounding noise exceeds the 1e-4 tolerance.  Two bins at a time:
                // the sample load, its conversion and its time value are shared by both.
                const bool two = bin + 1 < bin1;
                const float w32b = two ? a.w32[bin + 1] : w32;
#pragma unroll
                for (int j = 0; j < 16; ++j) { X[j] = cf{0.f, 0.f}; Y[j] = cf{0.f, 0.f}; }
                // The argument arithmetic runs on the packed instructions (gr_cpk.cuh: lane by lane the operations of tsec_of /
                // nco_fast2, bit-identical results): the time values of two samples per instruction, then the arguments of the
                // two bins per instruction -- this kernel is bound by its issue slots, not by the FP32 pipe.
                const cpk w2 = cpk_make(w32, w32b);
                float fn0 = (float)(base0 + 1);                    // n + 1 of this thread's first sample of block i (exact: < 2^24)
                for (int i = 0; i < a.tcoh; ++i, fn0 += (float)GR_N) {
                    const cpk fnp = cpk_make(fn0, fn0 + 128.0f);
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        float ts[2];
                        cpk_split(tsec_of2(cpk_add(fnp, cpk_bc((float)(128 * j)))), ts[0], ts[1]);
#pragma unroll
                        for (int o = 0; o < 2; ++o) {
                            const cf x = load_sample<IN_FMT>(src, base0 + (long long)i * GR_N + 128 * (j + o));
                            cf e, f;
                            nco_fast2_pair(cpk_mul(w2, cpk_bc(ts[o])), e, f);
                            X[j + o].x = fmaf(x.x, e.x, X[j + o].x); X[j + o].x = fmaf(-x.y, e.y, X[j + o].x);
                            X[j + o].y = fmaf(x.y, e.x, X[j + o].y); X[j + o].y = fmaf(x.x, e.y, X[j + o].y);
                            Y[j + o].x = fmaf(x.x, f.x, Y[j + o].x); Y[j + o].x = fmaf(-x.y, f.y, Y[j + o].x);
                            Y[j + o].y = fmaf(x.y, f.x, Y[j + o].y); Y[j + o].y = fmaf(x.x, f.y, Y[j + o].y);
                        }
                    }
                }
            } else {
            for (int i = t; i < a.tcoh; i += GR_FFT_THREADS)
                Rtab[i] = nco_fast(__fmul_rn(w32, __fdiv_rn((float)(i * GR_N), GR_FS)));
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 16; ++j) X[j] = cf{0.f, 0.f};
            for (int i = 0; i < a.tcoh; ++i) {                     // 16 independent accumulators per block
                const cf R = Rtab[i];
                cf x[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = load_sample<IN_FMT>(src, base0 + (long long)i * GR_N + 128 * j);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    X[j].x = fmaf(x[j].x, R.x, X[j].x); X[j].x = fmaf(-x[j].y, R.y, X[j].x);
                    X[j].y = fmaf(x[j].y, R.x, X[j].y); X[j].y = fmaf(x[j].x, R.y, X[j].y);
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) X[j] = cmul(X[j], nco_fast(nco_arg(w32, base0 + 128 * j)));
            }
        }
        // two copies in natural order, E0[m] = X[m] and E1[m] = X[m + 1]: the inverse kernel fetches a spectrum rotated by
        // any number of bins with 16-byte aligned bulk copies (even rotations of E0, odd ones as even rotations of E1; the
        // second copy is skipped when no bin of the plan has an odd shift)
        auto emit = [&](cf* V, int b) {
            fft2048<true>(V, smem, tw1, tw2, t);
            float2* e0 = a.spec + ((size_t)(rec * a.nbase + b) * a.nnoncoh + k) * (2 * GR_N);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 v = make_float2(V[j].x, V[j].y);
                e0[t + 128 * j] = v;
                if (a.need_e1) e0[GR_N + ((t + 128 * j + GR_N - 1) & (GR_N - 1))] = v;
            }
            __syncthreads();                                       // the FFT buffers (and Rtab) are reused by the next bin
        };
        emit(X, bin);
        if (!kOneBlock && a.exact_nco && bin + 1 < bin1) emit(Y, bin + 1);
    }
}

// reduce one PRN's 2048 scaled lags (st[j] = lag nb + 128 j; nb = fftt_out_base(t)) to its gr_acq_cell.
// Two block barriers; the scratch arrays are free again after the caller's next block barrier.
// Per-(PRN, bin) reduction of the 2048 accumulated lags (16 per thread: value j = lag nb + 128 j) to one cell:
// mean, population std, first maximum, largest value farther than GR_SECOND_PEAK_GUARD lags from it, and the
// maximum's two neighbours.  Two block barriers.  It runs once per 10-20 transforms, but as a plain
// per-lag loop it was a fifth of the kernel's warp time (ncu), so it is written for instruction count:
//  * sums: float per thread and per warp (tree), double only across the four warps;
//  * arg-max: values are >= 0, so their bit patterns order like unsigned integers: two REDUX per warp
//    (max of the patterns, then min of the lags that hold it = first maximum) instead of five shuffle rounds;
//  * second peak: a thread's lags are 128 apart, so at most ONE of them lies in the guard window of the
//    maximum; with the thread's two largest values tracked in the first pass the answer is O(1) per thread;
//  * neighbours: only the two threads that own lags mx -/+ 1 select them.
struct AcqScratch {
    double d[8];
    float f[4];
    int i[4];
    float sec[4];
};
__device__ __forceinline__ float sel16(const float* st, int j) {
    float v = st[0];
#pragma unroll
    for (int q = 1; q < 16; ++q) v = (j == q) ? st[q] : v;
    return v;
}
template <int BAR = 0>      // BAR = 0: the block barrier; else named barrier BAR over 128 threads (one group of a larger CTA)
__device__ __forceinline__ void acq_cell_epilogue(const float* st, int nb, int t, gr_acq_cell* cell, AcqScratch* S, int bar_id = 0) {
    auto sync = [&]() { if (BAR == 0) __syncthreads(); else asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
    float s = 0.f, s2 = 0.f, m1 = -1.f, m2 = -1.f;
    int j1 = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float v = st[j];
        s += v;
        s2 = fmaf(v, v, s2);
        const bool gt = v > m1;                     // strict: the first of equal values stays the maximum
        m2 = gt ? m1 : fmaxf(m2, v);
        j1 = gt ? j : j1;
        m1 = gt ? v : m1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const unsigned wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(m1));
    const int widx = (int)__reduce_min_sync(0xffffffffu, __float_as_uint(m1) == wmax ? (unsigned)(nb + 128 * j1) : 0x7fffffffu);
    const int w = t >> 5;
    if ((t & 31) == 0) { S->d[w] = (double)s; S->d[4 + w] = (double)s2; S->f[w] = __uint_as_float(wmax); S->i[w] = widx; }
    sync();
    float peak = S->f[0];
    int mx = S->i[0];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        const float om = S->f[k];
        const int oi = S->i[k];
        if (om > peak || (om == peak && oi < mx)) { peak = om; mx = oi; }
    }
    // the one lag of this thread that can lie inside [mx - G, mx + G] (circular)
    const int r = (mx - GR_SECOND_PEAK_GUARD - nb) & (GR_N - 1);
    const int jc = ((r + 127) >> 7) & 15;
    const bool inside = ((nb + 128 * jc - (mx - GR_SECOND_PEAK_GUARD)) & (GR_N - 1)) <= 2 * GR_SECOND_PEAK_GUARD;
    float sec = (inside && jc == j1) ? m2 : m1;
    sec = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(sec, 0.f))));
    if ((t & 31) == 0) S->sec[w] = sec;
    const int lo = (mx + GR_N - 1) & (GR_N - 1), hi = (mx + 1) & (GR_N - 1);
    if (((lo - nb) & 127) == 0) cell->em1 = sel16(st, ((lo - nb) & (GR_N - 1)) >> 7);
    if (((hi - nb) & 127) == 0) cell->ep1 = sel16(st, ((hi - nb) & (GR_N - 1)) >> 7);
    sync();
    if (t == 64) {                                   // not warp 0: that one issues the TMA loads
        const double sum = (S->d[0] + S->d[1]) + (S->d[2] + S->d[3]), sum2 = (S->d[4] + S->d[5]) + (S->d[6] + S->d[7]);
        const double mean = sum * (1.0 / GR_N);
        double var = sum2 * (1.0 / GR_N) - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const float sd = sqrtf((float)var);          // float sqrt / divide: the double versions are ~500 cycles of one thread
        cell->mx = mx;
        cell->peak = peak;
        cell->mean = (float)mean;
        cell->std = sd;
        cell->z = (float)((double)peak - mean) / sd;
        cell->second = fmaxf(fmaxf(S->sec[0], S->sec[1]), fmaxf(S->sec[2], S->sec[3]));
    }
}

// ---- kernel 2: x conj(code spectrum), inverse FFT, non-coherent accumulation, cell statistics ------
// CTA = (recording, Doppler bin, group of G PRNs), PRN groups fastest so that the CTAs sharing a forward
// spectrum run together and hit it in L2.  The PRN loop and the interval loop are real loops around ONE
// FFT code path (it stays in the instruction cache).
//
// Register file extension in tensor memory: the butterflies need y[16] + acc[16]; the PRN's conjugate
// spectrum c[16] and the two twiddle sets are per-thread constants that would cost 92 more registers and
// cap the kernel at 3 CTAs / SM.  TMEM (256 KB per SM, otherwise idle: no MMA is issued here) is
// lane-private storage with its own datapath (~850 B/clk/SM measured, tools/ubench): each thread parks its
// constants in its own TMEM lane (tcgen05.st 32x32b) and streams them back 16 words at a time
// (tcgen05.ld) right where they are consumed.  The same TMEM also carries the FFT's second transpose
// (gr_fft2048t.cuh).  128 columns per CTA: c | exchange | stage-2 twiddles | stage-1 twiddles => 4 CTAs / SM.
__device__ __forceinline__ void tm_ld16(uint32_t taddr, float* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
          "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const float* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
        :: "r"(taddr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),
           "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]));
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// v[k] *= tw[k], k = k0..k0+7, twiddles either from registers/shared (ptr) or from TMEM columns
template <bool kTmem>
__device__ __forceinline__ void twiddle8(cf* v, int k0, const cf* tw, uint32_t taddr) {
    if (kTmem) {
        float w[16];
        tm_ld16(taddr + 2 * k0, w);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k0 + k != 0) v[k0 + k] = cmul(v[k0 + k], cf{w[2 * k], w[2 * k + 1]});
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k0 + k != 0) v[k0 + k] = cmul(v[k0 + k], tw[k0 + k]);
    }
}

// v[k0..k0+7] *= the eight twiddles in w (interleaved re, im); k = 0 is the trivial one
__device__ __forceinline__ void twiddle8_regs(cf* v, int k0, const float* w) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k0 + k != 0) v[k0 + k] = cmul(v[k0 + k], cf{w[2 * k], w[2 * k + 1]});
}

__device__ __forceinline__ void twiddle8_pk(cpk* v, int k0, const float* w) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k0 + k != 0) v[k0 + k] = cpk_cmul(v[k0 + k], w[2 * k], w[2 * k + 1]);
}

// Forward spectra staged by TMA: one thread issues a 16 KiB cp.async.bulk for X_{k+1} into a shared-memory
// stage right after the exchange-1 barrier of transform k (every thread has consumed X_k by then); completion
// is tracked by an mbarrier, and the conjugate code spectrum is fetched from TMEM before the wait.  No
// registers are held across the transform for the prefetch (ncu: the wait for X_k was the largest stall of
// the register-fed version).
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
        ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
// stage[m] = E[(m + rot) mod 2048] for an even rotation: one or two bulk copies completing on one mbarrier phase
__device__ __forceinline__ void tma_load_rot(void* stage, const char* E, int rot, uint64_t* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), dst = (uint32_t)__cvta_generic_to_shared(stage);
    const uint32_t head = (uint32_t)(GR_N - rot) * 8u, tail = (uint32_t)rot * 8u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((uint32_t)(GR_N * 8)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(E + tail), "r"(head), "r"(b) : "memory");
    if (rot)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + head), "l"(E), "r"(tail), "r"(b) : "memory");
}
// one lane of a converged warp (the pattern that keeps a TMA issue in the uniform datapath: behind `if (t == 0)` the
// compiler wraps the bulk copy into a leader-election loop and the issuing warp falls ~300 cycles behind its CTA)
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(p));
    return p != 0;
}
// c[] fetch split into issue and wait so that the X wait sits in between
__device__ __forceinline__ void tm_ld16_issue(uint32_t taddr, float* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
          "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tm_ld_wait16(float* r) {      // the registers become valid (and are tied to the wait)
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
                   "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]));
}

// The PRN's conjugate code spectrum of this thread (elements t + 128 j) in the order it is parked in TMEM.
// Scalar form: (re, im) of j = 0..15.  Packed form: the transform reads it as the DATA of the first radix-16 stage
// (the forward spectrum X supplies the "twiddles", see the kernel), i.e. (im, re) -- the swap of the swap-form
// inverse -- in the order the radix-4 butterflies consume it: half H = j0 >> 1, position 4 (j0 & 1) + m <-> j = j0 + 4 m.
template <bool PK>
__device__ __forceinline__ void load_conjspec(float* w, const float2* cs) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float2 v = __ldg(cs + 128 * j);
        if constexpr (PK) {
            const int j0 = j & 3, m = j >> 2, q = 16 * (j0 >> 1) + 2 * (4 * (j0 & 1) + m);
            w[q] = v.y; w[q + 1] = v.x;
        } else {
            w[2 * j] = v.x; w[2 * j + 1] = v.y;
        }
    }
}

// ... the same from this thread's private slots of a shared-memory copy (slot j = 8 bytes at (t + 128 j) * 8)
template <bool PK>
__device__ __forceinline__ void conjspec_from_smem(float* w, const float2* slots) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float2 v = slots[128 * j];
        if constexpr (PK) {
            const int j0 = j & 3, m = j >> 2, q = 16 * (j0 >> 1) + 2 * (4 * (j0 & 1) + m);
            w[q] = v.y; w[q + 1] = v.x;
        } else {
            w[2 * j] = v.x; w[2 * j + 1] = v.y;
        }
    }
}

// FFT: gr_fft2048t.cuh (exchange 1 in shared memory with 128-bit stores on a double-buffered 16 KiB buffer and
// ONE block barrier per transform; exchange 2 + radix-8 through TMEM).  TM bit 1 / bit 2: stage-2 / stage-1
// twiddles in TMEM (else registers).
#define GR_ACQ_INV_SMEM (3 * GR_W_BUF1_BYTES)      // 2 x exchange-1 buffer + forward-spectrum stage = 48 KiB
template <int G, int TM, int MINB, bool PK>
__global__ void __launch_bounds__(GR_FFT_THREADS, MINB) acq_inv_kernel(const AcqArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* buf1 = reinterpret_cast<float4*>(smem_raw);                         // 2 x 16 KiB
    const float2* xs = reinterpret_cast<const float2*>(smem_raw + 2 * GR_W_BUF1_BYTES);  // forward-spectrum stage, 16 KiB, natural order
    __shared__ __align__(8) uint64_t xbar;
    __shared__ AcqScratch scratch;
    __shared__ uint32_t tm_base_sh;
    constexpr int kCols = 64 + ((TM & 2) ? 32 : 0) + ((TM & 4) ? 32 : 0);
    constexpr int kAlloc = kCols <= 64 ? 64 : 128;
    constexpr int kColC = 0, kColX = 32, kColTw2 = 64, kColTw1 = kColTw2 + ((TM & 2) ? 32 : 0);

    const int t = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);       // warp-uniform for the compiler
    const int nwork = a.nrec * a.nbins * a.ngroups;
    int work = a.work0 + blockIdx.x;                           // host guarantees work0 + gridDim.x <= nwork

    if (t == 0) mbar_init(&xbar, 1);
    if (t < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tm_base_sh)), "r"(kAlloc));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = tm_base_sh + ((uint32_t)(32 * (t >> 5)) << 16);

    // work item -> (recording, bin, group); PRN groups fastest
    int grp = work % a.ngroups, bin = (work / a.ngroups) % a.nbins, rec = work / (a.ngroups * a.nbins);
    // The spectrum of (recording, bin, interval k) is the base spectrum of the bin's 1-kHz class rotated by the bin's
    // shift: copy (shift & 1) of the pair the forward kernel wrote, rotated by the even part (tma_load_rot).
    constexpr size_t kStrideK = 2 * (size_t)GR_N * 8;                            // bytes between intervals
    auto prn_of = [&](int i) -> int { return a.in_params ? a.prn_c[i] : a.prns[i]; };
    auto item_src = [&](int wk, int& rot_out) -> const char* {
        const int b = (wk / a.ngroups) % a.nbins, r = wk / (a.ngroups * a.nbins);
        const int code = a.in_params ? a.bin_c[b] : ((a.bin_base[b] << 16) | a.bin_shift[b]);
        const int sh = code & 0xffff;
        rot_out = sh & ~1;
        return reinterpret_cast<const char*>(a.spec) + ((size_t)(r * a.nbase + (code >> 16)) * a.nnoncoh * 2 + (sh & 1)) * (GR_N * 8);
    };
    int rot;
    const char* spec = item_src(work, rot);
    void* xstage = smem_raw + 2 * GR_W_BUF1_BYTES;
    if (warp == 0 && elect_one()) tma_load_rot(xstage, spec, rot, &xbar);

    cf tw1[(TM & 4) ? 1 : 16], tw2[(TM & 2) ? 1 : 16];
    if constexpr (PK) {
        // Packed form: every twiddle is applied on the INPUT side of the next stage, where it folds into the first
        // layer of additions (gr_cpk.cuh).  Stage 2, thread (k1, n3): input n2 carries W_2048^((8 n2 + n3) k1);
        // stage 3, thread (k1loc, k2 = k2lo + 8 h): input n3 of group h carries W_128^(n3 k2).
        static_assert((TM & 6) == 6, "the packed form keeps both twiddle sets in TMEM");
        float w[32];
        const int L = t & 31;
        const int k1 = 4 * (t >> 5) + 2 * (L >> 4) + (L & 1), n3 = (L >> 1) & 7;
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) {
            const float2 u = a.tab.tw1[(8 * n2 + n3) * 16 + k1];
            const int j0 = n2 & 3, m = n2 >> 2, q = 16 * (j0 >> 1) + 2 * (4 * (j0 & 1) + m);
            w[q] = u.x; w[q + 1] = u.y;
        }
        tm_st16(tm + kColTw1, w);
        tm_st16(tm + kColTw1 + 16, w + 16);
        const int k2lo = 4 * ((L >> 2) & 1) + 2 * ((L >> 4) & 1) + ((L >> 1) & 1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int n = 1; n < 8; ++n) {
                const float2 u = a.tab.tw2[n * 16 + k2lo + 8 * h];
                w[16 * h + 2 * (n - 1)] = u.x; w[16 * h + 2 * (n - 1) + 1] = u.y;
            }
            w[16 * h + 14] = 0.f; w[16 * h + 15] = 0.f;
        }
        tm_st16(tm + kColTw2, w);
        tm_st16(tm + kColTw2 + 16, w + 16);
        load_conjspec<true>(w, a.tab.conjspec + (size_t)prn_of(grp * G) * GR_N + t);
        tm_st16(tm + kColC, w);
        tm_st16(tm + kColC + 16, w + 16);
        tm_wait_st();
    } else {
        float w[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float2 u = a.tab.tw1[t * 16 + i];
            w[2 * i] = u.x; w[2 * i + 1] = u.y;
            if (!(TM & 4)) tw1[(TM & 4) ? 0 : i] = cf{u.x, u.y};
        }
        if (TM & 4) { tm_st16(tm + kColTw1, w); tm_st16(tm + kColTw1 + 16, w + 16); }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float2 u = a.tab.tw2[fftt_n3(t) * 16 + i];
            w[2 * i] = u.x; w[2 * i + 1] = u.y;
            if (!(TM & 2)) tw2[(TM & 2) ? 0 : i] = cf{u.x, u.y};
        }
        if (TM & 2) { tm_st16(tm + kColTw2, w); tm_st16(tm + kColTw2 + 16, w + 16); }
        load_conjspec<false>(w, a.tab.conjspec + (size_t)prn_of(grp * G) * GR_N + t);     // first job's conjugate code spectrum
        tm_st16(tm + kColC, w);
        tm_st16(tm + kColC + 16, w + 16);
        tm_wait_st();
    }

    const float sc = (a.mode == GR_ACQ_POW) ? a.scale * a.scale : a.scale;
    const int obase = fftt_out_base(t);
    int par = 0;                       // parity of the transforms done so far: exchange-1 buffer AND mbarrier phase
    int g = 0;

    while (true) {
        // first spectrum of the job after this one (same bin or next work item); fetched during this job's last transform
        int n_work = work, n_g = g + 1;
        if (n_g >= G || grp * G + n_g >= a.nprn) { n_g = 0; n_work = work + gridDim.x; }
        const bool has_next = n_work < nwork;
        const int n_grp = n_work % a.ngroups;
        const char* xjob_next = spec;
        int rot_next = rot;
        if (n_work != work) xjob_next = has_next ? item_src(n_work, rot_next) : nullptr;
        float acc[16];                                          // non-coherent accumulators, 16 lags per thread
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        for (int k = 0; k < a.nnoncoh; ++k) {
          if constexpr (PK) {
            // packed form: one 64-bit register pair per complex value, FADD2 / FMUL2 / FFMA2 (gr_cpk.cuh).
            // Stage 1 computes the radix-16 butterflies of (Im Y, Re Y), Y_j = X_j conj(C_j), as "data = (C.im, C.re),
            // twiddle = conj(X_j)": the spectrum product folds into the first layer of additions like a twiddle.
            cpk y[16];
            float cl[16], ch[16];
            tm_ld16_issue(tm + kColC, cl);
            tm_ld16_issue(tm + kColC + 16, ch);
            mbar_wait(&xbar, par);
            float xl[16], xh[16];                                // conj(X_j) in butterfly order, as the data above
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 v = xs[t + 128 * j];
                const int j0 = j & 3, m = j >> 2, q = 2 * (4 * (j0 & 1) + m);
                if (j0 >> 1) { xh[q] = v.x; xh[q + 1] = -v.y; } else { xl[q] = v.x; xl[q + 1] = -v.y; }
            }
            tm_ld_wait16(cl);
            tm_ld_wait16(ch);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int j0 = j & 3, m = j >> 2, q = 2 * (4 * (j0 & 1) + m);
                y[j] = (j0 >> 1) ? cpk_make(ch[q], ch[q + 1]) : cpk_make(cl[q], cl[q + 1]);
            }
            cpk_dft16_in_tw<0>(y, xl);
            cpk_dft16_in_tw<2>(y, xh);
            cpk_dft16_out(y);
            float4* b1 = buf1 + par * (GR_W_BUF1_BYTES / 16);
            par ^= 1;
            fftt_ex1_write_pk(b1, t, y);
            float wa[16], wb[16];
            tm_ld16_issue(tm + kColTw1, wa);                     // arrives while the block waits at the barrier
            __syncthreads();
            if (k + 1 == a.nnoncoh && has_next) {
                // next PRN's conjugate code spectrum: global -> this thread's private slots of the exchange buffer that stays
                // free until the next job's first transform (cp.async: no registers, no wait; same L1 traffic as a load)
                const float2* cs = a.tab.conjspec + (size_t)prn_of(n_grp * G + n_g) * GR_N + t;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_raw + par * GR_W_BUF1_BYTES) + t * 8;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 1024 * j), "l"(cs + 128 * j) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            {                                                    // every thread has consumed X_k: refill the stage
                const bool more = k + 1 < a.nnoncoh;
                const char* src = more ? spec + (size_t)(k + 1) * kStrideK : xjob_next;
                if (warp == 0 && src != nullptr && elect_one()) tma_load_rot(xstage, src, more ? rot : rot_next, &xbar);
            }
            fftt_ex1_read_pk(b1, t, y);
            tm_ld_wait16(wa);
            tm_ld16_issue(tm + kColTw1 + 16, wb);
            cpk_dft16_in_tw<0>(y, wa);
            tm_ld_wait16(wb);
            cpk_dft16_in_tw<2>(y, wb);
            cpk_dft16_out(y);
            fftt_ex2_stage3_pk(tm + kColX, tm + kColTw2, y);
            if (a.mode == GR_ACQ_POW) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { float yr, yi; cpk_split(y[j], yr, yi); acc[j] = fmaf(yr, yr, fmaf(yi, yi, acc[j])); }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) { float yr, yi; cpk_split(y[j], yr, yi); acc[j] += sqrtf(yr * yr + yi * yi); }
            }
          } else {
            cf y[16];
            float w0[16], w1[16];
            tm_ld16_issue(tm + kColC, w0);
            tm_ld16_issue(tm + kColC + 16, w1);
            mbar_wait(&xbar, par);
#pragma unroll
            for (int j = 0; j < 16; ++j) { const float2 v = xs[t + 128 * j]; y[j] = cf{v.x, v.y}; }
            tm_ld_wait16(w0);
            tm_ld_wait16(w1);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                // Y = X * conjC ; operand of the swap-form inverse = (Im Y, Re Y)
                const cf c0 = cf{w0[2 * j], w0[2 * j + 1]}, c1 = cf{w1[2 * j], w1[2 * j + 1]};
                const cf x0 = y[j], x1 = y[8 + j];
                y[j].x = x0.x * c0.y + x0.y * c0.x;
                y[j].y = x0.x * c0.x - x0.y * c0.y;
                y[8 + j].x = x1.x * c1.y + x1.y * c1.x;
                y[8 + j].y = x1.x * c1.x - x1.y * c1.y;
            }
            // ---- FFT-2048 (gr_fft2048t.cuh); TMEM-resident twiddles are fetched one step ahead of their use ----
            float wa[16], wb[16];
            if (TM & 4) tm_ld16_issue(tm + kColTw1, wa);
            dft16(y);
            if (TM & 4) {
                tm_ld_wait16(wa);
                tm_ld16_issue(tm + kColTw1 + 16, wb);
                twiddle8_regs(y, 0, wa);
                tm_ld_wait16(wb);
                twiddle8_regs(y, 8, wb);
            } else {
                twiddle8<false>(y, 0, tw1, 0);
                twiddle8<false>(y, 8, tw1, 0);
            }
            float4* b1 = buf1 + par * (GR_W_BUF1_BYTES / 16);
            par ^= 1;
            fftw_ex1_write(b1, t, y);
            if (TM & 2) tm_ld16_issue(tm + kColTw2, wa);         // arrives while the block waits at the barrier
            __syncthreads();
            if (k + 1 == a.nnoncoh && has_next) {
                // next PRN's conjugate code spectrum: global -> this thread's private slots of the exchange buffer that stays
                // free until the next job's first transform (cp.async: no registers, no wait; same L1 traffic as a load)
                const float2* cs = a.tab.conjspec + (size_t)prn_of(n_grp * G + n_g) * GR_N + t;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_raw + par * GR_W_BUF1_BYTES) + t * 8;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 1024 * j), "l"(cs + 128 * j) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            {                                                    // every thread has consumed X_k: refill the stage
                const bool more = k + 1 < a.nnoncoh;
                const char* src = more ? spec + (size_t)(k + 1) * kStrideK : xjob_next;
                if (warp == 0 && src != nullptr && elect_one()) tma_load_rot(xstage, src, more ? rot : rot_next, &xbar);
            }
            fftt_ex1_read(b1, t, y);
            dft16(y);
            if (TM & 2) {
                tm_ld_wait16(wa);
                tm_ld16_issue(tm + kColTw2 + 16, wb);
                twiddle8_regs(y, 0, wa);
                tm_ld_wait16(wb);
                twiddle8_regs(y, 8, wb);
            } else {
                twiddle8<false>(y, 0, tw2, 0);
                twiddle8<false>(y, 8, tw2, 0);
            }
            fftt_ex2_stage3(tm + kColX, y);
            if (a.mode == GR_ACQ_POW) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += y[j].x * y[j].x + y[j].y * y[j].y;
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += sqrtf(y[j].x * y[j].x + y[j].y * y[j].y);
            }
          }
        }
        if (has_next) {
            // park the next job's code spectrum (this warp is done with the current one); the copy was started half a
            // transform ago, and the cell reduction's barriers keep the buffer from being overwritten before every
            // thread has read its slots
            float w[32];
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            conjspec_from_smem<PK>(w, reinterpret_cast<const float2*>(smem_raw + par * GR_W_BUF1_BYTES) + t);
            tm_st16(tm + kColC, w);
            tm_st16(tm + kColC + 16, w + 16);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] *= sc;
        acq_cell_epilogue(acc, obase, t, a.out + ((size_t)rec * a.nprn + grp * G + g) * a.nbins + bin, &scratch);
        if (!has_next) break;
        tm_wait_st();
        if (n_work != work) {
            work = n_work; grp = n_grp; bin = (work / a.ngroups) % a.nbins; rec = work / (a.ngroups * a.nbins);
            spec = xjob_next; rot = rot_next;
        }
        g = n_g;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (t < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base_sh), "r"(kAlloc));
}

// ---- kernel 2, "quad" form: one 512-thread CTA per SM = four 128-thread groups sharing the forward spectrum ----------
// The four groups work on four PRNs of the SAME (recording, Doppler bin) and consume the same spectra X_k: one TMA fill of a
// double-buffered stage serves four transforms (a quarter of the TMA writes into shared memory and of the L2 reads of the
// 4-CTA form), and it is issued by whichever warp is the LAST to have read the previous contents (an acq_rel counter in
// shared memory: nobody waits for the stage to become empty, no fixed warp is the straggler).  Groups synchronise among
// themselves only through the stage (a group can run up to two transforms ahead of the slowest); inside a group everything
// is as in acq_inv_kernel with named barriers in place of the block barrier.  Twiddles sit in TMEM once per CTA (lanes are
// shared by warps w, w + 4, ...); 64 + 4 x 64 = 320 of the SM's 512 columns.
#define GR_ACQ_QUAD_SMEM(NS) (4 * 2 * GR_W_BUF1_BYTES + (NS) * GR_W_BUF1_BYTES)      // 4 x double exchange-1 buffer + NS stages (NS = 4: 192 KiB)
template <int NS>      // stages of the forward-spectrum ring (a power of two)
__global__ void __launch_bounds__(512, 1) acq_inv_quad_kernel(const AcqArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t xfull[NS];       // stage i holds its spectrum (TMA completion)
    __shared__ __align__(8) uint64_t xempty[NS];      // all 16 warps have read stage i
    constexpr int kLog = NS == 2 ? 1 : NS == 4 ? 2 : 3;
    static_assert((1 << kLog) == NS, "NS must be 2, 4 or 8");
    __shared__ AcqScratch scratch[4];
    __shared__ uint32_t tm_base_sh;
    constexpr int kAlloc = 512;
    constexpr int kColTw2 = 0, kColTw1 = 32;

    const int tid = threadIdx.x;
    const int t = tid & 127;                                    // thread of the group = FFT thread
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // 0..15, warp-uniform for the compiler
    const int g = warp >> 2;                                    // group 0..3
    const int bar_id = 1 + g;
    const int kColC = 64 + 64 * g, kColX = 96 + 64 * g;
    unsigned char* gsm = smem_raw + (size_t)g * (2 * GR_W_BUF1_BYTES);           // this group's two exchange-1 buffers
    float4* buf1 = reinterpret_cast<float4*>(gsm);
    unsigned char* stage0 = smem_raw + 4 * 2 * GR_W_BUF1_BYTES;                  // stage b at stage0 + b * 16 KiB
    const int nunits = a.quad_units;                            // the first quad_units of the nrec * nbins (recording, bin) units
    const int nrounds = (a.nprn + 3) >> 2;
    const int per_unit = nrounds * a.nnoncoh;                   // stage fills per unit
    const int my_units = ((int)blockIdx.x < nunits) ? (nunits - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nseq = my_units * per_unit;                       // stage fills of this CTA

    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 16); }
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tm_base_sh)), "r"(kAlloc));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = tm_base_sh + ((uint32_t)(32 * (warp & 3)) << 16);

    constexpr size_t kStrideK = 2 * (size_t)GR_N * 8;                            // bytes between intervals
    auto prn_of = [&](int i) -> int { return a.in_params ? a.prn_c[i] : a.prns[i]; };
    // source of stage fill number q of this CTA: unit blockIdx.x + (q / per_unit) gridDim.x, interval (q % per_unit) % nnoncoh
    auto fill_src = [&](int q, int& rot_out) -> const char* {
        const int unit = (int)blockIdx.x + (q / per_unit) * (int)gridDim.x;
        const int k = (q % per_unit) % a.nnoncoh;
        const int b = unit % a.nbins, r = unit / a.nbins;
        const int code = a.in_params ? a.bin_c[b] : ((a.bin_base[b] << 16) | a.bin_shift[b]);
        const int sh = code & 0xffff;
        rot_out = sh & ~1;
        return reinterpret_cast<const char*>(a.spec) + ((size_t)(r * a.nbase + (code >> 16)) * a.nnoncoh * 2 + (sh & 1)) * (GR_N * 8) +
               (size_t)k * kStrideK;
    };
    if (tid == 0) {
        for (int q = 0; q < NS && q < nseq; ++q) {
            int rot;
            const char* src = fill_src(q, rot);
            tma_load_rot(stage0 + q * GR_W_BUF1_BYTES, src, rot, &xfull[q]);
        }
    }
    {
        float w[32];
        if (g == 0) {                                            // twiddles, once per CTA (TMEM lanes are shared across the groups)
            const int L = t & 31;
            const int k1 = 4 * (t >> 5) + 2 * (L >> 4) + (L & 1), n3 = (L >> 1) & 7;
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                const float2 u = a.tab.tw1[(8 * n2 + n3) * 16 + k1];
                const int j0 = n2 & 3, m = n2 >> 2, q = 16 * (j0 >> 1) + 2 * (4 * (j0 & 1) + m);
                w[q] = u.x; w[q + 1] = u.y;
            }
            tm_st16(tm + kColTw1, w);
            tm_st16(tm + kColTw1 + 16, w + 16);
            const int k2lo = 4 * ((L >> 2) & 1) + 2 * ((L >> 4) & 1) + ((L >> 1) & 1);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int n = 1; n < 8; ++n) {
                    const float2 u = a.tab.tw2[n * 16 + k2lo + 8 * h];
                    w[16 * h + 2 * (n - 1)] = u.x; w[16 * h + 2 * (n - 1) + 1] = u.y;
                }
                w[16 * h + 14] = 0.f; w[16 * h + 15] = 0.f;
            }
            tm_st16(tm + kColTw2, w);
            tm_st16(tm + kColTw2 + 16, w + 16);
        }
        if (my_units > 0 && g < a.nprn) {                        // first job's conjugate code spectrum
            load_conjspec<true>(w, a.tab.conjspec + (size_t)prn_of(g) * GR_N + t);
            tm_st16(tm + kColC, w);
            tm_st16(tm + kColC + 16, w + 16);
        }
        tm_wait_st();
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    const float sc = (a.mode == GR_ACQ_POW) ? a.scale * a.scale : a.scale;
    const int obase = fftt_out_base(t);
    int par = 0;                       // exchange-1 buffer of this group's next transform
    int seq = 0;                       // stage fills consumed so far (all warps of the CTA count alike)

    for (int u = 0; u < my_units; ++u) {
        const int unit = (int)blockIdx.x + u * (int)gridDim.x;
        const int bin = unit % a.nbins, rec = unit / a.nbins;
        for (int r = 0; r < nrounds; ++r) {
            const int pi = 4 * r + g;                            // this group's PRN of the round
            const bool active = pi < a.nprn;
            // the job after this one of this group: its PRN of the next round, or (none left in this unit) of the next unit's round 0
            int n_pi = pi + 4;
            bool has_next = true;
            if (n_pi >= a.nprn) { n_pi = g; has_next = u + 1 < my_units; }
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.f;
            // this warp is done with the stage of fill sq
            auto release_stage = [&](int sq) {
                __syncwarp();
                if ((tid & 31) == 0)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&xempty[sq & (NS - 1)])) : "memory");
            };
            // Refill: at the top of fill sq, warp (sq mod 16) refills the stage that was read two fills ago (every warp has almost
            // certainly released it by now) with the spectrum NS - 2 fills ahead.  The duty rotates over the 16 warps: a fixed
            // issuer -- or the last warp to release a stage, which is by construction the slowest -- would carry the ~300 cycles
            // of address arithmetic and TMA issue on every transform and hold its whole group back at the next barrier.
            auto refill = [&](int sq) {
                const int tgt = sq - 2 + NS;
                if (warp == (sq & 15) && sq >= 2 && tgt < nseq) {
                    if (elect_one()) {
                        const int eb = (sq - 2) & (NS - 1);
                        mbar_wait(&xempty[eb], ((sq - 2) >> kLog) & 1);
                        int rot;
                        const char* src = fill_src(tgt, rot);
                        tma_load_rot(stage0 + eb * GR_W_BUF1_BYTES, src, rot, &xfull[eb]);
                    }
                    __syncwarp();
                }
            };
            if (!active) {                                       // a group without a PRN in this round only keeps the stage protocol going
                for (int k = 0; k < a.nnoncoh; ++k, ++seq) {
                    refill(seq);
                    mbar_wait(&xfull[seq & (NS - 1)], (seq >> kLog) & 1);
                    release_stage(seq);
                }
                continue;
            }
            for (int k = 0; k < a.nnoncoh; ++k, ++seq) {
                const int sb = seq & (NS - 1);
                const float2* xs = reinterpret_cast<const float2*>(stage0 + sb * GR_W_BUF1_BYTES);
                cpk y[16];
                float cl[16], ch[16];
                tm_ld16_issue(tm + kColC, cl);
                tm_ld16_issue(tm + kColC + 16, ch);
                refill(seq);
                mbar_wait(&xfull[sb], (seq >> kLog) & 1);
                // conj(X_j) in butterfly order; the two halves of the first butterfly layer one after the other
                {
                    float xl[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int j0 = j & 3, m = j >> 2, q = 2 * (4 * (j0 & 1) + m);
                        if (!(j0 >> 1)) { const float2 v = xs[t + 128 * j]; xl[q] = v.x; xl[q + 1] = -v.y; }
                    }
                    tm_ld_wait16(cl);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int j0 = j & 3, m = j >> 2, q = 2 * (4 * (j0 & 1) + m);
                        if (!(j0 >> 1)) y[j] = cpk_make(cl[q], cl[q + 1]);
                    }
                    cpk_dft16_in_tw<0>(y, xl);
                }
                {
                    float xh[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int j0 = j & 3, m = j >> 2, q = 2 * (4 * (j0 & 1) + m);
                        if (j0 >> 1) { const float2 v = xs[t + 128 * j]; xh[q] = v.x; xh[q + 1] = -v.y; }
                    }
                    tm_ld_wait16(ch);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int j0 = j & 3, m = j >> 2, q = 2 * (4 * (j0 & 1) + m);
                        if (j0 >> 1) y[j] = cpk_make(ch[q], ch[q + 1]);
                    }
                    release_stage(seq);                          // X_k is in registers
                    cpk_dft16_in_tw<2>(y, xh);
                }
                cpk_dft16_out(y);
                float4* b1 = buf1 + par * (GR_W_BUF1_BYTES / 16);
                par ^= 1;
                fftt_ex1_write_pk(b1, t, y);
                float wa[16], wb[16];
                tm_ld16_issue(tm + kColTw1, wa);                 // arrives while the group waits at its barrier
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                if (k + 1 == a.nnoncoh && has_next) {
                    // next job's conjugate code spectrum: global -> this thread's private slots of the group's idle exchange buffer
                    const float2* cs = a.tab.conjspec + (size_t)prn_of(n_pi) * GR_N + t;
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(gsm + par * GR_W_BUF1_BYTES) + t * 8;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 1024 * j), "l"(cs + 128 * j) : "memory");
                    asm volatile("cp.async.commit_group;" ::: "memory");
                }
                fftt_ex1_read_pk(b1, t, y);
                tm_ld_wait16(wa);
                tm_ld16_issue(tm + kColTw1 + 16, wb);
                cpk_dft16_in_tw<0>(y, wa);
                tm_ld_wait16(wb);
                cpk_dft16_in_tw<2>(y, wb);
                cpk_dft16_out(y);
                fftt_ex2_stage3_pk(tm + kColX, tm + kColTw2, y);
                if (a.mode == GR_ACQ_POW) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) { float yr, yi; cpk_split(y[j], yr, yi); acc[j] = fmaf(yr, yr, fmaf(yi, yi, acc[j])); }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) { float yr, yi; cpk_split(y[j], yr, yi); acc[j] += sqrtf(yr * yr + yi * yi); }
                }
            }
            if (has_next) {
                float w[32];
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                conjspec_from_smem<true>(w, reinterpret_cast<const float2*>(gsm + par * GR_W_BUF1_BYTES) + t);
                tm_st16(tm + kColC, w);
                tm_st16(tm + kColC + 16, w + 16);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] *= sc;
            acq_cell_epilogue<1>(acc, obase, t, a.out + ((size_t)rec * a.nprn + pi) * a.nbins + bin, &scratch[g], bar_id);
            if (has_next) tm_wait_st();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base_sh), "r"(kAlloc));
}

// ------------------------------------------------------------------------------------------
#define GR_ACQ_G 4

// Bins whose frequencies differ by a multiple of fs / 2048 = 1 kHz share ONE forward spectrum: the wipe-off factor
// exp(-i 2 pi q 1000 (n + 1) / fs) = exp(-i 2 pi q (n + 1) / 2048) is the same in every 1-ms block, so it commutes with
// the coherent fold and turns into a circular shift of the FFT by q bins (and a constant phase that |.| drops).  The
// +-10 kHz / 500 Hz grid needs 2 forward FFTs per interval instead of 41, the 50-Hz grid 20 instead of 401.  The base
// of a class is its canonical member in [-500, 500) Hz -- whether or not that frequency is itself a bin -- so that a
// cell does not depend on which other bins are in the plan (bin-sharded searches stay bit-identical to unsharded ones)
// and the float32 phase arguments are as small as they can be.
extern "C" int gr_acq_classify_bins(const double* bin_hz, int nbins, int share, int32_t* base, int32_t* shift, double* base_hz) {
    if (!bin_hz || !base || !shift || !base_hz || nbins < 1) { gr_set_error("gr_acq_classify_bins: invalid argument"); return GR_ERR_ARG; }
    const double df = (double)GR_FS / GR_N;
    int nb = 0;
    for (int b = 0; b < nbins; ++b) {
        const double q = share ? floor(bin_hz[b] / df + 0.5) : 0.0;
        const double f0 = bin_hz[b] - q * df;
        if (!(fabs(bin_hz[b]) < 0.5 * GR_FS)) { gr_set_error("gr_acq_classify_bins: bin %d out of range", b); return GR_ERR_ARG; }
        int found = -1;
        for (int i = 0; share && i < nb && found < 0; ++i)
            if (fabs(f0 - base_hz[i]) < 1e-6) found = i;
        if (found < 0) { found = nb; base_hz[nb++] = f0; }
        base[b] = found;
        shift[b] = (int32_t)((((long)q % GR_N) + GR_N) % GR_N);
    }
    return nb;
}

extern "C" int gr_acq_plan_create(const int32_t* prns, int nprn, const double* bin_hz, int nbins, int tcoh_ms,
                                  int nnoncoh, int mode, int in_format, gr_acq_plan** plan) {
    GR_REQUIRE_INIT();
    if (!prns || !bin_hz || !plan || nprn < 1 || nbins < 1 || tcoh_ms < 1 || nnoncoh < 1 ||
        (mode != GR_ACQ_ABS && mode != GR_ACQ_POW) || (in_format != GR_IN_U8IQ && in_format != GR_IN_CF32)) {
        gr_set_error("gr_acq_plan_create: invalid argument");
        return GR_ERR_ARG;
    }
    for (int i = 0; i < nprn; ++i)
        if (prns[i] < 1 || prns[i] > GR_MAX_PRN) { gr_set_error("gr_acq_plan_create: prn %d out of range", prns[i]); return GR_ERR_ARG; }
    gr_acq_plan* p = new gr_acq_plan();
    p->nprn = nprn; p->nbins = nbins; p->tcoh = tcoh_ms; p->nnoncoh = nnoncoh; p->mode = mode; p->in_format = in_format;
    p->d_in = nullptr; p->in_bytes = 0; p->d_out = nullptr; p->out_bytes = 0; p->last_launches = 0; p->last_inv_form = GR_ACQ_INV_4CTA;
    p->d_cells = nullptr; p->cells_bytes = 0; p->d_best = nullptr; p->best_bytes = 0;
    p->d_spec = nullptr; p->spec_bytes = 0;
    p->pipe_ready = false;
    std::vector<int32_t> bin_base(nbins), bin_shift(nbins);
    std::vector<double> base_f(nbins);
    // Form of the forward kernel.  The reference evaluates one float32 phase argument fl32(w32 * fl32((n+1)/fs)) per sample
    // and bin (gpsrecv.py:232-235); its rounding error grows with |w| T.  While that error stays far below the 1e-4
    // tolerance the fast form (one spectrum per 1-kHz class, block rotations) is indistinguishable from it; beyond --
    // largest argument x 2^-24 above 1e-4 rad, e.g. +-10 kHz over 200 ms: 7.5e-4 rad -- the plan reproduces the
    // reference's argument sample by sample and bin by bin ("exact" form).  GPSB200_ACQ_EXACT_NCO=1 / =0 force a form.
    {
        double fmax = 0.0;
        for (int b = 0; b < nbins; ++b) fmax = fmax > fabs(bin_hz[b]) ? fmax : fabs(bin_hz[b]);
        const double argmax = 2.0 * 3.141592653589793 * fmax * (double)tcoh_ms * (double)nnoncoh * 1e-3;
        const char* e = getenv("GPSB200_ACQ_EXACT_NCO");
        p->exact_nco = e ? (atoi(e) != 0) : (argmax * 5.9604644775390625e-8 > 1e-4);
        const char* q = getenv("GPSB200_ACQ_QUAD");              // form of the inverse kernel: chosen per call unless forced here
        p->force_quad = q ? atoi(q) : -1;                        // 0: 4-CTA form, 1: quad form, 2: quad for the whole waves + 4-CTA for the rest
    }
    const int nb = gr_acq_classify_bins(bin_hz, nbins, getenv("GPSB200_ACQ_NOSHARE") == nullptr && !p->exact_nco, bin_base.data(),
                                        bin_shift.data(), base_f.data());
    if (nb < 0) { delete p; return nb; }
    base_f.resize(nb);
    p->nbase = (int)base_f.size();
    p->h_prns.assign(prns, prns + nprn);
    p->h_bin_code.resize(nbins);
    for (int b = 0; b < nbins; ++b) p->h_bin_code[b] = (bin_base[b] << 16) | bin_shift[b];
    std::vector<float> w(p->nbase);
    for (int i = 0; i < p->nbase; ++i) w[i] = (float)(2.0 * 3.141592653589793 * base_f[i]);   // 2*np.pi*freq, then weak -> float32
    GR_CUDA(cudaSetDevice(gr_lib()->device));
    GR_CUDA(cudaMalloc(&p->d_prns, nprn * sizeof(int32_t)));
    GR_CUDA(cudaMalloc(&p->d_w32, p->nbase * sizeof(float)));
    GR_CUDA(cudaMalloc(&p->d_bin_base, nbins * sizeof(int32_t)));
    GR_CUDA(cudaMalloc(&p->d_bin_shift, nbins * sizeof(int32_t)));
    GR_CUDA(cudaMemcpy(p->d_prns, prns, nprn * sizeof(int32_t), cudaMemcpyHostToDevice));
    GR_CUDA(cudaMemcpy(p->d_w32, w.data(), p->nbase * sizeof(float), cudaMemcpyHostToDevice));
    GR_CUDA(cudaMemcpy(p->d_bin_base, bin_base.data(), nbins * sizeof(int32_t), cudaMemcpyHostToDevice));
    GR_CUDA(cudaMemcpy(p->d_bin_shift, bin_shift.data(), nbins * sizeof(int32_t), cudaMemcpyHostToDevice));
    GR_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    GR_CUDA(cudaEventCreateWithFlags(&p->ev_last, cudaEventDisableTiming));
    p->has_last = false;
    p->last_stream = nullptr;
    gr_lib()->live_handles += 1;
    *plan = p;
    return GR_OK;
}

extern "C" int gr_acq_plan_destroy(gr_acq_plan* p) {
    if (!p) return GR_OK;
    cudaFree(p->d_prns);
    cudaFree(p->d_w32);
    cudaFree(p->d_bin_base);
    cudaFree(p->d_bin_shift);
    if (p->d_in) cudaFree(p->d_in);
    if (p->d_out) cudaFree(p->d_out);
    if (p->d_cells) cudaFree(p->d_cells);
    if (p->d_best) cudaFree(p->d_best);
    if (p->d_spec) cudaFree(p->d_spec);
    if (p->pipe_ready) {
        cudaStreamDestroy(p->s_in);
        for (int i = 0; i < GR_ACQ_HOST_CHUNKS; ++i) cudaEventDestroy(p->ev_in[i]);
    }
    cudaStreamDestroy(p->stream);
    cudaEventDestroy(p->ev_last);
    gr_lib()->live_handles -= 1;
    delete p;
    return GR_OK;
}

extern "C" int gr_acq_last_launches(const gr_acq_plan* p) { return p ? p->last_launches : 0; }
extern "C" int gr_acq_last_inverse_form(const gr_acq_plan* p) { return p ? p->last_inv_form : GR_ACQ_INV_4CTA; }
extern "C" int gr_acq_plan_form(const gr_acq_plan* p) { return p && p->exact_nco ? GR_ACQ_FORM_EXACT : GR_ACQ_FORM_FAST; }

static int grow(void** ptr, size_t* have, size_t need) {
    if (need <= *have) return GR_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *have = 0;
    GR_CUDA(cudaMalloc(ptr, need));
    *have = need;
    return GR_OK;
}

#define GR_ACQ_SPEC_CAP (8ull << 30)     // scratch for forward spectra: at most 8 GiB per sub-batch

extern "C" int gr_acq_run_dev(gr_acq_plan* p, const void* d_samples, int nrec, int64_t rec_stride,
                              gr_acq_cell* d_out, void* stream) {
    GR_REQUIRE_INIT();
    if (!p || !d_samples || !d_out || nrec < 1) { gr_set_error("gr_acq_run_dev: invalid argument"); return GR_ERR_ARG; }
    if (nrec > 1 && rec_stride < (int64_t)p->tcoh * p->nnoncoh * GR_N) {
        gr_set_error("gr_acq_run_dev: rec_stride %lld shorter than one recording", (long long)rec_stride);
        return GR_ERR_ARG;
    }
    const size_t spec_per_rec = (size_t)p->nbase * p->nnoncoh * 2 * GR_N * sizeof(float2);
    static const size_t spec_cap = getenv("GPSB200_ACQ_SPEC_MB") ? (size_t)atoi(getenv("GPSB200_ACQ_SPEC_MB")) << 20 : GR_ACQ_SPEC_CAP;
    int sub = (int)(spec_cap / spec_per_rec);
    if (sub < 1) sub = 1;
    if (sub > nrec) sub = nrec;
    int rc = grow((void**)&p->d_spec, &p->spec_bytes, (size_t)sub * spec_per_rec);
    if (rc != GR_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    GR_CUDA(cudaSetDevice(gr_lib()->device));
    if (p->has_last && p->last_stream != s) GR_CUDA(cudaStreamWaitEvent(s, p->ev_last, 0));   // plan-owned scratch: calls of one plan serialise
    const bool one = p->tcoh == 1;
    void (*fwd)(const AcqArgs) = p->in_format == GR_IN_U8IQ ? (one ? acq_fwd_kernel<GR_IN_U8IQ, true> : acq_fwd_kernel<GR_IN_U8IQ, false>)
                                                            : (one ? acq_fwd_kernel<GR_IN_CF32, true> : acq_fwd_kernel<GR_IN_CF32, false>);
    // development switch: GPSB200_ACQ_SCALAR=1 selects the scalar-FP32 form of the same transform (A/B timing)
    static const bool scalar_fp = getenv("GPSB200_ACQ_SCALAR") != nullptr;
    void (*inv)(const AcqArgs) = scalar_fp ? acq_inv_kernel<GR_ACQ_G, 6, 4, false> : acq_inv_kernel<GR_ACQ_G, 6, 4, true>;
    const size_t fwd_smem = GR_FFT_SMEM_BYTES + (one ? 0 : (size_t)p->tcoh * sizeof(cf));
    if (fwd_smem > 200 * 1024) { gr_set_error("gr_acq_run_dev: tcoh too large"); return GR_ERR_ARG; }
    GR_CUDA(cudaFuncSetAttribute(fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem));
    // development switch: GPSB200_ACQ_CTAS=2|3 pads the dynamic shared memory so that fewer CTAs fit an SM (occupancy study)
    static const int ctas_per_sm = getenv("GPSB200_ACQ_CTAS") ? atoi(getenv("GPSB200_ACQ_CTAS")) : 4;
    const int inv_smem = ctas_per_sm == 3 ? 72 * 1024 : ctas_per_sm == 2 ? 110 * 1024 : ctas_per_sm == 1 ? 200 * 1024 : GR_ACQ_INV_SMEM;
    GR_CUDA(cudaFuncSetAttribute(inv, cudaFuncAttributeMaxDynamicSharedMemorySize, inv_smem));
    const size_t bps = p->in_format == GR_IN_U8IQ ? 2 : 8;
    p->last_launches = 0;
    for (int r0 = 0; r0 < nrec; r0 += sub) {
        const int nr = (r0 + sub <= nrec) ? sub : nrec - r0;
        AcqArgs a;
        a.samples = reinterpret_cast<const char*>(d_samples) + (size_t)r0 * (size_t)rec_stride * bps;
        a.rec_stride = rec_stride;
        a.prns = p->d_prns;
        a.w32 = p->d_w32;
        a.in_params = (p->nprn <= GR_ACQ_MAX_PRN_C && p->nbins <= GR_ACQ_MAX_BIN_C && p->nbase < 32768) ? 1 : 0;
        if (a.in_params) {
            memcpy(a.prn_c, p->h_prns.data(), p->nprn * sizeof(int32_t));
            memcpy(a.bin_c, p->h_bin_code.data(), p->nbins * sizeof(int32_t));
        }
        a.bin_base = p->d_bin_base;
        a.bin_shift = p->d_bin_shift;
        a.nbase = p->nbase;
        a.exact_nco = p->exact_nco;
        a.need_e1 = 0;
        for (int b = 0; b < p->nbins; ++b) a.need_e1 |= p->h_bin_code[b] & 1;
        a.nrec = nr;
        a.nprn = p->nprn;
        a.nbins = p->nbins;
        a.ngroups = (p->nprn + GR_ACQ_G - 1) / GR_ACQ_G;
        a.tcoh = p->tcoh;
        a.nnoncoh = p->nnoncoh;
        a.mode = p->mode;
        a.scale = 1.0f / ((float)p->tcoh * (float)GR_N);
        a.out = d_out + (size_t)r0 * p->nprn * p->nbins;
        a.spec = p->d_spec;
        a.tab = gr_lib()->tab;
        // forward grid: enough CTAs to fill the GPU a few times over, as many bins per CTA as that allows
        const long long units = (long long)nr * p->nnoncoh;
        long long nchunks = (4LL * gr_lib()->num_sms * 4 + units - 1) / units;
        if (nchunks > p->nbase) nchunks = p->nbase;
        if (nchunks < 1) nchunks = 1;
        a.bins_per_chunk = (int)((p->nbase + nchunks - 1) / nchunks);
        if (p->exact_nco && p->tcoh > 1) a.bins_per_chunk += a.bins_per_chunk & 1;      // that form takes its bins two at a time
        a.nchunks = (p->nbase + a.bins_per_chunk - 1) / a.bins_per_chunk;
        const long long nfwd = units * a.nchunks;
        const long long ninv = (long long)nr * p->nbins * a.ngroups;
        if (nfwd > 0x7fffffffLL || ninv > 0x7fffffffLL) { gr_set_error("gr_acq_run_dev: grid too large"); return GR_ERR_ARG; }
        fwd<<<(unsigned)nfwd, GR_FFT_THREADS, fwd_smem, s>>>(a);
        // Form of the inverse kernel.  The quad form (one 512-thread CTA per SM, four PRNs of a (recording, bin) unit sharing
        // the staged spectra) is 4 - 5 % faster per transform on launches that fill the GPU many times over (21.3 against
        // 22.4 ms per 512 recordings of configs[1], 5.41 against 5.63 ms per 128: profiles/acq_r02_quad_ab.log), but it hands
        // out work in units of all PRNs x all intervals on one CTA per SM, so its last wave costs a whole unit; the 4-CTA form
        // works in items of 4 PRNs on four CTAs per SM, and the CTAs of a thinly occupied last wave run faster (one CTA alone on
        // an SM runs at 58 % of the rate of four), so its tail is soft: about a quarter of an item.  Measured cross-overs
        // (profiles/acq_r02_quad_ab.log, acq_r02_units_ab.log): 64 recordings of configs[1] equal, 16 recordings 0.73 (4-CTA)
        // against 0.77 ms, a 51-bin shard of configs[3] on 16 recordings 2.05 against 2.16 ms.  Splitting the quad form's units
        // by PRN rounds, or the 4-CTA form's items down to 1 PRN, was built and measured: no gain on small launches over the
        // 4-CTA form as it is, and 5 % lost on large ones (registers of the quad kernel), so neither is kept.
        // Third choice: the quad form for the whole waves and the 4-CTA form for the units of the last, partial wave (one more
        // launch on the same stream).  GPSB200_ACQ_QUAD=0 / 1 (read when the plan is created) forces a single form, =2 this split.
        const long long nunits = (long long)nr * p->nbins;
        const long long sms = gr_lib()->num_sms;
        const long long n_cta = (long long)(ctas_per_sm >= 1 && ctas_per_sm <= 4 ? ctas_per_sm : 4) * sms;
        const long long nrounds = (p->nprn + 3) / 4;
        const double unit_q = (double)(nrounds * p->nnoncoh) * 0.95, item = (double)(GR_ACQ_G * p->nnoncoh);
        auto t_4cta = [&](long long items) { return items > 0 ? ((double)items / (double)n_cta + 0.25) * item : 0.0; };
        const double t_std = t_4cta(ninv);
        const double t_quad = (double)((nunits + sms - 1) / sms) * unit_q;
        // ... and both: the quad form for the whole waves, the 4-CTA form for the units of the last, partial one
        const long long u_full = nunits / sms * sms;
        const double t_both = (double)(u_full / sms) * unit_q + t_4cta((nunits - u_full) * a.ngroups) + 0.1 * item;
        long long uq = 0;                                           // units that go to the quad form
        if (p->force_quad >= 0) uq = p->force_quad == 1 ? nunits : p->force_quad == 2 ? u_full : 0;
        else if (ctas_per_sm == 4) uq = (t_quad <= t_std && t_quad <= t_both) ? nunits : (t_both < t_std ? u_full : 0);
        p->last_inv_form = uq == nunits ? GR_ACQ_INV_QUAD : uq == 0 ? GR_ACQ_INV_4CTA : GR_ACQ_INV_BOTH;
        a.quad_units = (int)uq;
        a.work0 = (int)(uq * a.ngroups);
        if (uq > 0) {
            const long long qgrid = uq < sms ? uq : sms;
            void (*qk)(const AcqArgs) = acq_inv_quad_kernel<4>;
            const int qsmem = GR_ACQ_QUAD_SMEM(4);
            GR_CUDA(cudaFuncSetAttribute(qk, cudaFuncAttributeMaxDynamicSharedMemorySize, qsmem));
            qk<<<(unsigned)qgrid, 512, qsmem, s>>>(a);
            p->last_launches += 1;
        }
        if (uq < nunits) {
            const long long items = ninv - a.work0;
            const long long ninv_grid = items < n_cta ? items : n_cta;
            inv<<<(unsigned)ninv_grid, GR_FFT_THREADS, inv_smem, s>>>(a);
            p->last_launches += 1;
        }
        GR_CUDA(cudaGetLastError());
        p->last_launches += 1;
    }
    GR_CUDA(cudaEventRecord(p->ev_last, s));
    p->last_stream = s;
    p->has_last = true;
    return GR_OK;
}

extern "C" int gr_acq_run_host(gr_acq_plan* p, const void* h_samples, int nrec, int64_t rec_stride,
                               gr_acq_cell* h_out) {
    GR_REQUIRE_INIT();
    if (!p || !h_samples || !h_out || nrec < 1) { gr_set_error("gr_acq_run_host: invalid argument"); return GR_ERR_ARG; }
    GR_CUDA(cudaSetDevice(gr_lib()->device));
    const size_t bps = p->in_format == GR_IN_U8IQ ? 2 : 8;
    const size_t rec_len = (size_t)p->tcoh * p->nnoncoh * GR_N;
    if (nrec > 1 && (size_t)rec_stride < rec_len) { gr_set_error("gr_acq_run_host: rec_stride too short"); return GR_ERR_ARG; }
    const size_t nsamp = (size_t)(nrec - 1) * (size_t)rec_stride + rec_len;
    const size_t in_bytes = nsamp * bps, out_bytes = (size_t)nrec * p->nprn * p->nbins * sizeof(gr_acq_cell);
    if (in_bytes > p->in_bytes) {
        if (p->d_in) cudaFree(p->d_in);
        p->d_in = nullptr; p->in_bytes = 0;
        GR_CUDA(cudaMalloc(&p->d_in, in_bytes));
        p->in_bytes = in_bytes;
    }
    if (out_bytes > p->out_bytes) {
        if (p->d_out) cudaFree(p->d_out);
        p->d_out = nullptr; p->out_bytes = 0;
        GR_CUDA(cudaMalloc((void**)&p->d_out, out_bytes));
        p->out_bytes = out_bytes;
    }
    GR_CUDA(cudaMemcpyAsync(p->d_in, h_samples, in_bytes, cudaMemcpyHostToDevice, p->stream));
    int rc = gr_acq_run_dev(p, p->d_in, nrec, rec_stride, p->d_out, (void*)p->stream);
    if (rc != GR_OK) return rc;
    GR_CUDA(cudaMemcpyAsync(h_out, p->d_out, out_bytes, cudaMemcpyDeviceToHost, p->stream));
    GR_CUDA(cudaStreamSynchronize(p->stream));
    return GR_OK;
}

// ---- best Doppler bin per (recording, PRN): the tuples that leave the device ------------------
__global__ void acq_best_kernel(const gr_acq_cell* __restrict__ cells, const int32_t* __restrict__ prns, int nrec,
                                int nprn, int nbins, gr_acq_best* __restrict__ best) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrec * nprn) return;
    const gr_acq_cell* row = cells + (size_t)i * nbins;
    int bb = 0;
    float bz = row[0].z;
    for (int b = 1; b < nbins; ++b) {
        const float z = row[b].z;
        if (z > bz) { bz = z; bb = b; }
    }
    gr_acq_best o;
    o.prn = prns[i % nprn];
    o.bin = bb;
    o.cell = row[bb];
    best[i] = o;
}

extern "C" int gr_acq_search_dev(gr_acq_plan* p, const void* d_samples, int nrec, int64_t rec_stride,
                                 gr_acq_best* d_best, void* stream) {
    GR_REQUIRE_INIT();
    if (!p || !d_samples || !d_best || nrec < 1) { gr_set_error("gr_acq_search_dev: invalid argument"); return GR_ERR_ARG; }
    int rc = grow((void**)&p->d_cells, &p->cells_bytes, (size_t)nrec * p->nprn * p->nbins * sizeof(gr_acq_cell));
    if (rc != GR_OK) return rc;
    rc = gr_acq_run_dev(p, d_samples, nrec, rec_stride, p->d_cells, stream);
    if (rc != GR_OK) return rc;
    const int n = nrec * p->nprn;
    acq_best_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p->d_cells, p->d_prns, nrec, p->nprn, p->nbins, d_best);
    GR_CUDA(cudaGetLastError());
    p->last_launches += 1;
    GR_CUDA(cudaEventRecord(p->ev_last, (cudaStream_t)stream));      // d_cells is read by acq_best_kernel
    return GR_OK;
}

extern "C" int gr_acq_search_host(gr_acq_plan* p, const void* h_samples, int nrec, int64_t rec_stride,
                                  gr_acq_best* h_best) {
    GR_REQUIRE_INIT();
    if (!p || !h_samples || !h_best || nrec < 1) { gr_set_error("gr_acq_search_host: invalid argument"); return GR_ERR_ARG; }
    GR_CUDA(cudaSetDevice(gr_lib()->device));
    const size_t bps = p->in_format == GR_IN_U8IQ ? 2 : 8;
    const size_t rec_len = (size_t)p->tcoh * p->nnoncoh * GR_N;
    if (nrec > 1 && (size_t)rec_stride < rec_len) { gr_set_error("gr_acq_search_host: rec_stride too short"); return GR_ERR_ARG; }
    const size_t nsamp = (size_t)(nrec - 1) * (size_t)rec_stride + rec_len;
    const size_t in_bytes = nsamp * bps, best_bytes = (size_t)nrec * p->nprn * sizeof(gr_acq_best);
    int rc = grow(&p->d_in, &p->in_bytes, in_bytes);
    if (rc != GR_OK) return rc;
    rc = grow((void**)&p->d_best, &p->best_bytes, best_bytes);
    if (rc != GR_OK) return rc;
    if (nrec < 2 * GR_ACQ_HOST_CHUNKS) {
        GR_CUDA(cudaMemcpyAsync(p->d_in, h_samples, in_bytes, cudaMemcpyHostToDevice, p->stream));
        rc = gr_acq_search_dev(p, p->d_in, nrec, rec_stride, p->d_best, (void*)p->stream);
        if (rc != GR_OK) return rc;
    } else {
        // batches: the recordings go in as GR_ACQ_HOST_CHUNKS pieces on a copy stream, each searched as soon as it has
        // arrived, so that only the first piece's host-to-device copy is not hidden behind the kernels
        if (!p->pipe_ready) {
            GR_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
            for (int i = 0; i < GR_ACQ_HOST_CHUNKS; ++i) GR_CUDA(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
            p->pipe_ready = true;
        }
        const int per = (nrec + GR_ACQ_HOST_CHUNKS - 1) / GR_ACQ_HOST_CHUNKS;
        int launches = 0;
        for (int c = 0, r0 = 0; r0 < nrec; ++c, r0 += per) {
            const int nr = r0 + per <= nrec ? per : nrec - r0;
            const size_t off = (size_t)r0 * (size_t)rec_stride * bps;
            const size_t len = ((size_t)(nr - 1) * (size_t)rec_stride + rec_len) * bps;
            GR_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(p->d_in) + off, reinterpret_cast<const char*>(h_samples) + off, len,
                                    cudaMemcpyHostToDevice, p->s_in));
            GR_CUDA(cudaEventRecord(p->ev_in[c], p->s_in));
            GR_CUDA(cudaStreamWaitEvent(p->stream, p->ev_in[c], 0));
            rc = gr_acq_search_dev(p, reinterpret_cast<char*>(p->d_in) + off, nr, rec_stride, p->d_best + (size_t)r0 * p->nprn, (void*)p->stream);
            if (rc != GR_OK) return rc;
            launches += p->last_launches;
        }
        p->last_launches = launches;
    }
    GR_CUDA(cudaMemcpyAsync(h_best, p->d_best, best_bytes, cudaMemcpyDeviceToHost, p->stream));
    GR_CUDA(cudaStreamSynchronize(p->stream));
    return GR_OK;
}

// ---- debug hook: FP32 FFMA peak (the acquisition roofline's denominator) -----------------------
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float a, float b) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (float)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    if (s == 123.456f) out[0] = s;       // never true; keeps the chains alive
}

extern "C" int gr_debug_fp32_peak(int iters, double* tflops) {
    GR_REQUIRE_INIT();
    if (iters < 1 || !tflops) { gr_set_error("gr_debug_fp32_peak: invalid argument"); return GR_ERR_ARG; }
    GR_CUDA(cudaSetDevice(gr_lib()->device));
    float* d;
    GR_CUDA(cudaMalloc(&d, 4));
    const int blocks = gr_lib()->num_sms * 8;
    cudaEvent_t e0, e1;
    GR_CUDA(cudaEventCreate(&e0));
    GR_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        GR_CUDA(cudaEventRecord(e0));
        fma_peak_kernel<<<blocks, 256>>>(d, iters, 0.999f, 0.001f);
        GR_CUDA(cudaEventRecord(e1));
        GR_CUDA(cudaEventSynchronize(e1));
        float ms;
        GR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    GR_CUDA(cudaGetLastError());
    *tflops = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks / ((double)best * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return GR_OK;
}

// ---- debug hook: the CTA FFT on its own ------------------------------------------------------
__global__ void __launch_bounds__(GR_FFT_THREADS) fft_debug_kernel(const float2* in, float2* out, int inverse, GrTables tab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf* smem = reinterpret_cast<cf*>(smem_raw);
    const int t = threadIdx.x;
    cf tw1[16], tw2[16], v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float2 u = tab.tw1[t * 16 + k];
        const float2 w = tab.tw2[(t & 7) * 16 + k];
        tw1[k] = cf{u.x, u.y};
        tw2[k] = cf{w.x, w.y};
    }
    const float2* src = in + (size_t)blockIdx.x * GR_N;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float2 s = src[t + 128 * j];
        v[j] = inverse ? cf{s.y, s.x} : cf{s.x, s.y};
    }
    fft2048<true>(v, smem, tw1, tw2, t);
    float2* dst = out + (size_t)blockIdx.x * GR_N;
#pragma unroll
    for (int j = 0; j < 16; ++j) dst[t + 128 * j] = inverse ? make_float2(v[j].y, v[j].x) : make_float2(v[j].x, v[j].y);
}

extern "C" int gr_debug_fft2048(const float* h_in, float* h_out, int batch, int inverse) {
    GR_REQUIRE_INIT();
    if (!h_in || !h_out || batch < 1) { gr_set_error("gr_debug_fft2048: invalid argument"); return GR_ERR_ARG; }
    GR_CUDA(cudaSetDevice(gr_lib()->device));
    float2 *d_in, *d_out;
    const size_t bytes = (size_t)batch * GR_N * sizeof(float2);
    GR_CUDA(cudaMalloc(&d_in, bytes));
    GR_CUDA(cudaMalloc(&d_out, bytes));
    GR_CUDA(cudaMemcpy(d_in, h_in, bytes, cudaMemcpyHostToDevice));
    GR_CUDA(cudaFuncSetAttribute(fft_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_FFT_SMEM_BYTES));
    fft_debug_kernel<<<batch, GR_FFT_THREADS, GR_FFT_SMEM_BYTES>>>(d_in, d_out, inverse, gr_lib()->tab);
    GR_CUDA(cudaGetLastError());
    GR_CUDA(cudaMemcpy(h_out, d_out, bytes, cudaMemcpyDeviceToHost));
    cudaFree(d_in);
    cudaFree(d_out);
    return GR_OK;
}
