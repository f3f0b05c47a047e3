// Batched channel tracker: one CTA (128 threads) per tracked satellite, persistent over
// `n_epochs` consecutive epochs, all loop state resident on the device.
//
// Replaces, for every channel of a bank in ONE launch,
//   gpslib.SatStream.process         src/gpslib.py:1141-1210   (epoch state machine)
//   SatStream.demodDoppler           src/gpslib.py:1343-1346   (NCO wipe-off, float32 time base)
//   SatStream.cacodeCorr             src/gpslib.py:1315-1327   (sum of CORR_AVG 1-ms FFTs, x conj spectrum, |ifft|)
//   SatStream.findCodePhase/fit      src/gpslib.py:1268-1304   (argmax, mean/std, sub-sample fit)
//   SatStream.corrQuality            src/gpslib.py:1331-1339
//   SatStream.decodeData             src/gpslib.py:1394-1446   (1-ms prompt integrate & dump, carry-over, edges)
//   SatStream.phaseLockedLoop        src/gpslib.py:1215-1262
//   SatStream.sweepFrequency/getCorrMax  src/gpslib.py:1350-1380 (per-channel re-acquisition)
//   gpsrecv worker pool / satCalc    src/gpsrecv.py:300-417    (one OS process per channel -> one CTA per channel)
//
// Numerical design (DESIGN.md "tracking kernel"):
//   * NCO factorisation.  The reference rotates sample n of the epoch by exp(-i(phi + w t_n)),
//     t_n = (n+1)/fs.  With n = 2048 b + i this is  r_i * R_b,  r_i = exp(-i(phi + w (i+1)/fs)),
//     R_b = exp(-i w b 1ms), and r_i = r_t * rho_j for i = t + 128 j.  So one epoch costs each
//     thread two sincosf instead of N_CYC*16, the per-sample work is one complex FMA, and the
//     coherent fold of the CORR_AVG blocks happens in the time domain (sum of FFTs = FFT of sum).
//   * uint8 I/Q enters the FMAs as integer-valued floats; the affine map x = b/127.5 - 1 of the
//     reader (gpsrecv.py:168-173) is applied once per sum:  sum x q = s sum b q - (1+i) sum q.
//   * Prompt integration runs in "row-rotated" layout: thread t, row j handles sample
//     128*(delay>>7) + 2048(k-1) + t + 128 j of pass k, so only row 0 straddles a code-period
//     boundary; the true 1-ms sums are  S_k - B_k + B_{k+1}  (B = the part of row 0 before the
//     boundary).  4 FMA per sample, coalesced loads, no divergence.
//   * The loop filter (arctan discriminator, pi-unwrap, lock detector, DF FIFO), the correlation
//     quality FIFO, the sweep state machine and the edge detector run in the same kernel; only a
//     gr_epoch_out record per channel and epoch leaves the SM.
//   * Data movement: the epoch's raw block arrives in shared memory by TMA (cp.async.bulk + mbarrier) while the
//     previous epoch's serial tail runs; the whole channel state (scalars + DF / CORRLST rings) is resident in
//     shared memory for the life of the launch; the 448-byte record is assembled in shared memory and leaves by a
//     TMA bulk store (no global store sits in front of a block barrier).
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <type_traits>
#include <vector>

#include "gr_fft2048.cuh"
#include "gr_cpk.cuh"
#include "gr_internal.h"

#define GR_DF_CAP 128            // NO_SEC = 1024 // N_CYC <= 128  (N_CYC >= 8)
#define GR_CL_CAP (60 * 128)     // CORRLST_NO = 60 * NO_SEC
#define GR_TWO_PI_D 6.283185307179586
#define GR_TWO_PI_F 6.2831855f   // float32(2*np.pi)
#define GR_PI_F 3.1415927f       // float32(np.pi)

// ---- per-channel state (device global memory) -------------------------------------------
struct GrChanHot {
    int32_t active, prn, rec, delay;
    int32_t locked, sweep, sweep_req, freq_weak;      // freq_weak: FREQ is a python float (not np.float32)
    int32_t std_weak, ms_time, edge0, edge_len;       // EDGES[0], len(EDGES)
    int32_t rep_sweep, df_len, df_head, df_save_len;
    int32_t freq_save_weak, cl_len, cl_head, cl_sum;  // CORRLST ring
    int32_t cl_sum_last, carry_cnt, pad0, pad1;
    int64_t prev_stream_no;
    double freq, freq_save, std_dev, prev_signal;
    double carry_re, carry_im;                         // sum of PREV_SAMPLES
    double max_corr, corr_q, corr_l;
    float phase, amplitude;
};
struct GrChanS {                 // the part of a channel's state that lives in shared memory while the kernel runs
    GrChanHot h;                 // scalars
    float df[GR_DF_CAP];         // DF FIFO (ring)
    float df_save[GR_DF_CAP];
};
struct GrChan {                  // global memory: GrChanS (same layout, copied in / out once per launch) + the CORRLST ring
    GrChanHot h;
    float df[GR_DF_CAP];
    float df_save[GR_DF_CAP];
    int8_t cl[GR_CL_CAP];        // CORRLST FIFO (ring) of +1/-1: one entry in, at most two out per epoch (exact integer running
                                 // sums in `h`), touched by thread 0 only -- it stays in global memory, the two outgoing entries
                                 // are fetched into registers at the start of the epoch
};
static_assert(offsetof(GrChan, cl) == sizeof(GrChanS), "GrChanS is the prefix of GrChan");

struct gr_track_bank {
    gr_track_cfg cfg;
    GrChan* d_state;
    int32_t* d_slots;
    std::vector<int> slot_used;      // host mirror
    std::vector<int> slot_rec;       // recording index of each used slot
    std::vector<int> active;         // sorted active slots
    bool slots_dirty;
    cudaStream_t stream;             // for the host entry point
    cudaEvent_t ev_last;             // recorded behind the last launch (on whatever stream the caller gave): host edits wait on it
    bool in_flight;
    int max_rec;                     // largest recording index among the active channels
    void* d_in;  size_t in_bytes;
    gr_epoch_out* d_out; size_t out_bytes;
    int last_launches;
    int exact_nco;                   // the reference's float32 phase argument per sample (default) | factorised NCO (GPSB200_TRK_FAST_NCO=1)
    bool pipe_ready;                 // streams/events of the host pipeline
    cudaStream_t s_in, s_out;
    cudaEvent_t ev_in[2], ev_run[2], ev_out[2];
};

struct TrackArgs {
    const void* samples;
    long long rec_stride;   // samples
    long long smp_time;     // SMP_TIME of the first epoch
    int n_epochs, n_active;
    int out_tma;            // records leave by cp.async.bulk (16-byte aligned output array)
    int stage;              // raw I/Q of each epoch staged in shared memory by TMA (u8 input, 16-byte aligned recordings)
    int buf_bytes;          // size of the scratch buffer in front of TrackSmem (FFT buffers / staged prompt rows)
    int part_rows;          // prompt rows staged per reduction round
    const int32_t* slots;
    GrChan* state;
    gr_epoch_out* out;
    gr_track_cfg cfg;
    GrTables tab;
};

// ---- shared-memory scratch -------------------------------------------------------------------
#define GR_PART_ROWS 17                                   // prompt passes staged per reduction round (at most)
#define GR_TRACK_BUF_BYTES (GR_PART_ROWS * 128 * 16)      // 34816 >= GR_FFT_SMEM_BYTES; FFT buffers alias it
// The "dense" form of the kernel (batches: more channels than 2 x SMs) fits THREE CTAs per SM at 8-ms epochs: the FFT
// runs in its one-buffer mode and the prompt rows are staged n_cyc + 1 at a time, so the scratch buffer shrinks from
// 34 KB to 18 KB (71 KB per CTA with the 32 KB raw stage), and the register cap drops from 190 to 168.

struct TrackSmem {
    gr_epoch_out out[2];                 // the epoch's record is assembled here (448 B each, 16-byte aligned) and leaves by TMA
    unsigned long long rawbar;           // mbarrier of the raw-sample stage (first member: 8-byte aligned)
    GrChanS CH;                          // this channel's scalars + DF rings for the life of the kernel
    cf rho[16];                          // exp(-i w 128 j / fs)
    cf Rm[GR_MAX_NCYC + 2];              // Rm[k] = R_{k-1} = exp(-i w (k-1) ms), k = 0..n_cyc
    float4 red[GR_MAX_NCYC + 2];         // reduced prompt rows: (S_k, B_k)
    float qred[8][6];                    // per-warp sums of q: all rows, masked row 0, wrapped rows
    cf sigma[8];                         // vector form: exp(-i w i / fs), i = 0..7
    cf qb[8];                            // vector form: the replica values of the samples between the window start 8 (d >> 3) and d
    cf qbsum;
    int dq;                              // vector form: chunk rotation of the fold pass = DELAY >> 3 at the start of the epoch
    double sh_d[8];
    float sh_f[8];
    int sh_i[8];
    float nb[2];                         // corr[mx-1], corr[mx+1]
    double pr_re[3][GR_MAX_PROMPT], pr_im[3][GR_MAX_PROMPT];   // prompt means (complex128 in the reference), one copy per tail warp
    float4 wxs[3][GR_MAX_NCYC + 2];      // (X_k, XB_k) per tail warp
    float qsum[6];                       // qred summed over the warps
    double carry0_re, carry0_im;         // PREV_SAMPLES at the start of the epoch (the tail warps read these, warp 0 writes the new ones)
    double min_edge;                     // 3 STD_DEV of the previous epoch
    int carry0_cnt, locked_in, do_sweep;
    float ph[GR_MAX_PROMPT + 2];         // phase / realPhase
    // epoch scalars (written by thread 0, read by all after a barrier)
    float w32, phase32;
    int branch_sweep, delay, corr_delay, n_prompt;
    double z, code_phase, cmean, cstd;
    float c3[3];
};

// ---- sample access ------------------------------------------------------------------------------
// u8 -> float without the (quarter-rate) I2F unit: splice the byte into the mantissa of 2^23 and subtract 2^23
// (exact for 0..255): one PRMT + one FADD per component.
__device__ __forceinline__ cf u8pair_to_cf(unsigned v16) {
    const float x = __uint_as_float(__byte_perm(v16, 0x4B000000u, 0x7540)) - 8388608.0f;
    const float y = __uint_as_float(__byte_perm(v16, 0x4B000000u, 0x7541)) - 8388608.0f;
    return cf{x, y};          // integer-valued; the affine map x/127.5 - 1 is applied per sum
}
// kStage: `base` is the epoch's raw block in shared memory (see track_kernel), else global memory
template <int IN_FMT, bool kStage>
__device__ __forceinline__ cf load_raw(const void* base, long long n) {
    if (IN_FMT == GR_IN_U8IQ && kStage) {
        return u8pair_to_cf(reinterpret_cast<const unsigned short*>(base)[(int)n]);
    } else if (IN_FMT == GR_IN_U8IQ) {
        return u8pair_to_cf(__ldg(reinterpret_cast<const unsigned short*>(base) + n));
    } else {
        const float2 v = __ldg(reinterpret_cast<const float2*>(base) + n);
        return cf{v.x, v.y};
    }
}

// x = s*b - (1+i)  applied to a sum:  sum x q = s * sum(b q) - (1+i) * sum(q)
template <int IN_FMT>
__device__ __forceinline__ cf affine_sum(cf sb, cf sq) {
    if (IN_FMT == GR_IN_U8IQ) {
        const float s = 1.0f / 127.5f;
        // (1+i) * (a+ib) = (a-b) + i(a+b)
        return cf{fmaf(sb.x, s, -(sq.x - sq.y)), fmaf(sb.y, s, -(sq.x + sq.y))};
    } else {
        return sb;
    }
}

__device__ __forceinline__ cf expmi(float a) {   // exp(-i a)
    float s, c;
    sincosf(a, &s, &c);
    return cf{c, -s};
}
__device__ __forceinline__ cf expmi_d(double a) {   // exp(-i a), argument reduced in double
    a -= GR_TWO_PI_D * rint(a * (1.0 / GR_TWO_PI_D));      // no FP64 division on the epoch's critical path
    return expmi((float)a);
}

// fmodf(p, float32(2 pi)) for |p| < 2^20 without the generic remainder loop: the quotient from one FP64 multiply (it may
// be off by one), the remainder p - q m exactly by one FP64 FMA (q m has <= 44 significant bits), then the off-by-one
// fix-up; the result is exactly representable in float32, like fmodf's.
__device__ __forceinline__ float fmod_2pi_f32(float p) {
    const double m = (double)GR_TWO_PI_F;
    const double q = trunc((double)p * (1.0 / (double)GR_TWO_PI_F));
    double r = fma(-q, m, (double)p);
    if (p >= 0.f) { if (r < 0.0) r += m; else if (r >= m) r -= m; }
    else { if (r > 0.0) r -= m; else if (r <= -m) r += m; }
    return (float)r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- correlation: folded samples F (natural FFT layout) -> statistics of |ifft(fft(F)/avg * conjC)| --
// Leaves mx in S->sh_i[4], z / mean / std / corr[mx-1..mx+1] in S, written by thread 0; the caller synchronises.
template <bool kOneBuf, bool kSub = false>
__device__ __forceinline__ void corr_and_stats(cf* F, const float2* __restrict__ cs, float scale, cf* fftbuf,
                                               const cf* tw1, const cf* tw2, int t, TrackSmem* S) {
    // kSub: the 128 threads of this transform are the first half of a 256-thread CTA: named barrier 1 instead of barrier 0
    auto sync = [&]() { if (kSub) asm volatile("bar.sync 1, 128;" ::: "memory"); else __syncthreads(); };
    // ONE FFT body run twice (forward, then the swap-form inverse) instead of two inlined copies: 1 100 fewer instructions per
    // kernel.  The epoch is ~10 000 instructions of straight-line code and the CTAs of an SM are never in the same place, so the
    // instruction cache is a shared resource here: 16.4 -> 15.8 us per epoch at 3 CTAs per SM, 9.09 -> 8.96 us for one recording
    // (profiles/track_r02_onefft_ab.log); same arithmetic in the same order, records bit-identical.
    cf y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] = F[j];
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        fft2048<!kSub, 1, kOneBuf>(y, fftbuf, tw1, tw2, t);
        if (pass == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 c = __ldg(cs + t + 128 * j);
                const cf f = y[j];
                y[j].x = f.x * c.y + f.y * c.x;      // swap form of the inverse transform
                y[j].y = f.x * c.x - f.y * c.y;
            }
        }
    }
    float st[16];
    float s = 0.f, s2 = 0.f, mx = -1.f;
    int idx = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float v = sqrtf(y[j].x * y[j].x + y[j].y * y[j].y) * scale;
        st[j] = v;
        s += v;
        s2 = fmaf(v, v, s2);
        if (v > mx) { mx = v; idx = t + 128 * j; }
    }
    double ds = (double)s, ds2 = (double)s2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        ds2 += __shfl_xor_sync(0xffffffffu, ds2, o);
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (om > mx || (om == mx && oi < idx)) { mx = om; idx = oi; }
    }
    const int w = t >> 5;
    if ((t & 31) == 0) { S->sh_d[w] = ds; S->sh_d[4 + w] = ds2; S->sh_f[w] = mx; S->sh_i[w] = idx; }
    sync();
    double sum = 0.0, sum2 = 0.0;
    float bm = -1.f;
    int bi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        sum += S->sh_d[k];
        sum2 += S->sh_d[4 + k];
        const float om = S->sh_f[k];
        const int oi = S->sh_i[k];
        if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
    }
    const int lo = (bi + GR_N - 1) & (GR_N - 1), hi = (bi + 1) & (GR_N - 1);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int n = t + 128 * j;
        if (n == lo) S->nb[0] = st[j];
        if (n == hi) S->nb[1] = st[j];
    }
    sync();
    if (t == 0) {
        const double mean = sum * (1.0 / GR_N);
        double var = sum2 * (1.0 / GR_N) - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const float sd = sqrtf((float)var);          // float sqrt / divide: the double versions are ~500 cycles of this one thread
        S->cmean = mean;
        S->cstd = (double)sd;
        S->z = (double)((float)((double)bm - mean) / sd);
        S->c3[0] = S->nb[0];
        S->c3[1] = bm;
        S->c3[2] = S->nb[1];
        S->sh_i[4] = bi;
    }
    // no barrier: thread 0 goes on with the decision; the caller synchronises
}

// gpslib.py:1268-1290 fitCodePhase (double arithmetic on the float32 correlation values)
// The three values are float32 correlation magnitudes; their differences are formed in double (exact), the two
// quotients in float32 (6e-8 relative: 3e-8 sample, against the 7e-4 sample = 0.1 m bound) -- two FP64 divisions would be
// ~300 cycles on the epoch's critical path.
__device__ __forceinline__ double fit_code_phase(int mx, double lo, double c, double hi) {
    const float num = (float)(0.5 * (hi - lo));
    const float tri = num / (float)(lo > hi ? c - hi : c - lo);
    const float par = num / (float)(2.0 * c - hi - lo);
    return (double)mx + 0.5 * ((double)tri + (double)par);
}

// Coherent fold of `nblk` 1-ms blocks starting at block `first` (natural FFT layout):
//   F[j] = r_t rho_j * sum_b R_b x_b[t + 128 j]
template <int IN_FMT, bool kStage>
__device__ __forceinline__ void fold_blocks(cf* F, const void* src, int first, int nblk, cf rt, int t,
                                            const TrackSmem* S) {
    cf A[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) A[j] = cf{0.f, 0.f};
    cf rsum = cf{0.f, 0.f};
    for (int b = first; b < first + nblk; ++b) {
        const cf R = S->Rm[b + 1];
        rsum = cadd(rsum, R);
        const long long base = (long long)b * GR_N + t;
        cf x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = load_raw<IN_FMT, kStage>(src, base + 128 * j);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            A[j].x = fmaf(x[j].x, R.x, A[j].x);
            A[j].x = fmaf(-x[j].y, R.y, A[j].x);
            A[j].y = fmaf(x[j].x, R.y, A[j].y);
            A[j].y = fmaf(x[j].y, R.x, A[j].y);
        }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const cf r = cmul(rt, S->rho[j]);
        F[j] = cmul(affine_sum<IN_FMT>(A[j], rsum), r);
    }
}

// ---- vector form of the two sample passes (staged uint8 I/Q) ---------------------------------------------------------
// The staged block is read 16 bytes (8 samples) at a time: a 2048-sample block is 256 chunks, thread t owns the chunks
// u0 = (t + dq) & 255 and u1 = u0 ^ 128 of EVERY block (dq = DELAY >> 3 puts the code-period boundary into thread 0's
// first chunk).  Arithmetic on the packed FP32 instructions (gr_cpk.cuh): a sample costs two PRMT + one FADD2 (bytes ->
// integer-valued floats) and two FFMA2 (complex multiply-accumulate), against one LDS.U16, two PRMT, two FADD, four FFMA
// and address / select work in the scalar form (ncu, profiles/track_r01_v5_regions.txt: 115 IADD3 + 100 FSEL + 96 PRMT per
// 126 FFMA in the prompt pass).
__device__ __forceinline__ void load8_u8(const unsigned char* p, cpk* x) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const cpk magic = cpk_make(-8388608.0f, -8388608.0f);
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        x[2 * k] = cpk_add(cpk_make(__uint_as_float(__byte_perm(w[k], 0x4B000000u, 0x7540)),
                                    __uint_as_float(__byte_perm(w[k], 0x4B000000u, 0x7541))), magic);
        x[2 * k + 1] = cpk_add(cpk_make(__uint_as_float(__byte_perm(w[k], 0x4B000000u, 0x7542)),
                                        __uint_as_float(__byte_perm(w[k], 0x4B000000u, 0x7543))), magic);
    }
}
// A[8 h + i] = sum_b R_b x_b[8 u_h + i]  (bytes as integers; the affine map is applied to the sum), rsum = sum_b R_b
template <int NC>
__device__ __forceinline__ void fold_blocks_vec(cpk* A, cf& rsum, const unsigned char* stage, int first, int nblk, int u0, int u1,
                                                const cf* Rm) {
#pragma unroll
    for (int i = 0; i < 8 * NC; ++i) A[i] = cpk_make(0.f, 0.f);
    rsum = cf{0.f, 0.f};
    for (int b = first; b < first + nblk; ++b) {
        const cf R = Rm[b + 1];
        rsum = cadd(rsum, R);
        const unsigned char* pb = stage + (size_t)b * (GR_N * 2);
        cpk x[8], y[8];
        load8_u8(pb + 16 * u0, x);
        if (NC == 2) load8_u8(pb + 16 * u1, y);
#pragma unroll
        for (int i = 0; i < 8; ++i) {                                      // A += x R: x (Rr, Rr) + swap(x) (-Ri, Ri), two FFMA2
            A[i] = cpk_mac(A[i], x[i], R.x, R.y);
            if (NC == 2) A[8 + i] = cpk_mac(A[8 + i], y[i], R.x, R.y);
        }
    }
}
// ---- reference-exact NCO (kExact): one float32 phase argument per sample, as the reference evaluates it -----------------
// exp(-i fl32(PHASE + fl32(w * SEC_TIME[n]))) (gpslib.py:1343-1346).  At 5 kHz x 32 ms the argument reaches 1000 rad, where
// one float32 ulp is 6e-5 rad: the factorised NCO above computes the mathematically exact rotation instead and is that far
// from the reference on every sample (oracle/parity_floor.py: FREQ bit-equal on half the epochs only, complex prompts 4e-4
// apart).  This form reproduces the argument itself (sin / cos of it to 5e-7) and pays N_CYC x 2048 x 2 sin / cos per epoch
// and channel instead of 2 per thread.
// reader's conversion, gpsrecv.py:168-173: complex64(raw) / 127.5 - (1 + 1j).  One FFMA2 here (b * fl32(1 / 127.5) - 1 with
// one rounding; the reference rounds the product first: the two differ by at most one ulp of the sample, 6e-8, on a
// few of the 256 byte values)
__device__ __forceinline__ cpk true_sample_pk(cpk b) {
    return cpk_fma(b, cpk_make(1.0f / 127.5f, 1.0f / 127.5f), cpk_make(-1.f, -1.f));
}
template <int IN_FMT, bool kStage>
__device__ __forceinline__ cf load_true(const void* base, long long n) {
    const cf v = load_raw<IN_FMT, kStage>(base, n);
    if (IN_FMT == GR_IN_U8IQ) return cf{__fsub_rn(__fmul_rn(v.x, 1.0f / 127.5f), 1.0f), __fsub_rn(__fmul_rn(v.y, 1.0f / 127.5f), 1.0f)};
    return v;
}
// F[j] = sum_b x_b[t + 128 j] e(2048 b + t + 128 j)   (natural FFT layout; no affine map, no rotation left to apply)
template <int IN_FMT, bool kStage>
__device__ __forceinline__ void fold_blocks_exact(cf* F, const void* src, int first, int nblk, float w32, float phase32, int t) {
#pragma unroll
    for (int j = 0; j < 16; ++j) F[j] = cf{0.f, 0.f};
    for (int b = first; b < first + nblk; ++b) {
        const int base = b * GR_N + t;
        const float f0 = (float)(base + 1);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const cf x = load_true<IN_FMT, kStage>(src, base + 128 * j);
            const cf e = nco_exact(w32, phase32, f0 + (float)(128 * j));
            F[j].x = fmaf(x.x, e.x, F[j].x); F[j].x = fmaf(-x.y, e.y, F[j].x);
            F[j].y = fmaf(x.x, e.y, F[j].y); F[j].y = fmaf(x.y, e.x, F[j].y);
        }
    }
}
// vector form: A[8 h + i] = sum_b x_b[8 u_h + i] e(2048 b + 8 u_h + i)
template <int NC>
__device__ __forceinline__ void fold_blocks_vec_exact(cpk* A, const unsigned char* stage, int first, int nblk, int u0, int u1,
                                                      float w32, float phase32) {
#pragma unroll
    for (int i = 0; i < 8 * NC; ++i) A[i] = cpk_make(0.f, 0.f);
    for (int b = first; b < first + nblk; ++b) {
        const unsigned char* pb = stage + (size_t)b * (GR_N * 2);
#pragma unroll
        for (int h = 0; h < NC; ++h) {
            const int u = h ? u1 : u0;
            cpk x[8];
            load8_u8(pb + 16 * u, x);
            const float f0 = (float)(b * GR_N + 8 * u + 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const cf e = nco_exact(w32, phase32, f0 + (float)i);
                A[8 * h + i] = cpk_mac(A[8 * h + i], true_sample_pk(x[i]), e.x, e.y);
            }
        }
    }
}
// Exchange "thread owns chunks" -> "thread t owns elements t + 128 j" (the FFT's input layout) through 16 KiB of shared
// memory: chunk u = 64 bytes, its four 16-byte quarters swizzled by (u >> 1) & 3 so that the 128-bit stores of a
// quarter-warp and the 64-bit loads of a half-warp are conflict-free (checked exhaustively on the host, all rotations).
__device__ __forceinline__ void aex_write(float4* ex, int u, const cf* F8) {
    const int sw = (u >> 1) & 3;
#pragma unroll
    for (int c = 0; c < 4; ++c) ex[4 * u + (c ^ sw)] = make_float4(F8[2 * c].x, F8[2 * c].y, F8[2 * c + 1].x, F8[2 * c + 1].y);
}
__device__ __forceinline__ void aex_read(const float4* ex, int t, cf* F) {
    const float2* e2 = reinterpret_cast<const float2*>(ex);
    const int i = t & 7;
    const int off = 8 * (t >> 3) + 2 * ((i >> 1) ^ ((t >> 4) & 3)) + (i & 1);
#pragma unroll
    for (int j = 0; j < 16; ++j) { const float2 v = e2[off + 128 * j]; F[j] = cf{v.x, v.y}; }
}
// r of the first sample of chunk u: the reference's float32 argument fl32(phase + fl32(w * fl32((8 u + 1) / fs)))
__device__ __forceinline__ cf chunk_rot(float w32, float phase32, int u) {
    const float tsec = __fdiv_rn((float)(8 * u + 1), GR_FS);
    return expmi(__fadd_rn(phase32, __fmul_rn(w32, tsec)));
}
// NCO tables of the vector form: sigma_i, R_{k-1}.  Caller syncs.
__device__ __forceinline__ void nco_setup_vec(float w32, int n_cyc, int t, TrackSmem* S) {
    if (t < 8) S->sigma[t] = expmi_d((double)w32 * (double)t * (1.0 / (double)GR_FS));
    if (t >= 32 && t < 32 + n_cyc + 1) {
        const int k = t - 32;
        S->Rm[k] = expmi_d((double)w32 * (double)(k - 1) * 1e-3);
    }
}

// NCO tables of one wipe-off: rho_j, R_{k-1}; returns this thread's r_t.  Caller syncs.
__device__ __forceinline__ cf nco_setup(float w32, float phase32, int n_cyc, int t, TrackSmem* S) {
    if (t < 16) S->rho[t] = expmi_d((double)w32 * (double)(128 * t) * (1.0 / (double)GR_FS));
    if (t >= 32 && t < 32 + n_cyc + 1) {
        const int k = t - 32;
        S->Rm[k] = expmi_d((double)w32 * (double)(k - 1) * 1e-3);
    }
    // a_t = fl32(phase + fl32(w * t_sec)),  t_sec = fl32((t+1)/fs)    (gpslib.py:1053-1054, 1344)
    const float tsec = __fdiv_rn((float)(t + 1), GR_FS);
    const float a = __fadd_rn(phase32, __fmul_rn(w32, tsec));
    return expmi(a);
}

__device__ __forceinline__ float weak_w32(double freq, int weak) {
    // 2*np.pi*freq: python-float product rounded once when FREQ is a python float,
    // float32 product when FREQ is np.float32 (numpy 2 scalar promotion)
    return weak ? (float)(GR_TWO_PI_D * freq) : __fmul_rn(GR_TWO_PI_F, (float)freq);
}

// ---- state helpers (thread 0; c = scalars in shared memory, g = rings in global memory) ------
__device__ __forceinline__ void st_erase_prev(GrChanHot* c) {      // gpslib.py:1095-1099
    c->edge0 = 0;
    c->edge_len = 1;
    c->carry_cnt = 0;
    c->carry_re = 0.0;
    c->carry_im = 0.0;
}
__device__ __forceinline__ void st_unlock(GrChanHot* c, int8_t* cl) {          // gpslib.py:1102-1107
    c->locked = 0;
    c->cl_len = 1; c->cl_head = 0; cl[0] = 0; c->cl_sum = 0; c->cl_sum_last = 0;
    c->ms_time = 0;
    c->phase = 0.f;
    st_erase_prev(c);
}
__device__ __forceinline__ void st_init_sweep(GrChanHot* c, GrChanS* g, int8_t* cl, const gr_track_cfg& cfg) {   // gpslib.py:1110-1116
    st_unlock(c, cl);
    c->freq_save = c->freq;
    c->freq_save_weak = c->freq_weak;
    c->df_save_len = c->df_len;
    for (int i = 0; i < c->df_len; ++i) g->df_save[i] = g->df[(c->df_head + i) % GR_DF_CAP];
    c->freq = (double)cfg.min_freq;
    c->freq_weak = 1;
    c->df_len = 1; c->df_head = 0; g->df[0] = 0.f;
    c->sweep = 1;
}
// gpslib.py:1331-1339 corrQuality: CORRLST.append(cpq); if len > CORRLST_NO: del CORRLST[0]; means of the whole list and
// of its last NO_SEC entries, kept as exact integer running sums.  At n_cyc = 8 the list's capacity equals the ring's
// (60 * 128), so the slot the new entry goes into can be the one that holds the oldest entry: both entries that leave a
// sum are read BEFORE the store.
__device__ __forceinline__ void st_corr_ratios(GrChanHot* c, int no_sec) {          // the two means (FP64 divisions)
    c->corr_q = (double)c->cl_sum / (double)c->cl_len;
    const int nl = c->cl_len < no_sec ? c->cl_len : no_sec;
    c->corr_l = (double)c->cl_sum_last / (double)nl;
}
// The two entries that may leave the sums when the next one is appended: fetched at the start of the epoch (their indices
// are known then), consumed by st_corr_quality several microseconds later.
struct ClOut { int all, win; };
__device__ __forceinline__ ClOut st_corr_prefetch(const GrChanHot* c, const int8_t* cl, int no_sec) {
    const int cap = 60 * no_sec;
    const int len = c->cl_len, head = c->cl_head;
    int iw = head + len - no_sec;
    iw -= iw >= GR_CL_CAP ? GR_CL_CAP : 0;
    ClOut o;
    o.all = (len + 1 > cap) ? (int)cl[head] : 0;
    o.win = (len + 1 > no_sec) ? (int)cl[iw] : 0;
    return o;
}
template <bool kRatios = true>
__device__ __forceinline__ void st_corr_quality(GrChanHot* c, int8_t* cl, ClOut out, double code_phase, int no_sec) {
    const int cap = 60 * no_sec;
    const int v = code_phase < 0.0 ? -1 : 1;
    const int len = c->cl_len, head = c->cl_head;
    const bool full = len + 1 > cap;                       // append, then pop the oldest
    // head, len <= GR_CL_CAP: the ring indices are below 2 GR_CL_CAP, one conditional subtraction instead of a modulo
    int is = head + len, ih = head + 1;
    is -= is >= GR_CL_CAP ? GR_CL_CAP : 0;
    ih -= ih >= GR_CL_CAP ? GR_CL_CAP : 0;
    cl[is] = (int8_t)v;
    c->cl_sum += v - out.all;
    c->cl_sum_last += v - out.win;
    if (full) c->cl_head = ih;
    else c->cl_len = len + 1;
    if (kRatios) st_corr_ratios(c, no_sec);
}

// ---- the kernel ------------------------------------------------------------------------------------
// mbarrier / TMA helpers of the raw-sample stage
__device__ __forceinline__ void trk_mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
        ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// one epoch's raw block (bytes, multiple of 16 KiB... of 4 KiB) global -> shared, completion on `bar`
__device__ __forceinline__ void trk_stage_issue(void* dst, const char* gsrc, unsigned bytes, unsigned long long* bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    for (unsigned o = 0; o < bytes; o += 16384u) {
        const unsigned n = bytes - o < 16384u ? bytes - o : 16384u;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d + o), "l"(gsrc + o), "r"(n), "r"(b) : "memory");
    }
}

// kStage (u8 input only): the epoch's raw block (n_cyc x 4 KiB) is brought into shared memory by TMA while
// the previous epoch's serial tail (prompt means, edge detector, PLL) runs; both sample passes then read
// shared memory instead of L2/HBM -- a single recording is a chain of dependent epochs, so load latency,
// not bandwidth, is what the epoch time is made of.
// NT = 256 ("wide" form, launches of at most one CTA per SM): the two sample passes run on 256 threads with one chunk per
// thread and block; the transforms and everything serial stay on the first 128 threads (named barrier 1).
// kExact: the reference's float32 phase argument for every sample (see nco_exact) instead of the factorised NCO.
template <int IN_FMT, bool kStage, bool kDense = false, int NT = GR_FFT_THREADS, bool kExact = false>
__global__ void __launch_bounds__(NT, kDense ? 3 : 1) track_kernel(const TrackArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf* fftbuf = reinterpret_cast<cf*>(smem_raw);
    float4* part = reinterpret_cast<float4*>(smem_raw);          // aliases the FFT buffers
    TrackSmem* S = reinterpret_cast<TrackSmem*>(smem_raw + a.buf_bytes);
    unsigned char* stage = smem_raw + a.buf_bytes + ((sizeof(TrackSmem) + 15) & ~(size_t)15);
    GrChanHot* C = &S->CH.h;
    constexpr bool kVec = (IN_FMT == GR_IN_U8IQ) && kStage;      // vector form of the sample passes (staged uint8 I/Q)
    constexpr int NC = 256 / NT;                                 // chunks per thread and block in the vector form
    constexpr bool kWide = NT != GR_FFT_THREADS;
    static_assert(NT == 128 || (NT == 256 && kVec), "the wide form exists for the vector passes only");

    const int t = threadIdx.x;
    const int slot = a.slots[blockIdx.x];
    GrChan* Gg = a.state + slot;        // global copy: read once, written back at the end
    GrChanS* G = &S->CH;                // scalars and DF rings are touched every epoch: keep them out of L2 latency
    int8_t* CL = Gg->cl;                // the CORRLST ring stays in global memory (see GrChan)
    const int n_cyc = a.cfg.n_cyc;
    const int ngps = n_cyc * GR_N;
    const int no_sec = 1024 / n_cyc;
    const int ngps_sh = 31 - __clz(ngps);
    const int corr_avg = a.cfg.corr_avg < n_cyc ? a.cfg.corr_avg : n_cyc;
    const int prn = Gg->h.prn;
    const float2* cs = a.tab.conjspec + (size_t)prn * GR_N;

    cf tw1[16], tw2[16];                // FFT twiddles, in registers for the life of the kernel
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float2 u = a.tab.tw1[(t & 127) * 16 + k];
        const float2 v = a.tab.tw2[(t & 7) * 16 + k];
        tw1[k] = cf{u.x, u.y};
        tw2[k] = cf{v.x, v.y};
    }
    const float* code_g = a.tab.code + (size_t)prn * GR_N;       // resampled C/A code of this PRN (8 KB, L1 / L2 resident)
    for (int i = t; i < (int)(sizeof(GrChanS) / 4); i += NT)
        reinterpret_cast<uint32_t*>(G)[i] = reinterpret_cast<const uint32_t*>(Gg)[i];
    if (t == 0) {
        if (kStage) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&S->rawbar)), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();

    const long long rec_off = (long long)C->rec * a.rec_stride;
    const char* rec_base = reinterpret_cast<const char*>(a.samples) + rec_off * (IN_FMT == GR_IN_U8IQ ? 2 : 8);
    const unsigned epoch_bytes = (unsigned)ngps * 2u;
    if (kStage && t == 0) trk_stage_issue(stage, rec_base, epoch_bytes, &S->rawbar);

    for (int e = 0; e < a.n_epochs; ++e) {
        const void* src = kStage ? (const void*)stage
                                 : (const void*)(rec_base + (long long)e * ngps * (IN_FMT == GR_IN_U8IQ ? 2 : 8));
        const long long smp_time = a.smp_time + (long long)e * ngps;
        const long long stream_no = smp_time >> ngps_sh;            // ngps and no_sec are powers of two (n_cyc = 8, 16, 32)
        gr_epoch_out* gO = a.out + ((size_t)e * a.n_active + blockIdx.x);
        gr_epoch_out* O = &S->out[e & 1];       // assembled in shared memory: no global store sits in front of a barrier
        if (t == 0 && a.out_tma) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // record e-2 has left
        const bool report = (stream_no & (long long)(no_sec - 1)) == 0;

        // ---- epoch prologue (gpslib.py:1142-1151) ----
        ClOut cl_out{0, 0};
        if (t == 0) {
            int erased = 0;
            if (stream_no - 1 != C->prev_stream_no) { st_erase_prev(C); erased = 1; }
            C->prev_stream_no = stream_no;
            const int req = C->sweep_req && !C->sweep;
            C->sweep_req = 0;
            if (req) { st_init_sweep(C, G, CL, a.cfg); erased = 1; }
            O->erased = erased;
            S->branch_sweep = C->sweep;
            S->carry0_cnt = C->carry_cnt; S->carry0_re = C->carry_re; S->carry0_im = C->carry_im;
            S->min_edge = C->std_weak ? 3.0 * C->std_dev : (double)__fmul_rn(3.0f, (float)C->std_dev);
            S->locked_in = C->locked;
            S->do_sweep = 0;
            S->dq = C->delay >> 3;
            S->w32 = weak_w32(C->freq, C->freq_weak);
            S->phase32 = C->phase;
            O->edge_mask = 0ull;
            O->n_prompt = 0;
            O->prompt_b1 = 0;
            O->prompt_st0 = 0;
            O->locked_in = C->locked;
            O->report_freq = 0.0;
            O->rep_sweep = 0;
            O->report = report ? 1 : 0;
            O->prn = prn;
            cl_out = st_corr_prefetch(C, CL, no_sec);       // two global loads in flight until the correlation decision
        }
        __syncthreads();
        if (kStage) trk_mbar_wait(&S->rawbar, e & 1);

        if (S->branch_sweep) {
            // =============== sweep branch (gpslib.py:1153-1173, 1350-1380) ===============
            int delay = -1;
            double code_phase = -1.0, z = 0.0;
            double freq = C->freq;                           // a python float during a sweep
            const int avg = a.cfg.sweep_corr_avg;
            int j = 0;
            while (delay < 0 && j < a.cfg.it_sweep) {
                const float w32 = (float)(GR_TWO_PI_D * freq);
                cf rt = cf{0.f, 0.f};
                if (!kExact) rt = nco_setup(w32, 0.f, n_cyc, t, S);
                __syncthreads();
                if (!kWide || t < GR_FFT_THREADS) {
                    cf F[16];
                    if (kExact) fold_blocks_exact<IN_FMT, kStage>(F, src, 0, avg, w32, 0.f, t);   // getCorrMax: phase = 0
                    else fold_blocks<IN_FMT, kStage>(F, src, 0, avg, rt, t, S);
                    corr_and_stats<kDense, kWide>(F, cs, 1.0f / ((float)avg * (float)GR_N), fftbuf, tw1, tw2, t, S);
                }
                __syncthreads();
                z = S->z;
                if (z > (double)a.cfg.corr_min) {
                    delay = S->sh_i[4];
                    code_phase = fit_code_phase(delay, (double)S->c3[0], (double)S->c3[1], (double)S->c3[2]);
                } else {
                    freq = freq + (double)a.cfg.step_freq;
                }
                ++j;
                __syncthreads();                              // S->z / rho / Rm are rewritten next round
            }
            if (kStage && t == 0 && e + 1 < a.n_epochs)      // all reads of this epoch's block are behind a barrier
                trk_stage_issue(stage, rec_base + (long long)(e + 1) * epoch_bytes, epoch_bytes, &S->rawbar);
            int running = 1;
            if (delay >= 0) running = 0;
            else if (freq > (double)a.cfg.max_freq) { freq = (double)a.cfg.min_freq; running = 0; }
            if (t == 0) {
                C->rep_sweep = 1;
                C->sweep = running;
                C->freq = freq;
                C->freq_weak = 1;
                C->max_corr = z;
                st_corr_quality(C, CL, cl_out, code_phase, no_sec);
                if (delay >= 0) C->delay = delay;
                else if (!running) {                          // restoreFreq, gpslib.py:1118-1120
                    C->freq = C->freq_save;
                    C->freq_weak = C->freq_save_weak;
                    C->df_len = C->df_save_len;
                    C->df_head = 0;
                    for (int i = 0; i < C->df_len; ++i) G->df[i] = G->df_save[i];
                }
                O->tracked = 0;
                for (int k = 0; k < 2 * GR_MAX_PROMPT; ++k) O->prompt[k] = 0.f;
                O->corr_delay = delay;
                O->code_phase = code_phase;
                if (report) { O->rep_sweep = C->rep_sweep; O->report_freq = C->freq; C->rep_sweep = 0; }
                O->corr3[0] = S->c3[0]; O->corr3[1] = S->c3[1]; O->corr3[2] = S->c3[2];
                O->corr_mean = (float)S->cmean;
                O->corr_std = (float)S->cstd;
            }
        } else {
            // =============== tracking branch (gpslib.py:1175-1208) ===============
            const float w32 = S->w32, phase32 = S->phase32;
            cf rt = cf{0.f, 0.f};
            int dq = 0, u0 = 0, u1 = 0;
            cf E0 = cf{0.f, 0.f}, E1 = cf{0.f, 0.f};
            const int d_spec = C->delay;                     // DELAY going into the epoch (thread 0 updates it behind the next barriers)
            if constexpr (kVec) {
                if (!kExact) nco_setup_vec(w32, n_cyc, t, S);
                dq = S->dq;
                u0 = (t + dq) & 255;
                u1 = u0 ^ 128;
                if (!kExact) {
                    E0 = chunk_rot(w32, phase32, u0);
                    if (NC == 2) E1 = chunk_rot(w32, phase32, u1);
                }
                __syncthreads();
                float4* ex = reinterpret_cast<float4*>(kDense ? fftbuf : fftbuf + GR_B1_ELEMS);   // free at this point (see fft2048)
                {
                    cpk A[8 * NC];
                    cf G[8 * NC];
                    if constexpr (kExact) {
                        // ONE pass over the epoch's samples: every sample is rotated once, y = x e(n), and goes into the coherent
                        // fold (blocks first .. first + corr_avg - 1) AND into the prompt sums, the latter formed for the DELAY the
                        // epoch started with.  The correlation below changes DELAY on a few per cent of the epochs only; then
                        // the prompt sums are formed again for the new DELAY (the two-pass code further down).
                        const int first = (n_cyc - corr_avg) / 2;
                        const int r8s = d_spec & 7;
                        const bool wr0 = t + dq >= 256, wr1 = t + dq + 128 >= 256;
                        float cc[8 * NC];
#pragma unroll
                        for (int h = 0; h < NC; ++h) {
                            const int uh = h ? u1 : u0;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                cc[8 * h + i] = __ldg(code_g + ((8 * uh + i - d_spec) & (GR_N - 1)));
                                if (h == 0 && t == 0) S->qb[i] = cf{i < r8s ? cc[i] : 0.f, 0.f};
                            }
                        }
                        float2* part2 = reinterpret_cast<float2*>(smem_raw);       // row k = pass k, this thread's column only
                        for (int k = 0; k <= n_cyc; ++k) part2[k * NT + t] = make_float2(0.f, 0.f);
#pragma unroll
                        for (int i = 0; i < 8 * NC; ++i) A[i] = cpk_make(0.f, 0.f);
                        auto block = [&](int b, auto fold_tag) {           // one 1-ms block; fold_tag: it belongs to the coherent fold
                            constexpr bool kFold = decltype(fold_tag)::value;
                            const unsigned char* pb = stage + (size_t)b * (GR_N * 2);
#pragma unroll
                            for (int h = 0; h < NC; ++h) {
                                const int uh = h ? u1 : u0;
                                cpk x[8];
                                load8_u8(pb + 16 * uh, x);
                                const float f0 = (float)(b * GR_N + 8 * uh + 1);
                                const cpk fn0 = cpk_make(f0, f0 + 1.0f);
                                cpk s0 = cpk_make(0.f, 0.f), s1 = s0;
#pragma unroll
                                for (int i = 0; i < 8; i += 2) {
                                    cf e[2];
                                    nco_exact2(w32, phase32, cpk_add(fn0, cpk_bc((float)i)), e[0], e[1]);   // samples i, i + 1
#pragma unroll
                                    for (int o = 0; o < 2; ++o) {
                                        const cpk y = cpk_cmul(true_sample_pk(x[i + o]), e[o].x, e[o].y);
                                        if (kFold) A[8 * h + i + o] = cpk_add(A[8 * h + i + o], y);
                                        const cpk c2 = cpk_make(cc[8 * h + i + o], cc[8 * h + i + o]);
                                        if (o) s1 = cpk_fma(y, c2, s1); else s0 = cpk_fma(y, c2, s0);
                                    }
                                }
                                float re, im;
                                cpk_split(cpk_add(s0, s1), re, im);
                                const int k = b + 1 - ((h ? wr1 : wr0) ? 1 : 0);
                                float2 acc = part2[k * NT + t];
                                acc.x += re; acc.y += im;
                                part2[k * NT + t] = acc;
                            }
                        };
                        for (int b = 0; b < first; ++b) block(b, std::false_type{});
                        for (int b = first; b < first + corr_avg; ++b) block(b, std::true_type{});
                        for (int b = first + corr_avg; b < n_cyc; ++b) block(b, std::false_type{});
                        __syncthreads();
                        if (t <= n_cyc) {                                     // B_k: the d & 7 samples in front of the boundary, pass k = t
                            cf bsum = cf{0.f, 0.f};
                            if (t >= 1) {
                                cpk x[8];
                                load8_u8(stage + (size_t)(t - 1) * (GR_N * 2) + 16 * dq, x);
                                const float f0 = (float)((t - 1) * GR_N + 8 * dq + 1);
#pragma unroll
                                for (int i = 0; i < 7; ++i) {
                                    const cf e = nco_exact(w32, phase32, f0 + (float)i);
                                    float yr, yi;
                                    cpk_split(cpk_cmul(true_sample_pk(x[i]), e.x, e.y), yr, yi);
                                    const float c = S->qb[i].x;
                                    bsum.x = fmaf(yr, c, bsum.x);
                                    bsum.y = fmaf(yi, c, bsum.y);
                                }
                            }
                            S->red[t].z = bsum.x;
                            S->red[t].w = bsum.y;
                        }
                        {   // reduce the rows: warp w takes rows w, w + NT / 32, ...
                            const int w = t >> 5, l = t & 31;
                            for (int r = w; r <= n_cyc; r += NT / 32) {
                                float2 v = part2[r * NT + l];
#pragma unroll
                                for (int m = 1; m < NT / 32; ++m) {
                                    const float2 u = part2[r * NT + l + 32 * m];
                                    v.x += u.x; v.y += u.y;
                                }
                                v.x = warp_sum(v.x); v.y = warp_sum(v.y);
                                if (l == 0) { S->red[r].x = v.x; S->red[r].y = v.y; }
                            }
                        }
                        __syncthreads();                                      // the rows are read: the buffer is free for the exchange
#pragma unroll
                        for (int i = 0; i < 8 * NC; ++i) cpk_split(A[i], G[i].x, G[i].y);
                    } else {
                        cf rsum;
                        fold_blocks_vec<NC>(A, rsum, stage, (n_cyc - corr_avg) / 2, corr_avg, u0, u1, S->Rm);
#pragma unroll
                        for (int i = 0; i < 8 * NC; ++i) {
                            float ar, ai;
                            cpk_split(A[i], ar, ai);
                            const cf r = cmul(i < 8 ? E0 : E1, S->sigma[i & 7]);
                            G[i] = cmul(affine_sum<IN_FMT>(cf{ar, ai}, rsum), r);
                        }
                    }
                    aex_write(ex, u0, G);
                    if (NC == 2) aex_write(ex, u1, G + 8 * (NC - 1));
                }
                __syncthreads();
                if (!kWide || t < GR_FFT_THREADS) {
                    cf F[16];
                    aex_read(ex, t, F);
                    corr_and_stats<kDense, kWide>(F, cs, 1.0f / ((float)corr_avg * (float)GR_N), fftbuf, tw1, tw2, t, S);
                }
            } else {
                if (!kExact) rt = nco_setup(w32, phase32, n_cyc, t, S);
                __syncthreads();
                cf F[16];
                if (kExact) fold_blocks_exact<IN_FMT, kStage>(F, src, (n_cyc - corr_avg) / 2, corr_avg, w32, phase32, t);
                else fold_blocks<IN_FMT, kStage>(F, src, (n_cyc - corr_avg) / 2, corr_avg, rt, t, S);
                corr_and_stats<kDense>(F, cs, 1.0f / ((float)corr_avg * (float)GR_N), fftbuf, tw1, tw2, t, S);
            }
            if (t == 0) {                                    // same thread as the statistics above: no barrier in between
                int delay = -1;
                double code_phase = -1.0;
                if (S->z > (double)a.cfg.corr_min) {
                    delay = S->sh_i[4];
                    code_phase = fit_code_phase(delay, (double)S->c3[0], (double)S->c3[1], (double)S->c3[2]);
                }
                st_corr_quality<false>(C, CL, cl_out, code_phase, no_sec);      // integer sums now, the two FP64 means by thread 96 later
                if (delay >= 0) C->delay = delay;
                S->corr_delay = delay;
                S->code_phase = code_phase;
                S->delay = C->delay;
                const int dd = C->delay;
                S->n_prompt = ((C->carry_cnt + dd > 0) ? 1 : 0) + (n_cyc - 1) + (dd == 0 ? 1 : 0);
            }
            __syncthreads();

            // ---- prompt integrate & dump (decodeData) ----
            const int d = S->delay;
            if constexpr (kVec) {
              if (!kExact || d != d_spec) {                   // exact form: the fused pass above already holds the sums for d_spec
                // Vector form: pass k sums the 2048 samples from 8 (d >> 3) + 2048 (k - 1) on, thread t its chunks u0, u1
                // (a chunk whose index ran past 255 lies in the next block: "wrapped").  The code-period boundary d sits
                // inside thread 0's first chunk: the d & 7 samples in front of it (B_k) are summed once more by thread k
                // after the passes, and the true 1-ms sums are S_k - B_k + B_{k+1} as in the scalar form.
                const int r8 = d & 7;
                if ((d >> 3) != dq) {                                   // DELAY moved to another chunk (rare; uniform)
                    dq = d >> 3;
                    u0 = (t + dq) & 255;
                    u1 = u0 ^ 128;
                    if (!kExact) {
                        E0 = chunk_rot(w32, phase32, u0);
                        if (NC == 2) E1 = chunk_rot(w32, phase32, u1);
                    }
                }
                const bool wr0 = t + dq >= 256, wr1 = t + dq + 128 >= 256;
                float qr[8 * NC], qi[8 * NC];                           // fast form: rotation x code; exact form: qr = code only
                if constexpr (kExact) {
#pragma unroll
                    for (int h = 0; h < NC; ++h) {
                        const int uh = h ? u1 : u0;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float c = __ldg(code_g + ((8 * uh + i - d) & (GR_N - 1)));
                            qr[8 * h + i] = c;
                            qi[8 * h + i] = 0.f;
                            if (h == 0 && t == 0) S->qb[i] = cf{i < r8 ? c : 0.f, 0.f};
                        }
                    }
                } else {
                    cf qall = cf{0.f, 0.f}, qw = cf{0.f, 0.f}, qbs = cf{0.f, 0.f};
                    const cf R1 = S->Rm[2];
#pragma unroll
                    for (int h = 0; h < NC; ++h) {
                        const int uh = h ? u1 : u0;
                        const bool wr = h ? wr1 : wr0;
                        cf Eh = h ? E1 : E0;
                        if (wr) Eh = cmul(Eh, R1);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float c = __ldg(code_g + ((8 * uh + i - d) & (GR_N - 1)));
                            const cf r = cmul(Eh, S->sigma[i]);
                            const cf q = cf{r.x * c, r.y * c};
                            qr[8 * h + i] = q.x;
                            qi[8 * h + i] = q.y;
                            qall = cadd(qall, q);
                            if (wr) qw = cadd(qw, q);
                            if (h == 0 && t == 0) {                     // thread 0's first chunk starts at 8 dq <= d: never wrapped
                                const cf qm = i < r8 ? q : cf{0.f, 0.f};
                                S->qb[i] = qm;
                                qbs = cadd(qbs, qm);
                            }
                        }
                    }
                    float v[6] = {qall.x, qall.y, qbs.x, qbs.y, qw.x, qw.y};
#pragma unroll
                    for (int m = 0; m < 6; ++m) {
                        v[m] = warp_sum(v[m]);
                        if ((t & 31) == 0) S->qred[t >> 5][m] = v[m];
                    }
                }
                float2* part2 = reinterpret_cast<float2*>(smem_raw);
                const int rows_cap = a.buf_bytes / (NT * 8);
                const int part_rows = rows_cap < n_cyc + 1 ? rows_cap : n_cyc + 1;
                const unsigned char* p0 = stage + 16 * u0;
                const unsigned char* p1 = stage + 16 * u1;
                for (int k0 = 0; k0 <= n_cyc; k0 += part_rows) {
                    const int k1 = (k0 + part_rows <= n_cyc + 1) ? k0 + part_rows : n_cyc + 1;
#pragma unroll 3
                    for (int k = k0; k < k1; ++k) {
                        const int b0 = k - 1 + (wr0 ? 1 : 0), b1 = k - 1 + (wr1 ? 1 : 0);
                        const bool v0 = (unsigned)b0 < (unsigned)n_cyc, v1 = (unsigned)b1 < (unsigned)n_cyc;
                        cpk x[8], y[8];
                        load8_u8(p0 + (size_t)(v0 ? b0 : 0) * (GR_N * 2), x);
                        if (NC == 2) load8_u8(p1 + (size_t)(v1 ? b1 : 0) * (GR_N * 2), y);
                        // sum x q per chunk: two FFMA2 per sample, two independent chains per chunk (even / odd samples)
                        cpk a0 = cpk_make(0.f, 0.f), a1 = a0, b0e = a0, b1e = a0;
                        if constexpr (kExact) {
                            // q = e(n) c: the reference's factor of this very sample (block b0 / b1 of the epoch)
                            const float f0 = (float)((v0 ? b0 : 0) * GR_N + 8 * u0 + 1);
                            const float f1 = (float)((v1 ? b1 : 0) * GR_N + 8 * u1 + 1);
#pragma unroll
                            for (int i = 0; i < 8; i += 2) {
                                cf e0, e1;
                                nco_exact2(w32, phase32, cpk_add(cpk_make(f0, f0 + 1.0f), cpk_bc((float)i)), e0, e1);
                                a0 = cpk_mac(a0, true_sample_pk(x[i]), e0.x * qr[i], e0.y * qr[i]);
                                b0e = cpk_mac(b0e, true_sample_pk(x[i + 1]), e1.x * qr[i + 1], e1.y * qr[i + 1]);
                                if (NC == 2) {
                                    cf g0, g1;
                                    nco_exact2(w32, phase32, cpk_add(cpk_make(f1, f1 + 1.0f), cpk_bc((float)i)), g0, g1);
                                    a1 = cpk_mac(a1, true_sample_pk(y[i]), g0.x * qr[8 * (NC - 1) + i], g0.y * qr[8 * (NC - 1) + i]);
                                    b1e = cpk_mac(b1e, true_sample_pk(y[i + 1]), g1.x * qr[8 * (NC - 1) + i + 1], g1.y * qr[8 * (NC - 1) + i + 1]);
                                }
                            }
                        } else {
#pragma unroll
                        for (int i = 0; i < 8; i += 2) {
                            a0 = cpk_mac(a0, x[i], qr[i], qi[i]);
                            b0e = cpk_mac(b0e, x[i + 1], qr[i + 1], qi[i + 1]);
                            if (NC == 2) {
                                a1 = cpk_mac(a1, y[i], qr[8 * (NC - 1) + i], qi[8 * (NC - 1) + i]);
                                b1e = cpk_mac(b1e, y[i + 1], qr[8 * (NC - 1) + i + 1], qi[8 * (NC - 1) + i + 1]);
                            }
                        }
                        }
                        a0 = cpk_add(a0, b0e);
                        a1 = cpk_add(a1, b1e);
                        float r0, i0, r1, i1;
                        cpk_split(a0, r0, i0);
                        cpk_split(a1, r1, i1);
                        const float re = (v0 ? r0 : 0.f) + (NC == 2 && v1 ? r1 : 0.f);
                        const float im = (v0 ? i0 : 0.f) + (NC == 2 && v1 ? i1 : 0.f);
                        part2[(k - k0) * NT + t] = make_float2(re, im);
                    }
                    __syncthreads();
                    if (k0 == 0 && t <= n_cyc) {                          // B_k: the d & 7 samples in front of the boundary, pass k = t
                        cf bsum = cf{0.f, 0.f};
                        if (t >= 1) {
                            cpk x[8];
                            load8_u8(stage + (size_t)(t - 1) * (GR_N * 2) + 16 * dq, x);
                            const float f0 = (float)((t - 1) * GR_N + 8 * dq + 1);
#pragma unroll
                            for (int i = 0; i < 7; ++i) {
                                float xr, xi;
                                cpk_split(kExact ? true_sample_pk(x[i]) : x[i], xr, xi);
                                cf qm = S->qb[i];
                                if (kExact) { const cf e = nco_exact(w32, phase32, f0 + (float)i); qm = cf{e.x * qm.x, e.y * qm.x}; }
                                bsum.x = fmaf(xr, qm.x, bsum.x); bsum.x = fmaf(-xi, qm.y, bsum.x);
                                bsum.y = fmaf(xr, qm.y, bsum.y); bsum.y = fmaf(xi, qm.x, bsum.y);
                            }
                        }
                        S->red[t].z = bsum.x;
                        S->red[t].w = bsum.y;
                    }
                    if (!kExact && k0 == 0 && t >= 96 && t < 102) {
                        float qv = (S->qred[0][t - 96] + S->qred[1][t - 96]) + (S->qred[2][t - 96] + S->qred[3][t - 96]);
                        if (kWide) qv += (S->qred[4][t - 96] + S->qred[5][t - 96]) + (S->qred[6][t - 96] + S->qred[7][t - 96]);
                        S->qsum[t - 96] = qv;
                    }
                    {   // reduce the staged rows: warp w takes rows w, w + NT / 32, ...
                        const int w = t >> 5, l = t & 31;
                        for (int r = w; r < k1 - k0; r += NT / 32) {
                            float2 v = part2[r * NT + l];
#pragma unroll
                            for (int m = 1; m < NT / 32; ++m) {
                                const float2 u = part2[r * NT + l + 32 * m];
                                v.x += u.x; v.y += u.y;
                            }
                            v.x = warp_sum(v.x); v.y = warp_sum(v.y);
                            if (l == 0) { S->red[k0 + r].x = v.x; S->red[k0 + r].y = v.y; }
                        }
                    }
                    __syncthreads();
                }
              }
            } else {
                const int jb = d >> 7, dlow = d & 127;
                const bool inB = t < dlow;                       // row 0, before the code-period boundary
                cf q[16];
                float cj[kExact ? 16 : 1];                       // exact form: the code values of this thread's 16 rows
                if constexpr (kExact) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) cj[kExact ? j : 0] = __ldg(code_g + ((t + 128 * ((j + jb) & 15) - d) & (GR_N - 1)));
                } else {
                    cf qall = cf{0.f, 0.f}, qw = cf{0.f, 0.f};
                    const cf R1 = S->Rm[2];
    #pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int row = (j + jb) & 15;
                        const bool wrapped = (j + jb) >= 16;
                        const int i = t + 128 * row;
                        const float c = __ldg(code_g + ((i - d) & (GR_N - 1)));
                        cf r = cmul(rt, S->rho[row]);
                        if (wrapped) r = cmul(r, R1);
                        q[j] = cf{r.x * c, r.y * c};
                        qall = cadd(qall, q[j]);
                        if (wrapped) qw = cadd(qw, q[j]);
                    }
                    float v[6] = {qall.x, qall.y, inB ? q[0].x : 0.f, inB ? q[0].y : 0.f, qw.x, qw.y};
    #pragma unroll
                    for (int m = 0; m < 6; ++m) {
                        v[m] = warp_sum(v[m]);
                        if ((t & 31) == 0) S->qred[t >> 5][m] = v[m];
                    }
                }
                for (int k0 = 0; k0 <= n_cyc; k0 += a.part_rows) {
                    const int k1 = (k0 + a.part_rows <= n_cyc + 1) ? k0 + a.part_rows : n_cyc + 1;
                    for (int k = k0; k < k1; ++k) {
                        const long long base = (long long)128 * jb + (long long)GR_N * (k - 1) + t;
                        // rows outside the epoch (pass 0: not yet wrapped, last pass: wrapped) read a clamped
                        // address and are zeroed with selects: no branches around the loads
                        const int vmode = (k == 0) ? 1 : (k == n_cyc ? 2 : 0);
                        cf x[16];
    #pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const bool wrapped = (j + jb) >= 16;
                            const bool valid = vmode == 0 || (vmode == 1 ? wrapped : !wrapped);
                            const long long n = valid ? base + 128 * j : (long long)t;
                            const cf v = kExact ? load_true<IN_FMT, kStage>(src, n) : load_raw<IN_FMT, kStage>(src, n);
                            x[j].x = valid ? v.x : 0.f;
                            x[j].y = valid ? v.y : 0.f;
                            if constexpr (kExact) {                 // q = e(n) c: the reference's factor of this very sample
                                const cf e = nco_exact(w32, phase32, (float)((int)n + 1));
                                q[j] = cf{e.x * cj[kExact ? j : 0], e.y * cj[kExact ? j : 0]};
                            }
                        }
                        cf p0 = cmul(x[0], q[0]);
                        cf acc = p0;
    #pragma unroll
                        for (int j = 1; j < 16; ++j) {
                            acc.x = fmaf(x[j].x, q[j].x, acc.x);
                            acc.x = fmaf(-x[j].y, q[j].y, acc.x);
                            acc.y = fmaf(x[j].x, q[j].y, acc.y);
                            acc.y = fmaf(x[j].y, q[j].x, acc.y);
                        }
                        if (!inB) p0 = cf{0.f, 0.f};
                        part[(k - k0) * 128 + t] = make_float4(acc.x, acc.y, p0.x, p0.y);
                    }
                    __syncthreads();
                    if (!kExact && k0 == 0 && t >= 96 && t < 102)
                        S->qsum[t - 96] = (S->qred[0][t - 96] + S->qred[1][t - 96]) + (S->qred[2][t - 96] + S->qred[3][t - 96]);
                    {   // reduce the staged rows: warp w takes rows w, w+4, ...
                        const int w = t >> 5, l = t & 31;
                        for (int r = w; r < k1 - k0; r += 4) {
                            float4 v = part[r * 128 + l];
    #pragma unroll
                            for (int m = 1; m < 4; ++m) {
                                const float4 u = part[r * 128 + l + 32 * m];
                                v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
                            }
                            v.x = warp_sum(v.x); v.y = warp_sum(v.y); v.z = warp_sum(v.z); v.w = warp_sum(v.w);
                            if (l == 0) S->red[k0 + r] = v;
                        }
                    }
                    __syncthreads();
                }
            }
            if (kStage && t == 0 && e + 1 < a.n_epochs)      // the prompt pass was the last reader of this epoch's block
                trk_stage_issue(stage, rec_base + (long long)(e + 1) * epoch_bytes, epoch_bytes, &S->rawbar);
            // ---- everything behind the prompt sums is split by TASK over the four warps, warp-level synchronisation only:
            //      warp 0 the carrier loop (the only part the next epoch waits for), warp 1 the edge detector, warp 2 the
            //      amplitude statistics and the prompt values of the record, warp 3 the remaining record fields.  Warps 0-2
            //      each form the 1-ms prompt means themselves (lane k <-> pass k, and k + 32) -- same arithmetic, same bits --
            //      so that no block barrier sits between the sums and their three consumers.  State is partitioned by
            //      warp; what one warp needs from another's previous value was snapshot in the prologue.
            const int np = S->n_prompt;
            const int wid = t >> 5, ln = t & 31;
            const unsigned FULL = 0xffffffffu;
            const int nps0 = S->carry0_cnt;
            const bool first_present = nps0 + d > 0;
            const int fp = first_present ? 1 : 0;
            double mr[2] = {0.0, 0.0}, mi[2] = {0.0, 0.0};              // this lane's prompts: index ln and ln + 32
            if (wid < 3) {
                float4* xsw = S->wxs[wid];
                double* prr = S->pr_re[wid];
                double* pri = S->pr_im[wid];
                // true-sample sums X_k, XB_k: affine map per sum, block rotation R_{k-1}
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = ln + 32 * h;
                    if (k <= n_cyc) {
                        const float4 v = S->red[k];
                        const cf Qk = (k == 0) ? cf{S->qsum[4], S->qsum[5]}
                                               : (k == n_cyc ? cf{S->qsum[0] - S->qsum[4], S->qsum[1] - S->qsum[5]} : cf{S->qsum[0], S->qsum[1]});
                        cf X = affine_sum<IN_FMT>(cf{v.x, v.y}, Qk);
                        cf XB = (k == 0) ? cf{0.f, 0.f} : affine_sum<IN_FMT>(cf{v.z, v.w}, cf{S->qsum[2], S->qsum[3]});
                        if constexpr (kExact) {                     // the sums are already those of the rotated true samples
                            X = cf{v.x, v.y};
                            XB = (k == 0) ? cf{0.f, 0.f} : cf{v.z, v.w};
                        } else {
                        const cf R = S->Rm[k];
                        X = cmul(X, R);
                        XB = cmul(XB, R);
                        }
                        xsw[k] = make_float4(X.x, X.y, XB.x, XB.y);
                    }
                }
                __syncwarp();
                // seg_m = X_m - XB_m + XB_{m+1} (m < n_cyc), tail = X_n - XB_n; prompt index = (first segment present ? 1 : 0) + m - 1
                // 1 / (number of samples of the first, partial segment): float reciprocal + two Newton steps in FP64 (the FP64
                // division is ~150 cycles; the quotient differs from the correctly rounded one by at most 1 ulp of a double)
                const double cnt = (double)(nps0 + d > 0 ? nps0 + d : 1);
                double rc = (double)__frcp_rn((float)cnt);
                rc = rc * (2.0 - cnt * rc);
                rc = rc * (2.0 - cnt * rc);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int m = ln + 32 * h;
                    if (m <= n_cyc) {
                        const float4 v = xsw[m];
                        double sre = (double)v.x - (double)v.z, sim = (double)v.y - (double)v.w;
                        if (m < n_cyc) { const float4 u = xsw[m + 1]; sre += (double)u.z; sim += (double)u.w; }
                        if (m == 0) {
                            if (first_present) { prr[0] = (S->carry0_re + sre) * rc; pri[0] = (S->carry0_im + sim) * rc; }
                        } else if (m < n_cyc || d == 0) {
                            prr[fp + m - 1] = sre * (1.0 / GR_N);                     // exact: a power of two
                            pri[fp + m - 1] = sim * (1.0 / GR_N);
                        } else if (wid == 0) {                                        // carry-over of the partial code period (gpslib.py:1440-1441)
                            C->carry_cnt = GR_N - d; C->carry_re = sre; C->carry_im = sim;
                        }
                        if (wid == 0 && m == n_cyc && d == 0) { C->carry_cnt = 0; C->carry_re = 0.0; C->carry_im = 0.0; }
                    }
                }
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = ln + 32 * h;
                    if (k < np) { mr[h] = prr[k]; mi[h] = pri[k]; }
                }
            }
            if (wid == 1) {
                // ---- edge detector (gpslib.py:1417-1436), threshold from the PREVIOUS epoch's STD_DEV.  The per-prompt
                //      comparisons run one per lane; what is sequential (the sign of the last accepted edge) is a loop over
                //      ballot masks in registers ----
                unsigned long long emask = 0ull;
                if (S->locked_in) {
                    const double* prr = S->pr_re[1];
                    const double min_edge = S->min_edge;
                    const double prev_signal0 = C->prev_signal;
                    unsigned long long pos = 0ull, neg = 0ull, ppos = 0ull, pneg = 0ull, big = 0ull;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int k = ln + 32 * h;
                        const bool valid = k < np;
                        const double prev = (k == 0) ? prev_signal0 : (valid ? prr[k - 1] : 0.0);
                        pos |= (unsigned long long)__ballot_sync(FULL, valid && mr[h] > 0.0) << (32 * h);
                        neg |= (unsigned long long)__ballot_sync(FULL, valid && mr[h] < 0.0) << (32 * h);
                        ppos |= (unsigned long long)__ballot_sync(FULL, valid && prev > 0.0) << (32 * h);
                        pneg |= (unsigned long long)__ballot_sync(FULL, valid && prev < 0.0) << (32 * h);
                        big |= (unsigned long long)__ballot_sync(FULL, valid && fabs(mr[h] - prev) > min_edge) << (32 * h);
                    }
                    int edge0 = C->edge0, elen = C->edge_len;
                    int prev_sign = (2 * (elen & 1) - 1) * edge0;
                    const unsigned long long all = np >= 64 ? ~0ull : ((1ull << np) - 1ull);
                    if (edge0 != 0 && (pos | neg) == all && (ppos | pneg) == all && np <= 32) {
                        // the usual case: every sign is +-1.  One bit of state (the sign of the last accepted edge), masks shifted
                        // down one position per prompt
                        unsigned s32 = (unsigned)pos, p32 = (unsigned)ppos, b32 = (unsigned)big, ps = prev_sign > 0 ? 1u : 0u, em = 0u;
                        for (int k = 0; k < np; ++k) {
                            const unsigned sb = s32 & 1u, edge = (sb ^ ps) & ~(ps ^ (p32 & 1u)) & b32 & 1u;
                            em |= edge << k;
                            ps ^= edge;                                   // an accepted edge flips the sign
                            s32 >>= 1; p32 >>= 1; b32 >>= 1;
                        }
                        emask = em;
                        elen += __popc(em);
                    } else {
                        for (int k = 0; k < np; ++k) {
                            const int sgn = (int)((pos >> k) & 1ull) - (int)((neg >> k) & 1ull);
                            if (edge0 == 0) {
                                edge0 = sgn;
                                prev_sign = sgn;
                            } else {
                                const bool same = prev_sign > 0 ? ((ppos >> k) & 1ull) != 0 : (prev_sign < 0 ? ((pneg >> k) & 1ull) != 0 : false);
                                if (sgn != prev_sign && same && ((big >> k) & 1ull)) {
                                    emask |= 1ull << k;
                                    ++elen;
                                    prev_sign = sgn;
                                }
                            }
                        }
                    }
                    if (report && elen > 2) {                        // evalEdges -> logicalBits, gpslib.py:1465-1487 (runs while still PHASE_LOCKED as of this epoch's start)
                        if ((elen - 2) & 1) edge0 = -edge0;
                        elen = 2;
                    }
                    if (ln == 0) {
                        C->edge0 = edge0;
                        C->edge_len = elen;
                        if (np > 0) C->prev_signal = prr[np - 1];
                        C->ms_time += np;
                    }
                }
                if (ln == 0) {
                    int n1 = nps0 + d;
                    long long st;
                    if (n1 == 0) { n1 = GR_N; st = smp_time; } else { st = smp_time + d - GR_N; }
                    O->edge_mask = emask;
                    O->prompt_b1 = n1;
                    O->prompt_st0 = st;
                    O->edge0 = C->edge0;
                    O->edge_len = C->edge_len;
                    O->ms_time = C->ms_time;
                }
            } else if (wid == 2) {
                // ---- amplitude statistics (gpslib.py:1186-1188) and the prompt values of the record ----
                float ab[2];
                float sum_ab = 0.f;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = ln + 32 * h;
                    ab[h] = 0.f;
                    if (k < np) {
                        const cf g = cf{(float)mr[h], (float)mi[h]};                   // gpsData = np.asarray(..., complex64)
                        O->prompt[2 * k] = g.x;
                        O->prompt[2 * k + 1] = g.y;
                        ab[h] = hypotf(g.x, g.y);
                        sum_ab += ab[h];
                    } else if (k < GR_MAX_PROMPT) {
                        O->prompt[2 * k] = 0.f;
                        O->prompt[2 * k + 1] = 0.f;
                    }
                }
                if (ln < 2 && ln + 32 >= np) { O->prompt[2 * (ln + 32)] = 0.f; O->prompt[2 * (ln + 32) + 1] = 0.f; }
                sum_ab = warp_sum(sum_ab);
                const float fn = (float)np;
                const float mean_ab = sum_ab / fn;
                float sq = 0.f;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = ln + 32 * h;
                    if (k < np) { const float dv = ab[h] - mean_ab; sq = fmaf(dv, dv, sq); }
                }
                sq = warp_sum(sq);
                if (ln == 0) {
                    const float std32 = sqrtf(sq / fn);
                    const float amp = mean_ab / std32;
                    C->std_dev = (double)std32;
                    C->std_weak = 0;
                    C->amplitude = amp;
                    O->amplitude = amp;
                    O->std_dev = std32;
                }
            } else if (wid == 3) {
                if (ln == 0) {                                       // CORR_Q, CORR_L: two FP64 divisions, off the critical path
                    st_corr_ratios(C, no_sec);
                    O->corr_q = C->corr_q;
                    O->corr_l = C->corr_l;
                } else if (ln == 1) {
                    O->tracked = 1;
                    O->corr_delay = S->corr_delay;
                    O->code_phase = S->code_phase;
                    O->corr3[0] = S->c3[0]; O->corr3[1] = S->c3[1]; O->corr3[2] = S->c3[2];
                    O->corr_mean = (float)S->cmean;
                    O->corr_std = (float)S->cstd;
                    O->n_prompt = np;
                    O->delay = d;
                    O->n_prev = d == 0 ? 0 : GR_N - d;
                    C->max_corr = S->z;
                    O->max_corr = S->z;
                    O->reserved[0] = 0; O->reserved[1] = 0;
                }
            } else if (wid == 0) {
                // ---- carrier loop: phaseLockedLoop, gpslib.py:1215-1262 ----
                float ph[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = ln + 32 * h;
                    ph[h] = 0.f;
                    if (k < np) {
                        ph[h] = atanf(__fdiv_rn((float)mi[h], (float)mr[h]));
                        S->ph[k] = ph[h];
                    }
                }
                // mean of the DF FIFO (its loads issue before the scan below)
                float dsum = 0.f;
                const int dfl = C->df_len, dfh = C->df_head;
                for (int i = ln; i < dfl; i += 32) dsum += G->df[(dfh + i) % GR_DF_CAP];
                __syncwarp();
                // unwrap: dp -= sign(delta) whenever |delta| > 2; realPhase[i] += dp*pi
                float inc[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = ln + 32 * h;
                    inc[h] = 0.f;
                    if (k >= 1 && k < np) {
                        const float dl = __fsub_rn(ph[h], S->ph[k - 1]);
                        if (fabsf(dl) > 2.0f) inc[h] = dl > 0.f ? -1.f : 1.f;
                    }
                }
                float sc = inc[0];                       // inclusive scan: elements 0..31
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float u = __shfl_up_sync(FULL, sc, o);
                    if (ln >= o) sc += u;
                }
                const float tot31 = __shfl_sync(FULL, sc, 31);
                const float inc32 = __shfl_sync(FULL, inc[1], 0);
                float turns[2];
                turns[0] = sc;
                turns[1] = tot31 + inc32 + (ln == 1 ? inc[1] : 0.f);     // element 32 (lane 0), 33 (lane 1)
                __syncwarp();
                float sum_rp = 0.f;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = ln + 32 * h;
                    if (k < np) {
                        const float rp = __fadd_rn(ph[h], __fmul_rn(turns[h], GR_PI_F));
                        S->ph[k] = rp;
                        sum_rp += rp;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {       // two independent butterflies interleaved
                    sum_rp += __shfl_xor_sync(FULL, sum_rp, o);
                    dsum += __shfl_xor_sync(FULL, dsum, o);
                }
                __syncwarp();
                if (ln == 0) {
                    const float fn = (float)np;
                    bool sweep = false;
                    if (report) {
                        O->rep_sweep = C->rep_sweep;
                        O->report_freq = C->freq;
                        C->rep_sweep = 0;
                        if (C->cl_len >= 60 * no_sec)                // checkCorrQuality (CORR_Q itself is warp 3's job: same quotient)
                            sweep = (double)C->cl_sum / (double)C->cl_len < -0.9;
                    }
                    if (sweep) {
                        S->do_sweep = 1;                             // initSweep touches every warp's state: after the barrier below
                    } else {
                        const float max_df = 20.0f / (float)no_sec;
                        const float dev = sum_rp / fn;
                        const int n4 = np < 4 ? np : 4;
                        float off = 0.f;
                        for (int i = np - n4; i < np; ++i) off += S->ph[i];
                        off = off / (float)n4;
                        float df;
                        if (C->locked) {
                            df = __fadd_rn(dev, dsum / (float)dfl);
                            if (fabsf(df) > max_df) df = df > 0.f ? max_df : -max_df;
                            if (C->df_len >= no_sec) { C->df_head = (C->df_head + 1) % GR_DF_CAP; C->df_len -= 1; }
                            G->df[(C->df_head + C->df_len) % GR_DF_CAP] = df;
                            C->df_len += 1;
                        } else {
                            df = __fmul_rn(10.0f, dev);
                            C->df_len = 1; C->df_head = 0; G->df[0] = df;
                        }
                        if (fabsf(dev) < 0.1f) C->locked = 1;
                        // demodDoppler's phase carry (gpslib.py:1345-1346), then PHASE += phaseshift
                        const float tlast = __fdiv_rn((float)ngps, GR_FS);
                        float p = fmod_2pi_f32(__fadd_rn(phase32, __fmul_rn(w32, tlast)));
                        if (p < 0.f) p += GR_TWO_PI_F;
                        C->phase = __fadd_rn(p, off);
                        const float f = __fadd_rn((float)C->freq, df);
                        if (f > a.cfg.max_freq) { C->freq = (double)a.cfg.max_freq; C->freq_weak = 1; }
                        else if (f < a.cfg.min_freq) { C->freq = (double)a.cfg.min_freq; C->freq_weak = 1; }
                        else { C->freq = (double)f; C->freq_weak = 0; }
                    }
                    O->sweep = 0;
                    O->locked = C->locked;
                    O->freq = C->freq;
                    O->freq_weak = C->freq_weak;
                    O->phase = (double)C->phase;
                }
            }
        }
        // sweep branch: thread 0 produced everything itself (no barrier needed in front of this copy); tracking branch: the
        // four warps have written their record fields directly
        if (t == 0 && S->branch_sweep) {
            // all loads first: C and O are both shared memory, the compiler will not reorder loads over stores
            const int v_sweep = C->sweep, v_delay = C->delay, v_locked = C->locked, v_ms = C->ms_time, v_prev = C->carry_cnt;
            const int v_weak = C->freq_weak, v_e0 = C->edge0, v_el = C->edge_len;
            const double v_mc = C->max_corr, v_cq = C->corr_q, v_cl = C->corr_l, v_f = C->freq, v_sd = C->std_dev;
            const float v_ph = C->phase, v_amp = C->amplitude;
            O->sweep = v_sweep;
            O->delay = v_delay;
            O->locked = v_locked;
            O->ms_time = v_ms;
            O->n_prev = v_prev;
            O->max_corr = v_mc;
            O->corr_q = v_cq;
            O->corr_l = v_cl;
            O->freq = v_f;
            O->freq_weak = v_weak;
            O->edge0 = v_e0;
            O->edge_len = v_el;
            O->phase = (double)v_ph;
            O->amplitude = v_amp;
            O->std_dev = (float)v_sd;
            O->reserved[0] = 0; O->reserved[1] = 0;
        }
        // the record was written through the generic proxy: every writer's fence orders its writes before the bulk copy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (S->do_sweep) {                                   // quality-triggered re-sweep (gpslib.py:1134-1138, 1198-1203): rare; uniform
            if (t == 0) {
                st_init_sweep(C, G, CL, a.cfg);
                O->erased |= 2;
                O->sweep = C->sweep;
                O->locked = C->locked;
                O->ms_time = C->ms_time;
                O->n_prev = C->carry_cnt;
                O->freq = C->freq;
                O->freq_weak = C->freq_weak;
                O->edge0 = C->edge0;
                O->edge_len = C->edge_len;
                O->phase = (double)C->phase;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            __syncthreads();
        }
        if (a.out_tma) {
            if (t == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(gO), "r"((unsigned)__cvta_generic_to_shared(O)), "r"((unsigned)sizeof(gr_epoch_out)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            for (int i = t; i < (int)(sizeof(gr_epoch_out) / 4); i += NT)
                reinterpret_cast<uint32_t*>(gO)[i] = reinterpret_cast<const uint32_t*>(O)[i];
        }
    }
    if (t == 0 && a.out_tma) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int i = t; i < (int)(sizeof(GrChanS) / 4); i += NT)
        reinterpret_cast<uint32_t*>(Gg)[i] = reinterpret_cast<const uint32_t*>(G)[i];
}

// ---- host API ------------------------------------------------------------------------------------------
static size_t track_smem_bytes(size_t buf_bytes = GR_TRACK_BUF_BYTES) { return buf_bytes + ((sizeof(TrackSmem) + 15) & ~(size_t)15); }
// dense form: one-buffer FFT, prompt rows staged min(n_cyc + 1, GR_PART_ROWS) at a time
static int track_dense_rows(int n_cyc) { return n_cyc + 1 < GR_PART_ROWS ? n_cyc + 1 : GR_PART_ROWS; }
static size_t track_dense_buf_bytes(int n_cyc) {
    // the dense form is the vector form (staged uint8 I/Q): its prompt rows are float2 per thread (8 bytes), and the FFT runs
    // in its one-buffer mode
    const size_t rows = (size_t)track_dense_rows(n_cyc) * 128 * 8;
    return rows > GR_ONEBUF_BYTES ? rows : (size_t)GR_ONEBUF_BYTES;
}
static size_t track_stage_bytes(int n_cyc) { return (size_t)n_cyc * GR_N * 2; }

extern "C" int gr_track_default_cfg(gr_track_cfg* cfg) {
    if (!cfg) { gr_set_error("gr_track_default_cfg: null"); return GR_ERR_ARG; }
    cfg->n_cyc = 32;            // gpsglob.py:122
    cfg->corr_avg = 8;          // gpsglob.py:68
    cfg->sweep_corr_avg = 4;    // gpsglob.py:71
    cfg->it_sweep = 40;         // gpsglob.py:41
    cfg->corr_min = 8.0f;       // gpsglob.py:69
    cfg->min_freq = -5000.0f;   // gpsglob.py:63-64
    cfg->max_freq = 5000.0f;
    cfg->step_freq = 200.0f;    // gpsglob.py:65
    cfg->in_format = GR_IN_U8IQ;
    cfg->max_channels = 16;
    return GR_OK;
}

extern "C" int gr_track_bank_create(const gr_track_cfg* cfg, gr_track_bank** bank) {
    GR_REQUIRE_INIT();
    if (!cfg || !bank) { gr_set_error("gr_track_bank_create: null argument"); return GR_ERR_ARG; }
    if (cfg->n_cyc < 8 || cfg->n_cyc > GR_MAX_NCYC || (1024 % cfg->n_cyc) != 0 || cfg->corr_avg < 1 ||
        cfg->sweep_corr_avg < 1 || cfg->sweep_corr_avg > cfg->n_cyc || cfg->it_sweep < 1 || cfg->max_channels < 1 ||
        (cfg->in_format != GR_IN_U8IQ && cfg->in_format != GR_IN_CF32)) {
        gr_set_error("gr_track_bank_create: invalid configuration (n_cyc must be 8, 16 or 32)");
        return GR_ERR_ARG;
    }
    GR_CUDA(cudaSetDevice(gr_lib()->device));
    gr_track_bank* b = new gr_track_bank();
    b->cfg = *cfg;
    b->slot_used.assign(cfg->max_channels, 0);
    b->slot_rec.assign(cfg->max_channels, 0);
    b->slots_dirty = true;
    b->d_in = nullptr; b->in_bytes = 0; b->d_out = nullptr; b->out_bytes = 0; b->last_launches = 0;
    b->in_flight = false;
    b->max_rec = 0;
    b->pipe_ready = false;
    GR_CUDA(cudaEventCreateWithFlags(&b->ev_last, cudaEventDisableTiming));
    GR_CUDA(cudaMalloc((void**)&b->d_state, sizeof(GrChan) * (size_t)cfg->max_channels));
    GR_CUDA(cudaMemset(b->d_state, 0, sizeof(GrChan) * (size_t)cfg->max_channels));
    GR_CUDA(cudaMalloc((void**)&b->d_slots, sizeof(int32_t) * (size_t)cfg->max_channels));
    GR_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    {
        const char* fast = getenv("GPSB200_TRK_FAST_NCO");
        b->exact_nco = !(fast && atoi(fast) != 0);
    }
    const int plain = (int)track_smem_bytes(), staged = (int)(track_smem_bytes() + track_stage_bytes(GR_MAX_NCYC));
#define GR_TRK_ATTR(EX)                                                                                                             \
    GR_CUDA(cudaFuncSetAttribute(track_kernel<GR_IN_U8IQ, false, false, 128, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, plain));   \
    GR_CUDA(cudaFuncSetAttribute(track_kernel<GR_IN_CF32, false, false, 128, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, plain));   \
    GR_CUDA(cudaFuncSetAttribute(track_kernel<GR_IN_U8IQ, true, false, 128, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, staged));   \
    GR_CUDA(cudaFuncSetAttribute(track_kernel<GR_IN_U8IQ, true, true, 128, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, staged));    \
    GR_CUDA(cudaFuncSetAttribute(track_kernel<GR_IN_U8IQ, true, false, 256, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, staged));
    if (b->exact_nco) { GR_TRK_ATTR(true) } else { GR_TRK_ATTR(false) }
#undef GR_TRK_ATTR
    gr_lib()->live_handles += 1;
    *bank = b;
    return GR_OK;
}

extern "C" int gr_track_bank_destroy(gr_track_bank* b) {
    if (!b) return GR_OK;
    cudaDeviceSynchronize();
    cudaFree(b->d_state);
    cudaFree(b->d_slots);
    if (b->d_in) cudaFree(b->d_in);
    if (b->d_out) cudaFree(b->d_out);
    cudaStreamDestroy(b->stream);
    cudaEventDestroy(b->ev_last);
    gr_lib()->live_handles -= 1;
    if (b->pipe_ready) {
        cudaStreamDestroy(b->s_in);
        cudaStreamDestroy(b->s_out);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(b->ev_in[i]); cudaEventDestroy(b->ev_run[i]); cudaEventDestroy(b->ev_out[i]); }
    }
    delete b;
    return GR_OK;
}

static int bank_quiesce(gr_track_bank* b) {
    // state edits from the host must not race a process call still in flight.  An event, not the caller's stream: the
    // stream of the last gr_track_process_dev call may have been destroyed by its owner since.
    if (b->in_flight) {
        GR_CUDA(cudaEventSynchronize(b->ev_last));
        b->in_flight = false;
    }
    return GR_OK;
}

extern "C" int gr_track_add(gr_track_bank* b, int rec, int prn, double freq, int delay) {
    GR_REQUIRE_INIT();
    if (!b || prn < 1 || prn > GR_MAX_PRN || rec < 0 || delay < 0 || delay >= GR_N) {
        gr_set_error("gr_track_add: invalid argument (prn %d rec %d delay %d)", prn, rec, delay);
        return GR_ERR_ARG;
    }
    int slot = -1;
    for (size_t i = 0; i < b->slot_used.size(); ++i)
        if (!b->slot_used[i]) { slot = (int)i; break; }
    if (slot < 0) { gr_set_error("gr_track_add: bank full (%d channels)", (int)b->slot_used.size()); return GR_ERR_STATE; }
    int rc = bank_quiesce(b);
    if (rc != GR_OK) return rc;
    std::vector<GrChan> staging(1);       // gpslib.py:1050-1091 (9 KB: off the stack, and not a shared static)
    GrChan& c = staging[0];
    memset(&c, 0, sizeof(c));
    c.h.active = 1; c.h.prn = prn; c.h.rec = rec; c.h.delay = delay;
    c.h.freq = freq; c.h.freq_weak = 1;
    c.h.std_dev = 0.005; c.h.std_weak = 1;
    c.h.edge0 = 0; c.h.edge_len = 1;
    c.h.df_len = 1; c.h.df_head = 0; c.df[0] = 0.f;
    c.h.cl_len = 1; c.h.cl_head = 0; c.cl[0] = 0;
    c.h.prev_stream_no = 0;
    GR_CUDA(cudaMemcpy(b->d_state + slot, &c, sizeof(GrChan), cudaMemcpyHostToDevice));
    b->slot_used[slot] = 1;
    b->slot_rec[slot] = rec;
    b->slots_dirty = true;
    return slot;
}

extern "C" int gr_track_remove(gr_track_bank* b, int slot) {
    GR_REQUIRE_INIT();
    if (!b || slot < 0 || slot >= (int)b->slot_used.size() || !b->slot_used[slot]) {
        gr_set_error("gr_track_remove: invalid slot %d", slot);
        return GR_ERR_ARG;
    }
    int rc = bank_quiesce(b);             // a launch still running reads this slot's state: let it finish before the slot can be re-used
    if (rc != GR_OK) return rc;
    b->slot_used[slot] = 0;
    b->slots_dirty = true;
    return GR_OK;
}

extern "C" int gr_track_request_sweep(gr_track_bank* b, int slot) {
    GR_REQUIRE_INIT();
    if (!b || slot < 0 || slot >= (int)b->slot_used.size() || !b->slot_used[slot]) {
        gr_set_error("gr_track_request_sweep: invalid slot %d", slot);
        return GR_ERR_ARG;
    }
    int rc = bank_quiesce(b);
    if (rc != GR_OK) return rc;
    const int32_t one = 1;
    GR_CUDA(cudaMemcpy(reinterpret_cast<char*>(b->d_state + slot) + offsetof(GrChan, h) + offsetof(GrChanHot, sweep_req), &one, sizeof(one),
                       cudaMemcpyHostToDevice));
    return GR_OK;
}

extern "C" int gr_track_num_active(const gr_track_bank* b) {
    if (!b) return 0;
    int n = 0;
    for (int u : b->slot_used) n += u;
    return n;
}

extern "C" int gr_track_last_launches(const gr_track_bank* b) { return b ? b->last_launches : 0; }
extern "C" int gr_track_bank_form(const gr_track_bank* b) { return b && b->exact_nco ? GR_TRK_FORM_EXACT : GR_TRK_FORM_FAST; }

extern "C" int gr_track_process_dev(gr_track_bank* b, const void* d_samples, int64_t rec_stride, int n_epochs,
                                    int64_t smp_time, gr_epoch_out* d_out, void* stream) {
    GR_REQUIRE_INIT();
    if (!b || !d_samples || !d_out || n_epochs < 1 || rec_stride < 0) {
        gr_set_error("gr_track_process_dev: invalid argument");
        return GR_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    GR_CUDA(cudaSetDevice(gr_lib()->device));
    if (b->slots_dirty) {
        b->active.clear();
        b->max_rec = 0;
        for (size_t i = 0; i < b->slot_used.size(); ++i)
            if (b->slot_used[i]) { b->active.push_back((int)i); if (b->slot_rec[i] > b->max_rec) b->max_rec = b->slot_rec[i]; }
        if (!b->active.empty()) {
            int rcq = bank_quiesce(b);
            if (rcq != GR_OK) return rcq;
            GR_CUDA(cudaMemcpy(b->d_slots, b->active.data(), b->active.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        }
        b->slots_dirty = false;
    }
    b->last_launches = 0;
    if (b->active.empty()) return GR_OK;
    if (b->max_rec > 0 && rec_stride != 0 && rec_stride < (int64_t)n_epochs * b->cfg.n_cyc * GR_N) {   // 0: all channels read the same samples
        gr_set_error("gr_track_process_dev: channels on recordings 0..%d need rec_stride >= %lld samples", b->max_rec,
                     (long long)n_epochs * b->cfg.n_cyc * GR_N);
        return GR_ERR_ARG;
    }
    TrackArgs a;
    a.samples = d_samples;
    a.rec_stride = rec_stride;
    a.smp_time = smp_time;
    a.n_epochs = n_epochs;
    a.n_active = (int)b->active.size();
    a.slots = b->d_slots;
    a.state = b->d_state;
    a.out = d_out;
    a.cfg = b->cfg;
    a.tab = gr_lib()->tab;
    // TMA needs 16-byte aligned sources: every recording's first sample, hence base and stride
    a.out_tma = ((uintptr_t)d_out % 16) == 0;
    a.stage = b->cfg.in_format == GR_IN_U8IQ && ((uintptr_t)d_samples % 16) == 0 && ((2 * rec_stride) % 16) == 0;
    a.buf_bytes = GR_TRACK_BUF_BYTES;
    a.part_rows = GR_PART_ROWS;
    // Form of the kernel.  wide (256 threads per channel): launches that leave SMs to spare anyway (at most one CTA per SM);
    // dense (three CTAs per SM): more channels than fit at two CTAs per SM, epoch short enough for three; else standard.
    // GPSB200_TRACK_FORM = std | dense | wide and GPSB200_TRACK_DENSE = 0 | 1 are development / test switches (a getenv per
    // launch, ~50 ns: the tests flip them between banks).
    const char* form_env = getenv("GPSB200_TRACK_FORM");
    const char* dense_env = getenv("GPSB200_TRACK_DENSE");
    const size_t dense_smem = track_smem_bytes(track_dense_buf_bytes(b->cfg.n_cyc)) + track_stage_bytes(b->cfg.n_cyc);
    const bool dense_fits = a.stage && 3 * (dense_smem + 1024) <= 228 * 1024;
    bool dense = dense_fits && (dense_env ? atoi(dense_env) != 0 : a.n_active > 2 * gr_lib()->num_sms);
    bool wide = a.stage && !dense && !dense_env && a.n_active <= gr_lib()->num_sms;
    if (form_env) {
        dense = dense_fits && !strcmp(form_env, "dense");
        wide = a.stage && !strcmp(form_env, "wide");
    }
    // the exact form's fused pass keeps one partial prompt sum per thread and pass in the scratch buffer
    if (b->exact_nco && (size_t)(b->cfg.n_cyc + 1) * 256 * 8 > (size_t)GR_TRACK_BUF_BYTES) wide = false;
    if (dense) {
        a.buf_bytes = (int)track_dense_buf_bytes(b->cfg.n_cyc);
        a.part_rows = track_dense_rows(b->cfg.n_cyc);
    }
    const size_t staged_smem = track_smem_bytes() + track_stage_bytes(b->cfg.n_cyc);
#define GR_TRK_LAUNCH(EX)                                                                                                   \
    if (dense) track_kernel<GR_IN_U8IQ, true, true, 128, EX><<<a.n_active, GR_FFT_THREADS, dense_smem, s>>>(a);                 \
    else if (wide) track_kernel<GR_IN_U8IQ, true, false, 256, EX><<<a.n_active, 256, staged_smem, s>>>(a);                       \
    else if (a.stage) track_kernel<GR_IN_U8IQ, true, false, 128, EX><<<a.n_active, GR_FFT_THREADS, staged_smem, s>>>(a);          \
    else if (b->cfg.in_format == GR_IN_U8IQ)                                                                                 \
        track_kernel<GR_IN_U8IQ, false, false, 128, EX><<<a.n_active, GR_FFT_THREADS, track_smem_bytes(), s>>>(a);               \
    else track_kernel<GR_IN_CF32, false, false, 128, EX><<<a.n_active, GR_FFT_THREADS, track_smem_bytes(), s>>>(a);
    if (b->exact_nco) { GR_TRK_LAUNCH(true) } else { GR_TRK_LAUNCH(false) }
#undef GR_TRK_LAUNCH
    GR_CUDA(cudaGetLastError());
    GR_CUDA(cudaEventRecord(b->ev_last, s));
    b->in_flight = true;
    b->last_launches = 1;
    return GR_OK;
}

// Host entry point: the recording(s) live in host memory (pinned for full speed).  The epochs
// are cut into chunks; chunk c+1 is copied in on one stream while chunk c is tracked on a
// second and chunk c-1's records are copied out on a third (double-buffered) -- the stream
// pipeline that replaces gpsrecv's deque + worker pool (gpsrecv.py:47-104, 404-417).
extern "C" int gr_track_process_host(gr_track_bank* b, const void* h_samples, int64_t rec_stride, int nrec,
                                     int n_epochs, int64_t smp_time, gr_epoch_out* h_out) {
    GR_REQUIRE_INIT();
    if (!b || !h_samples || !h_out || n_epochs < 1 || nrec < 1) {
        gr_set_error("gr_track_process_host: invalid argument");
        return GR_ERR_ARG;
    }
    GR_CUDA(cudaSetDevice(gr_lib()->device));
    const size_t bps = b->cfg.in_format == GR_IN_U8IQ ? 2 : 8;
    const size_t ngps = (size_t)b->cfg.n_cyc * GR_N;
    const size_t span = (size_t)n_epochs * ngps;
    if (nrec > 1 && (size_t)rec_stride < span) { gr_set_error("gr_track_process_host: rec_stride too short"); return GR_ERR_ARG; }
    const int nact = gr_track_num_active(b);
    if (nact == 0) return GR_OK;
    for (size_t i = 0; i < b->slot_used.size(); ++i)
        if (b->slot_used[i] && b->slot_rec[i] >= nrec) {
            gr_set_error("gr_track_process_host: slot %d tracks recording %d, but the call holds %d recording(s)", (int)i, b->slot_rec[i], nrec);
            return GR_ERR_ARG;
        }
    // chunk: about 16 MiB of samples over all recordings, at least one epoch
    size_t ce = (16u << 20) / (ngps * bps * (size_t)nrec);
    if (ce < 1) ce = 1;
    if (ce > (size_t)n_epochs) ce = (size_t)n_epochs;
    const size_t chunk_samples = ce * ngps;                     // per recording
    const size_t in_bytes = 2 * chunk_samples * bps * (size_t)nrec;
    const size_t out_bytes = 2 * ce * (size_t)nact * sizeof(gr_epoch_out);
    if (in_bytes > b->in_bytes) {
        if (b->d_in) cudaFree(b->d_in);
        b->d_in = nullptr; b->in_bytes = 0;
        GR_CUDA(cudaMalloc(&b->d_in, in_bytes));
        b->in_bytes = in_bytes;
    }
    if (out_bytes > b->out_bytes) {
        if (b->d_out) cudaFree(b->d_out);
        b->d_out = nullptr; b->out_bytes = 0;
        GR_CUDA(cudaMalloc((void**)&b->d_out, out_bytes));
        b->out_bytes = out_bytes;
    }
    if (!b->pipe_ready) {
        GR_CUDA(cudaStreamCreateWithFlags(&b->s_in, cudaStreamNonBlocking));
        GR_CUDA(cudaStreamCreateWithFlags(&b->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            GR_CUDA(cudaEventCreateWithFlags(&b->ev_in[i], cudaEventDisableTiming));
            GR_CUDA(cudaEventCreateWithFlags(&b->ev_run[i], cudaEventDisableTiming));
            GR_CUDA(cudaEventCreateWithFlags(&b->ev_out[i], cudaEventDisableTiming));
        }
        b->pipe_ready = true;
    }
    const char* hin = reinterpret_cast<const char*>(h_samples);
    int launches = 0;
    int c = 0;
    for (size_t e0 = 0; e0 < (size_t)n_epochs; e0 += ce, ++c) {
        const int buf = c & 1;
        const size_t ne = (e0 + ce <= (size_t)n_epochs) ? ce : (size_t)n_epochs - e0;
        char* din = reinterpret_cast<char*>(b->d_in) + (size_t)buf * chunk_samples * bps * (size_t)nrec;
        gr_epoch_out* dout = b->d_out + (size_t)buf * ce * (size_t)nact;
        if (c >= 2) GR_CUDA(cudaStreamWaitEvent(b->s_in, b->ev_run[buf], 0));      // kernel of chunk c-2 has read this buffer
        for (int r = 0; r < nrec; ++r)
            GR_CUDA(cudaMemcpyAsync(din + (size_t)r * chunk_samples * bps, hin + ((size_t)r * (size_t)rec_stride + e0 * ngps) * bps,
                                    ne * ngps * bps, cudaMemcpyHostToDevice, b->s_in));
        GR_CUDA(cudaEventRecord(b->ev_in[buf], b->s_in));
        GR_CUDA(cudaStreamWaitEvent(b->stream, b->ev_in[buf], 0));
        if (c >= 2) GR_CUDA(cudaStreamWaitEvent(b->stream, b->ev_out[buf], 0));    // records of chunk c-2 have left
        int rc = gr_track_process_dev(b, din, (int64_t)chunk_samples, (int)ne, smp_time + (int64_t)(e0 * ngps), dout, (void*)b->stream);
        if (rc != GR_OK) return rc;
        launches += b->last_launches;
        GR_CUDA(cudaEventRecord(b->ev_run[buf], b->stream));
        GR_CUDA(cudaStreamWaitEvent(b->s_out, b->ev_run[buf], 0));
        GR_CUDA(cudaMemcpyAsync(h_out + e0 * (size_t)nact, dout, ne * (size_t)nact * sizeof(gr_epoch_out), cudaMemcpyDeviceToHost, b->s_out));
        GR_CUDA(cudaEventRecord(b->ev_out[buf], b->s_out));
    }
    GR_CUDA(cudaStreamSynchronize(b->s_in));
    GR_CUDA(cudaStreamSynchronize(b->stream));
    GR_CUDA(cudaStreamSynchronize(b->s_out));
    b->last_launches = launches;
    return GR_OK;
}
