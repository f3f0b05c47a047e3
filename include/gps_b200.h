/* gps_b200.h -- C ABI of the B200-native GPS L1 C/A acquisition + tracking hot path.
 *
 * Plain C, plain pointers and sizes; no torch / C++ types.  This is the boundary
 * a maintainer of annappo/GPS-SDR-Receiver binds (ctypes) to
 * replace the numpy/scipy code on the path
 *     gpsrecv.sweepAllSats / demodDoppler / findCodePhase      (src/gpsrecv.py:217-274)
 *     gpslib.GPSCacode / GPSCacodeRep / SatStream.process      (src/gpslib.py:62-87, 1044-1446)
 *     gpsrecv.initMultiProcPool / satCalc worker pool          (src/gpsrecv.py:300-417)
 * Every entry point returns GR_OK (0) or a negative error code; gr_last_error()
 * gives the text.  There is NO CPU fallback: without a CUDA device gr_init fails.
 *
 * Conventions
 *   - raw I/Q is the RTL-SDR byte stream: uint8 I, uint8 Q per sample, 2.048 MS/s,
 *     2048 samples per 1-ms code period (gpsrecv.py:168-173, gpsglob.py:119-125);
 *   - "dev" entry points take device pointers (e.g. torch tensor .data_ptr()) and a
 *     cudaStream_t passed as void*; they are asynchronous on that stream;
 *   - "host" entry points take host pointers, copy in, run, copy out and return
 *     when the results are in the caller's buffer.
 */
#ifndef GPS_B200_H
#define GPS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GR_OK 0
#define GR_ERR_CUDA (-1)   /* a CUDA runtime call failed (no device, launch error, ...) */
#define GR_ERR_ARG (-2)    /* invalid argument */
#define GR_ERR_STATE (-3)  /* library not initialised / handle invalid */

#define GR_CODE_SAMPLES 2048
#define GR_NUM_PRN 37

/* ---- library ------------------------------------------------------------------ */
int gr_version(void);
/* Select `device`, build the Gold-code tables (replaces src/cacodes.py:5-80 and
 * gpslib.GPSCacode, gpslib.py:62-77) and their conjugate spectra
 * (gpsrecv.py:574-577) and make them device resident.  Idempotent. */
int gr_init(int device);
int gr_shutdown(void);
const char* gr_last_error(void);

/* ---- tables (host copies of what the kernels use) --------------------------------- */
/* 1023 chips, +1 / -1   (cacodes.py: cacodes[prn]) */
int gr_get_chips(int prn, int8_t* out1023);
/* gpslib.GPSCacode(prn): float64[2048], bit-identical to the reference's table */
int gr_get_cacode(int prn, double* out2048);
/* scipy.fft.fft(GPSCacode(prn)) as interleaved re,im float64[4096] (gpsrecv.py:577) */
int gr_get_code_spectrum(int prn, double* out4096);

/* ---- acquisition: PRN x Doppler x code-phase search -------------------------------- */
enum { GR_IN_U8IQ = 0,   /* uint8 I,Q pairs (2 bytes / sample)                  */
       GR_IN_CF32 = 1 }; /* complex64 samples as the reference's reader makes   */
enum { GR_ACQ_ABS = 0,   /* statistic |c|       (gpsrecv.py:258), nnoncoh = 1     */
       GR_ACQ_POW = 1 }; /* statistic sum_k |c_k|^2 over nnoncoh intervals        */
#define GR_SECOND_PEAK_GUARD 4 /* lags excluded each side of the peak for `second` */

/* One (PRN, Doppler bin) cell reduced over its 2048 code phases
 * (gpsrecv.findCodePhase, gpsrecv.py:217-227, plus neighbours and 2nd peak). */
typedef struct gr_acq_cell {
    int32_t mx;     /* argmax lag (first maximum), 0..2047  -> "delay"          */
    float peak;     /* statistic at mx                                          */
    float mean;     /* mean over the 2048 lags                                  */
    float std;      /* population standard deviation                            */
    float z;        /* (peak - mean) / std  -> "normMaxCorr"                    */
    float em1, ep1; /* statistic at mx-1, mx+1 (circular)                       */
    float second;   /* largest value farther than GR_SECOND_PEAK_GUARD from mx  */
} gr_acq_cell;

typedef struct gr_acq_plan gr_acq_plan;

/* A plan fixes the search grid.  prns[nprn] in 1..37; bin_hz[nbins] Doppler bins in
 * Hz (float64, as the reference's python floats); tcoh_ms 1-ms blocks are summed
 * coherently (gpsrecv.py:250-254), nnoncoh such intervals are accumulated as |.|^2.
 * A recording must hold tcoh_ms*nnoncoh*2048 samples.  Wipe-off uses phase 0 and the
 * reference's float32 time base t[n] = (n+1)/fs (gpsrecv.py:32-33, 232-235).
 * Bins that differ by a multiple of fs/2048 = 1 kHz share one forward FFT (their spectra
 * are circular shifts of each other); a cell's value does not depend on which other
 * bins the plan holds.  |bin_hz| must stay below fs/2.
 * Form of the forward kernel (gr_acq_plan_form): the reference's float32 phase argument
 * fl32(w32 * fl32((n+1)/fs)) carries a rounding error that grows with |f| T.  Searches where it
 * stays far below the 1e-4 tolerance (largest argument x 2^-24 <= 1e-4 rad: e.g. +-10 kHz over
 * 10 ms) run in the FAST form (shared spectra, block rotations); longer / wider searches (e.g.
 * +-10 kHz over 200 ms) run in the EXACT form, which reproduces the reference's argument for
 * every sample of every bin.  Environment (read at creation): GPSB200_ACQ_EXACT_NCO=1 / =0
 * forces the exact / fast form; GPSB200_ACQ_NOSHARE=1 one forward FFT per bin in the fast form. */
#define GR_ACQ_FORM_FAST 0
#define GR_ACQ_FORM_EXACT 1
/* The classification gr_acq_plan_create applies to its Doppler bins (host only, needs no GPU): base[b] = index of the
 * forward spectrum bin b uses, shift[b] = its circular shift in FFT bins (0..2047), base_hz[i] = frequency the i-th base
 * spectrum is computed for (in [-500, 500) Hz when sharing is on).  Returns the number of base spectra or a negative
 * error code.  share = 0: one spectrum per bin. */
int gr_acq_classify_bins(const double* bin_hz, int nbins, int share, int32_t* base, int32_t* shift, double* base_hz);
int gr_acq_plan_create(const int32_t* prns, int nprn, const double* bin_hz, int nbins,
                       int tcoh_ms, int nnoncoh, int mode, int in_format, gr_acq_plan** plan);
int gr_acq_plan_destroy(gr_acq_plan* plan);
/* GR_ACQ_FORM_FAST or GR_ACQ_FORM_EXACT: which form gr_acq_plan_create chose for this grid */
int gr_acq_plan_form(const gr_acq_plan* plan);
/* nrec independent recordings, `rec_stride` samples apart; d_out[nrec][nprn][nbins]. */
int gr_acq_run_dev(gr_acq_plan* plan, const void* d_samples, int nrec, int64_t rec_stride,
                   gr_acq_cell* d_out, void* stream);
int gr_acq_run_host(gr_acq_plan* plan, const void* h_samples, int nrec, int64_t rec_stride,
                    gr_acq_cell* h_out);
/* kernel launches issued by the last run of this plan (for bench accounting) */
int gr_acq_last_launches(const gr_acq_plan* plan);
/* Form of the inverse kernel the last run launched.  Both compute the same cells bit for bit; the launcher picks per call:
 * GR_ACQ_INV_4CTA = acq_inv_kernel, four 128-thread CTAs per SM, work handed out in items of 4 PRNs x all intervals;
 * GR_ACQ_INV_QUAD = acq_inv_quad_kernel, one 512-thread CTA per SM whose four groups share every staged forward spectrum,
 * work handed out per (recording, bin): faster on launches that fill the GPU many times over, slower on small ones;
 * GR_ACQ_INV_BOTH = the quad form for the whole waves of the launch, then the 4-CTA form for the rest.
 * Environment (read when the plan is created): GPSB200_ACQ_QUAD=0 / =1 / =2 forces a form. */
#define GR_ACQ_INV_4CTA 0
#define GR_ACQ_INV_QUAD 1
#define GR_ACQ_INV_BOTH 2 /* quad form for the whole waves of the launch, 4-CTA form for the rest */
int gr_acq_last_inverse_form(const gr_acq_plan* plan);

/* The search result proper: for every recording and PRN the Doppler bin with the largest
 * z (first maximum), i.e. the (PRN, Doppler, code-phase, metric) tuple.  The full cell grid
 * stays in device memory (plan-owned scratch); only nrec*nprn tuples are written. */
typedef struct gr_acq_best {
    int32_t prn;
    int32_t bin;        /* index into the plan's bin_hz[]                                 */
    gr_acq_cell cell;   /* the winning (PRN, bin) cell                                    */
} gr_acq_best;
int gr_acq_search_dev(gr_acq_plan* plan, const void* d_samples, int nrec, int64_t rec_stride,
                      gr_acq_best* d_best, void* stream);
int gr_acq_search_host(gr_acq_plan* plan, const void* h_samples, int nrec, int64_t rec_stride,
                       gr_acq_best* h_best);

/* ---- tracking: a bank of channels (replaces the multiprocessing pool) ------------- */
typedef struct gr_track_bank gr_track_bank;

typedef struct gr_track_cfg {      /* gpsglob.py:38-42, 63-75, 119-125                 */
    int32_t n_cyc;                 /* N_CYC: 1-ms code periods per epoch (8, 16 or 32) */
    int32_t corr_avg;              /* CORR_AVG (clamped to n_cyc, gpslib.py:1072)      */
    int32_t sweep_corr_avg;        /* SWEEP_CORR_AVG                                   */
    int32_t it_sweep;              /* IT_SWEEP                                         */
    float corr_min;                /* CORR_MIN                                         */
    float min_freq, max_freq;      /* MIN_FREQ, MAX_FREQ                               */
    float step_freq;               /* STEP_FREQ                                        */
    int32_t in_format;             /* GR_IN_U8IQ | GR_IN_CF32                          */
    int32_t max_channels;          /* capacity of the bank                             */
} gr_track_cfg;

#define GR_MAX_NCYC 32
#define GR_MAX_PROMPT (GR_MAX_NCYC + 2)

/* Per channel and epoch: everything SatStream.process leaves behind that a caller can
 * observe (gpslib.py:1141-1210).  Doubles where the reference holds python floats /
 * float64, floats where it holds float32.  448 bytes (a multiple of 16: the kernel
 * stores records with one TMA bulk copy each). */
typedef struct gr_epoch_out {
    int32_t prn;
    int32_t sweep;            /* SWEEP after this epoch                                  */
    int32_t tracked;          /* 1: tracking branch ran, 0: sweep branch                 */
    int32_t delay;            /* DELAY                                                   */
    int32_t corr_delay;       /* delay found by this epoch's correlation, -1 if z<=min   */
    int32_t locked;           /* PHASE_LOCKED after the epoch                            */
    int32_t locked_in;        /* PHASE_LOCKED while decodeData ran (edges, MS_TIME)      */
    int32_t report;           /* 1 if streamNo % NO_SEC == 0 (frameLst due)              */
    int32_t rep_sweep;        /* 'SWP' flag of that report (gpslib.py:1124-1131)         */
    int32_t n_prompt;         /* number of 1-ms prompt values (N_CYC or N_CYC+1)         */
    int32_t ms_time;          /* MS_TIME after the epoch                                 */
    int32_t n_prev;           /* len(PREV_SAMPLES)                                       */
    int32_t prompt_b1;        /* start of the 2nd prompt segment in decodeData's index   */
    int32_t freq_weak;        /* 1: FREQ is a python float, 0: numpy float32             */
    int32_t edge0;            /* EDGES[0] after the epoch: sign of the first signal, 0 = unset */
    int32_t edge_len;         /* len(EDGES) after the epoch (and after evalEdges' trim)  */
    int64_t prompt_st0;       /* ST: sample time of prompt 0 (gpslib.py:1408-1412);      */
                              /* prompt k>0 starts at ST + prompt_b1 + 2048 (k-1)        */
    uint64_t edge_mask;       /* bit k set: EDGES.append((MS_TIME_k, ST + n0_k)) at      */
                              /* prompt k, MS_TIME_k = ms_time - n_prompt + k            */
    double code_phase;        /* codePhase (-1.0 if none)                                */
    double max_corr;          /* MAX_CORR (normMaxCorr) -> 'CRM'                         */
    double corr_q, corr_l;    /* CORR_Q, CORR_L                                          */
    double freq;              /* FREQ after the PLL update                               */
    double report_freq;       /* FREQ at report time -> 'FRQ' (valid when report)        */
    double phase;             /* PHASE after the PLL update                              */
    float amplitude, std_dev; /* AMPLITUDE -> 'AMP', STD_DEV                             */
    float corr3[3];           /* corr[mx-1], corr[mx], corr[mx+1]                        */
    float corr_mean, corr_std;
    int32_t erased;           /* bit 0: EDGES/PREV_SAMPLES erased before this epoch (stream  */
                              /* gap or sweep request), bit 1: initSweep ran at its end   */
    float prompt[2 * GR_MAX_PROMPT];   /* gpsData re,im (complex64)                      */
    int32_t reserved[2];               /* pads the record to 448 bytes                   */
} gr_epoch_out;

int gr_track_default_cfg(gr_track_cfg* cfg);
int gr_track_bank_create(const gr_track_cfg* cfg, gr_track_bank** bank);
int gr_track_bank_destroy(gr_track_bank* bank);
/* gpslib.SatStream(satNo, freq, delay=delay, ...) (gpslib.py:1050-1091) on recording
 * `rec` of the sample buffer: returns the slot (>= 0). */
int gr_track_add(gr_track_bank* bank, int rec, int prn, double freq, int delay);
int gr_track_remove(gr_track_bank* bank, int slot);           /* `del inst`             */
int gr_track_request_sweep(gr_track_bank* bank, int slot);    /* process(..., sweep=True) */
/* Run `n_epochs` consecutive epochs for every active channel in ONE launch (replaces
 * gpsrecv.satCalc's fan-out/fan-in over the process pool, gpsrecv.py:404-417).
 * d_samples: recordings `rec_stride` samples apart, each n_epochs * n_cyc * 2048 samples,
 * shared by all channels of that recording; smp_time: SMP_TIME of the first epoch
 * (sample count of data[0], gpsrecv.py:470); out[n_epochs][n_active] ordered by slot. */
int gr_track_process_dev(gr_track_bank* bank, const void* d_samples, int64_t rec_stride, int n_epochs,
                         int64_t smp_time, gr_epoch_out* d_out, void* stream);
int gr_track_process_host(gr_track_bank* bank, const void* h_samples, int64_t rec_stride, int nrec,
                          int n_epochs, int64_t smp_time, gr_epoch_out* h_out);
int gr_track_num_active(const gr_track_bank* bank);
int gr_track_last_launches(const gr_track_bank* bank);
/* NCO form of the bank's kernel.  EXACT (default): every sample is rotated by the reference's own float32 phase argument
 * exp(-i fl32(PHASE + fl32(w * SEC_TIME[n]))) (gpslib.py:1343-1346).  FAST (GPSB200_TRK_FAST_NCO=1 when the bank is
 * created): the rotation is factorised (2 sin/cos per thread and epoch) and mathematically exact, i.e. it does not
 * carry the reference's per-sample float32 argument noise (6e-5 rad at 5 kHz x 32 ms): same decisions, FREQ a few float32
 * ulp apart, carrier phase / complex prompts up to 1e-4 / 5e-4 rad apart. */
#define GR_TRK_FORM_FAST 0
#define GR_TRK_FORM_EXACT 1
int gr_track_bank_form(const gr_track_bank* bank);

/* ---- synthetic recordings (measurement / test infrastructure) ------------ */
typedef struct gr_synth_sat {
    int32_t prn;
    int32_t bit_offset_ms;   /* first nav-bit boundary, ms since the code start nearest n=0 */
    uint32_t bit_seed;       /* nav-bit stream                                             */
    float amp;               /* amplitude relative to full scale 1.0                       */
    double doppler;          /* Hz at n = 0                                                */
    double doppler_rate;     /* Hz/s                                                       */
    double delay;            /* code phase in samples, 0 <= delay < 2048                   */
    double phi0;             /* carrier phase at t = 0, rad                                */
} gr_synth_sat;
/* uint8 I,Q bytes of `nsamples` samples starting at sample `start_sample` of a recording
 * defined by (sats, noise_sigma, seed): reproducible in pieces.  At most 16 satellites. */
int gr_synth_iq_dev(uint8_t* d_out, int64_t nsamples, int64_t start_sample, const gr_synth_sat* sats,
                    int nsat, float noise_sigma, uint64_t seed, void* stream);

/* Geometry-consistent variant: every satellite's signal is a function of its own clock
 * reading t_sat = t_rx - tau(t_rx), tau given at nodes `node_dt` seconds apart (node 0 one step before sample 0,
 * 4-point Lagrange interpolation): code phase = t_sat mod 1 ms, nav bit = bits[floor((t_sat - bit_t0) / 20 ms)],
 * carrier phase = -2 pi f_L1 tau.  t_rx = t0_ms * 1e-3 + t0_frac + n / fs is the receiver clock reading. */
typedef struct gr_synth_geo_sat {
    int32_t prn;
    int32_t n_nodes;
    int32_t n_bits;
    float amp;
    int64_t bit_t0_ms;       /* satellite-clock time of bits[0], ms                                */
    const double* d_tau;     /* device pointer: n_nodes values of (receiver clock - satellite clock), s */
    const int8_t* d_bits;    /* device pointer: n_bits nav bits, 0 / 1                             */
} gr_synth_geo_sat;
int gr_synth_geo_dev(uint8_t* d_out, int64_t nsamples, int64_t start_sample, const gr_synth_geo_sat* sats, int nsat,
                     int64_t t0_ms, double t0_frac, double node_dt, float noise_sigma, uint64_t seed, void* stream);

/* ---- debug / test hooks ----------------------------------------------------------------- */
/* forward (inverse=0) or unnormalised inverse FFT-2048 of `batch` vectors, complex64 */
int gr_debug_fft2048(const float* h_in, float* h_out, int batch, int inverse);
/* FP32 FFMA throughput of the device (all SMs, register-resident chains): the measured
 * denominator of the acquisition roofline.  Returns TFLOP/s (2 flop per FFMA). */
int gr_debug_fp32_peak(int iters, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* GPS_B200_H */
