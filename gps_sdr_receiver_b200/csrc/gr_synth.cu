// Device-side generator of synthetic RTL-SDR recordings (uint8 I,Q) with known GPS L1 C/A
// satellites.  MEASUREMENT / TEST INFRASTRUCTURE: it makes the
// multi-gigabyte inputs of the throughput configurations directly in HBM from a seed; it
// is not on the receiver's product path.  Byte layout and scaling are the ones the
// reference reader expects (src/gpsrecv.py:168-173): byte 0 = I, byte 1 = Q,
// value = round((x + 1) * 127.5) clipped to [0, 255].
//
// Signal model per satellite (same as gps_sdr_receiver_b200/synth.py, different random
// streams):  a * chip[floor(((n - tau) mod 2048) * 1023/2048)] * navbit(n)
//            * exp(i (2 pi (f t + fdot t^2 / 2) + phi0)),  t = (n+1)/fs,
// plus complex white Gaussian noise of standard deviation `noise_sigma` per component.
#include <math.h>
#include <stdio.h>

#include "gr_internal.h"

#define GR_SYNTH_MAX_SAT 16

struct SynthArgs {
    uint8_t* out;
    long long nsamples, start_sample;
    int nsat;
    float noise_sigma;
    unsigned long long seed;
    gr_synth_sat sats[GR_SYNTH_MAX_SAT];
    const int8_t* chips;   // [GR_MAX_PRN+1][1024]
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {   // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) synth_kernel(const SynthArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.nsamples) return;
    const long long n = a.start_sample + i;
    const double t = (double)(n + 1) / 2048000.0;
    float re = 0.f, im = 0.f;
    for (int s = 0; s < a.nsat; ++s) {
        const gr_synth_sat& S = a.sats[s];
        const double rel = (double)n - S.delay;
        const double code_no_f = floor(rel / 2048.0);
        const double ph = rel - 2048.0 * code_no_f;                       // (n - tau) mod 2048
        int ci = (int)floor(ph * (1023.0 / 2048.0));
        ci = ci >= 1023 ? ci - 1023 : ci;
        const float chip = (float)a.chips[S.prn * 1024 + ci];
        const long long code_no = (long long)code_no_f;
        long long bit_no = code_no - S.bit_offset_ms;
        bit_no = bit_no >= 0 ? bit_no / 20 : -((-bit_no + 19) / 20);      // floor division
        const unsigned long long h = mix64(((unsigned long long)S.prn << 40) ^ ((unsigned long long)S.bit_seed << 48) ^
                                           (unsigned long long)(bit_no + (1ll << 38)));
        const float nav = (h & 1ull) ? 1.f : -1.f;
        double cyc = S.doppler * t + 0.5 * S.doppler_rate * t * t;
        cyc -= floor(cyc);
        float sn, cs;
        sincosf((float)(6.283185307179586 * cyc) + (float)S.phi0, &sn, &cs);
        const float amp = S.amp * chip * nav;
        re = fmaf(amp, cs, re);
        im = fmaf(amp, sn, im);
    }
    // Box-Muller on a counter-based hash of (seed, n)
    const unsigned long long r = mix64(a.seed * 0xD1342543DE82EF95ull + (unsigned long long)n);
    const float u1 = ((float)(unsigned)(r >> 40) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(unsigned)((r >> 8) & 0xFFFFFFu)) * (1.0f / 16777216.0f);
    const float rad = a.noise_sigma * sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    re = fmaf(rad, cs, re);
    im = fmaf(rad, sn, im);
    const float qi = fminf(fmaxf(rintf((re + 1.0f) * 127.5f), 0.f), 255.f);
    const float qq = fminf(fmaxf(rintf((im + 1.0f) * 127.5f), 0.f), 255.f);
    reinterpret_cast<uchar2*>(a.out)[i] = make_uchar2((unsigned char)qi, (unsigned char)qq);
}

extern "C" int gr_synth_iq_dev(uint8_t* d_out, int64_t nsamples, int64_t start_sample, const gr_synth_sat* sats,
                               int nsat, float noise_sigma, uint64_t seed, void* stream) {
    GR_REQUIRE_INIT();
    if (!d_out || nsamples < 1 || nsat < 0 || nsat > GR_SYNTH_MAX_SAT || (nsat > 0 && !sats)) {
        gr_set_error("gr_synth_iq_dev: invalid argument (at most %d satellites)", GR_SYNTH_MAX_SAT);
        return GR_ERR_ARG;
    }
    SynthArgs a;
    a.out = d_out;
    a.nsamples = nsamples;
    a.start_sample = start_sample;
    a.nsat = nsat;
    a.noise_sigma = noise_sigma;
    a.seed = seed;
    for (int s = 0; s < nsat; ++s) {
        if (sats[s].prn < 1 || sats[s].prn > GR_MAX_PRN) { gr_set_error("gr_synth_iq_dev: prn %d out of range", sats[s].prn); return GR_ERR_ARG; }
        a.sats[s] = sats[s];
    }
    a.chips = gr_lib()->tab.chips;
    const long long nblk = (nsamples + 255) / 256;
    if (nblk > 0x7fffffffLL) { gr_set_error("gr_synth_iq_dev: too many samples for one call"); return GR_ERR_ARG; }
    synth_kernel<<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(a);
    GR_CUDA(cudaGetLastError());
    return GR_OK;
}


// ---- geometry-consistent generator (gr_synth_geo_dev) ---------------------------------------------------
struct SynthGeoArgs {
    uint8_t* out;
    long long nsamples, start_sample;
    long long t0_ms;
    double t0_frac, inv_node_dt;
    int nsat;
    float noise_sigma;
    unsigned long long seed;
    gr_synth_geo_sat sats[GR_SYNTH_MAX_SAT];
    const int8_t* chips;
};

__global__ void __launch_bounds__(256) synth_geo_kernel(const SynthGeoArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.nsamples) return;
    const long long n = a.start_sample + i;
    const double trel = (double)n / 2048000.0;                      // receiver clock since sample 0
    const double x = trel * a.inv_node_dt + 1.0;                    // node 0 sits one step before sample 0
    float re = 0.f, im = 0.f;
    for (int s = 0; s < a.nsat; ++s) {
        const gr_synth_geo_sat& S = a.sats[s];
        int k = (int)floor(x);
        k = k < 1 ? 1 : (k > S.n_nodes - 3 ? S.n_nodes - 3 : k);
        const double u = x - (double)k;
        const double y0 = S.d_tau[k - 1], y1 = S.d_tau[k], y2 = S.d_tau[k + 1], y3 = S.d_tau[k + 2];
        // 4-point Lagrange through nodes k-1 .. k+2 at offset u from node k
        const double tau = -u * (u - 1.0) * (u - 2.0) / 6.0 * y0 + (u + 1.0) * (u - 1.0) * (u - 2.0) / 2.0 * y1 -
                           (u + 1.0) * u * (u - 2.0) / 2.0 * y2 + (u + 1.0) * u * (u - 1.0) / 6.0 * y3;
        const double ysat = (a.t0_frac + (trel - tau)) * 1e3;       // satellite clock, ms past t0_ms
        const double ms_f = floor(ysat);
        const double frac = ysat - ms_f;
        int ci = (int)(frac * 1023.0);
        ci = ci >= 1023 ? 1022 : ci;
        const float chip = (float)a.chips[S.prn * 1024 + ci];
        const long long ms_abs = a.t0_ms + (long long)ms_f;
        long long b = ms_abs - S.bit_t0_ms;
        b = b >= 0 ? b / 20 : -((-b + 19) / 20);
        b = b < 0 ? 0 : (b >= S.n_bits ? S.n_bits - 1 : b);
        const float nav = S.d_bits[b] ? 1.f : -1.f;
        double cyc = -1575.42e6 * tau;
        cyc -= floor(cyc);
        float sn, cs;
        sincospif((float)(2.0 * cyc), &sn, &cs);
        const float amp = S.amp * chip * nav;
        re = fmaf(amp, cs, re);
        im = fmaf(amp, sn, im);
    }
    const unsigned long long r = mix64(a.seed * 0xD1342543DE82EF95ull + (unsigned long long)n);
    const float u1 = ((float)(unsigned)(r >> 40) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(unsigned)((r >> 8) & 0xFFFFFFu)) * (1.0f / 16777216.0f);
    const float rad = a.noise_sigma * sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    re = fmaf(rad, cs, re);
    im = fmaf(rad, sn, im);
    const float qi = fminf(fmaxf(rintf((re + 1.0f) * 127.5f), 0.f), 255.f);
    const float qq = fminf(fmaxf(rintf((im + 1.0f) * 127.5f), 0.f), 255.f);
    reinterpret_cast<uchar2*>(a.out)[i] = make_uchar2((unsigned char)qi, (unsigned char)qq);
}

extern "C" int gr_synth_geo_dev(uint8_t* d_out, int64_t nsamples, int64_t start_sample, const gr_synth_geo_sat* sats, int nsat,
                                int64_t t0_ms, double t0_frac, double node_dt, float noise_sigma, uint64_t seed, void* stream) {
    GR_REQUIRE_INIT();
    if (!d_out || nsamples < 1 || nsat < 1 || nsat > GR_SYNTH_MAX_SAT || !sats || !(node_dt > 0.0)) {
        gr_set_error("gr_synth_geo_dev: invalid argument (1..%d satellites)", GR_SYNTH_MAX_SAT);
        return GR_ERR_ARG;
    }
    SynthGeoArgs a;
    a.out = d_out; a.nsamples = nsamples; a.start_sample = start_sample;
    a.t0_ms = t0_ms; a.t0_frac = t0_frac; a.inv_node_dt = 1.0 / node_dt;
    a.nsat = nsat; a.noise_sigma = noise_sigma; a.seed = seed;
    const double t_end = (double)(start_sample + nsamples) / 2048000.0;
    for (int s = 0; s < nsat; ++s) {
        if (sats[s].prn < 1 || sats[s].prn > GR_MAX_PRN || !sats[s].d_tau || !sats[s].d_bits || sats[s].n_bits < 1 ||
            sats[s].n_nodes < 4 || (double)(sats[s].n_nodes - 3) * node_dt < t_end) {
            gr_set_error("gr_synth_geo_dev: satellite %d: bad prn / pointers / too few nodes for the requested span", s);
            return GR_ERR_ARG;
        }
        a.sats[s] = sats[s];
    }
    a.chips = gr_lib()->tab.chips;
    const long long nblk = (nsamples + 255) / 256;
    if (nblk > 0x7fffffffLL) { gr_set_error("gr_synth_geo_dev: too many samples for one call"); return GR_ERR_ARG; }
    synth_geo_kernel<<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(a);
    GR_CUDA(cudaGetLastError());
    return GR_OK;
}
