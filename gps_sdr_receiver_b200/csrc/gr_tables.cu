// Constant tables: Gold codes, the resampled 2048-point code, conjugate code
// spectra and FFT twiddles; library lifetime (gr_init / gr_shutdown).
//
// Replaces  src/cacodes.py:5-80          (literal chip table -> G1/G2 shift registers)
//           src/gpslib.py:62-77          (doubledCacode + GPSCacode: 1023 -> 2046 -> 2048)
//           src/gpsrecv.py:574-577       (FFT_CACODE = fft(GPSCacode(prn)))
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <complex>
#include <vector>

#include "gr_internal.h"

static GrLib g_lib;
static thread_local char g_err[512] = "";      // per calling thread, like errno

GrLib* gr_lib() { return &g_lib; }

void gr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* gr_last_error(void) { return g_err; }
extern "C" int gr_version(void) { return 200; }

// G2 output taps for PRN 1..37 (IS-GPS-200, table 3-Ia).
static const int kG2Taps[GR_MAX_PRN][2] = {
    {2, 6}, {3, 7}, {4, 8}, {5, 9}, {1, 9}, {2, 10}, {1, 8}, {2, 9}, {3, 10}, {2, 3}, {3, 4}, {5, 6}, {6, 7},
    {7, 8}, {8, 9}, {9, 10}, {1, 4}, {2, 5}, {3, 6}, {4, 7}, {5, 8}, {6, 9}, {1, 3}, {4, 6}, {5, 7}, {6, 8},
    {7, 9}, {8, 10}, {1, 6}, {2, 7}, {3, 8}, {4, 9}, {5, 10}, {4, 10}, {1, 7}, {2, 8}, {4, 10}};

static void gold_chips(int prn, int8_t* out) {
    int g1[10], g2[10];
    for (int i = 0; i < 10; ++i) g1[i] = g2[i] = 1;
    const int a = kG2Taps[prn - 1][0] - 1, b = kG2Taps[prn - 1][1] - 1;
    for (int i = 0; i < 1023; ++i) {
        out[i] = (g1[9] ^ g2[a] ^ g2[b]) ? 1 : -1;
        const int f1 = g1[2] ^ g1[9];
        const int f2 = g2[1] ^ g2[2] ^ g2[5] ^ g2[7] ^ g2[8] ^ g2[9];
        for (int k = 9; k > 0; --k) { g1[k] = g1[k - 1]; g2[k] = g2[k - 1]; }
        g1[0] = f1;
        g2[0] = f2;
    }
}

// gpslib.py:62-77.  The reference interpolates the doubled code (2046 points at
// x = 0..2045) onto np.linspace(0, 2045, 2048, dtype=float32).  numpy evaluates that
// grid in float32 as  fl32(fl32(i) * fl32(2045/2047))  with the last point forced to
// 2045, and np.interp then works in float64:  y[j] + (y[j+1]-y[j]) * (x - j).
static void resample_code(const int8_t* chips, double* out) {
    volatile float step = 2045.0f / 2047.0f;
    for (int i = 0; i < GR_N; ++i) {
        volatile float xf = (float)i * step;
        if (i == GR_N - 1) xf = 2045.0f;
        const double x = (double)xf;
        const int j = (int)floor(x);
        if (j >= 2045) { out[i] = (double)chips[1022]; continue; }
        const double y0 = (double)chips[j >> 1], y1 = (double)chips[(j + 1) >> 1];
        out[i] = (y1 - y0) * (x - (double)j) + y0;
    }
}

// in-place iterative radix-2 FFT, double precision (host, init time only)
static void fft_host(std::vector<std::complex<double>>& a) {
    const int n = (int)a.size();
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    const double PI = 3.14159265358979323846;
    for (int len = 2; len <= n; len <<= 1) {
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; ++k) {
                const double ang = -2.0 * PI * (double)k / (double)len;
                const std::complex<double> w(cos(ang), sin(ang));
                const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

extern "C" int gr_init(int device) {
    GrLib* L = gr_lib();
    if (L->ready && L->device == device) return GR_OK;
    if (L->ready && L->live_handles > 0) {
        // plans and banks hold device memory and use the tables of the device they were created on: one GPU per process
        gr_set_error("gr_init: library is bound to device %d with %d live plan/bank handle(s); destroy them first "
                     "(one GPU per process: run one process per GPU)", L->device, L->live_handles);
        return GR_ERR_STATE;
    }
    if (L->ready) gr_shutdown();
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        gr_set_error("gr_init: no CUDA device available (%s); this library has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return GR_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { gr_set_error("gr_init: device %d out of range", device); return GR_ERR_ARG; }
    GR_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GR_CUDA(cudaGetDeviceProperties(&prop, device));
    L->device = device;
    L->num_sms = prop.multiProcessorCount;
    gr_build_host_tables();

    const size_t ncode = (size_t)(GR_MAX_PRN + 1) * GR_N;
    std::vector<float> code_f(ncode, 0.f);
    std::vector<float2> cs_f(ncode, make_float2(0.f, 0.f));
    for (int p = 1; p <= GR_MAX_PRN; ++p)
        for (int i = 0; i < GR_N; ++i) {
            code_f[(size_t)p * GR_N + i] = (float)L->code[p][i];     // exact: values are float32-representable
            cs_f[(size_t)p * GR_N + i] = make_float2((float)L->spec_re[p][i], (float)-L->spec_im[p][i]);
        }
    std::vector<float2> tw1(128 * 16), tw2(8 * 16);
    const double PI = 3.14159265358979323846;
    for (int t = 0; t < 128; ++t)
        for (int k = 0; k < 16; ++k) {
            const double a = -2.0 * PI * (double)(t * k) / 2048.0;
            tw1[t * 16 + k] = make_float2((float)cos(a), (float)sin(a));
        }
    for (int n3 = 0; n3 < 8; ++n3)
        for (int k = 0; k < 16; ++k) {
            const double a = -2.0 * PI * (double)(n3 * k) / 128.0;
            tw2[n3 * 16 + k] = make_float2((float)cos(a), (float)sin(a));
        }
    std::vector<int8_t> chips_h((size_t)(GR_MAX_PRN + 1) * 1024, 0);
    for (int p = 1; p <= GR_MAX_PRN; ++p) memcpy(&chips_h[(size_t)p * 1024], L->chips[p], 1023);
    int8_t* d_chips;
    GR_CUDA(cudaMalloc(&d_chips, chips_h.size()));
    GR_CUDA(cudaMemcpy(d_chips, chips_h.data(), chips_h.size(), cudaMemcpyHostToDevice));
    L->tab.chips = d_chips;
    float *d_code; float2 *d_cs, *d_tw1, *d_tw2;
    GR_CUDA(cudaMalloc(&d_code, ncode * sizeof(float)));
    GR_CUDA(cudaMalloc(&d_cs, ncode * sizeof(float2)));
    GR_CUDA(cudaMalloc(&d_tw1, tw1.size() * sizeof(float2)));
    GR_CUDA(cudaMalloc(&d_tw2, tw2.size() * sizeof(float2)));
    GR_CUDA(cudaMemcpy(d_code, code_f.data(), ncode * sizeof(float), cudaMemcpyHostToDevice));
    GR_CUDA(cudaMemcpy(d_cs, cs_f.data(), ncode * sizeof(float2), cudaMemcpyHostToDevice));
    GR_CUDA(cudaMemcpy(d_tw1, tw1.data(), tw1.size() * sizeof(float2), cudaMemcpyHostToDevice));
    GR_CUDA(cudaMemcpy(d_tw2, tw2.data(), tw2.size() * sizeof(float2), cudaMemcpyHostToDevice));
    L->tab.code = d_code;
    L->tab.conjspec = d_cs;
    L->tab.tw1 = d_tw1;
    L->tab.tw2 = d_tw2;
    L->ready = true;
    return GR_OK;
}

void gr_build_host_tables() {
    GrLib* L = gr_lib();
    if (L->host_tables) return;
    for (int p = 1; p <= GR_MAX_PRN; ++p) {
        gold_chips(p, L->chips[p]);
        resample_code(L->chips[p], L->code[p]);
        std::vector<std::complex<double>> a(GR_N);
        for (int i = 0; i < GR_N; ++i) a[i] = L->code[p][i];
        fft_host(a);
        for (int i = 0; i < GR_N; ++i) { L->spec_re[p][i] = a[i].real(); L->spec_im[p][i] = a[i].imag(); }
    }
    L->host_tables = true;
}

extern "C" int gr_shutdown(void) {
    GrLib* L = gr_lib();
    if (!L->ready) return GR_OK;
    cudaSetDevice(L->device);
    cudaFree((void*)L->tab.code);
    cudaFree((void*)L->tab.conjspec);
    cudaFree((void*)L->tab.tw1);
    cudaFree((void*)L->tab.tw2);
    cudaFree((void*)L->tab.chips);
    L->tab = GrTables{};
    L->ready = false;
    return GR_OK;
}

// The table getters work without a GPU (host tables only): they are the drop-in for
// gpslib.GPSCacode and for the FFT_CACODE list, and let the CPU-only test suite check
// the tables bit for bit.
extern "C" int gr_get_chips(int prn, int8_t* out) {
    if (prn < 1 || prn > GR_MAX_PRN || !out) { gr_set_error("gr_get_chips: bad prn %d", prn); return GR_ERR_ARG; }
    gr_build_host_tables();
    memcpy(out, gr_lib()->chips[prn], 1023);
    return GR_OK;
}
extern "C" int gr_get_cacode(int prn, double* out) {
    if (prn < 1 || prn > GR_MAX_PRN || !out) { gr_set_error("gr_get_cacode: bad prn %d", prn); return GR_ERR_ARG; }
    gr_build_host_tables();
    memcpy(out, gr_lib()->code[prn], GR_N * sizeof(double));
    return GR_OK;
}
extern "C" int gr_get_code_spectrum(int prn, double* out) {
    if (prn < 1 || prn > GR_MAX_PRN || !out) { gr_set_error("gr_get_code_spectrum: bad prn %d", prn); return GR_ERR_ARG; }
    gr_build_host_tables();
    for (int i = 0; i < GR_N; ++i) { out[2 * i] = gr_lib()->spec_re[prn][i]; out[2 * i + 1] = gr_lib()->spec_im[prn][i]; }
    return GR_OK;
}
