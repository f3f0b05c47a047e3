// Throughput of the packed FP32 instructions of sm_100a (FADD2 / FMUL2 / FFMA2 = PTX add/mul/fma .f32x2) against
// their scalar forms, and of a radix-16 butterfly written with them (a complex number = one 64-bit register pair;
// ptxas folds the re<->im swap and the per-half sign into operand modifiers .LO_HI / .NP / .PN).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gps_sdr_receiver_b200/csrc tools/ubench/f32x2.cu -o tools/ubench/f32x2.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "gr_fft2048.cuh"
#include "gr_cpk.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// MODE 0: scalar FFMA, 16 chains; 1: FFMA2, 8 chains of pairs; 2: FADD2; 3: FMUL2+FFMA2 with swapped/negated operand
// (a complex multiply); 4: scalar FADD 16 chains; 5: FFMA2 16 chains of pairs (32 registers)
template <int MODE>
__global__ void __launch_bounds__(128) k_pipe(float* out, int iters, float a, float b) {
    float r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = 1.0f + 1e-3f * (threadIdx.x + i);
    cpk p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) p[i] = cpk_make(r[2 * i], r[2 * i + 1]);
    const cpk A = cpk_make(a, b), B = cpk_make(b, a);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], a, b);
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = r[i] + a;
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = cpk_fma(p[i], A, B);
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = cpk_fma(p[i], A, B);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = cpk_add(p[i], A);
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = cpk_cmul(p[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i] + cpk_re(p[i]) + cpk_im(p[i]);
    if (s == 1.2345f) out[0] = s;
}

// radix-16 stage (dft16 + 15 twiddles + rescale): scalar form (as ubench.cu test 1) and packed form
template <bool PACKED>
__global__ void __launch_bounds__(128) k_dft16(float* out, const float2* twp, int iters) {
    cf tw[16];
    for (int i = 0; i < 16; ++i) { const float2 u = twp[threadIdx.x * 16 + i]; tw[i] = cf{u.x, u.y}; }
    float s = 0.f;
    if (!PACKED) {
        cf v[16];
        for (int i = 0; i < 16; ++i) v[i] = cf{(float)(threadIdx.x + i), (float)i};
        for (int it = 0; it < iters; ++it) {
            dft16(v);
#pragma unroll
            for (int k = 1; k < 16; ++k) v[k] = cmul(v[k], tw[k]);
#pragma unroll
            for (int k = 0; k < 16; ++k) { v[k].x *= 0.25f; v[k].y *= 0.25f; }
        }
        for (int i = 0; i < 16; ++i) s += v[i].x + v[i].y;
    } else {
        cpk v[16];
        for (int i = 0; i < 16; ++i) v[i] = cpk_make((float)(threadIdx.x + i), (float)i);
        for (int it = 0; it < iters; ++it) {
            cpk_dft16(v);
#pragma unroll
            for (int k = 1; k < 16; ++k) v[k] = cpk_cmul(v[k], tw[k].x, tw[k].y);
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = cpk_scale(v[k], 0.25f);
        }
        for (int i = 0; i < 16; ++i) s += cpk_re(v[i]) + cpk_im(v[i]);
    }
    if (s == 1.2345f) out[0] = s;
}

// correctness of the packed butterflies against the scalar ones, plain and with twiddled inputs
__global__ void k_check(const float2* in, const float2* twd, float2* o_ref, float2* o_pk, float2* o8_ref, float2* o8_pk,
                        float2* ot_ref, float2* ot_pk, float2* o8t_ref, float2* o8t_pk) {
    cf v[16], tw[16];
    cpk p[16];
    const int T = threadIdx.x;
    for (int i = 0; i < 16; ++i) tw[i] = cf{twd[T * 16 + i].x, twd[T * 16 + i].y};
    auto load = [&](int n) { for (int i = 0; i < n; ++i) { v[i] = cf{in[T * 16 + i].x, in[T * 16 + i].y}; p[i] = cpk_make(v[i].x, v[i].y); } };
    auto store = [&](int n, float2* a, float2* b) { for (int i = 0; i < n; ++i) { a[T * n + i] = make_float2(v[i].x, v[i].y); b[T * n + i] = make_float2(cpk_re(p[i]), cpk_im(p[i])); } };
    load(16); dft16(v); cpk_dft16(p); store(16, o_ref, o_pk);
    load(8); dft8(v); cpk_dft8(p); store(8, o8_ref, o8_pk);
    load(16);
    for (int i = 0; i < 16; ++i) v[i] = cmul(v[i], tw[i]);
    dft16(v);
    {
        float wlo[16], whi[16];
        for (int d = 0; d < 2; ++d)
            for (int m = 0; m < 4; ++m) {
                wlo[8 * d + 2 * m] = tw[d + 4 * m].x; wlo[8 * d + 2 * m + 1] = tw[d + 4 * m].y;
                whi[8 * d + 2 * m] = tw[2 + d + 4 * m].x; whi[8 * d + 2 * m + 1] = tw[2 + d + 4 * m].y;
            }
        cpk_dft16_in_tw<0>(p, wlo);
        cpk_dft16_in_tw<2>(p, whi);
        cpk_dft16_out(p);
    }
    store(16, ot_ref, ot_pk);
    load(8);
    for (int i = 1; i < 8; ++i) v[i] = cmul(v[i], tw[i]);
    dft8(v);
    {
        float w[14];
        for (int n = 1; n < 8; ++n) { w[2 * (n - 1)] = tw[n].x; w[2 * (n - 1) + 1] = tw[n].y; }
        cpk_dft8_tw(p, w);
    }
    store(8, o8t_ref, o8t_pk);
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int clk_khz; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const double ghz = clk_khz * 1e-6;
    const int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs, %.3f GHz (nominal max)\n", pr.name, sms, ghz);
    float* d_out; CK(cudaMalloc(&d_out, 16));
    std::vector<float2> tw(128 * 16);
    for (int i = 0; i < 128 * 16; ++i) { double a = -2.0 * 3.14159265358979 * (i / 16) * (i % 16) / 2048.0; tw[i] = make_float2((float)cos(a), (float)sin(a)); }
    float2* d_tw; CK(cudaMalloc(&d_tw, tw.size() * 8)); CK(cudaMemcpy(d_tw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice));

    {   // correctness
        std::vector<float2> in(128 * 16);
        srand(1);
        for (auto& x : in) x = make_float2((rand() % 2001 - 1000) * 1e-3f, (rand() % 2001 - 1000) * 1e-3f);
        float2* d_in; CK(cudaMalloc(&d_in, in.size() * 8));
        CK(cudaMemcpy(d_in, in.data(), in.size() * 8, cudaMemcpyHostToDevice));
        float2* d_o[8];
        for (auto& q : d_o) CK(cudaMalloc(&q, in.size() * 8));
        k_check<<<1, 128>>>(d_in, d_tw, d_o[0], d_o[1], d_o[2], d_o[3], d_o[4], d_o[5], d_o[6], d_o[7]);
        CK(cudaDeviceSynchronize());
        const char* names[4] = {"dft16", "dft8", "dft16 with twiddled inputs", "dft8 with twiddled inputs"};
        for (int c = 0; c < 4; ++c) {
            const size_t n = (c & 1) ? in.size() / 2 : in.size();
            std::vector<float2> a(n), b(n);
            CK(cudaMemcpy(a.data(), d_o[2 * c], n * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), d_o[2 * c + 1], n * 8, cudaMemcpyDeviceToHost));
            double e = 0, mx = 0; size_t ne = 0;
            for (size_t i = 0; i < n; ++i) {
                e = fmax(e, fmax(fabs(a[i].x - b[i].x), fabs(a[i].y - b[i].y))); ne += (a[i].x != b[i].x) + (a[i].y != b[i].y);
                mx = fmax(mx, fmax(fabs(a[i].x), fabs(a[i].y)));
            }
            printf("packed vs scalar %-28s max |diff| %.3g of max |value| %.3g (%zu of %zu values differ)\n", names[c], e, mx, ne, n * 2);
        }
    }

    const int iters = 4000;
    for (int cps = 1; cps <= 4; ++cps) {
        const int grid = sms * cps;
        auto rep = [&](const char* name, double flop_per_iter_thread, double inst_per_iter, float ms) {
            const double cyc = ms * 1e-3 * ghz * 1e9;
            printf("%-22s warps/SMSP=%d  %.3f ms  inst/clk/SMSP=%.3f  flop/clk/SM=%.1f  (%.1f TFLOP/s)\n", name, cps, ms,
                   inst_per_iter * cps * iters / cyc, flop_per_iter_thread * 128 * cps * iters / cyc,
                   flop_per_iter_thread * 128.0 * grid * iters / (ms * 1e-3) * 1e-12);
        };
        rep("FFMA x16", 32, 16, time_ms([&] { k_pipe<0><<<grid, 128>>>(d_out, iters, 0.999f, 1e-3f); }));
        rep("FADD x16", 16, 16, time_ms([&] { k_pipe<4><<<grid, 128>>>(d_out, iters, 0.999f, 1e-3f); }));
        rep("FFMA2 x8", 32, 8, time_ms([&] { k_pipe<1><<<grid, 128>>>(d_out, iters, 0.999f, 1e-3f); }));
        rep("FFMA2 x16", 64, 16, time_ms([&] { k_pipe<5><<<grid, 128>>>(d_out, iters, 0.999f, 1e-3f); }));
        rep("FADD2 x8", 16, 8, time_ms([&] { k_pipe<2><<<grid, 128>>>(d_out, iters, 0.999f, 1e-3f); }));
        rep("cmul (FMUL2+FFMA2) x8", 48, 16, time_ms([&] { k_pipe<3><<<grid, 128>>>(d_out, iters, 0.6f, 0.8f); }));
        const float s = time_ms([&] { k_dft16<false><<<grid, 128>>>(d_out, d_tw, iters); });
        const float p = time_ms([&] { k_dft16<true><<<grid, 128>>>(d_out, d_tw, iters); });
        const double cyc_s = s * 1e-3 * ghz * 1e9 / iters, cyc_p = p * 1e-3 * ghz * 1e9 / iters;
        printf("radix-16 stage          warps/SMSP=%d  scalar %.1f cycles/iter/warp-slot, packed %.1f  (x%.2f)\n", cps, cyc_s, cyc_p, cyc_s / cyc_p);
    }
    return 0;
}
