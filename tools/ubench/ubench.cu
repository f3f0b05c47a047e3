// Microbenchmarks behind DESIGN.md's acquisition-kernel ceiling analysis (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gps_sdr_receiver_b200/csrc tools/ubench/ubench.cu -o gpurun_out/ubench
// Each test prints achieved warp-instructions / cycle / SMSP (from SASS-counted instructions of
// the loop body and the elapsed SM cycles) for a few occupancies.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gr_fft2048.cuh"
#include "gr_fft2048w.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// ---- 1. pure-register radix-16 stage: dft16 + 15 twiddle multiplies, data stays in registers ----
__global__ void __launch_bounds__(128) k_dft16(float* out, const float2* twp, int iters) {
    cf v[16], tw[16];
    for (int i = 0; i < 16; ++i) { v[i] = cf{(float)(threadIdx.x + i), (float)i}; const float2 u = twp[threadIdx.x * 16 + i]; tw[i] = cf{u.x, u.y}; }
    for (int it = 0; it < iters; ++it) {
        dft16(v);
#pragma unroll
        for (int k = 1; k < 16; ++k) v[k] = cmul(v[k], tw[k]);
#pragma unroll
        for (int k = 0; k < 16; ++k) { v[k].x *= 0.25f; v[k].y *= 0.25f; }      // keep magnitudes bounded (32 FMUL)
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += v[i].x + v[i].y;
    if (s == 1.2345f) out[0] = s;
}

// ---- 2. FFMA with three distinct, rotating register operands (no operand reuse) ----
__global__ void __launch_bounds__(128) k_ffma3(float* out, int iters) {
    float r[24];
    for (int i = 0; i < 24; ++i) r[i] = 1.0f + 1e-3f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 24; ++i) r[i] = fmaf(r[(i + 7) % 24], r[(i + 13) % 24], r[i]) * 0.5f;   // FFMA + FMUL
    }
    float s = 0.f;
    for (int i = 0; i < 24; ++i) s += r[i];
    if (s == 1.2345f) out[0] = s;
}
// ---- 3. FADD with two distinct rotating operands ----
__global__ void __launch_bounds__(128) k_fadd2(float* out, int iters) {
    float r[24];
    for (int i = 0; i < 24; ++i) r[i] = 1.0f + 1e-3f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 24; ++i) r[i] = r[(i + 7) % 24] - r[(i + 13) % 24];
    }
    float s = 0.f;
    for (int i = 0; i < 24; ++i) s += r[i];
    if (s == 1.2345f) out[0] = s;
}

// ---- 4. shared-memory exchanges of the gen-3 FFT alone (no math) ----
__global__ void __launch_bounds__(128) k_exch(float* out, int iters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* buf1 = reinterpret_cast<float4*>(smem_raw);
    const int t = threadIdx.x, l = t & 31;
    cf* buf2w = reinterpret_cast<cf*>(smem_raw + 2 * GR_W_BUF1_BYTES) + (t >> 5) * GR_W_WARP_UNITS;
    cf v[16];
    for (int i = 0; i < 16; ++i) v[i] = cf{(float)(t + i), (float)i};
    int par = 0;
    for (int it = 0; it < iters; ++it) {
        float4* b1 = buf1 + par * (GR_W_BUF1_BYTES / 16);
        par ^= 1;
        fftw_ex1_write(b1, t, v);
        __syncthreads();
        fftw_ex1_read(b1, t, v);
        fftw_ex2_write(buf2w, l, v);
        __syncwarp();
        const cf* p = buf2w + GR_W_A * (l & 3) + GR_W_B * (l >> 3) + ((l >> 2) & 1);
#pragma unroll
        for (int n3 = 0; n3 < 8; ++n3) { v[2 * n3] = p[2 * n3]; v[2 * n3 + 1] = p[2 * n3 + 4 * GR_W_B]; }
        __syncwarp();
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += v[i].x + v[i].y;
    if (s == 1.2345f) out[0] = s;
}

// ---- 5. the whole gen-3 FFT (math + exchanges), twiddles in registers ----
__global__ void __launch_bounds__(128) k_fftw(float* out, const float2* tw1p, const float2* tw2p, int iters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* buf1 = reinterpret_cast<float4*>(smem_raw);
    const int t = threadIdx.x, l = t & 31;
    cf* buf2w = reinterpret_cast<cf*>(smem_raw + 2 * GR_W_BUF1_BYTES) + (t >> 5) * GR_W_WARP_UNITS;
    cf v[16], tw1[16], tw2[16];
    for (int i = 0; i < 16; ++i) {
        v[i] = cf{(float)(t + i) * 1e-3f, (float)i * 1e-3f};
        const float2 u = tw1p[t * 16 + i]; tw1[i] = cf{u.x, u.y};
        const float2 w = tw2p[(t & 7) * 16 + i]; tw2[i] = cf{w.x, w.y};
    }
    int par = 0;
    for (int it = 0; it < iters; ++it) {
        dft16(v);
#pragma unroll
        for (int k = 1; k < 16; ++k) v[k] = cmul(v[k], tw1[k]);
        float4* b1 = buf1 + par * (GR_W_BUF1_BYTES / 16);
        par ^= 1;
        fftw_ex1_write(b1, t, v);
        __syncthreads();
        fftw_ex1_read(b1, t, v);
        dft16(v);
#pragma unroll
        for (int k = 1; k < 16; ++k) v[k] = cmul(v[k], tw2[k]);
        fftw_ex2_write(buf2w, l, v);
        __syncwarp();
        fftw_ex2_read_stage3(buf2w, l, v);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k) { v[k].x *= (1.0f / 64.0f); v[k].y *= (1.0f / 64.0f); }
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += v[i].x + v[i].y;
    if (s == 1.2345f) out[0] = s;
}


// ---- 6. warp shuffles alone, and shuffles next to the shared-memory exchange ----
template <bool kWithSmem>
__global__ void __launch_bounds__(128) k_shfl(float* out, int iters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* buf1 = reinterpret_cast<float4*>(smem_raw);
    const int t = threadIdx.x;
    float r[32];
    for (int i = 0; i < 32; ++i) r[i] = (float)(t + i);
    cf v[16];
    for (int i = 0; i < 16; ++i) v[i] = cf{(float)(t + i), (float)i};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __shfl_xor_sync(0xffffffffu, r[i], 1);
        if (kWithSmem) {
            fftw_ex1_write(buf1, t, v);
            __syncthreads();
            fftw_ex1_read(buf1, t, v);
            __syncthreads();
        }
    }
    float s = 0.f;
    for (int i = 0; i < 32; ++i) s += r[i];
    for (int i = 0; i < 16; ++i) s += v[i].x + v[i].y;
    if (s == 1.2345f) out[0] = s;
}

// ---- 7. TMEM loads: 64 columns (x16 four times) per thread and iteration ----
__global__ void __launch_bounds__(128) k_tmem(float* out, int iters) {
    __shared__ uint32_t base_sh;
    const int t = threadIdx.x;
    if (t < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_sh)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = base_sh + ((uint32_t)(32 * (t >> 5)) << 16);
    float w[16];
    for (int i = 0; i < 16; ++i) w[i] = (float)(t + i);
    for (int c = 0; c < 4; ++c)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tm + 16 * c),
                     "f"(w[0]), "f"(w[1]), "f"(w[2]), "f"(w[3]), "f"(w[4]), "f"(w[5]), "f"(w[6]), "f"(w[7]), "f"(w[8]), "f"(w[9]), "f"(w[10]),
                     "f"(w[11]), "f"(w[12]), "f"(w[13]), "f"(w[14]), "f"(w[15]));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                         "tcgen05.wait::ld.sync.aligned;"
                         : "=f"(w[0]), "=f"(w[1]), "=f"(w[2]), "=f"(w[3]), "=f"(w[4]), "=f"(w[5]), "=f"(w[6]), "=f"(w[7]), "=f"(w[8]),
                           "=f"(w[9]), "=f"(w[10]), "=f"(w[11]), "=f"(w[12]), "=f"(w[13]), "=f"(w[14]), "=f"(w[15])
                         : "r"(tm + 16 * c));
            acc += w[0] + w[15];
        }
    }
    if (acc == 1.2345f) out[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_sh), "r"(64));
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
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    int clk_khz; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    const double ghz = clk_khz * 1e-6;
    const int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs, %.3f GHz (nominal max)\n", pr.name, sms, ghz);
    float* d_out; CK(cudaMalloc(&d_out, 16));
    std::vector<float2> tw(128 * 16);
    for (int i = 0; i < 128 * 16; ++i) { double a = -2.0 * 3.14159265358979 * (i / 16) * (i % 16) / 2048.0; tw[i] = make_float2((float)cos(a), (float)sin(a)); }
    float2* d_tw; CK(cudaMalloc(&d_tw, tw.size() * 8)); CK(cudaMemcpy(d_tw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_exch, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_W_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_shfl<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_W_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_shfl<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_W_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_fftw, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_W_SMEM_BYTES));
    const int iters = 2000;
    // (name, instructions per loop iteration per warp [from SASS], launcher)
    for (int cps = 1; cps <= 6; ++cps) {          // CTAs (of 4 warps) per SM = warps per SMSP
        const int grid = sms * cps;
        auto rep = [&](const char* name, double inst_per_iter, float ms) {
            const double cyc = ms * 1e-3 * ghz * 1e9;
            printf("%-10s warps/SMSP=%d  %.3f ms  cycles/iter/warp-slot=%.1f  ipc/SMSP=%.3f\n", name, cps, ms, cyc / iters,
                   inst_per_iter * cps * iters / cyc);
        };
        rep("dft16+tw", 249, time_ms([&] { k_dft16<<<grid, 128>>>(d_out, d_tw, iters); }));
        rep("ffma3", 48.75, time_ms([&] { k_ffma3<<<grid, 128>>>(d_out, iters); }));
        rep("fadd2", 24.75, time_ms([&] { k_fadd2<<<grid, 128>>>(d_out, iters); }));
        rep("shfl32", 35, time_ms([&] { k_shfl<false><<<grid, 128, GR_W_SMEM_BYTES>>>(d_out, iters); }));
        rep("tmem64col", 16, time_ms([&] { k_tmem<<<grid, 128>>>(d_out, iters); }));
        if (cps <= 4) {
            rep("shfl+ex1", 62, time_ms([&] { k_shfl<true><<<grid, 128, GR_W_SMEM_BYTES>>>(d_out, iters); }));
            rep("exch", 57, time_ms([&] { k_exch<<<grid, 128, GR_W_SMEM_BYTES>>>(d_out, iters); }));
            rep("fftw", 620, time_ms([&] { k_fftw<<<grid, 128, GR_W_SMEM_BYTES>>>(d_out, d_tw, d_tw, iters); }));
        }
    }
    return 0;
}
