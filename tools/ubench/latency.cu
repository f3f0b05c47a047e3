// Unloaded latencies (SM cycles, one CTA of 128 threads alone on an SM) of the data-movement steps of one transform of
// the inverse acquisition kernel: what a warp waits for when no other warp covers it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gps_sdr_receiver_b200/csrc tools/ubench/latency.cu -o tools/ubench/latency.bin
#include <cstdio>
#include <cstdlib>
#include "gr_fft2048t.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)::"memory"); return c; }
__device__ __forceinline__ void ld16(uint32_t taddr, float* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
          "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(taddr));
}

__global__ void __launch_bounds__(128) k_lat(long long* out, int reps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* buf1 = reinterpret_cast<float4*>(smem_raw);
    __shared__ uint32_t tm_base_sh;
    const int t = threadIdx.x;
    if (t < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tm_base_sh)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = tm_base_sh + ((uint32_t)(32 * (t >> 5)) << 16);
    float r[32];
    for (int i = 0; i < 32; ++i) r[i] = (float)(t * 32 + i);
    cpk y[16];
    for (int i = 0; i < 16; ++i) y[i] = cpk_make(r[2 * i], r[2 * i + 1]);
    long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int it = 0; it < reps; ++it) {
        long long c0 = clk();
        tm_round(tm + 32, r);                                   // one TMEM transpose round (2 st, wait, 2 ld, wait)
        long long c1 = clk();
        ld16(tm, r);                                            // one 16-column fetch, issue + wait
        long long c2 = clk();
        fftt_ex1_write_pk(buf1, t, y);
        __syncthreads();
        fftt_ex1_read_pk(buf1, t, y);
        float s = 0.f;
        for (int i = 0; i < 16; ++i) s += cpk_re(y[i]);         // consume the loads
        long long c3 = clk();
        __syncthreads();
        long long c4 = clk();
        float q = s;
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __shfl_xor_sync(0xffffffffu, r[i], 4);   // 16 independent shuffles
        for (int i = 0; i < 16; ++i) q += r[i];
        long long c5 = clk();
#pragma unroll
        for (int m = 0; m < 8; ++m) { const float4 v = buf1[128 * m + t]; q += v.x + v.y + v.z + v.w; }    // 8 LDS.128 + consume
        long long c6 = clk();
        r[31] = q;
        acc[0] += c1 - c0; acc[1] += c2 - c1; acc[2] += c3 - c2; acc[3] += c4 - c3; acc[4] += c5 - c4; acc[5] += c6 - c5;
    }
    if (t == 0) for (int i = 0; i < 6; ++i) out[blockIdx.x * 8 + i] = acc[i];
    if (r[31] == 1.2345f) out[7] = 1;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base_sh), "r"(64));
}

int main() {
    CK(cudaSetDevice(0));
    long long* d; CK(cudaMalloc(&d, 8 * 8 * 1024)); CK(cudaMemset(d, 0, 8 * 8 * 1024));
    CK(cudaFuncSetAttribute(k_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * GR_W_BUF1_BYTES));
    const int reps = 200;
    const char* names[6] = {"TMEM transpose round (2 st, wait, 2 ld, wait)", "TMEM fetch 16 columns (ld + wait)",
                            "exchange 1 (8 STS.128, block barrier, 16 LDS.64, use)", "block barrier alone (4 warps)",
                            "16 independent SHFL + use", "8 LDS.128 + use"};
    for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm += 3) {
        k_lat<<<148 * ctas_per_sm, 128, 2 * GR_W_BUF1_BYTES>>>(d, reps);
        CK(cudaDeviceSynchronize());
        long long h[8];
        CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        printf("%d CTA(s) per SM, all running the same sequence (cycles include ~20 for the clock reads):\n", ctas_per_sm);
        for (int i = 0; i < 6; ++i) printf("  %-58s %7.1f cycles\n", names[i], (double)h[i] / reps);
    }
    return 0;
}
