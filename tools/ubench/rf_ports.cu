// Do packed-FP32 math and shared-memory / tensor-memory traffic overlap on one SM sub-partition, or do they share a resource?
//
// The inverse acquisition kernel needs 572 FP32-pipe cycles and ~416 shared-memory-pipe cycles per transform and runs at
// ~911 cycles per transform and SM: close to the SUM, although the two pipes are different units
// (profiles/ubench_r01_inverse_loop_ablation.log: "math only" 499 + "data movement only" 519 against 965 for the full loop).
// This test separates the two: every CTA has 4 "math" warps (one per scheduler: register-only FFMA2 / FADD2 chains on 16
// independent packed accumulators) and 4 "memory" warps (one per scheduler: conflict-free 128-bit shared-memory loads +
// stores, or tensor-memory 32x32b.x16 loads + stores, all into / out of registers).  Times: math alone, memory alone,
// both together.  Independent units: both = max(math, memory).  Shared resource (register-file ports, dispatch): both -> sum.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/rf_ports.bin tools/ubench/rf_ports.cu && tools/ubench/rf_ports.bin
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef unsigned long long cpk;
__device__ __forceinline__ cpk pk(float a, float b) { cpk r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ cpk pfma(cpk a, cpk b, cpk c) { cpk r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ cpk padd(cpk a, cpk b) { cpk r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// MATH: 1 = FFMA2 with three register operands, 2 = FADD2.  MEM: 0 none, 1 LDS.128 + STS.128, 2 LDS.128 only, 3 tcgen05.ld + st
template <int MATH, int MEM>
__global__ void __launch_bounds__(256, 2) k(float* out, int iters, int math_on, int mem_on) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t tm_base;
    const int t = threadIdx.x, warp = t >> 5;
    if (MEM == 3) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&tm_base)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    float sink = 0.f;
    if (warp < 4) {
        if (math_on) {
            cpk a[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = pk(1.0f + t + i, 0.5f * i);
            const cpk m = pk(0.999f, 1.001f), c = pk(1e-3f, -1e-3f);
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int r = 0; r < 18; ++r)           // 18 x 16 = 288 packed instructions per iteration = one transform's math
#pragma unroll
                    for (int i = 0; i < 16; ++i) a[i] = MATH == 1 ? pfma(a[i], m, a[(i + 1) & 15]) : padd(a[i], c);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); sink += x + y; }
        }
    } else if (mem_on) {
        const int u = t - 128;                          // 0..127
        if (MEM == 1 || MEM == 2) {
            float4* buf = reinterpret_cast<float4*>(smem);
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = make_float4(u, i, 1.f, 2.f);
            for (int it = 0; it < iters; ++it) {
                // per iteration and thread: 3 x 8 128-bit loads (X + exchange-1 read + one more 16 KiB) and 8 128-bit stores:
                // 64 KiB through the shared-memory pipe per 128 threads
#pragma unroll
                for (int rep = 0; rep < 3; ++rep) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 w = buf[i * 128 + ((u + it + rep) & 127)];
                        v[i].x += w.x; v[i].y += w.y; v[i].z += w.z; v[i].w += w.w;
                    }
                }
                if (MEM == 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) buf[i * 128 + u] = v[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) sink += v[i].x + v[i].y + v[i].z + v[i].w;
        } else if (MEM == 3) {
            const uint32_t tm = tm_base + ((uint32_t)(32 * (warp & 3)) << 16);
            float r[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = (float)(u + i);
            for (int it = 0; it < iters; ++it) {
                // per iteration and thread: 8 x (16-column store + 16-column load) = 8 KiB each way per warp
#pragma unroll
                for (int rep = 0; rep < 8; ++rep) {
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                                 :: "r"(tm + 16 * (rep & 1)), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),
                                    "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]));
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                        "tcgen05.wait::ld.sync.aligned;\n"
                        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
                          "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
                        : "r"(tm + 16 * (rep & 1)));
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) sink += r[i];
        }
    }
    if (sink == 123.456f) out[t] = sink;
    if (MEM == 3) {
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm_base));
    }
}

template <int MATH, int MEM>
static void run(const char* name, float* out, int sms, double ghz) {
    const int iters = 2000;
    const size_t smem = 16 * 1024;
    cudaFuncSetAttribute(k<MATH, MEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms[3];
    const int cfg[3][2] = {{1, 0}, {0, 1}, {1, 1}};
    for (int c = 0; c < 3; ++c) {
        k<MATH, MEM><<<2 * sms, 256, smem>>>(out, 10, cfg[c][0], cfg[c][1]);
        cudaEventRecord(e0);
        k<MATH, MEM><<<2 * sms, 256, smem>>>(out, iters, cfg[c][0], cfg[c][1]);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms[c], e0, e1);
    }
    const double cyc = ghz * 1e6 / iters;       // cycles per iteration of one warp (2 warps of each kind per scheduler)
    printf("%-44s math alone %7.1f  memory alone %7.1f  both %7.1f  (max %7.1f, sum %7.1f) cycles/iteration\n", name, ms[0] * cyc,
           ms[1] * cyc, ms[2] * cyc, (ms[0] > ms[1] ? ms[0] : ms[1]) * cyc, (ms[0] + ms[1]) * cyc);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("device %s, %d SMs, %.3f GHz; 2 CTAs x (4 math + 4 memory warps) per SM = 2 + 2 warps per scheduler\n", p.name, p.multiProcessorCount, ghz);
    float* out;
    cudaMalloc(&out, 4096);
    run<1, 1>("FFMA2 (3 reg operands) | LDS.128 x24 + STS.128 x8", out, p.multiProcessorCount, ghz);
    run<2, 1>("FADD2 (reg + const)    | LDS.128 x24 + STS.128 x8", out, p.multiProcessorCount, ghz);
    run<1, 2>("FFMA2 (3 reg operands) | LDS.128 x24", out, p.multiProcessorCount, ghz);
    run<1, 3>("FFMA2 (3 reg operands) | TMEM st.x16 + ld.x16  x8", out, p.multiProcessorCount, ghz);
    run<2, 3>("FADD2 (reg + const)    | TMEM st.x16 + ld.x16  x8", out, p.multiProcessorCount, ghz);
    cudaDeviceSynchronize();
    return 0;
}
