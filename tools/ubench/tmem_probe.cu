// Probe of the tcgen05.st / tcgen05.ld fragment layouts (one warp): which (thread, register) of a
// 16x256b / 16x128b store lands in which (lane, column) as seen by a 32x32b load.  Used to design the
// warp-local transpose in gr_fft2048t.cuh.   nvcc -gencode arch=compute_100a,code=sm_100a tmem_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

#define LD32(taddr, r) asm volatile( \
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n" \
    "tcgen05.wait::ld.sync.aligned;" \
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), \
      "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), \
      "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr))
#define ST16(shape, taddr, r, o) asm volatile( \
    "tcgen05.st.sync.aligned." shape ".b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
    :: "r"(taddr), "r"(r[o+0]), "r"(r[o+1]), "r"(r[o+2]), "r"(r[o+3]), "r"(r[o+4]), "r"(r[o+5]), "r"(r[o+6]), "r"(r[o+7]), "r"(r[o+8]), "r"(r[o+9]), \
       "r"(r[o+10]), "r"(r[o+11]), "r"(r[o+12]), "r"(r[o+13]), "r"(r[o+14]), "r"(r[o+15]))

template <int MODE>
__global__ void probe(uint32_t* out) {
    __shared__ uint32_t base_sh;
    const int t = threadIdx.x;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_sh)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = base_sh;
    uint32_t r[32], z[32];
    for (int i = 0; i < 32; ++i) { r[i] = 100 * t + i; z[i] = 9999; }
    // clear 64 columns
    ST16("32x32b.x16", tm, z, 0); ST16("32x32b.x16", tm + 16, z, 0); ST16("32x32b.x16", tm + 32, z, 0); ST16("32x32b.x16", tm + 48, z, 0);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (MODE == 0) {            // 16x256b.x4: regs 0..15 -> lanes 0..15, regs 16..31 -> lanes 16..31
        ST16("16x256b.x4", tm, r, 0);
        ST16("16x256b.x4", tm + (16u << 16), r, 16);
    } else if (MODE == 1) {     // 16x128b.x8
        ST16("16x128b.x8", tm, r, 0);
        ST16("16x128b.x8", tm + (16u << 16), r, 16);
    } else {                    // 16x64b.x16
        ST16("16x64b.x16", tm, r, 0);
        ST16("16x64b.x16", tm + (16u << 16), r, 16);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t g[32];
    LD32(tm, g);
    for (int i = 0; i < 32; ++i) out[t * 64 + i] = g[i];
    LD32(tm + 32, g);
    for (int i = 0; i < 32; ++i) out[t * 64 + 32 + i] = g[i];
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_sh), "r"(64));
}

int main() {
    uint32_t* d; CK(cudaMalloc(&d, 32 * 64 * 4));
    uint32_t h[32 * 64];
    const char* names[3] = {"16x256b.x4", "16x128b.x8", "16x64b.x16"};
    for (int mode = 0; mode < 3; ++mode) {
        if (mode == 0) probe<0><<<1, 32>>>(d); else if (mode == 1) probe<1><<<1, 32>>>(d); else probe<2><<<1, 32>>>(d);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        printf("== store %s (value = 100*thread + reg), seen by 32x32b loads: rows = TMEM lane, 64 columns\n", names[mode]);
        for (int lane = 0; lane < 32; ++lane) {
            printf("lane %2d:", lane);
            for (int c = 0; c < 64; ++c) if (h[lane * 64 + c] != 9999) printf(" c%d=%u", c, h[lane * 64 + c]);
            printf("\n");
        }
    }
    return 0;
}
