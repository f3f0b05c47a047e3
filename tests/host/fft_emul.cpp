// Host emulation of the CTA-level FFT-2048 (gr_fft2048.cuh): runs the three
// stages "thread by thread" with ordinary arrays standing in for shared memory,
// so the index/twiddle mapping is verified on a machine without a GPU.
// Reads 2048 complex64 from stdin (binary), writes 2048 complex64 to stdout.
#include <cstdio>
#include <cmath>
#include <vector>
#include "../../gps_sdr_receiver_b200/csrc/gr_fft2048.cuh"

int main() {
    std::vector<cf> x(2048), out(2048);
    if (fread(x.data(), sizeof(cf), 2048, stdin) != 2048) return 1;
    std::vector<cf> buf1(GR_B1_ELEMS), buf2(GR_B2_ELEMS);
    std::vector<cf> regs(128 * 16);
    const double PI = 3.14159265358979323846;
    for (int t = 0; t < 128; ++t) {
        cf* v = &regs[t * 16];
        cf tw1[16];
        for (int j = 0; j < 16; ++j) {
            v[j] = x[t + 128 * j];
            double a = -2.0 * PI * (double)(t * j) / 2048.0;
            tw1[j] = cf{(float)cos(a), (float)sin(a)};
        }
        fft_stage1(v, tw1);
        fft_ex1_write(buf1.data(), t, v);
    }
    for (int t = 0; t < 128; ++t) {
        cf* v = &regs[t * 16];
        cf tw2[16];
        for (int j = 0; j < 16; ++j) {
            double a = -2.0 * PI * (double)((t & 7) * j) / 128.0;
            tw2[j] = cf{(float)cos(a), (float)sin(a)};
        }
        fft_ex1_read(buf1.data(), t, v);
        fft_stage2(v, tw2);
        fft_ex2_write(buf2.data(), t, v);
    }
    for (int t = 0; t < 128; ++t) {
        cf* v = &regs[t * 16];
        fft_ex2_read_stage3(buf2.data(), t, v);
        for (int j = 0; j < 16; ++j) out[t + 128 * j] = v[j];
    }
    fwrite(out.data(), sizeof(cf), 2048, stdout);
    return 0;
}
