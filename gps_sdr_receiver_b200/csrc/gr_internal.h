// Library-internal state shared by the .cu translation units.
#pragma once
#include <stdarg.h>
#include <stdint.h>

#include "../../include/gps_b200.h"
#include "gr_common.cuh"

struct GrLib {
    bool ready = false;
    bool host_tables = false;
    int device = -1;
    int num_sms = 0;
    int live_handles = 0;        // acquisition plans + tracking banks alive (they pin the library to its device)
    GrTables tab{};
    int8_t chips[GR_MAX_PRN + 1][1023];
    double code[GR_MAX_PRN + 1][GR_N];
    double spec_re[GR_MAX_PRN + 1][GR_N];
    double spec_im[GR_MAX_PRN + 1][GR_N];
};

GrLib* gr_lib();
void gr_set_error(const char* fmt, ...);
void gr_build_host_tables();

#define GR_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            gr_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return GR_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define GR_REQUIRE_INIT()                                                       \
    do {                                                                        \
        if (!gr_lib()->ready) {                                                 \
            gr_set_error("library not initialised: call gr_init(device) first"); \
            return GR_ERR_STATE;                                                \
        }                                                                       \
    } while (0)
