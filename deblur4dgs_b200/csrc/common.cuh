// common.cuh -- shared helpers for the libd4gs.so translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/d4gs.h"

namespace d4 {

void set_error(const char *fmt, ...);

#define D4_CHECK_ARG(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            d4::set_error(__VA_ARGS__);    \
            return 2;                      \
        }                                  \
    } while (0)

#define D4_CHECK_LAUNCH(name)                                                          \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            d4::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

static inline cudaStream_t as_stream(d4_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kAlphaMax = 0.999f;
constexpr float kTMin = 1e-4f;

}  // namespace d4
