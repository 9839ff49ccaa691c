// common.cuh -- shared host/device helpers of libdedflow_b200 (sm_100a only, no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/dedflow_b200.h"

typedef int32_t i32;
typedef int64_t i64;
typedef uint32_t u32;
typedef uint8_t u8;
typedef double f64;

namespace dfb {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define DFB_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      dfb::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,  \
                     cudaGetErrorString(_e));                                                 \
      return DFB_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define DFB_LAUNCH_CHECK()                                                                    \
  do {                                                                                        \
    dfb::count_launch();                                                                      \
    DFB_CUDA(cudaGetLastError());                                                             \
  } while (0)

#define DFB_CHECK(expr)                 \
  do {                                  \
    int _s = (expr);                    \
    if (_s != DFB_OK) return _s;        \
  } while (0)

inline int ceil_div(i64 a, i64 b) { return (int)((a + b - 1) / b); }
inline cudaStream_t as_stream(void* s) { return (cudaStream_t)s; }

// number of SMs of the current device (B200: 148); cached
int num_sms();

// physics / time-integration constants: reference src/assemble.cu:23-40, src/main.c:23-27
constexpr f64 kRHOC = 0.5;
constexpr f64 kDT = 5e-2;
constexpr f64 kALPHAM = (3.0 - kRHOC) / (1.0 + kRHOC);
constexpr f64 kALPHAF = 1.0 / (1.0 + kRHOC);
constexpr f64 kGAMMA = 0.5 + kALPHAM - kALPHAF;
constexpr f64 kRHO = 1.0e3;
constexpr f64 kCP = 1.0;
constexpr f64 kKAPPA = 0.66;
constexpr f64 kMU = 10.0 / 3.0;
// 4-point tet rule, reference src/assemble.cu:43-47
constexpr f64 kGW = 0.0416666666666667;
constexpr f64 kSHA = 0.5854101966249685;
constexpr f64 kSHB = 0.1381966011250105;

}  // namespace dfb
