// ptb_host_shim.h — TEST INFRASTRUCTURE.  Lets g++ compile the product's device header (csrc/ptb_device.cuh) for the host so that the
// traversal state machine, the analytic-light tests and the camera rays — the hit-deciding code, written only with IEEE round-to-nearest
// intrinsics — can be checked against the oracle bit for bit WITHOUT a GPU (tests/test_host_traversal.py).  It is never linked into
// libptb200.so and nothing in the product can reach it; the parity gates proper still run the CUDA build on the GPU.
// Compile with -ffp-contract=off: the __f*_rn intrinsics below then are exactly one IEEE operation each, as on the device.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector_types.h>
#include <vector_functions.h>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __global__
#ifndef __restrict__
#define __restrict__
#endif

// bytes fetched through the read-only path by the calling thread: a work measure for order studies of the any-hit hierarchy (scripts/anyhit_order.py)
static thread_local unsigned long long g_hh_ldg_bytes = 0;
template <class T> static inline T __ldg(const T* p) { g_hh_ldg_bytes += sizeof(T); return *p; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline float __uint2float_rn(uint32_t x) { return (float)x; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
#include <algorithm>
using std::min; using std::max;
// sincosf: glibc's (declared by <cmath> with _GNU_SOURCE, which g++ defines)
#define CUDART_INF_F (__builtin_inff())
static inline int __ffs(int x) { return __builtin_ffs(x); }
