/* TEST INFRASTRUCTURE ONLY.  Just enough of the CUDA device vocabulary to compile the body of a
 * simple, barrier-free kernel with g++ and run its threads one after another on the host
 * (tests/test_kernel_emulation.py).  It checks a kernel's index arithmetic, border handling and
 * rounding order against the oracle without a GPU; it says nothing about performance or about
 * kernels that use shared memory, shuffles or atomics.  Never part of the product library. */
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(x)
struct float4 { float x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct dim3 { unsigned x, y, z; };
static dim3 blockIdx, threadIdx, blockDim, gridDim;
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
