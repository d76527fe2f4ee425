/* TEST INFRASTRUCTURE ONLY.  Just enough of the CUDA device vocabulary to compile the body of a
 * simple kernel with g++ and run its threads one after another on the host
 * (tests/test_kernel_emulation.py).  It checks a kernel's index arithmetic, border handling and
 * rounding order against the oracle without a GPU; it says nothing about performance, races or
 * memory ordering.  Barriers are no-ops and __shared__ becomes a function-local static: a kernel
 * that stages data through shared memory is run twice per block (the first pass fills the staging
 * buffers, the second produces the output), which is valid when every thread makes one trip through
 * its loop.  Compile with -ffp-contract=off so that only the explicit fmaf() calls fuse.  Never part
 * of the product library. */
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
struct float4 { float x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct dim3 { unsigned x, y, z; };
static dim3 blockIdx, threadIdx, blockDim, gridDim;
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
struct uint4 { unsigned x, y, z, w; };
#define __shared__ static
static inline void __syncthreads() {}
static inline void __syncwarp() {}
static inline uchar4 make_uchar4(unsigned char a, unsigned char b, unsigned char c, unsigned char d) { return {a, b, c, d}; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
/* x86-64 SSE arithmetic rounds every operation to nearest even in its own precision */
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __frcp_rn(float a) { volatile float r = 1.0f / a; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline int __float2int_rz(float f) { return (int)f; }
static inline unsigned atomicMin(unsigned* p, unsigned v) { unsigned o = *p; if (v < o) *p = v; return o; }
typedef unsigned long long prg_emu_ull;
static inline prg_emu_ull atomicMin(prg_emu_ull* p, prg_emu_ull v) { prg_emu_ull o = *p; if (v < o) *p = v; return o; }
static inline prg_emu_ull atomicCAS(prg_emu_ull* p, prg_emu_ull cmp, prg_emu_ull v) { prg_emu_ull o = *p; if (o == cmp) *p = v; return o; }
static inline prg_emu_ull atomicAdd(prg_emu_ull* p, prg_emu_ull v) { prg_emu_ull o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline int atomicExch(int* p, int v) { int o = *p; *p = v; return o; }
static inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; memcpy(&d, &u, 8); return d; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline long long __double2ll_rn(double a) { return llrint(a); }

/* round 2: packed fp32x2 arithmetic (each lane rounded like the scalar op), F2I.RN with saturation */
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return {__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return {__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return {fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
static inline int __float2int_rn(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return (-2147483647 - 1);
  return (int)nearbyintf(f);            /* default rounding mode: half to even */
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline void __threadfence() {}
static inline void __nanosleep(unsigned) {}
