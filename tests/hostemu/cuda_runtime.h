// Minimal stand-in for <cuda_runtime.h>: lets g++ parse the .cuh sources of libpyflow_b200 so that the bodies of
// the element-wise kernels can be executed on the host, one emulated thread at a time (tests only).
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#define __device__
#define __host__
#define __global__ static inline
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct double2 { double x, y; };
static inline double2 make_double2(double a, double b) { double2 v; v.x = a; v.y = b; return v; }
typedef void* cudaStream_t;
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
static inline void __syncthreads() {}
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
static inline double __shfl_up_sync(unsigned, double v, int) { return v; }
static inline double __shfl_down_sync(unsigned, double v, int) { return v; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
