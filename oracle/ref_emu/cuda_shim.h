// Host shim that lets g++ compile the reference's kernels3.cu AS IT LIES under /root/reference and run its
// kernels on the CPU (TEST INFRASTRUCTURE: used to pin oracle/ against the reference's own code).
//
// Execution model: one CUDA block at a time, every CUDA thread of the block a real std::thread; `__shared__`
// variables become function-local statics (shared by the threads of the block), `__syncthreads()` a std::barrier.
// Nothing of the reference is copied: emu_main.cpp #includes the .cu from its original location.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local emu_dim3 threadIdx, blockIdx;
static emu_dim3 blockDim, gridDim;
static std::barrier<>* emu_barrier = nullptr;

#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __restrict
#define __syncthreads() emu_barrier->arrive_and_wait()

struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct int4 { int x, y, z, w; };
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline int3 make_int3(int x, int y, int z) { return {x, y, z}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }

// CUDA's global overloads
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline float min(float a, float b) { return fminf(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
static inline double min(double a, double b) { return fmin(a, b); }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline float int2float(int v) { return (float)v; }
static inline float __int2float_rn(int v) { return (float)v; }

static std::mutex emu_atomic_mutex;
static inline double atomicAdd(double* a, double v) { std::lock_guard<std::mutex> g(emu_atomic_mutex); double o = *a; *a = o + v; return o; }
static inline float atomicAdd(float* a, float v) { std::lock_guard<std::mutex> g(emu_atomic_mutex); float o = *a; *a = o + v; return o; }
static inline int atomicAdd(int* a, int v) { std::lock_guard<std::mutex> g(emu_atomic_mutex); int o = *a; *a = o + v; return o; }

// legacy texture reference of the OpenGL path (never executed here)
template <class T, int D> struct texture {};
template <class T> static inline T tex2D(texture<T, 2>, float, float) { return T(); }

// run `body` once per CUDA thread of a (gx, gy) x (bx) launch
template <class F> static void emu_launch(unsigned gx, unsigned gy, unsigned bx, F&& body) {
    gridDim.x = gx; gridDim.y = gy; gridDim.z = 1; blockDim.x = bx; blockDim.y = 1; blockDim.z = 1;
    for (unsigned by_ = 0; by_ < gy; by_++) for (unsigned bx_ = 0; bx_ < gx; bx_++) {
        std::barrier<> bar(bx);
        emu_barrier = &bar;
        std::vector<std::thread> th;
        th.reserve(bx);
        for (unsigned t = 0; t < bx; t++)
            th.emplace_back([&, t]() { threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0; blockIdx.x = bx_; blockIdx.y = by_; blockIdx.z = 0; body(); });
        for (auto& x : th) x.join();
    }
    emu_barrier = nullptr;
}
