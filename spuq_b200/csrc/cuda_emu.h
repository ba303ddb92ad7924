// cuda_emu.h -- host emulation of the tiny CUDA subset spuq_sg.cu uses, for the CPU-only
// development container.  Compiled ONLY into tests/_emu/libspuq_sg_emu.so (g++ -DSPUQ_EMU), which
// the `-m "not gpu"` tests use to exercise the plan builder, the index arithmetic of the kernels
// and the C ABI without a GPU.  It is never part of libspuq_sg.so and never loaded by spuq_b200.
//
// Model: a launch runs every block, and inside a block every thread, sequentially.  That is exact
// for kernels whose threads do not communicate.  Kernels that use __syncthreads / shuffles are
// written with an `#ifdef SPUQ_EMU` serial body instead.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
static inline double __hiloint2double(int hi, int lo) {
    unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
    double d; std::memcpy(&d, &u, 8); return d;
}
static inline double2 make_double2(double a, double b) { return double2{a, b}; }

extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };

static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline int emu_sm_count() { return 4; }

// round-to-nearest products and sums that the compiler may not contract (host: plain operators, -ffp-contract is off for this build)
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
template <typename T> static inline T __ldg(const T* p) { return *p; }

template <typename F>
static inline void emu_launch(dim3 grid, dim3 block, F&& body) {
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx = dim3(bx, by, bz);
        for (unsigned tz = 0; tz < block.z; ++tz)
        for (unsigned ty = 0; ty < block.y; ++ty)
        for (unsigned tx = 0; tx < block.x; ++tx) {
            threadIdx = dim3(tx, ty, tz);
            body();
        }
    }
}
namespace spuq { int& launch_counter(); }
#define SPUQ_LAUNCH(kernel, grid, block, smem, stream, ...) \
    do { ++spuq::launch_counter(); emu_launch((grid), (block), [&]() { kernel(__VA_ARGS__); }); } while (0)
