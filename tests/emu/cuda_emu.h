// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE ONLY (never part of the product library).
//
// A tiny SIMT emulator: just enough of the CUDA execution model (threads, warps, CTAs,
// warp collectives, barriers, atomics, the handful of runtime calls the host side uses) to
// compile tess_b200/csrc/*.cu with g++ and run the *kernel logic* on a CPU.  It exists because
// the build container has no GPU: the warp-level choreography (candidate filters, ballots,
// per-lane moment caches) is debugged here in seconds, under -fsanitize if wanted, and only
// then sent to a B200.  The product (libtess_b200.so) is built by nvcc and has no CPU path;
// nothing under tess_b200/ includes this header unless TESS_EMU is defined by tests/emu/.
//
// Model: one CTA at a time; every CUDA thread is an OS thread; __syncthreads / *_sync
// collectives are std::barrier phases.  Only full-mask (0xffffffff) warp collectives are
// supported, which is also the rule the kernels follow on the GPU.
#pragma once
#include <atomic>
#include <barrier>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) double2 { double x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
static inline int2 make_int2(int a, int b) { int2 r; r.x = a; r.y = b; return r; }
static inline int4 make_int4(int a, int b, int c, int d) { int4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }

namespace emu {
struct ThreadCtx {
    std::barrier<>* cta_bar;
    std::barrier<>* warp_bar;
    uint64_t* xchg;  // 32 slots of this thread's warp
    int lane;
};
inline thread_local ThreadCtx tctx;
inline dim3 g_blockDim, g_gridDim;
}  // namespace emu

inline thread_local uint3 threadIdx, blockIdx;
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

// ------------------------------------------------------------------ barriers & collectives
static inline void __syncthreads() { emu::tctx.cta_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { assert(mask == 0xffffffffu); emu::tctx.warp_bar->arrive_and_wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <class T>
static inline T emu_exchange(T v, int src) {
    static_assert(sizeof(T) <= 8, "emu shuffle supports <= 64-bit types");
    auto& c = emu::tctx;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    c.xchg[c.lane] = raw;
    c.warp_bar->arrive_and_wait();
    uint64_t r = c.xchg[src];
    c.warp_bar->arrive_and_wait();
    T out;
    std::memcpy(&out, &r, sizeof(T));
    return out;
}
template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    assert(mask == 0xffffffffu && width == 32);
    return emu_exchange(v, src & 31);
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int m, int width = 32) {
    assert(mask == 0xffffffffu && width == 32);
    return emu_exchange(v, (emu::tctx.lane ^ m) & 31);
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    assert(mask == 0xffffffffu && width == 32);
    int src = emu::tctx.lane + (int)d;
    return emu_exchange(v, src < 32 ? src : emu::tctx.lane);
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    assert(mask == 0xffffffffu && width == 32);
    int src = emu::tctx.lane - (int)d;
    return emu_exchange(v, src >= 0 ? src : emu::tctx.lane);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    assert(mask == 0xffffffffu);
    auto& c = emu::tctx;
    c.xchg[c.lane] = pred ? 1u : 0u;
    c.warp_bar->arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)(c.xchg[i] & 1u) << i;
    c.warp_bar->arrive_and_wait();
    return r;
}
static inline int __reduce_add_sync(unsigned mask, int v) {
    assert(mask == 0xffffffffu);
    auto& c = emu::tctx;
    c.xchg[c.lane] = (uint64_t)(int64_t)v;
    c.warp_bar->arrive_and_wait();
    int r = 0;
    for (int i = 0; i < 32; ++i) r += (int)(int64_t)c.xchg[i];
    c.warp_bar->arrive_and_wait();
    return r;
}
static inline int __reduce_max_sync(unsigned mask, int v) {
    assert(mask == 0xffffffffu);
    auto& c = emu::tctx;
    c.xchg[c.lane] = (uint64_t)(int64_t)v;
    c.warp_bar->arrive_and_wait();
    int r = (int)(int64_t)c.xchg[0];
    for (int i = 1; i < 32; ++i) { int x = (int)(int64_t)c.xchg[i]; if (x > r) r = x; }
    c.warp_bar->arrive_and_wait();
    return r;
}
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
template <class T>
static inline unsigned __match_any_sync(unsigned mask, T v) {
    assert(mask == 0xffffffffu);
    auto& c = emu::tctx;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    c.xchg[c.lane] = raw;
    c.warp_bar->arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)(c.xchg[i] == raw) << i;
    c.warp_bar->arrive_and_wait();
    return r;
}

// ------------------------------------------------------------------ atomics
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline double atomicAdd(double* p, double v) {
    uint64_t* q = reinterpret_cast<uint64_t*>(p);
    uint64_t old = __atomic_load_n(q, __ATOMIC_SEQ_CST);
    for (;;) {
        double o;
        std::memcpy(&o, &old, 8);
        double n = o + v;
        uint64_t nb;
        std::memcpy(&nb, &n, 8);
        if (__atomic_compare_exchange_n(q, &old, nb, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return o;
    }
}
static inline int atomicSub(int* p, int v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicMax(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline int atomicMin(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

// ------------------------------------------------------------------ intrinsics
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long l) { double r; std::memcpy(&r, &l, 8); return r; }
static inline float __int_as_float(int i) { float r; std::memcpy(&r, &i, 4); return r; }
static inline int __float_as_int(float f) { int r; std::memcpy(&r, &f, 4); return r; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline int __double2hiint(double d) { long long r; std::memcpy(&r, &d, 8); return (int)(r >> 32); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
template <class T>
static inline T __ldg(const T* p) { return *p; }
using std::isfinite;
using std::fmin;
using std::fmax;
using std::fabs;
using std::floor;
using std::sqrt;
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }

// ------------------------------------------------------------------ runtime API subset
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* d) { *d = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 2; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) {
    *p = std::aligned_alloc(256, (n + 255) / 256 * 256 + 256);
    if (*p) std::memset(*p, 0xCD, n);   // poison: catches reads of unset workspace
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }

// ------------------------------------------------------------------ kernel launch
namespace emu {
template <class K, class... A>
void launch(K kern, dim3 grid, dim3 block, A... args) {
    assert(block.x % 32 == 0 && block.y == 1 && block.z == 1 && grid.y == 1 && grid.z == 1);
    g_gridDim = grid;
    g_blockDim = block;
    const int nt = (int)block.x, nw = nt / 32;
    for (unsigned b = 0; b < grid.x; ++b) {
        std::barrier<> cta(nt);
        std::vector<std::unique_ptr<std::barrier<>>> wb;
        for (int w = 0; w < nw; ++w) wb.emplace_back(new std::barrier<>(32));
        std::vector<uint64_t> xchg((size_t)nw * 32, 0);
        std::vector<std::thread> th;
        th.reserve(nt);
        for (int t = 0; t < nt; ++t) {
            th.emplace_back([&, t]() {
                threadIdx = uint3{(unsigned)t, 0, 0};
                blockIdx = uint3{b, 0, 0};
                tctx.cta_bar = &cta;
                tctx.warp_bar = wb[t / 32].get();
                tctx.xchg = xchg.data() + (size_t)(t / 32) * 32;
                tctx.lane = t % 32;
                kern(args...);
                tctx.warp_bar->arrive_and_drop();
                tctx.cta_bar->arrive_and_drop();
            });
        }
        for (auto& x : th) x.join();
    }
}
}  // namespace emu
