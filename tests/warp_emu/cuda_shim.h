// TEST HARNESS ONLY — CPU warp emulator.
//
// Compiles blokus-engine_b200/csrc/*.cu with g++ so that the kernel SOURCE and the host-side C ABI
// can be exercised by `pytest -m "not gpu"` on a box without a GPU.  Every CUDA thread of a CTA is
// an OS thread; warp collectives (__shfl_*, __ballot, __reduce_*) exchange values through a
// per-warp mailbox guarded by a std::barrier; CTAs of a launch run one after another.  It is slow
// (thousands of barrier waits per game) and exists only to check LOGIC early.  It is not a CPU
// fallback: the package (blokus-engine_b200/blokus_self_play) only ever loads the sm_100a library,
// and nothing outside tests/ builds or loads this.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define BK_WARP_EMU 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __constant__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint4 { uint32_t x, y, z, w; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct uint2 { uint32_t x, y; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1) : x(a) {} };
struct EmuIdx { int x = 0, y = 0, z = 0; };

struct EmuWarp {
    std::barrier<> bar{32};
    uint64_t slot[2][32];
};
struct EmuCta {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<EmuWarp>> warps;
};

extern thread_local EmuIdx threadIdx, blockIdx, blockDim, gridDim;
extern thread_local EmuWarp* emu_warp;
extern thread_local EmuCta* emu_cta;
extern thread_local int emu_lane;
extern thread_local unsigned emu_phase;

void emu_launch(int grid, int block, const std::function<void()>& body);

template <class T>
inline uint64_t emu_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "emu value too wide"); std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T>
inline T emu_from(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

// every lane posts v; returns the mailbox row (valid until this lane's next-but-one collective)
template <class T>
inline const uint64_t* emu_post(T v) {
    const unsigned ph = emu_phase & 1u;
    emu_phase += 1u;
    emu_warp->slot[ph][emu_lane] = emu_bits(v);
    emu_warp->bar.arrive_and_wait();
    return emu_warp->slot[ph];
}
#define EMU_FULLMASK(m) do { if ((m) != 0xffffffffu) std::abort(); } while (0)

template <class T> inline T __shfl_sync(unsigned m, T v, int src) { EMU_FULLMASK(m); const uint64_t* s = emu_post(v); return emu_from<T>(s[src & 31]); }
template <class T> inline T __shfl_up_sync(unsigned m, T v, int d) { EMU_FULLMASK(m); const uint64_t* s = emu_post(v); int src = emu_lane - d; return emu_from<T>(s[src < 0 ? emu_lane : src]); }
template <class T> inline T __shfl_down_sync(unsigned m, T v, int d) { EMU_FULLMASK(m); const uint64_t* s = emu_post(v); int src = emu_lane + d; return emu_from<T>(s[src > 31 ? emu_lane : src]); }
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int d) { EMU_FULLMASK(m); const uint64_t* s = emu_post(v); return emu_from<T>(s[(emu_lane ^ d) & 31]); }
inline unsigned __ballot_sync(unsigned m, int pred) { EMU_FULLMASK(m); const uint64_t* s = emu_post<uint32_t>(pred ? 1u : 0u); unsigned r = 0; for (int i = 0; i < 32; ++i) r |= unsigned(s[i] & 1u) << i; return r; }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline unsigned __reduce_or_sync(unsigned m, unsigned v) { EMU_FULLMASK(m); const uint64_t* s = emu_post(v); unsigned r = 0; for (int i = 0; i < 32; ++i) r |= unsigned(s[i]); return r; }
inline unsigned __reduce_add_sync(unsigned m, unsigned v) { EMU_FULLMASK(m); const uint64_t* s = emu_post(v); unsigned r = 0; for (int i = 0; i < 32; ++i) r += unsigned(s[i]); return r; }
inline unsigned __reduce_max_sync(unsigned m, unsigned v) { EMU_FULLMASK(m); const uint64_t* s = emu_post(v); unsigned r = 0; for (int i = 0; i < 32; ++i) r = unsigned(s[i]) > r ? unsigned(s[i]) : r; return r; }
inline unsigned __reduce_min_sync(unsigned m, unsigned v) { EMU_FULLMASK(m); const uint64_t* s = emu_post(v); unsigned r = 0xffffffffu; for (int i = 0; i < 32; ++i) r = unsigned(s[i]) < r ? unsigned(s[i]) : r; return r; }
inline void __syncwarp(unsigned m = 0xffffffffu) { EMU_FULLMASK(m); emu_post<uint32_t>(0u); }
inline void __syncthreads() { emu_cta->bar->arrive_and_wait(); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs(int(v)); }
inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline unsigned __umulhi(unsigned a, unsigned b) { return unsigned((uint64_t(a) * b) >> 32); }
inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    if (offset <= 0) std::abort();
    int seen = 0;
    for (unsigned i = base; i < 32; ++i)
        if ((mask >> i) & 1u) { if (++seen == offset) return i; }
    return 0xffffffffu;
}
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }

// IEEE round-to-nearest arithmetic (this TU is built with -ffp-contract=off)
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline long long __double_as_longlong(double d) { long long b; std::memcpy(&b, &d, 8); return b; }
inline double __longlong_as_double(long long b) { double d; std::memcpy(&d, &b, 8); return d; }
inline unsigned __float_as_uint(float f) { unsigned b; std::memcpy(&b, &f, 4); return b; }
inline float __uint_as_float(unsigned b) { float f; std::memcpy(&f, &b, 4); return f; }
inline float __ldg(const float* p) { return *p; }
inline unsigned __ldg(const unsigned* p) { return *p; }

// ---- runtime API stand-ins ---------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
typedef struct EmuEvent { double t; }* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum { cudaStreamNonBlocking = 1 };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetDeviceCount(int* c) { *c = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = static_cast<T*>(std::calloc(1, n ? n : 1)); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<void*>(1); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
