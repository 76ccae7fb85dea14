// cuda_emu.h -- TEST-ONLY shim that lets the product's CUDA sources (s2k_kernels.cuh, s2k_api.cu) be
// compiled by g++ and executed on the host, one fiber (or std::thread, see below) per CUDA thread, with barriers for
// __syncthreads and lock-step warp collectives.  It exists so that the kernel LOGIC (tile stitching, halos,
// look-back, tail rules, edge cases) can be checked against the oracle in the CPU-only test tier, where no
// GPU is available.  It is never part of the product: rust-seq2kminmers_b200 loads libs2k_b200.so (nvcc,
// sm_100a) and nothing else; this header is only ever included when tests/emu/build_emu.sh defines S2K_EMU.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <sched.h>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__
#define __align__(n) alignas(n)

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct alignas(16) ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };

// Two execution models.  Default: FIBERS -- every emulated CUDA thread is a ucontext fiber, the fibers of a CTA share
// one OS thread (CTAs of a concurrent launch get one OS thread each) and switch at barriers and warp collectives.
// Hundreds of OS threads meeting at futex barriers cost minutes of system time per test on a small VM; fibers switch
// in user space.  Under AddressSanitizer (tests/test_emu_asan.py), or with -DS2K_EMU_THREADS, every CUDA thread is a
// real std::thread instead (ASan does not follow swapcontext without annotations).
#if !defined(__SANITIZE_ADDRESS__) && !defined(S2K_EMU_THREADS)
#define S2K_EMU_FIBERS 1
#include <ucontext.h>
#endif

namespace emu {
#ifdef S2K_EMU_FIBERS
struct Bar { unsigned n = 0, count = 0, gen = 0; };
struct Block {
    Bar bar;
    std::vector<Bar> wbar;
    std::vector<uint64_t> slot;      // 32 per warp
    uint8_t *smem = nullptr;
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::vector<char *> stacks;
    std::vector<char> done;
    unsigned cur = 0;
    const std::function<void()> *body = nullptr;
};
#else
struct Block {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<std::barrier<>>> wbar;
    std::vector<uint64_t> slot;      // 32 per warp
    uint8_t *smem = nullptr;
};
#endif
inline thread_local Block *blk = nullptr;
inline thread_local emu_dim3 tIdx, bIdx, bDim, gDim;

#ifdef S2K_EMU_FIBERS
inline void fiber_yield() { Block *B = blk; swapcontext(&B->ctx[B->cur], &B->sched); }
inline void bar_wait(Bar &x)
{
    const unsigned g = x.gen;
    if (++x.count == x.n) { x.count = 0; ++x.gen; }
    else while (x.gen == g) fiber_yield();
}
inline void block_barrier() { bar_wait(blk->bar); }
inline void warp_barrier(unsigned w) { bar_wait(blk->wbar[w]); }
inline void fiber_entry()
{
    Block *B = blk;
    (*B->body)();
    B = blk;
    B->done[B->cur] = 1;
    swapcontext(&B->ctx[B->cur], &B->sched);
}
constexpr size_t FIBER_STACK = 256 * 1024;
// runs one CTA to completion on the calling OS thread
inline void run_block(unsigned b, unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body)
{
    Block B;
    const unsigned nw = (block + 31) / 32;
    B.bar.n = block;
    B.wbar.resize(nw);
    for (unsigned w = 0; w < nw; ++w) B.wbar[w].n = std::min(32u, block - 32 * w);
    B.slot.assign((size_t)nw * 32, 0);
    B.smem = smem_bytes ? (uint8_t *)aligned_alloc(128, (smem_bytes + 127) & ~size_t(127)) : nullptr;
    B.ctx.resize(block); B.stacks.resize(block); B.done.assign(block, 0); B.body = &body;
    blk = &B; bIdx.x = b; bDim.x = block; gDim.x = grid;
    for (unsigned t = 0; t < block; ++t) {
        B.stacks[t] = (char *)malloc(FIBER_STACK);
        getcontext(&B.ctx[t]);
        B.ctx[t].uc_stack.ss_sp = B.stacks[t];
        B.ctx[t].uc_stack.ss_size = FIBER_STACK;
        B.ctx[t].uc_link = &B.sched;
        makecontext(&B.ctx[t], fiber_entry, 0);
    }
    for (unsigned remaining = block; remaining;)
        for (unsigned t = 0; t < block; ++t)
            if (!B.done[t]) {
                B.cur = t; tIdx.x = t;
                swapcontext(&B.sched, &B.ctx[t]);
                if (B.done[t]) --remaining;
            }
    for (char *st : B.stacks) free(st);
    free(B.smem);
    blk = nullptr;
}
// Runs `body` once per emulated CUDA thread.  concurrent=false runs the blocks one after the other (needed for
// kernels that keep static __shared__ state, which the shim maps to plain statics).
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, bool concurrent, const std::function<void()> &body)
{
    if (concurrent && grid > 1) {
        std::vector<std::thread> threads;
        for (unsigned b = 0; b < grid; ++b) threads.emplace_back([&, b]() { run_block(b, grid, block, smem_bytes, body); });
        for (auto &t : threads) t.join();
    } else {
        Block *outer = blk;
        for (unsigned b = 0; b < grid; ++b) run_block(b, grid, block, smem_bytes, body);
        blk = outer;
    }
}
#else
inline void block_barrier() { blk->bar->arrive_and_wait(); }
inline void warp_barrier(unsigned w) { blk->wbar[w]->arrive_and_wait(); }
inline void fiber_yield() { sched_yield(); }
// Runs `body` once per emulated CUDA thread.  concurrent=false runs the blocks one after the other (needed for
// kernels that keep static __shared__ state, which the shim maps to plain statics).
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, bool concurrent, const std::function<void()> &body)
{
    auto run_block = [&](unsigned b, std::vector<std::thread> &threads, Block &B) {
        B.bar = std::make_unique<std::barrier<>>(block);
        const unsigned nw = (block + 31) / 32;
        for (unsigned w = 0; w < nw; ++w) B.wbar.push_back(std::make_unique<std::barrier<>>(std::min(32u, block - 32 * w)));
        B.slot.assign((size_t)nw * 32, 0);
        B.smem = smem_bytes ? (uint8_t *)aligned_alloc(128, (smem_bytes + 127) & ~size_t(127)) : nullptr;
        for (unsigned t = 0; t < block; ++t)
            threads.emplace_back([&, b, t]() {
                blk = &B; tIdx.x = t; bIdx.x = b; bDim.x = block; gDim.x = grid;
                body();
            });
    };
    if (concurrent) {
        std::vector<Block> blocks(grid);
        std::vector<std::thread> threads;
        threads.reserve((size_t)grid * block);
        for (unsigned b = 0; b < grid; ++b) run_block(b, threads, blocks[b]);
        for (auto &t : threads) t.join();
        for (auto &B : blocks) free(B.smem);
    } else {
        for (unsigned b = 0; b < grid; ++b) {
            Block B;
            std::vector<std::thread> threads;
            threads.reserve(block);
            run_block(b, threads, B);
            for (auto &t : threads) t.join();
            free(B.smem);
        }
    }
}
#endif
} // namespace emu

#define threadIdx emu::tIdx
#define blockIdx emu::bIdx
#define blockDim emu::bDim
#define gridDim emu::gDim

static inline void __syncthreads() { emu::block_barrier(); }

template <typename T> static inline T emu_warp_exchange(T v, int src_lane_or_neg, bool take)
{
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    emu::blk->slot[w * 32 + lane] = bits;
    emu::warp_barrier(w);
    T r = v;
    if (take && src_lane_or_neg >= 0 && src_lane_or_neg < 32) {
        uint64_t o = emu::blk->slot[w * 32 + (unsigned)src_lane_or_neg];
        std::memcpy(&r, &o, sizeof(T));
    }
    emu::warp_barrier(w);
    return r;
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int o)
{
    const int lane = (int)(threadIdx.x & 31);
    return emu_warp_exchange(v, lane - o, lane >= o);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int o)
{
    const int lane = (int)(threadIdx.x & 31);
    return emu_warp_exchange(v, lane + o, lane + o < 32);
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int o)
{
    const int lane = (int)(threadIdx.x & 31);
    return emu_warp_exchange(v, lane ^ o, true);
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src)
{
    return emu_warp_exchange(v, src & 31, true);
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(threadIdx.x >> 5); }
static inline unsigned __ballot_sync(unsigned, bool pred)
{
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    emu::blk->slot[w * 32 + lane] = pred ? 1 : 0;
    emu::warp_barrier(w);
    unsigned r = 0;
    const unsigned n = std::min(32u, blockDim.x - 32 * w);
    for (unsigned i = 0; i < n; ++i) r |= (unsigned)emu::blk->slot[w * 32 + i] << i;
    emu::warp_barrier(w);
    return r;
}
static inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0; }
static inline unsigned __activemask() { return 0xffffffffu; }     // kernels here run full warps

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __vcmpne4(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) if (((a >> (8 * i)) & 0xff) != ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i);
    return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s)
{
    s &= 31;
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s)
{
    s &= 31;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
    const unsigned long long pool = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned n = (sel >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(pool >> (8 * (n & 7))) & 0xff;
        if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline void __nanosleep(unsigned) { sched_yield(); }     // spin-waits are on other CTAs (own OS threads)
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v)
{
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
static inline unsigned atomicMin(unsigned *p, unsigned v)
{
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned atomicMax(unsigned *p, unsigned v)
{
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v)
{
    unsigned long long cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return cur;
}
template <typename T> static inline T min(T a, T b) { return b < a ? b : a; }
template <typename T> static inline T max(T a, T b) { return a < b ? b : a; }

// ------------------------------------------------------------------------------------------------ runtime API
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaDevAttrMultiProcessorCount = 16, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = aligned_alloc(256, (n + 255) & ~size_t(255)); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *c) { *c = 1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { const char *e = getenv("S2K_EMU_SMS"); *v = e ? atoi(e) : 1; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
