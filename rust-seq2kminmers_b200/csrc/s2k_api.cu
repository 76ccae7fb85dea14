// s2k_api.cu -- C ABI (include/seq2kminmers.h) over the sm_100a kernels in s2k_kernels.cuh.
// Host side of the drop-in boundary: parameter validation mirroring the reference's panics, the exact
// float recipe for the selection bounds, table construction, buffer management and launch sequencing.
// There is no CPU fallback: every entry point that computes anything launches CUDA kernels or fails.
#include "../../include/seq2kminmers.h"
#include "s2k_kernels.cuh"
#include "s2k_count.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace s2k;

// Build-time experiments, both bit-exact and both measured SLOWER than k_minimizers on B200:

// Kernel launch.  `conc` only matters to the test-tier host emulation (tests/emu): kernels with static shared
// state run their blocks one after the other there.
#ifdef S2K_EMU
#define S2K_LAUNCH(kfn, grid, block, smem, stream, conc, ...) \
    emu::launch((unsigned)(grid), (unsigned)(block), (size_t)(smem), conc, [&]() { kfn(__VA_ARGS__); })
#else
#define S2K_LAUNCH(kfn, grid, block, smem, stream, conc, ...) kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

namespace {

// ------------------------------------------------------------------------------------------------ bounds
// src/lib.rs:91: ((density as f64) * (u32::MAX as f64)) as u32   (Rust `as` saturates, NaN -> 0)
uint32_t bound_scalar(double density)
{
    const double v = density * 4294967295.0;
    if (!(v > 0.0)) return 0u;
    if (v >= 4294967295.0) return 0xffffffffu;
    return (uint32_t)v;
}
// src/nthash_avx512_32.rs:47-48: density = bound/u32::MAX (f64); ((density as f32) * (u32::MAX as f32)) as u32
uint32_t bound_simd(uint32_t b)
{
    const double dd = (double)b / 4294967295.0;
    const float f = (float)dd;
    volatile float prod = f * 4294967296.0f;
    if (!(prod > 0.0f)) return 0u;
    if (prod >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)prod;
}

const uint64_t SEED64[4] = {0x3c8bfbb395c60474ull, 0x3193c18562a02b4cull, 0x20323ed082572324ull,
                            0x295549f54be24456ull};   // A C G T, src/nthash_hpc.rs:31-34

uint32_t rolw(uint32_t x, unsigned r, int w)
{
    r %= (unsigned)w;
    if (!r) return x;
    const uint32_t m = w == 32 ? 0xffffffffu : 0x7fffffffu;
    return ((x << r) | (x >> (w - r))) & m;
}
uint32_t rorw(uint32_t x, unsigned r, int w) { r %= (unsigned)w; return rolw(x, (unsigned)w - r, w); }

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    bool host = false;
};

struct Timing {
    bool enabled = false;
    cudaEvent_t ev[64][2];
    int n = 0;           // minimizer-kernel launches recorded
    cudaEvent_t wv[2];   // window-kernel
    bool created = false;
    double min_ms = 0, win_ms = 0;
    uint32_t min_launches = 0;
};

} // namespace

struct s2k_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t flags = 0;
    std::string err;
    uint64_t launches = 0;
    int sm_count = 148;
    double rate_hint = 0.0;         // observed minimizers per base (grow-only)
    // device buffers
    Buf d_bases, d_seq_off, d_tile_lb, d_status, d_small, d_mins, d_min_off, d_min_loc, d_tile_pre, d_hpc_off, d_km_off, d_min_cnt;
    Buf d_hash, d_start, d_end, d_rev, d_rle_hpc, d_rle_pos, d_hscr, d_tmp, d_tile_info, d_tile_base, d_tile_src;
    Buf d_tmp_hi, d_mins_hi;           // H = u64 flavour: high halves of the minimizer hashes
    bool min_hi_valid = false;
    Buf d_ct_keys, d_ct_cnt, d_ct_first, d_ct_side, d_co_hash, d_co_cnt, d_co_first, h_ct_side;   // s2k_count_device
    // pinned host result buffers
    Buf h_hash, h_start, h_end, h_rev, h_km_off, h_mins, h_min_off, h_min_cnt, h_small, h_rle_hpc, h_rle_pos;
    Timing tm;
    bool attr_set = false;
    // pipelined host path (s2k_run on large batches)
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_in[3] = {nullptr, nullptr, nullptr}, ev_free[3] = {nullptr, nullptr, nullptr}, ev_out = nullptr, ev_done = nullptr;
    bool pipe_ready = false;
    uint64_t slab_bytes = 0;        // 0 = default
    Buf d_in[3], d_in_off[3], h_off_stage[3];   // input slabs: one being computed, two in flight behind it
    Buf d_piece, h_piece;           // long sequences in pieces: kept-base counter and cut results
    Buf d_stage;                    // a slab's results parked for the D2H while the next slab is computed
    Buf d_pack[3], h_pack[3];       // 2-bit transport: device landing buffers, ring of pinned staging buffers
    cudaEvent_t ev_pack[3] = {nullptr, nullptr, nullptr};
    int host_threads = 0;           // 0 = 3/4 of the hardware threads, at most 16 (measured best on a 16-core host)
    double pack_ratio = 0.7;        // share of slabs that travel packed (the rest keep PCIe busy with plain ASCII)
    uint64_t tr_h2d_bytes = 0, tr_packed = 0, tr_plain = 0;   // last s2k_run: bytes copied to the device, slabs by kind
    Buf h_fx_bases, h_fx_off;       // s2k_run_fastx: parsed file in pinned memory
    Buf h_ascii[3];                 // ... slabs that cannot travel packed, gathered as ASCII
    uint64_t fx_n_seqs = 0, fx_n_bases = 0;
    bool fx_have_bases = false;     // the last s2k_run_fastx materialised the parsed bases (small file, or S2K_FASTX_KEEP_BASES)
};

namespace {

int fail(s2k_ctx *c, int code, const std::string &msg)
{
    if (c) c->err = msg;
    return code;
}
#define CU_PLACEHOLDER
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? S2K_ERR_OOM : S2K_ERR_CUDA,          \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                        \
    } while (0)

int ensure(s2k_ctx *ctx, Buf &b, size_t bytes, bool host)
{
    if (bytes <= b.cap && b.p) return S2K_OK;
    if (b.p) {
        if (b.host) cudaFreeHost(b.p); else cudaFree(b.p);
        b.p = nullptr; b.cap = 0;
    }
    size_t want = std::max<size_t>(bytes + bytes / 8, 256);
    want = (want + 255) & ~size_t(255);
    cudaError_t e = host ? cudaMallocHost(&b.p, want) : cudaMalloc(&b.p, want);
    if (e != cudaSuccess && want > bytes) {          // retry without slack
        (void)cudaGetLastError();
        want = (std::max<size_t>(bytes, 256) + 255) & ~size_t(255);
        e = host ? cudaMallocHost(&b.p, want) : cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        b.p = nullptr;
        return fail(ctx, S2K_ERR_OOM, std::string("allocation of ") + std::to_string(want) + " bytes failed: " +
                                          cudaGetErrorString(e));
    }
    b.cap = want; b.host = host;
    return S2K_OK;
}
// Grow a pinned host buffer, keeping its first `keep` bytes.
int ensure_keep(s2k_ctx *ctx, Buf &b, size_t bytes, size_t keep)
{
    if (bytes <= b.cap && b.p) return S2K_OK;
    Buf nb;
    int rc = ensure(ctx, nb, bytes + bytes / 4, true);
    if (rc != S2K_OK) return rc;
    if (b.p && keep) std::memcpy(nb.p, b.p, keep);
    if (b.p) cudaFreeHost(b.p);
    b = nb;
    return S2K_OK;
}
void release(Buf &b)
{
    if (b.p) { if (b.host) cudaFreeHost(b.p); else cudaFree(b.p); }
    b.p = nullptr; b.cap = 0;
}

} // namespace
#include "s2k_fastx.inc"
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <atomic>
#include <chrono>
#include <thread>
namespace {

// ---- 2-bit transport: host side.  Packs n bases (4 per byte, code (b>>1)&3: A0 C1 T2 G3) and counts the bytes that
// are not upper-case A/C/G/T; a slab with any such byte travels as plain ASCII instead, so nothing is ever lost.
bool host_has_avx512()
{
#if defined(__x86_64__)
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
#else
    return false;
#endif
}
#if defined(__x86_64__)
__attribute__((target("avx512f,avx512bw,avx512vl")))
uint64_t pack2_avx512(const uint8_t *src, uint8_t *dst, size_t n)      // n multiple of 64
{
    const __m512i three = _mm512_set1_epi8(3);
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i m1 = _mm512_set1_epi16(0x0401), m2 = _mm512_set1_epi32(0x00100001);
    __mmask64 bad = 0;
    // dst is 16-byte aligned here (slab staging buffers, 64-base granules): streaming stores keep the packed bytes out
    // of the caches and spare the read-for-ownership -- the packers are bound by host memory traffic, not by arithmetic
    const bool nt = ((uintptr_t)dst & 15) == 0;
    for (size_t i = 0; i < n; i += 64) {
        const __m512i v = _mm512_loadu_si512(src + i);
        const __m512i c = _mm512_and_si512(_mm512_srli_epi16(v, 1), three);
        bad |= _mm512_cmpneq_epi8_mask(_mm512_shuffle_epi8(lut, c), v);
        const __m512i p32 = _mm512_madd_epi16(_mm512_maddubs_epi16(c, m1), m2);     // c0 + 4 c1 + 16 c2 + 64 c3 per 4 bases
        const __m128i o = _mm512_cvtepi32_epi8(p32);
        if (nt) _mm_stream_si128((__m128i *)(dst + (i >> 2)), o);
        else _mm_storeu_si128((__m128i *)(dst + (i >> 2)), o);
    }
    if (nt) _mm_sfence();
    return bad != 0;
}
#endif
uint64_t pack2_range(const uint8_t *src, uint8_t *dst, size_t lo, size_t hi)   // lo multiple of 64; packs bases [lo, hi)
{
    uint64_t bad = 0;
    size_t i = lo;
#if defined(__x86_64__)
    if (host_has_avx512()) {
        const size_t body = (hi - lo) & ~size_t(63);
        bad += pack2_avx512(src + lo, dst + (lo >> 2), body);
        i = lo + body;
    }
#endif
    for (; i < hi; i += 4) {
        uint8_t o = 0;
        for (size_t j = 0; j < 4 && i + j < hi; ++j) {
            const uint8_t b = src[i + j], c = (b >> 1) & 3;
            bad += "ACTG"[c] != b;
            o |= (uint8_t)(c << (2 * j));
        }
        dst[i >> 2] = o;
    }
    return bad;
}

} // namespace
namespace {

struct Plan {
    bool hpc, simd, w31, quirk;
    bool h64 = false;            // H = u64 flavour (S2K_HASH_NT1_64): 64-bit ntHash1, `hash <= bound` on 64 bits
    uint64_t thr64 = 0;
    bool h16 = false;            // H = u16 flavour (S2K_HASH_NT1_16): see make_plan
    bool trunc16 = false;        //   ... in mode Regular: 32-bit hash truncated before the test
    uint4 xy64[XYN];
    uint32_t l, k, d, need, thr, halo, tile;
    bool none;       // threshold selects nothing
    bool dense;      // >= ~0.5 % of the owners selected: bookkeeping per owner instead of a test per group of four
    uint8_t lut[256];   // raw byte -> code of its base class
    uint2 xy[XYN], xf[XFN], x2[XFN];
};

// Validation mirrors the reference's panics: assert!(k<=31) (src/nthash_avx512_32.rs:33) for the SIMD modes,
// KSizeTooBig / assert!(k<256) (src/nthash_hpc.rs:123-133) for the scalar ones.
int make_plan(s2k_ctx *ctx, const s2k_params *p, Plan &P)
{
    if (!p) return fail(ctx, S2K_ERR_NULL, "params is null");
    if (p->mode < S2K_MODE_REGULAR || p->mode > S2K_MODE_HPCSIMD) return fail(ctx, S2K_ERR_BAD_PARAM, "unknown hash mode");
    if (p->variant != S2K_HASH_NT1_32 && p->variant != S2K_HASH_NT2_31 && p->variant != S2K_HASH_NT1_64 &&
        p->variant != S2K_HASH_NT1_16)
        return fail(ctx, S2K_ERR_BAD_PARAM, "unknown hash variant");
    if (p->l == 0 || p->k == 0) return fail(ctx, S2K_ERR_BAD_PARAM, "l and k must be >= 1");
    P.hpc = p->mode == S2K_MODE_HPC || p->mode == S2K_MODE_HPCSIMD;
    P.simd = p->mode == S2K_MODE_SIMD || p->mode == S2K_MODE_HPCSIMD;
    P.w31 = p->variant == S2K_HASH_NT2_31;
    if (P.w31 && !P.simd) return fail(ctx, S2K_ERR_BAD_PARAM, "the 31-bit variant replaces the SIMD iterator only (modes Simd/HpcSimd)");
    P.h64 = p->variant == S2K_HASH_NT1_64;
    if (P.h64 && P.simd) return fail(ctx, S2K_ERR_BAD_PARAM, "H = u64 exists for the scalar iterators only (modes Regular/Hpc): the SIMD iterators are 32-bit (src/nthash_avx512_32.rs)");
    P.h16 = p->variant == S2K_HASH_NT1_16;
    if (P.h16 && P.simd) return fail(ctx, S2K_ERR_BAD_PARAM, "H = u16 exists for the scalar iterators only (modes Regular/Hpc): the SIMD iterators take a u32 bound (src/nthash_avx512_32.rs:32)");
    P.trunc16 = P.h16 && !P.hpc;
    if (P.simd && p->l > 31) return fail(ctx, S2K_ERR_L_TOO_BIG, "l must be <= 31 in the Simd/HpcSimd modes");
    if (!P.simd && p->l >= 256) return fail(ctx, S2K_ERR_L_TOO_BIG, "l must be < 256 in the Regular/Hpc modes");
    if (p->k > 0x7fffffffu) return fail(ctx, S2K_ERR_BAD_PARAM, "k too large");
    P.l = p->l; P.k = p->k;
    P.d = (p->mode == S2K_MODE_HPC) ? 1u : 0u;       // Hpc: the l-mer is emitted when the NEXT kept base shows up
    P.need = P.l - 1 + P.d;
    P.quirk = P.simd && !P.w31 && !(ctx && (ctx->flags & S2K_NO_TAIL_RULE));
    P.halo = P.need >= 128 ? 512u : 256u;
    P.tile = P.hpc ? (uint32_t)WIN - P.halo : std::min<uint32_t>((uint32_t)CAP, (uint32_t)WIN - P.halo);
    const uint32_t bs = bound_scalar(p->density);
    uint64_t excl;                                   // select iff hash < excl
    if (P.simd) { uint32_t b = bound_simd(bs); if (P.w31) b /= 2; excl = b; }
    else excl = (uint64_t)bs + 1;
    P.none = excl == 0;
    P.thr = P.none ? 0u : (uint32_t)(excl - 1);
    P.dense = (double)excl / (P.w31 ? 2147483648.0 : 4294967296.0) >= 0.0027;   // the canonical minimum doubles the rate
    if (P.h64) {                                     // src/lib.rs:91 with H = u64: ((density as f64) * (u64::MAX as f64)) as u64
        P.thr64 = s2k_bound_u64(p->density);
        P.none = false;                              // `hash <= 0` still selects a hash of 0
        P.dense = (double)P.thr64 / 18446744073709551616.0 >= 0.0027;
    }
    if (P.h16) {
        // src/lib.rs:91 with H = u16.  Mode Hpc (src/nthash_hpc.rs with 16-bit seeds and rotations): the state lives
        // DUPLICATED in both halves of a 32-bit word -- a 32-bit rotation of x * 0x10001 is the 16-bit rotation of x,
        // duplicated; XOR keeps the form; x -> x * 0x10001 is increasing, so min and `<=` agree with the 16-bit ones.
        // The 32-bit kernels run unchanged on duplicated seeds and a duplicated bound; k_finalize masks the records.
        // Mode Regular (`hash = x as H`, src/lib.rs:224): 32-bit tables, truncation inside the byte-form hash stage.
        const uint32_t b16 = s2k_bound_u16(p->density);
        P.thr = P.trunc16 ? b16 : b16 * 0x10001u;
        P.none = false;
        P.dense = ((double)b16 + 1.0) / 65536.0 >= (P.trunc16 ? 0.0054 : 0.0027);   // truncation: no doubling by the minimum
    }
    // base classes: 0..3 = A C T G (bits 1-2 of the ASCII byte: the packed compaction classifies with one AND),
    // 4 = seed 0, 5 = seed 1.  SEED64 is in the order A C G T; the complement of seed s is seed 3 - s.
    const int w = P.w31 ? 31 : 32;
    uint32_t h[8] = {0}, rc[8] = {0};
    static const int seed_of[4] = {0, 1, 3, 2};
    for (int b = 0; b < 4; ++b) {
        h[b] = P.w31 ? (uint32_t)(SEED64[seed_of[b]] >> 33) : (uint32_t)SEED64[seed_of[b]];
        rc[b] = P.w31 ? (uint32_t)(SEED64[3 - seed_of[b]] >> 33) : (uint32_t)SEED64[3 - seed_of[b]];
    }
    h[5] = rc[5] = 1;
    if (P.h16 && !P.trunc16) for (int b = 0; b < 6; ++b) { h[b] = (h[b] & 0xffffu) * 0x10001u; rc[b] = (rc[b] & 0xffffu) * 0x10001u; }
    static const uint8_t code_of[6] = {0, 8, 16, 24, 32, 40};     // see s2k_kernels.cuh (table layout)
    if (P.simd) {                                    // low nibble, src/nthash_avx512_32.rs:178-193
        static const uint8_t nib[16] = {4, 0, 4, 1, 2, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4};
        for (int i = 0; i < 256; ++i) P.lut[i] = code_of[nib[i & 15]];
    } else {                                         // src/nthash_hpc.rs:29-49
        for (int i = 0; i < 256; ++i) P.lut[i] = code_of[5];
        P.lut['A'] = code_of[0]; P.lut['C'] = code_of[1]; P.lut['T'] = code_of[2]; P.lut['G'] = code_of[3];
        P.lut['N'] = code_of[4];
    }
    std::memset(P.xy, 0, sizeof(P.xy));
    std::memset(P.xf, 0, sizeof(P.xf));
    std::memset(P.x2, 0, sizeof(P.x2));
    for (int o = 0; o < 6; ++o)
        for (int i = 0; i < 6; ++i) {
            const uint2 e = make_uint2(rolw(h[o], P.l, w) ^ h[i], rorw(rc[o], 1, w) ^ rolw(rc[i], P.l - 1, w));
            P.xy[(8 * code_of[o] + code_of[i]) / 8] = e;
            if (o < 4 && i < 4) {
                P.xf[(4 * code_of[o] + code_of[i]) / 8] = e;
                // two warm-up steps at once: nothing leaves, bases o then i enter
                P.x2[(4 * code_of[o] + code_of[i]) / 8] =
                    make_uint2(rolw(h[o], 1, w) ^ h[i], rorw(rolw(rc[o], P.l - 1, w), 1, w) ^ rolw(rc[i], P.l - 1, w));
            }
        }
    if (P.h64) {                                     // 64-bit seeds (src/nthash_hpc.rs:30-49), scalar base map: N -> 0, other -> 1
        uint64_t h6[8] = {0}, r6[8] = {0};
        for (int b = 0; b < 4; ++b) { h6[b] = SEED64[seed_of[b]]; r6[b] = SEED64[3 - seed_of[b]]; }
        h6[5] = r6[5] = 1;
        auto rol = [](uint64_t x, unsigned r) { r &= 63u; return r ? (x << r) | (x >> (64u - r)) : x; };
        std::memset(P.xy64, 0, sizeof(P.xy64));
        for (int o = 0; o < 6; ++o)
            for (int i = 0; i < 6; ++i) {
                const uint64_t f = rol(h6[o], P.l) ^ h6[i], r = rol(r6[o], 63) ^ rol(r6[i], P.l - 1);
                P.xy64[(8 * code_of[o] + code_of[i]) / 8] = make_uint4((uint32_t)f, (uint32_t)(f >> 32), (uint32_t)r, (uint32_t)(r >> 32));
            }
    }
    return S2K_OK;
}

// Fraction of the hash space at or below the bound (capacity of the minimizer stream, overlap of the pieces of a long
// sequence): the bound lives on 64 bits (H = u64), 16 bits (H = u16, mode Regular) or w bits.
double selection_fraction(const Plan &P)
{
    if (P.h64) return std::min(1.0, ((double)P.thr64 + 1.0) / 18446744073709551616.0);
    if (P.trunc16) return std::min(1.0, ((double)P.thr + 1.0) / 65536.0);
    return std::min(1.0, ((double)P.thr + 1.0) / (P.w31 ? 2147483648.0 : 4294967296.0));
}

template <typename T> T *ptr(Buf &b) { return reinterpret_cast<T *>(b.p); }

typedef void (*MinimizerKernel)(const K1Args);
MinimizerKernel minimizer_kernel(bool hpc, bool w31, bool dense, bool h64 = false)
{
    static const MinimizerKernel k[8] = {
        k_minimizers<false, false, false>, k_minimizers<true, false, false>, k_minimizers<false, true, false>, k_minimizers<true, true, false>,
        k_minimizers<false, false, true>,  k_minimizers<true, false, true>,  k_minimizers<false, true, true>,  k_minimizers<true, true, true>};
    static const MinimizerKernel k64[4] = {
        k_minimizers<false, false, false, true>, k_minimizers<true, false, false, true>,
        k_minimizers<false, false, true, true>,  k_minimizers<true, false, true, true>};
    if (h64) return k64[(hpc ? 1 : 0) | (dense ? 2 : 0)];
    return k[(hpc ? 1 : 0) | (w31 ? 2 : 0) | (dense ? 4 : 0)];
}


int set_attrs(s2k_ctx *ctx)
{
    if (ctx->attr_set) return S2K_OK;
    const int smem = (int)sizeof(Smem);
    for (int v = 0; v < 8; ++v) CU(cudaFuncSetAttribute(minimizer_kernel(v & 1, v & 2, v & 4), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int v = 0; v < 4; ++v) CU(cudaFuncSetAttribute(minimizer_kernel(v & 1, false, v & 2, true), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ctx->attr_set = true;
    return S2K_OK;
}

void timing_prepare(s2k_ctx *ctx)
{
    Timing &T = ctx->tm;
    if (T.enabled && !T.created) {
        for (auto &e : T.ev) { cudaEventCreate(&e[0]); cudaEventCreate(&e[1]); }
        cudaEventCreate(&T.wv[0]); cudaEventCreate(&T.wv[1]);
        T.created = true;
    }
    T.n = 0;
}

// d_small layout (uint64 words): [0] record cursor, [4] ticket(u32), [5] err(u32), [6] ticket2(u32)
// Runs the whole device pipeline on `st`.  On return the totals have been read back (one sync).
// in_place_ok: the caller does not need the ordered minimizer stream on the device.  With S2K_NO_MINIMIZER_STREAM set the
// records then stay where k_minimizers appended them and the window stage reads them through a per-tile index.
int run_device(s2k_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_seq_off, uint64_t n_seqs, uint64_t n_bases,
               const Plan &P, cudaStream_t st, s2k_result *out, bool in_place_ok = false)
{
    const bool in_place = in_place_ok && (ctx->flags & S2K_NO_MINIMIZER_STREAM) && P.k <= (uint32_t)KW_MAX && !P.h64 && !P.h16;
    ctx->min_hi_valid = false;
    int rc;
    if ((rc = set_attrs(ctx)) != S2K_OK) return rc;
    timing_prepare(ctx);
    std::memset(out, 0, sizeof(*out));
    out->n_seqs = n_seqs;
    out->location = S2K_LOC_DEVICE;

    if ((rc = ensure(ctx, ctx->d_small, 64, false))) return rc;
    if ((rc = ensure(ctx, ctx->h_small, 64, true))) return rc;
    if ((rc = ensure(ctx, ctx->d_min_off, (n_seqs + 1) * 8, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_km_off, (n_seqs + 1) * 8, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_min_cnt, std::max<uint64_t>(n_seqs, 1) * 4, false))) return rc;
    const bool want_hpc_off = P.hpc && P.quirk;
    if (want_hpc_off && (rc = ensure(ctx, ctx->d_hpc_off, (n_seqs + 1) * 8, false))) return rc;

    uint64_t *small = ptr<uint64_t>(ctx->d_small);
    uint64_t *hsmall = ptr<uint64_t>(ctx->h_small);
    uint64_t n_min = 0;

    if (n_bases == 0 || n_seqs == 0 || P.none) {
        CU(cudaMemsetAsync(ctx->d_min_off.p, 0, (n_seqs + 1) * 8, st));
        CU(cudaMemsetAsync(ctx->d_km_off.p, 0, (n_seqs + 1) * 8, st));
        CU(cudaMemsetAsync(ctx->d_min_cnt.p, 0, std::max<uint64_t>(n_seqs, 1) * 4, st));
        CU(cudaStreamSynchronize(st));
        out->km_off = ptr<uint64_t>(ctx->d_km_off);
        out->min_off = ptr<uint64_t>(ctx->d_min_off);
        out->min_cnt = ptr<uint32_t>(ctx->d_min_cnt);
        return S2K_OK;
    }

    if ((rc = ensure(ctx, ctx->d_status, ((n_seqs + RT * RPT - 1) / (RT * RPT)) * 8 + 8, false))) return rc;

    // capacity of the minimizer stream: expected 2*density*(kept bases), with head room; grown and rerun on overflow
    uint64_t cap;
    {
        const double frac = selection_fraction(P);
        double rate = std::min(1.0, 2.0 * frac) * 1.15 + 0.0005;
        rate = std::max(rate, ctx->rate_hint * 1.05);
        cap = std::min<uint64_t>(n_bases, (uint64_t)((double)n_bases * rate) + 65536);
        if (ctx->flags & S2K_DEBUG_TINY_CAP) cap = std::min<uint64_t>(cap, 1000);
    }

    // The whole launch sequence is enqueued without a host round trip: every buffer is sized from `cap`, the kernels
    // after k_minimizers read the minimizer total from the device and do nothing once the record store has overflowed
    // (ERR_CAP); ONE synchronisation at the end brings back the totals and the error word.  An overflow (the batch is
    // denser than the expected selection rate, or one CTA's append region ran out) reruns the sequence once with the
    // exact size and a single global allocator.
    const uint64_t tile_eff = P.tile;
    const uint64_t n_tiles64 = (n_bases + tile_eff - 1) / tile_eff;
    if (n_tiles64 >= 0xfffffff0ull) return fail(ctx, S2K_ERR_BAD_PARAM, "batch too large");
    const uint32_t n_tiles = (uint32_t)n_tiles64;
    const int max_grid = ctx->sm_count * S2K_MINB;
    const size_t hscr_words = (size_t)max_grid * WIN * (P.h64 ? 2 : 1), smem = sizeof(Smem);
    void (*kfn)(const K1Args) = minimizer_kernel(P.hpc, P.w31, P.dense, P.h64);
    const uint32_t n_chunks = (n_tiles + ST - 1) / ST;
    const uint64_t rtiles = (n_seqs + RT * RPT - 1) / (RT * RPT);
    if ((rc = ensure(ctx, ctx->d_tile_lb, ((uint64_t)n_tiles + 1) * 4, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_tile_info, (uint64_t)n_tiles * 16, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_hscr, hscr_words * 4, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_min_loc, (n_seqs + 1) * 8, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_tile_pre, ((uint64_t)n_tiles + 1) * sizeof(ulonglong2), false))) return rc;
    if (in_place && (rc = ensure(ctx, ctx->d_tile_src, (uint64_t)n_tiles * sizeof(ulonglong2), false))) return rc;
    // d_tile_base: [tile_loc u64 x n_tiles][chunk_tot u64 x n_chunks][chunk_base u64x2 x (n_chunks+1)]
    if ((rc = ensure(ctx, ctx->d_tile_base, ((uint64_t)n_tiles + n_chunks + 2) * 8 + ((uint64_t)n_chunks + 1) * 16 + 16, false))) return rc;

    bool regions = true;                               // per-CTA append regions first; one global allocator on the rerun
    for (int attempt = 0;; ++attempt) {
        if (attempt == 4) return fail(ctx, S2K_ERR_INTERNAL, "minimizer kernel did not converge");
        // Per-CTA append regions get 3/4 + 1024 records of slack on top of an even share.  Tiles are handed out
        // dynamically, but the CTAs of one SM do not progress at the same pace when the ALU pipe is saturated (the warp
        // scheduler prefers the higher warp slots): measured on config-2 shaped batches, the busiest CTA appends more
        // than 1.33 x the average (HPC off, d = 0.01: an 1/8 slack overflowed on every run and the batch ran twice).
        const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)max_grid);
        const uint64_t slack = (ctx->flags & S2K_DEBUG_TINY_CAP) ? 1 : 3ull * cap / (4ull * (uint64_t)grid) + 1024;
        const uint64_t region_cap = regions ? cap / (uint64_t)grid + slack : 0;
        const uint64_t tmp_cap = regions ? region_cap * (uint64_t)grid : cap;
        const uint64_t item_cap = std::max<uint64_t>(cap, 1);
        if ((rc = ensure(ctx, ctx->d_tmp, tmp_cap * sizeof(uint4), false))) return rc;
        if ((rc = ensure(ctx, ctx->d_mins, (in_place ? 1 : item_cap) * sizeof(uint4), false))) return rc;
        if (P.h64 && (rc = ensure(ctx, ctx->d_tmp_hi, tmp_cap * 4, false))) return rc;
        if (P.h64 && (rc = ensure(ctx, ctx->d_mins_hi, item_cap * 4, false))) return rc;
        if ((rc = ensure(ctx, ctx->d_hash, item_cap * 8, false))) return rc;
        if ((rc = ensure(ctx, ctx->d_start, item_cap * 4, false))) return rc;
        if ((rc = ensure(ctx, ctx->d_end, item_cap * 4, false))) return rc;
        if ((rc = ensure(ctx, ctx->d_rev, item_cap, false))) return rc;

        K1Args A;
        A.bases = d_bases; A.seq_off = d_seq_off;
        A.tile_lb = ptr<uint32_t>(ctx->d_tile_lb);
        A.ticket = reinterpret_cast<uint32_t *>(small + 4);
        A.cursor = reinterpret_cast<unsigned long long *>(small + 0);
        A.tile_info = ptr<uint4>(ctx->d_tile_info);
        A.min_out = ptr<uint4>(ctx->d_tmp); A.min_cap = tmp_cap;
        A.region_cap = region_cap;
        A.min_off = ptr<uint64_t>(ctx->d_min_loc);
        A.hpc_off = want_hpc_off ? ptr<uint64_t>(ctx->d_hpc_off) : nullptr;
        A.hscr = ptr<uint32_t>(ctx->d_hscr);
        A.err = reinterpret_cast<uint32_t *>(small + 5);
        A.n_seqs = n_seqs; A.n_bases = n_bases; A.n_tiles = n_tiles;
        A.tile = (uint32_t)tile_eff; A.halo = P.halo; A.l = P.l; A.d = P.d; A.need = P.need; A.thr = P.thr;
        A.vmask = P.simd ? 0x0f0f0f0fu : 0xffffffffu; A.one = 1u;
        A.trunc16 = P.trunc16 ? 1u : 0u;
        A.thr64 = P.thr64; A.min_hi = P.h64 ? ptr<uint32_t>(ctx->d_tmp_hi) : nullptr;
        if (P.h64) std::memcpy(A.xy64, P.xy64, sizeof(P.xy64)); else std::memset(A.xy64, 0, sizeof(A.xy64));
        std::memcpy(A.cls_lut, P.lut, 256);
        std::memcpy(A.xy, P.xy, sizeof(P.xy));
        std::memcpy(A.xf, P.xf, sizeof(P.xf));
        std::memcpy(A.x2, P.x2, sizeof(P.x2));
        CU(cudaMemsetAsync(small, 0, 64, st));
        CU(cudaMemsetAsync(ctx->d_status.p, 0, rtiles * 8, st));
        S2K_LAUNCH(k_tile_bounds, (n_tiles + 1 + 255) / 256, 256, 0, st, false, d_seq_off, n_seqs, n_bases, (uint32_t)tile_eff, n_tiles,
                   ptr<uint32_t>(ctx->d_tile_lb));
        Timing &T = ctx->tm;
        const bool rec = T.enabled && T.n < 64;
        if (rec) cudaEventRecord(T.ev[T.n][0], st);
        S2K_LAUNCH(kfn, grid, NT, smem, st, true, A);
        if (rec) { cudaEventRecord(T.ev[T.n][1], st); ++T.n; }
        CU(cudaGetLastError());
        ctx->launches += 2;

        // ---- tile order: exclusive prefix of the per-tile counts, where each tile's records belong
        unsigned long long *tile_loc = ptr<unsigned long long>(ctx->d_tile_base);
        unsigned long long *chunk_tot = tile_loc + n_tiles;
        ulonglong2 *chunk_base = reinterpret_cast<ulonglong2 *>(
            (reinterpret_cast<uintptr_t>(chunk_tot + n_chunks) + 15) & ~uintptr_t(15));
        S2K_LAUNCH(k_tile_scan_a, n_chunks, ST, 0, st, false, ptr<uint4>(ctx->d_tile_info), n_tiles, tile_loc, chunk_tot);
        S2K_LAUNCH(k_tile_scan_b, 1, ST, 0, st, false, chunk_tot, n_chunks, chunk_base, (unsigned long long)cap, A.err);
        KFArgs F;
        F.tile_info = ptr<uint4>(ctx->d_tile_info); F.tile_loc = tile_loc; F.chunk_base = chunk_base;
        F.tmp = ptr<uint4>(ctx->d_tmp); F.mins = ptr<uint4>(ctx->d_mins);
        F.min_cap = tmp_cap; F.n_tiles = n_tiles;
        F.tile_pre = ptr<ulonglong2>(ctx->d_tile_pre);
        F.tile_src = in_place ? ptr<ulonglong2>(ctx->d_tile_src) : nullptr; F.copy = in_place ? 0 : 1;
        F.tmp_hi = P.h64 ? ptr<uint32_t>(ctx->d_tmp_hi) : nullptr; F.mins_hi = P.h64 ? ptr<uint32_t>(ctx->d_mins_hi) : nullptr;
        F.hmask = P.h16 ? 0xffffu : 0xffffffffu;
        F.err = A.err;
        const int gridf = (int)std::min<uint64_t>(((uint64_t)n_tiles + 8) / 8, (uint64_t)ctx->sm_count * 8);
        S2K_LAUNCH(k_finalize, gridf, 256, 0, st, false, F);
        CU(cudaGetLastError());
        ctx->launches += 3;

        // ---- per-sequence prefixes and counts, then the window stage
        K2Args B;
        B.mins = in_place ? ptr<uint4>(ctx->d_tmp) : ptr<uint4>(ctx->d_mins);
        B.tile_src = in_place ? ptr<ulonglong2>(ctx->d_tile_src) : nullptr; B.n_tiles = n_tiles; B.tile = (uint32_t)tile_eff;
        B.tile_magic = ~0ull / tile_eff + 1;
        B.min_loc = ptr<uint64_t>(ctx->d_min_loc);
        B.hpc_loc = want_hpc_off ? ptr<uint64_t>(ctx->d_hpc_off) : nullptr;
        B.tile_pre = ptr<ulonglong2>(ctx->d_tile_pre);
        B.seq_off = d_seq_off; B.bases = d_bases; B.n_seqs = n_seqs; B.n_bases = n_bases;
        B.l = P.l; B.k = P.k; B.quirk = P.quirk; B.hpc = P.hpc;
        B.min_off = ptr<uint64_t>(ctx->d_min_off);
        B.km_off = ptr<uint64_t>(ctx->d_km_off);
        B.min_cnt = ptr<uint32_t>(ctx->d_min_cnt);
        B.status = ptr<uint64_t>(ctx->d_status);
        B.ticket = reinterpret_cast<uint32_t *>(small + 6);
        B.err = A.err;
        if (T.enabled) cudaEventRecord(T.wv[0], st);
        const int grid2 = (int)std::min<uint64_t>(rtiles + 1, (uint64_t)ctx->sm_count * 8);
        S2K_LAUNCH(k_read_counts, grid2, RT, 0, st, false, B);
        K2SArgs Sc;
        Sc.min_cnt = B.min_cnt; Sc.km_off = B.km_off; Sc.n_seqs = n_seqs; Sc.k = P.k;
        Sc.status = B.status; Sc.ticket = B.ticket; Sc.err = A.err;
        const uint64_t stiles = (n_seqs + RT * SPT - 1) / (RT * SPT);
        S2K_LAUNCH(k_item_scan, (int)std::min<uint64_t>(std::max<uint64_t>(stiles, 1), (uint64_t)ctx->sm_count * 8), RT, 0, st, false, Sc);
        CU(cudaGetLastError());
        ctx->launches += 1;
        K3Args C;
        C.mins = B.mins; C.min_off = B.min_off; C.km_off = B.km_off; C.n_min_p = A.cursor; C.err = A.err; C.k = P.k;
        C.hash = ptr<uint64_t>(ctx->d_hash); C.start = ptr<uint32_t>(ctx->d_start);
        C.end = ptr<uint32_t>(ctx->d_end); C.rev = ptr<uint8_t>(ctx->d_rev);
        C.hash_hi = P.h64 ? ptr<uint32_t>(ctx->d_mins_hi) : nullptr;
        C.mix16 = P.h16 ? 1u : 0u;
        if (in_place) {
            K3TArgs D;
            D.W = C; D.tile_info = ptr<uint4>(ctx->d_tile_info); D.tile_src = B.tile_src; D.n_tiles = n_tiles;
            // short reads (fewer than ~4k expected minimizers per sequence): most windows span two sequences
            D.ridtest = (double)cap / (double)n_seqs < 4.0 * (double)P.k ? 1u : 0u;
            // 16 CTAs per SM in the grid (2.7 waves at six resident): one-wave grids measured slower, the tiles are uneven
            const int g3 = (int)std::min<uint64_t>(((uint64_t)n_tiles + 7) / 8, (uint64_t)ctx->sm_count * 16);
            switch ((int)P.k) {
#define S2K_WINDOWS_CASE(K) case K: S2K_LAUNCH(k_windows_t<K>, g3, 256, 0, st, false, D); break;
                S2K_WINDOWS_CASE(1) S2K_WINDOWS_CASE(2) S2K_WINDOWS_CASE(3) S2K_WINDOWS_CASE(4) S2K_WINDOWS_CASE(5) S2K_WINDOWS_CASE(6)
                S2K_WINDOWS_CASE(7) S2K_WINDOWS_CASE(8) S2K_WINDOWS_CASE(9) S2K_WINDOWS_CASE(10) S2K_WINDOWS_CASE(11) S2K_WINDOWS_CASE(12)
#undef S2K_WINDOWS_CASE
            }
        } else {
            const int g3 = (int)std::min<uint64_t>((item_cap + 255) / 256, (uint64_t)ctx->sm_count * 16);
            switch (P.k <= (uint32_t)KW_MAX && !P.h64 && !P.h16 ? (int)P.k : 0) {
#define S2K_WINDOWS_CASE(K) case K: S2K_LAUNCH(k_windows_w<K>, g3, 256, 0, st, false, C); break;
                S2K_WINDOWS_CASE(1) S2K_WINDOWS_CASE(2) S2K_WINDOWS_CASE(3) S2K_WINDOWS_CASE(4) S2K_WINDOWS_CASE(5) S2K_WINDOWS_CASE(6)
                S2K_WINDOWS_CASE(7) S2K_WINDOWS_CASE(8) S2K_WINDOWS_CASE(9) S2K_WINDOWS_CASE(10) S2K_WINDOWS_CASE(11) S2K_WINDOWS_CASE(12)
#undef S2K_WINDOWS_CASE
                default: S2K_LAUNCH(k_windows, g3, 256, 0, st, false, C);
            }
        }
        CU(cudaGetLastError());
        if (T.enabled) cudaEventRecord(T.wv[1], st);
        ctx->launches += 2;
        CU(cudaMemcpyAsync(hsmall, small, 8, cudaMemcpyDeviceToHost, st));             // record cursor == total minimizers
        CU(cudaMemcpyAsync(hsmall + 2, small + 5, 8, cudaMemcpyDeviceToHost, st));     // error word
        CU(cudaMemcpyAsync(hsmall + 4, B.km_off + n_seqs, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        n_min = hsmall[0];
        const uint32_t errw = (uint32_t)hsmall[2];
        if (errw & ERR_ALIGN) return fail(ctx, S2K_ERR_INTERNAL, "shared-memory tables are not 256-byte aligned (or a warp of k_minimizers started partial)");
        if (errw & ERR_SPIN) return fail(ctx, S2K_ERR_INTERNAL, "scan look-back timed out");
        if (!(errw & ERR_CAP) && n_min <= cap) break;
        if (getenv("S2K_TRACE_RERUN"))
            fprintf(stderr, "[s2k] rerun: err %u, minimizers %llu, cap %llu, region_cap %llu x %d CTAs, tiles %u\n", errw,
                    (unsigned long long)n_min, (unsigned long long)cap, (unsigned long long)region_cap, grid, n_tiles);
        cap = std::max(cap, n_min);                    // exact size known now: rerun once, with one global allocator
        regions = false;
        T.n = 0;
    }
    ctx->rate_hint = std::max(ctx->rate_hint, (double)n_min / (double)n_bases);
    if (ctx->tm.enabled) {
        Timing &T = ctx->tm;
        double ms = 0;
        for (int i = 0; i < T.n; ++i) { float f = 0; cudaEventElapsedTime(&f, T.ev[i][0], T.ev[i][1]); ms += f; }
        float f = 0; cudaEventElapsedTime(&f, T.wv[0], T.wv[1]);
        T.min_ms = ms; T.win_ms = f; T.min_launches = (uint32_t)T.n;
    }
    ctx->min_hi_valid = P.h64;
    out->n_minimizers = n_min;
    out->n_items = hsmall[4];
    out->hash = ptr<uint64_t>(ctx->d_hash);
    out->start = ptr<uint32_t>(ctx->d_start);
    out->end = ptr<uint32_t>(ctx->d_end);
    out->rev = ptr<uint8_t>(ctx->d_rev);
    out->km_off = ptr<uint64_t>(ctx->d_km_off);
    out->minimizers = in_place ? nullptr : reinterpret_cast<const s2k_minimizer *>(ctx->d_mins.p);
    out->min_off = ptr<uint64_t>(ctx->d_min_off);
    out->min_cnt = ptr<uint32_t>(ctx->d_min_cnt);
    return S2K_OK;
}

int check_offsets(s2k_ctx *ctx, const uint64_t *seq_off, uint64_t n_seqs)
{
    if (n_seqs >= 0xfffffff0ull) return fail(ctx, S2K_ERR_BAD_OFFSETS, "too many sequences in one batch (limit 2^32-16)");
    if (seq_off[0] != 0) return fail(ctx, S2K_ERR_BAD_OFFSETS, "seq_off[0] must be 0");
    for (uint64_t i = 0; i < n_seqs; ++i) {
        if (seq_off[i + 1] < seq_off[i]) return fail(ctx, S2K_ERR_BAD_OFFSETS, "seq_off must be non-decreasing");
        if (seq_off[i + 1] - seq_off[i] >= 0xffffffffull) return fail(ctx, S2K_ERR_BAD_OFFSETS, "a sequence has >= 2^32-1 bases");
    }
    return S2K_OK;
}

} // namespace

// ================================================================================================ C ABI
extern "C" {

int s2k_abi_version(void) { return S2K_ABI_VERSION; }

const char *s2k_strerror(int status)
{
    switch (status) {
    case S2K_OK: return "ok";
    case S2K_ERR_BAD_PARAM: return "bad parameter";
    case S2K_ERR_L_TOO_BIG: return "l out of range for the selected hash mode";
    case S2K_ERR_CUDA: return "CUDA error";
    case S2K_ERR_OOM: return "out of memory";
    case S2K_ERR_BAD_OFFSETS: return "invalid sequence offsets";
    case S2K_ERR_INTERNAL: return "internal device-side check failed";
    case S2K_ERR_NULL: return "null pointer";
    case S2K_ERR_IO: return "input file error";
    default: return "unknown status";
    }
}

const char *s2k_last_error(const s2k_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

uint64_t s2k_bound_u64(double density)
{
    const double v = density * 18446744073709551615.0;       // u64::MAX as f64 == 2^64
    if (!(v > 0.0)) return 0ull;
    if (v >= 18446744073709551616.0) return ~0ull;           // Rust `as` saturates
    return (uint64_t)v;
}
uint32_t s2k_bound_u16(double density)
{
    const double v = density * 65535.0;                      // ((density as f64) * (u16::MAX as f64)) as u16
    if (!(v > 0.0)) return 0u;
    if (v >= 65535.0) return 0xffffu;                        // Rust `as` saturates
    return (uint32_t)v;
}
int s2k_last_minimizer_hash_hi(const s2k_ctx *ctx, const uint32_t **d_hi)
{
    if (!ctx || !d_hi) return S2K_ERR_NULL;
    *d_hi = ctx->min_hi_valid ? reinterpret_cast<const uint32_t *>(ctx->d_mins_hi.p) : nullptr;
    return S2K_OK;
}
void s2k_bounds(double density, uint32_t *b_scalar, uint32_t *b_simd, uint32_t *b_31)
{
    const uint32_t bs = bound_scalar(density), bv = bound_simd(bs);
    if (b_scalar) *b_scalar = bs;
    if (b_simd) *b_simd = bv;
    if (b_31) *b_31 = bv / 2;
}

int s2k_ctx_create(int device, s2k_ctx **out)
{
    if (!out) return S2K_ERR_NULL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { (void)cudaGetLastError(); return S2K_ERR_CUDA; }
    if (device < 0 || device >= count) return S2K_ERR_BAD_PARAM;
    s2k_ctx *ctx = new (std::nothrow) s2k_ctx();
    if (!ctx) return S2K_ERR_OOM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        (void)cudaGetLastError();
        delete ctx;
        return S2K_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = ctx;
    return S2K_OK;
}

void s2k_ctx_destroy(s2k_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); }
    Buf *all[] = {&ctx->d_bases, &ctx->d_seq_off, &ctx->d_tile_lb, &ctx->d_status, &ctx->d_small, &ctx->d_mins,
                  &ctx->d_min_off, &ctx->d_min_loc, &ctx->d_tile_pre, &ctx->d_hpc_off, &ctx->d_km_off, &ctx->d_min_cnt, &ctx->d_hash, &ctx->d_start,
                  &ctx->d_end, &ctx->d_rev, &ctx->d_ct_keys, &ctx->d_ct_cnt, &ctx->d_ct_first, &ctx->d_ct_side, &ctx->d_co_hash, &ctx->d_co_cnt, &ctx->d_co_first, &ctx->h_ct_side, &ctx->d_rle_hpc, &ctx->d_rle_pos, &ctx->d_hscr, &ctx->d_tmp, &ctx->d_tmp_hi, &ctx->d_mins_hi, &ctx->d_tile_info, &ctx->d_tile_base, &ctx->d_tile_src, &ctx->h_hash, &ctx->h_start, &ctx->h_end,
                  &ctx->h_rev, &ctx->h_km_off, &ctx->h_mins, &ctx->h_min_off, &ctx->h_min_cnt, &ctx->h_small,
                  &ctx->h_rle_hpc, &ctx->h_rle_pos, &ctx->d_in[0], &ctx->d_in[1], &ctx->d_in[2], &ctx->d_in_off[0], &ctx->d_in_off[1], &ctx->d_in_off[2],
                  &ctx->h_off_stage[0], &ctx->h_off_stage[1], &ctx->h_off_stage[2], &ctx->d_piece, &ctx->h_piece, &ctx->d_stage, &ctx->h_fx_bases, &ctx->h_fx_off, &ctx->h_ascii[0], &ctx->h_ascii[1], &ctx->h_ascii[2], &ctx->d_pack[0], &ctx->d_pack[1], &ctx->d_pack[2],
                  &ctx->h_pack[0], &ctx->h_pack[1], &ctx->h_pack[2]};
    for (Buf *b : all) release(*b);
    if (ctx->tm.created) {
        for (auto &e : ctx->tm.ev) { cudaEventDestroy(e[0]); cudaEventDestroy(e[1]); }
        cudaEventDestroy(ctx->tm.wv[0]); cudaEventDestroy(ctx->tm.wv[1]);
    }
    if (ctx->pipe_ready) {
        cudaStreamDestroy(ctx->s_h2d); cudaStreamDestroy(ctx->s_d2h);
        for (int i = 0; i < 3; ++i) { cudaEventDestroy(ctx->ev_in[i]); cudaEventDestroy(ctx->ev_free[i]); }
        cudaEventDestroy(ctx->ev_out); cudaEventDestroy(ctx->ev_done);
        for (int i = 0; i < 3; ++i) cudaEventDestroy(ctx->ev_pack[i]);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int s2k_ctx_set_flags(s2k_ctx *ctx, uint32_t flags)
{
    if (!ctx) return S2K_ERR_NULL;
    ctx->flags = flags;
    return S2K_OK;
}

int s2k_ctx_set_timing(s2k_ctx *ctx, int enabled)
{
    if (!ctx) return S2K_ERR_NULL;
    ctx->tm.enabled = enabled != 0;
    return S2K_OK;
}

int s2k_last_kernel_ms(const s2k_ctx *ctx, double *minimizer_ms, double *window_ms, uint32_t *minimizer_launches)
{
    if (!ctx) return S2K_ERR_NULL;
    if (minimizer_ms) *minimizer_ms = ctx->tm.min_ms;
    if (window_ms) *window_ms = ctx->tm.win_ms;
    if (minimizer_launches) *minimizer_launches = ctx->tm.min_launches;
    return S2K_OK;
}

uint64_t s2k_launch_count(const s2k_ctx *ctx) { return ctx ? ctx->launches : 0; }
// ---------------------------------------------------------------------------------------------- consumer side: counting
uint32_t s2k_count_part(uint64_t hash, uint32_t n_parts)
{
    uint64_t x = hash;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (uint32_t)(((x >> 32) * (uint64_t)n_parts) >> 32);
}

int s2k_count_device(s2k_ctx *ctx, const uint64_t *d_hash, const uint64_t *d_id, uint64_t n_items, uint64_t id_base,
                     void *stream, s2k_count_result *out)
{
    if (!ctx) return S2K_ERR_NULL;
    if (!out || (n_items && !d_hash)) return fail(ctx, S2K_ERR_NULL, "null argument");
    CU(cudaSetDevice(ctx->device));
    ctx->err.clear();
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    std::memset(out, 0, sizeof(*out));
    out->location = S2K_LOC_DEVICE;
    int rc;
    uint64_t capacity = 1024;                              // load factor <= 2/3 even if every hash is distinct
    while (capacity * 2 < n_items * 3) capacity <<= 1;
    if ((rc = ensure(ctx, ctx->d_ct_keys, capacity * 8, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_ct_cnt, capacity * 4, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_ct_first, capacity * 8, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_ct_side, 32, false))) return rc;
    if ((rc = ensure(ctx, ctx->h_ct_side, 32, true))) return rc;
    const uint64_t n_out = std::max<uint64_t>(n_items, 1);
    if ((rc = ensure(ctx, ctx->d_co_hash, n_out * 8, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_co_cnt, n_out * 4, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_co_first, n_out * 8, false))) return rc;
    CU(cudaMemsetAsync(ctx->d_ct_keys.p, 0xff, capacity * 8, st));
    CU(cudaMemsetAsync(ctx->d_ct_first.p, 0xff, capacity * 8, st));
    CU(cudaMemsetAsync(ctx->d_ct_cnt.p, 0, capacity * 4, st));
    unsigned long long *side = ptr<unsigned long long>(ctx->d_ct_side);
    CU(cudaMemsetAsync(side, 0, 32, st));
    CU(cudaMemsetAsync(side + 1, 0xff, 8, st));
    KCArgs A;
    A.hash = d_hash; A.id = d_id; A.n_items = n_items; A.id_base = id_base;
    A.keys = ptr<unsigned long long>(ctx->d_ct_keys); A.cnt = ptr<uint32_t>(ctx->d_ct_cnt);
    A.first = ptr<unsigned long long>(ctx->d_ct_first); A.mask = capacity - 1; A.side = side;
    const int g1 = (int)std::min<uint64_t>((n_items + 255) / 256 + 1, (uint64_t)ctx->sm_count * 16);
    S2K_LAUNCH(k_count_insert, g1, 256, 0, st, false, A);
    KCOut O;
    O.keys = A.keys; O.cnt = A.cnt; O.first = A.first; O.capacity = capacity; O.side = side;
    O.out_hash = ptr<uint64_t>(ctx->d_co_hash); O.out_cnt = ptr<uint32_t>(ctx->d_co_cnt); O.out_first = ptr<uint64_t>(ctx->d_co_first);
    const int g2 = (int)std::min<uint64_t>((capacity + 255) / 256, (uint64_t)ctx->sm_count * 16);
    S2K_LAUNCH(k_count_compact, g2, 256, 0, st, false, O);
    CU(cudaGetLastError());
    ctx->launches += 2;
    CU(cudaMemcpyAsync(ctx->h_ct_side.p, side, 32, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    out->n_distinct = ptr<uint64_t>(ctx->h_ct_side)[2];
    out->n_items = n_items;
    out->hash = ptr<uint64_t>(ctx->d_co_hash);
    out->count = ptr<uint32_t>(ctx->d_co_cnt);
    out->first = ptr<uint64_t>(ctx->d_co_first);
    return S2K_OK;
}

int s2k_count_partition_device(s2k_ctx *ctx, const uint64_t *d_hash, uint64_t n_items, uint64_t id_base, uint32_t n_parts,
                               uint64_t *part_counts, uint64_t *d_out_hash, uint64_t *d_out_id, void *stream)
{
    if (!ctx) return S2K_ERR_NULL;
    if (!part_counts || (n_items && (!d_hash || !d_out_hash || !d_out_id))) return fail(ctx, S2K_ERR_NULL, "null argument");
    if (n_parts == 0 || n_parts > (uint32_t)CT_MAX_PARTS) return fail(ctx, S2K_ERR_BAD_PARAM, "n_parts must be 1..64");
    CU(cudaSetDevice(ctx->device));
    ctx->err.clear();
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    int rc;
    if ((rc = ensure(ctx, ctx->d_ct_side, 8 * 2 * CT_MAX_PARTS, false))) return rc;
    if ((rc = ensure(ctx, ctx->h_ct_side, 8 * 2 * CT_MAX_PARTS, true))) return rc;
    unsigned long long *cnt = ptr<unsigned long long>(ctx->d_ct_side), *cur = cnt + CT_MAX_PARTS;
    uint64_t *h = ptr<uint64_t>(ctx->h_ct_side);
    CU(cudaMemsetAsync(cnt, 0, 8 * CT_MAX_PARTS, st));
    const int g = (int)std::min<uint64_t>((n_items + 255) / 256 + 1, (uint64_t)ctx->sm_count * 8);
    S2K_LAUNCH(k_count_hist, g, 256, 0, st, false, d_hash, n_items, n_parts, cnt);
    CU(cudaMemcpyAsync(h, cnt, 8 * CT_MAX_PARTS, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    uint64_t acc = 0;
    for (uint32_t p = 0; p < n_parts; ++p) { part_counts[p] = h[p]; h[CT_MAX_PARTS + p] = acc; acc += h[p]; }
    CU(cudaMemcpyAsync(cur, h + CT_MAX_PARTS, 8 * n_parts, cudaMemcpyHostToDevice, st));
    S2K_LAUNCH(k_count_scatter, g, 256, 0, st, false, d_hash, n_items, id_base, n_parts, cur, d_out_hash, d_out_id);
    CU(cudaGetLastError());
    ctx->launches += 2;
    CU(cudaStreamSynchronize(st));
    return S2K_OK;
}

#if defined(S2K_PHASE_CLOCKS) && !defined(S2K_EMU)
// debug builds only (tools/phase_clocks.py): cycles thread 0 of every CTA spent per phase of k_minimizers since the last call
extern "C" int s2k_debug_phase_clocks(unsigned long long *out16)
{
    if (cudaMemcpyFromSymbol(out16, s2k::g_phase, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(s2k::g_phase, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif


int s2k_host_alloc(size_t bytes, void **out)
{
    if (!out) return S2K_ERR_NULL;
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return S2K_ERR_OOM; }
    return S2K_OK;
}
void s2k_host_free(void *p) { if (p) cudaFreeHost(p); }

int s2k_synth_device(s2k_ctx *ctx, uint64_t seed, uint64_t first, uint64_t count, uint8_t *d_out, void *stream)
{
    if (!ctx) return S2K_ERR_NULL;
    if (count && !d_out) return fail(ctx, S2K_ERR_NULL, "d_out is null");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : ctx->stream;
    if (count) {
        const int grid = (int)std::min<uint64_t>((count / 16 + 255) / 256 + 1, (uint64_t)ctx->sm_count * 16);
        S2K_LAUNCH(k_synth, grid, 256, 0, st, false, seed, first, count, d_out);
        CU(cudaGetLastError());
    }
    if (!stream) CU(cudaStreamSynchronize(st));
    return S2K_OK;
}

int s2k_run_device(s2k_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_seq_off, uint64_t n_seqs, uint64_t n_bases,
                   const s2k_params *params, void *stream, s2k_result *out)
{
    if (!ctx) return S2K_ERR_NULL;
    if (!out || (!d_seq_off)) return fail(ctx, S2K_ERR_NULL, "null argument");
    if (n_bases && !d_bases) return fail(ctx, S2K_ERR_NULL, "bases is null");
    if (n_seqs >= 0xfffffff0ull) return fail(ctx, S2K_ERR_BAD_OFFSETS, "too many sequences in one batch");
    if ((reinterpret_cast<uintptr_t>(d_bases) & 15u) != 0) return fail(ctx, S2K_ERR_BAD_PARAM, "device bases must be 16-byte aligned");
    Plan P;
    int rc = make_plan(ctx, params, P);
    if (rc != S2K_OK) return rc;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : ctx->stream;
    ctx->err.clear();
    return run_device(ctx, d_bases, d_seq_off, n_seqs, n_bases, P, st, out, true);
}

int s2k_ctx_set_transport(s2k_ctx *ctx, int host_threads, double pack_ratio)
{
    if (!ctx) return S2K_ERR_NULL;
    if (host_threads < 0 || host_threads > 256 || !(pack_ratio >= 0.0 && pack_ratio <= 1.0)) return fail(ctx, S2K_ERR_BAD_PARAM, "bad transport setting");
    ctx->host_threads = host_threads;
    ctx->pack_ratio = pack_ratio;
    return S2K_OK;
}

int s2k_last_transport(const s2k_ctx *ctx, uint64_t *h2d_bytes, uint64_t *packed_slabs, uint64_t *plain_slabs)
{
    if (!ctx) return S2K_ERR_NULL;
    if (h2d_bytes) *h2d_bytes = ctx->tr_h2d_bytes;
    if (packed_slabs) *packed_slabs = ctx->tr_packed;
    if (plain_slabs) *plain_slabs = ctx->tr_plain;
    return S2K_OK;
}

int s2k_ctx_set_slab_bytes(s2k_ctx *ctx, uint64_t bytes)
{
    if (!ctx) return S2K_ERR_NULL;
    ctx->slab_bytes = bytes;
    return S2K_OK;
}

// Large host batches: slabs flow through H2D -> kernels -> D2H on three streams, so that both PCIe directions and the
// compute overlap.  Input slabs are double-buffered on the device; the device result buffers are single (a slab's D2H
// is long finished when the next slab's H2D completes).
//
// A slab is a run of whole sequences, or a PIECE of one long sequence (longer than 1.5 slabs; a chromosome).  A piece
// owns the raw range [b0, b1) and is resident with a right overlap, [b0, hi): it runs as a sequence of its own (tail
// rule off) and keeps the minimizers that start in its range and the k-min-mers whose first minimizer it owns -- an
// l-mer depends on its own bases only, so nothing to the left is needed once the cut sits on a homopolymer-run boundary,
// and the overlap holds the k-1 further minimizers of the windows that begin near b1 (checked; if some stretch is so
// poor in minimizers that it does not, the batch is redone with an 8 times larger overlap).  The AVX-512 tail rule
// (src/nthash_avx512_32.rs:134-138) is a property of the whole sequence: kept bases are counted per piece and the last
// piece drops the minimizers ending in the final 16 kept bases when the rule fires.  (The multi-GPU split of one
// sequence, sharding.py, follows the same ownership rule.)
static std::atomic<long> g_stage{0}, g_aux{0};
static void stage_watchdog()
{
    static std::atomic<bool> started{false};
    if (!getenv("S2K_WATCHDOG") || started.exchange(true)) return;
    std::thread([]() {
        long last = -1;
        for (;;) {
            std::this_thread::sleep_for(std::chrono::seconds(5));
            const long st_now = g_stage.load();
            if (st_now == last && st_now != 0 && st_now != 600000)
                fprintf(stderr, "[s2k watchdog] run_pipelined has been at stage %ld (aux %ld) for 5 s\n", st_now, g_aux.load());
            last = st_now;
        }
    }).detach();
}
#define STAGE(x) g_stage.store(x)
struct Slab {
    uint64_t r0, r1;              // sequences [r0, r1); a piece: r1 = r0 + 1
    bool piece, first, last;      // piece of sequence r0; first / last piece of it
    uint64_t b0, b1, hi;          // absolute offsets into `bases`: first base copied, end of the owned range, end of the copy
};
// packed_in: `bases` is the caller's 2-bit packed batch (s2k_run_packed2) -- slabs are copied as they are and unpacked
// on the device; the host packers stay idle.
// fx: the bases come from a mapped FASTA/FASTQ file (s2k_run_fastx; `bases` is null): EVERY slab is gathered from the file
// and packed to 2 bits by the host threads straight into the pinned staging ring -- file bytes are read once, a quarter
// of a byte per base is written -- while earlier slabs are on the device; a slab holding a byte other than upper-case
// A/C/G/T is gathered as ASCII instead.
static int run_pipelined(s2k_ctx *ctx, const uint8_t *bases, const uint64_t *seq_off, uint64_t n_seqs, const Plan &P,
                         uint64_t slab_bytes, uint64_t overlap, int &overlap_short, s2k_result *out, bool packed_in,
                         const FxSource *fx = nullptr, int fx_threads = 0)
{
    int rc;
    auto base_at = [&](uint64_t i) -> uint8_t {             // equality of bases is all the host ever asks
        if (fx) return fx_base_at(*fx, i);
        return packed_in ? (uint8_t)((bases[i >> 2] >> (2 * (i & 3))) & 3u) : bases[i];
    };
    overlap_short = 0;                                     // 1: redo with a longer overlap, 2: only the one-shot run will do
    stage_watchdog();
    STAGE(1);
    if (!ctx->pipe_ready) {
        CU(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 3; ++i) { CU(cudaEventCreate(&ctx->ev_in[i])); CU(cudaEventCreate(&ctx->ev_free[i])); }
        CU(cudaEventCreate(&ctx->ev_out)); CU(cudaEventCreate(&ctx->ev_done));
        for (int i = 0; i < 3; ++i) CU(cudaEventCreate(&ctx->ev_pack[i]));
        ctx->pipe_ready = true;
    }
    cudaStream_t st = ctx->stream;
    const uint64_t n_bases = seq_off[n_seqs];
    const bool want_min = (ctx->flags & S2K_WANT_MINIMIZERS) != 0;
    Plan PP = P;                                           // pieces: the tail rule is applied here, not by the kernels
    PP.quirk = false;
    // slab plan
    std::vector<Slab> slabs;
    uint64_t max_b = 0, max_n = 0;
    bool any_piece = false;
    for (uint64_t r0 = 0; r0 < n_seqs;) {
        const uint64_t len0 = seq_off[r0 + 1] - seq_off[r0];
        if (len0 > slab_bytes + slab_bytes / 2 && len0 > (uint64_t)P.l) {
            const uint64_t s0 = seq_off[r0], s1 = seq_off[r0 + 1];
            // The AVX-512 tail rule (src/nthash_avx512_32.rs:134-138) drops the minimizers that END in the last 16 kept
            // bases; they START within the last l+15 kept bases.  Only the LAST piece knows whether the rule fired (it
            // needs the kept-base count of the whole sequence), so it must own every one of them, and k-1 surviving
            // minimizers before them for the windows the piece before it owns (checked when the rule fires): no cut
            // after `tail_lo` = the (l+16)-th kept base from the end, moved left by the overlap and onto a run start.
            // A homopolymer or no-run tail longer than a piece makes the last piece longer, not the result wrong.
            uint64_t tail_lo = s1;
            if (P.quirk) {
                uint64_t left = (uint64_t)P.l + 16;
                while (left && tail_lo > s0) {
                    --tail_lo;
                    if (!P.hpc || tail_lo == s0 || base_at(tail_lo) != base_at(tail_lo - 1)) --left;
                }
                tail_lo = tail_lo - s0 > overlap ? tail_lo - overlap : s0;
                if (P.hpc) while (tail_lo > s0 && base_at(tail_lo) == base_at(tail_lo - 1)) --tail_lo;
            }
            uint64_t b0 = s0;
            while (b0 < s1) {
                uint64_t b1 = std::min(s1, b0 + std::min<uint64_t>(slab_bytes, 1ull << 31));
                if (s1 - b1 < slab_bytes / 2) b1 = s1;                          // no tiny last piece
                if (P.hpc) while (b1 < s1 && base_at(b1) == base_at(b1 - 1)) ++b1;  // cut on a run boundary
                if (b1 < s1 && b1 > tail_lo) b1 = tail_lo > b0 ? tail_lo : s1;  // tail_lo is a kept base: a run boundary
                const uint64_t hi = b1 == s1 ? s1 : std::min(s1, b1 + overlap);
                slabs.push_back(Slab{r0, r0 + 1, true, b0 == s0, b1 == s1, b0, b1, hi});
                max_b = std::max(max_b, hi - b0);
                b0 = b1;
            }
            max_n = std::max<uint64_t>(max_n, 1);
            any_piece = true;
            ++r0;
            continue;
        }
        // whole sequences up to one slab of bases (binary search: batches of 10^8 short reads are common); a sequence
        // between 1 and 1.5 slabs travels alone, anything longer is met as r0 on a later round and goes in pieces
        const uint64_t *e = std::upper_bound(seq_off + r0 + 1, seq_off + n_seqs + 1, seq_off[r0] + slab_bytes);
        uint64_t r1 = (uint64_t)(e - seq_off) - 1;
        if (r1 <= r0) r1 = r0 + 1;
        if (r1 > n_seqs) r1 = n_seqs;
        const uint64_t nb = seq_off[r1] - seq_off[r0];
        slabs.push_back(Slab{r0, r1, false, false, false, seq_off[r0], seq_off[r1], seq_off[r1]});
        max_b = std::max(max_b, nb);
        max_n = std::max(max_n, r1 - r0);
        r0 = r1;
    }
    if (any_piece) {
        if ((rc = ensure(ctx, ctx->d_piece, 64, false))) return rc;
        if ((rc = ensure(ctx, ctx->h_piece, 64, true))) return rc;
        for (int i = 0; i < 3; ++i) if ((rc = ensure(ctx, ctx->h_off_stage[i], 16, true))) return rc;
    }
    const size_t n_slabs = slabs.size();
    for (int i = 0; i < 3; ++i) {
        if ((rc = ensure(ctx, ctx->d_in[i], max_b + 16, false))) return rc;
        if ((rc = ensure(ctx, ctx->d_in_off[i], (max_n + 1) * 8, false))) return rc;
    }
    if ((rc = ensure(ctx, ctx->h_km_off, (n_seqs + 1) * 8, true))) return rc;
    if ((rc = ensure(ctx, ctx->h_min_off, (n_seqs + 1) * 8, true))) return rc;
    if ((rc = ensure(ctx, ctx->h_min_cnt, std::max<uint64_t>(n_seqs, 1) * 4, true))) return rc;

    // ---- 2-bit transport: a share of the slabs is packed 4 bases/byte by host threads (pinned ring of 3 staging
    // buffers), copied (a quarter of the bytes) and unpacked on the device; the other slabs go as plain ASCII so that
    // PCIe and the packers work at the same time.  A slab with any byte outside upper-case ACGT goes as ASCII.
    int T = ctx->host_threads > 0 ? ctx->host_threads : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency() * 3 / 4));
    if (fx && fx_threads > 0) T = fx_threads;
    const bool can_pack = fx || (!packed_in && host_has_avx512() && ctx->pack_ratio > 0.0 && n_slabs >= 3);
    std::vector<int> ps_of(n_slabs, -1);
    std::vector<size_t> packed_slabs;
    if (can_pack) {
        double acc = 0.0;
        for (size_t s = 0; s < n_slabs; ++s) {
            acc += fx ? 1.0 : std::min(1.0, ctx->pack_ratio);
            if (acc >= 1.0 - 1e-9) { acc -= 1.0; ps_of[s] = (int)packed_slabs.size(); packed_slabs.push_back(s); }
        }
    }
    const size_t np = packed_slabs.size();
    if (fx) for (int i = 0; i < 3; ++i) if ((rc = ensure(ctx, ctx->h_ascii[i], 256, true))) return rc;   // grown on demand
    if (np) {
        for (int i = 0; i < 3; ++i) if ((rc = ensure(ctx, ctx->d_pack[i], max_b / 4 + 64, false))) return rc;
        for (int i = 0; i < 3; ++i) if ((rc = ensure(ctx, ctx->h_pack[i], max_b / 4 + 64, true))) return rc;
    }
    if (packed_in)
        for (int i = 0; i < 3; ++i) if ((rc = ensure(ctx, ctx->d_pack[i], max_b / 4 + 64, false))) return rc;
    std::vector<std::atomic<uint32_t>> pk_done(np);
    std::vector<std::atomic<uint64_t>> pk_bad(np);
    for (size_t i = 0; i < np; ++i) { pk_done[i].store(0); pk_bad[i].store(0); }
    std::vector<int> pk_state(np, 0);                      // 0 pending, 1 copy issued (ev_pack recorded), 2 buffer not used
    std::atomic<int64_t> pk_released{0};                   // staging buffers of packed slabs [0, released) are free again
    std::atomic<bool> pk_abort{false};
    struct Joiner {
        std::vector<std::thread> th; std::atomic<bool> *abort;
        ~Joiner() { abort->store(true); for (auto &t : th) t.join(); }
    } joiner{{}, &pk_abort};
    if (np) {
        for (int wi = 0; wi < T; ++wi)
            joiner.th.emplace_back([&, wi]() {
                for (size_t ps = 0; ps < np; ++ps) {
                    while ((int64_t)ps >= pk_released.load(std::memory_order_acquire) + 3) {
                        if (pk_abort.load()) return;
                        std::this_thread::yield();
                    }
                    const size_t sl = packed_slabs[ps];
                    const uint64_t b0 = slabs[sl].b0, nb = slabs[sl].hi - b0;
                    const uint64_t chunk = ((nb + T - 1) / T + 63) & ~uint64_t(63);
                    const uint64_t lo = std::min<uint64_t>(nb, chunk * wi), hi = std::min<uint64_t>(nb, lo + chunk);
                    if (hi > lo && fx) {
                        // gather 16 K bases at a time into a cache-resident block (line ends and headers drop out), pack it
                        uint8_t blk[16384];
                        uint64_t bad = 0;
                        for (uint64_t q = lo; q < hi; q += sizeof(blk)) {
                            const uint64_t e = std::min<uint64_t>(hi, q + sizeof(blk));
                            fx_gather(*fx, b0 + q, b0 + e, blk);
                            bad += pack2_range(blk, ptr<uint8_t>(ctx->h_pack[ps % 3]) + (q >> 2), 0, e - q);
                        }
                        pk_bad[ps].fetch_add(bad);
                    } else if (hi > lo) pk_bad[ps].fetch_add(pack2_range(bases + b0, ptr<uint8_t>(ctx->h_pack[ps % 3]), lo, hi));
                    pk_done[ps].fetch_add(1, std::memory_order_release);
                }
            });
    }
    auto poll_release = [&]() {
        int64_t r = pk_released.load();
        while ((size_t)r < np && (pk_state[r] == 2 || (pk_state[r] == 1 && cudaEventQuery(ctx->ev_pack[r % 3]) == cudaSuccess))) ++r;
        pk_released.store(r, std::memory_order_release);
    };

    auto issue_h2d = [&](size_t s) -> int {
        const int b = (int)(s % 3);
        const Slab &L = slabs[s];
        const uint64_t r0 = L.r0, r1 = L.r1, nb = L.hi - L.b0;
        const int ps = ps_of[s];
        bool packed = false;
        if (ps >= 0) {
            STAGE(100000 + (long)s); g_aux.store(ps * 1000 + (long)pk_released.load());
            const auto t_wait = std::chrono::steady_clock::now();
            while (pk_done[ps].load(std::memory_order_acquire) < (uint32_t)T) {
                poll_release();
                std::this_thread::yield();
                if (std::chrono::steady_clock::now() - t_wait > std::chrono::seconds(120))      // never seen; fail, do not hang
                    return fail(ctx, S2K_ERR_INTERNAL, "2-bit transport: host packers stalled");
            }
            packed = nb > 0 && pk_bad[ps].load() == 0;
        }
        STAGE(200000 + (long)s);
        CU(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_free[b], 0));          // the kernels of slab s-3 are done with it
        ctx->tr_h2d_bytes += (packed ? (nb + 3) / 4 : nb) + (r1 - r0 + 1) * 8;
        if (packed || packed_in) ++ctx->tr_packed; else ++ctx->tr_plain;
        if (packed_in) {                                                  // the caller's packed bytes, whatever base the slab starts on
            const uint64_t by0 = L.b0 >> 2, nby = ((L.hi + 3) >> 2) - by0;
            ctx->tr_h2d_bytes -= nb; ctx->tr_h2d_bytes += nby;
            if (nby) CU(cudaMemcpyAsync(ctx->d_pack[b].p, bases + by0, nby, cudaMemcpyHostToDevice, ctx->s_h2d));
            const int g = (int)std::min<uint64_t>((nb / 16 + 256) / 256, (uint64_t)ctx->sm_count * 8);
            if (nb) S2K_LAUNCH(k_unpack2, g, 256, 0, ctx->s_h2d, false, ptr<uint32_t>(ctx->d_pack[b]), nb, ptr<uint8_t>(ctx->d_in[b]), (uint32_t)(L.b0 & 3));
            ctx->launches += 1;
        } else if (packed) {
            CU(cudaMemcpyAsync(ctx->d_pack[b].p, ctx->h_pack[ps % 3].p, (nb + 3) / 4, cudaMemcpyHostToDevice, ctx->s_h2d));
            CU(cudaEventRecord(ctx->ev_pack[ps % 3], ctx->s_h2d));
            pk_state[ps] = 1;
            const int g = (int)std::min<uint64_t>((nb / 16 + 256) / 256, (uint64_t)ctx->sm_count * 8);
            S2K_LAUNCH(k_unpack2, g, 256, 0, ctx->s_h2d, false, ptr<uint32_t>(ctx->d_pack[b]), nb, ptr<uint8_t>(ctx->d_in[b]), 0u);
            ctx->launches += 1;
        } else if (fx) {                                                  // a byte outside ACGT: this slab travels as ASCII
            if (ps >= 0) pk_state[ps] = 2;
            CU(cudaEventSynchronize(ctx->ev_in[b]));                      // the copy of slab s-3 out of this staging buffer
            if ((rc = ensure(ctx, ctx->h_ascii[b], nb + 64, true))) return rc;
            {
                std::vector<std::thread> th;
                const uint64_t chunk = (nb + T - 1) / T;
                for (int wi = 0; wi < T; ++wi)
                    th.emplace_back([&, wi]() {
                        const uint64_t lo = std::min<uint64_t>(nb, chunk * wi), hi = std::min<uint64_t>(nb, lo + chunk);
                        fx_gather(*fx, L.b0 + lo, L.b0 + hi, ptr<uint8_t>(ctx->h_ascii[b]) + lo);
                    });
                for (auto &x : th) x.join();
            }
            if (nb) CU(cudaMemcpyAsync(ctx->d_in[b].p, ctx->h_ascii[b].p, nb, cudaMemcpyHostToDevice, ctx->s_h2d));
        } else {
            if (ps >= 0) pk_state[ps] = 2;
            if (nb) CU(cudaMemcpyAsync(ctx->d_in[b].p, bases + L.b0, nb, cudaMemcpyHostToDevice, ctx->s_h2d));
        }
        poll_release();
        uint64_t *d_off = ptr<uint64_t>(ctx->d_in_off[b]);
        if (L.piece) {                                                    // the piece is a sequence of its own: {0, nb}
            uint64_t *h_off = ptr<uint64_t>(ctx->h_off_stage[b]);         // free again: slab s-3 has been through the kernels
            h_off[0] = 0; h_off[1] = nb;
            CU(cudaMemcpyAsync(d_off, h_off, 16, cudaMemcpyHostToDevice, ctx->s_h2d));
        } else {
            // offsets: copy the caller's slice as it is, rebase it to the slab on the device
            CU(cudaMemcpyAsync(d_off, seq_off + r0, (r1 - r0 + 1) * 8, cudaMemcpyHostToDevice, ctx->s_h2d));
            if (r0) {
                const int g = (int)std::min<uint64_t>((r1 - r0 + 256) / 256, (uint64_t)ctx->sm_count * 4);
                S2K_LAUNCH(k_sub_first, g, 256, 0, ctx->s_h2d, false, d_off, r1 - r0 + 1);
                S2K_LAUNCH(k_zero_first, 1, 1, 0, ctx->s_h2d, false, d_off);
                ctx->launches += 2;
            }
        }
        CU(cudaEventRecord(ctx->ev_in[b], ctx->s_h2d));
        return S2K_OK;
    };

    uint64_t items = 0, mins = 0;
    ctx->tr_h2d_bytes = 0; ctx->tr_packed = 0; ctx->tr_plain = 0;
    for (int i = 0; i < 3; ++i) CU(cudaEventRecord(ctx->ev_free[i], st));
    CU(cudaEventRecord(ctx->ev_out, ctx->s_d2h));
    // Two slabs are kept in flight behind the one being computed: run_device ends with a host synchronisation, and
    // with a single slab queued the copy engine would idle from then until the next issue.
    if ((rc = issue_h2d(0))) return rc;
    if (n_slabs > 1 && (rc = issue_h2d(1))) return rc;
    uint64_t seq_cnt = 0, kept_plain = 0;                  // running values of the long sequence being pieced together
    for (size_t s = 0; s < n_slabs; ++s) {
        const int b = (int)(s % 3);
        const Slab &L = slabs[s];
        const uint64_t r0 = L.r0, r1 = L.r1, ns = r1 - r0, nb = L.hi - L.b0;
        if (s + 2 < n_slabs) {
            if ((rc = issue_h2d(s + 2))) return rc;
        }
        CU(cudaStreamWaitEvent(st, ctx->ev_in[b], 0));                    // inputs of this slab have landed
        s2k_result dev;
        STAGE(300000 + (long)s);
        rc = run_device(ctx, ptr<uint8_t>(ctx->d_in[b]), ptr<uint64_t>(ctx->d_in_off[b]), ns, nb, L.piece ? PP : P, st, &dev);
        if (rc != S2K_OK) return rc;
        STAGE(400000 + (long)s);
        uint64_t ni = dev.n_items, nm = dev.n_minimizers;
        if (!L.piece) {
            // per-slab prefixes / sequence indices -> global
            const int g1 = (int)std::min<uint64_t>((ns + 256) / 256, (uint64_t)ctx->sm_count * 8);
            if (items) S2K_LAUNCH(k_add_u64, g1, 256, 0, st, false, const_cast<uint64_t *>(dev.km_off), ns + 1, items);
            if (mins) S2K_LAUNCH(k_add_u64, g1, 256, 0, st, false, const_cast<uint64_t *>(dev.min_off), ns + 1, mins);
            if (want_min && r0 && dev.n_minimizers) {
                const int g2 = (int)std::min<uint64_t>((dev.n_minimizers + 255) / 256, (uint64_t)ctx->sm_count * 8);
                S2K_LAUNCH(k_add_seq, g2, 256, 0, st, false, reinterpret_cast<uint4 *>(const_cast<s2k_minimizer *>(dev.minimizers)),
                           dev.n_minimizers, (uint32_t)r0);
            }
        } else {
            // ---- a piece of a long sequence: what it owns, tail rule, coordinates of the whole sequence
            unsigned long long *dp = ptr<unsigned long long>(ctx->d_piece), *hp = ptr<unsigned long long>(ctx->h_piece);
            const uint64_t s0 = seq_off[r0], s1 = seq_off[r0 + 1], own = L.b1 - L.b0, n_min = dev.n_minimizers;
            uint4 *dmins = reinterpret_cast<uint4 *>(const_cast<s2k_minimizer *>(dev.minimizers));
            if (L.first) {
                seq_cnt = 0; kept_plain = 0;
                ptr<uint64_t>(ctx->h_km_off)[r0] = items;
                ptr<uint64_t>(ctx->h_min_off)[r0] = mins;
                CU(cudaMemsetAsync(dp, 0, 8, st));
            }
            if (P.quirk) {
                if (P.hpc) {
                    const int g = (int)std::min<uint64_t>((own + 255) / 256, (uint64_t)ctx->sm_count * 16);
                    if (own) S2K_LAUNCH(k_count_kept, g, 256, 0, st, false, ptr<uint8_t>(ctx->d_in[b]), own, dp);
                } else kept_plain += own;
            }
            bool rule = false;
            uint64_t e16 = 0xffffffffull;
            if (L.last && P.quirk) {
                uint64_t kept_total = kept_plain;
                if (P.hpc) {
                    CU(cudaMemcpyAsync(hp, dp, 8, cudaMemcpyDeviceToHost, st));
                    CU(cudaStreamSynchronize(st));
                    kept_total = hp[0];
                }
                const uint64_t S = kept_total >= P.l ? kept_total - P.l + 1 : 0;
                rule = S > 16 && S % 16 == 0;
                if (rule) {                                               // 16th kept base from the end of the sequence
                    uint64_t pos = s1, left = 16;
                    while (left && pos > s0) { --pos; if (!P.hpc || pos == s0 || base_at(pos) != base_at(pos - 1)) --left; }
                    if (pos <= L.b0) { overlap_short = 2; return S2K_OK; }     // cannot happen: no cut after tail_lo
                    e16 = pos - L.b0;
                }
            }
            S2K_LAUNCH(k_piece_cut, 1, 1, 0, st, false, dmins, n_min, L.last ? 0xffffffffu : (uint32_t)own, (uint32_t)e16, dp + 2);
            CU(cudaMemcpyAsync(hp + 2, dp + 2, 16, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            ctx->launches += 2;
            const uint64_t n_keep = rule ? hp[3] : n_min;
            const uint64_t i1 = L.last ? n_keep : hp[2];
            if (!L.last && i1 + P.k + 3 > n_min) { overlap_short = 1; return S2K_OK; }   // windows near b1 lack followers
            // the last windows of the piece before this one reach k-1 minimizers into it: all of them must have survived
            if (rule && !L.first && n_keep + 1 < P.k) { overlap_short = 2; return S2K_OK; }
            const uint64_t n_win = n_keep >= P.k ? n_keep - P.k + 1 : 0;
            ni = std::min(i1, n_win);
            nm = L.last ? n_min : i1;
            seq_cnt += i1;
            const uint32_t add = (uint32_t)(L.b0 - s0);
            if (add && ni) {
                const int g = (int)std::min<uint64_t>((ni + 255) / 256, (uint64_t)ctx->sm_count * 8);
                S2K_LAUNCH(k_piece_shift_items, g, 256, 0, st, false, const_cast<uint32_t *>(dev.start), const_cast<uint32_t *>(dev.end), ni, add);
            }
            if (want_min && nm && (add || r0)) {
                const int g = (int)std::min<uint64_t>((nm + 255) / 256, (uint64_t)ctx->sm_count * 8);
                S2K_LAUNCH(k_piece_shift_mins, g, 256, 0, st, false, dmins, nm, add, (uint32_t)r0);
            }
            if (L.last) {
                ptr<uint64_t>(ctx->h_km_off)[r0 + 1] = items + ni;
                ptr<uint64_t>(ctx->h_min_off)[r0 + 1] = mins + nm;
                ptr<uint32_t>(ctx->h_min_cnt)[r0] = (uint32_t)seq_cnt;
            }
        }
        CU(cudaEventRecord(ctx->ev_free[b], st));
        // host result buffers: sized from the first slab's rates, grown (rarely) if a later slab is denser
        uint64_t need_i = items + ni, need_m = mins + nm;
        if (s == 0 && nb) {
            const double scale = (double)n_bases / (double)nb * 1.03;
            need_i = std::max<uint64_t>(need_i, (uint64_t)((double)ni * scale) + 65536);
            need_m = std::max<uint64_t>(need_m, (uint64_t)((double)nm * scale) + 65536);
        }
        if (need_i * 8 > ctx->h_hash.cap || need_i * 4 > ctx->h_start.cap || need_i * 4 > ctx->h_end.cap || need_i > ctx->h_rev.cap ||
            (want_min && need_m * 16 > ctx->h_mins.cap)) {
            CU(cudaStreamSynchronize(ctx->s_d2h));                        // earlier copies must have landed before moving
            if ((rc = ensure_keep(ctx, ctx->h_hash, std::max<uint64_t>(need_i, 1) * 8, items * 8))) return rc;
            if ((rc = ensure_keep(ctx, ctx->h_start, std::max<uint64_t>(need_i, 1) * 4, items * 4))) return rc;
            if ((rc = ensure_keep(ctx, ctx->h_end, std::max<uint64_t>(need_i, 1) * 4, items * 4))) return rc;
            if ((rc = ensure_keep(ctx, ctx->h_rev, std::max<uint64_t>(need_i, 1), items))) return rc;
            if (want_min && (rc = ensure_keep(ctx, ctx->h_mins, std::max<uint64_t>(need_m, 1) * 16, mins * 16))) return rc;
        }
        // Park the slab's results in a staging buffer (device-to-device, microseconds) and send them home from there:
        // the kernels of the next slab then overlap this slab's D2H instead of waiting for it.
        const uint64_t tail = (!L.piece && s + 1 == n_slabs) ? 1 : 0;       // the last slab also brings the final prefix
        const uint64_t n_off = L.piece ? 0 : ns + tail, n_cnt = L.piece ? 0 : ns;
        const uint64_t nmw = want_min ? nm : 0;
        const uint64_t o_hash = 0, o_mins = o_hash + ni * 8, o_km = o_mins + nmw * 16, o_mo = o_km + n_off * 8,
                       o_start = o_mo + n_off * 8, o_end = o_start + ni * 4, o_cnt = o_end + ni * 4, o_rev = o_cnt + n_cnt * 4,
                       stage_bytes = o_rev + ni;
        CU(cudaStreamWaitEvent(st, ctx->ev_out, 0));                      // the previous slab's results have left the staging buffer
        if (stage_bytes > ctx->d_stage.cap) {
            CU(cudaStreamSynchronize(ctx->s_d2h));
            if ((rc = ensure(ctx, ctx->d_stage, stage_bytes + stage_bytes / 4 + 4096, false))) return rc;
        }
        uint8_t *sg = ptr<uint8_t>(ctx->d_stage);
        if (ni) {
            CU(cudaMemcpyAsync(sg + o_hash, dev.hash, ni * 8, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(sg + o_start, dev.start, ni * 4, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(sg + o_end, dev.end, ni * 4, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(sg + o_rev, dev.rev, ni, cudaMemcpyDeviceToDevice, st));
        }
        if (nmw) CU(cudaMemcpyAsync(sg + o_mins, dev.minimizers, nmw * 16, cudaMemcpyDeviceToDevice, st));
        if (n_off) {
            CU(cudaMemcpyAsync(sg + o_km, dev.km_off, n_off * 8, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(sg + o_mo, dev.min_off, n_off * 8, cudaMemcpyDeviceToDevice, st));
        }
        if (n_cnt) CU(cudaMemcpyAsync(sg + o_cnt, dev.min_cnt, n_cnt * 4, cudaMemcpyDeviceToDevice, st));
        CU(cudaEventRecord(ctx->ev_done, st));
        cudaStream_t so = ctx->s_d2h;
        CU(cudaStreamWaitEvent(so, ctx->ev_done, 0));
        if (ni) {
            CU(cudaMemcpyAsync(ptr<uint64_t>(ctx->h_hash) + items, sg + o_hash, ni * 8, cudaMemcpyDeviceToHost, so));
            CU(cudaMemcpyAsync(ptr<uint32_t>(ctx->h_start) + items, sg + o_start, ni * 4, cudaMemcpyDeviceToHost, so));
            CU(cudaMemcpyAsync(ptr<uint32_t>(ctx->h_end) + items, sg + o_end, ni * 4, cudaMemcpyDeviceToHost, so));
            CU(cudaMemcpyAsync(ptr<uint8_t>(ctx->h_rev) + items, sg + o_rev, ni, cudaMemcpyDeviceToHost, so));
        }
        if (n_off) {
            CU(cudaMemcpyAsync(ptr<uint64_t>(ctx->h_km_off) + r0, sg + o_km, n_off * 8, cudaMemcpyDeviceToHost, so));
            CU(cudaMemcpyAsync(ptr<uint64_t>(ctx->h_min_off) + r0, sg + o_mo, n_off * 8, cudaMemcpyDeviceToHost, so));
        }
        if (n_cnt) CU(cudaMemcpyAsync(ptr<uint32_t>(ctx->h_min_cnt) + r0, sg + o_cnt, n_cnt * 4, cudaMemcpyDeviceToHost, so));
        if (nmw) CU(cudaMemcpyAsync(ptr<uint8_t>(ctx->h_mins) + mins * 16, sg + o_mins, nmw * 16, cudaMemcpyDeviceToHost, so));
        CU(cudaEventRecord(ctx->ev_out, so));
        items += ni; mins += nm;
        poll_release();
    }
    STAGE(500000);
    CU(cudaStreamSynchronize(ctx->s_d2h));
    CU(cudaStreamSynchronize(ctx->s_h2d));
    STAGE(600000);
    std::memset(out, 0, sizeof(*out));
    out->n_seqs = n_seqs; out->n_items = items; out->n_minimizers = mins;
    out->location = S2K_LOC_HOST;
    out->hash = ptr<uint64_t>(ctx->h_hash);
    out->start = ptr<uint32_t>(ctx->h_start);
    out->end = ptr<uint32_t>(ctx->h_end);
    out->rev = ptr<uint8_t>(ctx->h_rev);
    out->km_off = ptr<uint64_t>(ctx->h_km_off);
    out->min_off = ptr<uint64_t>(ctx->h_min_off);
    out->min_cnt = ptr<uint32_t>(ctx->h_min_cnt);
    out->minimizers = want_min ? reinterpret_cast<const s2k_minimizer *>(ctx->h_mins.p) : nullptr;
    return S2K_OK;
}

static int run_host(s2k_ctx *ctx, const uint8_t *bases, const uint64_t *seq_off, uint64_t n_seqs, const s2k_params *params,
                    s2k_result *out, bool packed_in)
{
    if (!ctx) return S2K_ERR_NULL;
    if (!out || !seq_off) return fail(ctx, S2K_ERR_NULL, "null argument");
    Plan P;
    int rc = make_plan(ctx, params, P);
    if (rc != S2K_OK) return rc;
    if ((rc = check_offsets(ctx, seq_off, n_seqs)) != S2K_OK) return rc;
    const uint64_t n_bases = seq_off[n_seqs];
    if (n_bases && !bases) return fail(ctx, S2K_ERR_NULL, "bases is null");
    CU(cudaSetDevice(ctx->device));
    ctx->err.clear();
    cudaStream_t st = ctx->stream;
    {
        const uint64_t slab = ctx->slab_bytes ? ctx->slab_bytes : (256ull << 20);
        if (n_bases > slab + slab / 2 && n_seqs >= 1) {
            // right overlap of a piece of a long sequence: room for k-1 further minimizers at the selection rate, with a
            // wide margin (checked per piece; grown and redone if some stretch of the sequence is poorer than that)
            const double frac = std::max(1e-9, selection_fraction(P));
            uint64_t overlap = (uint64_t)std::min(4.0e9, 64.0 * (P.k + 8.0) / frac + 64.0 * P.l + 4096.0);
            for (int attempt = 0; attempt < 3; ++attempt) {
                int too_short = 0;
                rc = run_pipelined(ctx, bases, seq_off, n_seqs, P, slab, overlap, too_short, out, packed_in);
                if (rc != S2K_OK || !too_short) {
                    // on an error, copies from the caller's buffers may still be queued: do not hand the buffers back early
                    if (rc != S2K_OK && ctx->pipe_ready) { cudaStreamSynchronize(ctx->s_h2d); cudaStreamSynchronize(ctx->s_d2h); cudaStreamSynchronize(st); }
                    return rc;
                }
                STAGE(700000 + attempt);
                CU(cudaStreamSynchronize(ctx->s_h2d)); CU(cudaStreamSynchronize(ctx->s_d2h)); CU(cudaStreamSynchronize(st));
                if (too_short == 2) break;                 // a longer overlap would not help
                overlap = std::min<uint64_t>(n_bases, overlap * 8);
            }                                              // still not enough: the whole batch at once, below
        }
    }
    if ((rc = ensure(ctx, ctx->d_bases, n_bases + 32, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_seq_off, (n_seqs + 1) * 8, false))) return rc;
    if (packed_in) {
        const uint64_t nby = (n_bases + 3) / 4;
        if ((rc = ensure(ctx, ctx->d_pack[0], nby + 64, false))) return rc;
        if (nby) CU(cudaMemcpyAsync(ctx->d_pack[0].p, bases, nby, cudaMemcpyHostToDevice, st));
        const int g = (int)std::min<uint64_t>((n_bases / 16 + 256) / 256, (uint64_t)ctx->sm_count * 8);
        if (n_bases) S2K_LAUNCH(k_unpack2, g, 256, 0, st, false, ptr<uint32_t>(ctx->d_pack[0]), n_bases, ptr<uint8_t>(ctx->d_bases), 0u);
        ctx->launches += 1;
        ctx->tr_h2d_bytes = nby + (n_seqs + 1) * 8; ctx->tr_packed = 1; ctx->tr_plain = 0;
    } else {
        if (n_bases) CU(cudaMemcpyAsync(ctx->d_bases.p, bases, n_bases, cudaMemcpyHostToDevice, st));
        ctx->tr_h2d_bytes = n_bases + (n_seqs + 1) * 8; ctx->tr_packed = 0; ctx->tr_plain = 1;
    }
    CU(cudaMemcpyAsync(ctx->d_seq_off.p, seq_off, (n_seqs + 1) * 8, cudaMemcpyHostToDevice, st));
    s2k_result dev;
    rc = run_device(ctx, ptr<uint8_t>(ctx->d_bases), ptr<uint64_t>(ctx->d_seq_off), n_seqs, n_bases, P, st, &dev);
    if (rc != S2K_OK) return rc;
    const uint64_t ni = dev.n_items, nm = dev.n_minimizers;
    if ((rc = ensure(ctx, ctx->h_hash, std::max<uint64_t>(ni, 1) * 8, true))) return rc;
    if ((rc = ensure(ctx, ctx->h_start, std::max<uint64_t>(ni, 1) * 4, true))) return rc;
    if ((rc = ensure(ctx, ctx->h_end, std::max<uint64_t>(ni, 1) * 4, true))) return rc;
    if ((rc = ensure(ctx, ctx->h_rev, std::max<uint64_t>(ni, 1), true))) return rc;
    if ((rc = ensure(ctx, ctx->h_km_off, (n_seqs + 1) * 8, true))) return rc;
    if ((rc = ensure(ctx, ctx->h_min_off, (n_seqs + 1) * 8, true))) return rc;
    if ((rc = ensure(ctx, ctx->h_min_cnt, std::max<uint64_t>(n_seqs, 1) * 4, true))) return rc;
    if (ni) {
        CU(cudaMemcpyAsync(ctx->h_hash.p, dev.hash, ni * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ctx->h_start.p, dev.start, ni * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ctx->h_end.p, dev.end, ni * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ctx->h_rev.p, dev.rev, ni, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaMemcpyAsync(ctx->h_km_off.p, dev.km_off, (n_seqs + 1) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_min_off.p, dev.min_off, (n_seqs + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (n_seqs) CU(cudaMemcpyAsync(ctx->h_min_cnt.p, dev.min_cnt, n_seqs * 4, cudaMemcpyDeviceToHost, st));
    const bool want_min = (ctx->flags & S2K_WANT_MINIMIZERS) != 0;
    if (want_min) {
        if ((rc = ensure(ctx, ctx->h_mins, std::max<uint64_t>(nm, 1) * 16, true))) return rc;
        if (nm) CU(cudaMemcpyAsync(ctx->h_mins.p, dev.minimizers, nm * 16, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    *out = dev;
    out->location = S2K_LOC_HOST;
    out->hash = ptr<uint64_t>(ctx->h_hash);
    out->start = ptr<uint32_t>(ctx->h_start);
    out->end = ptr<uint32_t>(ctx->h_end);
    out->rev = ptr<uint8_t>(ctx->h_rev);
    out->km_off = ptr<uint64_t>(ctx->h_km_off);
    out->min_off = ptr<uint64_t>(ctx->h_min_off);
    out->min_cnt = ptr<uint32_t>(ctx->h_min_cnt);
    out->minimizers = want_min ? reinterpret_cast<const s2k_minimizer *>(ctx->h_mins.p) : nullptr;
    return S2K_OK;
}

int s2k_run(s2k_ctx *ctx, const uint8_t *bases, const uint64_t *seq_off, uint64_t n_seqs, const s2k_params *params,
            s2k_result *out)
{
    return run_host(ctx, bases, seq_off, n_seqs, params, out, false);
}

int s2k_run_packed2(s2k_ctx *ctx, const uint8_t *packed, const uint64_t *seq_off, uint64_t n_seqs, const s2k_params *params,
                    s2k_result *out)
{
    return run_host(ctx, packed, seq_off, n_seqs, params, out, true);
}

int64_t s2k_pack2(const uint8_t *bases, uint64_t n_bases, uint8_t *packed_out, int host_threads)
{
    if ((!bases || !packed_out) && n_bases) return -1;
    const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(host_threads, 1), n_bases / (1u << 20) + 1));
    std::vector<uint64_t> bad((size_t)T, 0);
    std::vector<std::thread> th;
    const uint64_t chunk = ((n_bases + T - 1) / T + 63) & ~uint64_t(63);
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
            const uint64_t lo = std::min<uint64_t>(n_bases, chunk * t), hi = std::min<uint64_t>(n_bases, lo + chunk);
            if (hi > lo) bad[(size_t)t] = pack2_range(bases, packed_out, lo, hi);
        });
    for (auto &t : th) t.join();
    uint64_t total = 0;
    for (uint64_t b : bad) total += b;
    return total ? 1 : 0;
}

int s2k_run_fastx(s2k_ctx *ctx, const char *path, int nb_threads, const s2k_params *params, s2k_result *out)
{
    if (!ctx) return S2K_ERR_NULL;
    if (!path || !out) return fail(ctx, S2K_ERR_NULL, "null argument");
    if (nb_threads < 1) nb_threads = 1;
    FxFile f;
    f.fd = open(path, O_RDONLY);
    if (f.fd < 0) return fail(ctx, S2K_ERR_IO, std::string("cannot open ") + path);
    struct stat st;
    if (fstat(f.fd, &st) != 0) return fail(ctx, S2K_ERR_IO, std::string("cannot stat ") + path);
    f.n = (size_t)st.st_size;
    if (f.n) {
        void *m = mmap(nullptr, f.n, PROT_READ, MAP_PRIVATE, f.fd, 0);
        if (m == MAP_FAILED) { f.n = 0; return fail(ctx, S2K_ERR_IO, std::string("cannot map ") + path); }
        f.p = (const char *)m;
    }
    size_t first = 0;
    while (first < f.n && (f.p[first] == '\n' || f.p[first] == '\r' || f.p[first] == ' ')) ++first;
    if (first < f.n && f.p[first] != '>' && f.p[first] != '@')
        return fail(ctx, S2K_ERR_IO, "not a FASTA/FASTQ file (first record must start with '>' or '@')");
    const bool fastq = first < f.n && f.p[first] == '@';
    const int T = (int)std::min<size_t>((size_t)nb_threads, std::max<size_t>(1, f.n >> 16));
    const bool fx_timing = getenv("S2K_FASTX_TIMING") != nullptr;
    const auto fx_t0 = std::chrono::steady_clock::now();
    auto fx_lap = [&](const char *what) {
        if (fx_timing) fprintf(stderr, "[s2k fastx] %-28s at %.1f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - fx_t0).count());
    };
    std::vector<std::vector<FxRec>> recs((size_t)T);
    std::vector<uint64_t> nb((size_t)T, 0);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t]() { fx_scan(f.p, f.n, f.n * (size_t)t / T, f.n * (size_t)(t + 1) / T, fastq, recs[t], nb[t]); });
        for (auto &x : th) x.join();
    }
    fx_lap("records indexed");
    uint64_t n_seqs = 0, n_bases = 0;
    std::vector<uint64_t> seq_base((size_t)T), byte_base((size_t)T);
    for (int t = 0; t < T; ++t) { seq_base[t] = n_seqs; byte_base[t] = n_bases; n_seqs += recs[t].size(); n_bases += nb[t]; }
    int rc;
    CU(cudaSetDevice(ctx->device));
    if ((rc = ensure(ctx, ctx->h_fx_off, (n_seqs + 1) * 8, true))) return rc;
    uint64_t *ho = ptr<uint64_t>(ctx->h_fx_off);
    ctx->fx_n_seqs = n_seqs; ctx->fx_n_bases = n_bases; ctx->fx_have_bases = false;

    // ---- streaming form: large files whose records are addressable in O(1) never exist as one ASCII batch on the host.
    // The host threads gather and pack slab i+1 straight from the mapping while slab i is on the device (run_pipelined).
    const uint64_t slab = ctx->slab_bytes ? ctx->slab_bytes : (256ull << 20);
    bool addressable = true;
    for (int t = 0; t < T && addressable; ++t)
        for (const FxRec &r : recs[t]) if (r.n && r.width == 0) { addressable = false; break; }
    const bool keep = (ctx->flags & S2K_FASTX_KEEP_BASES) != 0;
    if (addressable && !keep && n_bases > slab + slab / 2) {
        std::vector<FxRec> all;
        all.reserve(n_seqs);
        for (int t = 0; t < T; ++t) all.insert(all.end(), recs[t].begin(), recs[t].end());
        uint64_t o = 0;
        for (uint64_t i = 0; i < n_seqs; ++i) { ho[i] = o; o += all[i].n; }
        ho[n_seqs] = n_bases;
        Plan P;
        if ((rc = make_plan(ctx, params, P)) != S2K_OK) return rc;
        if ((rc = check_offsets(ctx, ho, n_seqs)) != S2K_OK) return rc;
        ctx->err.clear();
        FxSource src;
        src.file = f.p; src.recs = all.data(); src.seq_off = ho; src.n_recs = n_seqs;
        const double frac = std::max(1e-9, selection_fraction(P));
        uint64_t overlap = (uint64_t)std::min(4.0e9, 64.0 * (P.k + 8.0) / frac + 64.0 * P.l + 4096.0);
        fx_lap("offsets built");
        for (int attempt = 0; attempt < 3; ++attempt) {
            int too_short = 0;
            rc = run_pipelined(ctx, nullptr, ho, n_seqs, P, slab, overlap, too_short, out, false, &src, nb_threads);
            fx_lap("streamed through the device");
            if (rc != S2K_OK || !too_short) {
                if (rc != S2K_OK && ctx->pipe_ready) { cudaStreamSynchronize(ctx->s_h2d); cudaStreamSynchronize(ctx->s_d2h); cudaStreamSynchronize(ctx->stream); }
                return rc;
            }
            CU(cudaStreamSynchronize(ctx->s_h2d)); CU(cudaStreamSynchronize(ctx->s_d2h)); CU(cudaStreamSynchronize(ctx->stream));
            if (too_short == 2) break;
            overlap = std::min<uint64_t>(n_bases, overlap * 8);
        }                                                  // a long sequence that defeats the piece logic: materialise it, below
    }

    // ---- materialised form: the records copied to one pinned ASCII batch (small files, uneven lines, or on request)
    if ((rc = ensure(ctx, ctx->h_fx_bases, n_bases + 64, true))) return rc;
    uint8_t *hb = ptr<uint8_t>(ctx->h_fx_bases);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t]() {
                uint64_t o = byte_base[t], i = seq_base[t];
                for (const FxRec &r : recs[t]) { ho[i++] = o; fx_copy(f.p, r, hb + o, fastq); o += r.n; }
            });
        for (auto &x : th) x.join();
    }
    ho[n_seqs] = n_bases;
    ctx->fx_have_bases = true;
    return s2k_run(ctx, hb, ho, n_seqs, params, out);
}

int s2k_last_fastx(const s2k_ctx *ctx, uint64_t *n_seqs, uint64_t *n_bases, const uint8_t **bases, const uint64_t **seq_off)
{
    if (!ctx) return S2K_ERR_NULL;
    if (n_seqs) *n_seqs = ctx->fx_n_seqs;
    if (n_bases) *n_bases = ctx->fx_n_bases;
    if (bases) *bases = ctx->fx_have_bases ? reinterpret_cast<const uint8_t *>(ctx->h_fx_bases.p) : nullptr;
    if (seq_off) *seq_off = reinterpret_cast<const uint64_t *>(ctx->h_fx_off.p);
    return S2K_OK;
}

int s2k_encode_rle(s2k_ctx *ctx, const uint8_t *bases, const uint64_t *seq_off, uint64_t n_seqs, s2k_rle_result *out)
{
    if (!ctx) return S2K_ERR_NULL;
    if (!out || !seq_off) return fail(ctx, S2K_ERR_NULL, "null argument");
    int rc;
    if ((rc = check_offsets(ctx, seq_off, n_seqs)) != S2K_OK) return rc;
    const uint64_t n_bases = seq_off[n_seqs];
    if (n_bases && !bases) return fail(ctx, S2K_ERR_NULL, "bases is null");
    CU(cudaSetDevice(ctx->device));
    ctx->err.clear();
    cudaStream_t st = ctx->stream;
    std::memset(out, 0, sizeof(*out));
    out->n_seqs = n_seqs; out->location = S2K_LOC_HOST;
    if ((rc = ensure(ctx, ctx->h_min_off, (n_seqs + 1) * 8, true))) return rc;
    if (n_bases == 0) {
        std::memset(ctx->h_min_off.p, 0, (n_seqs + 1) * 8);
        out->hpc_off = ptr<uint64_t>(ctx->h_min_off);
        return S2K_OK;
    }
    const uint64_t n_tiles64 = (n_bases + RLE_TILE - 1) / RLE_TILE;
    if (n_tiles64 > 0x7fffffffull) return fail(ctx, S2K_ERR_BAD_PARAM, "batch too large for s2k_encode_rle");
    const uint32_t n_tiles = (uint32_t)n_tiles64;
    if ((rc = ensure(ctx, ctx->d_bases, n_bases + 16, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_seq_off, (n_seqs + 1) * 8, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_small, 64, false))) return rc;
    if ((rc = ensure(ctx, ctx->h_small, 64, true))) return rc;
    if ((rc = ensure(ctx, ctx->d_tile_lb, ((uint64_t)n_tiles + 1) * 4, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_status, (uint64_t)n_tiles * 8 + 8, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_rle_hpc, n_bases, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_rle_pos, n_bases * 4, false))) return rc;
    if ((rc = ensure(ctx, ctx->d_hpc_off, (n_seqs + 1) * 8, false))) return rc;
    CU(cudaMemcpyAsync(ctx->d_bases.p, bases, n_bases, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_seq_off.p, seq_off, (n_seqs + 1) * 8, cudaMemcpyHostToDevice, st));
    uint64_t *small = ptr<uint64_t>(ctx->d_small);
    CU(cudaMemsetAsync(small, 0, 64, st));
    CU(cudaMemsetAsync(ctx->d_status.p, 0, (uint64_t)n_tiles * 8, st));
    S2K_LAUNCH(k_tile_bounds, (n_tiles + 1 + 255) / 256, 256, 0, st, false, ptr<uint64_t>(ctx->d_seq_off), n_seqs,
               n_bases, (uint32_t)RLE_TILE, n_tiles, ptr<uint32_t>(ctx->d_tile_lb));
    K4Args A;
    A.bases = ptr<uint8_t>(ctx->d_bases); A.seq_off = ptr<uint64_t>(ctx->d_seq_off);
    A.tile_lb = ptr<uint32_t>(ctx->d_tile_lb); A.n_seqs = n_seqs; A.n_bases = n_bases; A.n_tiles = n_tiles;
    A.status = ptr<uint64_t>(ctx->d_status); A.ticket = reinterpret_cast<uint32_t *>(small + 4);
    A.err = reinterpret_cast<uint32_t *>(small + 5);
    A.hpc = ptr<uint8_t>(ctx->d_rle_hpc); A.pos = ptr<uint32_t>(ctx->d_rle_pos); A.hpc_off = ptr<uint64_t>(ctx->d_hpc_off);
    A.scalar_rule = (ctx->flags & S2K_RLE_SCALAR_RULE) ? 1u : 0u;
    timing_prepare(ctx);
    Timing &T = ctx->tm;
    if (T.enabled) cudaEventRecord(T.ev[0][0], st);
    S2K_LAUNCH(k_rle, (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count * 8), NT, 0, st, false, A);
    if (T.enabled) { cudaEventRecord(T.ev[0][1], st); T.n = 1; }
    CU(cudaGetLastError());
    ctx->launches += 2;
    CU(cudaMemcpyAsync(ctx->h_min_off.p, ctx->d_hpc_off.p, (n_seqs + 1) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_small.p, small + 5, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if ((uint32_t)ptr<uint64_t>(ctx->h_small)[0] & ERR_SPIN) return fail(ctx, S2K_ERR_INTERNAL, "tile look-back timed out");
    const uint64_t n_hpc = ptr<uint64_t>(ctx->h_min_off)[n_seqs];
    if ((rc = ensure(ctx, ctx->h_rle_hpc, std::max<uint64_t>(n_hpc, 1), true))) return rc;
    if ((rc = ensure(ctx, ctx->h_rle_pos, std::max<uint64_t>(n_hpc, 1) * 4, true))) return rc;
    CU(cudaMemcpyAsync(ctx->h_rle_hpc.p, ctx->d_rle_hpc.p, n_hpc, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_rle_pos.p, ctx->d_rle_pos.p, n_hpc * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (ctx->tm.enabled && ctx->tm.n == 1) {               // s2k_last_kernel_ms: minimizer_ms = k_rle here
        float f = 0; cudaEventElapsedTime(&f, ctx->tm.ev[0][0], ctx->tm.ev[0][1]);
        ctx->tm.min_ms = f; ctx->tm.win_ms = 0; ctx->tm.min_launches = 1;
    }
    out->n_hpc = n_hpc;
    out->hpc = ptr<uint8_t>(ctx->h_rle_hpc);
    out->pos = ptr<uint32_t>(ctx->h_rle_pos);
    out->hpc_off = ptr<uint64_t>(ctx->h_min_off);
    return S2K_OK;
}

} // extern "C"
