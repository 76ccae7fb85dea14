// s2k_kernels.cuh -- sm_100a kernels of the sequence -> k-min-mer path.
//
// Pipeline per launch sequence (all on one stream, no host round trip in between):
//   k_tile_bounds    : per tile, index of the first sequence starting at/after the tile      (tiny)
//   k_minimizers     : fused HPC keep-mask + in-smem compaction + canonical ntHash (32/31 bit) of every
//                      l-mer in HPC space + density threshold + append of (hash,start,end,seq) records to a
//                      per-CTA region, records of a tile contiguous and ordered            (the hot kernel)
//   k_tile_scan_a/b, k_finalize : exclusive prefix of per-tile counts -> where every tile's records belong in the
//                      ordered stream (tile_src), global per-sequence prefixes                        (per tile)
//   k_read_counts    : per sequence, minimizers feeding the window stage (AVX-512 tail rule) -> item counts,
//                      exclusive scan -> km_off                                              (n_seqs elements)
//   k_windows        : one thread per minimizer: k-window hash, canonical min, rev, start/end (~2*rho*d*N elements)
//
// Semantics follow SURVEY.md Appendix A; reference citations (file:line) are into rchikhi/rust-seq2kminmers:
//   keep mask            src/hpc.rs:86-95           (byte != previous byte, first byte of a sequence kept)
//   seeds / base classes src/nthash_hpc.rs:29-49    (scalar 256-entry tables: ACGT, N->0, other->1)
//                        src/nthash_avx512_32.rs:178-193,242-277 (low-nibble map, non-ACGT -> 0)
//   rolling update       src/nthash_hpc.rs:245-249, src/nthash_avx512_32.rs:348-509 (same recurrence)
//   31-bit variant       src/nthash2_avx512_32.rs:186-215,226-268
//   selection            `<= bound` src/nthash_hpc.rs:232,277, src/lib.rs:228; `< bound` src/nthash_avx512_32.rs:55,130
//   coordinates          Hpc: (run start of first base, run END of last base) src/nthash_hpc.rs:234,281
//                        HpcSimd: (pos[p], pos[p+l-1]) src/nthash_hpc_simd.rs:64; Simd/Regular: (p, p+l-1) src/lib.rs:202,226
//   final l-mer dropped  src/nthash_hpc.rs:220-222,265-267 (Hpc only)
//   len <= l -> nothing  src/lib.rs:97
//   tail rule            src/nthash_avx512_32.rs:134-138 (last 16 l-mers dropped when S>16 && S%16==0)
//   window stage         src/lib.rs:157-169 (mix), 231-261 (rolling k-window hash, canonical, rev, start/end/offset)
#pragma once
#include <cstddef>
#include <cstdint>
#ifdef S2K_EMU            // tests/emu: the same sources compiled by g++ onto host threads (test tier only)
#include "cuda_emu.h"
#define S2K_DYN_SMEM(name) uint8_t *name = emu::blk->smem
#define S2K_SHARED static
#else
#include <cuda_runtime.h>
#define S2K_DYN_SMEM(name) extern __shared__ __align__(1024) uint8_t name[]
#define S2K_SHARED __shared__
#endif

namespace s2k {

// ------------------------------------------------------------------------------------------------ geometry
#ifndef S2K_NT
#define S2K_NT 256
#endif
constexpr int NT    = S2K_NT;       // threads per CTA
constexpr int RAWPT = 64;           // raw bases per thread in the keep/compaction phase: four pieces of 16
constexpr int WIN   = NT * RAWPT;   // raw bases staged per tile (left halo + tile)
constexpr int NCHUNK = WIN / 32;    // 32-base chunks per window
#ifndef S2K_CH
#define S2K_CH 48
#endif
#ifndef S2K_HT
#define S2K_HT S2K_NT
#endif
constexpr int CH    = S2K_CH;       // owner positions per hash thread.  48 = three whole groups of 16 in the packed hash
                                    // stage and 12 288 owners per pass: an HPC tile of 16 128 random bases keeps
                                    // 12 096 +- 55, so one tile in 4 000 takes a second pass.  (With 52 the 9 % of idle
                                    // owner slots cost 3 % of the kernel; the byte form, whose word streams want CH/4
                                    // odd to avoid bank conflicts, only serves rare tiles.)
constexpr int HT    = S2K_HT;       // threads of the CTA that hash (the others wait at the barrier meanwhile)
constexpr int MW    = (CH + 63) / 64;   // 64-bit words of a thread's hit mask
constexpr int CAP   = HT * CH;      // owners hashed per pass (a second pass covers tiles that compress badly)
static_assert(CH % 4 == 0 && CH <= 128 && HT <= NT && 2 * CAP >= WIN, "hash geometry");
constexpr int XB    = 256;          // capacity of the left context, in kept (HPC) bases
constexpr int FW    = (XB + WIN) / 32 + 2;   // words of the owner-space flag bitmaps
constexpr int PKW   = (XB + WIN) / 16 + 8;   // words of the packed code array (a multiple of 4)
constexpr int PK_LMAX = 33;                  // packed tiles serve l <= 33: the warm-up reads 32 codes from two registers
#ifndef S2K_HL
#define S2K_HL 1024
#endif
constexpr int HL    = S2K_HL;         // hit-list entries emitted per round
constexpr int DIRTY_MAX = 62;
#ifndef S2K_SOC
#define S2K_SOC 128
#endif
constexpr int SOC   = S2K_SOC;          // sequence offsets cached per tile (tiles with more starts search global memory)
// Base classes are stored pre-scaled by 8 (the size of a table entry): code(A,C,G,T) = 0,8,16,24 and the two rare
// classes (seed 0 / seed 1) = 32,40.  Three tables of (forward, reverse) 32-bit pairs:
//   xy  general: entry (out,in) at byte 8*code(out)+code(in) -- collision-free for all 6x6 combinations
//   xf  ACGT only: entry (out,in) at byte 4*code(out)+code(in): the sixteen entries fill bytes 0..127 exactly, one
//       bank each, so loads never conflict; four such offsets are formed at once as the bytes of W_out*4+W_in
//   x2  ACGT only, warm-up: entry (c1,c2) at byte 4*code(c1)+code(c2) advances the hash state by two bases
// With a rare class in play 4*out+in still is a multiple of 8 below 208 (an aligned load of a meaningless entry);
// the thread notices (bit 5 of some code it touched) and redoes its owners through xy.
constexpr int XYN   = 46;           // (8*40 + 40) / 8 + 1 table slots
constexpr int XFN   = 32;           // 26 reachable slots, rounded up
constexpr int ZC8   = 32;           // code of the class whose forward and reverse seeds are both 0
constexpr uint32_t RARE4 = 0x20202020u;   // bit 5 of every byte: set only in the codes of the two rare classes

constexpr uint64_t FLAG_AGG  = 1ull << 62;
constexpr uint64_t FLAG_INCL = 2ull << 62;
constexpr uint64_t VALMASK   = (1ull << 62) - 1;
constexpr uint32_t SPIN_LIMIT = 1u << 24;

constexpr uint32_t ERR_SPIN = 1u, ERR_CAP = 2u, ERR_ALIGN = 4u;

struct K1Args {
    const uint8_t  *bases;
    const uint64_t *seq_off;     // n_seqs + 1
    const uint32_t *tile_lb;     // n_tiles + 1: first i with seq_off[i] >= tile start
    uint32_t *ticket;            // zeroed before launch
    unsigned long long *cursor;  // zeroed before launch: bump allocator over min_out
    uint4    *tile_info;         // n_tiles: (hits, kept bases, first record lo, hi)
    uint4    *min_out;           // minimizer records (hash, start, end, seq), grouped by tile, tiles in any order
    uint64_t  min_cap;
    uint64_t  region_cap;        // > 0: CTA b appends to records [b*region_cap, (b+1)*region_cap) without any atomic;
                                 // 0: one global bump allocator (the rerun after a region overflowed)
    uint64_t *min_off;           // n_seqs + 1: minimizers of the tile before the sequence start (tile-local)
    uint64_t *hpc_off;           // n_seqs + 1 or null: kept bases of the tile before the sequence start
    uint32_t *hscr;              // gridDim.x * WIN words: per-CTA stash of selected hashes (stays in L2)
    uint32_t *err;
    uint64_t  n_seqs, n_bases;
    uint32_t  n_tiles, tile, halo;
    uint32_t  l, d, need, thr;
    uint32_t  one;               // 1: a multiplier ptxas cannot fold keeps `x + bit` on the multiply-add pipe (one_r below)
    uint32_t  vmask;             // packed compaction: bits of (canonical base ^ base) that make a base a rare class
    uint8_t   cls_lut[256];      // raw byte -> 8 * base class (classes 0..5)
    uint2     xy[XYN];           // byte offset 8*code(out)+code(in) -> (rol(h[out],l)^h[in], ror(rc[out],1)^rol(rc[in],l-1))
    uint2     xf[XFN];           // the same for ACGT x ACGT at byte offset 4*code(out)+code(in)
    uint2     x2[XFN];           // (c1,c2) at 4*code(c1)+code(c2) -> (rol(h[c1],1)^h[c2], ror(r[c1],1)^r[c2]), r[c]=rol(rc[c],l-1)
    // H = u64 flavour (src/lib.rs:30-32 `pub type H = u64`, 64-bit ntHash1 seeds, src/nthash_hpc.rs:30-49): the same
    // (out,in) table with 64-bit entries (x,y = forward low/high, z,w = reverse low/high), the bound, and where the
    // high halves of the selected hashes go (min_out[i].x holds the low half).  Used by k_minimizers<..., H64 = true>.
    uint4     xy64[XYN];
    uint64_t  thr64;
    uint32_t *min_hi;
    // H = u16 flavour (src/lib.rs:29), mode Regular: the 32-bit canonical hash is truncated before the test (`x as H`,
    // src/lib.rs:224) -- byte form of the tile only.  (Mode Hpc of that flavour needs nothing here: a 16-bit state
    // duplicated in both halves of the word rotates, XORs and compares like the 32-bit one; see make_plan.)
    uint32_t  trunc16;
};

struct Smem {
    // lut and xf sit on 256-byte boundaries of the shared window: their (8-bit) offsets are spliced into the address
    // with one PRMT instead of an extract + add (smem_splice below); k_minimizers checks the alignment once.
    alignas(256) uint8_t lut[256];           // raw byte -> 8 * base class
    alignas(256) uint2 xf[XFN];              // the 16 live entries cover the 32 banks once
    alignas(256) uint2 gm[32];               // packed compaction: keep nibble -> (gather multiplier, 2 * kept); 16 live entries
    alignas(128) uint2 x2[XFN];
    uint2    x4[256];                        // warm-up: (c0,c1,c2,c3) at c0 + 4*c1 + 16*c2 + 64*c3 advances the state by 4 bases
    uint2    x4r[256];                       // the same entries rotated by 4 more: two bytes of codes per rotate (packed warm-up)
    uint2    xrm[64];                        // packed warm-up: what the 0..3 oldest codes of its first byte added (see hash_owners_packed)
    alignas(16) uint32_t pk[PKW];            // packed tiles: 2-bit class of every kept base, 16 per word, same index space as code[]
    uint32_t rare;                           // a thread met a rare class while packing: the tile is redone byte-wise
    uint32_t tma_tile;                       // tile whose window a bulk copy is bringing into code[] (packed tiles leave
    uint32_t code_raw;                       // code[] free), ~0u = none; code_raw: code[] holds raw bases, not class codes
    alignas(8) unsigned long long mbar;      // completion barrier of the bulk copy
    alignas(16) uint8_t pre[16];             // stays ZC8: with l = 255 and owner space shifted 3 bases into the halo the
                                             // word-wise hash stage reads up to 4 bytes below code[0]
    uint8_t  code[XB + WIN + 128];           // 8*class of every kept base, index XB + (kept index in the window)
    unsigned long long hitw[2][NT + 1][MW];  // per pass, per thread: selected owners (bit i = owner CH*t + i)
    uint32_t hitpre[2][NT + 1];              // ... and how many hits precede that thread in the tile
    uint32_t keepw[NCHUNK + 1];              // keep mask per 32-base chunk
    uint32_t qoff[NCHUNK + 1];               // kept bases before the chunk
    uint32_t startw[NCHUNK];                 // raw-space bitmap: sequence starts
    uint32_t shortw[NCHUNK];                 // ... of sequences with len <= l
    uint32_t f1[FW];                         // owner-space bitmap: kept bases that start a sequence
    uint32_t f2[FW];                         // ... of sequences with len <= l
    uint32_t ctxpos[XB];                     // walk-back context: distance below W0
    uint16_t qmap[(XB + WIN) / 64 + 2];      // chunk holding kept base 64*m (coarse inverse of qoff)
    uint2    xy[XYN];
    uint4    xy64[XYN];                      // H = u64 flavour only
    uint16_t hl[HL];
    uint8_t  sel8[256 * 8];                  // sel8[8*b + r] = position of the r-th set bit of byte b (r < popc(b))
    uint32_t wsum[3][NT / 32];               // warp totals: kept-count scan, hit-count scan of pass 0 / pass 1
    uint32_t hk;
    uint32_t next[2];                        // tile ticket by parity: the next one is drawn while this one is processed
    uint32_t nlb[2][2];                      // ... with its tile_lb pair
    unsigned long long soc[SOC + 1];         // seq_off[lb-1 .. lb-1+SOC]: the emission looks sequences up here
    uint32_t n_dirty[2];                     // flag words set during this tile (cleared at the next loop top);
    uint16_t dirty[2][DIRTY_MAX];            // double-buffered by tile parity.  bit 15: f1/f2, else startw/shortw
    unsigned long long s0, rec0, rec_lim, cur;   // cur: records this CTA has appended to its region
};
static_assert(offsetof(Smem, lut) == 0 && offsetof(Smem, xf) == 256 && offsetof(Smem, gm) == 512,
              "k_minimizers derives the lut addresses from xf's");
static_assert(offsetof(Smem, code) % 16 == 0 && offsetof(Smem, pk) % 16 == 0, "vector stores into code[] and pk[]");

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ uint64_t ld_relaxed(const uint64_t *p)
{
#ifdef S2K_EMU
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_relaxed(uint64_t *p, uint64_t v)
{
#ifdef S2K_EMU
    __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ uint64_t warp_sum64(uint64_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Exclusive block scan of one uint32 per thread (NT threads). Returns exclusive prefix; total via out-param.
// One barrier: the caller hands in a wsum buffer whose previous readers are already behind some other barrier.
__device__ __forceinline__ uint32_t warp_incl_scan_m(uint32_t v, int lane, uint32_t one);
__device__ __forceinline__ void warp_totals(const uint32_t *wsum, int warp, int lane, uint32_t one, uint32_t &below, uint32_t &all);
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *wsum, uint32_t &total, uint32_t one)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan_m(v, lane, one);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t pre;
    warp_totals(wsum, warp, lane, one, pre, total);
    return pre + incl - v;
}
__device__ __forceinline__ uint32_t lowmask(uint32_t n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }
__device__ __forceinline__ unsigned long long lowmask64(uint32_t n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }
__device__ __forceinline__ int nth_set_bit(uint32_t w, int r)      // position of the r-th (0-based) set bit
{
    int pos = 0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const int c = __popc(w & ((1u << s) - 1u));
        if (r >= c) { r -= c; pos += s; w >>= s; }
    }
    return pos;
}
// ---- shared-memory table look-ups whose byte offset is spliced into the address.
// A table on a 256-byte boundary of the shared window has an address whose low byte is zero, so
// PRMT(word, base, 0x765k) = (base & ~0xff) | byte k of word IS the address of entry `byte k`: one instruction where
// extract + add took two (the compaction does one such look-up per raw base, the hash stage one per l-mer).
#ifdef S2K_EMU
typedef const uint8_t *smem_tab_t;
__device__ __forceinline__ smem_tab_t smem_tab(const void *p) { return reinterpret_cast<const uint8_t *>(p); }
__device__ __forceinline__ bool smem_tab_ok(smem_tab_t) { return true; }
__device__ __forceinline__ uint32_t tab_u8(smem_tab_t t, uint32_t w, int k) { return t[(w >> (8 * k)) & 0xffu]; }
__device__ __forceinline__ uint2 tab_u64(smem_tab_t t, uint32_t w, int k)
{
    return *reinterpret_cast<const uint2 *>(t + ((w >> (8 * k)) & 0xffu));
}
#else
typedef uint32_t smem_tab_t;
__device__ __forceinline__ smem_tab_t smem_tab(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool smem_tab_ok(smem_tab_t t) { return (t & 0xffu) == 0u; }
__device__ __forceinline__ uint32_t tab_u8(smem_tab_t t, uint32_t w, int k)      // k: a constant after unrolling
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(__byte_perm(w, t, 0x7650u + (uint32_t)k)));
    return v;
}
__device__ __forceinline__ uint2 tab_u64(smem_tab_t t, uint32_t w, int k)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(__byte_perm(w, t, 0x7650u + (uint32_t)k)));
    return v;
}
#endif
template <bool W31> __device__ __forceinline__ uint32_t rol1(uint32_t x)
{
    if (W31) return ((x << 1) | (x >> 30)) & 0x7fffffffu;
    return __funnelshift_l(x, x, 1);
}
template <bool W31> __device__ __forceinline__ uint32_t ror1(uint32_t x)
{
    if (W31) return (x >> 1) | ((x & 1u) << 30);
    return __funnelshift_r(x, x, 1);
}
template <bool W31> __device__ __forceinline__ uint32_t rol2(uint32_t x)
{
    if (W31) return ((x << 2) | (x >> 29)) & 0x7fffffffu;
    return __funnelshift_l(x, x, 2);
}
template <bool W31> __device__ __forceinline__ uint32_t ror2(uint32_t x)
{
    if (W31) return (x >> 2) | ((x & 3u) << 29);
    return __funnelshift_r(x, x, 2);
}
// a * b + c kept on the multiply-add pipe (IMAD), away from the integer ALU
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
#ifdef S2K_EMU
    return a * b + c;
#else
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#endif
}
// warp_incl_scan with its adds on the multiply-add pipe (one = 1 in a register ptxas cannot fold, see k_minimizers)
__device__ __forceinline__ uint32_t warp_incl_scan_m(uint32_t v, int lane, uint32_t one)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = mad_u32(one, t, v);
    }
    return v;
}
// The NT / 32 warp totals of a block scan, turned into (sum of the warps below, sum of all) by a scan over lanes 0..7
// with its adds on the multiply-add pipe -- one shared load, five shuffles and a handful of ALU instructions where the
// plain loop spends eight loads and sixteen.  S2K_WARP_TOTALS_LOOP restores the loop.
__device__ __forceinline__ void warp_totals(const uint32_t *wsum, int warp, int lane, uint32_t one, uint32_t &below, uint32_t &all)
{
#if defined(S2K_WARP_TOTALS_LOOP) || defined(S2K_EMU)
    uint32_t pre = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) {
        const uint32_t s = wsum[i];
        if (i < warp) pre += s;
        tot += s;
    }
    (void)lane; (void)one;
    below = pre; all = tot;
#else
    static_assert(NT / 32 == 8, "warp_totals scans eight totals");
    const uint32_t x = wsum[lane & 7];
    uint32_t v = x;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if ((lane & 7) >= o) v = mad_u32(one, t, v);
    }
    all = __shfl_sync(0xffffffffu, v, 7);
    below = __shfl_sync(0xffffffffu, v - x, warp);
#endif
}
template <bool W31> __device__ __forceinline__ uint32_t rol8(uint32_t x)
{
    if (W31) return ((x << 8) | (x >> 23)) & 0x7fffffffu;
    return __funnelshift_l(x, x, 8);
}
template <bool W31> __device__ __forceinline__ uint32_t ror8(uint32_t x)
{
    if (W31) return (x >> 8) | ((x & 255u) << 23);
    return __funnelshift_r(x, x, 8);
}
template <bool W31> __device__ __forceinline__ uint32_t rol4(uint32_t x)
{
    if (W31) return ((x << 4) | (x >> 27)) & 0x7fffffffu;
    return __funnelshift_l(x, x, 4);
}
template <bool W31> __device__ __forceinline__ uint32_t ror4(uint32_t x)
{
    if (W31) return (x >> 4) | ((x & 15u) << 27);
    return __funnelshift_r(x, x, 4);
}
// Sequence-start flags are sparse: remember which bitmap words were touched so that the next tile clears those
// instead of zeroing four bitmaps.  More than DIRTY_MAX touched words -> the next tile zeroes everything.
__device__ __forceinline__ void flag_raw(Smem &S, int par, uint32_t x, bool is_short)
{
    atomicOr(&S.startw[x >> 5], 1u << (x & 31));
    if (is_short) atomicOr(&S.shortw[x >> 5], 1u << (x & 31));
    const uint32_t k = atomicAdd(&S.n_dirty[par], 1u);
    if (k < DIRTY_MAX) S.dirty[par][k] = (uint16_t)(x >> 5);
}
__device__ __forceinline__ void flag_owner(Smem &S, int par, int oo, bool is_short)
{
    atomicOr(&S.f1[oo >> 5], 1u << (oo & 31));
    if (is_short) atomicOr(&S.f2[oo >> 5], 1u << (oo & 31));
    const uint32_t k = atomicAdd(&S.n_dirty[par], 1u);
    if (k < DIRTY_MAX) S.dirty[par][k] = (uint16_t)(0x8000u | (uint32_t)(oo >> 5));
}
__device__ __forceinline__ uint2 xy_at(const Smem &S, uint32_t out8, uint32_t in8)
{
    return *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(S.xy) + (out8 << 3) + in8);
}

// Chunk (32 raw bases) that holds the kept base with window index q >= 0.
__device__ __forceinline__ int chunk_of(const Smem &S, int q)
{
    int c = S.qmap[q >> 6];
    while ((int)S.qoff[c + 1] <= q) ++c;
    return c;
}
// Original-space position (global index into `bases`) of kept base q that lies in chunk c.
template <bool LUT>
__device__ __forceinline__ int64_t pos_in_chunk(const Smem &S, int64_t W0, int c, int q)
{
    if (!LUT) return W0 + 32 * c + nth_set_bit(S.keepw[c], q - (int)S.qoff[c]);   // HPC off: measured faster (registers)
    // select in two halving steps and one table look-up (the five-step form costs about twice the instructions)
    uint32_t w = S.keepw[c];
    int r = q - (int)S.qoff[c], pos = 0;
    int n = __popc(w & 0xffffu);
    if (r >= n) { r -= n; pos = 16; w >>= 16; }
    n = __popc(w & 0xffu);
    if (r >= n) { r -= n; pos += 8; w >>= 8; }
    return W0 + 32 * c + pos + (int)S.sel8[((w & 0xffu) << 3) + (uint32_t)(r & 7)];
}

// ------------------------------------------------------------------------------------------------ hash stage
// One thread, CH consecutive owners.  cb[i] = class code of the last base of owner i's l-mer (cb is 4-aligned);
// cb[i-l] = the base leaving the window.  Selected owners: bit i of mask, hash to hs[i].
//
// hash_owners_bytes: any classes.  Per owner two byte loads, one IMAD (8*out+in), one 8-byte load from xy.
template <bool W31, bool T16 = false>
__device__ __forceinline__ void hash_owners_bytes(const Smem &S, const uint8_t *cb, int l, uint32_t thr, uint32_t *hs,
                                                  unsigned long long (&mask)[MW])
{
    uint32_t fh = 0, rh = 0;
    for (int j = 1 - l; j < 0; ++j) {                  // warm-up: first l-1 bases of owner 0's l-mer
        const uint2 tt = xy_at(S, ZC8, cb[j]);
        fh = rol1<W31>(fh) ^ tt.x;
        rh = ror1<W31>(rh) ^ tt.y;
    }
    const uint8_t *co = cb - l;
#pragma unroll 4
    for (int i = 0; i < CH; ++i) {
        const uint2 tt = xy_at(S, i > 0 ? (uint32_t)co[i] : (uint32_t)ZC8, cb[i]);
        fh = rol1<W31>(fh) ^ tt.x;
        rh = ror1<W31>(rh) ^ tt.y;
        const uint32_t hv = T16 ? (min(fh, rh) & 0xffffu) : min(fh, rh);
        if (hv <= thr) { mask[i >> 6] |= 1ull << (i & 63); hs[i] = hv; }
    }
}
// hash_owners_bytes64: the H = u64 flavour (64-bit state per strand, `hash <= bound` on 64 bits).  Low halves of the
// selected hashes go to hs[i], high halves to hs[WIN + i].  The byte form of the tile is the only one it has.
__device__ __forceinline__ uint64_t rol1_64(uint64_t x) { return (x << 1) | (x >> 63); }
__device__ __forceinline__ uint64_t ror1_64(uint64_t x) { return (x >> 1) | (x << 63); }
__device__ __forceinline__ void hash_owners_bytes64(const Smem &S, const uint8_t *cb, int l, uint64_t thr, uint32_t *hs,
                                                    unsigned long long (&mask)[MW])
{
    uint64_t fh = 0, rh = 0;
    for (int j = 1 - l; j < 0; ++j) {                  // warm-up: first l-1 bases of owner 0's l-mer
        const uint4 tt = S.xy64[ZC8 + (cb[j] >> 3)];
        fh = rol1_64(fh) ^ ((uint64_t)tt.y << 32 | tt.x);
        rh = ror1_64(rh) ^ ((uint64_t)tt.w << 32 | tt.z);
    }
    const uint8_t *co = cb - l;
#pragma unroll 4
    for (int i = 0; i < CH; ++i) {
        const uint4 tt = S.xy64[(i > 0 ? (uint32_t)co[i] : (uint32_t)ZC8) + (cb[i] >> 3)];
        fh = rol1_64(fh) ^ ((uint64_t)tt.y << 32 | tt.x);
        rh = ror1_64(rh) ^ ((uint64_t)tt.w << 32 | tt.z);
        const uint64_t hv = fh < rh ? fh : rh;
        if (hv <= thr) { mask[i >> 6] |= 1ull << (i & 63); hs[i] = (uint32_t)hv; hs[WIN + i] = (uint32_t)(hv >> 32); }
    }
}
// hash_owners_words: the common case, every class among A,C,G,T.  Class bytes are read as words; the four table
// offsets of a group of owners are the bytes of W_out*4 + W_in (one IMAD), split with PRMT; the warm-up advances two
// bases per table load.  Returns the OR of every code word it looked at: a bit of RARE4 set means the result is
// meaningless (loads stayed aligned and inside the tables) and the caller falls back to hash_owners_bytes.
// DENSE: at the usual densities some lane of the warp selects an owner in almost every group of four (2 % of the
// owners at d = 0.01: 1 - 0.98^128 = 92 %), so a "rare" branch around the bookkeeping would nearly always run; the
// three predicated instructions per owner are cheaper.  Sparse selections keep the test per group.
// Multiplier that gathers the four 2-bit classes of a code word (bytes 8*c) into one byte: the high half of
// w * X4MUL holds c0 | c1 << 2 | c2 << 4 | c3 << 6 in its low byte (no two partial products share a bit and the low
// half cannot carry: 3 * (2^30 + 2^28 + 2^26 + 2^22 + 2^20 + 2^14) < 2^32).
constexpr uint32_t X4MUL = (1u << 29) | (1u << 23) | (1u << 17) | (1u << 11);
template <bool W31, bool DENSE>
__device__ __forceinline__ uint32_t hash_owners_words(const Smem &S, smem_tab_t xft, const uint8_t *cb, int l, uint32_t thr,
                                                      uint32_t *hs, unsigned long long (&mask)[MW])
{
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(cb);
    const uint8_t *x2 = reinterpret_cast<const uint8_t *>(S.x2);
    uint32_t fh = 0, rh = 0, seen = 0;
    {   // warm-up over cb[1-l .. -1] (it ends on a word boundary): an odd first base alone (general table), one pair
        // (two-base table), then whole code words through the four-base table -- 7 steps of 9 instructions for l = 31
        int j = 1 - l;
        if ((l - 1) & 1) {
            const uint32_t c = cb[j];
            const uint2 tt = xy_at(S, ZC8, c);
            fh = tt.x; rh = tt.y; seen |= c;
            ++j;
        }
        if ((l - 1) & 2) {
            const uint32_t c1 = cb[j], c2 = cb[j + 1];
            const uint2 tt = *reinterpret_cast<const uint2 *>(x2 + 4u * c1 + c2);
            fh = rol2<W31>(fh) ^ tt.x; rh = ror2<W31>(rh) ^ tt.y; seen |= c1 | c2;
        }
        const int nw = (l - 1) >> 2;
#pragma unroll 8
        for (int wi = -nw; wi < 0; ++wi) {
            const uint32_t w = cw[wi];
            seen |= w;
            const uint2 tt = S.x4[__umulhi(w, X4MUL) & 0xffu];
            fh = rol4<W31>(fh) ^ tt.x;
            rh = ror4<W31>(rh) ^ tt.y;
        }
    }
    const int oq = -((l + 3) >> 2);                        // word of cb[-l] relative to cw, and its byte shift
    const uint32_t ob = (uint32_t)((-l) & 3) * 8u;
    const uint32_t *ow = cw + oq;
    uint32_t o_prev = ow[0];
#pragma unroll
    for (int g = 0; g < CH / 4; ++g) {
        const uint32_t w_in = cw[g], o_next = ow[g + 1];
        const uint32_t w_out = __funnelshift_r(o_prev, o_next, ob);
        o_prev = o_next;
        seen |= w_in | w_out;
        const uint32_t i4 = w_out * 4u + w_in;
        uint32_t hv[4];
        {
            uint2 tt;
            if (g == 0) tt = xy_at(S, ZC8, w_in & 0xffu);           // owner 0: nothing leaves yet
            else tt = tab_u64(xft, i4, 0);
            fh = rol1<W31>(fh) ^ tt.x; rh = ror1<W31>(rh) ^ tt.y; hv[0] = min(fh, rh);
            tt = tab_u64(xft, i4, 1);
            fh = rol1<W31>(fh) ^ tt.x; rh = ror1<W31>(rh) ^ tt.y; hv[1] = min(fh, rh);
            tt = tab_u64(xft, i4, 2);
            fh = rol1<W31>(fh) ^ tt.x; rh = ror1<W31>(rh) ^ tt.y; hv[2] = min(fh, rh);
            tt = tab_u64(xft, i4, 3);
            fh = rol1<W31>(fh) ^ tt.x; rh = ror1<W31>(rh) ^ tt.y; hv[3] = min(fh, rh);
        }
        if (DENSE || min(min(hv[0], hv[1]), min(hv[2], hv[3])) <= thr) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (hv[k] <= thr) { mask[(4 * g + k) >> 6] |= 1ull << ((4 * g + k) & 63); hs[4 * g + k] = hv[k]; }
        }
    }
    return seen;
}

// hash_owners_packed: the same CH owners from the PACKED code array (2 bits per kept base; tiles without rare classes
// and without a walk-back).  E = index (code[] space) of the last base of owner 0's l-mer.  Seven consecutive words of
// the array, funnel-shifted once, hold codes E-32 .. E+63 in registers: two words of warm-up context and the entering
// codes of all owners; five more give the leaving codes.  Per 16 owners the (out,in) pairs are interleaved into
// nibbles (X: even owners, Y: odd owners; two LOP3 + two shifts) and spread into four words of table offsets
// 8*(4*out+in), one per byte -- 12 instructions where the byte path spends 20 (two loads, a funnel shift, an IMAD and a
// rare-class OR per group of four).  Warm-up: the l-1 codes before owner 0 are rounded UP to whole bytes, each byte is
// one look-up in the four-base table (a packed byte IS its index), and what the 0..3 surplus codes of the oldest byte
// added is taken out again with one look-up in xrm (the hash is XOR-linear in the bases).
template <bool W31, bool DENSE>
__device__ __forceinline__ void hash_owners_packed(const Smem &S, smem_tab_t xft, int E, int l, uint32_t thr, uint32_t *hs,
                                                   unsigned long long (&mask)[MW], uint32_t one)
{
    constexpr int NG = (CH + 15) / 16;
    uint32_t mw[2 * MW];                                   // the hit mask as 32-bit words
#pragma unroll
    for (int x = 0; x < 2 * MW; ++x) mw[x] = 0u;
    uint32_t R[NG + 2];                                    // R[0], R[1]: codes E-32 .. E-1; R[2 + g]: codes E+16g ..
    {
        const int e0 = E - 32;
        const uint32_t *pw = S.pk + (e0 >> 4);
        const uint32_t sh = 2u * (uint32_t)(e0 & 15);
        uint32_t a = pw[0];
#pragma unroll
        for (int i = 0; i < NG + 2; ++i) {
            const uint32_t b = pw[i + 1];
            R[i] = __funnelshift_r(a, b, sh);
            a = b;
        }
    }
    uint32_t fh = 0, rh = 0;
    {   // warm-up: bytes nb-1 .. 0 counted back from code E-1 (static positions), oldest first
        const int n = l - 1, nb = (n + 3) >> 2, sur = 4 * nb - n;
#pragma unroll
        for (int k = 7; k >= 1; k -= 2) {                  // two bytes per rotate: the older one through the pre-rotated table
            const uint32_t b1 = ((k < 4 ? R[1] : R[0]) >> (8 * (3 - (k & 3)))) & 0xffu;
            const uint32_t b0 = ((k - 1 < 4 ? R[1] : R[0]) >> (8 * (3 - ((k - 1) & 3)))) & 0xffu;
            if (k < nb) {
                const uint2 t1 = S.x4r[b1], t0 = S.x4[b0];
                fh = rol8<W31>(fh) ^ t1.x ^ t0.x;
                rh = ror8<W31>(rh) ^ t1.y ^ t0.y;
            } else if (k - 1 < nb) {                       // nb odd: its oldest byte stands alone
                const uint2 t0 = S.x4[b0];
                fh = t0.x; rh = t0.y;
            }
        }
        if (sur) {
            const uint32_t ob = ((nb > 4 ? R[0] : R[1]) >> (8 * ((8 - nb) & 3))) & ((1u << (2 * sur)) - 1u);
            const uint2 tt = S.xrm[ob];
            fh ^= tt.x; rh ^= tt.y;
        }
    }
    uint32_t O[NG + 1];
    {
        const int o0 = E - l;
        const uint32_t *pw = S.pk + (o0 >> 4);
        const uint32_t sh = 2u * (uint32_t)(o0 & 15);
        uint32_t a = pw[0];
#pragma unroll
        for (int i = 0; i < NG; ++i) {
            const uint32_t b = pw[i + 1];
            O[i] = __funnelshift_r(a, b, sh);
            a = b;
        }
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const uint32_t in = R[2 + g], out = O[g];
        const uint32_t X = (in & 0x33333333u) | ((out << 2) & 0xccccccccu);        // nibble n: owner 2n   (4*out + in)
        const uint32_t Y = ((in >> 2) & 0x33333333u) | (out & 0xccccccccu);        // nibble n: owner 2n+1
        const uint32_t Wt[4] = {(X << 3) & 0x78787878u, (Y << 3) & 0x78787878u,    // byte r: owner 4r + t
                                (X >> 1) & 0x78787878u, (Y >> 1) & 0x78787878u};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (16 * g + 4 * r >= CH) break;
            uint32_t hv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                uint2 tt;
                if (g == 0 && r == 0 && t == 0) tt = xy_at(S, ZC8, (in & 3u) << 3);   // owner 0: nothing leaves yet
                else tt = tab_u64(xft, Wt[t], r);
                fh = rol1<W31>(fh) ^ tt.x; rh = ror1<W31>(rh) ^ tt.y; hv[t] = min(fh, rh);
            }
            if (DENSE || min(min(hv[0], hv[1]), min(hv[2], hv[3])) <= thr) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int i = 16 * g + 4 * r + t;
                    if (hv[t] <= thr) { mw[i >> 5] = mad_u32(one, 1u << (i & 31), mw[i >> 5]); hs[i] = hv[t]; }   // "|=" on the multiply-add pipe
                }
            }
        }
    }
#pragma unroll
    for (int x = 0; x < MW; ++x) mask[x] |= (unsigned long long)mw[2 * x] | ((unsigned long long)mw[2 * x + 1] << 32);
}

// Rotations by a run-time amount within w = 32 or 31 bits (table construction only).
template <bool W31> __device__ __forceinline__ uint32_t rolv(uint32_t x, uint32_t s)
{
    const uint32_t w = W31 ? 31u : 32u;
    s %= w;
    if (s == 0) return x;
    return ((x << s) | (x >> (w - s))) & (W31 ? 0x7fffffffu : 0xffffffffu);
}
template <bool W31> __device__ __forceinline__ uint32_t rorv(uint32_t x, uint32_t s)
{
    const uint32_t w = W31 ? 31u : 32u;
    s %= w;
    return rolv<W31>(x, w - s);
}

// ------------------------------------------------------------------------------------------------ tile bounds
__global__ void k_tile_bounds(const uint64_t *__restrict__ seq_off, uint64_t n_seqs, uint64_t n_bases, uint32_t tile,
                              uint32_t n_tiles, uint32_t *__restrict__ tile_lb)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    uint64_t pos = (uint64_t)t * tile;
    if (pos > n_bases) pos = n_bases;
    uint64_t lo = 0, hi = n_seqs + 1;       // first i in [0, n_seqs] with seq_off[i] >= pos
    while (lo < hi) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (seq_off[mid] < pos) lo = mid + 1; else hi = mid;
    }
    tile_lb[t] = (uint32_t)lo;
}

// ------------------------------------------------------------------------------------------------ minimizers
// One tile = `tile` raw bases plus a left halo.  Per tile:
//   S2  sequence starts falling into the window -> raw-space bitmaps
//   S3  64 raw bases per thread straight from global into registers; keep mask (byte != previous byte, or a
//       sequence start); block scan of kept counts
//   S4  branch-free compaction of the kept bases' classes into S.code (HPC order); start flags -> owner space;
//       if the halo holds fewer than l-1(+1) kept bases of the current sequence, warp 0 walks further back
//   S5  every kept base of the tile "owns" one l-mer (the one it completes).  60 owners per thread: l-1 warm-up
//       steps, then one rolling step per owner: 2 byte loads, one table load, 2 rotates, 2 three-input XORs,
//       min, compare.  Selected hashes go to a per-CTA scratch that lives in L2.
//   S6  block scan of hit counts, decoupled look-back across tiles (one 64-bit word: kept count | hit count)
//   S7  hits are listed in order in shared memory and handed out one per thread: positions by rank/select on
//       the keep masks, sequence index by binary search of seq_off, one 16-byte record store per minimizer
//   S8  per-sequence offsets (minimizers, kept bases) for every sequence that starts in the tile
#ifndef S2K_MINB
#define S2K_MINB 4                  // CTAs per SM the minimizer kernel is compiled for (register cap)
#endif
// -DS2K_PHASE_CLOCKS (tools/phase_clocks.py): thread 0 of every CTA sums the cycles between phase marks into g_phase.
#if defined(S2K_PHASE_CLOCKS) && !defined(S2K_EMU)
__device__ unsigned long long g_phase[16];
#define PHASE_DECL long long ph_t = clock64(); unsigned long long ph_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define PHASE(i) do { if (threadIdx.x == 0) { const long long n_ = clock64(); ph_acc[i] += (unsigned long long)(n_ - ph_t); ph_t = n_; } } while (0)
#define PHASE_FLUSH do { if (threadIdx.x == 0) for (int i_ = 0; i_ < 10; ++i_) atomicAdd(&g_phase[i_], ph_acc[i_]); } while (0)
#else
#define PHASE_DECL
#define PHASE(i)
#define PHASE_FLUSH
#endif
#if !defined(S2K_EMU) && !defined(S2K_NO_TMA)
#define S2K_TMA 1
#endif
// ---- 1-D bulk copy (TMA) of the next tile's window into shared memory, completion on an mbarrier.
// One thread arms the barrier with the byte count and issues cp.async.bulk; the copy engine fetches the 16 KB while the
// CTA hashes the current tile; at the next tile every thread waits on the barrier's phase and reads its pieces with
// four conflict-free 16-byte shared loads instead of four global loads.
#ifdef S2K_TMA
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // earlier generic-proxy accesses of dst come first
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tS2K_MBAR_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra S2K_MBAR_DONE;\n\t"
                 "bra S2K_MBAR_WAIT;\n\tS2K_MBAR_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif
// S3a of k_minimizers: this thread's four pieces of 16 raw bases (piece 32*j + lane of its warp's 2048 bases).
__device__ __forceinline__ void load_pieces(const K1Args &A, int64_t W0, int xw, bool full, uint32_t (&w)[16])
{
    if (full) {
        const uint4 *src = reinterpret_cast<const uint4 *>(A.bases + W0 + xw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 x = __ldg(src + 32 * j);
            w[4 * j] = x.x; w[4 * j + 1] = x.y; w[4 * j + 2] = x.z; w[4 * j + 3] = x.w;
        }
    } else {
#pragma unroll
        for (int v = 0; v < 16; ++v) {
            uint32_t x = 0;
            for (int b = 0; b < 4; ++b) {
                const int64_t gg = W0 + xw + 512 * (v >> 2) + 4 * (v & 3) + b;
                if (gg >= 0 && gg < (int64_t)A.n_bases) x |= (uint32_t)A.bases[gg] << (8 * b);
            }
            w[v] = x;
        }
    }
}
// S4, byte form: class bytes (8 * class) of the kept bases to S.code, predicated byte stores in HPC order.
template <bool HPC>
__device__ __forceinline__ void compact_bytes(Smem &S, smem_tab_t xft, const uint32_t (&w)[16], const uint32_t (&k16)[4],
                                              const uint32_t (&qj)[4])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (!HPC && k16[j] == 0xffffu && ((XB + qj[j]) & 15u) == 0) {
            // every base kept (no HPC): the 16 classes of the piece as one 16-byte store
            uint32_t o4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t x = w[4 * j + i];
                o4[i] = (uint32_t)S.lut[x & 0xffu] | ((uint32_t)S.lut[(x >> 8) & 0xffu] << 8) |
                        ((uint32_t)S.lut[(x >> 16) & 0xffu] << 16) | ((uint32_t)S.lut[x >> 24] << 24);
            }
            *reinterpret_cast<uint4 *>(S.code + XB + qj[j]) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        } else {
            uint8_t *cp = S.code + XB + qj[j];
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                if ((k16[j] >> b) & 1u) { *cp = (uint8_t)tab_u8(xft - 256, w[4 * j + (b >> 2)], b & 3); ++cp; }
            }
        }
    }
}
// S4, packed form, all arithmetic (no look-up per base).  Per word of four raw bases:
//   class    V = x & 0x06060606: bits 1-2 of an ASCII base are its class in the order A C T G = 0 1 2 3 (the class
//            numbering of every table; it also is the 2-bit transport format)
//   validity 2 * class indexes an 8-entry byte table held in two registers (PRMT with a data-dependent selector is a
//            4-way table look-up): the canonical base of that class, compared with the base itself.  The
//            differences are OR-ed up and tested once per thread against vmask (0xff per byte in the scalar profile:
//            exactly A C G T; 0x0f in the nibble profile: the low nibble decides, src/nthash_avx512_32.rs:178-193).
//            A set bit = some base of a rare class: the tile is redone in byte form.
//   gather   V * M[keep nibble] (low 32 bits) holds the classes of the kept bases of the word in its TOP 2*kept bits:
//            the multiplier has one term per kept base j, 2^(31 - 2*kept + 2*rank(j) - 8j), which moves class j to bit
//            32 - 2*kept + 2*rank(j); the other partial products fall off the top or below the field, no two on the
//            same bits (so no carry).  A funnel shift pushes the field into the accumulator from below, last word
//            first.
// Eight ALU instructions and one multiply per word; two shared-memory atomics per piece.  (A 256-entry table fetched
// once per two words measured slower: 11.62 against 11.54 ms per 10 Gbp.)
__device__ __forceinline__ uint32_t compact_packed(Smem &S, const uint32_t (&w)[16], const uint32_t (&k16)[4],
                                                   const uint32_t (&qj)[4], uint32_t vmask)
{
    const uint8_t *gm = reinterpret_cast<const uint8_t *>(S.gm);
    uint32_t bad = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 3; i >= 0; --i) {                     // last word first: the first kept base ends up in the low bits
            const uint32_t x = w[4 * j + i];
            const uint32_t v = x & 0x06060606u;                                    // 2 * class of the four bases
            const uint32_t sel = v + (v >> 12);                                    // nibbles: v0 v2 v1 v3
            uint32_t canon;                                                        // A C T G at indices 0 2 4 6
#ifdef S2K_EMU
            canon = __byte_perm(0x00430041u, 0x00470054u, sel);
#else
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(canon) : "r"(0x00430041u), "r"(0x00470054u), "r"(sel));
#endif
            bad |= canon ^ __byte_perm(x, 0u, 0x3120u);
            const uint2 e = *reinterpret_cast<const uint2 *>(gm + (((k16[j] << 3) >> (4 * i)) & 0x78u));
            acc = __funnelshift_l(v * e.x, acc, e.y);
        }
        const uint32_t p0 = (uint32_t)XB + qj[j], sh = 2u * (p0 & 15u);
        const uint32_t lo = acc << sh, hi = __funnelshift_l(acc, 0u, sh);
        if (lo) atomicOr(&S.pk[p0 >> 4], lo);
        if (hi) atomicOr(&S.pk[(p0 >> 4) + 1], hi);
    }
    return bad & vmask;
}

// A tile in byte form after bulk copies have used code[] as their landing buffer: raw bases are not class codes (the
// hash stage forms table offsets from whatever lies around the kept bases), so the array is reset first.  Rare.
__device__ __forceinline__ void reset_codes(Smem &S)
{
    const uint32_t z = ZC8 * 0x01010101u;
    for (int i = threadIdx.x; i < (int)(sizeof(S.code) / 16); i += NT) reinterpret_cast<uint4 *>(S.code)[i] = make_uint4(z, z, z, z);
    __syncthreads();
    if (threadIdx.x == 0) S.code_raw = 0u;
}

template <bool HPC, bool W31, bool DENSE, bool H64 = false>
__global__ void __launch_bounds__(NT, S2K_MINB) k_minimizers(const __grid_constant__ K1Args A)
{
    S2K_DYN_SMEM(smem_raw);
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = (int)A.l, d = (int)A.d;
    uint32_t *const hs = A.hscr + (size_t)blockIdx.x * (H64 ? 2 * WIN : WIN);
    if (H64 && tid < XYN) S.xy64[tid] = A.xy64[tid];

    for (int i = tid; i < 256; i += NT) S.lut[i] = A.cls_lut[i];
    for (int i = tid; HPC && i < 256 * 8; i += NT) {        // the table-assisted select serves the HPC variants only
        uint32_t b = (uint32_t)i >> 3;
        for (int r = i & 7; r > 0 && b; --r) b &= b - 1u;      // drop the r lowest set bits
        S.sel8[i] = (uint8_t)(b ? __ffs((int)b) - 1 : 0);
    }
    if (tid < XYN) S.xy[tid] = A.xy[tid];
    if (tid < XFN) { S.xf[tid] = A.xf[tid]; S.x2[tid] = A.x2[tid]; }
    for (int i = tid; i < 256; i += NT) {                   // four warm-up bases at once = two steps of the pair table
        const uint2 a = A.x2[4 * (i & 3) + ((i >> 2) & 3)], b = A.x2[4 * ((i >> 4) & 3) + (i >> 6)];
        const uint2 e4 = make_uint2(rol2<W31>(a.x) ^ b.x, ror2<W31>(a.y) ^ b.y);
        S.x4[i] = e4;
        S.x4r[i] = make_uint2(rol4<W31>(e4.x), ror4<W31>(e4.y));
    }
    if (tid < 16) {                                        // gm: see compact_packed
        const int wd = 2 * __popc((uint32_t)tid);
        uint32_t m = 0, r = 0;
        for (int j = 0; j < 4; ++j)
            if ((tid >> j) & 1) { m |= 1u << (31 - wd + 2 * (int)r - 8 * j); ++r; }
        S.gm[tid] = make_uint2(m, (uint32_t)wd);
    }
    if (tid < 64) {                                        // xrm: see hash_owners_packed (A.xy[32 + c] = (h[c], rol(rc[c], l-1)))
        const int n = (int)A.l - 1, nb = (n + 3) >> 2, sur = 4 * nb - n;
        uint2 e = make_uint2(0u, 0u);
        for (int i = 0; i < sur; ++i) {
            const uint2 hc = A.xy[32 + ((tid >> (2 * i)) & 3)];
            e.x ^= rolv<W31>(hc.x, (uint32_t)(4 * nb - 1 - i));
            e.y ^= rorv<W31>(hc.y, (uint32_t)(4 * nb - 1 - i));
        }
        S.xrm[tid] = e;
    }
    for (int i = tid; i < PKW; i += NT) S.pk[i] = 0u;
    if (tid == 0) {
        S.rare = 0u; S.tma_tile = ~0u; S.code_raw = 0u;
#ifdef S2K_TMA
        mbar_init(&S.mbar, 1u);
#endif
    }
    uint32_t tma_phase = 0;                                // parity of the next bulk copy to complete
    // The integer ALU (LOP3, SHF, PRMT, ISETP, VIMNMX: half the issue rate) is the busiest unit of this kernel and the
    // multiply-add pipe is nearly idle, so additions that ptxas would place on the ALU are written as one_r * x + y:
    // one_r is 1 in a register ptxas can neither fold nor re-load from the constant bank (full warps: 32 >> 5).
    const uint32_t one_r = A.one & ((uint32_t)__popc(__activemask()) >> 5);
    const smem_tab_t xft = smem_tab(S.xf);                 // S.lut lies 256 bytes below: one register
    if ((tid == 0 && !smem_tab_ok(xft)) || one_r != 1u) atomicOr(A.err, ERR_ALIGN);   // (a partial warp would make one_r 0)
    for (int i = tid; i < (int)sizeof(S.code); i += NT) S.code[i] = ZC8;
    if (tid < 16) S.pre[tid] = ZC8;
    for (int i = tid; i < NCHUNK; i += NT) { S.startw[i] = 0; S.shortw[i] = 0; }
    for (int i = tid; i < FW; i += NT) { S.f1[i] = 0; S.f2[i] = 0; }
    if (tid == 0) { S.n_dirty[0] = 0; S.n_dirty[1] = 0; }
    if (tid == 0) {                                        // first ticket; later ones are drawn a tile ahead
        S.cur = 0ull;
        const uint32_t t0 = atomicAdd(A.ticket, 1u);
        S.next[0] = t0;
        if (t0 < A.n_tiles) {
            S.nlb[0][0] = A.tile_lb[t0];
            S.nlb[0][1] = t0 + 1 == A.n_tiles ? (uint32_t)(A.n_seqs + 1) : A.tile_lb[t0 + 1];
        }
    }
    int par = 0;                                           // parity of the tile being processed
    PHASE_DECL;

    for (;;) {
        __syncthreads();                                   // the previous tile is done and has wiped its flag words
        const uint32_t t = S.next[par];
        if (t >= A.n_tiles) break;
        uint32_t t_next = 0;
        if (tid == 0) {
            S.n_dirty[par ^ 1] = 0;                        // the previous tile's list has been consumed
            t_next = atomicAdd(A.ticket, 1u);              // in flight while this tile is staged; parked in S.next below
        }
        PHASE(0);

        const int64_t T0 = (int64_t)((uint64_t)t * A.tile);
        const int64_t T1 = min(T0 + (int64_t)A.tile, (int64_t)A.n_bases);
        const int64_t W0 = T0 - (int64_t)A.halo;
        const bool last_tile = (uint64_t)T1 == A.n_bases;
        const uint32_t lb = S.nlb[par][0], ub = S.nlb[par][1];

        // ---- S3a: this thread's 64 raw bases, global -> registers (issued first: the latency overlaps S2).
        // A warp covers 2048 consecutive bases of the window in four rounds of 32 PIECES of 16 bases: in round j lane L
        // holds piece 32*j + L, so a round's loads are one contiguous 512-byte row, and the kept bases of neighbouring
        // lanes land about 12 bytes = 3 words apart in S.code: the compaction's byte stores of a round fall into
        // different banks (64 consecutive bases per thread put the lanes 12 words apart: 4-way conflicts).
        const int xw = 2048 * warp + 16 * lane;            // window offset of this thread's piece of round 0
        const bool full = W0 >= 0 && W0 + (int64_t)WIN <= (int64_t)A.n_bases;      // uniform: the whole window is readable
        uint32_t w[16];
#ifdef S2K_TMA
        if (S.tma_tile == t) {                             // uniform: the bulk copy issued during the previous tile
            mbar_wait(&S.mbar, tma_phase & 1u);
            ++tma_phase;
            const uint4 *src = reinterpret_cast<const uint4 *>(S.code + xw);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 x = src[32 * j];
                w[4 * j] = x.x; w[4 * j + 1] = x.y; w[4 * j + 2] = x.z; w[4 * j + 3] = x.w;
            }
        } else
#endif
        load_pieces(A, W0, xw, full, w);

        // ---- S2: sequence starts inside the tile (and the start of the sequence containing T0, if in the window)
        for (uint32_t i = lb + tid; i < ub; i += NT) {
            const uint64_t so = A.seq_off[i];
            if (i - lb < (uint32_t)SOC) S.soc[i - lb + 1] = so;
            if (so < (uint64_t)T1) {
                const uint64_t len = A.seq_off[i + 1] - so;
                flag_raw(S, par, (uint32_t)((int64_t)so - W0), len > 0 && len <= (uint64_t)l);
            }
        }
        if (tid == 0) {
            const uint64_t so_lb = A.seq_off[lb];
            const uint64_t so_pr = lb ? A.seq_off[lb - 1] : 0ull;
            S.soc[0] = so_pr;
            unsigned long long s0 = (unsigned long long)T0;
            if (so_lb != (uint64_t)T0) {                   // the sequence containing T0 started earlier
                s0 = so_pr;
                if ((int64_t)s0 >= W0) {
                    flag_raw(S, par, (uint32_t)((int64_t)s0 - W0), so_lb - s0 <= (uint64_t)l);
                }
            }
            S.s0 = s0;
        }
        __syncthreads();
        PHASE(1);

        // ---- S3b: keep masks (16 bits per piece), block scan of kept counts
        uint32_t k16[4], st16[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            st16[j] = reinterpret_cast<const uint16_t *>(S.startw)[128 * warp + 32 * j + lane];   // the piece's half of its chunk word
        }
        if (HPC) {
            // keep bit = byte differs from the byte before it.  Per word: PRMT lines the previous bytes up, XOR, the
            // classic "byte is non-zero" carry trick leaves bit 7 of each byte, one multiply gathers the four bits into
            // the top nibble and a funnel shift appends it -- words are visited from high to low so that base 0 ends up
            // in bit 0.  The byte before a piece is the last byte of the lane below (lane 0: lane 31 of the round before;
            // round 0: the last byte of the warp before, from global memory).
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t send = (j > 0 && lane == 31) ? w[4 * j - 1] : w[4 * j + 3];
                uint32_t prevw = __shfl_sync(0xffffffffu, send, (lane + 31) & 31);
                if (j == 0 && lane == 0) {
                    const int64_t g = W0 + xw;
                    prevw = ((g > 0 && g <= (int64_t)A.n_bases) ? (uint32_t)A.bases[g - 1] : 0u) << 24;
                }
                uint32_t kk = 0u;
#pragma unroll
                for (int i = 3; i >= 0; --i) {
                    const uint32_t x = w[4 * j + i] ^ __byte_perm(i ? w[4 * j + i - 1] : prevw, w[4 * j + i], 0x6543u);
                    const uint32_t nz = (x | mad_u32(one_r, x & 0x7f7f7f7fu, 0x7f7f7f7fu)) & 0x80808080u;   // the add as an IMAD
                    kk = __funnelshift_l(nz * 0x00204081u, kk, 4);
                }
                k16[j] = kk | st16[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) k16[j] = 0xffffu;
        }
        const int x_hi = (int)(T1 - W0);                   // window offsets [x_lo, x_hi) hold the bases of [max(W0,0), T1)
        if (!full || x_hi < WIN) {
            const int x_lo = W0 < 0 ? (int)min((int64_t)WIN, -W0) : 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x0 = xw + 512 * j;
                const int a = min(max(x_lo - x0, 0), 16), e = min(max(x_hi - x0, 0), 16);
                k16[j] &= e > a ? (lowmask((uint32_t)e) & ~lowmask((uint32_t)a)) : 0u;
            }
        }
        // One warp scan per pair of rounds (counts of a round sum to at most 512: 16-bit fields), then the rounds'
        // totals: kept bases before piece (j, lane) = before the warp + rounds below j + lanes below in round j.
        const uint32_t c0 = __popc(k16[0]), c1 = __popc(k16[1]), c2 = __popc(k16[2]), c3 = __popc(k16[3]);
        const uint32_t i01 = warp_incl_scan_m(c0 | (c1 << 16), lane, one_r), i23 = warp_incl_scan_m(c2 | (c3 << 16), lane, one_r);
        const uint32_t t01 = __shfl_sync(0xffffffffu, i01, 31), t23 = __shfl_sync(0xffffffffu, i23, 31);
        uint32_t qj[4];
        qj[0] = (i01 & 0xffffu) - c0;
        qj[1] = (t01 & 0xffffu) + (i01 >> 16) - c1;
        qj[2] = (t01 & 0xffffu) + (t01 >> 16) + (i23 & 0xffffu) - c2;
        qj[3] = (t01 & 0xffffu) + (t01 >> 16) + (t23 & 0xffffu) + (i23 >> 16) - c3;
        if (lane == 31) S.wsum[0][warp] = (t01 & 0xffffu) + (t01 >> 16) + (t23 & 0xffffu) + (t23 >> 16);
        // The piece at the halo/tile boundary sits in warp 0 (halo <= 512 raw bases): its prefix needs no warp totals,
        // so the kept count of the halo is published before the scan's only barrier.
        if (warp == 0 && lane == (int)((A.halo >> 4) & 31u)) S.hk = (A.halo >> 9) ? qj[1] : qj[0];
        __syncthreads();
        uint32_t qw, wk;
        warp_totals(S.wsum[0], warp, lane, one_r, qw, wk);
        {
            uint16_t *kw16 = reinterpret_cast<uint16_t *>(S.keepw);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                qj[j] += qw;
                const int pc = 128 * warp + 32 * j + lane;  // piece index in the window; chunk = pc >> 1
                kw16[pc] = (uint16_t)k16[j];
                if (!(lane & 1)) S.qoff[pc >> 1] = qj[j];
                const uint32_t m = (qj[j] + 63u) & ~63u;    // a piece holds at most 16 kept bases: at most one multiple of 64
                if (m < qj[j] + __popc(k16[j])) S.qmap[m >> 6] = (uint16_t)(pc >> 1);
            }
        }
        if (tid == NT - 1) { S.qoff[NCHUNK] = wk; S.keepw[NCHUNK] = 0; }

        PHASE(2);
        // ---- S4: compaction.  Packed form (2 bits per kept base, register-gathered) unless the halo is too short for
        // the left context (walk-back below) or l is beyond its warm-up; a rare class met on the way turns the tile back
        // to the byte form after the barrier.
        const uint32_t hk_real = S.hk;
        const bool need_walk = HPC && (int64_t)S.s0 < W0 && hk_real < A.need;
        const bool try_packed = !H64 && !need_walk && l <= PK_LMAX && !A.trunc16;
        if (try_packed) {
            if (compact_packed(S, w, k16, qj, A.vmask)) S.rare = 1u;
        } else {
            if (S.code_raw) reset_codes(S);                // uniform; every thread read its pieces before the last barrier
            compact_bytes<HPC>(S, xft, w, k16, qj);
        }
        if (tid == 0) S.next[par ^ 1] = t_next;           // the ticket drawn at the top has long arrived
        PHASE(3);
        // Owner space starts up to 3 kept bases inside the halo (those pseudo-owners are masked out below) so that
        // every thread's class bytes begin on a word boundary of S.code: the hash stage reads them as words.
        const int dlt = (int)((hk_real - (uint32_t)d) & 3u);
        const int hk = (int)hk_real - dlt;
        if (st16[0] | st16[1] | st16[2] | st16[3]) {       // sequence starts among this thread's kept bases -> owner-space flags
#ifndef S2K_START_LOOP_PER_PIECE
            // One loop over the starts of all four pieces: on 150-bp reads a thread holds 0.4 starts, some lane of every
            // warp has one in every piece, and a loop per piece made each warp walk four divergent branches (3.9 of
            // config 3's 36.3 instructions per base; profiles/r2_c3_phases.txt).
            uint32_t s01 = (st16[0] & k16[0]) | ((st16[1] & k16[1]) << 16), s23 = (st16[2] & k16[2]) | ((st16[3] & k16[3]) << 16);
            while (s01 | s23) {
                const bool up = s01 == 0u;
                const int b = __ffs((int)(up ? s23 : s01)) - 1;
                if (up) s23 &= s23 - 1u; else s01 &= s01 - 1u;
                const int hi = b >> 4, bb = b & 15, j = (up ? 2 : 0) + hi;
                const uint32_t kj = up ? (hi ? k16[3] : k16[2]) : (hi ? k16[1] : k16[0]);
                const uint32_t qq = up ? (hi ? qj[3] : qj[2]) : (hi ? qj[1] : qj[0]);
                const uint32_t sh2 = S.shortw[64 * warp + 16 * j + (lane >> 1)] >> (16 * (lane & 1) + bb);
                const int oo = (int)qq + __popc(kj & lowmask((uint32_t)bb)) - hk + XB;
                if (oo >= 0) flag_owner(S, par, oo, sh2 & 1u);
            }
#else
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t sw = st16[j] & k16[j];
                if (sw) {
                    const uint32_t sh2 = S.shortw[64 * warp + 16 * j + (lane >> 1)] >> (16 * (lane & 1));
                    while (sw) {
                        const int b = __ffs((int)sw) - 1;
                        sw &= sw - 1;
                        const int oo = (int)qj[j] + __popc(k16[j] & lowmask((uint32_t)b)) - hk + XB;
                        if (oo >= 0) flag_owner(S, par, oo, (sh2 >> b) & 1u);
                    }
                }
            }
#endif
        }
        // ---- S4b: not enough context in the halo -> walk back through the sequence (rare: long homopolymers)
        if (need_walk && warp == 0) {
            uint32_t remaining = A.need - hk_real, taken = 0;
            const int64_t s0 = (int64_t)S.s0;
            int64_t hi = W0;
            while (remaining > 0 && hi > s0) {
                const int64_t lo = max(s0, hi - 32);
                const int64_t g = lo + lane;
                const bool valid = g < hi;
                uint8_t b = 0, pb = 0;
                if (valid) { b = A.bases[g]; if (g > s0) pb = A.bases[g - 1]; }
                const bool kp = valid && (g == s0 || b != pb);
                const uint32_t m = __ballot_sync(0xffffffffu, kp);
                const uint32_t above = (lane == 31) ? 0u : (m >> (lane + 1));
                const uint32_t rank = __popc(above);
                if (kp && rank < remaining) {
                    const uint32_t slot = taken + rank;            // 0 = nearest to the window
                    S.code[XB - 1 - (int)slot] = S.lut[b];
                    S.ctxpos[slot] = (uint32_t)(W0 - g);
                    if (g == s0) { const int oo = XB - 1 - (int)slot - hk; if (oo >= 0) flag_owner(S, par, oo, false); }
                }
                const uint32_t c = min((uint32_t)__popc(m), remaining);
                taken += c; remaining -= c; hi = lo;
            }
        }
        __syncthreads();
        const bool packed = try_packed && S.rare == 0u;
        if (try_packed && !packed) {                       // uniform, rare: N, IUPAC, lower case ... -> byte form after all
            if (S.code_raw) reset_codes(S);
            load_pieces(A, W0, xw, full, w);               // (the registers were not kept across the barrier)
            compact_bytes<HPC>(S, xft, w, k16, qj);
            __syncthreads();
        }
        PHASE(4);
        const uint32_t tn = S.next[par ^ 1];
#ifndef S2K_EMU
        {   // the next tile's window: one bulk copy (TMA) into code[], which a packed tile does not use; tiles in byte
            // form (and windows that stick out of the batch) only pull the bases into L2
            const int64_t wn = (int64_t)((uint64_t)tn * A.tile) - (int64_t)A.halo;
#ifdef S2K_TMA
            const bool bulk = packed && tn < A.n_tiles && wn >= 0 && wn + (int64_t)WIN <= (int64_t)A.n_bases;
#else
            const bool bulk = false;
#endif
#ifdef S2K_TMA
            if (tid == 0) {
                S.tma_tile = bulk ? tn : ~0u;
                if (bulk) S.code_raw = 1u;
                if (bulk) tma_load_1d(S.code, A.bases + wn, (uint32_t)WIN, &S.mbar);
            }
#endif
            if (!bulk && tn < A.n_tiles && tid < WIN / 128) {
                const int64_t a = wn + 128 * tid;
                if (a >= 0 && a < (int64_t)A.n_bases) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.bases + a));
            }
        }
#endif

        // ---- S5 + S6a: rolling canonical ntHash over the owners, CAP per pass; block scan of hit counts
        const uint32_t n_own = wk - (uint32_t)hk;         // kept bases in [T0, T1) (+dlt): each completes one l-mer
        uint32_t tile_min = 0;
        const int n_pass = n_own > (uint32_t)CAP ? 2 : 1;  // the second pass exists for badly compressing HPC tiles only
#pragma unroll 1
        for (int pass = 0; pass < n_pass; ++pass) {
            unsigned long long mask[MW];
#pragma unroll
            for (int x = 0; x < MW; ++x) mask[x] = 0ull;
            const int v0 = pass * CAP + CH * tid;
            if (tid < HT && (uint32_t)v0 < n_own) {
                const int n_u = min(CH, (int)n_own - v0);
                // owners invalidated by sequence starts: a start f kills owners [f, f+l-2+d] (+1 if len<=l)
                unsigned long long invalid[MW];
#pragma unroll
                for (int x = 0; x < MW; ++x) invalid[x] = 0ull;
                if (S.n_dirty[par] != 0) {                         // uniform: no flag was set in this tile -> nothing to mask
                    const int L1 = l - 1 + d, o0 = v0 + XB;
                    const int w_hi = (o0 + CH - 1) >> 5, w_lo = (o0 - L1 - 1) >> 5;
                    for (int wi = w_lo; wi <= w_hi; ++wi) {
                        uint32_t fw = S.f1[wi];
                        if (fw) {
                            const uint32_t sw2 = S.f2[wi];
                            while (fw) {
                                const int b = __ffs(fw) - 1;
                                fw &= fw - 1;
                                int lo = wi * 32 + b - o0;
                                int hi = lo + L1 + (int)((sw2 >> b) & 1u);
                                lo = max(lo, 0); hi = min(hi, CH);
#pragma unroll
                                for (int x = 0; x < MW; ++x) {
                                    const int a = max(lo - 64 * x, 0), e = min(hi - 64 * x, 64);
                                    if (e > a) invalid[x] |= lowmask64(e - a) << a;
                                }
                            }
                        }
                    }
                }
                if (v0 == 0) invalid[0] |= lowmask64((uint32_t)dlt);   // the pseudo-owners inside the halo
                if (packed) {
                    hash_owners_packed<W31, DENSE>(S, xft, XB + hk + v0 - d, l, A.thr, hs + v0, mask, one_r);
                } else {
                    const uint8_t *cb = S.code + XB + hk + v0 - d; // cb[i]: last base of owner i's l-mer; 4-aligned
                    if (H64) {
                        hash_owners_bytes64(S, cb, l, A.thr64, hs + v0, mask);
                    } else if (!W31 && A.trunc16) {
                        hash_owners_bytes<false, true>(S, cb, l, A.thr, hs + v0, mask);
                    } else {
                        const uint32_t rare = hash_owners_words<W31, DENSE>(S, xft, cb, l, A.thr, hs + v0, mask);
                        if (rare & RARE4) {                        // a rare class among the bytes touched: redo via xy
#pragma unroll
                            for (int x = 0; x < MW; ++x) mask[x] = 0ull;
                            hash_owners_bytes<W31>(S, cb, l, A.thr, hs + v0, mask);
                        }
                    }
                }
#pragma unroll
                for (int x = 0; x < MW; ++x)                       // owners >= n_u hashed garbage
                    mask[x] &= ~invalid[x] & lowmask64((uint32_t)max(min(n_u - 64 * x, 64), 0));
            }
            uint32_t cnt = 0;
#pragma unroll
            for (int x = 0; x < MW; ++x) cnt += __popcll(mask[x]);
            uint32_t tot;
            const uint32_t ex = block_excl_scan(cnt, S.wsum[1 + pass], tot, one_r);
#pragma unroll
            for (int x = 0; x < MW; ++x) S.hitw[pass][tid][x] = mask[x];
            S.hitpre[pass][tid] = tile_min + ex;
            tile_min += tot;
            if (tid == NT - 1) {
#pragma unroll
                for (int x = 0; x < MW; ++x) S.hitw[pass][NT][x] = 0ull;
                S.hitpre[pass][NT] = tile_min;
            }
        }

        PHASE(5);
        // ---- S6b: claim a contiguous run of records for this tile (tiles land in any order; k_finalize sorts them out).
        // Thread 0 fires the atomic and the loads of the next tile's sequence bounds here and picks the results up
        // after the barrier / at the end of the tile, so that nobody waits for their latency.
        unsigned long long r0 = 0ull, lim = A.min_cap;
        uint32_t nlb0 = 0, nlb1 = 0;
        if (tid == 0) {
            if (A.region_cap) {
                const unsigned long long c = S.cur;
                r0 = (unsigned long long)blockIdx.x * A.region_cap + c;
                lim = ((unsigned long long)blockIdx.x + 1ull) * A.region_cap;
                S.cur = c + tile_min;
            } else if (tile_min) r0 = atomicAdd(A.cursor, (unsigned long long)tile_min);
            if (tn < A.n_tiles) {
                nlb0 = A.tile_lb[tn];
                nlb1 = tn + 1 == A.n_tiles ? (uint32_t)(A.n_seqs + 1) : A.tile_lb[tn + 1];
            }
        }
        __syncthreads();                                   // hit masks and prefixes of all threads are in place
        if (try_packed) {                                  // the packed array has been read: clean for the next tile's ORs
            for (int i = tid; i < PKW / 4; i += NT) reinterpret_cast<uint4 *>(S.pk)[i] = make_uint4(0u, 0u, 0u, 0u);
            if (tid == 0) S.rare = 0u;
        }
        if (tid == 0) {
            S.rec0 = r0; S.rec_lim = lim;                  // read after the barrier inside the emission loop
            A.tile_info[t] = make_uint4(tile_min, wk - hk_real, (uint32_t)r0, (uint32_t)(r0 >> 32));
            if (r0 + tile_min > lim) atomicOr(A.err, ERR_CAP);
        }
        PHASE(6);

        // ---- S7: ordered hit list in shared memory, then one thread per minimizer
        for (uint32_t base = 0; base < tile_min; base += HL) {
            if (base) __syncthreads();                     // previous round has been consumed
#pragma unroll 1
            for (int pass = 0; pass < n_pass; ++pass) {
                uint32_t o = S.hitpre[pass][tid];
#pragma unroll
                for (int x = 0; x < MW; ++x) {
                    unsigned long long m = S.hitw[pass][tid][x];
                    if (tile_min <= (uint32_t)HL) {        // the usual case, one round: no range test, 32-bit halves
                        uint32_t mlo = (uint32_t)m, mhi = (uint32_t)(m >> 32);
                        const uint32_t v0 = (uint32_t)(pass * CAP + CH * tid + 64 * x);
                        while (mlo) { S.hl[o++] = (uint16_t)(v0 + (uint32_t)__ffs((int)mlo) - 1u); mlo &= mlo - 1u; }
                        while (mhi) { S.hl[o++] = (uint16_t)(v0 + 31u + (uint32_t)__ffs((int)mhi)); mhi &= mhi - 1u; }
                        continue;
                    }
                    while (m) {
                        const int i = __ffsll((long long)m) - 1 + 64 * x;
                        m &= m - 1;
                        if (o >= base && o < base + HL) S.hl[o - base] = (uint16_t)(pass * CAP + CH * tid + i);
                        ++o;
                    }
                }
            }
            __syncthreads();
            const uint64_t rec0 = S.rec0, rec_lim = S.rec_lim;
            const bool cached = ub - lb <= (uint32_t)SOC;
            const uint32_t n_round = min((uint32_t)HL, tile_min - base);
            for (uint32_t j = tid; j < n_round; j += NT) {
                const int v = S.hl[j];
                const int qo = hk + v;                     // window index of the owner base
                const uint32_t h = hs[v];
                int c = chunk_of(S, qo);
                const int64_t g_own = pos_in_chunk<HPC>(S, W0, c, qo);
                const int qs = qo - (l - 1 + d);           // first base of the l-mer: a little further left
                int64_t g_start;
                if (qs < 0) g_start = W0 - (int64_t)S.ctxpos[-1 - qs];
                else { while ((int)S.qoff[c] > qs) --c; g_start = pos_in_chunk<HPC>(S, W0, c, qs); }
                uint32_t lo = lb, hi = ub;                 // first i in [lb,ub) with seq_off[i] > g_own
                while (lo < hi) {
                    const uint32_t mid = lo + ((hi - lo) >> 1);
                    const uint64_t sm = cached ? S.soc[mid - lb + 1] : A.seq_off[mid];
                    if (sm <= (uint64_t)g_own) lo = mid + 1; else hi = mid;
                }
                const uint32_t rid = lo - 1;
                const uint64_t so = cached ? S.soc[lo - lb] : A.seq_off[rid];
                const uint64_t idx = rec0 + base + j;
                if (idx < rec_lim) {
                    A.min_out[idx] = make_uint4(h, (uint32_t)((uint64_t)g_start - so),
                                                (uint32_t)((uint64_t)g_own - (uint64_t)d - so), rid);
                    if (H64) A.min_hi[idx] = hs[WIN + v];
                }
            }
        }
        PHASE(7);
        // ---- S8: per-sequence offsets for every sequence starting in this tile
        for (uint32_t i = lb + tid; i < ub; i += NT) {
            const uint64_t so = i - lb < (uint32_t)SOC ? S.soc[i - lb + 1] : A.seq_off[i];
            const uint32_t x = (uint32_t)((int64_t)so - W0);
            const uint32_t qx = S.qoff[x >> 5] + __popc(S.keepw[x >> 5] & lowmask(x & 31));
            const uint32_t v = qx - (uint32_t)hk;
            uint32_t hb = tile_min;                        // a start at the very end of the tile: every hit precedes it
            if (v < n_own) {
                const uint32_t pass = v >= (uint32_t)CAP ? 1u : 0u;
                const uint32_t vv = v - pass * CAP, u = vv / CH, bit = vv - u * CH;
                hb = S.hitpre[pass][u];
#pragma unroll
                for (int x = 0; x < MW; ++x)
                    hb += __popcll(S.hitw[pass][u][x] & lowmask64((uint32_t)max(min((int)bit - 64 * x, 64), 0)));
            }
            A.min_off[i] = hb;
            if (A.hpc_off) A.hpc_off[i] = qx - hk_real;
        }
        {   // wipe the flag words this tile touched (their last readers passed a barrier in the hash stage)
            const uint32_t nd = S.n_dirty[par];
            if (nd > DIRTY_MAX) {
                for (int i = tid; i < NCHUNK; i += NT) { S.startw[i] = 0; S.shortw[i] = 0; }
                for (int i = tid; i < FW; i += NT) { S.f1[i] = 0; S.f2[i] = 0; }
            } else if ((uint32_t)tid < nd) {
                const uint32_t e = S.dirty[par][tid];
                if (e & 0x8000u) { S.f1[e & 0x7fffu] = 0; S.f2[e & 0x7fffu] = 0; }
                else { S.startw[e] = 0; S.shortw[e] = 0; }
            }
        }
        if (tid == 0) { S.nlb[par ^ 1][0] = nlb0; S.nlb[par ^ 1][1] = nlb1; }
        par ^= 1;
        PHASE(8);
    }
    if (tid == 0 && A.region_cap) atomicAdd(A.cursor, S.cur);     // the host reads the total from the cursor
    PHASE_FLUSH;
}


// ------------------------------------------------------------------------------------------------ tile order
// k_minimizers leaves each tile's records contiguous but the tiles in completion order.  Two-level exclusive prefix
// of the per-tile (hits, kept bases): k_tile_scan_a = one thread per tile, block scan over 1024 tiles;
// k_tile_scan_b = one CTA over the chunk totals.  k_finalize = one warp per tile: copy the records to their final,
// ordered place and turn the tile-local per-sequence prefixes into global ones.
constexpr int ST = 1024;
__global__ void __launch_bounds__(ST) k_tile_scan_a(const uint4 *__restrict__ tile_info, uint32_t n_tiles,
                                                    unsigned long long *__restrict__ tile_loc,
                                                    unsigned long long *__restrict__ chunk_tot)
{
    S2K_SHARED unsigned long long ws[ST / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t t = blockIdx.x * ST + tid;
    unsigned long long v = 0;                              // hits | kept << 32: a chunk's sums stay below 2^32
    if (t < n_tiles) { const uint4 x = tile_info[t]; v = (unsigned long long)x.x | ((unsigned long long)x.y << 32); }
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += y;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    unsigned long long pre = 0, tot = 0;
    for (int i = 0; i < ST / 32; ++i) { const unsigned long long y = ws[i]; if (i < (int)warp) pre += y; tot += y; }
    if (t < n_tiles) tile_loc[t] = pre + incl - v;
    if (tid == 0) chunk_tot[blockIdx.x] = tot;
}
// `cap`: records the result buffers were sized for; more minimizers than that raise ERR_CAP here, before any kernel
// writes through the ordered indices.
__global__ void __launch_bounds__(ST) k_tile_scan_b(const unsigned long long *__restrict__ chunk_tot, uint32_t n_chunks,
                                                    ulonglong2 *__restrict__ chunk_base, unsigned long long cap, uint32_t *err)
{
    S2K_SHARED unsigned long long sa[ST], sb[ST];
    S2K_SHARED unsigned long long carry[2];
    const uint32_t tid = threadIdx.x;
    if (tid == 0) { carry[0] = 0; carry[1] = 0; }
    for (uint32_t c0 = 0; c0 < n_chunks; c0 += ST) {
        __syncthreads();
        const uint32_t c = c0 + tid;
        const unsigned long long v = c < n_chunks ? chunk_tot[c] : 0ull;
        const unsigned long long a = v & 0xffffffffull, b = v >> 32;
        sa[tid] = a; sb[tid] = b;
        __syncthreads();
        for (uint32_t o = 1; o < ST; o <<= 1) {
            unsigned long long xa = 0, xb = 0;
            if (tid >= o) { xa = sa[tid - o]; xb = sb[tid - o]; }
            __syncthreads();
            sa[tid] += xa; sb[tid] += xb;
            __syncthreads();
        }
        if (c < n_chunks) chunk_base[c] = make_ulonglong2(carry[0] + sa[tid] - a, carry[1] + sb[tid] - b);
        __syncthreads();
        if (tid == ST - 1) { carry[0] += sa[tid]; carry[1] += sb[tid]; }
    }
    __syncthreads();
    if (tid == 0) {
        chunk_base[n_chunks] = make_ulonglong2(carry[0], carry[1]);
        if (carry[0] > cap) atomicOr(err, ERR_CAP);
    }
}

struct KFArgs {
    const uint4 *tile_info;
    const unsigned long long *tile_loc;
    const ulonglong2 *chunk_base;
    const uint4 *tmp;
    uint4 *mins;
    uint64_t min_cap;
    uint32_t n_tiles;
    ulonglong2 *tile_pre;        // n_tiles + 1: (minimizers, kept bases) before the tile; entry n_tiles = the totals
    ulonglong2 *tile_src;        // or null: per tile (ordered index of its first record, where its records sit in tmp)
    const uint32_t *tmp_hi;      // H = u64 flavour: high halves of the hashes, beside tmp / mins (else null)
    uint32_t *mins_hi;
    int32_t copy;                // 0: leave the records where they are (the window stage reads them through tile_src)
    uint32_t hmask;              // AND-ed into the hash of every copied record (0xffff for the H = u16 flavour, else ~0)
    const uint32_t *err;         // ERR_CAP set: the record store overflowed, the host reruns -- touch nothing
};
__global__ void __launch_bounds__(256) k_finalize(const __grid_constant__ KFArgs A)
{
    if (*A.err & ERR_CAP) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t <= A.n_tiles; t += nwarps) {
        if (t == A.n_tiles) {                              // the totals, for sequences that start at the very end
            const ulonglong2 tot = A.chunk_base[(A.n_tiles + ST - 1) / ST];
            if (lane == 0) A.tile_pre[t] = tot;
            continue;
        }
        const uint4 info = A.tile_info[t];
        const ulonglong2 cb = A.chunk_base[t / ST];
        const unsigned long long loc = A.tile_loc[t];
        const uint64_t bm = cb.x + (loc & 0xffffffffull), bk = cb.y + (loc >> 32);
        const uint64_t src = ((uint64_t)info.w << 32) | info.z;
        if (lane == 0) {
            A.tile_pre[t] = make_ulonglong2(bm, bk);
            if (A.tile_src) A.tile_src[t] = make_ulonglong2(bm, src);
        }
        if (A.copy && src + info.x <= A.min_cap)
            for (uint32_t j = lane; j < info.x; j += 32) {
                uint4 rec = A.tmp[src + j];
                rec.x &= A.hmask;
                A.mins[bm + j] = rec;
                if (A.tmp_hi) A.mins_hi[bm + j] = A.tmp_hi[src + j];
            }
    }
}

// ------------------------------------------------------------------------------------------------ read counts
constexpr int RT = 256;            // threads per CTA
constexpr int RPT = 4;             // sequences per thread
struct K2Args {
    const uint4    *mins;
    const uint64_t *min_loc, *hpc_loc;   // n_seqs + 1, tile-LOCAL prefixes written by k_minimizers (hpc_loc may be null)
    const uint64_t *seq_off;
    const uint8_t  *bases;
    const ulonglong2 *tile_pre;          // n_tiles + 1: (minimizers, kept bases) before each tile (k_finalize)
    uint64_t  n_seqs, n_bases;
    uint32_t  l, k;
    int32_t   quirk, hpc;
    uint64_t *min_off;             // n_seqs + 1, out: global exclusive prefix of per-sequence minimizer counts
    uint64_t *km_off;              // n_seqs + 1
    uint32_t *min_cnt;             // n_seqs
    uint64_t *status;              // per tile, zeroed
    uint32_t *ticket;
    uint32_t *err;
    const ulonglong2 *tile_src;    // non-null: `mins` is the unordered record store of k_minimizers (tiles contiguous,
    uint32_t  n_tiles, tile;       // in completion order) and ordered index m is looked up through tile_src
    uint64_t  tile_magic;          // floor(2^64 / tile) + 1: position / tile as one 64-bit multiply-high (+ a fix-up)
};
// x / tile without the 64-bit division sequence (five of them per thread made k_read_counts 3x slower).
__device__ __forceinline__ uint32_t tile_of_pos(const K2Args &A, uint64_t x)
{
#ifdef S2K_EMU
    uint64_t q = (uint64_t)(((unsigned __int128)x * A.tile_magic) >> 64);
#else
    uint64_t q = __umul64hi(x, A.tile_magic);
#endif
    if (q * A.tile > x) --q;                               // the estimate is exact or one too large
    return (uint32_t)q;
}

// Tile holding ordered record m: the largest t with tile_src[t].x <= m.  `hint` is a tile near it (the caller knows a
// base position close to the minimizer): a step or two, except across runs of tiles without minimizers (a homopolymer
// of megabases), where the walk turns into a binary search after 8 steps.
__device__ __forceinline__ uint32_t tile_of_record(const ulonglong2 *tile_src, uint32_t n_tiles, uint64_t m, uint32_t hint)
{
    uint32_t t = hint < n_tiles ? hint : n_tiles - 1;
    for (int steps = 0; t > 0 && tile_src[t].x > m; --t) {
        if (++steps > 8) {
            uint32_t lo = 0, hi = t;                       // tile_src[lo].x <= m < tile_src[hi].x
            while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (tile_src[mid].x <= m) lo = mid; else hi = mid; }
            return lo;                                     // nothing above lo qualifies
        }
    }
    for (int steps = 0; t + 1 < n_tiles && tile_src[t + 1].x <= m; ++t) {
        if (++steps > 8) {
            uint32_t lo = t + 1, hi = n_tiles;             // tile_src[lo].x <= m; hi = one past the last tile
            while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (tile_src[mid].x <= m) lo = mid; else hi = mid; }
            return lo;
        }
    }
    return t;
}

// Minimizers of a sequence that reach the window stage.  In the AVX-512 profile of ntHash1 the iterator
// masks the last block with (1 << (S % 16)) - 1 (src/nthash_avx512_32.rs:134-138): when S > 16 and
// S % 16 == 0 the last 16 l-mer positions are lost.  Their minimizers are a suffix of the sequence's list:
// exactly those whose last base is one of the final 16 kept bases.  Called by the threads whose sequence is under the
// rule (one in sixteen): the lanes of a warp walk their own sequences side by side, the ~21 bases of a walk sit in one
// or two cache lines.  (Measured alternatives, config 3: the last 64 bases as unrolled independent loads 8.8 ms, one
// warp per sequence with ballots 5.4 ms, this 2.x ms for the whole kernel.)
__device__ __forceinline__ uint32_t tail_rule_count(const K2Args &A, uint64_t so, uint64_t se, uint64_t m0, uint64_t cnt)
{
    uint64_t e16;                                       // local position of the 16th kept base from the end
    if (!A.hpc) {
        e16 = (se - so) - 16;
    } else {
        uint64_t g = se;
        int found = 0;
        // Fast form: the last 48 bases as six aligned words fetched at once (the byte-wise walk below is a chain of some
        // twenty dependent loads, and with short reads one sequence in sixteen comes here), keep bits of the 47 bases
        // whose predecessor is among them, the 16th set bit from the top.  Fewer than 16 kept bases in that stretch, a
        // sequence shorter than it or the end of the batch: the walk.
        const uint64_t p = (se - 41) & ~7ull;
        if (se >= so + 48 && p + 48 <= A.n_bases && (reinterpret_cast<uintptr_t>(A.bases) & 7u) == 0) {
            const unsigned long long *wp = reinterpret_cast<const unsigned long long *>(A.bases + p);
            unsigned long long wv[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) wv[i] = wp[i];
            unsigned long long km = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const unsigned long long x = wv[i] ^ ((wv[i] << 8) | (i ? wv[i - 1] >> 56 : 0ull));
                const unsigned long long nz = (((x & 0x7f7f7f7f7f7f7f7full) + 0x7f7f7f7f7f7f7f7full) | x) & 0x8080808080808080ull;
                km |= (((nz >> 7) * 0x0102040810204080ull) >> 56) << (8 * i);
            }
            km &= ~1ull;                                // the first byte has no predecessor here
            const uint32_t n_in = (uint32_t)(se - p);   // bytes [p, se) belong to the sequence: 41 .. 48 of them
            if (n_in < 48) km &= (1ull << n_in) - 1ull;
            if (__popcll(km) >= 16) {
#pragma unroll 1
                for (int i = 0; i < 15; ++i) km &= ~(1ull << (63 - __clzll((long long)km)));
                g = p + (uint64_t)(63 - __clzll((long long)km));
                found = 16;
            }
        }
        while (found < 16) {                            // M >= 16 guarantees termination above `so`
            --g;
            if (g == so || A.bases[g] != A.bases[g - 1]) ++found;
        }
        e16 = g - so;
    }
    if (!A.tile_src) {
        while (cnt > 0 && (uint64_t)A.mins[m0 + cnt - 1].z >= e16) --cnt;
    } else {
        uint32_t t = tile_of_pos(A, se - 1);
        while (cnt > 0) {
            const uint64_t m = m0 + cnt - 1;
            t = tile_of_record(A.tile_src, A.n_tiles, m, t);
            const ulonglong2 ts = A.tile_src[t];
            if ((uint64_t)A.mins[ts.y + (m - ts.x)].z < e16) break;
            --cnt;
        }
    }
    return (uint32_t)cnt;
}

// Per sequence: the tile-local prefixes of k_minimizers made global (the tile of a sequence = the tile of its first
// base), minimizers that feed the window stage, item counts, and their exclusive scan (decoupled look-back over tiles
// of RT * RPT sequences).  One pass over the per-sequence arrays: 16-24 bytes read, 20 written per sequence.
// k_read_counts: per sequence, the global minimizer prefix (min_off) and the number of minimizers that reach the window
// stage (min_cnt: the AVX-512 tail rule applied).  A warp takes 32 * RPT consecutive sequences, lane L the sequences
// base + 32 j + L (coalesced loads; the next sequence's prefixes arrive by shuffle); no block barrier anywhere.  The
// sequences the rule applies to (one in sixteen with short reads) are listed per warp in shared memory and settled one
// per lane: inside the per-sequence loop some lane of almost every warp would take that chain of dependent loads and the
// whole warp would wait for it once per sequence.  The exclusive scan of the item counts is k_item_scan's.
__global__ void __launch_bounds__(RT) k_read_counts(const __grid_constant__ K2Args A)
{
    if (*A.err & ERR_CAP) return;                          // the record store overflowed: the host reruns
    constexpr int WT = 32 * RPT;                           // sequences per warp pass
    S2K_SHARED uint16_t q_slot[RT / 32][WT];               // 32 j + lane of a listed sequence
    S2K_SHARED uint32_t q_cnt[RT / 32][WT];                // its minimizers before the rule
    S2K_SHARED uint32_t q_res[RT / 32][WT];                // by slot: minimizers after it
    S2K_SHARED unsigned long long q_m0[RT / 32][WT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t n_pass = A.n_seqs / WT + 1;             // the last pass also writes min_off[n_seqs]
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t ps = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ps < n_pass; ps += n_warps) {
        const uint64_t base = ps * WT;
        uint64_t gm[RPT + 1], gk[RPT + 1];                 // global prefixes of sequences base + 32 j + lane (j = RPT: lane 0 only)
#pragma unroll
        for (int j = 0; j <= RPT; ++j) {
            gm[j] = 0; gk[j] = 0;
            const uint64_t r = base + 32 * j + lane;
            if (r <= A.n_seqs && (j < RPT || lane == 0)) {
                const uint64_t so = A.seq_off[r];
                uint32_t ts = tile_of_pos(A, so);
                if (ts >= A.n_tiles) ts = A.n_tiles - 1;
                const ulonglong2 pre = A.tile_pre[ts];
                gm[j] = A.min_loc[r] + pre.x;
                gk[j] = A.hpc_loc ? A.hpc_loc[r] + pre.y : so;
            }
        }
        uint32_t cj[RPT], ruled = 0, n_list = 0;
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const uint64_t r = base + 32 * j + lane;
            // prefixes of sequence r + 1: the lane above, or lane 0 of the next row
            const uint64_t gm_up = __shfl_down_sync(0xffffffffu, gm[j], 1), gk_up = __shfl_down_sync(0xffffffffu, gk[j], 1);
            const uint64_t gm_row = __shfl_sync(0xffffffffu, gm[j + 1], 0), gk_row = __shfl_sync(0xffffffffu, gk[j + 1], 0);
            const uint64_t gm1 = lane == 31 ? gm_row : gm_up, gk1 = lane == 31 ? gk_row : gk_up;
            cj[j] = r < A.n_seqs ? (uint32_t)(gm1 - gm[j]) : 0u;
            bool need = false;
            if (A.quirk && r < A.n_seqs && cj[j] > 0) {
                const uint64_t M = gk1 - gk[j];
                need = M >= (uint64_t)A.l + 16 && ((M - A.l + 1) & 15) == 0;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, need);
            if (need) {
                const uint32_t e = n_list + (uint32_t)__popc(m & ((1u << lane) - 1u));
                q_slot[warp][e] = (uint16_t)(32 * j + lane); q_cnt[warp][e] = cj[j]; q_m0[warp][e] = gm[j];
                ruled |= 1u << j;
            }
            n_list += (uint32_t)__popc(m);
        }
        if (n_list) {                                      // warp-uniform
            __syncwarp();
            for (uint32_t e = lane; e < n_list; e += 32) {
                const uint32_t sl = q_slot[warp][e];
                const uint64_t r = base + sl;
                q_res[warp][sl] = tail_rule_count(A, A.seq_off[r], A.seq_off[r + 1], q_m0[warp][e], q_cnt[warp][e]);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < RPT; ++j)
                if ((ruled >> j) & 1u) cj[j] = q_res[warp][32 * j + lane];
            __syncwarp();                                  // the lists are reused by the next pass
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const uint64_t r = base + 32 * j + lane;
            if (r < A.n_seqs) { A.min_cnt[r] = cj[j]; A.min_off[r] = gm[j]; }
            else if (r == A.n_seqs) A.min_off[r] = gm[j];
        }
    }
}

// k_item_scan: items per sequence = max(0, min_cnt - k + 1) (src/lib.rs:231-261: a window needs k minimizers), exclusive
// scan -> km_off[0 .. n_seqs].  One pass, decoupled look-back over tiles of RT * SPT sequences.
constexpr int SPT = 8;             // sequences per thread of the scan
struct K2SArgs {
    const uint32_t *min_cnt;
    uint64_t *km_off;
    uint64_t  n_seqs;
    uint32_t  k;
    uint64_t *status;              // per tile, zeroed
    uint32_t *ticket;
    uint32_t *err;
};
__global__ void __launch_bounds__(RT) k_item_scan(const __grid_constant__ K2SArgs A)
{
    if (*A.err & ERR_CAP) return;
    S2K_SHARED uint32_t wsum[RT / 32];
    S2K_SHARED unsigned long long s_excl;
    S2K_SHARED uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t n_tiles = (A.n_seqs + RT * SPT - 1) / (RT * SPT);
    if (n_tiles == 0) { if (blockIdx.x == 0 && tid == 0) A.km_off[0] = 0; return; }
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(A.ticket, 1u);
        __syncthreads();
        const uint32_t t = s_tile;
        if (t >= n_tiles) break;
        const uint64_t r0 = (uint64_t)t * (RT * SPT) + (uint64_t)tid * SPT;
        uint32_t items[SPT], sum = 0;
        if (r0 + SPT <= A.n_seqs) {
            const uint4 a = *reinterpret_cast<const uint4 *>(A.min_cnt + r0), b = *reinterpret_cast<const uint4 *>(A.min_cnt + r0 + 4);
            items[0] = a.x; items[1] = a.y; items[2] = a.z; items[3] = a.w; items[4] = b.x; items[5] = b.y; items[6] = b.z; items[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < SPT; ++j) items[j] = r0 + j < A.n_seqs ? A.min_cnt[r0 + j] : 0u;
        }
#pragma unroll
        for (int j = 0; j < SPT; ++j) { items[j] = items[j] >= A.k ? items[j] - A.k + 1 : 0u; sum += items[j]; }
        uint32_t incl = warp_incl_scan(sum, lane);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t pre = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < RT / 32; ++i) { const uint32_t sw = wsum[i]; if (i < warp) pre += sw; tot += sw; }
        const uint32_t excl_local = pre + incl - sum;
        if (warp == 0) {
            uint64_t excl = 0;
            if (t > 0) {
                if (lane == 0) st_relaxed(&A.status[t], FLAG_AGG | (uint64_t)tot);
                int64_t j = (int64_t)t - 1;
                for (;;) {
                    const int64_t idx = j - lane;
                    uint64_t sv = FLAG_INCL;
                    if (idx >= 0) {
                        uint32_t spins = 0;
                        while (((sv = ld_relaxed(&A.status[idx])) >> 62) == 0) {
                            if (++spins > SPIN_LIMIT) { atomicOr(A.err, ERR_SPIN); sv = FLAG_INCL; break; }
                            __nanosleep(40);
                        }
                    }
                    const uint32_t im = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
                    const int first = im ? (__ffs(im) - 1) : 32;
                    excl += warp_sum64(lane <= first ? (sv & VALMASK) : 0ull);
                    if (im) break;
                    j -= 32;
                }
            }
            if (lane == 0) {
                st_relaxed(&A.status[t], FLAG_INCL | (excl + tot));
                s_excl = excl;
                if (t == n_tiles - 1) A.km_off[A.n_seqs] = excl + tot;
            }
        }
        __syncthreads();
        uint64_t o = s_excl + excl_local;
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
            const uint64_t r = r0 + j;
            if (r < A.n_seqs) A.km_off[r] = o;
            o += items[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------ windows
struct K3Args {
    const uint4    *mins;
    const uint64_t *min_off, *km_off;
    const unsigned long long *n_min_p;   // total minimizers, on the device (the record cursor of k_minimizers)
    const uint32_t *err;                  // ERR_CAP set: the record store overflowed, the host reruns -- touch nothing
    uint32_t  k;
    uint64_t *hash;
    uint32_t *start, *end;
    uint8_t  *rev;
    const uint32_t *hash_hi;              // H = u64 flavour (k_windows only): high halves, MixHash<u64> = identity (src/lib.rs:171-177)
    uint32_t mix16;                       // H = u16 flavour (k_windows only): MixHash<u16>, src/lib.rs:142-155
};
__device__ __forceinline__ uint64_t mix32(uint32_t h)      // MixHash for u32, src/lib.rs:157-169
{
    uint64_t x = h;
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    return x;
}
__device__ __forceinline__ uint64_t mix16(uint32_t h)      // MixHash for u16, src/lib.rs:142-155 (wrapping multiplies)
{
    uint64_t x = h & 0xffffu;
    x ^= (x << 33) | (x >> 31); x *= 0xff51afd7ed558ccdull;
    x ^= (x << 33) | (x >> 31); x *= 0xc4ceb9fe1a85ec53ull;
    x ^= (x << 33) | (x >> 31);
    return x;
}
__device__ __forceinline__ uint64_t rol64(uint64_t x, uint32_t r)
{
    r &= 63u;
    return r ? ((x << r) | (x >> (64u - r))) : x;
}
__global__ void __launch_bounds__(256) k_windows(const __grid_constant__ K3Args A)
{
    if (*A.err & ERR_CAP) return;
    const uint64_t n_min = *A.n_min_p;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_min; g += stride) {
        const uint4 first = A.mins[g];
        const uint32_t rid = first.w;
        const uint64_t c = g - A.min_off[rid];             // window index inside the sequence == offset
        const uint64_t k0 = A.km_off[rid];
        if (c >= A.km_off[rid + 1] - k0) continue;         // fewer than k minimizers left (or tail rule)
        uint64_t f = 0, r = 0;
        uint32_t end = first.z;
        for (uint32_t tt = 0; tt < A.k; ++tt) {
            const uint4 mrec = tt ? A.mins[g + tt] : first;
            const uint64_t m = A.hash_hi ? ((uint64_t)A.hash_hi[g + tt] << 32 | mrec.x) : A.mix16 ? mix16(mrec.x) : mix32(mrec.x);
            f ^= rol64(m, A.k - 1 - tt);
            r ^= rol64(m, tt);
            end = mrec.z;
        }
        const uint64_t o = k0 + c;
        A.hash[o] = f < r ? f : r;
        A.start[o] = first.y;
        A.end[o] = end;
        A.rev[o] = r < f;
    }
}

// Warp-cooperative form for small k (the usual 3..10): every lane loads ONE record and mixes its hash once; the k
// values of a window come from the k-1 lanes above by shuffle, with compile-time rotation amounts.  A warp pass covers
// 32 consecutive minimizers and emits the 32-(k-1) windows that start in its first lanes.  (k_windows recomputes
// mix() k times per item and rotates by run-time amounts: ~190 instructions per item against ~90 here.)
template <int R> __device__ __forceinline__ uint64_t rol64c(uint64_t x)
{
    constexpr int r = R & 63;
    return r ? ((x << r) | (x >> (64 - r))) : x;
}
template <int K, int T> struct WindowFold {
    static __device__ __forceinline__ void run(uint64_t m, uint64_t &f, uint64_t &r)
    {
        const uint64_t mt = T ? __shfl_down_sync(0xffffffffu, m, T) : m;
        f ^= rol64c<K - 1 - T>(mt);
        r ^= rol64c<T>(mt);
        WindowFold<K, T + 1>::run(m, f, r);
    }
};
template <int K> struct WindowFold<K, K> {
    static __device__ __forceinline__ void run(uint64_t, uint64_t &, uint64_t &) {}
};
template <int K>
__global__ void __launch_bounds__(256) k_windows_w(const __grid_constant__ K3Args A)
{
    constexpr uint32_t OUT = 32 - (K - 1);                  // windows emitted per warp pass
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    if (*A.err & ERR_CAP) return;
    const uint64_t n_min = *A.n_min_p;
    const uint64_t n_pass = (n_min + OUT - 1) / OUT;
    for (uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n_pass; p += nwarps) {
        const uint64_t g = p * OUT + lane;
        uint4 rec = make_uint4(0u, 0u, 0u, 0u);
        if (g < n_min) rec = A.mins[g];
        uint64_t f = 0, r = 0;
        WindowFold<K, 0>::run(mix32(rec.x), f, r);
        const uint32_t end = K > 1 ? __shfl_down_sync(0xffffffffu, rec.z, K - 1) : rec.z;
        if (lane < OUT && g < n_min) {
            const uint32_t rid = rec.w;
            const uint64_t c = g - A.min_off[rid];          // window index inside the sequence == offset
            const uint64_t k0 = A.km_off[rid];
            if (c < A.km_off[rid + 1] - k0) {               // else: fewer than k minimizers left (or tail rule)
                const uint64_t o = k0 + c;
                A.hash[o] = f < r ? f : r;
                A.start[o] = rec.y;
                A.end[o] = end;
                A.rev[o] = r < f;
            }
        }
    }
}
// The same window stage reading the records where k_minimizers left them (no ordered copy, S2K_NO_MINIMIZER_STREAM):
// one warp per tile; the last K-1 windows of a tile look into the following tiles.
struct K3TArgs {
    K3Args W;                       // W.mins = the unordered record store
    const uint4 *tile_info;
    const ulonglong2 *tile_src;
    uint32_t n_tiles;
    uint32_t ridtest;               // 1: test that a window's first and last record share the sequence before the per-
                                    // sequence loads (short reads; on long reads the extra shuffle costs more than it spares:
                                    // config 3 window stage 3.73 -> 3.47 ms, config 2 1.265 -> 1.280 ms)
};
// Six CTAs per SM (40 registers): with no minimum ptxas settles on 40 registers AND spills 12 bytes; 8 CTAs (32 registers)
// and one-wave grids measured no better (profiles/r2_ab_log.txt: window stage 1.297 -> 1.265 ms on config 2).
#ifndef S2K_WIN_MINB
#define S2K_WIN_MINB 6
#endif
#define S2K_WIN_BOUNDS __launch_bounds__(256, S2K_WIN_MINB)
template <int K>
__global__ void S2K_WIN_BOUNDS k_windows_t(const __grid_constant__ K3TArgs A)
{
    constexpr uint32_t OUT = 32 - (K - 1);
    if (*A.W.err & ERR_CAP) return;
    const uint64_t n_min = *A.W.n_min_p;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < A.n_tiles; t += nwarps) {
        const uint32_t h = A.tile_info[t].x;
        if (h == 0) continue;
        const ulonglong2 ts = A.tile_src[t];
#if !defined(S2K_EMU) && !defined(S2K_WIN_NO_PREFETCH)
        // The stage is latency-bound (one 512-byte load in flight per warp, long-scoreboard stalls 14 of 20 cycles per
        // issue): pull the tile's records (~4 KB; the first 256 of them) into L2, one prefetch per 32-byte sector, so
        // that the passes after the first wait for L2 instead of DRAM.  No register is held for it.
#pragma unroll
        for (uint32_t m = 0; m < 4; ++m) {
            const uint32_t idx = 2u * lane + 64u * m;
            if (idx < h) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.W.mins + ts.y + idx));
        }
#endif
        for (uint32_t p0 = 0; p0 < h; p0 += OUT) {
            const uint32_t j = p0 + lane;
            uint4 rec = make_uint4(0u, 0u, 0u, 0u);
            if (j < h) {
                rec = A.W.mins[ts.y + j];
            } else if (j - h < (uint32_t)(K - 1) && ts.x + j < n_min) {
                const uint64_t m = ts.x + j;
                const uint32_t t2 = tile_of_record(A.tile_src, A.n_tiles, m, t + 1);
                const ulonglong2 ts2 = A.tile_src[t2];
                rec = A.W.mins[ts2.y + (m - ts2.x)];
            }
            uint64_t f = 0, r = 0;
            WindowFold<K, 0>::run(mix32(rec.x), f, r);
            const uint32_t end = K > 1 ? __shfl_down_sync(0xffffffffu, rec.z, K - 1) : rec.z;
            // a window lies inside one sequence: on short reads (config 3: 1.65 minimizers per read) this spares 98 % of
            // the lanes the three dependent, scattered loads below (A.ridtest; the exact test follows either way)
            const uint32_t rid_last = (K > 1 && A.ridtest) ? __shfl_down_sync(0xffffffffu, rec.w, K - 1) : rec.w;   // uniform
            if (lane < OUT && j < h && rid_last == rec.w) {
                const uint32_t rid = rec.w;
                const uint64_t c = ts.x + j - A.W.min_off[rid];
                const uint64_t k0 = A.W.km_off[rid];
                if (c < A.W.km_off[rid + 1] - k0) {
                    const uint64_t o = k0 + c;
                    A.W.hash[o] = f < r ? f : r;
                    A.W.start[o] = rec.y;
                    A.W.end[o] = end;
                    A.W.rev[o] = r < f;
                }
            }
        }
    }
}
constexpr int KW_MAX = 12;          // largest k with a k_windows_w instantiation

// ------------------------------------------------------------------------------------------------ RLE (encode_rle_simd)
// src/hpc.rs:44-147 over a batch: kept bytes + run starts, ordered.  Same keep rule and tile order machinery,
// no hashing.  One thread handles 32 bases; ordered by a decoupled look-back on kept counts.
struct K4Args {
    const uint8_t  *bases;
    const uint64_t *seq_off;
    const uint32_t *tile_lb;
    uint64_t n_seqs, n_bases;
    uint32_t n_tiles;
    uint32_t scalar_rule;          // S2K_RLE_SCALAR_RULE: only runs of "ACTGactgNn" collapse (src/hpc.rs:14)
    uint64_t *status; uint32_t *ticket; uint32_t *err;
    uint8_t  *hpc; uint32_t *pos; uint64_t *hpc_off;
};
constexpr int RLE_TILE = NT * 32;
// bit b of the result: byte b of the 32 bytes in w[0..7] is one of "ACTGactgNn" (upper-cased: A C G T N)
__device__ __forceinline__ uint32_t rle_collapsible(const uint32_t (&w)[8])
{
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = ((w[i] >> (8 * j)) & 0xffu) & 0xdfu;            // fold the case bit
            const bool in = c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N';
            m |= (in ? 1u : 0u) << (4 * i + j);
        }
    }
    return m;
}
__global__ void __launch_bounds__(NT) k_rle(const __grid_constant__ K4Args A)
{
    S2K_SHARED uint32_t startw[NT];
    S2K_SHARED uint32_t keepw[NT + 1], qoff[NT + 1];
    S2K_SHARED uint32_t wsum[NT / 32];
    S2K_SHARED uint32_t s_tile;
    S2K_SHARED unsigned long long s_excl;
    // the tile's output, staged so that it leaves in whole sectors: a thread's ~24 kept bases are consecutive in the
    // output but 24 bytes / 96 bytes away from its neighbours', and byte / word stores straight to global memory made
    // every store instruction touch 32 sectors (148 Gbp/s; profiles/r2_rle.txt)
    S2K_SHARED uint32_t s_pos[RLE_TILE];
    S2K_SHARED uint32_t s_hpc[RLE_TILE / 4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(A.ticket, 1u);
        __syncthreads();
        const uint32_t t = s_tile;
        if (t >= A.n_tiles) break;
        const uint64_t T0 = (uint64_t)t * RLE_TILE, T1 = min(T0 + (uint64_t)RLE_TILE, A.n_bases);
        const bool last_tile = T1 == A.n_bases;
        const uint32_t lb = A.tile_lb[t], ub = last_tile ? (uint32_t)(A.n_seqs + 1) : A.tile_lb[t + 1];
        startw[tid] = 0;
        // this thread's 32 bases: two 16-byte loads (the batch is 16-byte aligned, RLE_TILE a multiple of 32)
        const uint64_t g0 = T0 + 32ull * tid;
        uint32_t w[8];
        if (g0 + 32 <= A.n_bases) {
            const uint4 x = __ldg(reinterpret_cast<const uint4 *>(A.bases + g0)), y = __ldg(reinterpret_cast<const uint4 *>(A.bases + g0) + 1);
            w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w; w[4] = y.x; w[5] = y.y; w[6] = y.z; w[7] = y.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t x = 0;
                for (int j = 0; j < 4; ++j) { const uint64_t g = g0 + 4 * i + j; if (g < A.n_bases) x |= (uint32_t)A.bases[g] << (8 * j); }
                w[i] = x;
            }
        }
        __syncthreads();
        for (uint32_t i = lb + tid; i < ub; i += NT) {
            const uint64_t so = A.seq_off[i];
            if (so < T1) { const uint32_t x = (uint32_t)(so - T0); atomicOr(&startw[x >> 5], 1u << (x & 31)); }
        }
        __syncthreads();
        // keep bit = byte differs from the byte before it (word-wise, as in k_minimizers), or starts a sequence
        uint32_t prevw = __shfl_up_sync(0xffffffffu, w[7], 1);
        if (lane == 0) prevw = ((g0 > 0 && g0 <= A.n_bases) ? (uint32_t)A.bases[g0 - 1] : 0u) << 24;
        uint32_t keep = 0;
#pragma unroll
        for (int i = 7; i >= 0; --i) {
            const uint32_t x = w[i] ^ __byte_perm(i ? w[i - 1] : prevw, w[i], 0x6543u);
            const uint32_t nz = (x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;
            keep = __funnelshift_l(nz * 0x00204081u, keep, 4);
        }
        if (A.scalar_rule) keep |= ~rle_collapsible(w);      // a repeated byte outside "ACTGactgNn" is kept
        keep |= startw[tid];
        if (g0 + 32 > T1) keep &= g0 < T1 ? lowmask((uint32_t)(T1 - g0)) : 0u;
        uint32_t incl = warp_incl_scan(__popc(keep), lane);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t pre = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < NT / 32; ++i) { const uint32_t s = wsum[i]; if (i < warp) pre += s; tot += s; }
        const uint32_t q = pre + incl - __popc(keep);
        keepw[tid] = keep; qoff[tid] = q;
        if (tid == NT - 1) { keepw[NT] = 0; qoff[NT] = tot; }
        if (warp == 0) {
            uint64_t excl = 0;
            if (t > 0) {
                if (lane == 0) st_relaxed(&A.status[t], FLAG_AGG | (uint64_t)tot);
                int64_t j = (int64_t)t - 1;
                for (;;) {
                    const int64_t idx = j - lane;
                    uint64_t s = FLAG_INCL;
                    if (idx >= 0) {
                        uint32_t spins = 0;
                        while (((s = ld_relaxed(&A.status[idx])) >> 62) == 0) {
                            if (++spins > SPIN_LIMIT) { atomicOr(A.err, ERR_SPIN); s = FLAG_INCL; break; }
                            __nanosleep(40);
                        }
                    }
                    const uint32_t im = __ballot_sync(0xffffffffu, (s >> 62) == 2);
                    const int first = im ? (__ffs(im) - 1) : 32;
                    excl += warp_sum64(lane <= first ? (s & VALMASK) : 0ull);
                    if (im) break;
                    j -= 32;
                }
            }
            if (lane == 0) { st_relaxed(&A.status[t], FLAG_INCL | (excl + tot)); s_excl = excl; }
        }
        __syncthreads();
        const uint64_t base = s_excl;
        // sequence of each kept base: walk the starts of this tile
        if (keep) {
            uint32_t lo = lb, hi = ub;                     // first i with seq_off[i] > g0
            while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (A.seq_off[mid] <= g0) lo = mid + 1; else hi = mid; }
            uint32_t rid = lo - 1;
            uint64_t so = A.seq_off[rid], nx = A.seq_off[rid + 1];
            uint32_t o = q;
            uint8_t *const sb = reinterpret_cast<uint8_t *>(s_hpc);
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                if ((keep >> b) & 1u) {
                    const uint64_t g = g0 + b;
                    while (g >= nx) { ++rid; so = A.seq_off[rid]; nx = A.seq_off[rid + 1]; }
                    sb[o] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
                    s_pos[o] = (uint32_t)(g - so);
                    ++o;
                }
            }
        }
        __syncthreads();
        for (uint32_t i = tid; i < tot; i += NT) A.pos[base + i] = s_pos[i];
        {   // kept bytes: the destination starts at any byte -- bytes up to a word boundary, whole words, bytes again
            const uint8_t *const sb = reinterpret_cast<const uint8_t *>(s_hpc);
            const uint32_t head = min(tot, (uint32_t)((4u - (uint32_t)(base & 3ull)) & 3u));
            if ((uint32_t)tid < head) A.hpc[base + tid] = sb[tid];
            const uint32_t nw = (tot - head) >> 2;
            uint32_t *const dw = reinterpret_cast<uint32_t *>(A.hpc + base + head);
            for (uint32_t i = tid; i < nw; i += NT) {
                const uint32_t a = head + 4u * i, sh = 8u * (a & 3u);      // shared side: unaligned by `head`
                dw[i] = __funnelshift_r(s_hpc[a >> 2], s_hpc[(a >> 2) + (sh ? 1u : 0u)], sh);
            }
            const uint32_t tail0 = head + 4u * nw;
            if ((uint32_t)tid < tot - tail0) A.hpc[base + tail0 + tid] = sb[tail0 + tid];
        }
        for (uint32_t i = lb + tid; i < ub; i += NT) {
            const uint64_t s = A.seq_off[i];
            const uint32_t x = (uint32_t)(s - T0);
            A.hpc_off[i] = base + qoff[x >> 5] + __popc(keepw[x >> 5] & lowmask(x & 31));
        }
    }
}

// ------------------------------------------------------------------------------------------------ slab stitching
// s2k_run streams a large host batch through the device in slabs; per-slab prefixes and sequence indices are made
// global on the device before they are copied back.
__global__ void k_add_u64(uint64_t *__restrict__ a, uint64_t n, uint64_t add)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] += add;
}
__global__ void k_sub_first(uint64_t *__restrict__ a, uint64_t n)     // a[i] -= a[0], a[0] last
{
    const uint64_t base = a[0];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = 1 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] -= base;
}
__global__ void k_zero_first(uint64_t *a) { a[0] = 0; }
__global__ void k_add_seq(uint4 *__restrict__ mins, uint64_t n, uint32_t add)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) mins[i].w += add;
}

// ---- one long sequence travelling through s2k_run in pieces (run_pipelined): bookkeeping kernels.
// A piece is the owned range [b0, b1) of the sequence plus a right overlap, processed as a sequence of its own.
// k_piece_cut: i1 = minimizers whose start lies in the owned range (all of them for the last piece), n_keep =
// minimizers left after the AVX-512 tail rule (last piece only: those ending at or beyond e16 are dropped).
__global__ void k_piece_cut(const uint4 *__restrict__ mins, uint64_t n_min, uint32_t own_len, uint32_t e16,
                            unsigned long long *__restrict__ out)
{
    uint64_t lo = 0, hi = n_min;                 // first minimizer with start >= own_len
    while (lo < hi) { const uint64_t mid = lo + ((hi - lo) >> 1); if (mins[mid].y < own_len) lo = mid + 1; else hi = mid; }
    out[0] = lo;
    lo = 0; hi = n_min;                          // first minimizer with end >= e16
    while (lo < hi) { const uint64_t mid = lo + ((hi - lo) >> 1); if (mins[mid].z < e16) lo = mid + 1; else hi = mid; }
    out[1] = lo;
}
// kept (HPC) bases among b[0, n): b[i] != b[i-1], b[0] counts (pieces are cut at run boundaries)
__global__ void __launch_bounds__(256) k_count_kept(const uint8_t *__restrict__ b, uint64_t n, unsigned long long *__restrict__ acc)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) c += (i == 0 || b[i] != b[i - 1]) ? 1u : 0u;
    c = warp_sum64(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(acc, c);
}
__global__ void k_piece_shift_items(uint32_t *__restrict__ start, uint32_t *__restrict__ end, uint64_t n, uint32_t add)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { start[i] += add; end[i] += add; }
}
__global__ void k_piece_shift_mins(uint4 *__restrict__ mins, uint64_t n, uint32_t add, uint32_t seq)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint4 m = mins[i];
        m.y += add; m.z += add; m.w = seq;
        mins[i] = m;
    }
}

// 2-bit transport (s2k_run): the host packs ACGT-only slabs 4 bases per byte (base i in bits 2*(i%4), code = (b>>1)&3:
// A0 C1 T2 G3) to quarter the PCIe traffic; this kernel restores the ASCII bytes the rest of the path works on.
// `shift` (0..3): the first base wanted is base `shift` of packed[0] (packed-input batches are cut at arbitrary bases;
// reads packed[nvec] then, which the callers keep inside their buffers).
__global__ void __launch_bounds__(256) k_unpack2(const uint32_t *__restrict__ packed, uint64_t n_bases, uint8_t *__restrict__ out,
                                                 uint32_t shift)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nvec = (n_bases + 15) / 16;                    // 16 bases = one packed word = one 16-byte store
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const uint32_t pw = shift ? __funnelshift_r(packed[v], packed[v + 1], 2 * shift) : packed[v];
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) x |= ((0x47544341u >> (8 * ((pw >> (8 * i + 2 * j)) & 3u))) & 0xffu) << (8 * j);   // "ACTG"
            w[i] = x;
        }
        *reinterpret_cast<uint4 *>(out + 16 * v) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ------------------------------------------------------------------------------------------------ synthetic reads
// Workload generator of SURVEY.md 8(d): 32 bases per 64-bit word, word(j) = splitmix64 finalizer of
// seed + (j+1)*0x9E3779B97F4A7C15, base(i) = "ACGT"[(word(i>>5) >> 2*(i&31)) & 3].  Same data on host
// (oracle/s2k_oracle.c: s2k_oracle_synth) and device without a transfer.
__device__ __forceinline__ uint64_t synth_word(uint64_t seed, uint64_t j)
{
    uint64_t z = seed + (j + 1) * 0x9E3779B97F4A7C15ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
__global__ void __launch_bounds__(256) k_synth(uint64_t seed, uint64_t first, uint64_t count, uint8_t *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nvec = (count + 15) / 16;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const uint64_t t0 = v * 16;
        uint32_t w[4] = {0, 0, 0, 0};
        uint64_t cur_j = ~0ull, word = 0;
        for (int b = 0; b < 16; ++b) {
            const uint64_t i = first + t0 + b;
            if ((i >> 5) != cur_j) { cur_j = i >> 5; word = synth_word(seed, cur_j); }
            const uint32_t c = (uint32_t)(word >> (2 * (i & 31))) & 3u;
            w[b >> 2] |= ((0x54474341u >> (8 * c)) & 0xffu) << (8 * (b & 3));   // "ACGT"
        }
        if (t0 + 16 <= count && ((reinterpret_cast<uintptr_t>(out) + t0) & 15u) == 0) {
            *reinterpret_cast<uint4 *>(out + t0) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (int b = 0; b < 16 && t0 + b < count; ++b) out[t0 + b] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
        }
    }
}

} // namespace s2k
