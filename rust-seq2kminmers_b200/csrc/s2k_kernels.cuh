// s2k_kernels.cuh -- sm_100a kernels of the sequence -> k-min-mer path.
//
// Pipeline per launch sequence (all on one stream, no host round trip in between):
//   k_tile_bounds    : per tile, index of the first sequence starting at/after the tile      (tiny)
//   k_minimizers     : fused HPC keep-mask + in-smem compaction + canonical ntHash (32/31 bit) of every
//                      l-mer in HPC space + density threshold + ordered append of (hash,start,end,seq)
//                      records, tile order kept by a decoupled look-back                     (the hot kernel)
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
#include <cstdint>
#ifdef S2K_EMU            // tests/emu: the same sources compiled by g++ onto host threads (test tier only)
#include "cuda_emu.h"
#define S2K_DYN_SMEM(name) uint8_t *name = emu::blk->smem
#define S2K_SHARED static
#else
#include <cuda_runtime.h>
#define S2K_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#define S2K_SHARED __shared__
#endif

namespace s2k {

// ------------------------------------------------------------------------------------------------ geometry
constexpr int NT    = 256;          // threads per CTA
constexpr int CH    = 32;           // owner positions per thread in the hash phase
constexpr int WIN   = NT * CH;      // raw bases staged per tile (left halo + tile)
constexpr int XB    = 256;          // capacity of the left context, in kept (HPC) bases
constexpr int XC    = XB / CH;      // ... in columns of the transposed code array
constexpr int NTP   = NT + XC + 4;  // column pitch of the transposed code array (268 = 4*67: conflict-free rows)
constexpr int NWORD = WIN / 32;     // 32-base chunks per window == NT
constexpr int OOW   = NT + XC + 1;  // words of the owner-space flag bitmaps
constexpr int ZCLS  = 4;            // base class whose forward and reverse seeds are both 0

constexpr uint64_t FLAG_AGG  = 1ull << 62;
constexpr uint64_t FLAG_INCL = 2ull << 62;
constexpr uint64_t VALMASK   = (1ull << 62) - 1;
constexpr uint32_t SPIN_LIMIT = 1u << 24;

constexpr uint32_t ERR_SPIN = 1u, ERR_CAP = 2u;

struct K1Args {
    const uint8_t  *bases;
    const uint64_t *seq_off;     // n_seqs + 1
    const uint32_t *tile_lb;     // n_tiles + 1: first i with seq_off[i] >= tile start
    uint64_t *status;            // n_tiles, zeroed before launch
    uint32_t *ticket;            // zeroed before launch
    const uint64_t *carry_in;    // [2] minimizers / kept bases before this slab
    uint64_t *carry_out;         // [2]
    uint4    *min_out;           // minimizer records (hash, start, end, seq)
    uint64_t  min_cap;
    uint64_t *min_off;           // n_seqs + 1
    uint64_t *hpc_off;           // n_seqs + 1 or null
    uint32_t *err;
    uint64_t  n_seqs, n_bases, slab_begin, slab_end;
    uint32_t  n_tiles, tile, halo;
    uint32_t  l, d, need, thr;
    uint8_t   cls_lut[256];      // raw byte -> base class (0..5)
    uint2     xy[64];            // [out*8+in] -> (rol(h[out],l)^h[in], ror(rc[out],1)^rol(rc[in],l-1))
};

struct Smem {
    uint8_t  raw[16 + WIN];
    uint8_t  code[32 * NTP];
    uint32_t hh[WIN];
    uint32_t keepw[NWORD + 1];
    uint32_t qoff[NWORD + 1];
    uint32_t startw[NWORD];
    uint32_t shortw[NWORD];
    uint32_t f1[OOW + 1];
    uint32_t f2[OOW + 1];
    uint32_t hitw[NT + 1];
    uint32_t hitpre[NT + 1];
    uint32_t ctxpos[XB];
    uint2    xy[64];
    uint8_t  lut[256];
    uint32_t wsum[8];
    uint32_t tile_id, hk, min_ex, kept_ex;
    unsigned long long s0;
};

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
// Uses S.wsum; callers must not touch wsum until the trailing barrier inside has been passed.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *wsum, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(v, lane);
    __syncthreads();                       // previous users of wsum are done
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t pre = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) {
        uint32_t s = wsum[i];
        if (i < warp) pre += s;
        tot += s;
    }
    total = tot;
    return pre + incl - v;
}
__device__ __forceinline__ uint32_t lowmask(uint32_t n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }
__device__ __forceinline__ int nth_set_bit(uint32_t w, int r)
{
    for (int t = 0; t < r; ++t) w &= w - 1;
    return __ffs(w) - 1;
}
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
__device__ __forceinline__ int code_idx(int ee) { return (ee & 31) * NTP + (ee >> 5); }

// Original-space position (global index into `bases`) of the kept base with window index q
// (q < 0: context gathered by the walk-back).
__device__ __forceinline__ int64_t pos_of(const Smem &S, int64_t W0, int q)
{
    if (q < 0) return W0 - (int64_t)S.ctxpos[-1 - q];
    int lo = 0, hi = NT;                    // qoff[lo] <= q < qoff[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if ((int)S.qoff[mid] <= q) lo = mid; else hi = mid;
    }
    return W0 + 32 * lo + nth_set_bit(S.keepw[lo], q - (int)S.qoff[lo]);
}

// ------------------------------------------------------------------------------------------------ tile bounds
__global__ void k_tile_bounds(const uint64_t *__restrict__ seq_off, uint64_t n_seqs, uint64_t slab_begin,
                              uint64_t slab_end, uint32_t tile, uint32_t n_tiles, uint32_t *__restrict__ tile_lb)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    uint64_t pos = slab_begin + (uint64_t)t * tile;
    if (pos > slab_end) pos = slab_end;
    uint64_t lo = 0, hi = n_seqs + 1;       // first i in [0, n_seqs] with seq_off[i] >= pos
    while (lo < hi) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (seq_off[mid] < pos) lo = mid + 1; else hi = mid;
    }
    tile_lb[t] = (uint32_t)lo;
}

// ------------------------------------------------------------------------------------------------ minimizers
template <bool HPC, bool W31>
__global__ void __launch_bounds__(NT, 3) k_minimizers(const __grid_constant__ K1Args A)
{
    S2K_DYN_SMEM(smem_raw);
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = (int)A.l, d = (int)A.d;

    for (int i = tid; i < 256; i += NT) S.lut[i] = A.cls_lut[i];
    if (tid < 64) S.xy[tid] = A.xy[tid];
    for (int i = tid; i < 32 * NTP; i += NT) S.code[i] = ZCLS;

    for (;;) {
        __syncthreads();                                   // everyone is done with the previous tile
        if (tid == 0) S.tile_id = atomicAdd(A.ticket, 1u);
        __syncthreads();
        const uint32_t t = S.tile_id;
        if (t >= A.n_tiles) break;

        const int64_t T0 = (int64_t)(A.slab_begin + (uint64_t)t * A.tile);
        const int64_t T1 = min(T0 + (int64_t)A.tile, (int64_t)A.slab_end);
        const int64_t W0 = T0 - (int64_t)A.halo;
        const bool last_tile = (uint64_t)T1 == A.n_bases;
        const uint32_t lb = A.tile_lb[t];
        const uint32_t ub = last_tile ? (uint32_t)(A.n_seqs + 1) : A.tile_lb[t + 1];

        // ---- S1: stage raw bytes [W0-16, W0+WIN); bytes outside [0, n_bases) read as 0
        for (int v = tid; v < (WIN + 16) / 16; v += NT) {
            const int64_t g = W0 - 16 + (int64_t)v * 16;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (g >= 0 && g + 16 <= (int64_t)A.n_bases) {
                val = __ldg(reinterpret_cast<const uint4 *>(A.bases + g));
            } else if (g + 16 > 0 && g < (int64_t)A.n_bases) {
                uint32_t w[4] = {0, 0, 0, 0};
                for (int j = 0; j < 16; ++j) {
                    const int64_t gg = g + j;
                    if (gg >= 0 && gg < (int64_t)A.n_bases) w[j >> 2] |= (uint32_t)A.bases[gg] << (8 * (j & 3));
                }
                val = make_uint4(w[0], w[1], w[2], w[3]);
            }
            *reinterpret_cast<uint4 *>(S.raw + v * 16) = val;
        }
        for (int i = tid; i < NWORD; i += NT) { S.startw[i] = 0; S.shortw[i] = 0; }
        for (int i = tid; i < OOW + 1; i += NT) { S.f1[i] = 0; S.f2[i] = 0; }
        __syncthreads();

        // ---- S2: sequence starts inside the tile (and the start of the sequence containing T0, if staged)
        for (uint32_t i = lb + tid; i < ub; i += NT) {
            const uint64_t so = A.seq_off[i];
            if (so < (uint64_t)T1) {
                const uint32_t x = (uint32_t)((int64_t)so - W0);
                atomicOr(&S.startw[x >> 5], 1u << (x & 31));
                const uint64_t len = A.seq_off[i + 1] - so;
                if (len > 0 && len <= (uint64_t)l) atomicOr(&S.shortw[x >> 5], 1u << (x & 31));
            }
        }
        if (tid == 0) {
            const uint64_t so_lb = A.seq_off[lb];
            unsigned long long s0 = (unsigned long long)T0;
            if (so_lb != (uint64_t)T0) {                   // the sequence containing T0 started earlier
                s0 = A.seq_off[lb - 1];
                if ((int64_t)s0 >= W0) {
                    const uint32_t x = (uint32_t)((int64_t)s0 - W0);
                    atomicOr(&S.startw[x >> 5], 1u << (x & 31));
                    if (so_lb - s0 <= (uint64_t)l) atomicOr(&S.shortw[x >> 5], 1u << (x & 31));
                }
            }
            S.s0 = s0;
        }
        __syncthreads();

        // ---- S3: keep mask of this thread's 32 raw bases, block scan of kept counts
        uint32_t w[8];
        {
            const uint4 a = *reinterpret_cast<const uint4 *>(S.raw + 16 + 32 * tid);
            const uint4 b = *reinterpret_cast<const uint4 *>(S.raw + 32 + 32 * tid);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        }
        uint32_t keep;
        if (HPC) {
            uint32_t prevb = S.raw[16 + 32 * tid - 1];
            keep = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t sh = (w[i] << 8) | prevb;
                prevb = w[i] >> 24;
                const uint32_t neq = __vcmpne4(w[i], sh);
                keep |= (((neq & 0x08040201u) * 0x01010101u) >> 24) << (4 * i);
            }
            keep |= S.startw[tid];
        } else {
            keep = 0xffffffffu;
        }
        {
            const int64_t g0 = W0 + 32 * tid;
            uint32_t vmask = 0xffffffffu;
            if (g0 < 0) vmask = (g0 <= -32) ? 0u : (0xffffffffu << (int)(-g0));
            const int64_t rem = T1 - g0;
            if (rem <= 0) vmask = 0u; else if (rem < 32) vmask &= (1u << (int)rem) - 1u;
            keep &= vmask;
        }
        uint32_t wk;
        const uint32_t q = block_excl_scan(__popc(keep), S.wsum, wk);
        S.keepw[tid] = keep;
        S.qoff[tid] = q;
        if (tid == NT - 1) { S.qoff[NT] = wk; S.keepw[NT] = 0; }
        if (tid == (int)(A.halo >> 5)) S.hk = q;
        __syncthreads();
        const uint32_t hk = S.hk;

        // ---- S4: compaction of base classes into the transposed code array, start flags into owner space
        {
            int ee = (int)q - (int)hk + d + 256;           // code index of this thread's first kept base
            const uint32_t sw = S.startw[tid] & keep, sh2 = S.shortw[tid];
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                if ((keep >> b) & 1u) {
                    if (ee >= 0) S.code[code_idx(ee)] = S.lut[(w[b >> 2] >> (8 * (b & 3))) & 0xffu];
                    if ((sw >> b) & 1u) {
                        const int oo = ee - d;
                        if (oo >= 0) {
                            atomicOr(&S.f1[oo >> 5], 1u << (oo & 31));
                            if ((sh2 >> b) & 1u) atomicOr(&S.f2[oo >> 5], 1u << (oo & 31));
                        }
                    }
                    ++ee;
                }
            }
        }
        // ---- S4b: not enough context in the halo -> walk back through the sequence (rare: long homopolymers)
        const bool need_walk = HPC && (int64_t)S.s0 < W0 && hk < A.need;
        if (need_walk && warp == 0) {
            uint32_t remaining = A.need - hk, taken = 0;
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
                    const int ee = -1 - (int)slot - (int)hk + d + 256;
                    if (ee >= 0) S.code[code_idx(ee)] = S.lut[b];
                    S.ctxpos[slot] = (uint32_t)(W0 - g);
                    if (g == s0) { const int oo = ee - d; if (oo >= 0) atomicOr(&S.f1[oo >> 5], 1u << (oo & 31)); }
                }
                const uint32_t c = min((uint32_t)__popc(m), remaining);
                taken += c; remaining -= c; hi = lo;
            }
        }
        __syncthreads();

        // ---- S5: rolling canonical ntHash over this thread's 32 owner positions
        const uint32_t n_own = wk - hk;                   // kept bases in [T0, T1): one l-mer is owned by each
        const int ubase = 32 * tid;
        uint32_t mask = 0;
        if ((uint32_t)ubase < n_own) {
            const int n_u = min(32, (int)n_own - ubase);
            // owners invalidated by sequence starts: every start f kills owners [f, f+l-2+d] (+1 if len<=l)
            uint32_t invalid = 0;
            {
                const int L1 = l - 1 + d;
                const int w_hi = tid + XC, w_lo = (ubase + 256 - L1 - 1) >> 5;
                for (int wi = w_lo; wi <= w_hi; ++wi) {
                    uint32_t fw = S.f1[wi];
                    if (fw) {
                        const uint32_t sw2 = S.f2[wi];
                        while (fw) {
                            const int b = __ffs(fw) - 1;
                            fw &= fw - 1;
                            int lo = wi * 32 + b - (ubase + 256);
                            int hi = lo + L1 + (int)((sw2 >> b) & 1u);
                            lo = max(lo, 0); hi = min(hi, 32);
                            if (hi > lo) invalid |= lowmask(hi - lo) << lo;
                        }
                    }
                }
            }
            uint32_t fh = 0, rh = 0;
            {
                int ee = ubase + 256 - l + 1;              // warm-up: first l-1 bases of owner 0's l-mer
                for (int j = 0; j < l - 1; ++j, ++ee) {
                    const uint32_t in = S.code[code_idx(ee)];
                    const uint2 tt = S.xy[ZCLS * 8 + in];
                    fh = rol1<W31>(fh) ^ tt.x;
                    rh = ror1<W31>(rh) ^ tt.y;
                }
            }
            const uint8_t *cin = S.code + tid + XC;
            const int eo = ubase + 256 - l;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint32_t in = cin[i * NTP];
                uint32_t out = ZCLS;
                if (i > 0) out = S.code[code_idx(eo + i)];
                const uint2 tt = S.xy[out * 8 + in];
                fh = rol1<W31>(fh) ^ tt.x;
                rh = ror1<W31>(rh) ^ tt.y;
                const uint32_t h = min(fh, rh);
                if (h <= A.thr && i < n_u) { mask |= 1u << i; S.hh[ubase + i] = h; }
            }
            mask &= ~invalid;
        }

        // ---- S6: ordered placement: block scan of hit counts, decoupled look-back across tiles
        uint32_t tile_min;
        const uint32_t hpre = block_excl_scan(__popc(mask), S.wsum, tile_min);
        S.hitw[tid] = mask;
        S.hitpre[tid] = hpre;
        if (tid == NT - 1) { S.hitpre[NT] = tile_min; S.hitw[NT] = 0; }
        if (warp == 0) {
            const uint64_t agg = ((uint64_t)n_own << 31) | (uint64_t)tile_min;
            uint64_t excl = 0;
            if (t > 0) {
                if (lane == 0) st_relaxed(&A.status[t], FLAG_AGG | agg);
                int64_t j = (int64_t)t - 1;
                for (;;) {
                    const int64_t idx = j - lane;
                    uint64_t s = FLAG_INCL;                // tiles before the slab: inclusive prefix 0
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
            if (lane == 0) {
                st_relaxed(&A.status[t], FLAG_INCL | (excl + agg));
                S.min_ex = (uint32_t)(excl & 0x7fffffffu);
                S.kept_ex = (uint32_t)((excl >> 31) & 0x7fffffffu);
                if (t == A.n_tiles - 1) {
                    A.carry_out[0] = A.carry_in[0] + (excl & 0x7fffffffu) + tile_min;
                    A.carry_out[1] = A.carry_in[1] + ((excl >> 31) & 0x7fffffffu) + n_own;
                }
            }
        }
        __syncthreads();
        const uint64_t min_base = A.carry_in[0] + S.min_ex;
        const uint64_t kept_base = A.carry_in[1] + S.kept_ex;

        // ---- S7: emit minimizer records in order
        {
            uint32_t m = mask, kk = 0;
            while (m) {
                const int i = __ffs(m) - 1;
                m &= m - 1;
                const int qo = (int)hk + ubase + i;        // window index of the owner base
                const uint32_t h = S.hh[ubase + i];
                const int64_t g_own = pos_of(S, W0, qo);
                const int64_t g_start = pos_of(S, W0, qo - (l - 1 + d));
                uint32_t lo = lb, hi = ub;                 // first i in [lb,ub) with seq_off[i] > g_own
                while (lo < hi) {
                    const uint32_t mid = lo + ((hi - lo) >> 1);
                    if (A.seq_off[mid] <= (uint64_t)g_own) lo = mid + 1; else hi = mid;
                }
                const uint32_t rid = lo - 1;
                const uint64_t so = A.seq_off[rid];
                const uint64_t idx = min_base + hpre + kk;
                if (idx < A.min_cap)
                    A.min_out[idx] = make_uint4(h, (uint32_t)((uint64_t)g_start - so),
                                                (uint32_t)((uint64_t)g_own - (uint64_t)d - so), rid);
                else
                    atomicOr(A.err, ERR_CAP);
                ++kk;
            }
        }
        // ---- S8: per-sequence offsets for every sequence starting in this tile
        for (uint32_t i = lb + tid; i < ub; i += NT) {
            const uint64_t so = A.seq_off[i];
            const uint32_t x = (uint32_t)((int64_t)so - W0);
            const uint32_t qx = S.qoff[x >> 5] + __popc(S.keepw[x >> 5] & lowmask(x & 31));
            const uint32_t v = qx - hk;
            const uint32_t hb = S.hitpre[v >> 5] + __popc(S.hitw[v >> 5] & lowmask(v & 31));
            A.min_off[i] = min_base + hb;
            if (A.hpc_off) A.hpc_off[i] = kept_base + v;
        }
    }
}

// ------------------------------------------------------------------------------------------------ read counts
constexpr int RT = 256;            // threads per CTA
constexpr int RPT = 4;             // sequences per thread
struct K2Args {
    const uint4    *mins;
    const uint64_t *min_off, *hpc_off, *seq_off;
    const uint8_t  *bases;
    uint64_t  n_seqs;
    uint32_t  l, k;
    int32_t   quirk, hpc;
    uint64_t *km_off;              // n_seqs + 1
    uint32_t *min_cnt;             // n_seqs
    uint64_t *status;              // per tile, zeroed
    uint32_t *ticket;
    uint32_t *err;
};

// Minimizers of sequence r that reach the window stage.  In the AVX-512 profile of ntHash1 the iterator
// masks the last block with (1 << (S % 16)) - 1 (src/nthash_avx512_32.rs:134-138): when S > 16 and
// S % 16 == 0 the last 16 l-mer positions are lost.  Their minimizers are a suffix of the sequence's list:
// exactly those whose last base is one of the final 16 kept bases.
__device__ __forceinline__ uint32_t window_feed_count(const K2Args &A, uint64_t r)
{
    const uint64_t m0 = A.min_off[r], m1 = A.min_off[r + 1];
    uint64_t cnt = m1 - m0;
    if (A.quirk && cnt > 0) {
        const uint64_t so = A.seq_off[r], se = A.seq_off[r + 1];
        const uint64_t M = A.hpc ? (A.hpc_off[r + 1] - A.hpc_off[r]) : (se - so);
        if (M >= (uint64_t)A.l + 16 && ((M - A.l + 1) & 15) == 0) {
            uint64_t e16;                                   // local position of the 16th kept base from the end
            if (!A.hpc) {
                e16 = (se - so) - 16;
            } else {
                uint64_t g = se;
                int found = 0;
                while (found < 16) {                        // M >= 16 guarantees termination above `so`
                    --g;
                    if (g == so || A.bases[g] != A.bases[g - 1]) ++found;
                }
                e16 = g - so;
            }
            while (cnt > 0 && (uint64_t)A.mins[m0 + cnt - 1].z >= e16) --cnt;
        }
    }
    return (uint32_t)cnt;
}

__global__ void __launch_bounds__(RT) k_read_counts(const __grid_constant__ K2Args A)
{
    S2K_SHARED uint32_t wsum[RT / 32];
    S2K_SHARED unsigned long long s_excl;
    S2K_SHARED uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t n_tiles = (A.n_seqs + RT * RPT - 1) / (RT * RPT);
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(A.ticket, 1u);
        __syncthreads();
        const uint32_t t = s_tile;
        if (t >= n_tiles) break;
        const uint64_t r0 = (uint64_t)t * (RT * RPT) + (uint64_t)tid * RPT;
        uint32_t items[RPT], sum = 0;
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            items[j] = 0;
            const uint64_t r = r0 + j;
            if (r < A.n_seqs) {
                const uint32_t c = window_feed_count(A, r);
                A.min_cnt[r] = c;
                items[j] = c >= A.k ? c - A.k + 1 : 0;
            }
            sum += items[j];
        }
        // block scan
        uint32_t incl = warp_incl_scan(sum, lane);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t pre = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < RT / 32; ++i) { const uint32_t s = wsum[i]; if (i < warp) pre += s; tot += s; }
        const uint32_t excl_local = pre + incl - sum;
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
            if (lane == 0) {
                st_relaxed(&A.status[t], FLAG_INCL | (excl + tot));
                s_excl = excl;
                if (t == n_tiles - 1) A.km_off[A.n_seqs] = excl + tot;
            }
        }
        __syncthreads();
        uint64_t o = s_excl + excl_local;
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
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
    uint64_t  n_min;
    uint32_t  k;
    uint64_t *hash;
    uint32_t *start, *end;
    uint8_t  *rev;
};
__device__ __forceinline__ uint64_t mix32(uint32_t h)      // MixHash for u32, src/lib.rs:157-169
{
    uint64_t x = h;
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    return x;
}
__device__ __forceinline__ uint64_t rol64(uint64_t x, uint32_t r)
{
    r &= 63u;
    return r ? ((x << r) | (x >> (64u - r))) : x;
}
__global__ void __launch_bounds__(256) k_windows(const __grid_constant__ K3Args A)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.n_min; g += stride) {
        const uint4 first = A.mins[g];
        const uint32_t rid = first.w;
        const uint64_t c = g - A.min_off[rid];             // window index inside the sequence == offset
        const uint64_t k0 = A.km_off[rid];
        if (c >= A.km_off[rid + 1] - k0) continue;         // fewer than k minimizers left (or tail rule)
        uint64_t f = 0, r = 0;
        uint32_t end = first.z;
        for (uint32_t tt = 0; tt < A.k; ++tt) {
            const uint4 mrec = tt ? A.mins[g + tt] : first;
            const uint64_t m = mix32(mrec.x);
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

// ------------------------------------------------------------------------------------------------ RLE (encode_rle_simd)
// src/hpc.rs:44-147 over a batch: kept bytes + run starts, ordered.  Same keep rule and tile order machinery,
// no hashing.  One thread handles 32 bases; ordered by a decoupled look-back on kept counts.
struct K4Args {
    const uint8_t  *bases;
    const uint64_t *seq_off;
    const uint32_t *tile_lb;
    uint64_t n_seqs, n_bases;
    uint32_t n_tiles;
    uint64_t *status; uint32_t *ticket; uint32_t *err;
    uint8_t  *hpc; uint32_t *pos; uint64_t *hpc_off;
};
constexpr int RLE_TILE = NT * 32;
__global__ void __launch_bounds__(NT) k_rle(const __grid_constant__ K4Args A)
{
    S2K_SHARED uint32_t startw[NT];
    S2K_SHARED uint32_t keepw[NT + 1], qoff[NT + 1];
    S2K_SHARED uint32_t wsum[NT / 32];
    S2K_SHARED uint32_t s_tile;
    S2K_SHARED unsigned long long s_excl;
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
        __syncthreads();
        for (uint32_t i = lb + tid; i < ub; i += NT) {
            const uint64_t so = A.seq_off[i];
            if (so < T1) { const uint32_t x = (uint32_t)(so - T0); atomicOr(&startw[x >> 5], 1u << (x & 31)); }
        }
        __syncthreads();
        const uint64_t g0 = T0 + 32ull * tid;
        uint32_t keep = 0;
        uint8_t by[32];
        uint8_t prev = (g0 > 0 && g0 <= A.n_bases) ? A.bases[g0 - 1] : 0;
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const uint64_t g = g0 + b;
            by[b] = g < T1 ? A.bases[g] : 0;
            if (g < T1 && (by[b] != prev || ((startw[tid] >> b) & 1u))) keep |= 1u << b;
            prev = by[b];
        }
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
            uint64_t o = base + q;
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                if ((keep >> b) & 1u) {
                    const uint64_t g = g0 + b;
                    while (g >= nx) { ++rid; so = A.seq_off[rid]; nx = A.seq_off[rid + 1]; }
                    A.hpc[o] = by[b];
                    A.pos[o] = (uint32_t)(g - so);
                    ++o;
                }
            }
        }
        for (uint32_t i = lb + tid; i < ub; i += NT) {
            const uint64_t s = A.seq_off[i];
            const uint32_t x = (uint32_t)(s - T0);
            A.hpc_off[i] = base + qoff[x >> 5] + __popc(keepw[x >> 5] & lowmask(x & 31));
        }
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
