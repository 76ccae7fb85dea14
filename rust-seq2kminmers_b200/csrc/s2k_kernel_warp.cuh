// s2k_kernel_warp.cuh -- k_minimizers_w: the minimizer kernel with WARP-independent sub-tiles.
//
// Same semantics, helpers, arguments and outputs as k_minimizers (s2k_kernels.cuh); what changes is who waits for whom.
// k_minimizers runs a 16 K tile through ~13 block-wide phases; ncu showed block barriers as its top stall reason.  Here
// a CTA still claims a 16 384-base tile, but each of its 8 warps owns a 2 048-base sub-tile with a private halo
// (64 raw bases, 256 for long l) and private shared-memory state, and does keep mask -> scan -> compaction -> flags ->
// walk-back -> rolling hash with __syncwarp() only.  Block barriers remain for: the tile ticket, the per-tile cache of
// sequence offsets, and the prefix of the 8 warps' hit counts before records are written (5 per tile).
#pragma once

namespace s2k {

constexpr int WSUB    = 2048;                         // raw bases per warp sub-tile
constexpr int WTILE   = 8 * WSUB;                     // raw bases per CTA tile
constexpr int WHB_MAX = 256;                          // largest per-warp halo, raw bytes
constexpr int WXB     = 256;                          // context capacity in kept bases (halo + walk-back)
constexpr int WCODE   = WXB + WHB_MAX + WSUB + 128;
constexpr int WSW     = (WHB_MAX + WSUB) / 32;        // raw-space flag words of a warp window
constexpr int WFW     = (WXB + WHB_MAX + WSUB) / 32 + 2;   // owner-space flag words
constexpr int WHL     = 256;                          // hit-list entries per emission round
constexpr int SOC     = 512;                          // sequence offsets cached per CTA tile
constexpr int WDIRTY  = 30;
constexpr int HSW     = 2 * 32 * 68;                  // hash-stash words per warp (two passes of the widest geometry)

template <int CHW> struct WarpS {
    static constexpr int MWW = (CHW + 63) / 64;
    uint8_t  code[WCODE];                             // class code of kept base q of the window at [WXB + q]
    unsigned long long hitw[2][33][MWW];
    uint32_t hitpre[2][33];
    uint32_t keepw[65];                               // keep mask of main chunk c (32 raw bases); [64] = 0
    uint32_t qoff[65];                                // kept bases of the window (halo included) before chunk c; [64] = all
    uint32_t startw[WSW], shortw[WSW];                // raw-space flags over the window (halo words first)
    uint32_t f1[WFW], f2[WFW];                        // owner-space flags
    uint32_t ctxpos[WXB];                             // walk-back context: distance below the window start
    uint16_t hpos[WHB_MAX];                           // window offset of kept halo base q
    uint16_t qmap[(WHB_MAX + WSUB) / 64 + 2];
    uint16_t hl[WHL];
    uint16_t dirty[WDIRTY];
    uint32_t n_dirty;
};

template <int CHW> struct SmemW {
    WarpS<CHW> w[8];
    unsigned long long soc[SOC];                      // seq_off[c0 .. c0 + soc_n)
    uint2    xy[XYN];
    uint8_t  lut[256];
    uint32_t wtot_h[8], wtot_k[8], wpre_h[9], wpre_k[9];
    uint32_t tile_id;
    unsigned long long rec0;
};

template <int CHW> __device__ __forceinline__ void wflag_raw(WarpS<CHW> &W, uint32_t x, bool is_short)
{
    atomicOr(&W.startw[x >> 5], 1u << (x & 31));
    if (is_short) atomicOr(&W.shortw[x >> 5], 1u << (x & 31));
    const uint32_t k = atomicAdd(&W.n_dirty, 1u);
    if (k < WDIRTY) W.dirty[k] = (uint16_t)(x >> 5);
}
template <int CHW> __device__ __forceinline__ void wflag_owner(WarpS<CHW> &W, int oo, bool is_short)
{
    atomicOr(&W.f1[oo >> 5], 1u << (oo & 31));
    if (is_short) atomicOr(&W.f2[oo >> 5], 1u << (oo & 31));
    const uint32_t k = atomicAdd(&W.n_dirty, 1u);
    if (k < WDIRTY) W.dirty[k] = (uint16_t)(0x8000u | (uint32_t)(oo >> 5));
}
__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v, int lane) { return warp_incl_scan(v, lane); }

template <bool HPC, bool W31, int CHW>
__global__ void __launch_bounds__(256, 3) k_minimizers_w(const __grid_constant__ K1Args A)
{
    S2K_DYN_SMEM(smem_raw);
    SmemW<CHW> &S = *reinterpret_cast<SmemW<CHW> *>(smem_raw);
    constexpr int MWW = WarpS<CHW>::MWW;
    constexpr int CAPW = 32 * CHW;
    static_assert(CHW % 4 == 0 && ((CHW / 4) & 1) == 1 && 2 * CAPW >= WSUB && 2 * CAPW <= HSW, "warp hash geometry");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = (int)A.l, d = (int)A.d;
    const int HB = (int)A.halo, hb = HB >> 5;              // per-warp halo: 64 or 256 raw bytes, 2 or 8 per lane
    WarpS<CHW> &W = S.w[warp];
    uint32_t *const hs = A.hscr + ((size_t)blockIdx.x * 8 + warp) * HSW;

    for (int i = tid; i < 256; i += 256) S.lut[i] = A.cls_lut[i];
    if (tid < XYN) S.xy[tid] = A.xy[tid];
    for (int i = lane; i < WCODE; i += 32) W.code[i] = ZC8;
    for (int i = lane; i < WSW; i += 32) { W.startw[i] = 0; W.shortw[i] = 0; }
    for (int i = lane; i < WFW; i += 32) { W.f1[i] = 0; W.f2[i] = 0; }
    if (lane == 0) W.n_dirty = 0;

    for (;;) {
        __syncthreads();                                   // B1: everyone is done with the previous tile
        if (tid == 0) S.tile_id = atomicAdd(A.ticket, 1u);
        __syncthreads();                                   // B2
        const uint32_t t = S.tile_id;
        if (t >= A.n_tiles) break;
        const int64_t T0 = (int64_t)((uint64_t)t * WTILE);
        const int64_t T1 = min(T0 + (int64_t)WTILE, (int64_t)A.n_bases);
        const bool last_tile = (uint64_t)T1 == A.n_bases;
        const uint32_t lb = A.tile_lb[t];
        const uint32_t ub = last_tile ? (uint32_t)(A.n_seqs + 1) : A.tile_lb[t + 1];
        // cache of the sequence offsets this tile can touch: indices [c0, min(ub, n_seqs)]
        const uint32_t c0 = lb > 0 ? lb - 1 : 0;
        const uint32_t c_hi = ub < (uint32_t)A.n_seqs ? ub : (uint32_t)A.n_seqs;
        const bool cached = c_hi - c0 + 1 <= (uint32_t)SOC;
        if (cached) for (uint32_t i = tid; i <= c_hi - c0; i += 256) S.soc[i] = A.seq_off[c0 + i];
        __syncthreads();                                   // B3
        auto so_at = [&](uint32_t i) -> uint64_t { return cached ? S.soc[i - c0] : A.seq_off[i]; };

        // ================================================================ warp-independent part 1
        const int64_t s0w = T0 + (int64_t)WSUB * warp;     // this warp's sub-tile [s0w, s1w), window starts HB earlier
        const int64_t s1w = min(s0w + (int64_t)WSUB, T1);
        const int64_t W0w = s0w - HB;
        const bool active = s0w < T1;
        uint32_t hk = 0, n_own = 0, warp_hits = 0, lbw = lb, ubw = lb;
        if (active) {
            // ---- a. clear the flag words the previous tile touched
            {
                const uint32_t nd = W.n_dirty;
                if (nd > (uint32_t)WDIRTY) {
                    for (int i = lane; i < WSW; i += 32) { W.startw[i] = 0; W.shortw[i] = 0; }
                    for (int i = lane; i < WFW; i += 32) { W.f1[i] = 0; W.f2[i] = 0; }
                } else if ((uint32_t)lane < nd) {
                    const uint32_t e = W.dirty[lane];
                    if (e & 0x8000u) { W.f1[e & 0x7fffu] = 0; W.f2[e & 0x7fffu] = 0; }
                    else { W.startw[e] = 0; W.shortw[e] = 0; }
                }
                __syncwarp();
                if (lane == 0) W.n_dirty = 0;
                __syncwarp();
            }
            // ---- c. raw bases: 64 per lane of the sub-tile, hb per lane of the halo (issued before the flag work)
            const int64_t g0 = s0w + 64 * lane;
            uint32_t w[16];
            if (g0 + 64 <= (int64_t)A.n_bases) {
                const uint4 *src = reinterpret_cast<const uint4 *>(A.bases + g0);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const uint4 x = __ldg(src + v);
                    w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
                }
            } else {
#pragma unroll
                for (int v = 0; v < 16; ++v) {
                    uint32_t x = 0;
                    for (int j = 0; j < 4; ++j) {
                        const int64_t gg = g0 + 4 * v + j;
                        if (gg < (int64_t)A.n_bases) x |= (uint32_t)A.bases[gg] << (8 * j);
                    }
                    w[v] = x;
                }
            }
            const int64_t gh = W0w + hb * lane;
            unsigned long long hv = 0ull;
            for (int j = 0; j < hb; ++j) {
                const int64_t gg = gh + j;
                if (gg >= 0) hv |= (unsigned long long)A.bases[gg] << (8 * j);
            }
            // ---- b. sequences starting in this window
            {
                uint32_t lo = lb, hi = ub;                 // first i in [lb,ub) with seq_off[i] >= s0w
                while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (so_at(mid) < (uint64_t)s0w) lo = mid + 1; else hi = mid; }
                lbw = lo;
                if (last_tile && s1w == T1) ubw = ub;
                else {
                    hi = ub;                               // first i in [lbw,ub) with seq_off[i] >= s1w
                    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (so_at(mid) < (uint64_t)s1w) lo = mid + 1; else hi = mid; }
                    ubw = lo;
                }
            }
            int64_t s0r = s0w;                             // start of the sequence containing s0w
            if (!(lbw < ub && lbw <= (uint32_t)A.n_seqs && so_at(lbw) == (uint64_t)s0w)) s0r = (int64_t)so_at(lbw - 1);
            for (uint32_t i = lbw + lane; i < ubw; i += 32) {
                const uint64_t so = so_at(i);
                if (so < (uint64_t)s1w) {
                    const uint64_t len = so_at(i + 1) - so;
                    wflag_raw(W, (uint32_t)((int64_t)so - W0w), len > 0 && len <= (uint64_t)l);
                }
            }
            if (lane == 0 && s0r < s0w && s0r >= W0w)
                wflag_raw(W, (uint32_t)(s0r - W0w), so_at(lbw) - (uint64_t)s0r <= (uint64_t)l);
            __syncwarp();

            // ---- d. keep masks
            uint32_t klo, khi, kh;
            {
                const uint32_t hmask = hb == 2 ? 0x3u : 0xffu;
                const uint32_t hx = (uint32_t)(hb * lane);
                const uint32_t sb_h = (W.startw[hx >> 5] >> (hx & 31)) & hmask;
                uint32_t vm_h = hmask;                     // halo validity: only g >= 0 can fail
                if (gh < 0) vm_h = (gh <= -(int64_t)hb) ? 0u : ((hmask << (int)(-gh)) & hmask);
                if (HPC) {
                    const uint32_t last_h = (uint32_t)(hv >> (8 * (hb - 1))) & 0xffu;
                    uint32_t prevh = __shfl_up_sync(0xffffffffu, last_h, 1);
                    if (lane == 0) prevh = (W0w > 0) ? (uint32_t)A.bases[W0w - 1] : 0u;
                    kh = 0;
                    for (int j = 0; j < hb; ++j) {
                        const uint32_t b = (uint32_t)(hv >> (8 * j)) & 0xffu;
                        if (b != prevh) kh |= 1u << j;
                        prevh = b;
                    }
                    kh |= sb_h;
                    uint32_t prevb = __shfl_up_sync(0xffffffffu, w[15] >> 24, 1);
                    const uint32_t halo_last = __shfl_sync(0xffffffffu, last_h, 31);
                    if (lane == 0) prevb = halo_last;
                    uint32_t kk[2] = {0u, 0u};
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t sh = (w[i] << 8) | prevb;
                        prevb = w[i] >> 24;
                        const uint32_t neq = __vcmpne4(w[i], sh);
                        kk[i >> 3] |= (((neq & 0x08040201u) * 0x01010101u) >> 24) << (4 * (i & 7));
                    }
                    const int wi0 = (HB >> 5) + 2 * lane;
                    klo = kk[0] | W.startw[wi0];
                    khi = kk[1] | W.startw[wi0 + 1];
                } else {
                    kh = hmask; klo = khi = 0xffffffffu;
                }
                kh &= vm_h;
                unsigned long long vmask = ~0ull;
                const int64_t rem = s1w - g0;
                if (rem <= 0) vmask = 0ull; else if (rem < 64) vmask = (1ull << (int)rem) - 1ull;
                klo &= (uint32_t)vmask; khi &= (uint32_t)(vmask >> 32);
            }
            // ---- e. one warp scan of (halo count | main count << 16)
            const uint32_t clo = __popc(klo), cm = clo + __popc(khi), chh = __popc(kh);
            const uint32_t pk = chh | (cm << 16);
            const uint32_t incl = warp_incl_scan(pk, lane);
            const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t ex = incl - pk;
            hk = tot & 0xffffu;
            const uint32_t qh = ex & 0xffffu, qm = hk + (ex >> 16), wk = hk + (tot >> 16);
            n_own = wk - hk;
            W.keepw[2 * lane] = klo; W.keepw[2 * lane + 1] = khi;
            W.qoff[2 * lane] = qm; W.qoff[2 * lane + 1] = qm + clo;
            if (lane == 31) { W.qoff[64] = wk; W.keepw[64] = 0; }
            if (lane == 0) W.qmap[hk >> 6] = 0;
            __syncwarp();
            for (uint32_t m = (qm + 63u) & ~63u; m < qm + cm; m += 64)
                W.qmap[m >> 6] = (uint16_t)(m < qm + clo ? 2 * lane : 2 * lane + 1);
            // ---- f. compaction: halo (hb predicated steps), sub-tile (64 predicated steps); flags -> owner space
            {
                uint32_t q = qh;
                for (int j = 0; j < hb; ++j)
                    if ((kh >> j) & 1u) {
                        W.code[WXB + q] = S.lut[(uint32_t)(hv >> (8 * j)) & 0xffu];
                        W.hpos[q] = (uint16_t)(hb * lane + j);
                        ++q;
                    }
                uint8_t *cp = W.code + WXB + qm;
#pragma unroll
                for (int b = 0; b < 64; ++b) {                 // LUT load unconditional, store predicated: no branches
                    const uint32_t kb = (b < 32 ? (klo >> b) : (khi >> (b - 32))) & 1u;
                    const uint8_t cls = S.lut[(w[b >> 2] >> (8 * (b & 3))) & 0xffu];
                    if (kb) *cp = cls;
                    cp += kb;
                }
                // sequence starts among the kept bases -> owner space (rare)
                const uint32_t hx = (uint32_t)(hb * lane);
                uint32_t swh = ((W.startw[hx >> 5] >> (hx & 31)) & (hb == 2 ? 0x3u : 0xffu)) & kh;
                const uint32_t shh = W.shortw[hx >> 5] >> (hx & 31);
                while (swh) {
                    const int b = __ffs(swh) - 1;
                    swh &= swh - 1;
                    wflag_owner(W, (int)qh + __popc(kh & lowmask(b)) - (int)hk + WXB, (shh >> b) & 1u);
                }
                const int wi0 = (HB >> 5) + 2 * lane;
                const unsigned long long keep = ((unsigned long long)khi << 32) | klo;
                unsigned long long sw = (((unsigned long long)W.startw[wi0 + 1] << 32) | W.startw[wi0]) & keep;
                const unsigned long long sh2 = ((unsigned long long)W.shortw[wi0 + 1] << 32) | W.shortw[wi0];
                while (sw) {
                    const int b = __ffsll((long long)sw) - 1;
                    sw &= sw - 1;
                    wflag_owner(W, (int)qm + __popcll(keep & lowmask64(b)) - (int)hk + WXB, (sh2 >> b) & 1ull);
                }
            }
            __syncwarp();
            // ---- g. not enough context in the halo -> walk back through the sequence (rare)
            if (HPC && s0r < W0w && hk < A.need) {
                uint32_t remaining = A.need - hk, taken = 0;
                int64_t hi = W0w;
                while (remaining > 0 && hi > s0r) {
                    const int64_t lo = max(s0r, hi - 32);
                    const int64_t g = lo + lane;
                    const bool valid = g < hi;
                    uint8_t b = 0, pb = 0;
                    if (valid) { b = A.bases[g]; if (g > s0r) pb = A.bases[g - 1]; }
                    const bool kp = valid && (g == s0r || b != pb);
                    const uint32_t m = __ballot_sync(0xffffffffu, kp);
                    const uint32_t above = (lane == 31) ? 0u : (m >> (lane + 1));
                    const uint32_t rank = __popc(above);
                    if (kp && rank < remaining) {
                        const uint32_t slot = taken + rank;        // 0 = nearest to the window
                        W.code[WXB - 1 - (int)slot] = S.lut[b];
                        W.ctxpos[slot] = (uint32_t)(W0w - g);
                        if (g == s0r) { const int oo = WXB - 1 - (int)slot - (int)hk; if (oo >= 0) wflag_owner(W, oo, false); }
                    }
                    const uint32_t c = min((uint32_t)__popc(m), remaining);
                    taken += c; remaining -= c; hi = lo;
                }
            }
            __syncwarp();

            // ---- h. rolling canonical ntHash over the sub-tile's owners
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                unsigned long long mask[MWW];
#pragma unroll
                for (int x = 0; x < MWW; ++x) mask[x] = 0ull;
                const int v0 = pass * CAPW + CHW * lane;
                if ((uint32_t)(pass * CAPW) < n_own && (uint32_t)v0 < n_own) {
                    const int n_u = min(CHW, (int)n_own - v0);
                    unsigned long long invalid[MWW];
#pragma unroll
                    for (int x = 0; x < MWW; ++x) invalid[x] = 0ull;
                    {
                        const int L1 = l - 1 + d, o0 = v0 + WXB;
                        const int w_hi = (o0 + CHW - 1) >> 5, w_lo = (o0 - L1 - 1) >> 5;
                        for (int wi = w_lo; wi <= w_hi; ++wi) {
                            uint32_t fw = W.f1[wi];
                            if (fw) {
                                const uint32_t sw2 = W.f2[wi];
                                while (fw) {
                                    const int b = __ffs(fw) - 1;
                                    fw &= fw - 1;
                                    int lo = wi * 32 + b - o0;
                                    int hi = lo + L1 + (int)((sw2 >> b) & 1u);
                                    lo = max(lo, 0); hi = min(hi, CHW);
#pragma unroll
                                    for (int x = 0; x < MWW; ++x) {
                                        const int a = max(lo - 64 * x, 0), e = min(hi - 64 * x, 64);
                                        if (e > a) invalid[x] |= lowmask64(e - a) << a;
                                    }
                                }
                            }
                        }
                    }
                    const uint8_t *cb = W.code + WXB + hk + v0 - d;    // cb[i]: last base of owner i's l-mer
                    uint32_t fh = 0, rh = 0;
                    for (int j = 1 - l; j < 0; ++j) {                  // warm-up: first l-1 bases of owner 0's l-mer
                        const uint2 tt = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(S.xy) + (ZC8 << 2) + cb[j]);
                        fh = rol1<W31>(fh) ^ tt.x;
                        rh = ror1<W31>(rh) ^ tt.y;
                    }
                    const uint8_t *co = cb - l;
#pragma unroll
                    for (int i0 = 0; i0 < CHW; i0 += 4) {
                        uint32_t hv4[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int i = i0 + k;
                            const uint32_t in8 = cb[i];
                            const uint32_t out8 = i > 0 ? (uint32_t)co[i] : (uint32_t)ZC8;
                            const uint2 tt = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(S.xy) + (out8 << 2) + in8);
                            fh = rol1<W31>(fh) ^ tt.x;
                            rh = ror1<W31>(rh) ^ tt.y;
                            hv4[k] = min(fh, rh);
                        }
                        if (min(min(hv4[0], hv4[1]), min(hv4[2], hv4[3])) <= A.thr) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (hv4[k] <= A.thr) { mask[(i0 + k) >> 6] |= 1ull << ((i0 + k) & 63); hs[v0 + i0 + k] = hv4[k]; }
                        }
                    }
#pragma unroll
                    for (int x = 0; x < MWW; ++x)
                        mask[x] &= ~invalid[x] & lowmask64((uint32_t)max(min(n_u - 64 * x, 64), 0));
                }
                uint32_t cnt = 0;
#pragma unroll
                for (int x = 0; x < MWW; ++x) cnt += __popcll(mask[x]);
                const uint32_t inc = warp_incl_scan(cnt, lane);
                const uint32_t ptot = __shfl_sync(0xffffffffu, inc, 31);
#pragma unroll
                for (int x = 0; x < MWW; ++x) W.hitw[pass][lane][x] = mask[x];
                W.hitpre[pass][lane] = warp_hits + inc - cnt;
                warp_hits += ptot;
                if (lane == 31) {
#pragma unroll
                    for (int x = 0; x < MWW; ++x) W.hitw[pass][32][x] = 0ull;
                    W.hitpre[pass][32] = warp_hits;
                }
            }
        }
        if (lane == 0) { S.wtot_h[warp] = warp_hits; S.wtot_k[warp] = n_own; }
        __syncthreads();                                   // B4
        if (tid == 0) {
            uint32_t th = 0, tk = 0;
            for (int i = 0; i < 8; ++i) { S.wpre_h[i] = th; S.wpre_k[i] = tk; th += S.wtot_h[i]; tk += S.wtot_k[i]; }
            S.wpre_h[8] = th; S.wpre_k[8] = tk;
            const unsigned long long r0 = th ? atomicAdd(A.cursor, (unsigned long long)th) : 0ull;
            S.rec0 = r0;
            A.tile_info[t] = make_uint4(th, tk, (uint32_t)r0, (uint32_t)(r0 >> 32));
            if (r0 + th > A.min_cap) atomicOr(A.err, ERR_CAP);
        }
        __syncthreads();                                   // B5
        // ================================================================ warp-independent part 2
        if (active) {
            const uint64_t rec0 = S.rec0 + S.wpre_h[warp];
            // ---- j. ordered hit list of the warp, one lane per minimizer
            for (uint32_t base = 0; base < warp_hits; base += WHL) {
                __syncwarp();
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    uint32_t o = W.hitpre[pass][lane];
#pragma unroll
                    for (int x = 0; x < MWW; ++x) {
                        unsigned long long m = W.hitw[pass][lane][x];
                        while (m) {
                            const int i = __ffsll((long long)m) - 1 + 64 * x;
                            m &= m - 1;
                            if (o >= base && o < base + WHL) W.hl[o - base] = (uint16_t)(pass * CAPW + CHW * lane + i);
                            ++o;
                        }
                    }
                }
                __syncwarp();
                const uint32_t n_round = min((uint32_t)WHL, warp_hits - base);
                for (uint32_t j = lane; j < n_round; j += 32) {
                    const int v = W.hl[j];
                    const int qo = (int)hk + v;            // window kept index of the owner base (always in the sub-tile)
                    const uint32_t h = hs[v];
                    int c = W.qmap[qo >> 6];
                    while ((int)W.qoff[c + 1] <= qo) ++c;
                    const int64_t g_own = s0w + 32 * c + nth_set_bit(W.keepw[c], qo - (int)W.qoff[c]);
                    const int qs = qo - (l - 1 + d);
                    int64_t g_start;
                    if (qs < 0) g_start = W0w - (int64_t)W.ctxpos[-1 - qs];
                    else if (qs < (int)hk) g_start = W0w + (int64_t)W.hpos[qs];
                    else { while ((int)W.qoff[c] > qs) --c; g_start = s0w + 32 * c + nth_set_bit(W.keepw[c], qs - (int)W.qoff[c]); }
                    uint32_t lo = lbw, hi = ubw;           // first i in [lbw,ubw) with seq_off[i] > g_own
                    while (lo < hi) {
                        const uint32_t mid = lo + ((hi - lo) >> 1);
                        if (so_at(mid) <= (uint64_t)g_own) lo = mid + 1; else hi = mid;
                    }
                    const uint32_t rid = lo - 1;
                    const uint64_t so = so_at(rid);
                    const uint64_t idx = rec0 + base + j;
                    if (idx < A.min_cap)
                        A.min_out[idx] = make_uint4(h, (uint32_t)((uint64_t)g_start - so),
                                                    (uint32_t)((uint64_t)g_own - (uint64_t)d - so), rid);
                }
            }
            // ---- k. tile-local prefixes for every sequence starting in this sub-tile
            for (uint32_t i = lbw + lane; i < ubw; i += 32) {
                const uint64_t so = so_at(i);
                const uint32_t x = (uint32_t)((int64_t)so - s0w);
                const uint32_t qx = W.qoff[x >> 5] + __popc(W.keepw[x >> 5] & lowmask(x & 31));
                const uint32_t v = qx - hk;
                const uint32_t pass = v >= (uint32_t)CAPW ? 1u : 0u;
                const uint32_t vv = v - pass * CAPW, u = vv / CHW, bit = vv - u * CHW;
                uint32_t hbf = W.hitpre[pass][u];
#pragma unroll
                for (int xw = 0; xw < MWW; ++xw)
                    hbf += __popcll(W.hitw[pass][u][xw] & lowmask64((uint32_t)max(min((int)bit - 64 * xw, 64), 0)));
                A.min_off[i] = S.wpre_h[warp] + hbf;
                if (A.hpc_off) A.hpc_off[i] = S.wpre_k[warp] + v;
            }
        }
    }
}

} // namespace s2k
