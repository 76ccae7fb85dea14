// s2k_kernel_fast.cuh -- k_minimizers_fast: the minimizer kernel for the common case, rolling in RAW space.
//
// ncu on k_minimizers (profiles/r1g_*) shows the L1/shared-memory data pipe at 81 % of peak: the byte stores of the
// HPC compaction and the two class-byte loads per rolling step.  This kernel has neither.  Each thread walks its own
// 256 raw bases (plus 64 bases of warm-up in front of them) straight from registers, 16 at a time; a base that equals
// its predecessor is skipped by predication; the classes of the last l kept bases live in a 64-bit register FIFO
// (2 bits each), so the leaving base of the rolling update needs no memory at all.  Per kept base that leaves one LUT
// byte load, one 8-byte table load and 12 ALU/FMA operations.  Sequence starts are rare events: a per-thread counter of
// kept bases since the last start decides, only when a hash passes the threshold, whether the l-mer lies inside one
// sequence.  Hits go to a short per-thread list (hash, raw offset, kept index) in an L2-resident scratch; after one
// block-wide prefix of the per-thread counts every thread writes its own records.
//
// Exactness: the results are identical to k_minimizers' (same tile_info / record / per-sequence-prefix outputs, checked
// by the same parity tests).  The kernel covers HPC modes with 20 <= l <= 31 and declines -- sets ERR_FAST, the host
// reruns the batch with k_minimizers -- when it meets something outside its model: a kept base that is not A/C/G/T under
// the mode's base map, a thread whose 64 warm-up bases hold fewer than l-1(+1) kept bases (long homopolymers), or more
// hits in one thread's 256 bases than its list holds (very high density).
#pragma once

namespace s2k {

constexpr int FRC   = 256;                  // raw bases owned per thread
constexpr int FWU   = 64;                   // warm-up bases in front of them
constexpr int FTILE = NT * FRC;             // 65 536 raw bases per CTA tile
constexpr int FBW   = (FWU + FTILE) / 32 + 2;   // words of the raw-space start bitmaps (window starts FWU before the tile)
constexpr int FSOC  = 1024;                 // sequence offsets cached per tile
constexpr int FDIRTY = 1024;
constexpr uint32_t ERR_FAST = 4u;

struct SmemF {
    uint32_t startw[FBW], shortw[FBW];      // bit x: a sequence starts at raw offset x - FWU of the tile (short: len <= l)
    uint32_t keepw[FTILE / 32];             // keep mask of the tile (thread t: words 8t .. 8t+7)
    unsigned long long soc[FSOC];
    uint32_t kpre[NT + 1], hpre[NT + 1];    // kept bases / hits of the tile before thread t
    uint16_t dirty[FDIRTY];
    uint2    xy[16];                        // ACGT x ACGT part of the (out,in) table
    uint8_t  lut[256];
    uint32_t wsum[8];
    uint32_t wkeep[2], wcnt;                // thread 0's warm-up: keep mask and count (l-mers that start before the tile)
    uint32_t n_dirty, tile_id;
    unsigned long long rec0;
};

__device__ __forceinline__ void fflag(SmemF &S, uint32_t x, bool is_short)
{
    atomicOr(&S.startw[x >> 5], 1u << (x & 31));
    if (is_short) atomicOr(&S.shortw[x >> 5], 1u << (x & 31));
    const uint32_t k = atomicAdd(&S.n_dirty, 1u);
    if (k < FDIRTY) S.dirty[k] = (uint16_t)(x >> 5);
}
// position of the r-th (0-based) kept base inside thread ts' 256-base chunk
__device__ __forceinline__ int fselect(const SmemF &S, int ts, int r)
{
    int wi = 8 * ts;
    for (;;) {
        const uint32_t kw = S.keepw[wi];
        const int c = __popc(kw);
        if (r < c) return (wi - 8 * ts) * 32 + nth_set_bit(kw, r);
        r -= c;
        ++wi;
    }
}

template <bool W31, int D>
__global__ void __launch_bounds__(NT, 4) k_minimizers_fast(const __grid_constant__ K1Args A)
{
    S2K_DYN_SMEM(smem_raw);
    SmemF &S = *reinterpret_cast<SmemF *>(smem_raw);
    const int tid = threadIdx.x;
    const int l = (int)A.l;
    const uint32_t need = (uint32_t)(l - 1 + D);
    const uint32_t osh = (uint32_t)(2 * l - 39);           // FIFO: class of the base leaving the l-mer, as (class << 5)
    const uint32_t hcap = A.fast_hcap;
    uint32_t *const list_h = A.hscr + ((size_t)blockIdx.x * NT + tid) * 2 * hcap;
    uint32_t *const list_x = list_h + hcap;

    for (int i = tid; i < 256; i += NT) S.lut[i] = A.cls_lut[i];
    if (tid < 16) S.xy[tid] = A.xy[tid];
    for (int i = tid; i < FBW; i += NT) { S.startw[i] = 0; S.shortw[i] = 0; }
    if (tid == 0) S.n_dirty = 0;

    for (;;) {
        __syncthreads();                                   // B1
        if (tid == 0) S.tile_id = atomicAdd(A.ticket, 1u);
        {
            const uint32_t nd = S.n_dirty;
            if (nd > (uint32_t)FDIRTY) { for (int i = tid; i < FBW; i += NT) { S.startw[i] = 0; S.shortw[i] = 0; } }
            else for (uint32_t i = tid; i < nd; i += NT) { const uint32_t e = S.dirty[i]; S.startw[e] = 0; S.shortw[e] = 0; }
        }
        __syncthreads();                                   // B2
        if (tid == 0) S.n_dirty = 0;
        const uint32_t t = S.tile_id;
        if (t >= A.n_tiles) break;
        const int64_t T0 = (int64_t)((uint64_t)t * FTILE);
        const int64_t T1 = min(T0 + (int64_t)FTILE, (int64_t)A.n_bases);
        const bool last_tile = (uint64_t)T1 == A.n_bases;
        const uint32_t lb = A.tile_lb[t];
        const uint32_t ub = last_tile ? (uint32_t)(A.n_seqs + 1) : A.tile_lb[t + 1];
        const uint32_t c0 = lb > 0 ? lb - 1 : 0;
        const uint32_t c_hi = ub < (uint32_t)A.n_seqs ? ub : (uint32_t)A.n_seqs;
        const bool cached = c_hi - c0 + 1 <= (uint32_t)FSOC;
        if (cached) for (uint32_t i = tid; i <= c_hi - c0; i += NT) S.soc[i] = A.seq_off[c0 + i];
        __syncthreads();                                   // B3 (n_dirty reset and cache visible)
        auto so_at = [&](uint32_t i) -> uint64_t { return cached ? S.soc[i - c0] : A.seq_off[i]; };
        // sequence starts of the window [T0 - FWU, T1)
        for (uint32_t i = lb + tid; i < ub; i += NT) {
            const uint64_t so = so_at(i);
            if (so < (uint64_t)T1) {
                const uint64_t len = so_at(i + 1) - so;
                fflag(S, (uint32_t)((int64_t)so - T0 + FWU), len > 0 && len <= (uint64_t)l);
            }
        }
        if (tid == 0) {                                    // starts inside the warm-up bases in front of the tile
            for (uint32_t i = lb; i > 0;) {
                --i;
                const uint64_t so = A.seq_off[i];
                if ((int64_t)so < T0 - FWU) break;
                const uint64_t len = A.seq_off[i + 1] - so;
                if ((int64_t)so < T0) fflag(S, (uint32_t)((int64_t)so - T0 + FWU), len > 0 && len <= (uint64_t)l);
            }
        }
        __syncthreads();                                   // B4

        // ================================================================ rolling pass: no block barrier inside
        const int64_t gT = T0 + (int64_t)FRC * tid;        // this thread's 256 bases start here, warm-up FWU earlier
        uint32_t fh = A.fast_fh0, rh = A.fast_rh0;         // state and FIFO of l bases 'A': consistent by construction
        uint32_t fhi = 0, flo = 0;
        uint32_t h = min(fh, rh);
        uint32_t run = 0, shortcur = 0, kept = 0, wkept = 0, hits = 0, acc = 0, wu_start = 0;
        uint32_t kw_acc = 0;
        uint32_t prevb = 0;
        {
            const int64_t gp = gT - FWU - 1;
            if (gp >= 0 && gp < (int64_t)A.n_bases) prevb = A.bases[gp];
        }
#pragma unroll 1
        for (int p = 0; p < (FWU + FRC) / 16; ++p) {
            const int64_t g = gT - FWU + 16 * p;
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            if (g >= 0 && g + 16 <= (int64_t)A.n_bases) {
                const uint4 x = __ldg(reinterpret_cast<const uint4 *>(A.bases + g));
                w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
            } else if (g + 16 > 0 && g < (int64_t)A.n_bases) {
                for (int j = 0; j < 16; ++j) {
                    const int64_t gg = g + j;
                    if (gg >= 0 && gg < (int64_t)A.n_bases) w[j >> 2] |= (uint32_t)A.bases[gg] << (8 * (j & 3));
                }
            }
            // keep mask of the 16 bases
            uint32_t keep16 = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t sh = (w[i] << 8) | prevb;
                prevb = w[i] >> 24;
                const uint32_t neq = __vcmpne4(w[i], sh);
                keep16 |= (((neq & 0x08040201u) * 0x01010101u) >> 24) << (4 * i);
            }
            const uint32_t bx = (uint32_t)(FRC * tid + 16 * p);        // window offset of the piece (multiple of 16)
            const uint32_t sbits = (S.startw[bx >> 5] >> (bx & 31)) & 0xffffu;
            uint32_t vmask = 0xffffu;
            if (g < 0) vmask = (g <= -16) ? 0u : ((0xffffu << (int)(-g)) & 0xffffu);
            const int64_t rem = T1 - g;
            if (rem <= 0) vmask = 0u; else if (rem < 16) vmask &= (1u << (int)rem) - 1u;
            keep16 = (keep16 | sbits) & vmask;
            const uint32_t s16 = sbits & keep16;
            const bool owned = p >= FWU / 16;

            uint32_t hv[4];
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                if (D == 1) hv[b & 3] = h;                 // Hpc: the kept base that shows up completes the PREVIOUS l-mer
                if ((keep16 >> b) & 1u) {
                    const uint32_t in8 = S.lut[(w[b >> 2] >> (8 * (b & 3))) & 0xffu];
                    acc |= in8;
                    const uint32_t idx = ((fhi >> osh) & 0x60u) | in8;
                    const uint2 tt = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(S.xy) + idx);
                    fh = rol1<W31>(fh) ^ tt.x;
                    rh = ror1<W31>(rh) ^ tt.y;
                    h = min(fh, rh);
                    fhi = (fhi << 2) | (flo >> 30);
                    flo = (flo << 2) | (in8 >> 3);
                }
                if (D == 0) hv[b & 3] = h;
                if ((b & 3) == 3 && owned && min(min(hv[0], hv[1]), min(hv[2], hv[3])) <= A.thr) {
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) {
                        const int bb = b - 3 + k2;
                        if (((keep16 >> bb) & 1u) && hv[k2] <= A.thr) {
                            const uint32_t below = keep16 & lowmask(bb + 1);      // kept bases up to and including bb
                            const uint32_t sl = s16 & lowmask(bb + 1);
                            uint32_t run_at, short_at;
                            if (sl) {
                                const int ls = 31 - __clz(sl);
                                run_at = __popc(below >> ls);
                                short_at = (S.shortw[bx >> 5] >> ((bx & 31) + ls)) & 1u;
                            } else { run_at = run + __popc(below); short_at = shortcur; }
                            if (run_at >= (uint32_t)(l + D) + short_at) {
                                if (hits < hcap) {
                                    list_h[hits] = hv[k2];
                                    list_x[hits] = (uint32_t)(FRC * tid + 16 * (p - FWU / 16) + bb) |
                                                   ((kept + __popc(below) - 1u) << 16);
                                }
                                ++hits;
                            }
                        }
                    }
                }
            }
            // counters, keep bitmap
            if (s16) {
                const int ls = 31 - __clz(s16);
                run = __popc(keep16 >> ls);
                shortcur = (S.shortw[bx >> 5] >> ((bx & 31) + ls)) & 1u;
            } else run += __popc(keep16);
            if (owned) {
                kept += __popc(keep16);
                kw_acc |= keep16 << (16 * (p & 1));
                if (p & 1) { S.keepw[8 * tid + ((p - FWU / 16) >> 1)] = kw_acc; kw_acc = 0; }
            } else {
                wkept += __popc(keep16);
                wu_start |= s16;
                if (tid == 0) {
                    kw_acc |= keep16 << (16 * (p & 1));
                    if (p & 1) { S.wkeep[p >> 1] = kw_acc; kw_acc = 0; }
                }
            }
        }
        // outside the model -> the host reruns the batch with the general kernel
        if ((acc & 0x80u) || hits > hcap || (!wu_start && wkept < need && gT > 0 && gT < T1)) atomicOr(A.err, ERR_FAST);
        if (tid == 0) S.wcnt = wkept;

        // ================================================================ prefix of per-thread counts, allocation
        uint32_t tile_kept, tile_hits;
        const uint32_t kp = block_excl_scan(kept, S.wsum, tile_kept);
        const uint32_t hp = block_excl_scan(min(hits, hcap), S.wsum, tile_hits);
        S.kpre[tid] = kp; S.hpre[tid] = hp;
        if (tid == NT - 1) { S.kpre[NT] = tile_kept; S.hpre[NT] = tile_hits; }
        if (tid == 0) {
            const unsigned long long r0 = tile_hits ? atomicAdd(A.cursor, (unsigned long long)tile_hits) : 0ull;
            S.rec0 = r0;
            A.tile_info[t] = make_uint4(tile_hits, tile_kept, (uint32_t)r0, (uint32_t)(r0 >> 32));
            if (r0 + tile_hits > A.min_cap) atomicOr(A.err, ERR_CAP);
        }
        __syncthreads();                                   // B5

        // ================================================================ records: every thread writes its own hits
        {
            const uint64_t rec0 = S.rec0 + hp;
            const uint32_t nh = min(hits, hcap);
            for (uint32_t i = 0; i < nh; ++i) {
                const uint32_t hsh = list_h[i], xq = list_x[i];
                const uint32_t x = xq & 0xffffu;
                const int64_t g_own = T0 + x;
                const int Qs = (int)(kp + (xq >> 16)) - (int)need;        // tile kept index of the l-mer's first base
                int64_t g_start;
                if (Qs >= 0) {
                    int lo = 0, hi = NT;                                   // kpre[lo] <= Qs < kpre[hi]
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((int)S.kpre[mid] <= Qs) lo = mid; else hi = mid; }
                    g_start = T0 + FRC * lo + fselect(S, lo, Qs - (int)S.kpre[lo]);
                } else {
                    int r = (int)S.wcnt + Qs;                              // kept base inside the 64 bases before the tile
                    uint32_t kw = S.wkeep[0];
                    int base = 0;
                    if (r >= __popc(kw)) { r -= __popc(kw); kw = S.wkeep[1]; base = 32; }
                    g_start = T0 - FWU + base + nth_set_bit(kw, r < 0 ? 0 : r);
                }
                uint32_t lo2 = lb, hi2 = ub;                               // first i in [lb,ub) with seq_off[i] > g_own
                while (lo2 < hi2) {
                    const uint32_t mid = lo2 + ((hi2 - lo2) >> 1);
                    if (so_at(mid) <= (uint64_t)g_own) lo2 = mid + 1; else hi2 = mid;
                }
                const uint32_t rid = lo2 - 1;
                const uint64_t so = rid >= c0 ? so_at(rid) : A.seq_off[rid];
                const uint64_t idx = rec0 + i;
                if (idx < A.min_cap)
                    A.min_out[idx] = make_uint4(hsh, (uint32_t)((uint64_t)g_start - so),
                                                (uint32_t)((uint64_t)g_own - (uint64_t)D - so), rid);
            }
        }
        // tile-local prefixes for every sequence starting in the tile
        for (uint32_t i = lb + tid; i < ub; i += NT) {
            const uint64_t so = so_at(i);
            const uint32_t x = (uint32_t)((int64_t)so - T0);               // 0 .. FTILE
            const uint32_t ts = x >> 8, xr = x & 255u;
            uint32_t v = S.kpre[ts], hb = S.hpre[ts];
            if (ts < (uint32_t)NT) {
                for (uint32_t wi = 0; wi < (xr >> 5); ++wi) v += __popc(S.keepw[8 * ts + wi]);
                if (xr & 31u) v += __popc(S.keepw[8 * ts + (xr >> 5)] & lowmask(xr & 31u));
                const uint32_t nh = S.hpre[ts + 1] - S.hpre[ts];
                const uint32_t *lx = A.hscr + ((size_t)blockIdx.x * NT + ts) * 2 * hcap + hcap;
                for (uint32_t j = 0; j < nh; ++j) if ((lx[j] & 0xffffu) < x) ++hb;
            }
            A.min_off[i] = hb;
            if (A.hpc_off) A.hpc_off[i] = v;
        }
    }
}

} // namespace s2k
