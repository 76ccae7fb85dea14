/*
 * s2k_oracle.c -- CPU ORACLE for the sequence -> k-min-mer hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (rust-seq2kminmers_b200/csrc) never links, calls or falls back to anything here.
 *
 * What it is: a plain-C restatement of the reference crate's algorithm, written to
 * follow the reference's *procedural* logic (ring buffers, 16-lane blocks, tail masks)
 * so that its quirks are reproduced by construction, plus an independent closed-form
 * evaluator (s2k_oracle_closed_*) used by the tests to cross-check the restatement.
 * Reference citations are file:line into rchikhi/rust-seq2kminmers.
 *
 * Parity pinning: the reference cannot be built in this image (no rustc/cargo, nightly
 * features, un-vendored git crates), so the oracle is pinned by the reference's own
 * golden vectors: tests/main.rs:18-57 (KAT-1/KAT-2), src/old/nthash_hpc.rs.opt4:90-97
 * (KAT-3) and the structural equalities of tests/main.rs:76-89.  The 31-bit variant
 * (src/nthash2_avx512_32.rs) is NOT compiled or tested by the reference: for it,
 * "parity unpinned" -- the restatement below is the only spec.
 *
 * Third-party arithmetic absent from the tree: `nthash32` (git rchikhi/rust-nthash32,
 * unpinned, Cargo.toml:16), used by HashMode::Regular at src/lib.rs:108,217-229.  It is
 * restated from the published ntHash1 recurrence with u32 state and pinned by KAT-1.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define S2K_MODE_REGULAR 0
#define S2K_MODE_HPC 1
#define S2K_MODE_SIMD 2
#define S2K_MODE_HPCSIMD 3
#define S2K_NT1_32 0
#define S2K_NT2_31 1

/* ---------------------------------------------------------------- seeds / tables */
/* src/nthash_hpc.rs:29-49 and src/nthash_avx512_32.rs:233-236 */
static const uint64_t SEED64[4] = { 0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL,
                                    0x20323ed082572324ULL, 0x295549f54be24456ULL }; /* A C G T */

static inline uint32_t rol32(uint32_t x, unsigned r) { r &= 31; return r ? (x << r) | (x >> (32 - r)) : x; }
static inline uint32_t ror32(uint32_t x, unsigned r) { r &= 31; return r ? (x >> r) | (x << (32 - r)) : x; }
static inline uint64_t rol64(uint64_t x, unsigned r) { r &= 63; return r ? (x << r) | (x >> (64 - r)) : x; }
static inline uint64_t ror64(uint64_t x, unsigned r) { r &= 63; return r ? (x >> r) | (x << (64 - r)) : x; }

/* scalar 256-entry tables, src/nthash_hpc.rs:29-49: ACGT upper-case -> seed, N -> 0, else 1 */
static uint32_t H_LOOKUP[256], RC_LOOKUP[256];
static uint64_t H_LOOKUP64[256], RC_LOOKUP64[256];
static int tables_ready = 0;
static void init_tables(void)
{
    if (tables_ready) return;
    for (int i = 0; i < 256; i++) { H_LOOKUP[i] = RC_LOOKUP[i] = 1; H_LOOKUP64[i] = RC_LOOKUP64[i] = 1; }
    const char acgt[4] = { 'A', 'C', 'G', 'T' };
    for (int b = 0; b < 4; b++) {
        H_LOOKUP[(int)acgt[b]] = (uint32_t)SEED64[b];
        RC_LOOKUP[(int)acgt[b]] = (uint32_t)SEED64[3 - b];
        H_LOOKUP64[(int)acgt[b]] = SEED64[b];
        RC_LOOKUP64[(int)acgt[b]] = SEED64[3 - b];
    }
    H_LOOKUP['N'] = RC_LOOKUP['N'] = 0;
    H_LOOKUP64['N'] = RC_LOOKUP64['N'] = 0;
    tables_ready = 1;
}

/* SIMD base coding, src/nthash_avx512_32.rs:178-193: low nibble 1->0(A) 3->1(C) 7->2(G) 4->3(T) else 4 */
static inline int ckx(uint8_t b)
{
    static const int8_t table[16] = { 4, 0, 4, 1, 3, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4 };
    return table[b & 0x0f];
}
/* seed permutes, src/nthash_avx512_32.rs:242-277 (shift 0) and src/nthash2_avx512_32.rs:226-268 (>>33) */
static inline uint32_t lkf(uint8_t b, int shift) { int c = ckx(b); return c < 4 ? (uint32_t)(SEED64[c] >> shift) : 0; }
static inline uint32_t lkr(uint8_t b, int shift) { int c = ckx(b); return c < 4 ? (uint32_t)(SEED64[3 - c] >> shift) : 0; }

/* ---------------------------------------------------------------- bounds (A.1) */
/* src/lib.rs:91  hash_bound = ((density as f64) * (u32::MAX as f64)) as u32  (saturating cast) */
uint32_t s2k_oracle_bound_scalar(double density)
{
    double v = density * 4294967295.0;
    if (!(v > 0.0)) return 0;            /* NaN and negatives saturate to 0 */
    if (v >= 4294967295.0) return 0xffffffffu;
    return (uint32_t)v;
}
/* src/nthash_avx512_32.rs:47-48  density=(bound as f64)/(u32::MAX as f64); ((density as f32)*(u32::MAX as f32)) as u32 */
uint32_t s2k_oracle_bound_simd(uint32_t bound_scalar)
{
    double d = (double)bound_scalar / 4294967295.0;
    float f = (float)d;
    volatile float prod = f * 4294967296.0f; /* u32::MAX as f32 rounds to 2^32 */
    if (!(prod > 0.0f)) return 0;
    if (prod >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)prod;
}

/* ---------------------------------------------------------------- output sink */
typedef struct {
    uint64_t *start, *end;
    uint32_t *hash;
    size_t cap, n;
} minsink_t;
static inline void sink_push(minsink_t *s, uint64_t st, uint64_t en, uint32_t h)
{
    if (s->n < s->cap) {
        if (s->start) s->start[s->n] = st;
        if (s->end) s->end[s->n] = en;
        if (s->hash) s->hash[s->n] = h;
    }
    s->n++;
}

/* ---------------------------------------------------------------- HPC primitives */
/* src/hpc.rs:28-41 hpc(): collapse runs of any byte.  Returns hpc length. */
size_t s2k_oracle_hpc(const uint8_t *s, size_t n, uint8_t *out)
{
    size_t m = 0;
    int have = 0; uint8_t prev = 0;
    for (size_t i = 0; i < n; i++) {
        if (have && s[i] == prev) continue;
        if (have) out[m++] = prev;
        prev = s[i]; have = 1;
    }
    if (have) out[m++] = prev;
    return m;
}
/* src/hpc.rs:7-25 encode_rle(): collapses only runs of ACTGactgNn; returns bytes + run starts */
size_t s2k_oracle_encode_rle(const uint8_t *s, size_t n, uint8_t *out, uint64_t *pos)
{
    size_t m = 0, prev_i = 0;
    int have = 0; uint8_t prev = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t c = s[i];
        if (have && c == prev && strchr("ACTGactgNn", c) != NULL && c != 0) continue;
        if (have) { out[m] = prev; pos[m] = prev_i; m++; prev_i = i; }
        prev = c; have = 1;
    }
    if (have) { out[m] = prev; pos[m] = prev_i; m++; }
    return m;
}
/* src/hpc.rs:44-147 encode_rle_simd(): 16 bytes per step; mask = byte != previous byte, lane 0 of
 * block i>0 compares with the last byte of the previous block (:86-91), block 0 lane 0 always kept (:93-95);
 * tail block masked to len%16 lanes (:117-136).  For len<16 the reference compares s[0] with the byte in
 * front of the buffer (:125, undefined behaviour); the oracle keeps position 0, which is what
 * encode_rle returns and what tests/main.rs:77-78 asserts equal. */
size_t s2k_oracle_encode_rle_simd(const uint8_t *s, size_t n, uint8_t *out, uint32_t *pos)
{
    const size_t width = 16;
    size_t end_idx = n / width, m = 0;
    for (size_t i = 0; i <= end_idx; i++) {
        size_t lanes = width;
        if (i == end_idx) { lanes = n & (width - 1); if (lanes == 0) break; }
        uint16_t mask = 0;
        for (size_t j = 1; j < lanes; j++)
            if (s[i * width + j] != s[i * width + j - 1]) mask |= (uint16_t)(1u << j);
        if (i > 0) mask |= (uint16_t)(s[i * width] != s[i * width - 1]);
        else mask |= 1;
        for (size_t j = 0; j < lanes; j++)
            if (mask & (1u << j)) { out[m] = s[i * width + j]; pos[m] = (uint32_t)(i * width + j); m++; }
    }
    return m;
}

/* ---------------------------------------------------------------- HashMode::Hpc  (scalar, fused) */
/* Restates NtHashHPCIterator::new (src/nthash_hpc.rs:115-189) and ::next (:196-283), including the
 * 256-entry ring buffers, the backward walk for the reverse strand, `hash <= hash_bound` (:232,:277),
 * the item (run-start of first base, current_idx_plus_k-1, hash) (:234,:281) and the early `return None`
 * when the sequence is exhausted (:220-222, :265-267) which drops the final HPC l-mer. */
#define BUFLEN 256
static void minimizers_hpc_scalar(const uint8_t *seq, size_t seq_len, size_t k, uint32_t hash_bound, minsink_t *out)
{
    if (k > seq_len || k >= BUFLEN || k == 0) return; /* :117-125,:133 are errors; callers validate */
    uint32_t fh = 0, rh = 0;
    size_t j = 0, i = 0, prev_j = 0;
    uint32_t hbuf[BUFLEN], rcbuf[BUFLEN];
    size_t idxbuf[BUFLEN];
    memset(hbuf, 0, sizeof hbuf); memset(rcbuf, 0, sizeof rcbuf); memset(idxbuf, 0, sizeof idxbuf);
    uint8_t v, prev;
    while (i < k && j < seq_len) {                      /* :141-153 */
        v = seq[j];
        uint32_t hv = H_LOOKUP[v];
        hbuf[i] = hv; rcbuf[i] = 0; idxbuf[i] = j;
        fh ^= rol32(hv, (unsigned)(k - i - 1));
        i++;
        prev = v; prev_j = j;
        while (j < seq_len && seq[j] == prev) j++;
    }
    i -= 1; j = prev_j;                                 /* :161-163 */
    size_t cur = j;                                     /* current_idx_plus_k */
    for (;;) {                                          /* :165-177 */
        v = seq[j];
        uint32_t rcv = RC_LOOKUP[v];
        rcbuf[i] = rcv;
        rh ^= rol32(rcv, (unsigned)i);
        if (i == 0) break;
        i--;
        prev = v;
        while (j > 0 && seq[j] == prev) j--;
    }
    size_t buffer_pos = 0;
    int first = 1;
    for (;;) {                                          /* one iteration == one call to next() */
        uint8_t prevc = seq[cur], curc = prevc;
        uint32_t h_seqk, rc_seqk, hash;
        if (first) {                                    /* :208-236 */
            first = 0;
            for (;;) {
                cur++;
                if (cur >= seq_len) break;              /* the reference reads seq[cur] first (OOB by one, :214) */
                curc = seq[cur];
                if (curc != prevc) break;
            }
            if (cur >= seq_len) return;
            h_seqk = H_LOOKUP[curc]; rc_seqk = RC_LOOKUP[curc];
            size_t pos = (buffer_pos + k) % BUFLEN;
            hbuf[pos] = h_seqk; rcbuf[pos] = rc_seqk; idxbuf[pos] = cur;
            buffer_pos++;
            hash = rh < fh ? rh : fh;
            if (hash <= hash_bound) {
                sink_push(out, idxbuf[(buffer_pos + BUFLEN - 1) % BUFLEN], cur - 1, hash);
                continue;
            }
        } else {                                        /* :237-239 */
            h_seqk = hbuf[(buffer_pos + k - 1) % BUFLEN];
            rc_seqk = rcbuf[(buffer_pos + k - 1) % BUFLEN];
        }
        for (;;) {                                      /* :241-278 */
            uint32_t h_seqi = hbuf[(buffer_pos - 1) % BUFLEN], rc_seqi = rcbuf[(buffer_pos - 1) % BUFLEN];
            fh = rol32(fh, 1) ^ rol32(h_seqi, (unsigned)k) ^ h_seqk;
            rh = ror32(rh, 1) ^ ror32(rc_seqi, 1) ^ rol32(rc_seqk, (unsigned)k - 1);
            prevc = curc;
            for (;;) {
                cur++;
                if (cur >= seq_len) break;
                curc = seq[cur];
                if (curc != prevc) break;
            }
            if (cur >= seq_len) return;
            h_seqk = H_LOOKUP[curc]; rc_seqk = RC_LOOKUP[curc];
            size_t pos = (buffer_pos + k) % BUFLEN;
            hbuf[pos] = h_seqk; rcbuf[pos] = rc_seqk; idxbuf[pos] = cur;
            buffer_pos++;
            hash = rh < fh ? rh : fh;
            if (hash <= hash_bound) break;
        }
        sink_push(out, idxbuf[(buffer_pos + BUFLEN - 1) % BUFLEN], cur - 1, hash);
    }
}

/* ---------------------------------------------------------------- HashMode::Regular */
/* nthash32::NtHashIterator (external, absent): canonical ntHash1 with u32 state for every position
 * 0..N-l; threshold `hash <= hash_bound`, j = seq_pos, jend = j+l-1 at src/lib.rs:217-229. */
static void minimizers_regular(const uint8_t *seq, size_t n, size_t l, uint32_t hash_bound, minsink_t *out)
{
    if (l == 0 || l > n) return;
    uint32_t fh = 0, rh = 0;
    for (size_t i = 0; i < l; i++) {
        fh ^= rol32(H_LOOKUP[seq[i]], (unsigned)(l - 1 - i));
        rh ^= rol32(RC_LOOKUP[seq[i]], (unsigned)i);
    }
    for (size_t p = 0;; p++) {
        uint32_t h = rh < fh ? rh : fh;
        if (h <= hash_bound) sink_push(out, p, p + l - 1, h);
        if (p + l >= n) break;
        fh = rol32(fh, 1) ^ rol32(H_LOOKUP[seq[p]], (unsigned)l) ^ H_LOOKUP[seq[p + l]];
        rh = ror32(rh, 1) ^ ror32(RC_LOOKUP[seq[p]], 1) ^ rol32(RC_LOOKUP[seq[p + l]], (unsigned)l - 1);
    }
}

/* ---------------------------------------------------------------- HashMode::Simd (16-lane blocks) */
/* 31-bit rotates of src/nthash2_avx512_32.rs:186-215 (valid for inputs with bit 31 clear) */
static inline uint32_t srlv(uint32_t v, unsigned s) { return s >= 32 ? 0 : v >> s; }
static inline uint32_t sllv(uint32_t v, unsigned s) { return s >= 32 ? 0 : v << s; }
static inline uint32_t rorv31(uint32_t v, unsigned s) { return srlv(v, s) | (sllv(v, 32 - s) >> 1); }

/* bytes past the end of the slice are read by the reference (16-byte loads at arbitrary offsets,
 * src/nthash_avx512_32.rs:225-230); every lane that depends on them is masked later, so the oracle
 * substitutes 0 (-> code 4 -> seed 0). */
static inline uint8_t at(const uint8_t *s, size_t n, size_t i) { return i < n ? s[i] : 0; }

typedef struct { uint32_t f[16], r[16], h[16]; } lanes_t;

/* _mm512_NTC_epu32_initial: src/nthash_avx512_32.rs:281-341 / src/nthash2_avx512_32.rs:272-327 */
static void simd_initial(const uint8_t *s, size_t n, size_t k, int w31, lanes_t *L)
{
    int shift = w31 ? 33 : 0;
    unsigned ck = w31 ? (unsigned)(31 - (k % 31)) : (unsigned)(32 - (k % 32));
    for (int lane = 0; lane < 16; lane++) {
        uint32_t f = 0, r = 0;
        for (size_t i = 0; i < k; i++) {
            f = w31 ? rorv31(f, 30) : rol32(f, 1);
            f ^= lkf(at(s, n, i + lane), shift);
            uint32_t km = lkr(at(s, n, i + lane), shift);
            km = w31 ? rorv31(km, ck) : ror32(km, ck);
            r ^= km;
            r = w31 ? rorv31(r, 1) : ror32(r, 1);
        }
        L->f[lane] = f; L->r[lane] = r; L->h[lane] = f > r ? r : f;
    }
}
/* Hillis-Steele lane scan, `maskz_expand(0xfffe/0xfffc/0xfff0/0xff00, rot(x,1/2/4/8))` = shift lanes up by 1/2/4/8 */
static void lane_scan(uint32_t *x, int w31, int left)
{
    for (int d = 1; d < 16; d <<= 1) {
        uint32_t y[16];
        for (int lane = 0; lane < 16; lane++) {
            uint32_t v = x[lane];
            if (w31) y[lane] = left ? rorv31(v, (unsigned)(31 - d)) : rorv31(v, (unsigned)d);
            else y[lane] = left ? rol32(v, (unsigned)d) : ror32(v, (unsigned)d);
        }
        for (int lane = 15; lane >= d; lane--) x[lane] ^= y[lane - d];
    }
}
/* _mm512_NTC_epu32_sliding: src/nthash_avx512_32.rs:348-525 / src/nthash2_avx512_32.rs:331-478 */
static void simd_sliding(const uint8_t *s, size_t n, size_t off_out, size_t off_in, size_t k, int w31, lanes_t *L)
{
    int shift = w31 ? 33 : 0;
    unsigned ck = (unsigned)(31 - (k % 31));
    uint32_t kf[16], kr[16];
    for (int lane = 0; lane < 16; lane++) {
        uint32_t in_f = lkf(at(s, n, off_in + lane), shift), out_f = lkf(at(s, n, off_out + lane), shift);
        uint32_t in_r = lkr(at(s, n, off_in + lane), shift), out_r = lkr(at(s, n, off_out + lane), shift);
        if (w31) {
            kf[lane] = in_f ^ rorv31(out_f, ck);             /* nthash2:337-343 */
            kr[lane] = rorv31(in_r, ck) ^ out_r;             /* nthash2:403-411 */
        } else {
            kf[lane] = in_f ^ rol32(out_f, (unsigned)k);     /* nthash1:351-364 */
            kr[lane] = rol32(in_r, (unsigned)k - 1) ^ ror32(out_r, 1); /* nthash1:435-449 */
        }
    }
    lane_scan(kf, w31, 1);
    lane_scan(kr, w31, 0);
    uint32_t cf = L->f[15], cr = L->r[15];
    for (int lane = 0; lane < 16; lane++) {
        uint32_t f, r;
        if (w31) {
            f = rorv31(cf, (unsigned)(30 - lane)) ^ kf[lane];          /* shifts 30..15, nthash2:381-392 */
            r = rorv31(rorv31(cr, (unsigned)lane) ^ kr[lane], 1);      /* shifts 0..15 then ror31 1, nthash2:448-462 */
        } else {
            f = rol32(cf, (unsigned)(lane + 1)) ^ kf[lane];            /* shifts 1..16, nthash1:404-424 */
            r = ror32(cr, (unsigned)(lane + 1)) ^ kr[lane];            /* nthash1:487-504 */
        }
        L->f[lane] = f; L->r[lane] = r; L->h[lane] = f > r ? r : f;
    }
}
/* NtHashSIMDIterator::{new,next}: src/nthash_avx512_32.rs:32-77, 84-164 (and the nthash2 twin :37-154).
 * Emits (pos, hash) in order.  bound_in is the scalar bound handed to ::new. */
typedef void (*pos_emit_fn)(void *ctx, size_t pos, uint32_t hash);
static void simd_iterate(const uint8_t *s, size_t length, size_t k, uint32_t bound_in, int w31, pos_emit_fn emit, void *ctx)
{
    if (k == 0 || k > 31) return;                       /* assert!(k<=31), :33 */
    if (length < k) return;                             /* :87 */
    uint32_t bound = s2k_oracle_bound_simd(bound_in);   /* :47-48 */
    if (w31) bound /= 2;                                /* nthash2:54 */
    size_t sentinel = length - k + 1;
    lanes_t L;
    simd_initial(s, length, k, w31, &L);
    for (int lane = 0; lane < 16; lane++) {             /* buffered first block, :55-58 then :93-104 */
        if (!(L.h[lane] < bound)) continue;
        if ((size_t)lane >= sentinel) return;           /* `pos >= sentinel -> None`, :97-98.  The nthash2 file
                                                           has no such check (parity unpinned): the oracle masks p<S. */
        emit(ctx, (size_t)lane, L.h[lane]);
    }
    size_t i = 16;
    for (;;) {                                          /* :117-151 */
        if (i >= sentinel) return;
        simd_sliding(s, length, i - 1, i - 1 + k, k, w31, &L);
        size_t base = i;
        i += 16;
        uint16_t mask = 0;
        for (int lane = 0; lane < 16; lane++) if (L.h[lane] < bound) mask |= (uint16_t)(1u << lane);
        if (mask) {
            if (!w31) {
                if (i >= sentinel) mask &= (uint16_t)((1u << (sentinel % 16)) - 1); /* :134-138 (drops all 16 if S%16==0) */
            } else {
                /* nthash2:124-131 applies no mask; lanes >= S would be garbage. Deliberate choice: keep p<S only. */
                for (int lane = 0; lane < 16; lane++) if (base + lane >= sentinel) mask &= (uint16_t)~(1u << lane);
            }
        }
        for (int lane = 0; lane < 16; lane++) if (mask & (1u << lane)) emit(ctx, base + lane, L.h[lane]);
    }
}

typedef struct { minsink_t *out; size_t l; const uint32_t *hpc_pos; } simd_ctx_t;
static void emit_simd(void *c, size_t pos, uint32_t h)
{   /* src/lib.rs:202: jend = j + l - 1 */
    simd_ctx_t *x = (simd_ctx_t *)c;
    sink_push(x->out, pos, pos + x->l - 1, h);
}
static void emit_hpcsimd(void *c, size_t pos, uint32_t h)
{   /* src/nthash_hpc_simd.rs:61-68: (hpc_pos[p], hpc_pos[p+l-1], hash) */
    simd_ctx_t *x = (simd_ctx_t *)c;
    sink_push(x->out, x->hpc_pos[pos], x->hpc_pos[pos + x->l - 1], h);
}

/* ---------------------------------------------------------------- public: minimizers of one sequence */
/* Mirrors the dispatch in KminmersIterator::new (src/lib.rs:89-131): inner iterator only if seq.len() > l.
 * Returns the number of minimizers (may exceed cap; only cap are stored), or <0 on invalid parameters. */
long s2k_oracle_minimizers(const uint8_t *seq, size_t n, int l, double density, int mode, int variant,
                           uint64_t *start, uint64_t *end, uint32_t *hash, size_t cap)
{
    init_tables();
    if (l <= 0) return -1;
    if (variant == S2K_NT2_31 && !(mode == S2K_MODE_SIMD || mode == S2K_MODE_HPCSIMD)) return -1;
    if ((mode == S2K_MODE_SIMD || mode == S2K_MODE_HPCSIMD) && l > 31) return -2;
    if ((mode == S2K_MODE_HPC || mode == S2K_MODE_REGULAR) && l >= 256) return -2;
    minsink_t out = { start, end, hash, cap, 0 };
    uint32_t hash_bound = s2k_oracle_bound_scalar(density);  /* src/lib.rs:91 */
    if (!(n > (size_t)l)) return 0;                          /* src/lib.rs:97 */
    int w31 = variant == S2K_NT2_31;
    if (mode == S2K_MODE_HPC) minimizers_hpc_scalar(seq, n, (size_t)l, hash_bound, &out);
    else if (mode == S2K_MODE_REGULAR) minimizers_regular(seq, n, (size_t)l, hash_bound, &out);
    else if (mode == S2K_MODE_SIMD) {
        simd_ctx_t c = { &out, (size_t)l, NULL };
        simd_iterate(seq, n, (size_t)l, hash_bound, w31, emit_simd, &c);
    } else if (mode == S2K_MODE_HPCSIMD) {
        uint8_t *hs = (uint8_t *)malloc(n ? n : 1);
        uint32_t *hp = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
        size_t m = s2k_oracle_encode_rle_simd(seq, n, hs, hp);   /* src/nthash_hpc_simd.rs:36 */
        simd_ctx_t c = { &out, (size_t)l, hp };
        simd_iterate(hs, m, (size_t)l, hash_bound, w31, emit_hpcsimd, &c);
        free(hs); free(hp);
    } else return -1;
    return (long)out.n;
}

/* ---------------------------------------------------------------- window stage */
/* MixHash for u32, src/lib.rs:157-169 */
static inline uint64_t mix32(uint32_t h) { uint64_t x = h; x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; }

/* KminmersIterator::next, src/lib.rs:181-269, in its rolling form (:235-249), fed by a minimizer list.
 * Returns number of k-min-mers. */
long s2k_oracle_windows(const uint64_t *mstart, const uint64_t *mend, const uint32_t *mhash, size_t nmin, int k,
                        uint64_t *hash, uint64_t *start, uint64_t *end, uint64_t *offset, uint8_t *rev, size_t cap)
{
    if (k <= 0) return -1;
    size_t K = (size_t)k, count = 0;
    uint64_t f = 0, r = 0;
    uint64_t *sk = (uint64_t *)malloc((nmin ? nmin : 1) * sizeof(uint64_t));
    for (size_t len = 1; len <= nmin; len++) {
        uint64_t h = mix32(mhash[len - 1]);
        sk[len - 1] = h;
        if (len >= K) {
            if (len == K) {
                f ^= rol64(h, (unsigned)(K - 1 - (len - 1)));
                r ^= rol64(h, (unsigned)(len - 1));
            } else {
                f = rol64(f, 1) ^ h ^ rol64(sk[count - 1], (unsigned)K);
                r = ror64(r, 1) ^ rol64(h, (unsigned)(K - 1)) ^ ror64(sk[count - 1], 1);
            }
            if (count < cap) {
                if (hash) hash[count] = f < r ? f : r;
                if (rev) rev[count] = r < f;
                if (start) start[count] = mstart[count];
                if (end) end[count] = mend[len - 1];
                if (offset) offset[count] = count;
            }
            count++;
        } else {
            f ^= rol64(h, (unsigned)(K - 1 - (len - 1)));
            r ^= rol64(h, (unsigned)(len - 1));
        }
    }
    free(sk);
    return (long)count;
}

/* KminmersIterator::new + collect for one sequence.  Returns #k-min-mers (<0 invalid params). */
long s2k_oracle_kminmers(const uint8_t *seq, size_t n, int l, int k, double density, int mode, int variant,
                         uint64_t *hash, uint64_t *start, uint64_t *end, uint64_t *offset, uint8_t *rev, size_t cap,
                         long *n_minimizers)
{
    if (k <= 0) return -1;
    size_t mcap = n ? n : 1;
    uint64_t *ms = (uint64_t *)malloc(mcap * sizeof(uint64_t)), *me = (uint64_t *)malloc(mcap * sizeof(uint64_t));
    uint32_t *mh = (uint32_t *)malloc(mcap * sizeof(uint32_t));
    long nm = s2k_oracle_minimizers(seq, n, l, density, mode, variant, ms, me, mh, mcap);
    long res = nm;
    if (nm >= 0) res = s2k_oracle_windows(ms, me, mh, (size_t)nm, k, hash, start, end, offset, rev, cap);
    if (n_minimizers) *n_minimizers = nm;
    free(ms); free(me); free(mh);
    return res;
}

/* ---------------------------------------------------------------- closed form (independent cross-check) */
/* st[] = run starts, c[] = bytes at run starts; fh/rh by direct XOR over the l-mer.
 * width = 32 (nt1), 31 (nt2) or 64 (H=u64, used only to reproduce the reference's u64 golden vectors:
 * tests/main.rs:18-39 and src/old/nthash_hpc.rs.opt4:90-97).  strict: 0 -> `<=`, 1 -> `<`.
 * width = 16: the crate built with `pub type H = u16` (src/lib.rs:29), 16-bit ntHash1 state of NtHashHPCIterator
 * (seeds `as u16`, src/nthash_hpc.rs:30-49; u16 rotations); width = 3216: that build in mode Regular, where the
 * 32-bit nthash32 hash is truncated, `x as H` (src/lib.rs:224), before `hash <= bound`.  Parity unpinned for both:
 * the reference holds no vector for H = u16.
 * Output hashes are u64 so that width 64 fits. */
static inline uint64_t rolw(uint64_t x, unsigned r, int w)
{
    if (w == 64) return rol64(x, r);
    r %= (unsigned)w;
    uint64_t m = (w == 32) ? 0xffffffffULL : (w == 16) ? 0xffffULL : 0x7fffffffULL;
    return r ? (((x << r) | (x >> (w - r))) & m) : x;
}
long s2k_oracle_closed_minimizers(const uint8_t *seq, size_t n, int l, int hpc, int simd_tables, int width,
                                  uint64_t bound, int strict, long last_rule, int end_rule,
                                  uint64_t *start, uint64_t *end, uint64_t *hash, size_t cap)
{
    /* last_rule: 0 = all S l-mers; 1 = drop final l-mer (Hpc scalar); 2 = drop last 16 if S>16 && S%16==0 (nt1 simd)
     * end_rule : 0 = st[p+l-1]+0 in raw space == p+l-1 when !hpc; 1 = st[p+l]-1 (run end, scalar Hpc); */
    init_tables();
    if (!(n > (size_t)l) || l <= 0) return 0;
    size_t *st = (size_t *)malloc(n * sizeof(size_t));
    size_t M = 0;
    for (size_t i = 0; i < n; i++) if (!hpc || i == 0 || seq[i] != seq[i - 1]) st[M++] = i;
    long cnt = 0;
    if (M >= (size_t)l) {
        size_t S = M - (size_t)l + 1, last = S;
        if (last_rule == 1) last = S - 1;
        else if (last_rule == 2 && S > 16 && S % 16 == 0) last = S - 16;
        int shift = width == 31 ? 33 : 0;
        const int trunc16 = width == 3216;
        if (trunc16) width = 32;
        for (size_t p = 0; p < last; p++) {
            uint64_t fh = 0, rh = 0;
            for (int i = 0; i < l; i++) {
                uint8_t b = seq[st[p + i]];
                uint64_t hv, rv;
                if (simd_tables) {
                    int c = ckx(b);
                    hv = c < 4 ? SEED64[c] >> shift : 0; rv = c < 4 ? SEED64[3 - c] >> shift : 0;
                    if (width == 32) { hv &= 0xffffffffULL; rv &= 0xffffffffULL; }
                } else if (width == 64) { hv = H_LOOKUP64[b]; rv = RC_LOOKUP64[b]; }
                else { hv = H_LOOKUP[b]; rv = RC_LOOKUP[b]; }
                if (width == 16) { hv &= 0xffffULL; rv &= 0xffffULL; }
                fh ^= rolw(hv, (unsigned)(l - 1 - i), width);
                rh ^= rolw(rv, (unsigned)i, width);
            }
            uint64_t h = fh < rh ? fh : rh;
            if (trunc16) h &= 0xffffULL;
            int sel = strict ? (h < bound) : (h <= bound);
            if (!sel) continue;
            if ((size_t)cnt < cap) {
                if (start) start[cnt] = st[p];
                if (end) end[cnt] = end_rule == 1 ? st[p + l] - 1 : st[p + l - 1];
                if (hash) hash[cnt] = h;
            }
            cnt++;
        }
    }
    free(st);
    return cnt;
}
/* Closed-form window stage (Appendix A.2).  mix_u32: 0 = identity (MixHash<u64>), 1 = xorshift (MixHash<u32>),
 * 2 = MixHash<u16> (src/lib.rs:142-155; wrapping multiplies as in a release build). */
static inline uint64_t mix16(uint16_t h)
{
    uint64_t x = h;
    x ^= rol64(x, 33); x *= 0xff51afd7ed558ccdULL;
    x ^= rol64(x, 33); x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= rol64(x, 33);
    return x;
}
long s2k_oracle_closed_windows(const uint64_t *mhash, size_t nmin, int k, int mix_u32,
                               uint64_t *hash, uint8_t *rev, size_t cap)
{
    if (k <= 0) return -1;
    long cnt = 0;
    for (size_t c = 0; c + (size_t)k <= nmin; c++) {
        uint64_t f = 0, r = 0;
        for (int t = 0; t < k; t++) {
            uint64_t m = mix_u32 == 2 ? mix16((uint16_t)mhash[c + t]) : mix_u32 ? mix32((uint32_t)mhash[c + t]) : mhash[c + t];
            f ^= rol64(m, (unsigned)(k - 1 - t));
            r ^= rol64(m, (unsigned)t);
        }
        if ((size_t)cnt < cap) { if (hash) hash[cnt] = f < r ? f : r; if (rev) rev[cnt] = r < f; }
        cnt++;
    }
    return cnt;
}

/* ---------------------------------------------------------------- synthetic generator */
static inline uint64_t s2k_word(uint64_t seed, uint64_t j)
{
    uint64_t z = seed + (j + 1) * 0x9E3779B97F4A7C15ULL;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
uint64_t s2k_oracle_synth_word(uint64_t seed, uint64_t j) { return s2k_word(seed, j); }
void s2k_oracle_synth(uint64_t seed, uint64_t first, uint64_t count, uint8_t *out)
{
    for (uint64_t t = 0; t < count; t++) {
        uint64_t i = first + t;
        out[t] = (uint8_t)"ACGT"[(s2k_word(seed, i >> 5) >> (2 * (i & 31))) & 3];
    }
}

/* ---------------------------------------------------------------- batch + threads (cpu_baseline / tests) */
/* Mirrors src/main.rs:65-79: one iterator per record, records split over nb_threads workers.
 * Fills per-read k-min-mer counts (km_cnt[n]) and an order-sensitive digest per read when requested. */
typedef struct {
    const uint8_t *bases; const uint64_t *seq_off; uint64_t r0, r1;
    int l, k, mode, variant; double density;
    uint64_t *km_cnt, *digest, *min_cnt;
    uint64_t total_km, total_min;
} job_t;

static uint64_t fold(uint64_t acc, uint64_t v) { acc ^= v; acc *= 0x100000001b3ULL; acc ^= acc >> 29; return acc; }

static void *job_run(void *p)
{
    job_t *j = (job_t *)p;
    size_t cap = 0;
    uint64_t *h = NULL, *s = NULL, *e = NULL, *o = NULL; uint8_t *rv = NULL;
    for (uint64_t r = j->r0; r < j->r1; r++) {
        size_t n = (size_t)(j->seq_off[r + 1] - j->seq_off[r]);
        if (j->digest && n > cap) {
            cap = n * 2;
            h = (uint64_t *)realloc(h, cap * 8); s = (uint64_t *)realloc(s, cap * 8);
            e = (uint64_t *)realloc(e, cap * 8); o = (uint64_t *)realloc(o, cap * 8);
            rv = (uint8_t *)realloc(rv, cap);
        }
        long nm = 0;
        long c = s2k_oracle_kminmers(j->bases + j->seq_off[r], n, j->l, j->k, j->density, j->mode, j->variant,
                                     j->digest ? h : NULL, j->digest ? s : NULL, j->digest ? e : NULL,
                                     j->digest ? o : NULL, j->digest ? rv : NULL, j->digest ? cap : 0, &nm);
        if (c < 0) c = 0;
        if (nm < 0) nm = 0;
        if (j->km_cnt) j->km_cnt[r] = (uint64_t)c;
        if (j->min_cnt) j->min_cnt[r] = (uint64_t)nm;
        if (j->digest) {
            uint64_t d = 0xcbf29ce484222325ULL;
            for (long t = 0; t < c; t++) {
                d = fold(d, h[t]); d = fold(d, s[t]); d = fold(d, e[t]); d = fold(d, rv[t]);
            }
            j->digest[r] = d;
        }
        j->total_km += (uint64_t)c; j->total_min += (uint64_t)nm;
    }
    free(h); free(s); free(e); free(o); free(rv);
    return NULL;
}

long s2k_oracle_batch(const uint8_t *bases, const uint64_t *seq_off, uint64_t n_seqs, int l, int k, double density,
                      int mode, int variant, int nb_threads, uint64_t *km_cnt, uint64_t *min_cnt, uint64_t *digest,
                      uint64_t *total_min)
{
    init_tables();
    if (nb_threads < 1) nb_threads = 1;
    if ((uint64_t)nb_threads > n_seqs && n_seqs > 0) nb_threads = (int)n_seqs;
    job_t *jobs = (job_t *)calloc((size_t)nb_threads, sizeof(job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)nb_threads, sizeof(pthread_t));
    uint64_t total_bases = n_seqs ? seq_off[n_seqs] - seq_off[0] : 0;
    uint64_t r = 0;
    for (int t = 0; t < nb_threads; t++) {          /* balance by bases */
        uint64_t target = seq_off[0] + (total_bases / (uint64_t)nb_threads) * (uint64_t)(t + 1);
        uint64_t r1 = r;
        if (t == nb_threads - 1) r1 = n_seqs;
        else while (r1 < n_seqs && seq_off[r1 + 1] <= target) r1++;
        jobs[t] = (job_t){ bases, seq_off, r, r1, l, k, mode, variant, density, km_cnt, digest, min_cnt, 0, 0 };
        r = r1;
    }
    for (int t = 0; t < nb_threads; t++) pthread_create(&th[t], NULL, job_run, &jobs[t]);
    uint64_t tot = 0, totm = 0;
    for (int t = 0; t < nb_threads; t++) { pthread_join(th[t], NULL); tot += jobs[t].total_km; totm += jobs[t].total_min; }
    if (total_min) *total_min = totm;
    free(jobs); free(th);
    return (long)tot;
}

/* Per-read digest of an item stream laid out as the product's SoA result (u32 coordinates), folded
 * exactly like job_run() above, so tests can compare full tuples of millions of items cheaply. */
void s2k_oracle_digest_items(const uint64_t *hash, const uint32_t *start, const uint32_t *end, const uint8_t *rev,
                             const uint64_t *km_off, uint64_t n_seqs, uint64_t *digest)
{
    for (uint64_t r = 0; r < n_seqs; r++) {
        uint64_t d = 0xcbf29ce484222325ULL;
        for (uint64_t t = km_off[r]; t < km_off[r + 1]; t++) {
            d = fold(d, hash[t]); d = fold(d, start[t]); d = fold(d, end[t]); d = fold(d, rev[t]);
        }
        digest[r] = d;
    }
}
