/*
 * s2k_cpu_avx512.c -- CPU BASELINE for bench.py: an AVX-512 restatement of the reference's HashMode::HpcSimd /
 * HashMode::Simd path, threaded like src/main.rs:65-79 (one iterator per record, records spread over host threads,
 * items are produced and counted, not stored).
 *
 * TEST/BENCH INFRASTRUCTURE ONLY (like the rest of oracle/): the product never links or calls it.  It exists because
 * the reference (Rust nightly + un-vendored git crates) cannot be built in this image; it is labelled "port" wherever
 * reported.  It follows the reference's vector algorithm step by step:
 *   encode_rle_simd          src/hpc.rs:44-147        16 bytes per step, mask = byte != previous byte, compress
 *   base -> 0..4 code        src/nthash_avx512_32.rs:178-193  (pshufb on the low nibble)
 *   seed permutes            src/nthash_avx512_32.rs:242-277
 *   first 16 hashes          src/nthash_avx512_32.rs:281-341
 *   sliding, 16 positions    src/nthash_avx512_32.rs:348-525  (Hillis-Steele lane scan with maskz_expand, lane-15 carry)
 *   threshold + compress     src/nthash_avx512_32.rs:47-58,125-141 (f32-rederived bound, tail mask)
 *   position remap           src/nthash_hpc_simd.rs:61-68
 *   window stage             src/lib.rs:157-169,231-261
 * Results are checked bit-exactly against the scalar oracle by tests/test_oracle_kats.py::test_avx512_baseline.
 */
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

uint32_t s2k_oracle_bound_scalar(double density);
uint32_t s2k_oracle_bound_simd(uint32_t bound_scalar);

#define TGT __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq")))

int s2k_cpu_has_avx512(void)
{
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl");
}

/* src/hpc.rs:74-136 */
TGT static size_t rle16(const uint8_t *s, size_t len, uint8_t *res, uint32_t *pos)
{
    const size_t width = 16, end_idx = len / width;
    size_t m = 0;
    __m512i positions = _mm512_set_epi32(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m512i sixteen = _mm512_set1_epi32(16);
    for (size_t i = 0; i <= end_idx; i++) {
        size_t lanes = width;
        if (i == end_idx) { lanes = len & (width - 1); if (!lanes) break; }
        const __m128i v = _mm_loadu_si128((const __m128i *)(s + i * width));       /* buffers are padded by the caller */
        const __m128i sh = _mm_slli_si128(v, 1);
        __mmask16 mask = (__mmask16)(~_mm_cmpeq_epi8_mask(v, sh)) & 0xFFFE;
        if (i > 0) mask |= (__mmask16)(s[i * width] != s[i * width - 1]); else mask |= 1;
        if (lanes < width) mask &= (__mmask16)((1u << lanes) - 1);
        const __m512i wide = _mm512_cvtepi8_epi32(v);
        const __m128i packed = _mm512_cvtepi32_epi8(_mm512_maskz_compress_epi32(mask, wide));
        const unsigned cnt = (unsigned)__builtin_popcount(mask);
        _mm_mask_storeu_epi8(res + m, (__mmask16)((1u << cnt) - 1), packed);
        _mm512_mask_compressstoreu_epi32(pos + m, mask, positions);
        positions = _mm512_add_epi32(positions, sixteen);
        m += cnt;
    }
    return m;
}

static const uint64_t SEED[4] = {0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL, 0x20323ed082572324ULL, 0x295549f54be24456ULL};

TGT static inline __m512i lkx(const uint8_t *p)
{
    const __m128i table = _mm_set_epi8(4, 4, 4, 4, 4, 4, 4, 4, 2, 4, 4, 3, 1, 4, 0, 4);
    const __m128i v = _mm_and_si128(_mm_loadu_si128((const __m128i *)p), _mm_set1_epi8(0x0f));
    return _mm512_cvtepu8_epi32(_mm_shuffle_epi8(table, v));
}
TGT static inline __m512i lkf(const uint8_t *p)
{
    const __m512i seed = _mm512_set_epi32(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, (int)(uint32_t)SEED[3], (int)(uint32_t)SEED[2],
                                          (int)(uint32_t)SEED[1], (int)(uint32_t)SEED[0]);
    return _mm512_permutexvar_epi32(lkx(p), seed);
}
TGT static inline __m512i lkr(const uint8_t *p)
{
    const __m512i seed = _mm512_set_epi32(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, (int)(uint32_t)SEED[0], (int)(uint32_t)SEED[1],
                                          (int)(uint32_t)SEED[2], (int)(uint32_t)SEED[3]);
    return _mm512_permutexvar_epi32(lkx(p), seed);
}
#define SCAN(x, ROT)                                                                  \
    x = _mm512_xor_epi32(x, _mm512_maskz_expand_epi32(0xfffe, ROT(x, 1)));            \
    x = _mm512_xor_epi32(x, _mm512_maskz_expand_epi32(0xfffc, ROT(x, 2)));            \
    x = _mm512_xor_epi32(x, _mm512_maskz_expand_epi32(0xfff0, ROT(x, 4)));            \
    x = _mm512_xor_epi32(x, _mm512_maskz_expand_epi32(0xff00, ROT(x, 8)));

typedef struct { uint64_t f, r; uint64_t count, n_min; uint64_t *ring; size_t k; uint64_t digest; int want_digest;
                 uint64_t *sk; size_t sk_cap; size_t *spos; } win_t;

static inline uint64_t rol64(uint64_t x, unsigned r) { r &= 63; return r ? (x << r) | (x >> (64 - r)) : x; }
static inline uint64_t ror64(uint64_t x, unsigned r) { r &= 63; return r ? (x >> r) | (x << (64 - r)) : x; }
static inline uint64_t fold(uint64_t acc, uint64_t v) { acc ^= v; acc *= 0x100000001b3ULL; acc ^= acc >> 29; return acc; }

/* KminmersIterator::next for one minimizer (src/lib.rs:231-261) */
static inline void win_push(win_t *w, size_t j, size_t jend, uint32_t hash)
{
    uint64_t x = hash; x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    const size_t len = (size_t)w->n_min + 1, K = w->k;
    if (len > w->sk_cap) { w->sk_cap = w->sk_cap * 2 + 64; w->sk = (uint64_t *)realloc(w->sk, w->sk_cap * 8); w->spos = (size_t *)realloc(w->spos, w->sk_cap * sizeof(size_t)); }
    w->sk[len - 1] = x; w->spos[len - 1] = j;
    w->n_min = len;
    if (len >= K) {
        if (len == K) { w->f ^= rol64(x, (unsigned)(K - 1 - (len - 1))); w->r ^= rol64(x, (unsigned)(len - 1)); }
        else {
            w->f = rol64(w->f, 1) ^ x ^ rol64(w->sk[w->count - 1], (unsigned)K);
            w->r = ror64(w->r, 1) ^ rol64(x, (unsigned)(K - 1)) ^ ror64(w->sk[w->count - 1], 1);
        }
        if (w->want_digest) {
            const uint64_t h = w->f < w->r ? w->f : w->r;
            w->digest = fold(fold(fold(fold(w->digest, h), w->spos[w->count]), jend), (uint64_t)(w->r < w->f));
        }
        w->count++;
    } else { w->f ^= rol64(x, (unsigned)(K - 1 - (len - 1))); w->r ^= rol64(x, (unsigned)(len - 1)); }
}

/* NtHashSIMDIterator (src/nthash_avx512_32.rs:32-164) driving the window stage; hpc_pos == NULL: Simd mode */
TGT static void simd_kminmers(const uint8_t *s, size_t length, size_t k, uint32_t bound_in, const uint32_t *hpc_pos, win_t *w)
{
    if (length < k || k == 0 || k > 31) return;
    const uint32_t bound = s2k_oracle_bound_simd(bound_in);
    const __m512i vbound = _mm512_set1_epi32((int)bound);
    const size_t sentinel = length - k + 1;
    __m512i fh = _mm512_setzero_si512(), rh = _mm512_setzero_si512();
    const __m512i ck = _mm512_set1_epi32((int)(32 - (k % 32)));
    for (size_t i = 0; i < k; i++) {                                   /* :281-341 */
        fh = _mm512_xor_epi32(_mm512_rol_epi32(fh, 1), lkf(s + i));
        rh = _mm512_ror_epi32(_mm512_xor_epi32(rh, _mm512_rorv_epi32(lkr(s + i), ck)), 1);
    }
    __m512i h = _mm512_min_epu32(fh, rh);
    uint32_t hb[16], pb[16];
    __m512i positions = _mm512_set_epi32(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m512i sixteen = _mm512_set1_epi32(16);
    const __m512i vk = _mm512_set1_epi32((int)k), vkm = _mm512_set1_epi32((int)k - 1);
    const __m512i sh1 = _mm512_set_epi32(16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1);
    const __m512i l15 = _mm512_set1_epi32(15);
    __mmask16 mask = _mm512_cmplt_epu32_mask(h, vbound);
    size_t i = 16;
    for (;;) {
        unsigned n = (unsigned)__builtin_popcount(mask);
        if (n) {
            _mm512_mask_compressstoreu_epi32(hb, mask, h);
            _mm512_mask_compressstoreu_epi32(pb, mask, positions);
            for (unsigned t = 0; t < n; t++) {
                const size_t p = pb[t];
                if (p >= sentinel) return;                              /* :97-98 */
                if (hpc_pos) win_push(w, hpc_pos[p], hpc_pos[p + k - 1], hb[t]);   /* nthash_hpc_simd.rs:64 */
                else win_push(w, p, p + k - 1, hb[t]);                               /* lib.rs:202 */
            }
        }
        positions = _mm512_add_epi32(positions, sixteen);
        if (i >= sentinel) return;                                      /* :119-123 */
        {   /* :348-509 */
            const uint8_t *pin = s + i - 1 + k, *pout = s + i - 1;
            __m512i kf = _mm512_xor_epi32(lkf(pin), _mm512_rolv_epi32(lkf(pout), vk));
            SCAN(kf, _mm512_rol_epi32)
            fh = _mm512_xor_epi32(_mm512_rolv_epi32(_mm512_permutexvar_epi32(l15, fh), sh1), kf);
            __m512i kr = _mm512_xor_epi32(_mm512_rolv_epi32(lkr(pin), vkm), _mm512_ror_epi32(lkr(pout), 1));
            SCAN(kr, _mm512_ror_epi32)
            rh = _mm512_xor_epi32(_mm512_rorv_epi32(_mm512_permutexvar_epi32(l15, rh), sh1), kr);
            h = _mm512_min_epu32(fh, rh);
        }
        i += 16;
        mask = _mm512_cmplt_epu32_mask(h, vbound);
        if (mask && i >= sentinel) mask &= (__mmask16)((1u << (sentinel % 16)) - 1);    /* :134-138 */
    }
}

typedef struct {
    const uint8_t *bases; const uint64_t *seq_off; uint64_t r0, r1;
    int l, k, hpc; double density; uint64_t *km_cnt, *digest; uint64_t total;
} job_t;

TGT static void *job_run(void *p)
{
    job_t *j = (job_t *)p;
    size_t cap = 0;
    uint8_t *buf = NULL, *hs = NULL; uint32_t *hp = NULL;
    win_t w; memset(&w, 0, sizeof w);
    const uint32_t bound = s2k_oracle_bound_scalar(j->density);
    for (uint64_t r = j->r0; r < j->r1; r++) {
        const size_t n = (size_t)(j->seq_off[r + 1] - j->seq_off[r]);
        if (n + 128 > cap) { cap = 2 * n + 256; buf = realloc(buf, cap); hs = realloc(hs, cap); hp = realloc(hp, cap * 4); }
        w.f = w.r = 0; w.count = w.n_min = 0; w.k = (size_t)j->k; w.digest = 0xcbf29ce484222325ULL; w.want_digest = j->digest != NULL;
        if (n > (size_t)j->l) {                                           /* src/lib.rs:97 */
            memcpy(buf, j->bases + j->seq_off[r], n); memset(buf + n, 0, 64);   /* the reference over-reads; pad instead */
            if (j->hpc) {
                const size_t m = rle16(buf, n, hs, hp);
                memset(hs + m, 0, 64);
                simd_kminmers(hs, m, (size_t)j->l, bound, hp, &w);
            } else simd_kminmers(buf, n, (size_t)j->l, bound, NULL, &w);
        }
        if (j->km_cnt) j->km_cnt[r] = w.count;
        if (j->digest) j->digest[r] = w.digest;
        j->total += w.count;
    }
    free(buf); free(hs); free(hp); free(w.sk); free(w.spos);
    return NULL;
}

/* mode: 2 = Simd, 3 = HpcSimd.  Returns total k-min-mers, or -1 if the CPU lacks AVX-512 / bad mode. */
long s2k_cpu_avx512_batch(const uint8_t *bases, const uint64_t *seq_off, uint64_t n_seqs, int l, int k, double density,
                          int mode, int nb_threads, uint64_t *km_cnt, uint64_t *digest)
{
    if (!s2k_cpu_has_avx512() || (mode != 2 && mode != 3) || l < 1 || l > 31 || k < 1) return -1;
    if (nb_threads < 1) nb_threads = 1;
    if ((uint64_t)nb_threads > n_seqs && n_seqs > 0) nb_threads = (int)n_seqs;
    job_t *jobs = calloc((size_t)nb_threads, sizeof(job_t));
    pthread_t *th = calloc((size_t)nb_threads, sizeof(pthread_t));
    const uint64_t total_bases = n_seqs ? seq_off[n_seqs] : 0;
    uint64_t r = 0;
    for (int t = 0; t < nb_threads; t++) {
        const uint64_t target = (total_bases / (uint64_t)nb_threads) * (uint64_t)(t + 1);
        uint64_t r1 = r;
        if (t == nb_threads - 1) r1 = n_seqs; else while (r1 < n_seqs && seq_off[r1 + 1] <= target) r1++;
        jobs[t] = (job_t){bases, seq_off, r, r1, l, k, mode == 3, density, km_cnt, digest, 0};
        r = r1;
    }
    for (int t = 0; t < nb_threads; t++) pthread_create(&th[t], NULL, job_run, &jobs[t]);
    uint64_t tot = 0;
    for (int t = 0; t < nb_threads; t++) { pthread_join(th[t], NULL); tot += jobs[t].total; }
    free(jobs); free(th);
    return (long)tot;
}
