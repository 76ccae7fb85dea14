// Microbenchmark: how fast can the host pack ASCII bases to 2 bits (with detection of non-ACGT bytes)?
// gcc -O3 -march=native -pthread tools/bench_pack.c -o /tmp/bench_pack && /tmp/bench_pack [threads] [MiB]
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { const uint8_t *src; uint8_t *dst; size_t n; uint64_t bad; } job_t;

__attribute__((target("avx512f,avx512bw,avx512vl")))
static void *pack_job(void *p)
{
    job_t *j = (job_t *)p;
    const __m512i three = _mm512_set1_epi8(3);
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i m1 = _mm512_set1_epi16(0x0401), m2 = _mm512_set1_epi32(0x00100001);
    uint64_t bad = 0;
    size_t i = 0;
    for (; i + 64 <= j->n; i += 64) {
        const __m512i v = _mm512_loadu_si512(j->src + i);
        const __m512i c = _mm512_and_si512(_mm512_srli_epi16(v, 1), three);          // (b >> 1) & 3 : A0 C1 T2 G3
        const __mmask64 ne = _mm512_cmpneq_epi8_mask(_mm512_shuffle_epi8(lut, c), v);
        bad += (uint64_t)__builtin_popcountll(ne);
        const __m512i p16 = _mm512_maddubs_epi16(c, m1);                             // c0 + 4 c1 per 16-bit
        const __m512i p32 = _mm512_madd_epi16(p16, m2);                              // + 16 (c2 + 4 c3)
        _mm_storeu_si128((__m128i *)(j->dst + (i >> 2)), _mm512_cvtepi32_epi8(p32));
    }
    j->bad = bad;
    return NULL;
}

int main(int argc, char **argv)
{
    int T = argc > 1 ? atoi(argv[1]) : 16;
    size_t n = (size_t)(argc > 2 ? atoi(argv[2]) : 2048) << 20;
    uint8_t *src = aligned_alloc(64, n), *dst = aligned_alloc(64, n / 4 + 64);
    for (size_t i = 0; i < n; i++) src[i] = "ACGT"[(i * 2654435761u >> 13) & 3];
    memset(dst, 0, n / 4 + 64);
    pthread_t th[256]; job_t jobs[256];
    for (int rep = 0; rep < 4; rep++) {
        struct timespec a, b;
        clock_gettime(CLOCK_MONOTONIC, &a);
        for (int t = 0; t < T; t++) {
            size_t lo = (n / T * t) & ~(size_t)63, hi = t == T - 1 ? n : (n / T * (t + 1)) & ~(size_t)63;
            jobs[t] = (job_t){src + lo, dst + lo / 4, hi - lo, 0};
            pthread_create(&th[t], NULL, pack_job, &jobs[t]);
        }
        uint64_t bad = 0;
        for (int t = 0; t < T; t++) { pthread_join(th[t], NULL); bad += jobs[t].bad; }
        clock_gettime(CLOCK_MONOTONIC, &b);
        double s = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
        printf("threads %d: %.1f GB/s (bad=%llu)\n", T, n / s / 1e9, (unsigned long long)bad);
    }
    return 0;
}
