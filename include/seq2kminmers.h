/*
 * seq2kminmers.h -- C ABI of the B200-native sequence -> k-min-mer path.
 *
 * This is the drop-in boundary for rchikhi/rust-seq2kminmers' hot path.  Each entry point
 * cites the reference interface it replaces (file:line into the reference crate).  The
 * reference exposes a per-sequence iterator,
 *
 *     KminmersIterator::new(seq:&[u8], l, k, density:f64, mode:HashMode) -> io::Result<Self>   src/lib.rs:89
 *     impl Iterator for KminmersIterator { type Item = KminmerHash; fn next() }                src/lib.rs:179-270
 *     KminmerHash { hash:u64, start:usize, end:usize, offset:usize, rev:bool }                 src/kminmer.rs:128-135
 *
 * and the replacement is a BATCHED form of it: many sequences per call, results returned as
 * structure-of-arrays in the same per-sequence order the iterator would yield them.
 * `offset` (index of the k-min-mer inside its sequence) is implicit: item i of sequence r has
 * offset = i - km_off[r].
 *
 * Plain pointers and sizes only; no C++ or torch types.  Nothing unwinds across this ABI:
 * every call returns an int status (0 = ok, negative = error, see S2K_ERR_*).
 *
 * Precedent inside the reference for an opaque-handle C ABI: src/nthash_c.rs:7-29.
 */
#ifndef SEQ2KMINMERS_H
#define SEQ2KMINMERS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2K_ABI_VERSION 1

/* HashMode, src/lib.rs:21-27 (same discriminant order). */
typedef enum s2k_hash_mode {
    S2K_MODE_REGULAR = 0,  /* scalar ntHash1-32 in original space (nthash32 crate, src/lib.rs:108,217-229) */
    S2K_MODE_HPC     = 1,  /* scalar fused HPC + ntHash1-32 (src/nthash_hpc.rs) */
    S2K_MODE_SIMD    = 2,  /* AVX-512 ntHash semantics in original space (src/nthash_avx512_32.rs) */
    S2K_MODE_HPCSIMD = 3   /* AVX-512 HPC then AVX-512 ntHash (src/hpc.rs + src/nthash_hpc_simd.rs) */
} s2k_hash_mode;

/* Which rolling hash.  NT1_32 is what src/lib.rs compiles; NT2_31 is the 31-bit hybrid of
 * src/nthash2_avx512_32.rs (commented out of the reference build, src/lib.rs:8-9) and is only
 * defined for the SIMD modes, whose iterator it would replace. */
typedef enum s2k_hash_variant {
    S2K_HASH_NT1_32 = 0,
    S2K_HASH_NT2_31 = 1,
    S2K_HASH_NT1_64 = 2   /* the crate built with `pub type H = u64` (src/lib.rs:30-32): 64-bit ntHash1 seeds
                             (src/nthash_hpc.rs:30-49), bound = (density * u64::MAX as f64) as u64 (src/lib.rs:91),
                             `hash <= bound`, MixHash<u64> = identity (src/lib.rs:171-177).  Modes Regular and Hpc only
                             (the SIMD iterators are 32-bit).  Pinned by the reference's own golden vector for that
                             build, tests/main.rs:18-39.  result.minimizers[i].hash then holds the LOW half of the
                             minimizer's hash; s2k_last_minimizer_hash_hi gives the high halves. */
    ,
    S2K_HASH_NT1_16 = 3   /* the crate built with `pub type H = u16` (src/lib.rs:29): bound = (density * u16::MAX as f64)
                             as u16 (src/lib.rs:91), `hash <= bound`, MixHash<u16> = the murmur-style mix of
                             src/lib.rs:142-155.  Mode Hpc: ntHash1 on a 16-bit state, seeds `as u16`
                             (src/nthash_hpc.rs:30-49).  Mode Regular: the nthash32 hash truncated, `x as H`
                             (src/lib.rs:224).  Modes Regular and Hpc only (the SIMD iterators take a u32 bound).
                             Parity unpinned: the reference holds no vector for this build. */
} s2k_hash_variant;

/* Status codes.  The reference panics (unwrap/assert) where these are returned:
 *   S2K_ERR_L_TOO_BIG   assert!(k<=31) src/nthash_avx512_32.rs:33; KSizeTooBig / assert!(k<256) src/nthash_hpc.rs:123-133
 *   S2K_ERR_BAD_PARAM   l==0 or k==0 (usize underflow panics in src/lib.rs:238-247), unknown mode/variant
 * Sequences with len <= l are not errors: they yield nothing (src/lib.rs:97). */
#define S2K_OK                0
#define S2K_ERR_BAD_PARAM    -1
#define S2K_ERR_L_TOO_BIG    -2
#define S2K_ERR_CUDA         -3
#define S2K_ERR_OOM          -4
#define S2K_ERR_BAD_OFFSETS  -5   /* seq_off[0] != 0, decreasing offsets, or a sequence of >= 2^32-1 bases (host entry points;
                                     s2k_run_device trusts its device-resident offsets: see there) */
#define S2K_ERR_INTERNAL     -6   /* device-side consistency check failed */
#define S2K_ERR_NULL         -7
#define S2K_ERR_IO           -8

/* Arguments of KminmersIterator::new after `seq` (src/lib.rs:89). */
typedef struct s2k_params {
    uint32_t l;        /* minimizer length */
    uint32_t k;        /* k-min-mer length (minimizers per window) */
    double   density;  /* FH = f64, src/lib.rs:34 */
    int32_t  mode;     /* s2k_hash_mode */
    int32_t  variant;  /* s2k_hash_variant */
} s2k_params;

/* One universe minimizer: the item of NtHashHPCIterator / NtHashHPCSIMDIterator
 * (start, end, hash) -- src/nthash_hpc.rs:193, src/nthash_hpc_simd.rs:59 -- plus the index of the
 * sequence it belongs to.  start/end are in ORIGINAL coordinates of that sequence. */
typedef struct s2k_minimizer {
    uint32_t hash;
    uint32_t start;
    uint32_t end;
    uint32_t seq;
} s2k_minimizer;

#define S2K_LOC_HOST   0
#define S2K_LOC_DEVICE 1

/* Result of one batched run.  All pointers are owned by the context and stay valid until the next
 * s2k_run / s2k_run_device call on it or s2k_ctx_destroy.  `location` says whether they are host (pinned) or device
 * pointers.  Items of sequence r are [km_off[r], km_off[r+1]) in iterator order. */
typedef struct s2k_result {
    uint64_t n_seqs;
    uint64_t n_items;                /* k-min-mers (KminmerHash items) over the whole batch */
    uint64_t n_minimizers;           /* universe minimizers over the whole batch (before the simd tail rule) */
    const uint64_t *hash;            /* [n_items]  KminmerHash.hash  */
    const uint32_t *start;           /* [n_items]  KminmerHash.start */
    const uint32_t *end;             /* [n_items]  KminmerHash.end   */
    const uint8_t  *rev;             /* [n_items]  KminmerHash.rev   */
    const uint64_t *km_off;          /* [n_seqs+1] exclusive prefix of per-sequence item counts */
    const s2k_minimizer *minimizers; /* [n_minimizers] ordered minimizer stream (device runs; NULL for host runs
                                        unless S2K_WANT_MINIMIZERS was set) */
    const uint64_t *min_off;         /* [n_seqs+1] exclusive prefix of per-sequence minimizer counts (same rule) */
    const uint32_t *min_cnt;         /* [n_seqs]   minimizers per sequence that feed the window stage (i.e. after the
                                        `(len-l+1) % 16 == 0` tail rule of src/nthash_avx512_32.rs:134-138) */
    int32_t location;
    int32_t reserved;
} s2k_result;

typedef struct s2k_ctx s2k_ctx;

/* Flags for s2k_ctx_set_flags. */
#define S2K_WANT_MINIMIZERS 1u  /* s2k_run (host) also copies the minimizer stream back */
#define S2K_RLE_SCALAR_RULE 4u  /* s2k_encode_rle: the rule of the scalar `encode_rle` (src/hpc.rs:14): a repeated byte is
                                   dropped only if it is one of "ACTGactgNn"; default = `encode_rle_simd` / `hpc`
                                   (src/hpc.rs:33,86-95): every repeated byte is dropped */
#define S2K_FASTX_KEEP_BASES 32u /* s2k_run_fastx: keep the parsed bases as one ASCII batch on the host (s2k_last_fastx returns
                                   it).  Without it large files stream through the device slab by slab and are never
                                   materialised (s2k_last_fastx then returns NULL for the bases; offsets are always kept) */
#define S2K_DEBUG_TINY_CAP  8u  /* tests only: start with room for 1000 minimizers so that the grow-and-rerun path runs */
#define S2K_NO_MINIMIZER_STREAM 16u /* s2k_run_device: do not materialise the ordered minimizer stream (result.minimizers
                                   is NULL; items, km_off, min_off, min_cnt are unchanged).  The records stay where the
                                   minimizer kernel appended them and the window stage reads them in place: one pass over
                                   the records less.  Ignored for k > 12. */
#define S2K_NO_TAIL_RULE    2u  /* do not apply the `(len-l+1) % 16 == 0` tail rule of src/nthash_avx512_32.rs:134-138:
                                   for callers that process one sequence in pieces (sharding.py) and apply the rule
                                   themselves from the length of the whole sequence */

/* Lifetime.  One context = one CUDA device + one stream + grow-only device/pinned buffers.
 * A context is single-threaded (like one KminmersIterator, src/main.rs:65-79); distinct contexts may be
 * used concurrently from distinct host threads; multi-GPU = one context per device. */
int  s2k_ctx_create(int device, s2k_ctx **out);
void s2k_ctx_destroy(s2k_ctx *ctx);
int  s2k_ctx_set_flags(s2k_ctx *ctx, uint32_t flags);

/* KminmersIterator::new + collect over a batch, HOST buffers (src/lib.rs:89, 179-270).
 *   bases   : concatenated sequences, ASCII, seq_off[n_seqs] bytes
 *   seq_off : n_seqs+1 offsets, seq_off[0]==0, non-decreasing
 * Includes H2D of inputs and D2H of results; result pointers are pinned host memory. */
int s2k_run(s2k_ctx *ctx, const uint8_t *bases, const uint64_t *seq_off, uint64_t n_seqs,
            const s2k_params *params, s2k_result *out);

/* Large host batches are streamed through the device in slabs (H2D, kernels and D2H overlap on three streams): runs of
 * whole sequences, and pieces of any sequence longer than 1.5 slabs (a chromosome) -- a piece owns a base range and
 * carries a right overlap that completes its last windows; results are identical to a one-shot run.
 * bytes = target slab size (0 = default 256 MiB); batches up to 1.5 slabs go in one piece. */
int s2k_ctx_set_slab_bytes(s2k_ctx *ctx, uint64_t bytes);

/* Host ingest + run: the loop of the reference's driver, `parallel_fastx(&filename, nb_threads, task)` with
 * `task = |seq, id| KminmersIterator::new(seq, l, k, density, mode)` (src/main.rs:65-79; rust-parallelfastx is an
 * external crate).  `path` is a plain-text FASTA (multi-line allowed) or 4-line FASTQ file; it is mapped and indexed on
 * nb_threads host threads (records in file order).  Files larger than 1.5 slabs whose records have lines of one width
 * (any FASTQ, the usual FASTA) STREAM: the host threads gather slab i+1 from the mapping, line ends stripped, and pack it
 * to 2 bits per base straight into pinned staging while slab i is on the device -- the file is read once and never
 * exists as one batch on the host.  Other files (small, uneven lines) and callers that set S2K_FASTX_KEEP_BASES get
 * the records copied into one pinned ASCII batch that goes through s2k_run.  Results are identical either way.
 * s2k_last_fastx returns the parsed batch (pointers owned by the context, valid until the next fastx call; `bases` is
 * NULL after a streamed run). */
int s2k_run_fastx(s2k_ctx *ctx, const char *path, int nb_threads, const s2k_params *params, s2k_result *out);
int s2k_last_fastx(const s2k_ctx *ctx, uint64_t *n_seqs, uint64_t *n_bases, const uint8_t **bases, const uint64_t **seq_off);

/* 2-bit packed input (what the reference's src/old/hpc_2bit.rs:1-16 set out to do).  `packed` holds
 * the concatenated bases of the batch four per byte: base i in bits 2*(i%4) of byte i/4, code (ascii >> 1) & 3, i.e.
 * A=0 C=1 T=2 G=3; sequences follow each other without padding and seq_off counts BASES, as in s2k_run.  Only
 * upper-case A/C/G/T can be expressed -- input on which the ASCII form gives the same results in every mode.  A quarter
 * of the bytes cross PCIe and no host thread packs; everything else (slabs, pieces of long sequences, results) is
 * s2k_run.
 * s2k_pack2 is the matching helper: packs n_bases ASCII bases on host_threads threads; returns 0, or 1 if some byte was
 * not upper-case A/C/G/T (the output is then unusable), -1 on null arguments. */
int s2k_run_packed2(s2k_ctx *ctx, const uint8_t *packed, const uint64_t *seq_off, uint64_t n_seqs,
                    const s2k_params *params, s2k_result *out);
int64_t s2k_pack2(const uint8_t *bases, uint64_t n_bases, uint8_t *packed_out, int host_threads);

/* Transport of large host batches (s2k_run, s2k_run_fastx).  PCIe, not the GPU, bounds the end-to-end rate, so a share
 * `pack_ratio` of the slabs (default 0.7; 0 = never) is packed to 2 bits per base by `host_threads` host threads
 * (default 0 = 3/4 of the hardware threads, at most 16; AVX-512) into pinned staging, copied and unpacked on the device, while the
 * remaining slabs travel as plain ASCII so that both the packers and PCIe stay busy.  Results are identical: a slab
 * holding any byte other than upper-case A/C/G/T always travels as ASCII. */
int s2k_ctx_set_transport(s2k_ctx *ctx, int host_threads, double pack_ratio);
/* Last s2k_run / s2k_run_fastx: bytes actually copied host -> device, slabs that travelled packed / as ASCII. */
int s2k_last_transport(const s2k_ctx *ctx, uint64_t *h2d_bytes, uint64_t *packed_slabs, uint64_t *plain_slabs);

/* Same, DEVICE buffers already resident in HBM (bases 16-byte aligned); results stay on the device.
 * `stream` is a cudaStream_t (NULL = the context's own stream).  Returns after the launch sequence has been
 * enqueued and the scalar totals have been read back.  d_seq_off is TRUSTED (it lives on the device): it must start at
 * 0, be non-decreasing, end at n_bases, and every sequence must be shorter than 2^32-1 bases (coordinates are u32);
 * the host entry points check the same conditions and return S2K_ERR_BAD_OFFSETS. */
int s2k_run_device(s2k_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_seq_off, uint64_t n_seqs,
                   uint64_t n_bases, const s2k_params *params, void *stream, s2k_result *out);

/* encode_rle_simd over a batch (src/hpc.rs:44-147): homopolymer-compressed bytes and the run-start index of
 * each kept byte (relative to its sequence), plus per-sequence offsets into both.  Host buffers in, pinned
 * host pointers out (valid until the next call on the context). */
typedef struct s2k_rle_result {
    uint64_t n_seqs;
    uint64_t n_hpc;            /* total kept bytes */
    const uint8_t  *hpc;       /* [n_hpc] */
    const uint32_t *pos;       /* [n_hpc] run starts, original coordinates inside the sequence */
    const uint64_t *hpc_off;   /* [n_seqs+1] */
    int32_t location;
    int32_t reserved;
} s2k_rle_result;
int s2k_encode_rle(s2k_ctx *ctx, const uint8_t *bases, const uint64_t *seq_off, uint64_t n_seqs,
                   s2k_rle_result *out);

/* Synthetic workload of SURVEY.md 8(d), written straight into device memory: bases [first, first+count) of the
 * stream base(i) = "ACGT"[(splitmix64(seed, i>>5) >> 2*(i&31)) & 3].  Benchmark/test utility, not part of the
 * reference's API.  With stream == NULL the call runs on the context's stream and waits for it. */
int s2k_synth_device(s2k_ctx *ctx, uint64_t seed, uint64_t first, uint64_t count, uint8_t *d_out, void *stream);

/* Selection bounds exactly as the reference derives them (src/lib.rs:91, src/nthash_avx512_32.rs:47-48,
 * src/nthash2_avx512_32.rs:52-54). */
void s2k_bounds(double density, uint32_t *bound_scalar, uint32_t *bound_simd, uint32_t *bound_31);
/* ... and with H = u64 (src/lib.rs:91): (density * (u64::MAX as f64)) as u64, saturating. */
uint64_t s2k_bound_u64(double density);
/* The same for H = u16 (S2K_HASH_NT1_16): ((density as f64) * (u16::MAX as f64)) as u16, src/lib.rs:91. */
uint32_t s2k_bound_u16(double density);
/* After a DEVICE run with S2K_HASH_NT1_64: device pointer to the high 32 bits of every minimizer's hash, index-aligned
 * with result.minimizers (valid until the next run on the context); NULL after any other run. */
int s2k_last_minimizer_hash_hi(const s2k_ctx *ctx, const uint32_t **d_hi);

/* Pinned host memory for callers that want s2k_run's H2D copies to run at full PCIe rate. */
int  s2k_host_alloc(size_t bytes, void **out);
void s2k_host_free(void *p);

/* ---- Consumer side: abundance of the distinct k-min-mer hashes of an item stream.
 * rust-mdbg inserts the iterator's items into a concurrent map keyed by the k-min-mer hash (KminmerHash is equal/ordered by
 * `hash` alone, src/kminmer.rs:181-203; "Dashmap level ... kminmer hash collision", src/lib.rs:256-258).  The device table
 * does the same in HBM, so that only the distinct (hash, count, id of the first item carrying it) triples leave the GPU. */
typedef struct s2k_count_result {
    uint64_t n_distinct;
    uint64_t n_items;                /* items counted: sum of count[] */
    const uint64_t *hash;            /* [n_distinct] distinct k-min-mer hashes, in no particular order */
    const uint32_t *count;           /* [n_distinct] occurrences */
    const uint64_t *first;           /* [n_distinct] smallest item id among the occurrences (id = id_base + index into
                                        d_hash, or d_id[index]): leads back to start / end / rev of that item */
    int32_t location;                /* S2K_LOC_DEVICE */
    int32_t reserved;
} s2k_count_result;
/* d_hash: n_items k-min-mer hashes on the device (e.g. s2k_result.hash of s2k_run_device); d_id: their ids or NULL.
 * Buffers of the result belong to the context and stay valid until its next s2k_count_device call. */
int s2k_count_device(s2k_ctx *ctx, const uint64_t *d_hash, const uint64_t *d_id, uint64_t n_items, uint64_t id_base,
                     void *stream, s2k_count_result *out);
/* Several GPUs: every distinct hash is counted by rank s2k_count_part(hash, n_parts).  Buckets the items by that rank:
 * d_out_hash / d_out_id [n_items] receive the items of part 0, then part 1, ...; part_counts[n_parts] (HOST) their sizes --
 * the send buffers and split sizes of an all-to-all (NCCL), after which every rank calls s2k_count_device on what it
 * received.  n_parts <= 64. */
int s2k_count_partition_device(s2k_ctx *ctx, const uint64_t *d_hash, uint64_t n_items, uint64_t id_base, uint32_t n_parts,
                               uint64_t *part_counts, uint64_t *d_out_hash, uint64_t *d_out_id, void *stream);
uint32_t s2k_count_part(uint64_t hash, uint32_t n_parts);

/* Diagnostics. */
const char *s2k_last_error(const s2k_ctx *ctx);
const char *s2k_strerror(int status);
int s2k_abi_version(void);
/* Kernels launched by the context since creation (for benchmark bookkeeping). */
uint64_t s2k_launch_count(const s2k_ctx *ctx);
/* Name and average duration (ms) of the context's dominant kernel over the last s2k_run_device call,
 * measured with CUDA events on the launching stream when timing is enabled. */
int s2k_ctx_set_timing(s2k_ctx *ctx, int enabled);
int s2k_last_kernel_ms(const s2k_ctx *ctx, double *minimizer_ms, double *window_ms, uint32_t *minimizer_launches);

#ifdef __cplusplus
}
#endif
#endif /* SEQ2KMINMERS_H */
