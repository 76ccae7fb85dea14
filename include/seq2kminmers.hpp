// seq2kminmers.hpp -- header-only C++ host side above the C ABI (seq2kminmers.h).
//
// The reference is a Rust crate; this image has no Rust toolchain, so the host-side mirror of its operator API
// is written in C++ (the Rust binding a maintainer would add is shown in INTEGRATION.md).  Names, argument
// order and error behaviour follow the reference:
//   HashMode                         src/lib.rs:21-27
//   KminmerHash{hash,start,end,offset,rev}, equality/order on .hash only     src/kminmer.rs:128-135,181-203
//   KminmerVec{mers,start,end,offset,rev}, canonical vector of minimizer hashes   src/kminmer.rs:17-126
//   KminmersIterator::new(seq,l,k,density,mode) -> iterator of KminmerHash   src/lib.rs:89,179-270
//   encode_rle_simd / hpc                                                    src/hpc.rs:28-41,44-147
// Where the reference panics (unwrap/assert) this wrapper throws s2k::Error carrying the C status code.
#pragma once
#include "seq2kminmers.h"

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace s2k {

enum class HashMode : int { Regular = 0, Hpc = 1, Simd = 2, HpcSimd = 3 };
enum class HashVariant : int { NT1_32 = 0, NT2_31 = 1, NT1_64 = 2, NT1_16 = 3 };

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string &m) : std::runtime_error("s2k status " + std::to_string(st) + ": " + m), status(st) {}
};

struct KminmerHash {
    uint64_t hash;
    size_t start, end, offset;
    bool rev;
    uint64_t get_hash() const { return hash; }                                     // trait Kminmer
    bool operator==(const KminmerHash &o) const { return hash == o.hash; }         // src/kminmer.rs:181-185
    bool operator<(const KminmerHash &o) const { return hash < o.hash; }           // src/kminmer.rs:193-197
};
using KminmerType = KminmerHash;

// KminmerVec (src/kminmer.rs:17-126): the k-min-mer as the vector of its minimizer hashes in canonical orientation
// (the reversed vector if it is lexicographically smaller, then rev = true); equality and order by the vector.  Built on
// the host from the minimizer stream of a result (kminmer_vecs below).
struct KminmerVec {
    std::vector<uint32_t> mers_;
    size_t start = 0, end = 0, offset = 0;
    bool rev = false;
    KminmerVec() = default;
    KminmerVec(const uint32_t *mers, size_t k, size_t start_, size_t end_, size_t offset_)      // Kminmer::new, src/kminmer.rs:27-38
        : mers_(mers, mers + k), start(start_), end(end_), offset(offset_) { normalize(); }
    void normalize()                                                                            // src/kminmer.rs:53-60
    {
        std::vector<uint32_t> r(mers_.rbegin(), mers_.rend());
        if (r < mers_) { mers_.swap(r); rev = true; }
    }
    bool is_normalized() const { return mers_ <= std::vector<uint32_t>(mers_.rbegin(), mers_.rend()); }   // src/kminmer.rs:63-67
    std::vector<uint32_t> mers() const { return mers_; }                                        // src/kminmer.rs:80-82
    bool operator==(const KminmerVec &o) const { return mers_ == o.mers_; }
    bool operator<(const KminmerVec &o) const { return mers_ < o.mers_; }
};
// The items of sequence r of a HOST result that carries the minimizer stream (S2K_WANT_MINIMIZERS), KminmerVec flavour.
inline std::vector<KminmerVec> kminmer_vecs(const s2k_result &res, uint64_t r, size_t k)
{
    std::vector<KminmerVec> out;
    if (!res.minimizers || res.location != S2K_LOC_HOST) return out;
    const s2k_minimizer *m = res.minimizers + res.min_off[r];
    const size_t n = res.min_cnt[r];
    std::vector<uint32_t> h(k);
    for (size_t c = 0; c + k <= n; ++c) {
        for (size_t t = 0; t < k; ++t) h[t] = m[c + t].hash;
        out.emplace_back(h.data(), k, m[c].start, m[c + k - 1].end, c);
    }
    return out;
}

// One CUDA device + stream + buffers; single-threaded like one iterator (src/main.rs:65-79).
class Context {
  public:
    explicit Context(int device = 0)
    {
        int st = s2k_ctx_create(device, &ctx_);
        if (st != S2K_OK) throw Error(st, std::string(s2k_strerror(st)) + " (no CUDA device? there is no CPU fallback)");
    }
    ~Context() { s2k_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;

    // Batched KminmersIterator::new + collect.  Result pointers stay valid until the next run on this context.
    s2k_result run(const uint8_t *bases, const uint64_t *seq_off, uint64_t n_seqs, size_t l, size_t k, double density,
                   HashMode mode, HashVariant variant = HashVariant::NT1_32)
    {
        s2k_params p{(uint32_t)l, (uint32_t)k, density, (int32_t)mode, (int32_t)variant};
        s2k_result r;
        check(s2k_run(ctx_, bases, seq_off, n_seqs, &p, &r));
        return r;
    }
    // 2-bit packed input (4 bases per byte, A0 C1 T2 G3; seq_off counts bases): a quarter of the PCIe bytes, no host packing.
    s2k_result run_packed2(const uint8_t *packed, const uint64_t *seq_off, uint64_t n_seqs, size_t l, size_t k, double density,
                           HashMode mode, HashVariant variant = HashVariant::NT1_32)
    {
        s2k_params p{(uint32_t)l, (uint32_t)k, density, (int32_t)mode, (int32_t)variant};
        s2k_result r;
        check(s2k_run_packed2(ctx_, packed, seq_off, n_seqs, &p, &r));
        return r;
    }
    // The driver's file mode (src/main.rs:50-81): FASTA/FASTQ file -> k-min-mers of every record, nb_threads parser threads.
    s2k_result run_fastx(const std::string &path, int nb_threads, size_t l, size_t k, double density, HashMode mode,
                         HashVariant variant = HashVariant::NT1_32)
    {
        s2k_params p{(uint32_t)l, (uint32_t)k, density, (int32_t)mode, (int32_t)variant};
        s2k_result r;
        check(s2k_run_fastx(ctx_, path.c_str(), nb_threads, &p, &r));
        return r;
    }
    // Host-side knobs of run(): slab size of the three-stream pipeline, share of slabs packed to 2 bits/base.
    void set_slab_bytes(uint64_t bytes) { check(s2k_ctx_set_slab_bytes(ctx_, bytes)); }
    void set_transport(int host_threads, double pack_ratio) { check(s2k_ctx_set_transport(ctx_, host_threads, pack_ratio)); }
    s2k_rle_result encode_rle(const uint8_t *bases, const uint64_t *seq_off, uint64_t n_seqs)
    {
        s2k_rle_result r;
        check(s2k_encode_rle(ctx_, bases, seq_off, n_seqs, &r));
        return r;
    }
    s2k_ctx *raw() { return ctx_; }

  private:
    void check(int st)
    {
        if (st != S2K_OK) throw Error(st, s2k_last_error(ctx_));
    }
    s2k_ctx *ctx_ = nullptr;
};

// KminmersIterator::new(seq, l, k, density, mode): the sequence is processed on the GPU at construction (a batch
// of one); begin()/end() then walk the items in the reference's order.  For throughput use Context::run.
class KminmersIterator {
  public:
    KminmersIterator(Context &ctx, const uint8_t *seq, size_t len, size_t l, size_t k, double density, HashMode mode,
                     HashVariant variant = HashVariant::NT1_32)
    {
        const uint64_t off[2] = {0, len};
        const s2k_result r = ctx.run(seq, off, 1, l, k, density, mode, variant);
        items_.reserve(r.n_items);
        for (uint64_t i = 0; i < r.n_items; ++i)
            items_.push_back(KminmerHash{r.hash[i], r.start[i], r.end[i], (size_t)i, r.rev[i] != 0});
    }
    std::vector<KminmerHash>::const_iterator begin() const { return items_.begin(); }
    std::vector<KminmerHash>::const_iterator end() const { return items_.end(); }
    const std::vector<KminmerHash> &collect() const { return items_; }

  private:
    std::vector<KminmerHash> items_;
};

// encode_rle_simd(s) -> (hpc string, run-start positions)   src/hpc.rs:44
inline std::pair<std::string, std::vector<uint32_t>> encode_rle_simd(Context &ctx, const uint8_t *seq, size_t len)
{
    const uint64_t off[2] = {0, len};
    const s2k_rle_result r = ctx.encode_rle(seq, off, 1);
    return {std::string(reinterpret_cast<const char *>(r.hpc), r.n_hpc), std::vector<uint32_t>(r.pos, r.pos + r.n_hpc)};
}
inline std::string hpc(Context &ctx, const uint8_t *seq, size_t len) { return encode_rle_simd(ctx, seq, len).first; }

} // namespace s2k
