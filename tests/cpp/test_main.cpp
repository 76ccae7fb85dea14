// C++ twin of the reference's integration test (tests/main.rs:13-90) on top of include/seq2kminmers.hpp:
// golden k-min-mer hashes (tests/main.rs:41-57), RLE on the fixture, and the mode-equivalence sweep
// (tests/main.rs:82-89).  Needs a CUDA device; run by tests/test_gpu_parity.py.
#include "../../include/seq2kminmers.hpp"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>

static std::vector<uint8_t> load_fixture(const char *path)
{
    std::ifstream f(path, std::ios::binary);
    std::vector<uint8_t> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (raw.size() < 16 || std::memcmp(raw.data(), "S2KFIX01", 8) != 0) { std::fprintf(stderr, "bad fixture\n"); std::exit(2); }
    uint64_t n;
    std::memcpy(&n, raw.data() + 8, 8);
    std::vector<uint8_t> seq(n);
    for (uint64_t i = 0; i < n; ++i) seq[i] = (uint8_t)"ACGT"[(raw[16 + (i >> 2)] >> (2 * (i & 3))) & 3];
    return seq;
}
#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char **argv)
{
    using namespace s2k;
    const std::vector<uint8_t> contents = load_fixture(argc > 1 ? argv[1] : "tests/golden/ecoli100k.2bit");
    const std::vector<uint64_t> hashes32 = {143479479014703ull, 1415094313937202ull, 7085699921625713ull,
        2731023262850893ull, 3529660833839258ull, 2520689800435504ull, 3515165585325381ull, 2855190423625803ull,
        5122855536061684ull, 244022361441902ull, 2856446528761135ull, 906939906227534ull, 2115341643533671ull,
        246274980452770ull, 159737436030657ull};
    Context ctx(0);
    {
        KminmersIterator iter(ctx, contents.data(), contents.size(), 10, 5, 0.0001, HashMode::Regular);
        size_t count = 0;
        for (const KminmerHash &kminmer : iter) { CHECK(count < hashes32.size()); CHECK(kminmer.get_hash() == hashes32[count]); ++count; }
        CHECK(count == hashes32.size());
    }
    {
        auto rle = encode_rle_simd(ctx, contents.data(), contents.size());
        CHECK(rle.first == hpc(ctx, contents.data(), contents.size()));
        CHECK(rle.first.size() == 72873 && rle.second.size() == 72873 && rle.second[0] == 0);
        for (size_t i = 1; i < rle.first.size(); ++i) CHECK(rle.first[i] != rle.first[i - 1] && rle.second[i] > rle.second[i - 1]);
    }
    for (size_t l : {5, 7, 11, 17, 25, 31})
        for (size_t k : {2, 5, 8}) {
            auto a = KminmersIterator(ctx, contents.data(), contents.size(), l, k, 0.01, HashMode::Regular).collect();
            auto b = KminmersIterator(ctx, contents.data(), contents.size(), l, k, 0.01, HashMode::Simd).collect();
            auto c = KminmersIterator(ctx, contents.data(), contents.size(), l, k, 0.01, HashMode::Hpc).collect();
            auto d = KminmersIterator(ctx, contents.data(), contents.size(), l, k, 0.01, HashMode::HpcSimd).collect();
            CHECK(a == b);
            CHECK(c == d);
            CHECK(!a.empty() && !c.empty());
        }
    {   // H = u64 build of the crate (src/lib.rs:30-32): its golden vector, tests/main.rs:18-39, through the same wrapper
        const std::vector<uint64_t> hashes = {6097375827354318ull, 5077268723048817ull, 17093614815813553ull, 13932651659877218ull,
            2254626575123847ull, 4725847317728813ull, 10971942364167709ull, 1406844240705087ull, 15284878278949327ull,
            13429516156719180ull, 10760699289819902ull, 11244197813995113ull, 6993910349997344ull, 22098843726082404ull,
            4944933674400292ull, 14212811059278321ull, 9310664830401458ull, 11232758307960192ull, 9720472733789719ull,
            13210101786532125ull};
        KminmersIterator iter(ctx, contents.data(), contents.size(), 10, 5, 0.0001, HashMode::Regular, HashVariant::NT1_64);
        size_t count = 0;
        for (const KminmerHash &kminmer : iter) { CHECK(count < hashes.size()); CHECK(kminmer.get_hash() == hashes[count]); ++count; }
        CHECK(count == hashes.size());
    }
    {   // KminmerVec (src/kminmer.rs:17-60): canonical orientation, order by the vector
        const uint32_t a[3] = {3, 1, 2}, b[3] = {2, 1, 3};
        KminmerVec x(a, 3, 0, 9, 0), y(b, 3, 0, 9, 0);
        CHECK(x == y && x.rev && !y.rev && x.is_normalized() && x.mers() == std::vector<uint32_t>({2, 1, 3}));
    }
    try {   // assert!(k<=31), src/nthash_avx512_32.rs:33
        KminmersIterator bad(ctx, contents.data(), contents.size(), 32, 5, 0.01, HashMode::Simd);
        CHECK(false);
    } catch (const Error &e) { CHECK(e.status == S2K_ERR_L_TOO_BIG); }
    std::printf("test_main ok\n");
    return 0;
}
