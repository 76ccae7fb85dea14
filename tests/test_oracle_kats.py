"""CPU tier: the oracle against every golden vector the reference holds for this path,
the reference's structural test assertions, and procedural-restatement == closed-form on adversarial input."""
import numpy as np
import pytest

KAT1 = [143479479014703, 1415094313937202, 7085699921625713, 2731023262850893, 3529660833839258, 2520689800435504,
        3515165585325381, 2855190423625803, 5122855536061684, 244022361441902, 2856446528761135, 906939906227534,
        2115341643533671, 246274980452770, 159737436030657]                      # reference tests/main.rs:41-57
KAT2 = [6097375827354318, 5077268723048817, 17093614815813553, 13932651659877218, 2254626575123847, 4725847317728813,
        10971942364167709, 1406844240705087, 15284878278949327, 13429516156719180, 10760699289819902,
        11244197813995113, 6993910349997344, 22098843726082404, 4944933674400292, 14212811059278321,
        9310664830401458, 11232758307960192, 9720472733789719, 13210101786532125]  # reference tests/main.rs:18-39
KAT3_SEQ = b"ACTGCACATGATGAGTAGATGATGATGATGATGATATGATGATAT"
KAT3 = [(0, 1693589515812555183), (6, 876319423165292601), (13, 771890730643629033), (16, 826464090118103095),
        (33, 1245321008145464903), (34, 1193606442387228521)]                     # src/old/nthash_hpc.rs.opt4:96-97
DEMO = b"AACTGCACTGCACTGCACTGCACACTGCACTGCACTGCACTGCACACTGCACTGCACTGACTGCACTGCACTGCACTGCACTGCCTGC"  # src/main.rs:15


def test_kat1_reference_golden_u32(O, fixture_seq):
    r = O.kminmers(fixture_seq, 10, 5, 0.0001, O.REGULAR)
    assert [int(x) for x in r["hash"]] == KAT1
    assert r["n_minimizers"] == 19
    assert (int(r["start"][0]), int(r["end"][0]), int(r["rev"][0])) == (2341, 9477, 1)
    assert (int(r["start"][-1]), int(r["end"][-1]), int(r["offset"][-1])) == (72729, 96866, 14)


def test_kat2_reference_golden_u64(O, fixture_seq):
    _, _, h = O.closed_minimizers(fixture_seq, 10, False, False, 64, int(0.0001 * (2 ** 64 - 1)), False, 0, 0)
    wh, _ = O.closed_windows(h, 5, mix_u32=False)
    assert len(h) == 24 and [int(x) for x in wh] == KAT2


def test_kat3_opt4_doc_golden(O):
    st, _, h = O.closed_minimizers(KAT3_SEQ, 4, True, False, 64, int(0.1 * (2 ** 64 - 1)), True, 1, 1)
    assert list(zip(map(int, st), map(int, h))) == KAT3


def test_kat4_config1_aggregates(O, fixture_seq):
    exp = {O.HPC: (1475, 1471, 11725885919757810115, 475087621700556523, 74148365, 74608025, 745),
           O.HPCSIMD: (1475, 1471, 11725885919757810115, 475087621700556523, 74148365, 74607500, 745),
           O.REGULAR: (1946, 1942, 14336893697199334593, 899118319923364179, 95949596, 96406930, 981),
           O.SIMD: (1946, 1942, 14336893697199334593, 899118319923364179, 95949596, 96406930, 981)}
    for mode, e in exp.items():
        r = O.kminmers(fixture_seq, 31, 5, 0.01, mode)
        h = r["hash"]
        got = (r["n_minimizers"], len(h), int(h.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(h)),
               int(r["start"].sum()), int(r["end"].sum()), int(r["rev"].sum()))
        assert got == e, mode


def test_kat5_reference_sweep_mode_equivalence(O, fixture_seq):
    """tests/main.rs:82-89: Regular == Simd and Hpc == HpcSimd as hash sequences, 6 x 3 parameter pairs."""
    counts = {5: (2055, 2351), 7: (1488, 1935), 11: (1412, 2079), 17: (1441, 2038), 25: (1488, 1966), 31: (1475, 1946)}
    for l, (n_hpc, n_reg) in counts.items():
        for k in (2, 5, 8):
            reg, simd = (O.kminmers(fixture_seq, l, k, 0.01, m) for m in (O.REGULAR, O.SIMD))
            hp, hs = (O.kminmers(fixture_seq, l, k, 0.01, m) for m in (O.HPC, O.HPCSIMD))
            assert np.array_equal(reg["hash"], simd["hash"]) and np.array_equal(hp["hash"], hs["hash"])
            assert reg["n_minimizers"] == n_reg and hp["n_minimizers"] == n_hpc


def test_rle_equalities_of_reference_test(O, fixture_seq):
    """tests/main.rs:76-78."""
    a, pa = O.encode_rle(fixture_seq)
    b, pb = O.encode_rle_simd(fixture_seq)
    assert a == O.hpc(fixture_seq) and a == b and np.array_equal(pa, pb.astype(np.uint64))
    assert len(a) == 72873


def test_kat6_demo_and_doc_examples(O):
    st, en, h = O.minimizers(DEMO, 28, 0.1, O.HPC)
    assert list(zip(map(int, st), map(int, en), map(int, h))) == [
        (4, 31, 237497718), (6, 33, 204567447), (26, 53, 237497718), (28, 55, 204567447), (45, 72, 68739741),
        (47, 74, 343349922)]
    r = O.kminmers(DEMO, 28, 5, 0.1, O.REGULAR)
    assert [(int(a), int(b), int(c), int(d)) for a, b, c, d in zip(r["hash"], r["start"], r["end"], r["rev"])] == [
        (2666914291542867539, 0, 55, 1), (265907413648215804, 4, 72, 1), (2874609099774383239, 6, 74, 0),
        (3388454018765070272, 26, 84, 0)]
    r = O.kminmers(DEMO, 10, 5, 0.1, O.HPC)                                      # src/lib.rs:61-62 doc example
    assert [(int(a), int(b), int(c)) for a, b, c in zip(r["hash"], r["start"], r["end"])] == [
        (1781245113506412305, 14, 61), (4056819534695535436, 20, 62)]


def test_kat7_profile_divergence(O):
    q = (b"AACTTTTTGGGGGGCAAAAAACCCCCCCTGCCCCCCAAACTTTTTGGGGGGCAAAAAACCCCCCCTGCCCCCCAAACTTTTTGGGGGGCAAAAAACCCCCCCTGCCCCCCA")
    s = O.minimizers(q, 5, 0.5, O.HPC)
    v = O.minimizers(q, 5, 0.5, O.HPCSIMD)
    assert len(s[0]) == 26 and len(v[0]) == 27
    assert [int(x) for x in s[1][:6]] == [14, 20, 27, 28, 29, 35] and [int(x) for x in v[1][:6]] == [14, 15, 21, 28, 29, 30]


def test_kat8_doc_example_u32(O):
    st, en, h = O.minimizers(KAT3_SEQ, 4, 0.1, O.HPC)
    assert len(h) == 14
    assert list(zip(map(int, st), map(int, en), map(int, h)))[:4] == [(0, 3, 223693230), (2, 5, 113835212),
                                                                      (4, 7, 312416605), (7, 10, 282161344)]


def test_kat9_generator_and_pipeline(O):
    assert O.synth_word(0x5EED0002, 0) == 0x8c44e9ef30ca4931
    rd = O.synth(0x5EED0002, 0, 20000)
    assert rd[:64].tobytes() == b"CATACGACGGATAATATTGTCGGTACACATAGTGCCTGAGTATGTGTAGACGCCGACCATTATA"
    exp = {(O.HPC, 0): (270, 266, 11666637108378545117, 2831243), (O.HPCSIMD, 0): (270, 266, 11666637108378545117, 2831144),
           (O.SIMD, 0): (408, 404, 2326486756798998148, 3810617), (O.SIMD, 1): (383, 379, 9878680082315566458, 3655795),
           (O.HPCSIMD, 1): (291, 287, 287700146607339001, 3010222)}
    for (mode, var), e in exp.items():
        r = O.kminmers(rd, 31, 5, 0.01, mode, var)
        assert (r["n_minimizers"], len(r["hash"]), int(r["hash"].sum(dtype=np.uint64)), int(r["end"].sum())) == e
    rd11 = O.synth(0x5EED0002, 220000, 20000)       # S % 16 == 0: the AVX-512 tail rule is live
    assert len(O.minimizers(rd11, 31, 0.01, O.HPC)[0]) == 327 and len(O.minimizers(rd11, 31, 0.01, O.HPCSIMD)[0]) == 326


def test_bounds_recipe(O):
    table = {0.0001: (429496, 429496), 0.001: (4294967, 4294967), 0.007: (30064771, 30064772), 0.01: (42949672, 42949672),
             0.02: (85899345, 85899344), 0.05: (214748364, 214748368), 0.1: (429496729, 429496736),
             0.5: (2147483647, 2147483648), 1.0: (4294967295, 4294967295)}
    for d, (bs, bv) in table.items():
        assert O.bound_scalar(d) == bs and O.bound_simd(bs) == bv


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_restatement_equals_closed_form(O, batches, mode):
    """The procedural restatement (ring buffers / 16-lane blocks) against the Appendix-A closed form."""
    cases = [batches.seq(n, alphabet=a, runp=p) for n, a, p in
             [(2000, b"ACGT", 0.0), (3000, b"ACGT", 0.6), (1500, b"ACGTNacgtXY", 0.3), (64, b"AC", 0.5), (47, b"ACGT", 0.0),
              (33, b"ACGT", 0.0), (32, b"ACGT", 0.0), (31, b"ACGT", 0.0), (5, b"ACGT", 0.0), (0, b"ACGT", 0.0)]]
    for l in (1, 2, 5, 16, 31) + ((64, 255) if mode in (0, 1) else ()):
        for d in (0.05, 0.5, 1.0):
            for s in cases:
                a = O.minimizers(s, l, d, mode)
                b = O.closed_profile(s, l, d, mode)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2].astype(np.uint32)), (l, d, len(s))
                if mode in (2, 3) and l <= 31:
                    a = O.minimizers(s, l, d, mode, O.NT2_31)
                    b = O.closed_profile(s, l, d, mode, O.NT2_31)
                    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2].astype(np.uint32))


def test_window_stage_rolling_equals_closed(O, batches):
    s = batches.seq(30000)
    for k in (1, 2, 5, 8, 10, 40, 70):
        r = O.kminmers(s, 11, k, 0.05, O.REGULAR)
        _, _, mh = O.minimizers(s, 11, 0.05, O.REGULAR)
        h, rv = O.closed_windows(mh, k)
        assert np.array_equal(h, r["hash"]) and np.array_equal(rv, r["rev"])


def test_avx512_baseline_equals_scalar_oracle(O, batches):
    """oracle/s2k_cpu_avx512.c (the CPU baseline bench.py reports) is bit-exact against the scalar restatement."""
    if not O.has_avx512():
        pytest.skip("host CPU lacks AVX-512")
    lens = [20000] * 10 + [150] * 100 + [0, 5, 31, 32, 33, 46, 47, 48, 62, 63, 64, 50000] + list(batches.rng.integers(0, 300, 200))
    bases, so = batches.pack([batches.seq(n, alphabet=b"ACGT" if i % 5 else b"ACGTNacgt", runp=0.3 if i % 3 == 0 else 0.0)
                              for i, n in enumerate(lens)])
    for mode in (O.SIMD, O.HPCSIMD):
        for (l, k, d) in [(31, 5, 0.01), (31, 2, 0.5), (5, 3, 0.2), (16, 1, 1.0)]:
            a = O.avx512_batch(bases, so, l, k, d, mode, threads=3, want_digest=True)
            b = O.batch(bases, so, l, k, d, mode, 0, threads=3, want_digest=True)
            assert a["total"] == b["total"] and np.array_equal(a["km_cnt"], b["km_cnt"])
            assert np.array_equal(a["digest"], b["digest"]), (mode, l, k, d)
