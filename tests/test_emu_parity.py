"""CPU tier: the product's kernel sources, compiled by g++ onto host threads (tests/emu), against the oracle.
This checks kernel LOGIC only (halos, tile stitching, look-back order, tail rules); the GPU tier repeats the
same cases through the nvcc-built library."""
import numpy as np
import pytest

from conftest import assert_batch_matches_oracle
from parity_cases import cases, empty_tile_cases, fuzz_cases, run_device_in_place


def test_emulated_kernels_match_oracle(S, O, emu_ctx, batches, fixture_seq):
    for label, bases, so, params in cases(batches, fixture_seq, scale=1):
        for (l, k, d, mode, var) in params:
            got = emu_ctx.run(bases, so, l, k, d, S.HashMode(mode), S.HashVariant(var), want_minimizers=True)
            assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, var)


def test_emulated_rle_matches_oracle(S, O, emu_ctx, batches, fixture_seq):
    seqs = [fixture_seq[:30000], batches.seq(0), batches.seq(1), batches.seq(15), batches.seq(16), batches.seq(17),
            batches.seq(9000, runp=0.8), np.full(10000, ord("A"), np.uint8), batches.seq(5000, alphabet=b"ACGTNacgt")]
    bases, so = batches.pack(seqs)
    h, p, off = emu_ctx.encode_rle(bases, so)
    for i, s in enumerate(seqs):
        eh, ep = O.encode_rle_simd(s)
        a, b = int(off[i]), int(off[i + 1])
        assert h[a:b].tobytes() == eh and np.array_equal(p[a:b], ep), i
    # the scalar encode_rle (src/hpc.rs:14) collapses runs of "ACTGactgNn" only: junk bytes repeat in its output
    seqs = [batches.seq(4000, alphabet=b"ACGTNacgtnXX--RY", runp=0.6), batches.seq(0), np.full(300, ord("X"), np.uint8),
            batches.seq(700, alphabet=b"AX", runp=0.5), fixture_seq[:5000]]
    bases, so = batches.pack(seqs)
    h, p, off = emu_ctx.encode_rle(bases, so, scalar_rule=True)
    for i, s in enumerate(seqs):
        if len(s) == 0:
            assert off[i] == off[i + 1]
            continue
        eh, ep = O.encode_rle(s)
        a, b = int(off[i]), int(off[i + 1])
        assert h[a:b].tobytes() == eh and np.array_equal(p[a:b].astype(np.uint64), np.asarray(ep, dtype=np.uint64)), i


def test_split_invariance_emulated(S, O, emu_ctx, batches):
    """Chunk-stitch invariance: a batch gives the same per-sequence results as its two halves."""
    bases, so = batches.batch([9000, 150, 0, 20000, 31, 7000])
    whole = emu_ctx.run(bases, so, 31, 5, 0.02, S.HashMode.HpcSimd)
    cut = 3
    left = emu_ctx.run(bases[:int(so[cut])], so[:cut + 1], 31, 5, 0.02, S.HashMode.HpcSimd)
    right = emu_ctx.run(bases[int(so[cut]):], so[cut:] - so[cut], 31, 5, 0.02, S.HashMode.HpcSimd)
    assert np.array_equal(whole.hash, np.concatenate([left.hash, right.hash]))
    assert np.array_equal(whole.start, np.concatenate([left.start, right.start]))
    assert np.array_equal(whole.end, np.concatenate([left.end, right.end]))


def test_error_codes_mirror_reference_panics(S, emu_ctx):
    b = np.frombuffer(b"ACGT" * 100, dtype=np.uint8)
    so = np.array([0, 400], dtype=np.uint64)
    with pytest.raises(S.S2KError) as e:       # assert!(k<=31), src/nthash_avx512_32.rs:33
        emu_ctx.run(b, so, 32, 5, 0.1, S.HashMode.Simd)
    assert e.value.status == -2
    with pytest.raises(S.S2KError) as e:       # KSizeTooBig / assert!(k<256), src/nthash_hpc.rs:123-133
        emu_ctx.run(b, so, 256, 5, 0.1, S.HashMode.Hpc)
    assert e.value.status == -2
    for bad in ((0, 5), (5, 0)):
        with pytest.raises(S.S2KError) as e:
            emu_ctx.run(b, so, bad[0], bad[1], 0.1, S.HashMode.Hpc)
        assert e.value.status == -1
    with pytest.raises(S.S2KError) as e:       # the 31-bit iterator only exists for the SIMD modes
        emu_ctx.run(b, so, 31, 5, 0.1, S.HashMode.Hpc, S.HashVariant.NT2_31)
    assert e.value.status == -1
    with pytest.raises(S.S2KError) as e:
        emu_ctx.run(b, np.array([0, 300, 200, 400], dtype=np.uint64), 5, 2, 0.1, S.HashMode.Hpc)
    assert e.value.status == -5
    empty = emu_ctx.run(b[:0], np.array([0], dtype=np.uint64), 5, 2, 0.1, S.HashMode.Hpc)
    assert empty.n_items == 0 and empty.n_seqs == 0


def test_pipelined_host_path_emulated(S, O, emu_ctx, batches):
    """s2k_run on a batch larger than 1.5 slabs: slabs cut at sequence boundaries, offsets and sequence indices stitched."""
    seqs = [batches.seq(n) for n in [9000, 150, 0, 20000, 31, 7000, 0, 0, 12000, 150, 150, 30000, 5, 9000, 9000]]
    seqs[8] = batches.seq(12000, alphabet=b"ACGTN")       # a slab with non-ACGT bytes must travel as ASCII
    seqs[13] = batches.seq(9000, alphabet=b"ACGTacgt")
    bases, so = batches.pack(seqs)
    emu_ctx.set_slab_bytes(10000)
    try:
        for ratio, modes in ((0.7, (S.HashMode.HpcSimd, S.HashMode.Regular)), (1.0, (S.HashMode.HpcSimd,)), (0.0, (S.HashMode.Regular,))):
            emu_ctx.set_transport(3, ratio)               # 2-bit transport for 70 % / all / none of the slabs
            for mode in modes:
                got = emu_ctx.run(bases, so, 31, 3, 0.03, mode, want_minimizers=True)
                assert_batch_matches_oracle(O, got, bases, so, 31, 3, 0.03, mode)
    finally:
        emu_ctx.set_slab_bytes(0)
        emu_ctx.set_transport(0, 0.7)


def _norun(rng, n):
    """Random ACGT without equal neighbours: the HPC length equals the raw length."""
    import numpy as np
    steps = rng.integers(1, 4, n)
    steps[0] = rng.integers(0, 4)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[np.cumsum(steps) % 4].copy()


def test_long_sequences_travel_in_pieces_emulated(S, O, emu_ctx, batches):
    """A sequence longer than 1.5 slabs is cut into pieces with a right overlap (run_pipelined): owned minimizers and
    windows, coordinates of the whole sequence, and the AVX-512 tail rule applied by the last piece (lengths chosen so
    that it fires: S = kept - l + 1 is a multiple of 16)."""
    import numpy as np
    rng = np.random.default_rng(77)
    l, k, d = 21, 3, 0.3
    fire = 16 * 700 + l - 1                                # S % 16 == 0 when every base is kept
    seqs = [batches.seq(150), batches.seq(14000, runp=0.4), _norun(rng, fire), batches.seq(0), _norun(rng, fire + 5),
            batches.seq(12000, alphabet=b"ACGTN"), batches.seq(fire),                     # Simd: the raw length decides
            np.concatenate([batches.seq(8000), np.full(3500, 65, np.uint8)])]             # ends in a long homopolymer
    bases, so = batches.pack(seqs)
    emu_ctx.set_slab_bytes(6000)
    try:
        for mode, var, ratio in ((S.HashMode.HpcSimd, 0, 0.7), (S.HashMode.Simd, 0, 0.0), (S.HashMode.Hpc, 0, 0.7),
                                 (S.HashMode.Regular, 0, 0.7), (S.HashMode.HpcSimd, 1, 0.0)):
            emu_ctx.set_transport(3, ratio)
            got = emu_ctx.run(bases, so, l, k, d, mode, S.HashVariant(var), want_minimizers=True)
            assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, var)
    finally:
        emu_ctx.set_slab_bytes(0)
        emu_ctx.set_transport(0, 0.7)


def _tail_rule_cases(rng, l):
    """Long sequences whose AVX-512 tail rule fires while the last 16 kept bases are reached by l-mers that START far
    to the left: a no-run stretch followed by a homopolymer tail longer than a piece (the minimizers to drop then
    start in what would have been the previous piece)."""
    import numpy as np
    cases = []
    for lead, tail, doubled in ((6019, 3200, 5), (6019, 3200, 0), (12000, 9000, 3), (8040, 5000, 0), (20000, 1, 0)):
        kept = lead + 1
        lead += (-(kept - l + 1)) % 16                     # S = kept - l + 1 a multiple of 16: the rule fires
        body = _norun(rng, lead)
        if doubled:                                        # a few runs of two early on: the first piece (cut after 6000 raw
            at = np.sort(rng.choice(4000, doubled, replace=False))   # bases) then ends `doubled` kept bases earlier
            body = np.insert(body, at, body[at])
        t = np.full(tail, b"ACGT"[(b"ACGT".index(bytes([body[-1]])) + 1) % 4], np.uint8)
        cases.append(np.concatenate([body, t]))
    return cases


def test_tail_rule_minimizers_belong_to_the_last_piece_emulated(S, O, emu_ctx, batches):
    """the minimizers the AVX-512 tail rule drops END in the last 16 kept bases but may START a
    whole piece earlier (homopolymer tail).  The piece plan keeps every such start inside the last piece."""
    import numpy as np
    rng = np.random.default_rng(5)
    l = 21
    for seq in _tail_rule_cases(rng, l):
        kept = 1 + int(np.count_nonzero(seq[1:] != seq[:-1]))
        assert (kept - l + 1) % 16 == 0 and kept - l + 1 > 16
        bases, so = batches.pack([batches.seq(500), seq, batches.seq(300)])
        emu_ctx.set_slab_bytes(6000)
        try:
            for k, d in ((2, 1.0), (3, 0.3)):
                emu_ctx.set_transport(3, 0.5)
                got = emu_ctx.run(bases, so, l, k, d, S.HashMode.HpcSimd, S.HashVariant(0), want_minimizers=True)
                assert_batch_matches_oracle(O, got, bases, so, l, k, d, int(S.HashMode.HpcSimd), 0)
        finally:
            emu_ctx.set_slab_bytes(0)
            emu_ctx.set_transport(0, 0.7)


def _pack2_numpy(bases):
    """Independent restatement of the packed layout: base i in bits 2*(i%4) of byte i/4, code (ascii>>1)&3."""
    import numpy as np
    c = ((np.asarray(bases, dtype=np.uint8) >> 1) & 3).astype(np.uint8)
    c = np.concatenate([c, np.zeros((-len(c)) % 4, np.uint8)]).reshape(-1, 4)
    return (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)


def test_packed2_input_emulated(S, O, emu_ctx, batches):
    """s2k_run_packed2 (2-bit packed input) == the oracle on the ASCII form: one-shot, slabs of whole
    reads and pieces of long sequences (slabs then start at bases that are not multiples of 4)."""
    import numpy as np
    seqs = [batches.seq(n) for n in [151, 0, 9001, 33, 20003, 150, 7, 26001, 2]] + [batches.seq(9002, runp=0.5)]
    bases, so = batches.pack(seqs)
    packed = emu_ctx.pack2(bases, 3)
    assert np.array_equal(packed[:(len(bases) + 3) // 4], _pack2_numpy(bases))
    with pytest.raises(ValueError):
        emu_ctx.pack2(batches.seq(1000, alphabet=b"ACGTN"))
    try:
        for slab in (0, 7000):
            emu_ctx.set_slab_bytes(slab)
            for mode, var in ((S.HashMode.HpcSimd, 0), (S.HashMode.Regular, 0), (S.HashMode.Hpc, 0), (S.HashMode.Simd, 1)):
                got = emu_ctx.run(packed, so, 21, 3, 0.2, mode, S.HashVariant(var), want_minimizers=True, packed2=True)
                assert_batch_matches_oracle(O, got, bases, so, 21, 3, 0.2, mode, var)
    finally:
        emu_ctx.set_slab_bytes(0)


def test_random_parameter_fuzz_emulated(S, O, emu_ctx, batches):
    for bases, so, (l, k, d, mode, var) in fuzz_cases(batches, 25, 40000):
        got = emu_ctx.run(bases, so, l, k, d, S.HashMode(mode), S.HashVariant(var), want_minimizers=True)
        assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, var)


def test_capacity_overflow_reruns_emulated(S, O, emu_ctx, batches):
    """The minimizer buffer is sized from the expected selection rate; when a batch is denser the kernel reports the
    exact count and the host reruns once with that size.  Forced here with a 1000-record start capacity."""
    bases, so = batches.batch([40000, 150, 9000])
    got = emu_ctx.run(bases, so, 15, 3, 0.3, S.HashMode.Hpc, want_minimizers=True, debug_tiny_cap=True)
    assert got.n_minimizers > 1000
    assert_batch_matches_oracle(O, got, bases, so, 15, 3, 0.3, 1)


def test_window_stage_in_place_emulated(S, O, emu_ctx, batches, fixture_seq):
    """S2K_NO_MINIMIZER_STREAM: the window stage and the tail rule read the minimizer records where the minimizer kernel
    left them (tiles in completion order) -- same items, offsets and counts as the ordered path."""
    import ctypes
    keep = []
    def to_device(a):
        keep.append(a)
        return a, a.ctypes.data
    to_host = lambda p, nb: np.frombuffer((ctypes.c_uint8 * nb).from_address(p), dtype=np.uint8)
    for label, bases, so, params in cases(batches, fixture_seq, scale=1):
        if label in ("big-l", "non-ACGT"):
            continue
        for (l, k, d, mode, var) in params[:5]:
            got = run_device_in_place(S, emu_ctx, bases, so, l, k, d, mode, var, to_device, to_host)
            assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, var)
            ref = emu_ctx.run(bases, so, l, k, d, S.HashMode(mode), S.HashVariant(var))
            assert np.array_equal(got.min_off, ref.min_off) and np.array_equal(got.min_cnt, ref.min_cnt), label
    # sparse selection over many tiles: windows whose k minimizers span several tiles, tiles without any minimizer
    bases, so = batches.batch([120000, 150, 90000])
    for (l, k, d, mode) in [(31, 12, 0.0005, 3), (31, 5, 0.0002, 2), (15, 9, 0.001, 1), (31, 3, 0.00005, 0)]:
        got = run_device_in_place(S, emu_ctx, bases, so, l, k, d, mode, 0, to_device, to_host)
        assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, 0)
    bases, so = empty_tile_cases(batches, 170000)         # long runs of tiles without minimizers (index walks -> binary search)
    for (l, k, d, mode) in [(31, 5, 0.01, 3), (31, 12, 0.02, 1)]:
        got = run_device_in_place(S, emu_ctx, bases, so, l, k, d, mode, 0, to_device, to_host)
        assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, 0, check_minimizers=False)


def test_h64_flavour_emulated(S, O, emu_ctx, batches, fixture_seq):
    """H = u64: KAT-2 of the reference (tests/main.rs:18-39) through the kernels, closed-form parity."""
    from parity_cases import check_h64_flavour
    check_h64_flavour(S, O, emu_ctx, batches, fixture_seq, scale=1)


def test_h16_flavour_emulated(S, O, emu_ctx, batches, fixture_seq):
    """H = u16 (parity unpinned): closed-form parity in the modes Hpc and Regular."""
    from parity_cases import check_h16_flavour
    check_h16_flavour(S, O, emu_ctx, batches, fixture_seq, scale=1)


def test_kminmer_vec_flavour_emulated(S, O, emu_ctx, batches, fixture_seq):
    """KminmerVec (src/kminmer.rs:17-126) built from the minimizer stream: the same windows as the KminmerHash items
    (start, end, offset), canonical orientation by lexicographic order of the hash vector."""
    b = emu_ctx.run(fixture_seq, np.array([0, len(fixture_seq)], dtype=np.uint64), 31, 5, 0.01, S.HashMode.HpcSimd,
                    want_minimizers=True)
    vecs, items = list(b.kminmer_vecs(0, 5)), list(b.items(0))
    assert len(vecs) == len(items) > 1000
    mins = b.minimizers_of(0)["hash"]
    n_rev = 0
    for c, (v, it) in enumerate(zip(vecs, items)):
        assert (v.start, v.end, v.offset) == (it.start, it.end, it.offset) and v.is_normalized()
        w = [int(x) for x in mins[c:c + 5]]
        assert v.mers() == (w[::-1] if v.rev else w) and v.rev == (w[::-1] < w)
        n_rev += v.rev
    assert 0 < n_rev < len(vecs)
    a, r = S.KminmerVec([3, 1, 2], 0, 9, 0), S.KminmerVec([2, 1, 3], 0, 9, 0)
    assert a == r and a.rev != r.rev and a.mers() == [2, 1, 3] and a.print() == "2 1 3 "
    assert a.get_hash_u64() == r.get_hash_u64() != S.KminmerVec([2, 1, 4], 0, 9, 0).get_hash_u64()
    assert sorted([S.KminmerVec([5, 1], 0, 1, 0), S.KminmerVec([1, 2], 0, 1, 1)])[0].mers() == [1, 2]


def test_minimizer_tsv_round_trip_emulated(S, O, emu_ctx, fixture_seq, tmp_path):
    """src/old/kminmers-readwrite.rs: `{prefix}-{l}-{density}.mers` written from the device's minimizers, k-min-mers
    rebuilt from the file alone; lines and items against the oracle's minimizers."""
    seq = bytes(fixture_seq)[:40000]
    for hpc, l, k, d in ((True, 31, 5, 0.01), (False, 12, 3, 0.02)):
        prefix = str(tmp_path / ("hpc" if hpc else "reg"))
        w = S.KminmersWriteIterator(seq, l, k, d, hpc, prefix, ctx=emu_ctx)
        assert w.path == f"{prefix}-{l}-{d}.mers"
        written = list(w)
        read = list(S.KminmersReadIterator(l, k, d, prefix))
        st, en, h = O.closed_profile(seq, l, d, 1 if hpc else 0)
        lines = open(w.path).read().splitlines()
        assert lines == [f"{int(p) + (0 if hpc else 1)}\t{int(x)}" for p, x in zip(st, h)] and len(lines) > 100
        assert len(read) == len(written) == len(h) - k + 1
        for c, (a, b) in enumerate(zip(written, read)):
            assert a == b and (a.start, a.end, a.offset, a.rev) == (b.start, b.end, b.offset, b.rev)
            assert a.offset == c and a.start == int(st[c]) + (0 if hpc else 1)
            assert a.end == int(st[c + k - 1]) + (0 if hpc else 1) + l - 1
            w5 = [int(x) for x in h[c:c + k]]
            assert a.mers() == min(w5, w5[::-1])
    assert S.mers_path("p", 31, 0.01) == "p-31-0.01.mers" and S.mers_path("p", 5, 1.0) == "p-5-1.mers"
    assert S.mers_path("p", 5, 1e-7) == "p-5-0.0000001.mers" and S.mers_path("p", 5, 0.5) == "p-5-0.5.mers"


def test_hash_width_flavours_through_slabs_emulated(S, emu_ctx):
    """H = u16 and H = u64 through the pipelined host path: slabs, a sequence cut into pieces, 2-bit transport -- the
    same items and minimizers as the one-shot run."""
    rng = np.random.default_rng(7)
    lens = [0, 31, 150, 9000, 20000, 47, 120000, 3, 700, 64000]
    seqs = [np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 4 if i % 3 else 5, n)] for i, n in enumerate(lens)]
    so = np.zeros(len(seqs) + 1, dtype=np.uint64)
    so[1:] = np.cumsum(lens)
    bases = np.concatenate(seqs)
    try:
        for var in (S.HashVariant.NT1_16, S.HashVariant.NT1_64):
            for mode in (S.HashMode.Hpc, S.HashMode.Regular):
                emu_ctx.set_slab_bytes(0)
                a = emu_ctx.run(bases, so, 21, 4, 0.03, mode, var, want_minimizers=True)
                emu_ctx.set_slab_bytes(30000)
                b = emu_ctx.run(bases, so, 21, 4, 0.03, mode, var, want_minimizers=True)
                assert a.n_items > 1000 and emu_ctx.last_transport()[2] > 1
                for f in ("hash", "start", "end", "rev", "km_off", "min_off", "min_cnt", "minimizers"):
                    assert np.array_equal(getattr(a, f), getattr(b, f)), (var, mode, f)
    finally:
        emu_ctx.set_slab_bytes(0)
