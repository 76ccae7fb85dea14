"""Parity cases shared by the emulation tier (CPU, tests/emu) and the GPU tier (C ABI on a B200).
Each case: (label, bases, seq_off, [(l, k, density, mode, variant), ...])."""
import numpy as np
import pytest

REG, HPC, SIMD, HPCSIMD = 0, 1, 2, 3


def all_modes(l, k, d, nt2=True):
    out = [(l, k, d, m, 0) for m in (REG, HPC, SIMD, HPCSIMD) if m in (REG, HPC) or l <= 31]
    if nt2 and l <= 31:
        out += [(l, k, d, SIMD, 1), (l, k, d, HPCSIMD, 1)]
    return out


def cases(B, fixture_seq, scale=1):
    """B: conftest.Batches.  scale=1 for the CPU emulation tier, larger on the GPU."""
    rng = B.rng
    out = []
    so1 = np.array([0, len(fixture_seq)], dtype=np.uint64)
    # reference fixture: config 1 and the sweep of tests/main.rs:82-89
    par = all_modes(31, 5, 0.01) + [(10, 5, 0.0001, REG, 0)]
    sweep_l = (5, 7, 11, 17, 25, 31) if scale > 1 else (5, 17)
    for l in sweep_l:
        for k in ((2, 5, 8) if scale > 1 else (2, 8)):
            par += [(l, k, 0.01, m, 0) for m in (REG, HPC, SIMD, HPCSIMD)]
    out.append(("fixture", fixture_seq, so1, par))
    # lengths 0..l+1 and friends, empty reads between, first/last empty
    lens = [0, 0] + list(rng.integers(0, 80, 300 * scale)) + [0, 0, 0, 31, 32, 33, 30, 1, 2, 47, 48, 0]
    b, so = B.batch(lens)
    out.append(("short+empty", b, so, all_modes(31, 5, 0.05) + all_modes(5, 2, 0.3) + all_modes(17, 3, 0.1, nt2=False)))
    # 150-bp reads (config 3 shape)
    b, so = B.batch([150] * (300 * scale))
    out.append(("150bp", b, so, all_modes(31, 5, 0.01) + all_modes(31, 2, 0.05, nt2=False)))
    # homopolymers far longer than the halo, runs crossing tile boundaries, N blocks
    parts = [B.seq(5000), np.full(20000, ord("A"), np.uint8), B.seq(3000), np.full(9000, ord("N"), np.uint8), B.seq(40),
             np.full(300, ord("T"), np.uint8), B.seq(12000, runp=0.7)]
    s1 = np.concatenate(parts)
    b, so = B.pack([s1, B.seq(100), s1[::-1].copy(), np.full(8000, ord("C"), np.uint8), B.seq(500)])
    out.append(("homopolymers", b, so, all_modes(31, 5, 0.05) + all_modes(7, 2, 0.2, nt2=False)))
    # sequences that do not compress at all in HPC space (more owners per tile than one hash pass covers)
    norun = B.seq(50000)
    for i in range(1, len(norun)):
        if norun[i] == norun[i - 1]:
            norun[i] = b"ACGT"[(b"ACGT".index(int(norun[i])) + 1 + int(rng.integers(0, 3))) % 4]
    b, so = B.pack([np.frombuffer(b"ACGT" * 10000, dtype=np.uint8).copy(), norun, B.seq(3000)])
    out.append(("incompressible", b, so, all_modes(31, 5, 0.02, nt2=False) + [(31, 3, 0.5, HPCSIMD, 0), (31, 3, 0.5, HPC, 0)]))
    # non-ACGT bytes: N, lower case, IUPAC, junk (scalar table vs nibble table)
    b, so = B.batch([3000, 5000, 100, 20000], alphabet=b"ACGTNacgtnXRY-")
    out.append(("non-ACGT", b, so, all_modes(21, 4, 0.1)))
    # clean ACGT sequences with ONE other byte each: the packed tiles' validity test (PRMT table look-up against the
    # base itself) has to turn exactly that tile back to the byte form -- at tile and halo boundaries, as the first and
    # last base of a word, a piece and a sequence; bytes that are a base in the nibble profile only (lower case, 'Q' =
    # 0x51 -> A, 'D' = 0x44 -> T, 0x81, 0xc3), bytes that are none in either (0, 'N', 'L' = 0x4c, 0xff, '@', 'E', 'U')
    specials = [ord("N"), ord("a"), ord("t"), 0x00, 0x81, 0xC3, 0x51, 0x44, 0x4C, 0xFF, 0x40, 0x45, 0x55, 0x57, 0x37, 0x14]
    places = [0, 1, 3, 4, 15, 16, 63, 64, 15871, 15872, 15873, 16127, 16128, 16129, 16383, 16384, 19999]
    seqs = []
    for i, pos in enumerate(places):
        q = B.seq(20000)
        q[pos] = specials[i % len(specials)]
        seqs.append(q)
    for i, sp in enumerate(specials):
        q = B.seq(600)
        q[37 + i] = sp
        seqs.append(q)
    seqs.append(B.seq(20000))
    b, so = B.pack(seqs)
    out.append(("one-odd-byte", b, so, all_modes(31, 3, 0.05) + [(12, 2, 0.2, HPCSIMD, 0), (12, 2, 0.2, HPC, 0)]))
    # large l (scalar profiles only), incl. halo 512 path
    b, so = B.batch([30000, 700, 255, 256, 257, 100], runp=0.3)
    out.append(("big-l", b, so, [(l, 3, 0.05, m, 0) for m in (REG, HPC) for l in (64, 127, 128, 200, 255)]))
    # AVX-512 tail rule: S % 16 == 0 and neighbours, at very high density so every l-mer is a minimizer
    lens = [31 + 15 + 16 * j for j in range(0, 20)] + [46, 47, 48, 62, 63, 64, 16, 31]
    b, so = B.batch(lens)
    out.append(("tail-rule", b, so, [(31, 2, 0.5, SIMD, 0), (31, 2, 1.0, SIMD, 0), (31, 2, 0.5, HPCSIMD, 0),
                                     (31, 2, 1.0, HPCSIMD, 0), (31, 2, 0.5, SIMD, 1), (16, 1, 1.0, SIMD, 0),
                                     (16, 1, 1.0, HPCSIMD, 0)]))
    # dense selection, l sweep, several tiles
    b, so = B.batch([20000, 20000, 9000])
    par = []
    for l in (1, 2, 5, 16, 31):
        par += all_modes(l, 3, 0.5 if l > 2 else 1.0, nt2=(l == 31))
    par += [(31, 5, 0.0, REG, 0), (31, 5, 0.0, HPCSIMD, 0), (31, 5, 1e-9, HPCSIMD, 0), (31, 1, 0.01, HPC, 0),
            (31, 70, 0.05, HPC, 0)]
    out.append(("dense", b, so, par))
    return out


def fuzz_cases(B, n_cases, max_len):
    """Random (batch, parameters): every mode/variant, l across the allowed range, k, density over four decades,
    batch shapes from many tiny reads to a few long ones, alphabets with and without non-ACGT bytes."""
    rng = B.rng
    out = []
    for _ in range(n_cases):
        mode = int(rng.integers(0, 4))
        variant = int(rng.integers(0, 2)) if mode in (SIMD, HPCSIMD) else 0
        l = int(rng.integers(1, 32)) if mode in (SIMD, HPCSIMD) else int(rng.choice([rng.integers(1, 32), rng.integers(32, 256)]))
        k = int(rng.choice([1, 2, 3, 5, 8, 10, 33]))
        d = float(10 ** rng.uniform(-3.3, 0.0))
        shape = int(rng.integers(0, 4))
        if shape == 0:
            lens = list(rng.integers(0, 3 * l + 40, int(rng.integers(1, 400))))
        elif shape == 1:
            lens = [150] * int(rng.integers(1, 300))
        elif shape == 2:
            lens = list(rng.integers(1000, max_len, int(rng.integers(1, 5))))
        else:
            lens = list(rng.integers(0, 2000, int(rng.integers(1, 60)))) + [int(rng.integers(20000, max_len))]
        alphabet = [b"ACGT", b"ACGT", b"ACGTN", b"ACGTacgtNnRY"][int(rng.integers(0, 4))]
        runp = float(rng.choice([0.0, 0.0, 0.3, 0.8]))
        bases, so = B.batch(lens, alphabet=alphabet, runp=runp)
        out.append((bases, so, (l, k, d, mode, variant)))
    return out


def run_device_in_place(S, ctx, bases, so, l, k, d, mode, var, to_device, to_host):
    """s2k_run_device with S2K_NO_MINIMIZER_STREAM on a batch held by the caller: `to_device(np array) -> (handle, ptr)`
    puts an array where the context's kernels can read it, `to_host(ptr, nbytes) -> np.uint8 array` reads results back.
    Returns a KminmersBatch without a minimizer stream."""
    n = len(so) - 1
    pad = np.zeros(len(bases) + 16, dtype=np.uint8)
    pad[:len(bases)] = bases
    hb, pb = to_device(pad)
    ho, po = to_device(np.ascontiguousarray(so, dtype=np.uint64))
    r = ctx.run_device(pb, po, n, len(bases), l, k, d, S.HashMode(mode), S.HashVariant(var), no_minimizer_stream=True)
    assert not r.minimizers or k > 12            # NULL: the ordered stream was not materialised
    ni = int(r.n_items)
    get = lambda p, c, dt: to_host(p, c * np.dtype(dt).itemsize).view(dt).copy() if c else np.zeros(0, dt)
    return S.KminmersBatch(n, get(r.hash, ni, np.uint64), get(r.start, ni, np.uint32), get(r.end, ni, np.uint32),
                           get(r.rev, ni, np.uint8), get(r.km_off, n + 1, np.uint64), get(r.min_off, n + 1, np.uint64),
                           get(r.min_cnt, n, np.uint32), int(r.n_minimizers), None)


def empty_tile_cases(B, n_poly=400000, l=31):
    """Homopolymers of hundreds of kilobases: more than eight consecutive tiles without any minimizer in the HPC modes
    (the index walks of the in-place window stage turn into binary searches) -- in the middle of a sequence (windows
    spanning the gap) and at the end of two sequences whose head lengths are chosen so that the AVX-512 tail rule
    (S > 16 and S % 16 == 0 in HPC space) fires on the first and not on the second."""
    poly = np.full(n_poly, ord("A"), np.uint8)
    head = B.seq(4000)
    kept = lambda s: 1 + int(np.count_nonzero(s[1:] != s[:-1]))
    seqs = [np.concatenate([head, poly, B.seq(3000)])]
    n = 3000
    while (kept(np.concatenate([head[:n], poly[:8]])) - l + 1) % 16 != 0:
        n += 1
    seqs += [np.concatenate([head[:n], poly]), np.concatenate([head[:n + 1], poly]), B.seq(500)]
    if head[n] == ord("A"):                                # n + 1 would not add a kept base
        seqs[2] = np.concatenate([head[:n + 2], poly])
    return B.pack(seqs)


def check_h64_flavour(S, O, ctx, B, fixture_seq, scale=1):
    """S2K_HASH_NT1_64 (the crate built with `pub type H = u64`, src/lib.rs:30-32): the reference's own golden vector
    for that build (KAT-2, tests/main.rs:18-39) through the ABI, then items against the oracle's closed form (64-bit
    seeds src/nthash_hpc.rs:30-49, `hash <= (density * u64::MAX as f64) as u64`, identity mix src/lib.rs:171-177) in the
    modes it exists for, on batches with empty and short reads, homopolymers, non-ACGT bytes and l beyond 32."""
    from test_oracle_kats import KAT2
    V = S.HashVariant.NT1_64
    it = S.KminmersIterator(fixture_seq, 10, 5, 0.0001, S.HashMode.Regular, variant=V, ctx=ctx)
    assert [km.get_hash() for km in it] == KAT2
    lens = [0, 5, 11, 12, 40, 300, 0, 2500, 16000 * scale, 33, 64, 65]
    b1, so1 = B.batch(lens, runp=0.3)
    b2, so2 = B.batch([4000, 100, 9000], alphabet=b"ACGTNacgtXRY", runp=0.2)
    for bases, so in ((b1, so1), (b2, so2)):
        for mode, l, k, d in ((REG, 10, 5, 0.01), (HPC, 10, 3, 0.02), (REG, 31, 2, 0.05), (HPC, 31, 5, 0.05),
                              (HPC, 64, 2, 0.1), (REG, 100, 1, 0.3), (HPC, 11, 4, 1.0), (REG, 17, 3, 0.0)):
            got = ctx.run(bases, so, l, k, d, S.HashMode(mode), V)
            bound = min(int(d * float(2 ** 64 - 1)), 2 ** 64 - 1)
            for r in range(len(so) - 1):
                seq = bases[int(so[r]):int(so[r + 1])]
                hpc = mode == HPC
                st, en, h = O.closed_minimizers(seq, l, hpc, False, 64, bound, False, 1 if hpc else 0, 1 if hpc else 0)
                a, e = int(got.km_off[r]), int(got.km_off[r + 1])
                if len(h) < k:
                    assert e == a, (mode, l, k, d, r)
                    continue
                wh, wr = O.closed_windows(h, k, mix_u32=False)
                assert e - a == len(wh), (mode, l, k, d, r, e - a, len(wh))
                assert np.array_equal(got.hash[a:e], wh) and np.array_equal(got.rev[a:e], wr), (mode, l, k, d, r)
                assert np.array_equal(got.start[a:e], st[:len(wh)].astype(np.uint32)), (mode, l, k, d, r)
                assert np.array_equal(got.end[a:e], en[k - 1:].astype(np.uint32)), (mode, l, k, d, r)
    with pytest.raises(S.S2KError):                    # the SIMD iterators are 32-bit
        ctx.run(b1, so1, 10, 3, 0.01, S.HashMode.HpcSimd, V)


def check_h16_flavour(S, O, ctx, B, fixture_seq, scale=1):
    """S2K_HASH_NT1_16 (the crate built with `pub type H = u16`, src/lib.rs:29): items against the oracle's closed form
    -- mode Hpc: 16-bit ntHash1 state (seeds `as u16`, src/nthash_hpc.rs:30-49); mode Regular: the nthash32 hash
    truncated (`x as H`, src/lib.rs:224); `hash <= (density * u16::MAX as f64) as u16` (src/lib.rs:91); MixHash<u16>
    (src/lib.rs:142-155).  Parity unpinned (no reference-held vector for this build); tied to the pinned 32-bit profile
    at density 1: the Regular minimizer hashes must be the low halves of the 32-bit ones (KAT-1's build)."""
    V = S.HashVariant.NT1_16
    lens = [0, 5, 11, 12, 40, 300, 0, 2500, 16000 * scale, 33, 64, 65]
    b1, so1 = B.batch(lens, runp=0.3)
    b2, so2 = B.batch([4000, 100, 9000], alphabet=b"ACGTNacgtXRY", runp=0.2)
    fx = np.frombuffer(bytes(fixture_seq), dtype=np.uint8)[:30000]
    b3, so3 = fx, np.array([0, len(fx)], dtype=np.uint64)
    for bases, so in ((b1, so1), (b2, so2), (b3, so3)):
        for mode, l, k, d in ((REG, 10, 5, 0.01), (HPC, 10, 3, 0.02), (REG, 31, 2, 0.05), (HPC, 31, 5, 0.05),
                              (HPC, 64, 2, 0.1), (REG, 100, 1, 0.3), (HPC, 11, 4, 1.0), (REG, 17, 3, 0.0),
                              (HPC, 16, 2, 0.01), (HPC, 33, 3, 0.004), (REG, 31, 5, 0.01)):
            got = ctx.run(bases, so, l, k, d, S.HashMode(mode), V, want_minimizers=True)
            bound = min(int(d * 65535.0), 65535)
            hpc = mode == HPC
            for r in range(len(so) - 1):
                seq = bases[int(so[r]):int(so[r + 1])]
                st, en, h = O.closed_minimizers(seq, l, hpc, False, 16 if hpc else 3216, bound, False,
                                                1 if hpc else 0, 1 if hpc else 0)
                m0, m1 = int(got.min_off[r]), int(got.min_off[r + 1])
                assert m1 - m0 == len(h), (mode, l, k, d, r, m1 - m0, len(h))
                assert np.array_equal(got.minimizers["hash"][m0:m1], h.astype(np.uint32)), (mode, l, k, d, r)
                a, e = int(got.km_off[r]), int(got.km_off[r + 1])
                if len(h) < k:
                    assert e == a, (mode, l, k, d, r)
                    continue
                wh, wr = O.closed_windows(h, k, mix_u32=2)
                assert e - a == len(wh), (mode, l, k, d, r, e - a, len(wh))
                assert np.array_equal(got.hash[a:e], wh) and np.array_equal(got.rev[a:e], wr), (mode, l, k, d, r)
                assert np.array_equal(got.start[a:e], st[:len(wh)].astype(np.uint32)), (mode, l, k, d, r)
                assert np.array_equal(got.end[a:e], en[k - 1:].astype(np.uint32)), (mode, l, k, d, r)
    # density 1: every l-mer is selected in both builds, so the u16 Regular stream is the u32 one truncated
    g16 = ctx.run(b2, so2, 12, 3, 1.0, S.HashMode.Regular, V, want_minimizers=True)
    g32 = ctx.run(b2, so2, 12, 3, 1.0, S.HashMode.Regular, S.HashVariant.NT1_32, want_minimizers=True)
    assert g16.n_minimizers == g32.n_minimizers > 0
    assert np.array_equal(g16.minimizers["hash"], g32.minimizers["hash"] & np.uint32(0xffff))
    b16 = ctx.lib.c.s2k_bound_u16
    assert b16(0.01) == 655 and b16(2.0) == 65535 and b16(-1.0) == 0 and b16(float("nan")) == 0
    with pytest.raises(S.S2KError):                    # the SIMD iterators take a u32 bound
        ctx.run(b1, so1, 10, 3, 0.01, S.HashMode.HpcSimd, V)
