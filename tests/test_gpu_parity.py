"""GPU tier (B200): the nvcc-built library through the C ABI, bit-exact against the CPU oracle.
Small/medium sizes compare full item tuples per sequence; BASELINE-scale shapes are covered by per-read
digests of full tuples on a slab plus size-independent properties (split invariance, order, offsets)."""
import os

import numpy as np
import pytest

from conftest import assert_batch_matches_oracle
from parity_cases import cases, empty_tile_cases, fuzz_cases, run_device_in_place

pytestmark = pytest.mark.gpu


def test_library_is_the_cuda_build(S, gpu_ctx):
    assert S.LIB_PATH.exists() and gpu_ctx.lib.path == S.LIB_PATH
    n0 = gpu_ctx.launch_count
    gpu_ctx.run(np.frombuffer(b"ACGT" * 50, dtype=np.uint8), np.array([0, 200], dtype=np.uint64), 5, 2, 0.5, S.HashMode.Hpc)
    assert gpu_ctx.launch_count > n0


def test_parity_cases_full_tuples(S, O, gpu_ctx, batches, fixture_seq):
    n0 = gpu_ctx.launch_count
    for label, bases, so, params in cases(batches, fixture_seq, scale=4):
        for (l, k, d, mode, var) in params:
            got = gpu_ctx.run(bases, so, l, k, d, S.HashMode(mode), S.HashVariant(var), want_minimizers=True)
            assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, var)
    assert gpu_ctx.launch_count > n0


def test_reference_golden_vectors_through_the_abi(S, gpu_ctx, fixture_seq):
    """tests/main.rs:41-57 (KAT-1) and the reference's own test structure (tests/main.rs:60-89)."""
    from test_oracle_kats import KAT1
    it = S.KminmersIterator(fixture_seq, 10, 5, 0.0001, S.HashMode.Regular, ctx=gpu_ctx)
    assert [km.get_hash() for km in it] == KAT1
    for l in (5, 7, 11, 17, 25, 31):
        for k in (2, 5, 8):
            a = list(S.KminmersIterator(fixture_seq, l, k, 0.01, S.HashMode.Regular, ctx=gpu_ctx))
            b = list(S.KminmersIterator(fixture_seq, l, k, 0.01, S.HashMode.Simd, ctx=gpu_ctx))
            c = list(S.KminmersIterator(fixture_seq, l, k, 0.01, S.HashMode.Hpc, ctx=gpu_ctx))
            d = list(S.KminmersIterator(fixture_seq, l, k, 0.01, S.HashMode.HpcSimd, ctx=gpu_ctx))
            assert a == b and c == d and len(a) > 0 and len(c) > 0


def test_rle_primitives(S, O, gpu_ctx, batches, fixture_seq):
    """tests/main.rs:76-78 on the GPU path."""
    h, p = S.encode_rle_simd(fixture_seq, ctx=gpu_ctx)
    eh, ep = O.encode_rle_simd(fixture_seq)
    assert h == eh and np.array_equal(p, ep) and S.hpc(fixture_seq, ctx=gpu_ctx) == O.hpc(fixture_seq)
    seqs = [batches.seq(n, runp=rp) for n, rp in [(0, 0), (1, 0), (15, 0), (16, 0), (17, 0), (100000, 0.7), (40000, 0)]]
    seqs.append(np.full(70000, ord("A"), np.uint8))
    bases, so = batches.pack(seqs)
    hh, pp, off = gpu_ctx.encode_rle(bases, so)
    for i, s in enumerate(seqs):
        eh, ep = O.encode_rle_simd(s)
        assert hh[int(off[i]):int(off[i + 1])].tobytes() == eh and np.array_equal(pp[int(off[i]):int(off[i + 1])], ep)
    # scalar encode_rle (src/hpc.rs:7-25): only runs of "ACTGactgNn" collapse (src/hpc.rs:14)
    junk = batches.seq(50000, alphabet=b"ACGTNacgtnXX--RY", runp=0.6)
    eh, ep = O.encode_rle(junk)
    h, p = S.encode_rle(junk, ctx=gpu_ctx)
    assert h == eh and np.array_equal(p, np.asarray(ep, dtype=np.uint64))
    assert S.encode_rle(fixture_seq, ctx=gpu_ctx)[0] == O.encode_rle(fixture_seq)[0]


@pytest.mark.parametrize("shape", ["hifi20k", "short150", "chromosome"])
def test_synthetic_shapes_digest(S, O, gpu_ctx, shape):
    """Synthetic reads of the BASELINE shapes (SURVEY.md 8d generator), every read verified through an
    order-sensitive digest of its full (hash, start, end, rev) tuples computed by the oracle on all host cores."""
    threads = os.cpu_count() or 1
    if shape == "hifi20k":
        L, n, seed, runs = 20000, 3000, 0x5EED0002, [(31, 5, 0.01, 3, 0), (31, 5, 0.01, 1, 0)]
    elif shape == "short150":
        L, n, seed, runs = 150, 400000, 0x5EED0003, [(31, 5, 0.01, 3, 0), (31, 5, 0.01, 1, 0), (31, 5, 0.01, 2, 0)]
    else:
        L, n, seed, runs = 30_000_000, 1, 0x5EED0004, [(31, 5, 0.01, 3, 1), (31, 5, 0.01, 2, 1), (31, 5, 0.01, 3, 0)]
    bases = O.synth(seed, 0, L * n)
    so = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    for (l, k, d, mode, var) in runs:
        got = gpu_ctx.run(bases, so, l, k, d, S.HashMode(mode), S.HashVariant(var), copy=False)
        want = O.batch(bases, so, l, k, d, mode, var, threads=threads, want_digest=True)
        assert got.n_items == want["total"] and got.n_minimizers >= want["total_min"]
        assert np.array_equal(np.diff(got.km_off), want["km_cnt"])
        assert np.array_equal(got.min_cnt.astype(np.uint64), want["min_cnt"])
        dg = O.digest_items(got.hash, got.start, got.end, got.rev, got.km_off)
        assert np.array_equal(dg, want["digest"]), (shape, mode, var)
        if got.n_items:
            assert bool(np.all(got.start <= got.end))


def test_kat9_first_synthetic_read(S, gpu_ctx, O):
    rd = O.synth(0x5EED0002, 0, 20000)
    so = np.array([0, 20000], dtype=np.uint64)
    r = gpu_ctx.run(rd, so, 31, 5, 0.01, S.HashMode.HpcSimd)
    assert (r.n_minimizers, r.n_items, int(r.hash.sum(dtype=np.uint64)), int(r.end.astype(np.uint64).sum())) == \
        (270, 266, 11666637108378545117, 2831144)
    r = gpu_ctx.run(rd, so, 31, 5, 0.01, S.HashMode.Simd, S.HashVariant.NT2_31)
    assert (r.n_minimizers, r.n_items, int(r.hash.sum(dtype=np.uint64))) == (383, 379, 9878680082315566458)


def test_device_api_and_full_scale_properties(S, O, gpu_ctx):
    """Device-resident path at a BASELINE-scale shape (config 2 geometry: 20-kb reads, >2^30 bases so that the
    launch is split into slabs): results equal the host-buffer path on a slab, are invariant to how the batch is
    split, and satisfy order/offset properties; a sample of reads is checked tuple by tuple against the oracle."""
    import torch
    L, n = 20000, 60000                                   # 1.2 Gbp: crosses the 2^30-base slab boundary
    seed = 0x5EED0002
    dev = torch.device("cuda:0")
    d_bases = torch.empty(L * n, dtype=torch.uint8, device=dev)
    gpu_ctx.synth_device(seed, 0, L * n, d_bases.data_ptr())
    d_so = (torch.arange(n + 1, dtype=torch.int64, device=dev) * L)
    torch.cuda.synchronize()
    assert bytes(d_bases[:64].cpu().numpy()) == b"CATACGACGGATAATATTGTCGGTACACATAGTGCCTGAGTATGTGTAGACGCCGACCATTATA"

    def run(first_read, n_reads):
        r = gpu_ctx.run_device(d_bases.data_ptr() + first_read * L, d_so.data_ptr(), n_reads, n_reads * L, 31, 5, 0.01,
                               S.HashMode.HpcSimd)
        t = lambda p, c, isz, dt: torch.as_tensor(S.DeviceArray(p, c * isz, "|u1"), device=dev).view(dt).clone()
        return dict(n=int(r.n_items), hash=t(r.hash, r.n_items, 8, torch.int64), start=t(r.start, r.n_items, 4, torch.int32),
                    end=t(r.end, r.n_items, 4, torch.int32), rev=t(r.rev, r.n_items, 1, torch.uint8),
                    km_off=t(r.km_off, n_reads + 1, 8, torch.int64))

    whole = run(0, n)
    assert whole["n"] > 0 and int(whole["km_off"][-1]) == whole["n"]
    assert bool((whole["km_off"][1:] >= whole["km_off"][:-1]).all())
    assert bool((whole["start"] <= whole["end"]).all())
    # split invariance at a read boundary that is not a tile/slab boundary
    cut = 33333
    left, right = run(0, cut), run(cut, n - cut)
    for key in ("hash", "start", "end", "rev"):
        assert torch.equal(whole[key], torch.cat([left[key], right[key]])), key
    assert torch.equal(whole["km_off"][:cut + 1], left["km_off"])
    # sample of reads against the oracle, including the reads around the slab boundary (2^30 / 20000 ~ 53687)
    tile_reads = (65536 * 16128) // L
    sample = [0, 1, 11, cut - 1, cut, tile_reads - 1, tile_reads, tile_reads + 1, n - 1]
    km = whole["km_off"].cpu().numpy()
    for rdx in sample:
        a, b = int(km[rdx]), int(km[rdx + 1])
        want = O.kminmers(O.synth(seed, rdx * L, L), 31, 5, 0.01, O.HPCSIMD)
        assert b - a == len(want["hash"]), rdx
        assert np.array_equal(whole["hash"][a:b].cpu().numpy().view(np.uint64), want["hash"]), rdx
        assert np.array_equal(whole["start"][a:b].cpu().numpy().view(np.uint32), want["start"].astype(np.uint32)), rdx
        assert np.array_equal(whole["end"][a:b].cpu().numpy().view(np.uint32), want["end"].astype(np.uint32)), rdx
        assert np.array_equal(whole["rev"][a:b].cpu().numpy(), want["rev"]), rdx


def test_error_codes(S, gpu_ctx):
    b = np.frombuffer(b"ACGT" * 100, dtype=np.uint8)
    so = np.array([0, 400], dtype=np.uint64)
    for args, status in [((32, 5, 0.1, S.HashMode.Simd), -2), ((256, 5, 0.1, S.HashMode.Hpc), -2),
                         ((0, 5, 0.1, S.HashMode.Hpc), -1), ((5, 0, 0.1, S.HashMode.Hpc), -1)]:
        with pytest.raises(S.S2KError) as e:
            gpu_ctx.run(b, so, *args)
        assert e.value.status == status


def test_cpp_host_wrapper_twin_of_reference_test(S, tmp_path):
    """include/seq2kminmers.hpp: the C++ twin of the reference's tests/main.rs, linked against the product library."""
    import subprocess
    from conftest import ROOT
    exe = tmp_path / "test_main"
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", str(exe), str(ROOT / "tests" / "cpp" / "test_main.cpp"),
                    f"-L{S.PKG_DIR}", "-l:libs2k_b200.so", f"-Wl,-rpath,{S.PKG_DIR}"], check=True)
    out = subprocess.run([str(exe), str(ROOT / "tests" / "golden" / "ecoli100k.2bit")], capture_output=True, text=True)
    assert out.returncode == 0 and "test_main ok" in out.stdout, out.stdout + out.stderr


def test_pipelined_host_path(S, O, gpu_ctx, batches):
    """s2k_run streaming a batch through the device in many small slabs (three streams) == the oracle."""
    lens = [9000, 150, 0, 20000, 31, 7000, 0, 0, 12000, 150, 150, 30000, 5] * 6
    seqs = [batches.seq(n) for n in lens]
    seqs[8] = batches.seq(12000, alphabet=b"ACGTN")       # slabs with non-ACGT bytes must travel as ASCII
    seqs[40] = batches.seq(7000, alphabet=b"ACGTacgt")
    bases, so = batches.pack(seqs)
    gpu_ctx.set_slab_bytes(40000)
    try:
        for ratio, threads in ((0.7, 0), (1.0, 5), (0.0, 0)):   # 2-bit transport for 70 % / all / none of the slabs
            gpu_ctx.set_transport(threads, ratio)
            for mode in (S.HashMode.HpcSimd, S.HashMode.Hpc, S.HashMode.Simd):
                got = gpu_ctx.run(bases, so, 31, 3, 0.03, mode, want_minimizers=True)
                assert_batch_matches_oracle(O, got, bases, so, 31, 3, 0.03, mode)
    finally:
        gpu_ctx.set_slab_bytes(0)
        gpu_ctx.set_transport(0, 0.7)


def test_long_sequences_travel_in_pieces(S, O, gpu_ctx, batches):
    """s2k_run cuts sequences longer than 1.5 slabs into pieces with a right overlap; results == one-shot run == oracle.
    Lengths are chosen so that the AVX-512 tail rule fires (S = kept - l + 1 multiple of 16) and does not fire."""
    rng = np.random.default_rng(5)

    def norun(n):
        steps = rng.integers(1, 4, n)
        steps[0] = rng.integers(0, 4)
        return np.frombuffer(b"ACGT", dtype=np.uint8)[np.cumsum(steps) % 4].copy()

    l, k, d = 31, 5, 0.05
    fire = 16 * 40000 + l - 1
    seqs = [batches.seq(150), batches.seq(900000, runp=0.4), norun(fire), batches.seq(0), norun(fire + 3),
            batches.seq(700000, alphabet=b"ACGTN"), batches.seq(fire), batches.seq(5000),
            np.concatenate([batches.seq(500000), np.full(70000, 84, np.uint8)])]
    bases, so = batches.pack(seqs)
    gpu_ctx.set_slab_bytes(100000)
    try:
        for mode, var, ratio in ((S.HashMode.HpcSimd, 0, 0.7), (S.HashMode.Simd, 0, 0.0), (S.HashMode.Hpc, 0, 0.5),
                                 (S.HashMode.Regular, 0, 0.7), (S.HashMode.HpcSimd, 1, 0.0)):
            gpu_ctx.set_transport(0, ratio)
            got = gpu_ctx.run(bases, so, l, k, d, mode, S.HashVariant(var), want_minimizers=True)
            assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, var)
    finally:
        gpu_ctx.set_slab_bytes(0)
        gpu_ctx.set_transport(0, 0.7)


def test_packed2_input(S, O, gpu_ctx, batches):
    """s2k_run_packed2: 2-bit packed host input == the oracle on the ASCII form (one-shot, slabs, pieces)."""
    seqs = [batches.seq(n) for n in [151, 0, 90001, 33, 400003, 150, 7, 260001, 2]] + [batches.seq(90002, runp=0.5)]
    bases, so = batches.pack(seqs)
    packed = gpu_ctx.pack2(bases)
    c = ((bases >> 1) & 3).astype(np.uint8)
    c = np.concatenate([c, np.zeros((-len(c)) % 4, np.uint8)]).reshape(-1, 4)
    assert np.array_equal(packed[:(len(bases) + 3) // 4], c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6))
    try:
        for slab in (0, 70000):
            gpu_ctx.set_slab_bytes(slab)
            for mode, var in ((S.HashMode.HpcSimd, 0), (S.HashMode.Regular, 0), (S.HashMode.Hpc, 0), (S.HashMode.Simd, 1)):
                got = gpu_ctx.run(packed, so, 31, 5, 0.05, mode, S.HashVariant(var), want_minimizers=True, packed2=True)
                assert_batch_matches_oracle(O, got, bases, so, 31, 5, 0.05, mode, var)
    finally:
        gpu_ctx.set_slab_bytes(0)


def test_one_sequence_split_across_ranks(S, O, gpu_ctx):
    """SURVEY 8(e), config 4 shape: a 30-Mbp sequence processed as 4 base ranges (overlap-and-trim by ownership) gives
    exactly the k-min-mers of the whole sequence, for the 31-bit hash, the scalar HPC profile and ntHash1 HpcSimd
    (whose AVX-512 tail rule needs the HPC length of the whole sequence)."""
    import importlib
    sharding = importlib.import_module("rust-seq2kminmers_b200.sharding")
    n = 30_000_000
    seq = O.synth(0x5EED0004, 0, n)
    so = np.array([0, n], dtype=np.uint64)
    l, k, d = 31, 5, 0.01
    for mode, var in [(3, 1), (1, 0), (3, 0)]:
        whole = gpu_ctx.run(seq, so, l, k, d, S.HashMode(mode), S.HashVariant(var))
        ranges = sharding.sequence_ranges(n, 4, sharding.default_overlap_right(l, k, d))
        kept = [sharding.kept_in_range(seq[lo:hi], lo, b0, b1, True) for (b0, b1, lo, hi) in ranges]
        parts = [sharding.run_sequence_part(gpu_ctx, seq[lo:hi], lo, b0, b1, n, l, k, d, S.HashMode(mode), S.HashVariant(var),
                                            kept_total=sum(kept)) for (b0, b1, lo, hi) in ranges]
        assert sum(len(p["hash"]) for p in parts) == whole.n_items
        assert np.array_equal(np.concatenate([p["hash"] for p in parts]), whole.hash)
        assert np.array_equal(np.concatenate([p["start"] for p in parts]), whole.start.astype(np.uint64))
        assert np.array_equal(np.concatenate([p["end"] for p in parts]), whole.end.astype(np.uint64))
        assert np.array_equal(np.concatenate([p["rev"] for p in parts]), whole.rev)


def test_random_parameter_fuzz(S, O, gpu_ctx, batches):
    """120 random (batch, mode, variant, l, k, density) draws, full tuples and minimizer streams against the oracle."""
    for bases, so, (l, k, d, mode, var) in fuzz_cases(batches, 120, 120000):
        got = gpu_ctx.run(bases, so, l, k, d, S.HashMode(mode), S.HashVariant(var), want_minimizers=True)
        assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, var)


def test_capacity_overflow_reruns(S, O, gpu_ctx, batches):
    bases, so = batches.batch([200000, 150, 9000, 150000])
    for mode in (S.HashMode.Hpc, S.HashMode.Simd):
        got = gpu_ctx.run(bases, so, 15, 3, 0.3, mode, want_minimizers=True, debug_tiny_cap=True)
        assert got.n_minimizers > 1000
        assert_batch_matches_oracle(O, got, bases, so, 15, 3, 0.3, int(mode))


def test_contexts_on_concurrent_host_threads(S, O, batches):
    """include/seq2kminmers.h: a context is single-threaded, distinct contexts may run concurrently from distinct host
    threads (one stream each) -- the pattern of one context per worker thread in INTEGRATION.md."""
    import threading
    work = []
    for i in range(4):
        bases, so = batches.batch([150] * 2000 + [20000] * 20 + [int(x) for x in batches.rng.integers(0, 3000, 200)])
        work.append((bases, so, [S.HashMode.HpcSimd, S.HashMode.Hpc, S.HashMode.Simd, S.HashMode.Regular][i]))
    results, errors = [None] * 4, []

    def worker(i):
        try:
            with S.Context(0) as ctx:
                for _ in range(5):
                    results[i] = ctx.run(work[i][0], work[i][1], 31, 5, 0.02, work[i][2], want_minimizers=True)
        except Exception as e:          # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i in range(4):
        assert_batch_matches_oracle(O, results[i], work[i][0], work[i][1], 31, 5, 0.02, int(work[i][2]))


def test_window_stage_in_place(S, O, gpu_ctx, batches, fixture_seq):
    """S2K_NO_MINIMIZER_STREAM through s2k_run_device: the window stage and the tail rule read the minimizer records
    where k_minimizers appended them (per-CTA regions, tiles in completion order).  Same tuples as the oracle, same
    min_off / min_cnt as the ordered path; then config-2 geometry against the ordered path item by item."""
    import torch
    dev = torch.device("cuda:0")

    def to_device(a):
        t = torch.from_numpy(a.view(np.uint8).copy()).to(dev)
        return t, t.data_ptr()
    to_host = lambda p, nb: torch.as_tensor(S.DeviceArray(p, nb, "|u1"), device=dev).cpu().numpy()
    for label, bases, so, params in cases(batches, fixture_seq, scale=4):
        if label == "big-l":
            continue
        for (l, k, d, mode, var) in params:
            got = run_device_in_place(S, gpu_ctx, bases, so, l, k, d, mode, var, to_device, to_host)
            assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, var)
            ref = gpu_ctx.run(bases, so, l, k, d, S.HashMode(mode), S.HashVariant(var))
            assert np.array_equal(got.min_off, ref.min_off) and np.array_equal(got.min_cnt, ref.min_cnt), label
    bases, so = batches.batch([1200000, 150, 900000])
    for (l, k, d, mode) in [(31, 12, 0.0005, 3), (31, 5, 0.0002, 2), (15, 9, 0.001, 1), (31, 3, 0.00005, 0)]:
        got = run_device_in_place(S, gpu_ctx, bases, so, l, k, d, mode, 0, to_device, to_host)
        assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, 0)
    bases, so = empty_tile_cases(batches)                 # long runs of tiles without minimizers (index walks -> binary search)
    for (l, k, d, mode) in [(31, 5, 0.01, 3), (31, 12, 0.02, 3), (21, 3, 0.01, 1)]:
        got = run_device_in_place(S, gpu_ctx, bases, so, l, k, d, mode, 0, to_device, to_host)
        assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, 0, check_minimizers=False)
    # config-2 geometry, 0.4 Gbp: every item equal to the ordered path
    L, n = 20000, 20000
    d_bases = torch.empty(L * n + 16, dtype=torch.uint8, device=dev)
    gpu_ctx.synth_device(0x5EED0002, 0, L * n, d_bases.data_ptr())
    d_so = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    outs = []
    for flag in (False, True):
        r = gpu_ctx.run_device(d_bases.data_ptr(), d_so.data_ptr(), n, n * L, 31, 5, 0.01, S.HashMode.HpcSimd,
                               no_minimizer_stream=flag)
        t = lambda p, c, isz: torch.as_tensor(S.DeviceArray(p, c * isz, "|u1"), device=dev).clone()
        outs.append((int(r.n_items), t(r.hash, r.n_items, 8), t(r.start, r.n_items, 4), t(r.end, r.n_items, 4),
                     t(r.rev, r.n_items, 1), t(r.km_off, n + 1, 8), t(r.min_off, n + 1, 8), t(r.min_cnt, n, 4)))
    assert outs[0][0] == outs[1][0] > 0
    for a, b in zip(outs[0][1:], outs[1][1:]):
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_h64_flavour(S, O, gpu_ctx, batches, fixture_seq):
    """H = u64: KAT-2 of the reference (tests/main.rs:18-39) through the kernels, closed-form parity."""
    from parity_cases import check_h64_flavour
    check_h64_flavour(S, O, gpu_ctx, batches, fixture_seq, scale=4)


@pytest.mark.gpu
def test_h16_flavour(S, O, gpu_ctx, batches, fixture_seq):
    """H = u16 (parity unpinned): closed-form parity in the modes Hpc and Regular."""
    from parity_cases import check_h16_flavour
    check_h16_flavour(S, O, gpu_ctx, batches, fixture_seq, scale=4)


@pytest.mark.gpu
def test_one_pass_per_run_at_baseline_densities(S, gpu_ctx):
    """run_device sizes its buffers from the expected selection rate and reruns a batch whose records overflow them.
    At BASELINE's densities that must not happen (HPC off at d = 0.01 once overflowed a per-CTA append region on every
    run and silently doubled the step): one run = 8 kernel launches, in every mode, on 1.2 Gbp of config-2 shaped reads."""
    import torch
    L, n = 20000, 60000
    dev = torch.device("cuda:0")
    d_bases = torch.empty(L * n, dtype=torch.uint8, device=dev)
    gpu_ctx.synth_device(0x5EED0005, 0, L * n, d_bases.data_ptr())
    d_so = (torch.arange(n + 1, dtype=torch.int64, device=dev) * L)
    torch.cuda.synchronize()
    for mode in (S.HashMode.HpcSimd, S.HashMode.Simd, S.HashMode.Hpc, S.HashMode.Regular):
        for d in (0.001, 0.01):
            for in_place in (True, False):
                before = gpu_ctx.launch_count
                r = gpu_ctx.run_device(d_bases.data_ptr(), d_so.data_ptr(), n, n * L, 31, 5, d, mode, no_minimizer_stream=in_place)
                assert r.n_items > 0 and gpu_ctx.launch_count - before == 8, (mode, d, in_place, gpu_ctx.launch_count - before)
