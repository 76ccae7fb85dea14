"""Host ingest: s2k_run_fastx parses FASTA/FASTQ on several host threads into pinned memory and runs
the batch.  CPU tier: through the host-emulated library (the parser is plain host C++); GPU tier: the product library."""
import numpy as np
import pytest

from conftest import assert_batch_matches_oracle


def write_fasta(path, seqs, width, crlf=False, trailing_newline=True, blank_lines=False):
    nl = "\r\n" if crlf else "\n"
    parts = []
    for i, s in enumerate(seqs):
        parts.append(f">read{i} some description{nl}")
        t = bytes(s).decode()
        if width:
            for a in range(0, len(t), width):
                parts.append(t[a:a + width] + nl)
        else:
            parts.append(t + nl)
        if blank_lines and i % 3 == 0:
            parts.append(nl)
    text = "".join(parts)
    if not trailing_newline and text.endswith(nl):
        text = text[:-len(nl)]
    path.write_text(text, newline="")


def write_fastq(path, seqs, rng, crlf=False):
    nl = "\r\n" if crlf else "\n"
    out = []
    for i, s in enumerate(seqs):
        q = "".join(chr(33 + int(x)) for x in rng.integers(0, 42, len(s)))     # includes '@' and '+' and '>' as quality chars
        if len(q):
            q = ("@" if i % 2 else "+") + q[1:]                                # adversarial: quality lines starting with @ / +
        out.append(f"@r{i}{nl}{bytes(s).decode()}{nl}+{nl}{q}{nl}")
    path.write_text("".join(out), newline="")


def check(S, O, ctx, path, seqs, threads, mode=3):
    batch, bases, so = ctx.run_fastx(path, threads, 31, 5, 0.02, S.HashMode(mode))
    want_so = np.zeros(len(seqs) + 1, dtype=np.uint64)
    want_so[1:] = np.cumsum([len(s) for s in seqs])
    assert np.array_equal(so, want_so), (len(so), len(want_so))
    assert bases.tobytes() == b"".join(bytes(s) for s in seqs)
    assert_batch_matches_oracle(O, batch, bases, so, 31, 5, 0.02, mode)


def check_streamed(S, O, ctx, path, seqs, threads, mode=3, slab=6000):
    """The same file through the streaming form (slabs gathered and packed from the mapping, keep_bases=False)."""
    ctx.set_slab_bytes(slab)
    try:
        batch, bases, so = ctx.run_fastx(path, threads, 31, 5, 0.02, S.HashMode(mode), keep_bases=False)
    finally:
        ctx.set_slab_bytes(0)
    want_so = np.zeros(len(seqs) + 1, dtype=np.uint64)
    want_so[1:] = np.cumsum([len(s) for s in seqs])
    assert np.array_equal(so, want_so)
    want_bases = np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8)
    assert_batch_matches_oracle(O, batch, want_bases, so, 31, 5, 0.02, mode)
    return bases


def _cases(batches, tmp_path, big):
    rng = batches.rng
    lens = [0, 1, 31, 32, 150, 150, 7000, 0, 20000, 59, 60, 61, 120] + list(rng.integers(0, 400, 300))
    if big:
        lens += [300000, 150000] + [150] * 20000
    seqs = [batches.seq(n, alphabet=b"ACGT" if i % 7 else b"ACGTNacgt") for i, n in enumerate(lens)]
    files = []
    for name, kw in [("w60.fa", dict(width=60)), ("w0.fa", dict(width=0)), ("crlf.fa", dict(width=70, crlf=True)),
                     ("notrail.fa", dict(width=60, trailing_newline=False)), ("blank.fa", dict(width=80, blank_lines=True))]:
        p = tmp_path / name
        write_fasta(p, seqs, **kw)
        files.append((p, seqs))
    # lines whose ends fall where a uniform layout wants them although two of them are not of that width (a short line
    # and a long one that make up for each other): the probing scan must not take the record for uniform
    odd = [batches.seq(200), batches.seq(95), batches.seq(64)]
    text = ""
    for i, q in enumerate(odd):
        t = bytes(q).decode()
        lines = [t[a:a + 10] for a in range(0, len(t), 10)]
        if i == 0:
            lines[1:3] = [(lines[1] + lines[2])[:4], (lines[1] + lines[2])[4:]]      # 4 + 16 bases: the same number of bytes
            lines[3:5] = [(lines[3] + lines[4])[:14], (lines[3] + lines[4])[14:]]    # 14 + 6
        if i == 1:                                         # "AC\nGTACGTA\n": 9 bases in the 11 bytes of a line, every line end
            lines[6:7] = [lines[6][:2], lines[6][3:]]      # of the uniform layout in place, one more in between
            odd[i] = np.frombuffer("".join(lines).encode(), dtype=np.uint8).copy()
        text += f">odd{i}\n" + "\n".join(lines) + "\n"
    p = tmp_path / "shifty.fa"
    p.write_text(text, newline="")
    files.append((p, odd))
    fq = [s for s in seqs if len(s) > 0]
    p = tmp_path / "reads.fq"
    write_fastq(p, fq, rng)
    files.append((p, fq))
    p = tmp_path / "reads_crlf.fq"
    write_fastq(p, fq[:50], rng, crlf=True)
    files.append((p, fq[:50]))
    return files


def test_fastx_ingest_emulated(S, O, emu_ctx, batches, tmp_path):
    streamed = 0
    for path, seqs in _cases(batches, tmp_path, big=False):
        for threads in (1, 3):
            check(S, O, emu_ctx, path, seqs, threads)
        streamed += check_streamed(S, O, emu_ctx, path, seqs, 3) is None
    assert streamed >= 5                                   # blank.fa and shifty.fa have lines of uneven width: materialised after all
    with pytest.raises(S.S2KError) as e:
        emu_ctx.run_fastx(tmp_path / "missing.fa", 2, 31, 5, 0.02, S.HashMode.Hpc)
    assert e.value.status == -8
    (tmp_path / "junk.txt").write_text("hello\n")
    with pytest.raises(S.S2KError):
        emu_ctx.run_fastx(tmp_path / "junk.txt", 2, 31, 5, 0.02, S.HashMode.Hpc)
    (tmp_path / "empty.fa").write_text("")
    b, bases, so = emu_ctx.run_fastx(tmp_path / "empty.fa", 2, 31, 5, 0.02, S.HashMode.Hpc)
    assert b.n_items == 0 and len(so) == 1


@pytest.mark.gpu
def test_fastx_ingest_gpu(S, O, gpu_ctx, batches, tmp_path):
    for path, seqs in _cases(batches, tmp_path, big=True):
        for threads in (1, 8):
            check(S, O, gpu_ctx, path, seqs, threads)
        check_streamed(S, O, gpu_ctx, path, seqs, 8, slab=200000)
        check_streamed(S, O, gpu_ctx, path, seqs, 5, mode=1, slab=70000)      # long reads in pieces, scalar profile


@pytest.mark.gpu
def test_config1_fixture_through_the_driver_path(S, O, gpu_ctx, fixture_seq, tmp_path):
    """BASELINE config 1: the reference's fixture FASTA through the driver path with main.rs' parameters
    (l=31, k=5, density=0.01; src/main.rs:53-57), every HashMode, expected counts from SURVEY App. B KAT-4."""
    path = tmp_path / "ecoli.genome.100k.fa"
    path.write_text(">NZ_CP027599.1 Escherichia coli strain 97-3250 chromosome, complete genome\n" + fixture_seq.tobytes().decode())
    for mode, n_items in [(S.HashMode.Regular, 1942), (S.HashMode.Simd, 1942), (S.HashMode.Hpc, 1471), (S.HashMode.HpcSimd, 1471)]:
        batch, bases, so = gpu_ctx.run_fastx(path, 4, 31, 5, 0.01, mode)
        assert batch.n_items == n_items and len(so) == 2 and int(so[1]) == 99925
        assert_batch_matches_oracle(O, batch, bases, so, 31, 5, 0.01, int(mode))
