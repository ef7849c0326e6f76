"""CPU-only checks of the product's host side: the byte-defining shape code and block merge against the
oracle, the C-ABI library (loads, exports every symbol include/gcz.h declares, refuses to compute without a
GPU), headers, FASTA records.  No compute entry point is exercised here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import gecoz_b200 as G
from gecoz_b200 import _native as N
from gecoz_b200 import synth
from gecoz_b200.geco_index import FastaSequence, merge_blocks, read_fasta
from oracle import gcz_oracle as O

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    lib = N.lib()
    for header_name, exports in (("gcz.h", N.EXPORTS), ("gcz_file.h", N.FILE_EXPORTS)):
        header = (ROOT / "include" / header_name).read_text()
        declared = set(re.findall(r"^\s*(?:int|void|int32_t|int64_t|const char\*)\s+(gcz_[a-z0-9_]+)\s*\(", header, re.M))
        assert declared == set(exports), header_name
        for name in declared:
            assert hasattr(lib, name), name
    assert b"sm_100a" in lib.gcz_version()


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    text = np.frombuffer(b"ACGT\0", np.uint8).copy()
    counts = np.zeros(256, np.int64)
    rc = N.lib().gcz_count_symbols(0, N.ptr(text), len(text), N.ptr(counts))
    assert rc == N.GCZ_E_NODEVICE
    with pytest.raises(G.GczError, match="no CPU path"):
        N.check(rc)


def _same_shape(counts):
    s, o = G.shape_from_counts(counts), O.shape_from_counts(counts)
    assert list(s.bit_lengths) == list(o.bit_lengths)
    assert list(s.codes) == list(o.table)
    assert (s.size, s.table_bytes, s.length) == (o.size, o.table_bytes, o.length)
    for k in range(s.n_nodes):
        assert s.node_bits[k] == o.node_bits[s.node_name[k]]
    assert sum(1 for b in o.node_bits if b > 0) == s.n_nodes
    buf = np.zeros(s.table_bytes + 8, np.uint8)
    w = N.lib().gcz_shape_write(C.byref(s), N.ptr(buf), len(buf))
    assert w == s.table_bytes and buf[:w].tobytes() == O.shape_write(o)
    r = G.Shape()
    assert N.lib().gcz_shape_read(N.ptr(buf), w, C.byref(r)) == 0
    assert list(r.bit_lengths) == list(s.bit_lengths) and list(r.codes) == list(s.codes) and r.table_bytes == w
    assert [r.node_name[k] for k in range(r.n_nodes)] == [s.node_name[k] for k in range(s.n_nodes)]
    return s


def test_shape_matches_oracle_dna():
    t = synth.cfg1_text(200_000)
    s = _same_shape(np.bincount(t, minlength=256).astype(np.int64))
    assert s.n_nodes == 5
    t = synth.block_of([np.frombuffer(b"ACGTNacgtnRYKM", np.uint8)[np.random.default_rng(1).integers(0, 14, 50_000)]])
    _same_shape(np.bincount(t, minlength=256).astype(np.int64))


def test_shape_matches_oracle_random_alphabets():
    rng = np.random.default_rng(0)
    refused = 0
    for trial in range(250):
        k = int(rng.integers(2, 120))
        c = np.zeros(256, np.int64)
        idx = rng.choice(256, k, replace=False)
        mode = trial % 4
        if mode == 0:
            c[idx] = rng.integers(1, 1000, k)
        elif mode == 1:
            c[idx] = rng.integers(1, 5, k)
        elif mode == 2:
            c[idx] = (2.0 ** rng.integers(0, 22, k)).astype(np.int64)
        else:
            f = [1, 1]
            while len(f) < k:
                f.append(min(f[-1] + f[-2], 2 ** 31 // (2 * k)))
            c[idx] = np.array(f[:k], dtype=np.int64)      # Fibonacci weights: lengths above 15 get limited
        c[0] = max(c[0], 1)
        try:
            _same_shape(c)
        except G.GczError as e:
            # the reference sizes the length table with a 15-bit code-length code and writes it with a 7-bit one;
            # when they differ its own block overflows, and the product refuses
            assert "length table size mismatch" in str(e)
            refused += 1
    assert refused < 40


def test_single_symbol_alphabets():
    c = np.zeros(256, np.int64)
    c[0] = 9
    s = _same_shape(c)
    assert s.n_nodes == 1 and s.node_bits[0] == 9
    c = np.zeros(256, np.int64)
    c[65] = 9
    with pytest.raises(G.GczError):          # the reference hits a null node here
        G.shape_from_counts(c)


def test_sizes():
    for n in (1, 8, 511, 512, 513, 65536, 65537, 16_000_001, 248_956_423):
        assert N.lib().gcz_ranked_bytes(n) == O.ranked_bytes(n)
        for f in (0, 1, 5, 7):
            assert N.lib().gcz_index_size(n, f) == O.index_size(n, f)
    assert N.lib().gcz_index_size(16_000_001, 5) == 3_289_370
    assert N.lib().gcz_index_size(248_956_423, 5) == 55_197_282


def test_merge_matches_oracle():
    rng = np.random.default_rng(3)
    seqs = [FastaSequence(h, l, None, i) for i, (h, l) in enumerate(zip(synth.HG38_NAMES, synth.HG38_LENGTHS))]
    got = [[s.id for s in b.sequences] for b in merge_blocks(seqs)]
    assert got == O.merge_blocks(synth.HG38_LENGTHS, synth.HG38_NAMES) and len(got) == 18
    for trial in range(200):
        k = int(rng.integers(1, 40))
        lens = (rng.integers(1, 2000, k) if trial % 2 else rng.integers(1, 6, k)).tolist()
        hdrs = [f"s{rng.integers(0, k)}" if trial % 3 == 0 else f"s{i}" for i in range(k)]     # duplicates collapse
        seqs = [FastaSequence(h, l, None, i) for i, (h, l) in enumerate(zip(hdrs, lens))]
        assert [[s.id for s in b.sequences] for b in merge_blocks(seqs)] == O.merge_blocks(lens, hdrs)


def test_headers_match_oracle():
    hs = ["chr13 Homo sapiens", "chr14"]
    h = G.GecozRefBlockHeader(hs, 123456789, 987654)
    assert h.to_bytes() == O.ref_header(hs, 123456789, 987654)
    assert h.getHeaderHash() == O.header_hash(hs)
    assert G.GecozSSABlockHeader(hs, 77).to_bytes() == O.ssa_header(hs, 77)
    p = G.GecozRefBlockHeader.parse(h.to_bytes() + b"rest")
    assert (p.headers, p.size, p.len) == (hs, 123456789, 987654)
    assert h.findHeader("chr14") == 1 and h.findHeader("chr1") == -1


def test_fasta_records(tmp_path):
    fa = tmp_path / "t.fa"
    fa.write_bytes(b">s1 desc\r\nACGT\r\nacgu\r\n\r\n>s2\nNNNN\n@r1\nACGT\n+\nIIII\n>s3\nA\n")
    assert list(read_fasta(fa)) == [("s1 desc", b"ACGTacgu"), ("s2", b"NNNN"), ("r1", b"ACGT"), ("s3", b"A")]


def test_synthetic_workloads_are_seeded():
    a, b = synth.cfg2_text(100_000), synth.cfg2_text(100_000)
    assert np.array_equal(a, b) and a[-1] == 0 and (a == ord("N")).sum() > 7000
    d1, o1 = synth.patterns(a, 1000, 15, 100, seed=5)
    d2, o2 = synth.patterns(a, 1000, 15, 100, seed=5)
    assert np.array_equal(d1, d2) and np.array_equal(o1, o2) and set(np.unique(d1)) <= set(b"ACGT")
    assert len(synth.hg38_shaped_records(1e-5)) == 25


# ---- callers' host logic (no device needed) ---------------------------------------------------------------------------
def test_gff_attribute_column_follows_java_split():
    """String.split("\\|") drops trailing empty strings (tools/SimpleGFFGenerator.java:146-154)."""
    from gecoz_b200.geco_match import _attributes
    assert _attributes("abc") == "ID=abc"
    assert _attributes("a|b|c") == "ID=a;Note=b;Note=c"
    assert _attributes("a||") == "ID=a"
    assert _attributes("|a") == "ID=;Note=a"
    assert _attributes("") == "ID="
    assert _attributes("|") == ""


def test_pattern_file_records(tmp_path):
    """The record loop of SimpleGFFGenerator.search :59-86: '>' / '@' open, '+' closes, empty records are dropped."""
    from gecoz_b200.geco_match import read_patterns
    p = tmp_path / "p.fq"
    p.write_bytes(b"ignored before any header\n>r1 desc\nACGT\r\nAC\n>empty\n@r2\nGGU\n+\nIII\n@IIIquality-looking header\nTT\n>r3\n\nA\n")
    assert list(read_patterns(p)) == [("r1 desc", b"ACGTAC"), ("r2", b"GGU"), ("IIIquality-looking header", b"TT"), ("r3", b"A")]


def test_fasta_record_layout():
    """fasta/FastaFileWriter.java:132-215 for a multi-line sequence: a break after every 50 symbols and one more."""
    from gecoz_b200.geco_read import fasta_record_bytes
    for n in (49, 50, 51, 100, 149, 150):
        seq = np.frombuffer(bytes((65 + i % 4) for i in range(n)), np.uint8)
        rec = fasta_record_bytes("h d", seq)
        assert rec.startswith(b">h d\n")
        body = rec[len(b">h d\n"):]
        assert len(body) == n + n // 50 + 1
        lines = body.split(b"\n")
        assert b"".join(lines) == seq.tobytes()
        assert all(len(x) == 50 for x in lines[:n // 50])


def test_headers_are_plain_c_and_the_c_example_links(tmp_path):
    """include/*.h must be usable from C (the Java/JNI side and any other FFI bind the same symbols); the example
    program links against the library and, without a device, fails loudly instead of computing on the CPU."""
    import shutil
    import subprocess
    import torch
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    for h in ("gcz.h", "gcz_file.h"):
        subprocess.run([gcc, "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                        str(ROOT / "include" / h)], check=True)
    exe = tmp_path / "index_and_count"
    subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "index_and_count.c"),
                    f"-L{ROOT / 'gecoz_b200'}", "-lgcz_b200", f"-Wl,-rpath,{ROOT / 'gecoz_b200'}", "-o", str(exe)], check=True)
    fa = tmp_path / "t.fa"
    fa.write_bytes(b">s\nACGTACGTTTGACA\n")
    r = subprocess.run([str(exe), str(fa), str(tmp_path / "t.gcz"), "ACGT"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and ">s found : 2" in r.stdout
    else:
        assert r.returncode == 1 and "no CPU path" in r.stderr


def test_native_host_layer_under_thread_sanitizer(tmp_path):
    """The host layer alone (FASTA scan and assembly threads, the writer's buffer pools and block workers) compiled with
    -fsanitize=thread against stand-ins for the CUDA runtime and entry points (tests/host_tsan): no data race reports."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    cuda_inc = Path("/usr/local/cuda/include")
    if gxx is None or not (cuda_inc / "cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA headers")
    src = ROOT / "tests" / "host_tsan"
    csrc = ROOT / "gecoz_b200" / "csrc"
    exe = tmp_path / "host_tsan"
    build = subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fsanitize=thread", f"-I{csrc}", f"-I{ROOT / 'include'}", f"-I{cuda_inc}",
                            "-o", str(exe), str(src / "main.cpp"), str(src / "stubs.cpp"), str(csrc / "host_file.cpp"), str(csrc / "shape.cpp"),
                            "-ldl", "-lpthread"], capture_output=True, text=True)
    if build.returncode != 0 and "tsan" in build.stderr.lower():
        pytest.skip("no ThreadSanitizer runtime")
    assert build.returncode == 0, build.stderr
    r = subprocess.run([str(exe), str(tmp_path / "o.gcz")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, r.stderr[-4000:]
    assert r.stdout.count("rc 0 blocks") == 2


def test_synthetic_workloads_are_pinned():
    """bench.py, tools/ and the committed profiles quote numbers on these generators: a change of their output (a numpy
    upgrade, an edit of synth.py) must be noticed, not silently move the workload."""
    import hashlib
    from gecoz_b200 import synth
    text = synth.cfg2_text(3_000_000, seed=3)
    data, off = synth.patterns(text, 10_000, 15, 100, seed=5)
    digest = lambda a: hashlib.sha256(a.tobytes()).hexdigest()[:16]
    assert (digest(text), digest(data), digest(off)) == ("165caa614f83a843", "61a4ceecdc0a6e7d", "0584b6814f516ac9")
