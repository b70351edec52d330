"""CPU-only tests of the CLI's host logic (text ingestion + dump writer): gz / zlib / plain
auto-detection, std::getline semantics at EOF, the reference's record rules, and the
multi-member gzip dump writer read back by Python's gzip and by our own reader."""
import gzip
import os
import subprocess
import zlib

import pytest

from oracle import oracle as orc
from tests import helpers as H

SRC = os.path.join(H.ROOT, "tests", "cpp", "fasta_host_test.cpp")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("bin") / "fasta_host_test"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O1", "-std=c++17", "-fopenmp", "-o", str(out), SRC, "-lz"])
    return str(out)


def run(exe, *args):
    return subprocess.run([exe] + [str(a) for a in args], capture_output=True, check=True).stdout


TEXT = b">r1 first\nACGT\nACGTT\n\n>r2\nGG\n>r3 no newline at end\nTTTTACGTAC"


@pytest.mark.parametrize("kind", ["plain", "gzip", "zlib", "multi-member"])
def test_line_reader_formats_and_eof(exe, tmp_path, kind):
    p = tmp_path / "x.fa"
    if kind == "plain":
        p.write_bytes(TEXT)
    elif kind == "gzip":
        p.write_bytes(gzip.compress(TEXT))
    elif kind == "zlib":
        p.write_bytes(zlib.compress(TEXT))
    else:
        p.write_bytes(gzip.compress(TEXT[:20]) + gzip.compress(TEXT[20:]))
    got = run(exe, "lines", p).decode().split("\n")[:-1]
    want = ["%d:%s" % (i, l) for i, l in enumerate(TEXT.decode().split("\n"))]
    assert got == want
    # a trailing newline yields one extra empty line, like while(!eof) getline in the reference
    p.write_bytes(TEXT + b"\n")
    got = run(exe, "lines", p).decode().split("\n")[:-1]
    assert got == want + ["%d:" % len(want)]


def test_genome_concat_and_records_follow_the_reference(exe, tmp_path):
    p = tmp_path / "g.fa"
    p.write_bytes(TEXT)
    assert run(exe, "concat", p) == H.genome_like_reference(str(p))
    for k in (2, 5, 9, 11, 30):
        want = orc.records_like_reference(TEXT.decode(), k)      # incl. quirk G16 (short record kept)
        got = [l for l in run(exe, "records", p, k).split(b"\n") if l]
        assert got == want, k
    # the golden multi-record genome (gz, wrapped lines, a record shorter than k)
    g4 = os.path.join(H.GOLDEN, "caseA", "gA4.fa.gz")
    assert run(exe, "concat", g4) == H.genome_like_reference(g4)
    want = orc.records_like_reference(H.read_text(g4), 31)
    assert [l for l in run(exe, "records", g4, 31).split(b"\n") if l] == want and len(want) == 2


@pytest.mark.parametrize("threads", [1, 4])
@pytest.mark.parametrize("n", [0, 5, 100_000, 70_000_000])
def test_parallel_gzip_writer_round_trips(exe, tmp_path, n, threads):
    """Members carry their sizes in a gzip FEXTRA field (like BGZF): any gzip reader skips it,
    ours inflates `threads` members at a time (threads = 1: plain sequential inflate)."""
    p = tmp_path / "d.gz"
    assert run(exe, "gzwrite", p, n, threads).strip() == b"roundtrip-ok"
    data = gzip.open(p, "rb").read()                 # concatenated members are one gzip file
    assert len(data) == n + 8 and data[:4] == b"HEAD" and data[-4:] == b"TAIL"
    raw = open(p, "rb").read()
    assert raw[:4] == b"\x1f\x8b\x08\x04" and raw[12:16] == b"MK\x08\x00"       # FEXTRA size record
    if n > 64 << 20:
        assert raw.count(b"\x1f\x8b\x08\x04") >= 3                 # several members


def test_corrupt_member_is_reported(exe, tmp_path):
    p = tmp_path / "d.gz"
    run(exe, "gzwrite", p, 100_000, 4)
    raw = bytearray(open(p, "rb").read())
    raw[len(raw) // 2] ^= 0x5A
    p.write_bytes(bytes(raw))
    r = subprocess.run([exe, "gzwrite-readonly", str(p), "100000", "4"], capture_output=True)
    assert r.returncode != 0 or b"roundtrip-ok" not in r.stdout


def test_host_sequence_packer(tmp_path):
    """miekki_b200/csrc/pack.cpp (2 bits per base for mk_index_add's upload): every word is either
    packed exactly or listed as an exception with its raw bytes -- AVX2 and portable kernels,
    random k, dirty and clean sequences, sub-ranges (tests/cpp/pack_host_test.cpp)."""
    out = tmp_path / "pack_host_test"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-o", str(out),
                           os.path.join(H.ROOT, "tests", "cpp", "pack_host_test.cpp"),
                           os.path.join(H.ROOT, "miekki_b200", "csrc", "pack.cpp")])
    r = subprocess.run([str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout + r.stderr
