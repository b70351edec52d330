"""The `miekki` command line on the CPU: the unmodified CLI source linked against a stand-in for
libmiekki_b200.so that answers the C ABI from the oracle (tests/cpp/abi_stub.c -- test
infrastructure).  This checks everything the host side owns -- flags and banners, FASTA record
rules, read batching, the heap chain over genome shards, hit-line and exact-line text, the
streamed gz dump and its loader, the <dump>.names side-car -- against the reference binary's
golden outputs without a GPU.  The same scenarios run against the real library in
tests/test_gpu_cli.py."""
import os
import subprocess
from collections import Counter

import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers as H

D = os.path.join(H.GOLDEN, "caseA")


@pytest.fixture(scope="module")
def cli(tmp_path_factory):
    out = tmp_path_factory.mktemp("stub")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    lib = out / "libmiekki_b200.so"
    subprocess.check_call([cc, "-O2", "-shared", "-fPIC", "-o", str(lib),
                           os.path.join(H.ROOT, "tests", "cpp", "abi_stub.c"),
                           os.path.join(H.ROOT, "oracle", "miekki_oracle.c"), "-lm", "-lpthread"])
    exe = out / "miekki"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fopenmp", "-I" + os.path.join(H.ROOT, "include"),
                           "-o", str(exe), os.path.join(H.ROOT, "miekki_b200", "cli", "miekki_cli.cpp"),
                           "-L" + str(out), "-lmiekki_b200", "-lz", "-Wl,-rpath," + str(out)])
    return str(exe)


def run(cli, args, env=None):
    r = subprocess.run([cli] + [str(a) for a in args], cwd=D, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, **env) if env else None)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.parametrize("s", [200, 0, 5000])
def test_hit_lines(cli, tmp_path, s):
    out = tmp_path / "hits.txt"
    stdout = run(cli, ["-l", "list.txt", "-a", "reads.fa", "-k", 31, "-h", 12, "-t", 4, "-s", s, "-o", out])
    assert "Reference indexed: 14" in stdout and stdout.count("elapsed time:") == 2
    assert "Using 8 bits per minimizer, 4,096 minimizers so 32,768 bits per sequences" in stdout
    assert out.read_text() == open(os.path.join(D, "hits_s%d.txt" % s)).read()


@pytest.mark.parametrize("devices", ["0", "0,0,0"])
def test_dump_in_slabs_load_and_names(cli, tmp_path, devices):
    """-d streams the matrix (here in 58 slabs of 1,000 bytes) from 1 or 3 shards; the payload
    equals the reference's dump, and -i (again sharded, again in slabs) answers the same."""
    env = {"MIEKKI_DUMP_SLAB_BYTES": "1000"}
    dump = tmp_path / "idx.gz"
    run(cli, ["-l", "list.txt", "-k", 31, "-h", 12, "-t", 4, "-d", dump, "-o", tmp_path / "o.txt",
              "--devices", devices], env)
    got = orc.parse_dump(str(dump))
    z = H.load_dump_npz(os.path.join(D, "dump.npz"))
    assert (got.k, got.h, got.nbm, got.nbmant, got.n, got.b) == (31, 12, 8, 5, 14, 33)
    assert got.bloom_bits == 1 << 33 and got.threshold == 200 and got.compressed == 1
    assert np.array_equal(got.rows, z["rows"])
    assert np.array_equal(got.genome_size, z["genome_size"]) and np.array_equal(got.sketch_size, z["sketch_size"])
    idx, val = H.bloom_nonzero(got.bloom)
    assert np.array_equal(idx, z["bloom_idx"]) and np.array_equal(val, z["bloom_val"])
    assert os.path.exists(str(dump) + ".names")
    out = tmp_path / "hits.txt"
    stdout = run(cli, ["-i", dump, "-a", os.path.join(D, "reads.fa"), "-o", out, "-k", 21, "-s", 1,
                       "--devices", devices], env)       # -k / -s are ignored with -i (quirk G6)
    assert "Load sucessful" in stdout
    assert out.read_text() == open(os.path.join(D, "hits_s200.txt")).read()
    # the reference binary reads the same file (its zlib skips the members' size records)
    ref = orc.RefBinary()
    if ref.available:
        outr = tmp_path / "ref_hits.txt"
        ref.run(["-i", str(dump), "-a", os.path.join(D, "reads.fa"), "-t", "1", "-o", str(outr)], cwd=D)
        assert outr.read_text() == open(os.path.join(D, "hits_s200.txt")).read()
    # exact mode from a loaded index works thanks to the side-car (the reference crashes: quirk G4)
    oute = tmp_path / "exact.txt"
    run(cli, ["-i", dump, "-a", os.path.join(D, "reads.fa"), "-e", "-o", oute, "--devices", devices])
    want = [l for l in open(os.path.join(D, "exact.txt")).read().split("\n") if l]
    assert Counter(l for l in oute.read_text().split("\n") if l) == Counter(want)


@pytest.mark.parametrize("devices", ["0", "0,0,0,0,0"])
def test_exact_and_whole_file_modes(cli, tmp_path, devices):
    for s, fname in ((200, "exact.txt"), (0, "exact_s0.txt")):
        out = tmp_path / ("e%d.txt" % s)
        run(cli, ["-l", "list.txt", "-a", "reads.fa", "-k", 31, "-h", 12, "-e", "-s", s, "-o", out,
                  "--devices", devices])
        want = [l for l in open(os.path.join(D, fname)).read().split("\n") if l]
        assert Counter(l for l in out.read_text().split("\n") if l) == Counter(want)
    outa = tmp_path / "a.txt"
    run(cli, ["-l", "list.txt", "-A", "alist.txt", "-k", 31, "-h", 12, "-t", 3, "-o", outa, "--devices", devices])
    assert outa.read_text() == open(os.path.join(D, "hits_A.txt")).read()
    oute = tmp_path / "ae.txt"
    run(cli, ["-l", "list.txt", "-A", "alist.txt", "-k", 31, "-h", 12, "-e", "-o", oute, "--devices", devices])
    want = [l for l in open(os.path.join(D, "exact_A.txt")).read().split("\n") if l]
    assert Counter(l for l in oute.read_text().split("\n") if l) == Counter(want) and len(want) > 10


def test_sharded_hit_lines_with_all_ties(cli, tmp_path):
    """-s 0 keeps every candidate, so the lists are decided by the heap's tie order: five uneven
    shards chained in id order must reproduce the single-shard file byte for byte."""
    out = tmp_path / "h.txt"
    run(cli, ["-l", "list.txt", "-a", "reads.fa", "-k", 31, "-h", 12, "-s", 0, "-o", out, "--devices", "0,0,0,0,0"])
    assert out.read_text() == open(os.path.join(D, "hits_s0.txt")).read()


def test_case_d_seventy_genomes_over_shards(cli, tmp_path):
    """70 genome files (more than one parse wave of the build), 1 and 4 shards: hit lines, dump
    and exact-mode lines against the reference's."""
    d, _ = H.case_d(str(tmp_path))
    z = H.load_dump_npz(os.path.join(d, "dump.npz"))
    want_exact = Counter(l for l in open(os.path.join(d, "exact.txt")).read().split("\n") if l)
    for devices in ("0", "0,0,0,0"):
        out, dump = tmp_path / "h.txt", tmp_path / "i.gz"
        r = subprocess.run([cli, "-l", "list.txt", "-a", os.path.join(d, "reads.fa"), "-k", "31", "-h", "12", "-t", "3",
                            "-o", str(out), "-d", str(dump), "--devices", devices], cwd=tmp_path,
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and "Reference indexed: 70" in r.stdout, r.stdout[-1000:] + r.stderr[-1000:]
        assert out.read_text() == open(os.path.join(d, "hits_s200.txt")).read()
        got = orc.parse_dump(str(dump))
        assert np.array_equal(got.rows, z["rows"]) and np.array_equal(got.genome_size, z["genome_size"])
        assert np.array_equal(got.sketch_size, z["sketch_size"])
        idx, val = H.bloom_nonzero(got.bloom)
        assert np.array_equal(idx, z["bloom_idx"]) and np.array_equal(val, z["bloom_val"])
        oute = tmp_path / "e.txt"
        r = subprocess.run([cli, "-l", "list.txt", "-a", os.path.join(d, "reads.fa"), "-k", "31", "-h", "12", "-e",
                            "-o", str(oute), "--devices", devices], cwd=tmp_path, capture_output=True, text=True,
                           timeout=900)
        assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
        assert Counter(l for l in oute.read_text().split("\n") if l) == want_exact


@pytest.mark.parametrize("batch", [1, 7])
def test_query_pipeline_keeps_the_order_of_the_file(cli, tmp_path, batch):
    """query_file overlaps parsing, scoring and writing of consecutive read batches; with 1 or 7
    reads per batch (MIEKKI_QUERY_BATCH_READS) the 56 lines must still come out in file order."""
    out = tmp_path / "hits.txt"
    stdout = run(cli, ["-l", "list.txt", "-a", "reads.fa", "-k", 31, "-h", 12, "-s", 0, "-o", out],
                 {"MIEKKI_QUERY_BATCH_READS": str(batch)})
    assert out.read_text() == open(os.path.join(D, "hits_s0.txt")).read()
    n_reads = open(os.path.join(D, "hits_s0.txt")).read().count("\n")
    assert stdout.count("-") >= (n_reads + batch - 1) // batch          # one tick per batch (:345)


@pytest.mark.parametrize("devices", ["0", "0,0"])
def test_loads_reference_style_dump(cli, tmp_path, devices):
    """A dump written the way the reference writes it: one gzip member, no size records, the
    full 1 GiB Bloom table (read past beyond the window), a garbage jaccard_estimation byte."""
    import gzip
    z = H.load_dump_npz(os.path.join(D, "dump.npz"))
    bloom = np.zeros((1 << 33) // 8, np.uint8)
    bloom[z["bloom_idx"]] = z["bloom_val"]
    p = tmp_path / "ref.gz"
    with gzip.open(p, "wb", compresslevel=1) as f:
        f.write(np.array([31, 12, 8, 5, 14, 33], "<u4").tobytes())
        f.write(np.array([1 << 33], "<u8").tobytes())
        f.write(bytes([7, 0]))
        f.write(np.array([200], "<u4").tobytes() + bytes([1]))
        f.write(z["rows"].tobytes() + z["genome_size"].astype("<u8").tobytes())
        f.write(bloom.tobytes())
        f.write(z["sketch_size"].astype("<u4").tobytes())
    del bloom
    out = tmp_path / "hits.txt"
    run(cli, ["-i", p, "-a", os.path.join(D, "reads.fa"), "-o", out, "--devices", devices])
    assert out.read_text() == open(os.path.join(D, "hits_s200.txt")).read()
    # a truncated file is reported, not loaded
    raw = open(p, "rb").read()
    (tmp_path / "cut.gz").write_bytes(raw[: len(raw) // 2])
    r = subprocess.run([cli, "-i", str(tmp_path / "cut.gz"), "-a", os.path.join(D, "reads.fa"), "-o", str(out)],
                       cwd=D, capture_output=True, text=True)
    assert r.returncode != 0


@pytest.mark.parametrize("devices", ["0", "0,0"])
def test_merge_of_two_dumps_equals_one_build(cli, tmp_path, devices):
    H.merge_scenario(cli, tmp_path, devices)


def test_unreadable_input_is_an_error_not_a_crash(cli, tmp_path):
    """A genome file that is not what its gzip header promises: `miekki: ...` on stderr and exit
    code 1 (the reader's exception may not escape an OpenMP region or a std::async task)."""
    import gzip
    good = gzip.compress(b">g\n" + b"ACGT" * 5000 + b"\n")
    (tmp_path / "bad.fa.gz").write_bytes(good[:200] + bytes(300))
    (tmp_path / "list.txt").write_text(str(tmp_path / "bad.fa.gz") + "\n")
    r = subprocess.run([cli, "-l", str(tmp_path / "list.txt"), "-k", "31", "-h", "10", "-o", str(tmp_path / "o.txt")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "miekki:" in r.stderr and "The end" not in r.stdout
    # a dump that cannot be written (its directory does not exist) fails the run as well
    r = subprocess.run([cli, "-l", "list.txt", "-k", "31", "-h", "10", "-d", str(tmp_path / "nodir" / "x.gz"),
                        "-o", str(tmp_path / "o.txt")], cwd=D, capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "miekki:" in r.stderr and "The end" not in r.stdout


def test_messages(cli, tmp_path):
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 0 and "-l" in r.stdout                       # no arguments: help, exit(0)
    r = subprocess.run([cli, "-a", "reads.fa"], cwd=D, capture_output=True, text=True)
    assert "What am I supposed to index" in r.stdout and r.returncode == 0
    r = subprocess.run([cli, "-l", "list.txt", "-f", "11", "-o", str(tmp_path / "x")], cwd=D, capture_output=True,
                       text=True)
    assert "not implemented" in r.stdout and r.returncode == 0          # quirk G12
    stdout = run(cli, ["-l", "list.txt", "-k", 31, "-h", 12, "-o", tmp_path / "y"])
    assert "No query file, No queries" in stdout and "The end" in stdout
