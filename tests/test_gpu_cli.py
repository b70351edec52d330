"""The `miekki` command line (C++ host over the C ABI) against the reference binary's golden
outputs: same flags, same hit lines, same exact-mode lines, and a gz dump the reference's own
format parser reads back byte for byte (header byte 32 masked: quirk G7)."""
import os
import subprocess
from collections import Counter

import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu

CLI = os.path.join(H.ROOT, "miekki_b200", "cli", "miekki")


def run_cli(args, cwd, env=None):
    assert os.path.exists(CLI), "build the CLI first: make"
    r = subprocess.run([CLI] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, **env) if env else None)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.parametrize("s", [200, 0, 5000])
def test_cli_build_query_dump(tmp_path, s):
    d = os.path.join(H.GOLDEN, "caseA")
    out, dump = tmp_path / "hits.txt", tmp_path / "idx.gz"
    stdout = run_cli(["-l", "list.txt", "-a", "reads.fa", "-k", 31, "-h", 12, "-t", 4, "-s", s,
                      "-o", out, "-d", dump], d)
    assert "Reference indexed: 14" in stdout and stdout.count("elapsed time:") == 2
    assert "Using 8 bits per minimizer, 4,096 minimizers so 32,768 bits per sequences" in stdout
    assert out.read_text() == open(os.path.join(d, "hits_s%d.txt" % s)).read()
    if s == 200:
        got = orc.parse_dump(str(dump))
        z = H.load_dump_npz(os.path.join(d, "dump.npz"))
        assert (got.k, got.h, got.nbm, got.nbmant, got.n, got.b) == (31, 12, 8, 5, 14, 33)
        assert got.bloom_bits == 1 << 33 and got.threshold == 200 and got.compressed == 1
        assert np.array_equal(got.rows, z["rows"])
        assert np.array_equal(got.genome_size, z["genome_size"])
        assert np.array_equal(got.sketch_size, z["sketch_size"])
        idx, val = H.bloom_nonzero(got.bloom)
        assert np.array_equal(idx, z["bloom_idx"]) and np.array_equal(val, z["bloom_val"])
        # -i: load our own dump, query again (flags -k/-h/-s ignored: quirk G6)
        out2 = tmp_path / "hits2.txt"
        stdout = run_cli(["-i", dump, "-a", os.path.join(d, "reads.fa"), "-o", out2, "-k", 21, "-s", 1], d)
        assert "Load sucessful" in stdout
        assert out2.read_text() == open(os.path.join(d, "hits_s200.txt")).read()


@pytest.mark.parametrize("devices", ["0", "0,0,0"])
def test_cli_dump_and_load_in_many_slabs(tmp_path, devices):
    """-d / -i stream the matrix a slab of bucket rows at a time (256 MB by default).  A 1,000-byte
    slab cuts this 4,096 x 14 matrix into 58 slabs: the payload and the answers must not change."""
    d = os.path.join(H.GOLDEN, "caseA")
    env = {"MIEKKI_DUMP_SLAB_BYTES": "1000"}
    run_cli(["-l", "list.txt", "-k", 31, "-h", 12, "-t", 4, "-d", tmp_path / "whole.gz"], d)
    run_cli(["-l", "list.txt", "-k", 31, "-h", 12, "-t", 4, "-d", tmp_path / "slabs.gz", "--devices", devices], d, env)
    assert orc.read_payload(str(tmp_path / "whole.gz")) == orc.read_payload(str(tmp_path / "slabs.gz"))
    out = tmp_path / "hits.txt"
    run_cli(["-i", tmp_path / "slabs.gz", "-a", os.path.join(d, "reads.fa"), "-o", out, "--devices", devices], d, env)
    assert out.read_text() == open(os.path.join(d, "hits_s200.txt")).read()


@pytest.mark.parametrize("s,fname", [(200, "exact.txt"), (0, "exact_s0.txt")])
def test_cli_exact_mode(tmp_path, s, fname):
    d = os.path.join(H.GOLDEN, "caseA")
    out = tmp_path / "exact.txt"
    run_cli(["-l", "list.txt", "-a", "reads.fa", "-k", 31, "-h", 12, "-e", "-s", s, "-o", out], d)
    got = [l for l in out.read_text().split("\n") if l]
    want = [l for l in open(os.path.join(d, fname)).read().split("\n") if l]
    assert Counter(got) == Counter(want)
    # `-i dump -e`: the reference crashes (file names are not in the dump, quirk G4); our -d
    # leaves a side-car <dump>.names, so exact mode also works from a loaded index
    dump, out2 = tmp_path / "idx.gz", tmp_path / "exact2.txt"
    run_cli(["-l", "list.txt", "-k", 31, "-h", 12, "-s", s, "-d", dump, "-o", tmp_path / "unused.txt"], d)
    assert os.path.exists(str(dump) + ".names")
    run_cli(["-i", dump, "-a", "reads.fa", "-e", "-o", out2], d)
    assert Counter(l for l in out2.read_text().split("\n") if l) == Counter(want)


def test_cli_whole_file_queries(tmp_path):
    """-A and -A -e: whole files as queries, long-sequence sketch path."""
    d = os.path.join(H.GOLDEN, "caseA")
    out = tmp_path / "a.txt"
    run_cli(["-l", "list.txt", "-A", "alist.txt", "-k", 31, "-h", 12, "-t", 3, "-o", out], d)
    assert out.read_text() == open(os.path.join(d, "hits_A.txt")).read()
    oute = tmp_path / "ae.txt"
    run_cli(["-l", "list.txt", "-A", "alist.txt", "-k", 31, "-h", 12, "-e", "-o", oute], d)
    got = [l for l in oute.read_text().split("\n") if l]
    want = [l for l in open(os.path.join(d, "exact_A.txt")).read().split("\n") if l]
    assert Counter(got) == Counter(want) and len(want) > 10


@pytest.mark.parametrize("devices", ["0,0", "0,0,0,0,0"])
def test_cli_sharded_outputs_are_identical(tmp_path, devices):
    """--gpus/--devices: genome shards (here several contexts on one GPU; 14 genomes over 2 and
    5 shards, uneven ranges) must reproduce every output byte of the single-GPU run: hit lines at
    -s 200 and the all-candidates -s 0, the dump (rows interleaved back, Bloom folded), -e, -A."""
    d = os.path.join(H.GOLDEN, "caseA")
    for s in (200, 0):
        out, dump = tmp_path / ("h%d.txt" % s), tmp_path / ("i%d.gz" % s)
        run_cli(["-l", "list.txt", "-a", "reads.fa", "-k", 31, "-h", 12, "-t", 4, "-s", s, "-o", out,
                 "-d", dump, "--devices", devices], d)
        assert out.read_text() == open(os.path.join(d, "hits_s%d.txt" % s)).read()
    got = orc.parse_dump(str(tmp_path / "i200.gz"))
    z = H.load_dump_npz(os.path.join(d, "dump.npz"))
    assert np.array_equal(got.rows, z["rows"])
    assert np.array_equal(got.genome_size, z["genome_size"]) and np.array_equal(got.sketch_size, z["sketch_size"])
    idx, val = H.bloom_nonzero(got.bloom)
    assert np.array_equal(idx, z["bloom_idx"]) and np.array_equal(val, z["bloom_val"])
    # load the dump onto shards again
    out2 = tmp_path / "again.txt"
    run_cli(["-i", tmp_path / "i200.gz", "-a", os.path.join(d, "reads.fa"), "-o", out2, "--devices", devices], d)
    assert out2.read_text() == open(os.path.join(d, "hits_s200.txt")).read()
    oute, outa = tmp_path / "e.txt", tmp_path / "a.txt"
    run_cli(["-l", "list.txt", "-a", "reads.fa", "-k", 31, "-h", 12, "-e", "-o", oute, "--devices", devices], d)
    want = [l for l in open(os.path.join(d, "exact.txt")).read().split("\n") if l]
    assert Counter(l for l in oute.read_text().split("\n") if l) == Counter(want)
    run_cli(["-l", "list.txt", "-A", "alist.txt", "-k", 31, "-h", 12, "-o", outa, "--devices", devices], d)
    assert outa.read_text() == open(os.path.join(d, "hits_A.txt")).read()


@pytest.mark.parametrize("devices", ["0", "0,0"])
def test_cli_merge_of_two_dumps_equals_one_build(tmp_path, devices):
    """Row f4: mk_index_merge (one shard) / shard concatenation (several) behind `-i a -i b`."""
    assert os.path.exists(CLI), "build the CLI first: make"
    H.merge_scenario(CLI, tmp_path, devices)


def test_cli_loads_reference_style_dump(tmp_path):
    """A dump written the way the reference writes it (gzip, full 1 GiB Bloom table) loads."""
    import gzip
    d = os.path.join(H.GOLDEN, "caseA")
    z = H.load_dump_npz(os.path.join(d, "dump.npz"))
    bloom = np.zeros((1 << 33) // 8, np.uint8)
    bloom[z["bloom_idx"]] = z["bloom_val"]
    p = tmp_path / "ref.gz"
    with gzip.open(p, "wb", compresslevel=1) as f:
        f.write(np.array([31, 12, 8, 5, 14, 33], "<u4").tobytes())
        f.write(np.array([1 << 33], "<u8").tobytes())
        f.write(bytes([7, 0]))                       # garbage jaccard_estimation byte
        f.write(np.array([200], "<u4").tobytes() + bytes([1]))
        f.write(z["rows"].tobytes() + z["genome_size"].astype("<u8").tobytes())
        f.write(bloom.tobytes())
        f.write(z["sketch_size"].astype("<u4").tobytes())
    out = tmp_path / "hits.txt"
    run_cli(["-i", p, "-a", os.path.join(d, "reads.fa"), "-o", out], d)
    assert out.read_text() == open(os.path.join(d, "hits_s200.txt")).read()


def test_cli_messages(tmp_path):
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 0 and "This is a help message" in r.stdout
    r = subprocess.run([CLI, "-a", "x.fa"], capture_output=True, text=True)
    assert "What am I supposed to index ?" in r.stdout
    r = subprocess.run([CLI, "-l", "nope.txt", "-f", "11", "-o", str(tmp_path / "o")], capture_output=True, text=True)
    assert "not implemented" in r.stdout            # quirk G12
