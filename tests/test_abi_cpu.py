"""CPU-only checks of the drop-in boundary: the shared library loads, exports exactly the
symbols include/miekki_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "miekki_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mk_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from miekki_b200 import binding
    L = binding.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "missing symbol " + n
    # and the binding declares a signature for each of them
    assert sorted(binding.SIGNATURES) == names
    assert L.mk_abi_version() == 1


def test_hit_struct_layout_matches_similarity_score():
    from miekki_b200 import binding
    # Miekki.h:27-31: {u32 genome, u32 matches, double jaccard, double intersection}
    assert binding.HIT_DTYPE.itemsize == 24
    assert binding.HIT_DTYPE.fields["jaccard"][1] == 8
    assert binding.HIT_DTYPE.fields["intersection"][1] == 16
    assert ctypes.sizeof(binding.Stats) == 13 * 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from miekki_b200 import Miekki, MiekkiError
    with pytest.raises(MiekkiError, match="no CPU fallback"):
        Miekki()


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "miekki_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|miekki_oracle|mko_", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_bloom_reach_bounds_every_canonical_kmer():
    """mk_bloom_reach (pure arithmetic, api.cu) lets the build ignore Bloom bytes no k-mer can
    probe.  Its bound -- canonical k-mer < 4^k - 4^(ceil(k/2)-1) -- is checked (a) exhaustively on
    the digit model of the two encoders (forward / reverse-strand digit pairs a character can
    produce: A (0,3), C (1,2), G (2,1), T (3,0), anything else (0,0); an invalid prefix is all
    (0,3)), and (b) on k-mers the oracle actually produces: every string of length k + 1 over
    ACGTNacgt for k = 3, 4 plus random and adversarial strings, recovered from `anc` with the
    inverse hash."""
    import itertools

    import numpy as np

    from miekki_b200 import binding
    from oracle import oracle as orc
    L, O = binding.lib(), orc.lib()

    def bound(k):
        return 4 ** k - 4 ** ((k + 1) // 2 - 1)

    for k in range(2, 32):
        for b in (32, 33, 36, 40):
            assert L.mk_bloom_reach(k, b) == ((bound(k) - 1 + 1023) >> (b + 3)) + 1
    # (a) digit model, exhaustive for k <= 7
    pairs = [(0, 3), (1, 2), (2, 1), (3, 0), (0, 0)]
    for k in range(2, 8):
        best = 0
        for combo in itertools.product(pairs, repeat=k):
            S = sum(f << (2 * (k - 1 - i)) for i, (f, _) in enumerate(combo))
            RC = sum(r << (2 * i) for i, (_, r) in enumerate(combo))
            best = max(best, min(S, RC))
        assert best < bound(k), k
    # (b) k-mers of the oracle (a fingerprint of 255 hides a k-mer at one h: try several)
    rng = np.random.default_rng(3)
    alphabet = b"ACGTNacgt"

    def canon_of(seq, k):
        out = set()
        for h in (1, 2, 3, 4):
            fp, anc, _ = orc.sketch(seq, k, h)
            for a in anc[fp != 255]:
                out.add(int(O.mko_unrevhash64(int(a))))
        return out

    for k in (3, 4):
        for combo in itertools.product(alphabet, repeat=k + 1):
            for c in canon_of(bytes(combo), k):
                assert c < bound(k), (k, bytes(combo))
    for k in (5, 8, 13, 21, 31):
        seqs = [bytes(rng.choice(np.frombuffer(alphabet, np.uint8), 3 * k + 7)) for _ in range(300)]
        seqs += [b"T" * a + b"A" * (2 * k - a) for a in range(0, 2 * k + 1)]
        seqs += [b"t" * a + b"T" * (k - a) + b"A" * k for a in range(0, k)]
        for s in seqs:
            for c in canon_of(s, k):
                assert c < bound(k), (k, s)
                assert ((c + 1023) >> 36) < L.mk_bloom_reach(k, 33)


def test_header_is_plain_c():
    """The boundary is a C ABI: include/miekki_b200.h must compile as C99 on its own (what a cgo /
    ctypes / JNI binding sees), with no C++ or CUDA types in it."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "miekki_b200.h")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
