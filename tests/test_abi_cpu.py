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
