"""miekki_b200 -- B200-native (sm_100a) sketch-and-query path of Miekki.

Layout
  csrc/      hand-written CUDA kernels + the C ABI (include/miekki_b200.h) -> libmiekki_b200.so
  cli/       the `miekki` command line (C++ host: FASTA/gz parsing, dump I/O, text output)
  binding.py ctypes binding used by tests/ and bench.py
  synth.py   synthetic genomes / reads
"""
from .binding import HIT_DTYPE, LIB_PATH, Miekki, MiekkiError, lib  # noqa: F401

__all__ = ["Miekki", "MiekkiError", "HIT_DTYPE", "LIB_PATH", "lib"]
