"""Synthetic genomes and reads.

Two generators, both deterministic:

* ``genome()`` / ``sample_reads()`` -- numpy ``default_rng`` streams, the
  recipe BASELINE.md names (seed 1000+g per genome, 2_000_000+block per read
  block).  Used for everything that fits the host: tests, config 1, the
  reference-binary baseline.
* ``cb_bases()`` -- a counter-based generator (splitmix64 of
  ``seed + g*2^40 + position``, top two bits -> ``ACGT``).  The same function is
  implemented on the device (``mk_synth_genomes`` in csrc/synth.cu) so that the
  full-size configs (10,000+ genomes of 5 Mbp = 50+ Gbp) can be materialised in
  HBM in seconds while the host can still cut reads out of any genome without
  ever holding it.
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", np.uint8)
_M = (1 << 64) - 1


def genome(g: int, length: int = 5_000_000) -> bytes:
    """i.i.d. uniform ACGT, numpy default_rng(1000 + g)."""
    rng = np.random.default_rng(1000 + g)
    return ACGT[rng.integers(0, 4, length, dtype=np.uint8)].tobytes()


def substitute(seq: np.ndarray, rate: float, rng) -> np.ndarray:
    """Each base independently replaced, with probability ``rate``, by one of the
    three other bases (uniformly)."""
    if rate <= 0:
        return seq
    codes = np.searchsorted(ACGT, seq) & 3  # ACGT is sorted in ASCII
    hit = (rng.random(len(seq)) < rate) & np.isin(seq, ACGT)  # other letters stay
    shift = rng.integers(1, 4, len(seq), dtype=np.uint8)
    return np.where(hit, ACGT[(codes + shift) & 3], seq)


def sample_reads(genomes, n_reads: int, read_len: int, sub_rate: float = 0.0, block: int = 0):
    """Reads cut from ``genomes`` (a list of bytes or a callable g -> bytes slice
    provider is not needed at this size).  Returns [(header, seq)] with the
    strict 2-line record naming ``>read<r>_g<g>_p<p>``."""
    rng = np.random.default_rng(2_000_000 + block)
    out = []
    for r in range(n_reads):
        g = int(rng.integers(len(genomes)))
        src = genomes[g]
        p = int(rng.integers(len(src) - read_len))
        seq = np.frombuffer(src, np.uint8, read_len, p)
        seq = substitute(seq, sub_rate, rng)
        out.append((">read%d_g%d_p%d" % (block * 10_000 + r, g, p), seq.tobytes()))
    return out


def write_fasta(path: str, header: str, seq: bytes) -> None:
    with open(path, "wb") as f:
        f.write(header.encode() + b"\n" + seq + b"\n")


def write_reads(path: str, reads) -> None:
    with open(path, "wb") as f:
        for head, seq in reads:
            f.write(head.encode() + b"\n" + seq + b"\n")


# ---- counter-based generator (mirrored on the device) ------------------------

def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def cb_bases(seed: int, g: int, start: int, length: int) -> np.ndarray:
    """ASCII bases [start, start+length) of counter-based genome ``g``."""
    with np.errstate(over="ignore"):
        ctr = np.arange(start, start + length, dtype=np.uint64)
        ctr = ctr + np.uint64(((seed & _M) + (g << 40)) & _M)
        return ACGT[(_splitmix64(ctr) >> np.uint64(62)).astype(np.uint8)]


def cb_reads(seed: int, n_genomes: int, genome_len: int, n_reads: int, read_len: int,
             sub_rate: float = 0.0, block: int = 0):
    """Reads cut from the counter-based genomes -> (uint8 [n_reads, read_len],
    source genome ids, offsets)."""
    rng = np.random.default_rng(2_000_000 + block)
    gs = rng.integers(n_genomes, size=n_reads)
    ps = rng.integers(genome_len - read_len, size=n_reads)
    out = np.empty((n_reads, read_len), np.uint8)
    for r in range(n_reads):
        out[r] = substitute(cb_bases(seed, int(gs[r]), int(ps[r]), read_len), sub_rate, rng)
    return out, gs, ps


def cb_reads_block(seed: int, n_genomes: int, genome_len: int, n_reads: int, read_len: int,
                   sub_rate: float = 0.0, block: int = 0, chunk: int = 1000):
    """Vectorised form of ``cb_reads`` for the long-read configurations (its own random stream:
    ``default_rng(3_000_000 + block)``, a chunk of reads per numpy call instead of one read).
    -> (uint8 [n_reads, read_len], source genome ids, offsets)."""
    rng = np.random.default_rng(3_000_000 + block)
    gs = rng.integers(n_genomes, size=n_reads)
    ps = rng.integers(genome_len - read_len, size=n_reads)
    out = np.empty((n_reads, read_len), np.uint8)
    span = np.arange(read_len, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for r0 in range(0, n_reads, chunk):
            m = min(chunk, n_reads - r0)
            base = (np.uint64(seed & _M) + (gs[r0:r0 + m].astype(np.uint64) << np.uint64(40))
                    + ps[r0:r0 + m].astype(np.uint64))
            codes = (_splitmix64(base[:, None] + span[None, :]) >> np.uint64(62)).astype(np.uint8)
            if sub_rate > 0:
                hit = rng.random((m, read_len), dtype=np.float32) < sub_rate
                shift = rng.integers(1, 4, (m, read_len), dtype=np.uint8)
                codes = np.where(hit, (codes + shift) & 3, codes)
            out[r0:r0 + m] = ACGT[codes]
    return out, gs, ps
