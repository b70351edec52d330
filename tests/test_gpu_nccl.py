"""The NCCL multi-GPU path (one process per GPU, torch.distributed) against the oracle: runs when
the box has at least two GPUs (`gpurun --gpus 2 ...`), skipped otherwise.  The host-side protocol
is also covered on CPU with gloo (tests/test_sharded_cpu.py)."""
import os
import socket
import subprocess
import sys

import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_nccl_sharded_hit_lists_equal_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs (run under `gpurun --gpus 2`)")
    for world in sorted({2, min(n, 4), min(n, 8)}):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                            "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                            "--master-port", str(free_port()), os.path.join(H.ROOT, "tests", "nccl_worker.py")],
                           capture_output=True, text=True, timeout=900, cwd=H.ROOT)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        assert r.stdout.count(" ok") == world


def test_one_process_cli_on_distinct_devices(tmp_path):
    """`miekki --gpus N` is one process with one context per GPU: the ring scan (> 1,024 genomes per
    shard, > 48 KB of dynamic shared memory) and the long-read sketch need their shared-memory
    opt-in on EVERY device (it is a per-device function attribute).  Output must equal the
    one-GPU run byte for byte."""
    import numpy as np
    import torch
    from miekki_b200 import synth
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs (run under `gpurun --gpus 2`)")
    cli = os.path.join(H.ROOT, "miekki_b200", "cli", "miekki")
    G, GL = 2200 * min(n, 2), 20_000              # > 1,024 genomes per shard at 2 GPUs
    names = []
    for g in range(G):
        p = tmp_path / ("g%d.fa" % g)
        synth.write_fasta(str(p), ">g%d" % g, synth.cb_bases(7, g, 0, GL).tobytes())
        names.append(str(p))
    (tmp_path / "list.txt").write_text("\n".join(names) + "\n")
    rng = np.random.default_rng(3)
    with open(tmp_path / "reads.fa", "wb") as f:
        for r in range(200):
            g = int(rng.integers(G))
            ln = 12_000 if r % 4 == 0 else 1500      # long reads: the 16,384-slot shared-memory table
            p = int(rng.integers(GL - ln))
            f.write(b">r%d\n" % r + synth.substitute(synth.cb_bases(7, g, p, ln), 0.02, rng).tobytes() + b"\n")
    outs = []
    for gpus in (1, min(n, 2), n):
        out = tmp_path / ("hits_%d.txt" % gpus)
        r = subprocess.run([cli, "-l", "list.txt", "-a", "reads.fa", "-k", "31", "-h", "12", "-s", "0", "-t", "8",
                            "-o", str(out), "--gpus", str(gpus)], cwd=tmp_path, capture_output=True, text=True,
                           timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs.append(out.read_text())
    assert outs[0].count("\n") == 200
    assert outs[1] == outs[0] and outs[2] == outs[0]
