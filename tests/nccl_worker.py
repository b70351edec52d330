"""One rank of the multi-GPU parity run (launched by tests/test_gpu_nccl.py under torchrun, one
process per GPU, NCCL): the production sharded path -- `sharded.merge_bloom` (all-gather + fold)
and `sharded.pipelined_query` (scan of batch i+1 in flight while batch i's heap is chained through
the ranks with NCCL send/recv) -- on real sketches, checked on the last rank against the oracle's
single in-order index: every Bloom byte, and every hit list as text at -s 200 and at the
all-ties -s 0 (Miekki.cpp:376-397: the heap is order dependent, quirk G5)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import miekki_b200  # noqa: E402
from miekki_b200 import sharded, synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    sharded.nccl_env_defaults()
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    k, h, G, GL = 31, 12, 75, 60_000
    rng = np.random.default_rng(42)
    genomes = [synth.genome(g, GL) for g in range(G - 15)]
    for j in range(15):      # relatives of earlier genomes: heap replacements and near ties across shards
        genomes.append(synth.substitute(np.frombuffer(genomes[3 * j], np.uint8), 0.01 + 0.002 * j, rng).tobytes())
    reads = [s for _, s in synth.sample_reads(genomes, 330, 1500, sub_rate=0.01, block=9)]
    first, count = sharded.shard_range(G, rank, world)
    ix = miekki_b200.Miekki(k=k, h=h, threshold=200, device=local)
    ix.set_shard(first)
    ix.set_scan_spare_sms(1)       # as bench.py does: room for the NCCL kernels of the chain
    ix.insert_sequences(genomes[first:first + count])
    w = ix.bloom_window()
    mine = torch.empty(w, dtype=torch.uint8, device="cuda")
    ix.bloom_get_ptr(mine.data_ptr(), w)
    merged = sharded.merge_bloom(mine).contiguous()
    torch.cuda.synchronize()       # the fold ran on torch's stream; the context copies on its own
    ix.bloom_set_ptr(merged.data_ptr(), w)

    o = orc.Oracle(k=k, h=h, cap=G)
    for s in genomes:
        o.insert(s)
    m = min(w, len(o.bloom))
    assert np.array_equal(merged.cpu().numpy()[:m], o.bloom[:m]), "merged Bloom table differs from the in-order build"

    K, nb = 10, 3
    per = len(reads) // nb
    for thr in (200, 0):
        batches = [ix.upload(reads[b * per:(b + 1) * per]) for b in range(nb)]
        heap = torch.zeros((per, K * 24), dtype=torch.uint8, device="cuda")
        lens = torch.zeros(per, dtype=torch.int32, device="cuda")
        got = {}

        def on_result(i):
            torch.cuda.current_stream().synchronize()
            got[i] = (heap.cpu().numpy().view(miekki_b200.HIT_DTYPE).reshape(per, K).copy(),
                      lens.cpu().numpy().view(np.uint32).copy())
        # the last batch is chained in 3 tiles of reads (mk_topk_slot_range), the others whole
        sharded.pipelined_query(ix, iter(batches), heap, lens, K, 10, 0.5 * thr, on_result=on_result, tiles=2, last_tiles=3)
        if rank == world - 1:
            assert sorted(got) == list(range(nb))
            for b in range(nb):
                hh, ll = got[b]
                for j in range(per):
                    s = reads[b * per + j]
                    oc, _ = o.counts(s)
                    want = o.filter(oc, K, 10, 0.5 * thr)
                    a = orc.format_hit_line(">r", hh[j, :ll[j]])
                    e = orc.format_hit_line(">r", want)
                    assert a == e, "rank %d batch %d read %d at -s %d:\n%s%s" % (rank, b, j, thr, a, e)
                    assert np.array_equal(hh[j, :ll[j]]["genome"], want["genome"])
        for bt in batches:
            bt.free()
        dist.barrier()
    ix.close()
    dist.destroy_process_group()
    print("rank %d of %d ok" % (rank, world))


if __name__ == "__main__":
    main()
