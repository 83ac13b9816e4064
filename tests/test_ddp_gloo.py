"""World-size-2 gloo tests (CPU) of the data-parallel host logic: per-rank synthetic shards differ, the
flat-gradient bucket average equals the mean over ranks, and the reference arm of bench.py prints on
rank 0 only."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(ROOT, "midi-model_b200"))
    from midi_b200 import ddp
    from midi_b200.synth import synth_batch
    from midi_b200.tokenizer_tables import TokenizerTables
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tok = TokenizerTables("v2")
    shard = synth_batch(tok, 2, 17, seed=1234 + rank)
    n = 3 * 1000 + 512
    g = torch.Generator().manual_seed(rank)
    flat = torch.randn(n, generator=g).to(torch.bfloat16).float()
    mine = flat.clone()
    sync = ddp.GradSync(flat, bucket_elems=700)
    sync.ready(2000, n)       # "inner stack + lm_head" slice first
    sync.ready(0, 2000)       # then the "outer stack" slice
    sync.wait()
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    shards = [torch.zeros_like(shard) for _ in range(world)]
    dist.all_gather(shards, shard)
    ok_avg = torch.allclose(flat, sum(gathered) / world, atol=1e-6)
    ok_diff = not torch.equal(shards[0], shards[1])
    same_after = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(same_after, flat)
    ok_same = torch.equal(same_after[0], same_after[1])
    if rank == 0:
        q.put((ok_avg, ok_diff, ok_same))
    dist.destroy_process_group()


def test_flat_gradient_average_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res == (True, True, True), res


def test_bucket_ranges_cover_exactly():
    sys.path.insert(0, os.path.join(ROOT, "midi-model_b200"))
    from midi_b200 import ddp
    r = ddp.bucket_ranges(1000, [300, 650], 256)
    assert r[0][0] == 0 and r[-1][1] == 1000
    assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
    assert all(e - s <= 256 for s, e in r)
    assert any(e == 300 for _, e in r) and any(e == 650 for _, e in r)
    assert ddp.bucket_ranges(0, [], 16) == []
