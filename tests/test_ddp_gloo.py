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


def _trainer_worker(rank, world, port, q, lora_mode):
    """The fused trainer's data-parallel path end to end on CPU: training_loss(grad_ready=GradSync.ready) ->
    GradSync.wait -> fused_optimizer_step, over the CPU stand-in for the kernel layer (tests/mock_kernels.py)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (os.path.join(ROOT, "midi-model_b200"), ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import pytest
    import mock_kernels
    mock_kernels.install(pytest.MonkeyPatch())
    import midi_model as mm
    from midi_b200 import ddp, lora
    from midi_b200.synth import synth_batch
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                     # identical weights on every rank (as bench.py / train.py)
    model = mm.MIDIModel(mm.MIDIModelConfig.get_config("v2", True, n_layer=4, n_head=4, n_embd=256, n_inner=512))
    model = model.to(torch.bfloat16).train()
    if lora_mode:
        model.requires_grad_(False)
        model.add_adapter(lora.LoraAdapterConfig(r=8, lora_alpha=16, target_modules=["q_proj", "v_proj", "down_proj"]))
        g = torch.Generator().manual_seed(3)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if ".lora_B." in n:
                    p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(torch.bfloat16))
    batch = synth_batch(model.tokenizer, 2, 6, seed=1234 + rank)          # per-rank shard
    rt = model._rt()
    st = rt.store
    model.training_loss(batch)                                            # local gradients, no averaging
    local = st.gflat.clone()
    calls = []
    sync = ddp.GradSync(st.gflat)

    def ready(a, b):
        calls.append((a, b))
        sync.ready(a, b)

    model.training_loss(batch, grad_ready=ready)
    sync.wait()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    mean = (sum(t.float() for t in gathered) / world)
    lo, hi = st.train_lo, st.train_hi
    # every trainable element was handed over exactly once; nothing outside the trainable span was touched
    cover = torch.zeros(st.numel, dtype=torch.int32)
    for a, b in calls:
        cover[a:b] += 1
    ok_cover = bool((cover[lo:hi] == 1).all()) and (not lora_mode or bool((cover[:lo] == 0).all()))
    ok_avg = torch.allclose(st.gflat[lo:hi].float(), mean[lo:hi], rtol=2e-2, atol=1e-6)
    before = st.flat.clone()
    model.fused_optimizer_step(lr=1e-2, step=1)
    flats = [torch.zeros_like(st.flat) for _ in range(world)]
    dist.all_gather(flats, st.flat)
    ok_same = torch.equal(flats[0], flats[1])                             # ranks stay bit-identical after the update
    # the update touched the trainable span and nothing else (the frozen base of a LoRA run stays bit-identical)
    moved = bool((st.flat[lo:hi] != before[lo:hi]).any()) and torch.equal(st.flat[:lo], before[:lo]) \
        and torch.equal(st.flat[hi:], before[hi:])
    if rank == 0:
        q.put((ok_cover, ok_avg, ok_same, len(calls), bool(moved)))
    dist.destroy_process_group()


def _run_trainer(lora_mode, port_off):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + port_off) % 500
    procs = [ctx.Process(target=_trainer_worker, args=(r, 2, port, q, lora_mode)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return res


def test_fused_trainer_world2_full_training():
    ok_cover, ok_avg, ok_same, n_calls, moved = _run_trainer(False, 7)
    assert (ok_cover, ok_avg, ok_same, moved) == (True, True, True, True)
    assert n_calls >= 3          # token-level stack + lm_head first, then event-level layers, then the embedding table


def test_fused_trainer_world2_lora():
    ok_cover, ok_avg, ok_same, n_calls, moved = _run_trainer(True, 13)
    assert (ok_cover, ok_avg, ok_same, moved) == (True, True, True, True)
    assert n_calls == 1          # LoRA: one hand-over, the adapter tail


def _dropin_worker(rank, world, port, q, lora_mode):
    """What unchanged train.py gets under Lightning's DDP strategy (train.py:168-188, 461-474): the module wrapped in torch
    DistributedDataParallel, forward -> forward_token -> F.cross_entropy -> backward.  The engine's autograd Functions take
    the parameters as inputs, so DDP's gradient hooks fire and average the gradients; here over gloo on the mock kernel layer."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (os.path.join(ROOT, "midi-model_b200"), ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import pytest
    import torch.nn.functional as F
    from torch.nn.parallel import DistributedDataParallel as DDP
    import mock_kernels
    mock_kernels.install(pytest.MonkeyPatch())
    import midi_model as mm
    from midi_b200 import lora
    from midi_b200.synth import synth_batch
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = mm.MIDIModel(mm.MIDIModelConfig.get_config("v2", True, n_layer=4, n_head=4, n_embd=256, n_inner=512))
    model = model.to(torch.bfloat16).train()
    if lora_mode:
        model.requires_grad_(False)
        model.add_adapter(lora.LoraAdapterConfig(r=8, lora_alpha=16, target_modules=["q_proj", "k_proj", "up_proj"]))
        g = torch.Generator().manual_seed(3)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if ".lora_B." in n:
                    p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(torch.bfloat16))
    tok = model.tokenizer

    class Step(torch.nn.Module):                                  # TrainMIDIModel.training_step's arithmetic
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, batch):
            x, y = batch[:, :-1].contiguous(), batch[:, 1:].contiguous()
            hidden = self.m.forward(x)
            hidden = hidden.reshape(-1, hidden.shape[-1])
            y = y.reshape(-1, y.shape[-1])
            logits = self.m.forward_token(hidden, y[:, :-1])
            return F.cross_entropy(logits.view(-1, tok.vocab_size), y.view(-1), reduction="mean", ignore_index=tok.pad_id)

    batch = synth_batch(tok, 2, 6, seed=1234 + rank)
    Step(model)(batch).backward()                                 # local gradients
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    local = torch.cat([dict(model.named_parameters())[n].grad.float().reshape(-1) for n in names])
    for p in model.parameters():
        p.grad = None
    ddp = DDP(Step(model))
    ddp(batch).backward()
    synced = torch.cat([dict(model.named_parameters())[n].grad.float().reshape(-1) for n in names])
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    mean = sum(gathered) / world
    ok_avg = torch.allclose(synced, mean, rtol=2e-2, atol=1e-6)
    both = [torch.zeros_like(synced) for _ in range(world)]
    dist.all_gather(both, synced)
    ok_same = torch.equal(both[0], both[1])
    ok_frozen = all(p.grad is None for n, p in model.named_parameters() if not p.requires_grad)
    if rank == 0:
        q.put((ok_avg, ok_same, ok_frozen, len(names)))
    dist.destroy_process_group()


def _run_dropin(lora_mode, port_off):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + port_off) % 500
    procs = [ctx.Process(target=_dropin_worker, args=(r, 2, port, q, lora_mode)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return res


def test_dropin_under_torch_ddp_world2():
    ok_avg, ok_same, ok_frozen, n = _run_dropin(False, 23)
    assert (ok_avg, ok_same, ok_frozen) == (True, True, True) and n == 50          # 5 layers x 9 + 2 final norms + 2 embeddings + lm_head


def test_dropin_under_torch_ddp_world2_lora():
    ok_avg, ok_same, ok_frozen, n = _run_dropin(True, 31)
    assert (ok_avg, ok_same, ok_frozen) == (True, True, True) and n == 2 * 3 * 5    # A and B of three projections, five layers
