"""2-rank check of the drop-in (autograd) path under torch DistributedDataParallel, the way Lightning runs the
reference's train.py (train.py:461-474): identical seeded weights, different data shards, loss through
forward -> forward_token -> F.cross_entropy (train.py:169-185), torch AdamW.  After two steps every parameter must be
bit-identical across ranks and the losses finite.  Launch: torchrun --nproc-per-node 2 tools/ddp_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "midi-model_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist, torch.nn.functional as F
from torch.nn.parallel import DistributedDataParallel as DDP
import midi_model as mm
from midi_b200.synth import synth_batch

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
cfg = mm.MIDIModelConfig.get_config("v2", True, n_layer=4, n_head=16, n_embd=1024, n_inner=4096)
model = mm.MIDIModel(cfg).to(f"cuda:{local}", dtype=torch.bfloat16).train()


class Wrap(torch.nn.Module):          # DDP calls .forward(); do the whole train.py training_step math in it
    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, batch):
        x, y = batch[:, :-1].contiguous(), batch[:, 1:].contiguous()
        hidden = self.m.forward(x)
        hidden = hidden.reshape(-1, hidden.shape[-1])
        y = y.reshape(-1, y.shape[-1])
        logits = self.m.forward_token(hidden, y[:, :-1])
        return F.cross_entropy(logits.view(-1, self.m.tokenizer.vocab_size), y.view(-1), reduction="mean",
                               ignore_index=self.m.tokenizer.pad_id)


ddp = DDP(Wrap(model), device_ids=[local])
opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4, betas=(0.9, 0.99))
losses = []
for step in range(2):
    batch = synth_batch(model.tokenizer, 2, 65, seed=100 * rank + step).to(f"cuda:{local}")
    loss = ddp(batch)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(ddp.parameters(), 1.0)
    opt.step()
    losses.append(float(loss))
chk = torch.stack([p.detach().float().sum() for p in model.parameters()] + [p.detach().float().abs().sum() for p in model.parameters()])
gathered = [torch.zeros_like(chk) for _ in range(dist.get_world_size())]
dist.all_gather(gathered, chk)
same = all(torch.equal(gathered[0], g) for g in gathered)
lt = torch.tensor(losses, device=f"cuda:{local}")
all_l = [torch.zeros_like(lt) for _ in range(dist.get_world_size())]
dist.all_gather(all_l, lt)
if rank == 0:
    print("DDP_CHECK params_identical_across_ranks=%s losses=%s" % (same, [[round(float(v), 4) for v in l] for l in all_l]))
    assert same and all(torch.isfinite(l).all() for l in all_l) and float(all_l[0][1]) != float(all_l[1][1])
dist.destroy_process_group()
