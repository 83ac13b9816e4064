"""Debug helper: compare host-driven (eager) and CUDA-graph generate loops with the oracle on a peaked model."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import gpu_checks as G
from oracle import midi_oracle as O

mm, model = G._model(4, seed=0)
ocfg = O.cfg_from_hf(model.config)
model = model.to("cuda", dtype=torch.bfloat16).train()
tok = model.tokenizer
for step in range(1, 241):
    batch = G._song_batch(tok, 16, 66, seed=step).to("cuda")
    loss = model.training_loss(batch)
    model.fused_optimizer_step(lr=3e-4 * min(1.0, step / 20), step=step, weight_decay=0.01)
print("final loss", float(loss))
model.eval()
sd16 = G._sd(model, torch.bfloat16)
prompt = G._song_batch(tok, 4, 9, seed=999).numpy()
res = {}
for mode in ("eager", "graph", "nograph", "eager"):
    os.environ["B200_GENERATE"] = mode
    res[mode] = model.generate(prompt=prompt, batch_size=4, max_len=24, top_k=1)
ref = O.generate(sd16, ocfg, tok, prompt, batch_size=4, max_len=24, top_k=1,
                 inv_freq_net=model.net.rotary_emb.inv_freq, inv_freq_tok=model.net_token.rotary_emb.inv_freq)
for mode, ids in res.items():
    neq = ids != ref
    print(mode, "mismatch vs oracle:", int(neq.sum()), "per token position:", neq.sum((0, 1)).tolist(), "per event:", neq.sum((0, 2)).tolist())
print("ref row0:\n", ref[0, 8:14])
print("eager row0:\n", res["eager"][0, 8:14])
