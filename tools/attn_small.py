"""Small-shape forward + backward of the tcgen05 attention kernels (compute-sanitizer target): shapes with one, two and three
query tiles, a ragged last tile and an idle second softmax group; checks against fp32 SDPA so a silent corruption also shows."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "midi-model_b200"))
from midi_b200 import ops  # noqa: E402

worst = 0.0
for (B, S, nh) in ((1, 128, 4), (1, 200, 4), (2, 384, 4)):
    D, H = 64, nh * 64
    g = torch.Generator(device="cuda").manual_seed(S)
    qkv = torch.randn(B * S, 3 * H, device="cuda", generator=g).to(torch.bfloat16)
    do = torch.randn(B * S, H, device="cuda", generator=g).to(torch.bfloat16)
    o, lse = ops.attn_causal_fwd(qkv, B, S, nh, D, want_lse=True, impl="tc")
    dqkv = ops.attn_causal_bwd(qkv, o, do, lse, B, S, nh, D, impl="tc")
    torch.cuda.synchronize()
    q32 = qkv.float().view(B, S, 3, nh, D).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)
    s = q32[0] @ q32[1].transpose(-1, -2) / math.sqrt(D)
    m = torch.triu(torch.ones(S, S, device="cuda", dtype=torch.bool), 1)
    ref = torch.softmax(s.masked_fill(m, float("-inf")), -1) @ q32[2]
    ref.backward(do.float().view(B, S, nh, D).transpose(1, 2))
    gref = q32.grad.permute(1, 3, 0, 2, 4).reshape(B * S, 3 * H)
    e_f = float((o.float().view(B, S, nh, D).transpose(1, 2) - ref).norm() / ref.norm())
    e_b = float((dqkv.float() - gref).norm() / gref.norm())
    worst = max(worst, e_f, e_b)
    print(f"B={B} S={S} heads={nh}: fwd rel {e_f:.2e}  bwd rel {e_b:.2e}")
print("worst", worst)
sys.exit(0 if worst < 1e-2 else 1)
