"""Launch the event-level attention forward a few times at the bench shape (B=8, S=2048, 16 heads, d=64): ncu target."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "midi-model_b200"))
from midi_b200 import ops  # noqa: E402

B, S, nh, D = 8, 2048, 16, 64
g = torch.Generator(device="cuda").manual_seed(1)
qkv = (torch.randn(B * S, 3 * nh * D, device="cuda", generator=g)).to(torch.bfloat16)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(n):
    o, lse = ops.attn_causal_fwd(qkv, B, S, nh, D, want_lse=True, impl="tc")
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
