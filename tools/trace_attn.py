import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_checks as G
from midi_b200 import ops, lib
B, S, nh, D, H = 8, 2048, 16, 64, 1024
qkv = G.randn(B * S, 3 * H, seed=1); do = G.randn(B * S, H, seed=2)
o, lse = ops.attn_causal_fwd(qkv, B, S, nh, D, want_lse=True, impl="mma")
for _ in range(3): ops.attn_causal_bwd(qkv, o, do, lse, B, S, nh, D, impl="tc")
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
lib.load().b200_attn_debug_trace(buf.data_ptr())
ops.attn_causal_bwd(qkv, o, do, lse, B, S, nh, D, impl="tc")
torch.cuda.synchronize()
lib.load().b200_attn_debug_trace(None)
t = buf.view(16, 8).cpu()
t0 = int(t[0, 0])
names = ["S issued", "MMA: ds_full seen", "MMA: next kv ready", "CW: s_full seen", "CW: tmem loaded", "CW: math+sts done", "CW: arrived"]
for j in range(16):
    print(j, {names[k]: int(t[j, k]) - t0 for k in range(7) if int(t[j, k])})
