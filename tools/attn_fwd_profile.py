"""Per-phase cycle profile of the softmax warps of the attention forward (third generation) at the bench shape.

The kernel's instrumented instantiation sums clock64 deltas of lane 0 of every softmax warp into a device buffer
(b200_attn_debug_trace); this prints cycles per tile and shares, per group."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "midi-model_b200"))
from midi_b200 import lib, ops  # noqa: E402

B, S, nh, D = 8, 2048, 16, 64
g = torch.Generator(device="cuda").manual_seed(1)
qkv = torch.randn(B * S, 3 * nh * D, device="cuda", generator=g).to(torch.bfloat16)
for _ in range(3):
    ops.attn_causal_fwd(qkv, B, S, nh, D, want_lse=True, impl="tc")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attn_causal_fwd(qkv, B, S, nh, D, want_lse=True, impl="tc")
e1.record()
torch.cuda.synchronize()
print(f"plain kernel: {e0.elapsed_time(e1) / 10:.4f} ms")
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
lib.load().b200_attn_debug_trace(buf.data_ptr())
ops.attn_causal_fwd(qkv, B, S, nh, D, want_lse=True, impl="tc")
torch.cuda.synchronize()
lib.load().b200_attn_debug_trace(None)
names = ["s_full wait", "pass 1 (row max)", "pv_done wait + rescale check", "token wait", "pass 2 (exp, P store)",
         "fence + p_full arrive", "item epilogue / bookkeeping"]
t = buf.cpu().view(8, 16)
for grp in range(2):
    r = t[grp]
    tiles, warps, total = int(r[8]), int(r[9]), int(r[7])
    print(f"group {grp}: {warps} warps, {tiles} warp-tiles, {total / warps:.0f} cycles per warp, {total / max(tiles, 1):.0f} cycles per warp-tile")
    for k, nme in enumerate(names):
        print(f"   {nme:32s} {int(r[k]) / max(tiles, 1):8.0f} cycles/tile  {100.0 * int(r[k]) / max(total, 1):5.1f} %")
