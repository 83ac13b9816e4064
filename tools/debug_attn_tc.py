"""Bring-up helper for the tcgen05 attention kernels: run shapes one by one with progress prints."""
import faulthandler, os, sys, time
faulthandler.enable()
faulthandler.dump_traceback_later(100, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_checks as G
from midi_b200 import ops

def P(*a):
    print(*a, flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
shapes = [(1, 128, 4), (1, 256, 4), (2, 384, 16), (1, 200, 16), (2, 2047, 16), (8, 2048, 16)]
for (B, S, nh) in shapes:
    D, H = 64, nh * 64
    qkv = G.randn(B * S, 3 * H, seed=S + B)
    o2, lse2 = ops.attn_causal_fwd(qkv, B, S, nh, D, want_lse=True, impl="mma")
    torch.cuda.synchronize()
    if which == "fwd":
        P("fwd launch", B, S, nh)
        t = time.time()
        o, lse = ops.attn_causal_fwd(qkv, B, S, nh, D, want_lse=True, impl="tc")
        torch.cuda.synchronize()
        P("  done in %.3fs" % (time.time() - t), "rel o", G.rel(o.float(), o2.float()), "rel lse", G.rel(lse, lse2))
    else:
        do = G.randn(B * S, H, seed=S + B + 8)
        g_mm = ops.attn_causal_bwd(qkv, o2, do, lse2, B, S, nh, D, impl="mma")
        torch.cuda.synchronize()
        P("bwd launch", B, S, nh)
        t = time.time()
        g_tc = ops.attn_causal_bwd(qkv, o2, do, lse2, B, S, nh, D, impl="tc")
        torch.cuda.synchronize()
        P("  done in %.3fs" % (time.time() - t), "dq", G.rel(g_tc[:, :H].float(), g_mm[:, :H].float()),
          "dk", G.rel(g_tc[:, H:2 * H].float(), g_mm[:, H:2 * H].float()), "dv", G.rel(g_tc[:, 2 * H:].float(), g_mm[:, 2 * H:].float()))
# timing at the benchmark shape
B, S, nh, D, H = 8, 2048, 16, 64, 1024
qkv = G.randn(B * S, 3 * H, seed=1)
do = G.randn(B * S, H, seed=2)
o, lse = ops.attn_causal_fwd(qkv, B, S, nh, D, want_lse=True, impl="mma")
for impl in ("tc", "mma"):
    f = (lambda: ops.attn_causal_fwd(qkv, B, S, nh, D, want_lse=True, impl=impl)) if which == "fwd" else \
        (lambda: ops.attn_causal_bwd(qkv, o, do, lse, B, S, nh, D, impl=impl))
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    P(which, impl, "ms per call: %.3f" % (e0.elapsed_time(e1) / 10))
P("ALL DONE")
