"""Run one bf16 GEMM shape a few times (for ncu captures): python tools/gemm_one.py M N K [a_mn b_mn]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "midi-model_b200"))
import torch
from midi_b200 import ops

M, N, K = (int(v) for v in sys.argv[1:4])
a_mn = len(sys.argv) > 4 and sys.argv[4] == "1"
b_mn = len(sys.argv) > 5 and sys.argv[5] == "1"
g = torch.Generator(device="cuda").manual_seed(0)
A = (torch.randn((K, M) if a_mn else (M, K), device="cuda", generator=g) * 0.05).to(torch.bfloat16)
B = (torch.randn((K, N) if b_mn else (N, K), device="cuda", generator=g) * 0.05).to(torch.bfloat16)
for _ in range(3):
    C = ops.gemm(A, B, M, N, K, lda=A.stride(0), ldb=B.stride(0), a_mn=a_mn, b_mn=b_mn, allow_split=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    C = ops.gemm(A, B, M, N, K, lda=A.stride(0), ldb=B.stride(0), a_mn=a_mn, b_mn=b_mn, allow_split=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("gemm %d %d %d: %.4f ms  %.1f TFLOP/s" % (M, N, K, ms, 2.0 * M * N * K / ms / 1e9))
