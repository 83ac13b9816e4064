"""Per-phase time of the persistent generate kernel (csrc/decode_persist.cu): SM cycles CTA 0 spends in each phase
(including the grid barrier that closes it), from the kernel's own clock64 hook (b200_decode_desc.prof).

    python tools/decode_profile.py [batch] [n_events]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "midi-model_b200"))
sys.path.insert(0, ROOT)
os.environ["B200_DECODE_PROFILE"] = "1"
import torch  # noqa: E402

import midi_model as mm  # noqa: E402
from midi_b200 import decode as dec  # noqa: E402

NAMES = ["qkv(event)", "attention(event)", "combine(event)", "o_proj(event)", "gate|up(event)", "down(event)", "qkv(token)",
         "attention(token)", "o_proj(token)", "gate|up(token)", "down(token)", "lm_head", "sample", "commit"]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n_new = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    dev = torch.device("cuda")
    torch.manual_seed(0)
    model = mm.MIDIModel(mm.MIDIModelConfig.from_name("tv2o-medium")).to(dev, dtype=torch.bfloat16).eval()
    rt = model._rt()
    tok = model.tokenizer
    gg = dec.GraphGenerator(model._cached_stack("outer"), model._cached_stack("inner"), rt.lm_head, rt.pitch, rt.V, tok,
                            dec.GrammarLUT(tok, dev), B, n_new + 1, 1.0, 0.98, 20, 1234)
    prompt = torch.full((B, 1, 8), tok.pad_id, dtype=torch.long, device=dev)
    prompt[:, 0, 0] = tok.bos_id
    gg.run(prompt, use_graph="persist", check_every=64, stop_on_eos=False)
    torch.cuda.synchronize()
    gg.prof.zero_()
    t0 = time.perf_counter()
    gg.run(prompt, use_graph="persist", check_every=64, stop_on_eos=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    prof = gg.prof.cpu().tolist()
    tot = sum(prof[:14])
    mhz = 1900.0        # nominal SM clock under this light load (sm_max 1965 MHz); cycles are exact, microseconds approximate
    print(f"batch {B}, {n_new} events: {1e3 * dt / n_new:.3f} ms/event, {B * n_new / dt:.0f} events/s; SM clock {mhz:.0f} MHz (assumed for us)")
    print(f"{'phase':18s} {'calls/event':>11s} {'cycles/call':>11s} {'stage':>7s} {'work':>7s} {'barrier':>8s} {'us/call':>8s} {'us/event':>9s} {'share':>6s}")
    for i, n in enumerate(NAMES):
        c, k = prof[i], prof[32 + i]
        if k == 0:
            continue
        st, wk = prof[64 + 2 * i] / k, prof[64 + 2 * i + 1] / k
        print(f"{n:18s} {k / n_new:11.1f} {c / k:11.0f} {st:7.0f} {wk:7.0f} {c / k - st - wk:8.0f} {c / k / mhz:8.2f} {c / n_new / mhz:9.1f} {c / tot:6.1%}")
    print(f"{'total':18s} {sum(prof[32:46]) / n_new:11.1f} {'':11s} {'':8s} {tot / n_new / mhz:9.1f}")


if __name__ == "__main__":
    main()
