"""Run a short generate() so that ncu can list the decode kernels (use B200_GENERATE=nograph for per-kernel launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "midi-model_b200")); sys.path.insert(0, ROOT)
import torch
import midi_model as mm
torch.manual_seed(0)
model = mm.MIDIModel(mm.MIDIModelConfig.from_name("tv2o-medium")).to("cuda", dtype=torch.bfloat16).eval()
B = int(os.environ.get("GEN_B", "1"))
ids = model.generate(batch_size=B, max_len=int(os.environ.get("GEN_LEN", "12")), generator=torch.Generator("cuda").manual_seed(0))
print(ids.shape)
