"""bench.py -- train-step throughput of the B200-native MIDIModel (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one full optimizer step of tv2o-medium in bf16 on a synthetic grammar-valid batch of
(8 per GPU, 2049 events, 8 tokens): embed -> 12-layer event stack -> 3-layer token stack ->
lm_head -> CE (train.py:168-185) -> backward of all of it -> [gradient all-reduce over NCCL when
N > 1] -> global-norm clip + AdamW (train.py:121-138, 464).  Metric: MIDI-event tokens/s
(= global_batch * 2048 * 8 / step time; all 8 token slots counted, pads included).

  value : device-timed (CUDA events, max over ranks), batches already resident in HBM
  e2e   : same step through the public API with the batch in pinned HOST memory: H2D copy of the
          batch and D2H read of the loss inside the timed region, every step
  roofline     : aggregate of all tcgen05 GEMM launches of the timed steps (CUDA events around each
                 launch on the launching stream): algorithmic FLOPs / measured time vs measured bf16 peak
  cpu_baseline : the oracle port of the reference's path (fp32 eager PyTorch, all host cores) on a
                 bounded sample (B=1, S=128 train step), rank 0 at N=1 only

--impl reference times that CPU oracle train step alone (the reference is pure Python and cannot
travel to the GPU box; oracle/midi_oracle.py is its pinned restatement).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "midi-model_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the single JSON line

import torch  # noqa: E402

METRIC = "train_tokens_per_sec"
UNIT = "MIDI-event tokens/s"
MODEL = "tv2o-medium"
B_PER_GPU, S_EVENTS, T_TOK = 8, 2048, 8
LR, WARMUP_STEPS = 2e-4, 1000.0


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def train_flops_per_token() -> float:
    """forward 861.6 MFLOP/event at S=2048 (GEMMs + causal attention), x3 for training."""
    H, V, S = 1024, 3406, S_EVENTS
    outer = 12 * (4 * 2 * H * H + 3 * 2 * H * 4096 + 2 * 2 * H * (S + 1) / 2)
    inner = 3 * (4 * 2 * H * H + 3 * 2 * H * 1024 + 2 * 2 * H * (T_TOK + 1) / 2)
    per_event = outer + T_TOK * (inner + 2 * H * V)
    return 3.0 * per_event / T_TOK


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.1 <= t <= t1 + 0.3 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        reasons = []
        for i, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle's train step (fp32 eager PyTorch + torch AdamW), bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_train_step_tokens_per_s(steps: int, warmup: int, batch: int = 1, n_events: int = 128):
    import midi_model as mm
    from midi_b200.synth import synth_batch
    from oracle import midi_oracle as O
    # eager PyTorch at B=1 does not scale past a few dozen threads (128 threads on the GPU box were 7x SLOWER than
    # 16 in round-1 measurements): use up to 16 and report that count
    cores = min(os.cpu_count() or 1, 16)
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    cfg = mm.MIDIModelConfig.from_name(MODEL)
    model = mm.MIDIModel(cfg)          # parameter container only; the arithmetic below is the oracle's
    ocfg = O.cfg_from_hf(cfg)
    params = dict(model.named_parameters())
    no_decay = [p for n, p in params.items() if "bias" in n or "norm" in n]
    decay = [p for n, p in params.items() if not ("bias" in n or "norm" in n)]
    opt = torch.optim.AdamW([dict(params=decay, weight_decay=0.01), dict(params=no_decay, weight_decay=0.0)], lr=LR,
                            betas=(0.9, 0.99), eps=1e-8)
    data = [synth_batch(model.tokenizer, batch, n_events + 1, seed=1234 + i) for i in range(2)]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = O.train_loss(params, ocfg, data[i % 2])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)
        opt.step()
        float(loss.detach())
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    tok = batch * n_events * T_TOK
    ms = 1e3 * sum(times) / len(times)
    return tok / (ms / 1e3), ms, cores, f"oracle fp32 train step (fwd+bwd+clip+AdamW), B={batch}, S={n_events}, {steps} steps after {warmup} warm-up"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = max(1, min(args.steps, 6)), max(1, min(args.warmup, 2))
    v, ms, cores, sample = cpu_train_step_tokens_per_s(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{MODEL} train step, bounded CPU sample of batch=8 x 2048 events", "sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_native(args):
    import torch.distributed as dist
    import midi_model as mm
    from midi_b200 import lib, ops
    from midi_b200.synth import synth_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()

    torch.manual_seed(0)                                   # identical seeded-init weights on every rank
    model = mm.MIDIModel(mm.MIDIModelConfig.from_name(MODEL)).to(dev, dtype=torch.bfloat16).train()
    rt = model._rt()
    tok = model.tokenizer
    n_batches = 4
    host = [synth_batch(tok, B_PER_GPU, S_EVENTS + 1, seed=1234 + rank + 101 * i).pin_memory() for i in range(n_batches)]
    resident = [b.to(dev) for b in host]
    from midi_b200 import ddp
    # gradient averaging: NCCL all-reduce of the flat bf16 gradient buffer in 128 MB buckets on a side stream; the
    # token-level stack + lm_head slice is reduced while the event-level stack's backward is still running
    sync = ddp.GradSync(rt.store.gflat) if world > 1 else None

    state = {"step": 0}

    def step(batch_dev):
        state["step"] += 1
        loss = model.training_loss(batch_dev, grad_ready=sync.ready if sync else None)
        if sync:
            sync.wait()
        lr = LR * min(1.0, state["step"] / WARMUP_STEPS)
        model.fused_optimizer_step(lr=lr, step=state["step"])
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.query("b200_launch_count")
        t0 = time.time()
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, lib.query("b200_launch_count") - l0, t0, time.time()

    W, K = max(3, args.warmup), max(1, args.steps)
    for i in range(W):
        step(resident[i % n_batches])
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ms_total, launches, t0, t1 = timed(lambda i: step(resident[i % n_batches]), K)      # the timed region (`value`)
    # Roofline pass: the same K steps again with CUDA events around every GEMM launch.  It runs single-stream (the
    # weight-gradient GEMMs of the timed region run on a side stream under the HBM-bound kernels, where an event pair
    # would time the co-running kernels too), so `achieved` is each GEMM's own duration inside a full step.
    from midi_b200 import engine as _engine
    wg_stream, _engine.WGRAD_STREAM = _engine.WGRAD_STREAM, False
    ops.GEMM_PROFILE = []
    ms_prof, _, _, t1 = timed(lambda i: step(resident[i % n_batches]), K)
    prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
    _engine.WGRAD_STREAM = wg_stream
    clocks = sampler.stop(t0, t1)
    torch.cuda.synchronize()
    gemm_ms = sum(p[0].elapsed_time(p[1]) for p in prof)
    gemm_flops = sum(p[2] for p in prof)
    shapes = {}
    for e0, e1, fl, key in prof:
        ent = shapes.setdefault(key, [0, 0.0, 0.0])
        ent[0] += 1
        ent[1] += e0.elapsed_time(e1)
        ent[2] += fl
    per_shape = sorted(([*k, c, round(ms / K, 4), round(fl / (ms / 1e3) / 1e12, 1)] for k, (c, ms, fl) in shapes.items()),
                       key=lambda r: -r[-2])
    if rank == 0 and os.environ.get("B200_BENCH_VERBOSE"):
        print("M N K a_mn b_mn block_n splits count ms/step TF/s", file=sys.stderr)
        for r in per_shape:
            print(*r, file=sys.stderr)

    losses = []
    # end to end through the public API: int16 token batches in pinned host memory (the dataset's own dtype,
    # train.py:71) -> midi_b200.data.Prefetcher (H2D of batch i+1 on a copy stream while step i runs, inside the timed
    # region) -> MIDIModel.training_loss (widening + x/y split on the device) -> D2H read of the loss every step
    from midi_b200 import data as _data
    host16 = [_data.collate(list(h.numpy()), pad_id=0) for h in host]
    feed = iter(_data.Prefetcher((host16[i % n_batches] for i in range(K)), dev))

    def e2e_step(i):
        losses.append(float(step(next(feed))))              # D2H read of the loss

    ms_e2e, _, _, _ = timed(e2e_step, K)

    tokens_per_step = world * B_PER_GPU * S_EVENTS * T_TOK
    value = tokens_per_step / (ms_total / K / 1e3)
    e2e_value = tokens_per_step / (ms_e2e / K / 1e3)
    pk, pk_kind = peaks()
    peak_tf = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1400.0)))
    ach_tf = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    step_tf = train_flops_per_token() * (B_PER_GPU * S_EVENTS * T_TOK) / (ms_total / K / 1e3) / 1e12

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"{MODEL} train step (fwd+bwd+allreduce+clip+AdamW) bf16, batch={B_PER_GPU}/GPU x {S_EVENTS} events x {T_TOK} tokens",
                   "global_batch": world * B_PER_GPU, "n_events": S_EVENTS, "tokens_per_event": T_TOK,
                   "parallelism": f"dp{world}", "weights": "seeded-init (torch.manual_seed(0))",
                   "l2": "per-step working set ~20 GB >> 126 MB L2 (no explicit flush needed); 4 distinct batches cycled"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": int(host16[0].numel() * host16[0].element_size()), "d2h_bytes_per_step": 4,
                "last_loss": losses[-1] if losses else None},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": ach_tf / peak_tf if peak_tf else None, "traffic": gemm_traffic(),
                     "traffic_unit": "DRAM bytes (read+write) per GEMM launch, ncu capture in profiles/r1_gemm_traffic.json",
                     "kernel": "gemm_tcgen05_kernel (every launch of K steps, CUDA events on the launching stream; single-stream pass of the same steps)", "peak_kind": f"{pk_kind} sustained cuBLAS bf16",
                     "gemm_ms_per_step": gemm_ms / K, "gemm_share_of_step": gemm_ms / ms_prof if ms_prof else None,
                     "profile_pass_ms_per_step": ms_prof / K, "wgrad_side_stream_in_timed_region": bool(wg_stream),
                     "whole_step_model_tflops": step_tf, "whole_step_frac": step_tf / peak_tf if peak_tf else None,
                     "per_shape": {"columns": "M,N,K,a_mn,b_mn,block_n,splits,launches,ms_per_step,TFLOP/s", "rows": per_shape}},
    }
    if not args.no_generate:
        line["generate"] = generate_leg(model, dev, world, rank)
    if world == 1 and not args.no_cpu_baseline:
        v, ms, cores, sample = cpu_train_step_tokens_per_s(3, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ms_per_step": ms}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def gemm_traffic():
    """DRAM bytes (read + write) per GEMM launch from the committed ncu capture of this same command
    (profiles/r1_gemm_traffic.json, derived from profiles/r1_launches_step.csv); None when the capture is absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def generate_leg(model, dev, world, rank):
    """BASELINE.json's second metric: events/s of the KV-cached generate loop (midi_model.py:167-250) with the
    reference's default sampling (temp 1.0, top-p 0.98, top-k 20), from one BOS event; replicas only across GPUs
    (each rank generates its own rows, no collective).  EOS stopping is disabled so the event count is fixed
    (seeded-init weights emit EOS at random)."""
    import torch.distributed as dist
    from midi_b200 import decode as dec
    rt = model._rt()
    tok = model.tokenizer
    res = {"unit": "events/s", "sampling": "temp=1.0 top_p=0.98 top_k=20", "note": "one CUDA-graph replay per event; wall clock incl. launches"}
    # batch 8 runs BASELINE config 3 in full: 4096-event context (1 BOS + 4095 generated events per row)
    for B, n_new in ((1, 512), (8, 4095)):
        gg = dec.GraphGenerator(model._cached_stack("outer"), model._cached_stack("inner"), rt.lm_head, rt.pitch, rt.V, tok,
                                dec.GrammarLUT(tok, dev), B, n_new + 1, 1.0, 0.98, 20, 1234 + rank)
        prompt = torch.full((B, 1, tok.max_token_seq), tok.pad_id, dtype=torch.long, device=dev)
        prompt[:, 0, 0] = tok.bos_id
        gg.run(prompt, check_every=1 << 30, stop_on_eos=False)      # captures the graph + warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        out = gg.run(prompt, check_every=1 << 30, stop_on_eos=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert out.shape[1] == n_new + 1
        res[f"batch{B}"] = {"events_per_s": world * B * n_new / dt, "ms_per_event_step": 1e3 * dt / n_new, "rows": world * B,
                            "new_events_per_row": n_new}
        del gg
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-generate", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
