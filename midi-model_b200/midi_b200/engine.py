"""Layer-by-layer execution of the two Llama stacks of MIDIModel on the sm_100a kernels.

`ParamStore` keeps every parameter (and gradient) of the model as a view into ONE flat bf16
buffer, in `named_parameters()` order.  Because q/k/v and gate/up projections are adjacent in
that order, the fused `[3H, H]` QKV and `[2I, H]` gate|up weights are *views* -- the reference's
`state_dict` keys stay the source of truth (checkpoint row) and no packed
copies exist.  One flat gradient buffer means one all-reduce and one AdamW launch per step.

`StackEngine` runs forward (saving exactly what backward needs) and an explicit backward; there
is no autograd graph inside -- `autograd.Function`s in midi_model.py wrap whole stacks.
Math per layer: hf modeling_llama.py:303-332; rounding points: SURVEY.md Appendix A.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from . import lib, ops

import os

BF16 = torch.bfloat16
# RoPE backward is fused into the attention backward kernels (B200_FUSE_ROPE=0 -> stand-alone kernel).
# The forward fusion into the QKV GEMM epilogue exists (ops.linear_rope, bit-identical, per-thread or TMA-staged stores)
# but is OFF by default: four epilogue warps doing the rotation outlast the K=1024 main loop.  Measured A/B on one box
# with the staged epilogue: +0.2 ms/step against the HBM-roofline stand-alone kernel (QKV GEMM 1213 -> 843 TFLOP/s with
# the earlier per-thread stores).
FUSE_ROPE = os.environ.get("B200_FUSE_ROPE", "1") != "0"
FUSE_ROPE_FWD = os.environ.get("B200_FUSE_ROPE_FWD", "0") != "0"
# SwiGLU formed in the epilogue of the gate|up GEMM (ops.linear_swiglu, bit-identical to GEMM + stand-alone kernel): ON.
# Round 1 measured it slower (+1.5 ms/step: four epilogue warps doing an IEEE division and two SFU ops per element outlasted
# the K=1024 main loop).  With two epilogue warp groups (8 warps, each group its own staging box and store leader) and
# silu through ex2.approx / rcp.approx the fused GEMM costs +1.0 ms of GEMM time per step and removes the 1.5 ms/step
# stand-alone kernel plus its 17.6 GB of traffic: 56.4 -> 54.8 ms/step on one box (profiles/r2_epilogue_ab.txt).
# B200_FUSE_SWIGLU=0 restores GEMM + kernel.  (The fused RoPE epilogue stays off: neutral within box noise.)
FUSE_SWIGLU = os.environ.get("B200_FUSE_SWIGLU", "1") != "0"
# Weight-gradient GEMMs on a second stream: dW = dY^T X is off the backward's critical path, so it can run
# under the HBM-bound kernels that follow on the main stream (SwiGLU / RMSNorm backward leave the tensor pipe idle, and an
# elementwise CTA fits next to a GEMM CTA on an SM).  B200_WGRAD_STREAM=0 keeps everything on one stream.
WGRAD_STREAM = os.environ.get("B200_WGRAD_STREAM", "1") != "0"
_side_streams: dict = {}


def _side_stream(device) -> torch.cuda.Stream:
    s = _side_streams.get(device.index)
    if s is None:
        s = torch.cuda.Stream(device=device, priority=0)     # main work keeps the default (same) priority class
        _side_streams[device.index] = s
    return s


def _wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, accumulate: bool, side, hold=None) -> None:
    """dW (+)= dY^T X, on `side` when given: ordered after everything queued so far on the current stream; the caller
    joins `side` back before the gradients are consumed.  The operands must outlive the side GEMM: `hold` (a list) takes a
    reference the caller drops only after an event recorded on `side` has been waited for; with hold=None the caller
    guarantees the lifetime itself.  (Tensor.record_stream would also do, but it makes the caching allocator unable to
    reuse those blocks while the host runs ahead of the device: the 0.9 GB logits / dlogits block of every step then
    triggers fresh cudaMallocs -- measured +5 ms per step.)"""
    if side is None:
        ops.linear_wgrad(dy, x, dw, accumulate)
        return
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.linear_wgrad(dy, x, dw, accumulate)
    if hold is not None:
        hold.append((dy, x))


ALIGN = 256   # elements; AdamW's no-decay flags are per 256-element block


class ParamStore:
    def __init__(self, module: torch.nn.Module):
        named = list(module.named_parameters())
        if not named:
            raise lib.B200Error("model has no parameters")
        dev = named[0][1].device
        for n, p in named:
            if p.device != dev or p.dtype != BF16 or not p.is_cuda:
                raise lib.B200Error(f"parameter {n} is {p.dtype} on {p.device}: the B200 path needs the whole model in "
                                    "bfloat16 on one CUDA device (model.to('cuda', dtype=torch.bfloat16)); no fallback")
        self.device = dev
        self.names = [n for n, _ in named]
        self.offsets = {}
        off = 0
        for n, p in named:
            self.offsets[n] = off
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.flat = torch.zeros(off, dtype=BF16, device=dev)
        self.gflat = torch.zeros(off, dtype=BF16, device=dev)
        self.views = {}
        self.gviews = {}
        with torch.no_grad():
            for n, p in named:
                o = self.offsets[n]
                v = self.flat[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                self.views[n] = v
                self.gviews[n] = self.gflat[o:o + p.numel()].view(p.shape)
        self._params = dict(named)
        # no-decay flags (train.py:123-131: names containing 'bias' or 'norm')
        flags = torch.zeros(off // ALIGN, dtype=torch.uint8)
        for n, p in named:
            if "bias" in n or "norm" in n:
                o = self.offsets[n] // ALIGN
                flags[o:o + (p.numel() + ALIGN - 1) // ALIGN] = 1
        self.nodecay = flags.to(dev)

    def valid(self) -> bool:
        """False once somebody re-created or re-pointed ANY parameter (.to(dtype), .cuda(), p.data = ..., ...): the
        engine reads the flat buffer, so a parameter living elsewhere would silently be ignored."""
        views, params = self.views, self._params
        for n in self.names:
            if params[n].data_ptr() != views[n].data_ptr():
                return False
        return True

    def fused(self, names: List[str]) -> torch.Tensor:
        """[sum(rows), cols] view over adjacent 2-D parameters (q|k|v or gate|up)."""
        first = self.views[names[0]]
        o = self.offsets[names[0]]
        rows = 0
        for n in names:
            v = self.views[n]
            if self.offsets[n] != o + rows * first.shape[1] or v.shape[1] != first.shape[1]:
                raise lib.B200Error(f"parameters {names} are not adjacent in the flat buffer")
            rows += v.shape[0]
        return self.flat[o:o + rows * first.shape[1]].view(rows, first.shape[1])

    def fused_grad(self, names: List[str]) -> torch.Tensor:
        first = self.views[names[0]]
        o = self.offsets[names[0]]
        rows = sum(self.views[n].shape[0] for n in names)
        return self.gflat[o:o + rows * first.shape[1]].view(rows, first.shape[1])

    def publish_grads(self):
        """Expose the flat gradient buffer as `.grad` of every parameter (for train.py / torch optimizers)."""
        for n, p in self._params.items():
            if p.requires_grad:
                p.grad = self.gviews[n]

    def zero_grad(self):
        self.gflat.zero_()


@dataclass
class StackCfg:
    prefix: str
    n_layer: int
    n_head: int
    hidden: int
    inner: int
    eps: float

    @property
    def head_dim(self) -> int:
        return self.hidden // self.n_head


class LayerW:
    __slots__ = ("qkv", "o", "gu", "down", "ln1", "ln2")


class LayerG:
    __slots__ = ("qkv", "o", "gu", "down", "ln1", "ln2")


class StackGrads:
    """Gradient views of one stack inside a flat buffer laid out like ParamStore.flat (offset by `base`)."""

    def __init__(self, store: "ParamStore", cfg: "StackCfg", buf: torch.Tensor, base: int):
        p = cfg.prefix
        H = cfg.hidden

        def view(name, rows=None):
            v = store.views[name]
            o = store.offsets[name] - base
            if rows is None:
                return buf[o:o + v.numel()].view(v.shape)
            return buf[o:o + rows * v.shape[1]].view(rows, v.shape[1])

        self.layers = []
        for l in range(cfg.n_layer):
            a = f"{p}.layers.{l}.self_attn."
            m = f"{p}.layers.{l}.mlp."
            g = LayerG()
            g.qkv = view(a + "q_proj.weight", 3 * H)
            g.o = view(a + "o_proj.weight")
            g.gu = view(m + "gate_proj.weight", 2 * cfg.inner)
            g.down = view(m + "down_proj.weight")
            g.ln1 = view(f"{p}.layers.{l}.input_layernorm.weight")
            g.ln2 = view(f"{p}.layers.{l}.post_attention_layernorm.weight")
            self.layers.append(g)
        self.norm = view(f"{p}.norm.weight")
        self.embed = view(f"{p}.embed_tokens.weight")
        self.buf, self.base = buf, base

    def named(self, store: "ParamStore", names):
        """name -> gradient view, for handing gradients back to autograd."""
        out = []
        for n in names:
            v = store.views[n]
            o = store.offsets[n] - self.base
            out.append(self.buf[o:o + v.numel()].view(v.shape))
        return out


class StackEngine:
    """One Llama stack (outer `net`: causal over events; inner `net_token`: causal over <= 8 tokens per event)."""

    def __init__(self, store: ParamStore, cfg: StackCfg, tiny_attention: bool):
        self.cfg = cfg
        self.store = store
        self.tiny = tiny_attention
        p = cfg.prefix
        self.layers: List[LayerW] = []
        for l in range(cfg.n_layer):
            a = f"{p}.layers.{l}.self_attn."
            m = f"{p}.layers.{l}.mlp."
            w = LayerW()
            qkv_n = [a + "q_proj.weight", a + "k_proj.weight", a + "v_proj.weight"]
            gu_n = [m + "gate_proj.weight", m + "up_proj.weight"]
            w.qkv = store.fused(qkv_n)
            w.gu = store.fused(gu_n)
            w.o = store.views[a + "o_proj.weight"]
            w.down = store.views[m + "down_proj.weight"]
            w.ln1 = store.views[f"{p}.layers.{l}.input_layernorm.weight"]
            w.ln2 = store.views[f"{p}.layers.{l}.post_attention_layernorm.weight"]
            self.layers.append(w)
        self.norm = store.views[f"{p}.norm.weight"]
        self.embed = store.views[f"{p}.embed_tokens.weight"]
        self.names = [n for n in store.names if n.startswith(p + ".")]
        self.seg_start = store.offsets[self.names[0]]
        last = self.names[-1]
        self.seg_end = store.offsets[last] + (store.views[last].numel() + ALIGN - 1) // ALIGN * ALIGN
        self.main_grads = StackGrads(store, cfg, store.gflat, 0)
        if tiny_attention and cfg.head_dim != 256:
            raise lib.B200Error(f"inner stack head_dim {cfg.head_dim} unsupported (kernels are built for 256)")
        if not tiny_attention and cfg.head_dim != 64:
            raise lib.B200Error(f"outer stack head_dim {cfg.head_dim} unsupported (kernels are built for 64)")

    def fresh_grads(self) -> StackGrads:
        """A private gradient buffer for this stack (autograd mode hands these tensors to torch)."""
        buf = torch.empty(self.seg_end - self.seg_start, dtype=BF16, device=self.store.device)
        return StackGrads(self.store, self.cfg, buf, self.seg_start)

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, n_seq: int, S: int, inv_freq: torch.Tensor, save: bool):
        """x: [n_seq * S, H] inputs_embeds (row-major, sequences contiguous) -> (final-normed hidden, saved)."""
        c = self.cfg
        H, D, nh = c.hidden, c.head_dim, c.n_head
        cos, sin = ops.rope_table(inv_freq, S)
        saved = [] if save else None
        # Residual adds are fused into the norm that follows them (x + y is formed, rounded to bf16 and written by
        # the norm kernel), so every GEMM keeps the plain store epilogue.
        pending = None                                   # output of the previous layer's down_proj, not yet added
        for w in self.layers:
            if pending is None:
                n1, rstd1 = ops.rmsnorm(x, w.ln1, c.eps, want_rstd=True)
            else:
                x, n1, rstd1 = ops.add_rmsnorm(x, pending, w.ln1, c.eps)
            fuse_tiny = self.tiny and FUSE_ROPE and not FUSE_ROPE_FWD     # token-level stack: RoPE inside the attention kernel
            if FUSE_ROPE_FWD:
                qkv = ops.linear_rope(n1, w.qkv, cos, sin, S, D)      # QKV GEMM with RoPE in the epilogue
            else:
                qkv = ops.linear(n1, w.qkv)
                if not fuse_tiny:
                    ops.rope_qk_(qkv, cos, sin, S, H, D)
            if self.tiny:
                attn, lse = ops.attn_tiny_fwd(qkv, n_seq, S, nh, D, rope=(cos, sin) if fuse_tiny else None), None
            else:
                attn, lse = ops.attn_causal_fwd(qkv, n_seq, S, nh, D, want_lse=save)
            y1 = ops.linear(attn, w.o)
            h, n2, rstd2 = ops.add_rmsnorm(x, y1, w.ln2, c.eps)
            del y1
            if FUSE_SWIGLU and c.inner % 128 == 0:
                gu, act = ops.linear_swiglu(n2, w.gu)
            else:
                gu = ops.linear(n2, w.gu)
                act = ops.swiglu(gu)
            pending = ops.linear(act, w.down)
            if save:
                saved.append((x, n1, rstd1, qkv, attn, lse, h, n2, rstd2, gu, act))
            x = h
        x, y, rstd_f = ops.add_rmsnorm(x, pending, self.norm, c.eps)
        sv = dict(layers=saved, x_last=x, rstd_f=rstd_f, n_seq=n_seq, S=S, cos=cos, sin=sin) if save else None
        return y, sv

    # ------------------------------------------------------------------ backward
    def layer_range(self, li: int):
        """[start, end) of layer `li`'s parameters inside the flat parameter / gradient buffers."""
        p = self.cfg.prefix
        first = self.store.offsets[f"{p}.layers.{li}.self_attn.q_proj.weight"]
        if li + 1 < self.cfg.n_layer:
            end = self.store.offsets[f"{p}.layers.{li + 1}.self_attn.q_proj.weight"]
        else:
            end = self.store.offsets[f"{p}.norm.weight"]
        return first, end

    def backward(self, sv: dict, dy: torch.Tensor, grads: StackGrads, accumulate: bool = False,
                 layer_done=None) -> torch.Tensor:
        """dy: grad of the final-normed output.  Writes (or accumulates) every weight gradient of the stack into
        `grads` and returns the gradient w.r.t. the stack input (inputs_embeds)."""
        c = self.cfg
        H, D, nh = c.hidden, c.head_dim, c.n_head
        if sv is None or sv["layers"] is None:
            raise lib.B200Error("backward called without a saved forward")
        n_seq, S, cos, sin = sv["n_seq"], sv["S"], sv["cos"], sv["sin"]
        side = _side_stream(dy.device) if WGRAD_STREAM else None
        pending = []          # (event on the side stream, operands of the wgrads issued before it), released one layer late
        dx = ops.rmsnorm_bwd(dy, sv["x_last"], self.norm, sv["rstd_f"], None, grads.norm, accumulate)
        for li in range(len(self.layers) - 1, -1, -1):
            w = self.layers[li]
            g = grads.layers[li]
            hold = []
            x, n1, rstd1, qkv, attn, lse, h, n2, rstd2, gu, act = sv["layers"][li]
            sv["layers"][li] = None   # free as we go
            # ---- MLP block: x_out = h + down(act)
            dact = ops.linear_dgrad(dx, w.down)
            _wgrad(dx, act, g.down, accumulate, side, hold)
            del act
            dgu = ops.swiglu_bwd(gu, dact)
            del dact, gu
            dn2 = ops.linear_dgrad(dgu, w.gu)
            _wgrad(dgu, n2, g.gu, accumulate, side, hold)
            del dgu, n2
            dh = ops.rmsnorm_bwd(dn2, h, w.ln2, rstd2, dx, g.ln2, accumulate)
            del dn2, h, dx
            # ---- attention block: h = x + o(attn)
            dattn = ops.linear_dgrad(dh, w.o)
            _wgrad(dh, attn, g.o, accumulate, side, hold)
            rope = (cos, sin) if FUSE_ROPE else None
            if self.tiny:
                dqkv = ops.attn_tiny_bwd(qkv, dattn, n_seq, S, nh, D, rope=rope)
            else:
                dqkv = ops.attn_causal_bwd(qkv, attn, dattn, lse, n_seq, S, nh, D, rope=rope)
            del dattn, attn, qkv
            if not FUSE_ROPE:
                ops.rope_qk_(dqkv, cos, sin, S, H, D, backward=True)
            dn1 = ops.linear_dgrad(dqkv, w.qkv)
            _wgrad(dqkv, n1, g.qkv, accumulate, side, hold)
            del dqkv, n1
            dx = ops.rmsnorm_bwd(dn1, x, w.ln1, rstd1, dh, g.ln1, accumulate)
            del dn1, dh, x
            if side is not None:
                ev = torch.cuda.Event()
                ev.record(side)
                pending.append((ev, hold))
                if len(pending) > 1:          # the previous layer's wgrads have long finished: no stall, memory bounded
                    old_ev, old_hold = pending.pop(0)
                    torch.cuda.current_stream().wait_event(old_ev)
                    old_hold.clear()
            if layer_done is not None:
                if side is not None:
                    torch.cuda.current_stream().wait_stream(side)     # this layer's weight gradients are complete
                layer_done(li)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
            pending.clear()
        sv["layers"] = None
        return dx
