"""KV-cached inference of the two stacks and the device-side generate() loop.

Replaces hf DynamicCache's per-step `torch.cat` (cache_utils.py:102-121) with a paged KV cache
(pages of 64 positions for the outer stack, 8 for the inner one; per-row block tables) that the
QKV/RoPE step appends to and a split-T single-query attention kernel reads.  The sampling step
(midi_model.py:202-223) -- grammar mask, temperature softmax, top-p / top-k, draw -- is one
kernel per token with the grammar held on the device as id ranges, so the only host<->device
traffic per generated event is the 8-byte-per-row read of the event-type token that the
reference's `end` / early-exit logic needs (midi_model.py:224-237, 248).
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import lib, ops
from .engine import StackCfg, StackEngine

BF16 = torch.bfloat16


class PagedKV:
    """Per-stack paged KV cache for `batch` rows and up to `capacity` positions."""

    def __init__(self, cfg: StackCfg, batch: int, capacity: int, page: int, device):
        self.cfg, self.batch, self.page = cfg, batch, page
        self.max_pages = (capacity + page - 1) // page
        self.capacity = self.max_pages * page
        n_pages = batch * self.max_pages
        shape = (n_pages, cfg.n_head, page, cfg.head_dim)
        self.k = [torch.empty(shape, dtype=BF16, device=device) for _ in range(cfg.n_layer)]
        self.v = [torch.empty(shape, dtype=BF16, device=device) for _ in range(cfg.n_layer)]
        # identity block table: row b owns pages [b*max_pages, (b+1)*max_pages)
        self.block_table = torch.arange(n_pages, dtype=torch.int32, device=device).view(batch, self.max_pages).contiguous()
        self.length = 0

    def reset(self):
        self.length = 0


def _linear(x: torch.Tensor, w: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Linear on a few rows: weight-streaming skinny GEMM for <= 16 rows, tcgen05 GEMM otherwise."""
    M, K = x.shape
    N = w.shape[0]
    if M > 16:
        return ops.linear(x, w, residual=residual)
    y = torch.empty((M, N), dtype=BF16, device=x.device)
    lib.call("b200_gemv_bf16", x.data_ptr(), w.data_ptr(), lib.ptr(residual), y.data_ptr(), M, N, K, x.stride(0),
             w.stride(0), residual.stride(0) if residual is not None else 0, N, lib.stream())
    return y


def _lm_head(x: torch.Tensor, w: torch.Tensor, pitch: int) -> torch.Tensor:
    M, K = x.shape
    N = w.shape[0]
    if M > 16:
        return ops.linear(x, w, pitch=pitch)
    y = torch.empty((M, pitch), dtype=BF16, device=x.device)
    lib.call("b200_gemv_bf16", x.data_ptr(), w.data_ptr(), None, y.data_ptr(), M, N, K, x.stride(0), w.stride(0), 0, pitch,
             lib.stream())
    return y


class CachedStack:
    """Incremental forward of one stack over a PagedKV (inference only)."""

    def __init__(self, eng: StackEngine, max_pos: int, inv_freq: torch.Tensor):
        self.eng = eng
        self.cos, self.sin = ops.rope_table(inv_freq, max_pos)
        self.max_pos = max_pos

    def step(self, x: torch.Tensor, kv: PagedKV, s_new: int) -> torch.Tensor:
        """x: [batch * s_new, H] new inputs_embeds; appends to kv; returns final-normed hidden for the new rows."""
        c = self.eng.cfg
        H, D, nh = c.hidden, c.head_dim, c.n_head
        B = kv.batch
        past = kv.length
        T = past + s_new
        if T > kv.capacity or T > self.max_pos:
            raise lib.B200Error(f"KV cache overflow: {T} positions > capacity {min(kv.capacity, self.max_pos)}")
        scale = 1.0 / math.sqrt(D)
        n_split = max(1, min(32, (T + 255) // 256)) if D == 64 else 1
        ws_bytes = lib.query("b200_attn_decode_workspace_bytes", B * s_new, nh, D, n_split)
        for li, w in enumerate(self.eng.layers):
            n1 = ops.rmsnorm(x, w.ln1, c.eps)
            qkv = _linear(n1, w.qkv)
            ops.rope_qk_(qkv, self.cos, self.sin, s_new, H, D, pos0=past)
            lib.call("b200_kv_append", qkv.data_ptr(), kv.k[li].data_ptr(), kv.v[li].data_ptr(), kv.block_table.data_ptr(),
                     kv.max_pages, kv.page, nh, D, B, s_new, past, None, qkv.stride(0), lib.stream())
            if past == 0 and s_new > 1 and D == 64:
                attn, _ = ops.attn_causal_fwd(qkv, B, s_new, nh, D, want_lse=False)
            elif past == 0 and s_new > 1 and D == 256 and s_new <= 8:
                attn = ops.attn_tiny_fwd(qkv, B, s_new, nh, D)
            else:
                attn = torch.empty((B * s_new, H), dtype=BF16, device=x.device)
                ws = ops._ws("attn_decode", ws_bytes, x.device)
                lib.call("b200_attn_decode", qkv.data_ptr(), kv.k[li].data_ptr(), kv.v[li].data_ptr(),
                         kv.block_table.data_ptr(), kv.max_pages, kv.page, attn.data_ptr(), B, s_new, nh, D, past, None, T,
                         qkv.stride(0), H, scale, n_split, ws.data_ptr(), ws.numel(), lib.stream())
            h = _linear(attn, w.o, residual=x)
            n2 = ops.rmsnorm(h, w.ln2, c.eps)
            gu = _linear(n2, w.gu)
            act = ops.swiglu(gu)
            x = _linear(act, w.down, residual=h)
        kv.length = T
        return ops.rmsnorm(x, self.eng.norm, c.eps)


class GrammarLUT:
    """Device copy of the tokenizer grammar as id ranges (midi_tokenizer.py:517-535)."""

    def __init__(self, tok, device):
        self.eos, self.pad = tok.eos_id, tok.pad_id
        ev_ids = sorted(tok.event_ids.values())
        if ev_ids != list(range(self.eos + 1, self.eos + 1 + len(ev_ids))):
            raise lib.B200Error("event ids are not contiguous after eos: the range-based grammar does not apply")
        self.n_event_types = len(ev_ids)
        lut = np.zeros((self.n_event_types, 8, 2), dtype=np.int32)
        self.n_params = {}
        for name, params in tok.events.items():
            e = tok.event_ids[name] - (self.eos + 1)
            self.n_params[tok.event_ids[name]] = len(params)
            for i, pn in enumerate(params):
                ids = tok.parameter_ids[pn]
                if list(ids) != list(range(ids[0], ids[0] + len(ids))):
                    raise lib.B200Error(f"parameter ids of {pn} are not contiguous")
                lut[e, i] = (ids[0], ids[-1] + 1)
        self.lut = torch.from_numpy(lut).to(device)


def sample_from_logits(logits: torch.Tensor, V: int, temp: float, top_p: float, top_k: int, step: int,
                       event_tok: torch.Tensor, g: GrammarLUT, uniforms: torch.Tensor, out: torch.Tensor,
                       dense_mask: Optional[torch.Tensor] = None):
    """logits [B, pitch] bf16 -> out[:, step] (int64 [B, 8] event buffer)."""
    B = logits.shape[0]
    lib.call("b200_sample_from_logits", logits.data_ptr(), B, V, logits.stride(0), temp, top_p, top_k, step,
             event_tok.data_ptr(), g.lut.data_ptr(), g.n_event_types, g.eos, g.pad, lib.ptr(dense_mask), uniforms.data_ptr(),
             out.data_ptr() + 8 * step, out.stride(0), lib.stream())
