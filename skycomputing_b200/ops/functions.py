"""Autograd functions that run the BERT layer library on the sm_100a kernels.

One ``torch.autograd.Function`` per *span* of registered layers, so a whole transformer block
(Head + Body + Tail of the reference's layer list) is 7 forward and ~17 backward kernel launches
with no framework work in between; spans that are cut by a stage boundary (Head | Body | Tail
alone or pairs) run the corresponding subset.

Conventions
-----------
* activations are bf16 ``[B, S, H]`` (viewed as ``[B*S, H]``), parameters are fp32 masters with
  bf16 shadows (``ParamBank``); weight/bias gradients are accumulated by the kernels straight into
  the fp32 ``.grad`` buffers (wgrad GEMM epilogue / colsum / LN param-grad kernels), so autograd
  receives ``None`` for every parameter input;
* dropout masks are never stored: every site has an RNG stream id and the backward kernels
  regenerate the forward mask from ``{seed, step}`` kept in device memory;
* stage-boundary fusion: ``in_ch`` / ``out_ch`` (``parallel.p2p.FusedChannel``) make the first
  GEMM wait on panel flags of the received activation, the last LayerNorm write into the next
  stage's HBM, the first dgrad GEMM write its input-gradient tiles into the previous stage's HBM,
  and the last LayerNorm-backward wait on the panels of the gradient it receives.

Reference parity: the math is that of scaelum/model/bert_layers.py:171-395 (see the per-op
citations in csrc/kernels/*.cu); tests/test_layers_gpu.py checks every span against the fp32
PyTorch oracle in ``skycomputing_b200.models.bert_layers``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import native as nat


# --------------------------------------------------------------------------------------------
# parameter bank: fp32 masters (possibly several nn.Parameters fused into one flat buffer) with a
# bf16 shadow and a flat fp32 gradient buffer
# --------------------------------------------------------------------------------------------
class ParamBank:
    """Fuses ``params`` (same trailing shape) along dim 0 into one flat fp32 master + grad buffer.

    The original ``nn.Parameter`` objects keep their names/shapes (state_dict compatibility) but
    become views into the flat storage; ``shadow()`` returns the bf16 compute copy, refreshed when
    a master was modified through torch (version counter) - the fused optimizer refreshes it
    itself in the same kernel that applies the update.
    """

    def __init__(self, params: Sequence[torch.nn.Parameter], need_shadow: bool = True):
        self.params = list(params)
        self.need_shadow = need_shadow
        self.flat: Optional[torch.Tensor] = None
        self.flat_grad: Optional[torch.Tensor] = None
        self._shadow: Optional[torch.Tensor] = None
        self._versions: List[int] = []
        self._ptrs: List[int] = []
        # weight banks whose gradient is produced by the wgrad GEMM: the fused optimizer may leave
        # the gradient un-zeroed (`overwrite_first`) and flag it `fresh`; the first wgrad of the
        # next step then STORES instead of accumulating (saves 4 B/parameter of SGD traffic and
        # turns that epilogue's red.add into plain stores)
        self.wgrad_target = False
        self.overwrite_first = False
        self.fresh = False

    def _materialise(self) -> None:
        p0 = self.params[0]
        rows = sum(p.shape[0] for p in self.params)
        shape = (rows,) + tuple(p0.shape[1:])
        flat = torch.empty(shape, dtype=torch.float32, device=p0.device)
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.shape[0]
                flat[off:off + n].copy_(p.data.float())
                p.data = flat[off:off + n]
                off += n
        self.flat = flat
        self.flat_grad = torch.zeros_like(flat)
        self._bind_grads()
        self._ptrs = [p.data_ptr() for p in self.params]
        self._shadow = None

    def _bind_grads(self) -> None:
        off = 0
        for p in self.params:
            n = p.shape[0]
            p.grad = self.flat_grad[off:off + n]
            off += n

    def ensure(self) -> None:
        if self.flat is None or any(p.data_ptr() != q for p, q in zip(self.params, self._ptrs)) \
                or self.flat.device != self.params[0].device:
            self._materialise()
        elif any(p.grad is None or p.grad.data_ptr() < self.flat_grad.data_ptr()
                 or p.grad.data_ptr() >= self.flat_grad.data_ptr() + self.flat_grad.numel() * 4
                 for p in self.params):
            # optimizer.zero_grad(set_to_none=True) or a foreign grad tensor: rebind (zeroed)
            self.flat_grad.zero_()
            self._bind_grads()

    def master(self) -> torch.Tensor:
        self.ensure()
        return self.flat

    def grad(self) -> torch.Tensor:
        self.ensure()
        return self.flat_grad

    def shadow(self) -> torch.Tensor:
        self.ensure()
        versions = [p._version for p in self.params]
        if self._shadow is None or versions != self._versions:
            if self._shadow is None:
                self._shadow = torch.empty(self.flat.shape, dtype=torch.bfloat16,
                                           device=self.flat.device)
            nat.cast_f32_to_bf16_(self.flat, self._shadow)
            self._versions = versions
        return self._shadow

    def mark_shadow_fresh(self) -> None:
        self._versions = [p._version for p in self.params]

    def sgd_descriptor(self, momentum_buf: Optional[torch.Tensor] = None,
                       second_moment: Optional[torch.Tensor] = None):
        self.ensure()
        if self.need_shadow and self._shadow is None:
            self.shadow()
        return (self.flat.data_ptr(), self.flat_grad.data_ptr(),
                0 if momentum_buf is None else momentum_buf.data_ptr(),
                self._shadow.data_ptr() if (self.need_shadow and self._shadow is not None) else 0,
                self.flat.numel(), 1 if self.overwrite_first else 0,
                0 if second_moment is None else second_moment.data_ptr())


class SpanParams:
    """Parameter banks + hyper-parameters of one transformer-block span (any of Head/Body/Tail)."""

    def __init__(self, head=None, body=None, tail=None):
        self.has_head = head is not None
        self.has_body = body is not None
        self.has_tail = tail is not None
        self.banks: List[ParamBank] = []
        self.heads = 0
        self.p_attn = 0.0
        self.p_hidden = 0.0
        self.eps = 1e-12
        self.rng: Optional[nat.RngState] = None
        self.rng_base = 0
        if head is not None:
            att = head.attention
            self.heads = att.self.num_attention_heads
            self.p_attn = float(att.self.dropout.p)
            self.p_hidden = float(att.output.dropout.p)
            self.eps = float(att.output.LayerNorm.eps)
            self.wqkv = self._bank([att.self.query.weight, att.self.key.weight, att.self.value.weight])
            self.bqkv = self._bank([att.self.query.bias, att.self.key.bias, att.self.value.bias], False)
            self.wo = self._bank([att.output.dense.weight])
            self.bo = self._bank([att.output.dense.bias], False)
            self.g1 = self._bank([att.output.LayerNorm.weight], False)
            self.b1n = self._bank([att.output.LayerNorm.bias], False)
        if body is not None:
            lin = body.intermediate.dense_act
            self.w1 = self._bank([lin.weight])
            self.b1 = self._bank([lin.bias], False)
        if tail is not None:
            out = tail.output
            self.p_hidden = float(out.dropout.p)
            self.eps = float(out.LayerNorm.eps)
            self.w2 = self._bank([out.dense.weight])
            self.b2 = self._bank([out.dense.bias], False)
            self.g2 = self._bank([out.LayerNorm.weight], False)
            self.b2n = self._bank([out.LayerNorm.bias], False)

    def _bank(self, params, need_shadow: bool = True) -> ParamBank:
        b = ParamBank(params, need_shadow)
        self.banks.append(b)
        return b

    def all_params(self) -> List[torch.nn.Parameter]:
        return [p for b in self.banks for p in b.params]


def _flat2d(t: torch.Tensor) -> torch.Tensor:
    return t.reshape(-1, t.shape[-1])


def _ch_kwargs_wait(ch, mb: int) -> dict:
    return dict(wait_flags=ch.act_flags_ptr(mb), wait_epoch=ch.epoch_ptr, wait_mult=ch.act_wait_mult,
                error_flag=ch.error_ptr)


class BertSpanFn(torch.autograd.Function):
    """forward(ctx, sp, training, in_ch, out_ch, mb, n_in, *inputs_then_params)."""

    @staticmethod
    def forward(ctx, sp: SpanParams, training: bool, in_ch, out_ch, mb: int, n_in: int, *tensors):
        inputs = tensors[:n_in]
        # slot index on the inbound link / on the outbound link: the same number in a plain
        # pipeline (micro-batch j), different ones on the wrap-around link of a looped pipeline
        mb_in, mb_out = mb if isinstance(mb, tuple) else (mb, mb)
        rng = sp.rng
        p_attn = sp.p_attn if training else 0.0
        p_hid = sp.p_hidden if training else 0.0
        if (p_attn > 0 or p_hid > 0) and rng is None:
            raise RuntimeError("dropout requested but the span has no RngState attached")
        saved = {}
        if sp.has_head:
            x3, mask = inputs[0], inputs[1]
            B, S, H = x3.shape
            x = _flat2d(x3)
            mask2 = None if mask is None else mask.reshape(B, S)
            wait = _ch_kwargs_wait(in_ch, mb_in) if in_ch is not None else {}
            qkv = nat.gemm(x, sp.wqkv.shadow(), bias=sp.bqkv.master(), **wait)
            ctxt, lse = nat.attention_fwd(qkv, mask2, B, S, sp.heads, dropout_p=p_attn, rng=rng,
                                          rng_stream=sp.rng_base + 1)
            ln_out = {}
            if out_ch is not None and not sp.has_body:
                ln_out = dict(y_ptr=out_ch.peer_act_ptr(mb_out), signal_flags=out_ch.peer_act_flags_ptr(mb_out))
            y1, z1, mean1, rstd1 = _gemm_ln(ctxt, sp.wo, sp.bo, x, sp.g1, sp.b1n, sp.eps, p_hid, rng,
                                            sp.rng_base + 2, ln_out)
            saved.update(x=x, mask2=mask2, qkv=qkv, ctxt=ctxt, lse=lse, z1=z1, mean1=mean1,
                         rstd1=rstd1)
            a = y1
            shape3 = (B, S, H)
        else:
            mask = inputs[-1]
            if sp.has_body:
                a3 = inputs[0]
                B, S, H = a3.shape
                a = _flat2d(a3)
            else:
                inter3, a3 = inputs[0], inputs[1]
                B, S, H = a3.shape
                a = _flat2d(a3)
            shape3 = (B, S, H)
        if sp.has_body:
            wait = _ch_kwargs_wait(in_ch, mb_in) if (in_ch is not None and not sp.has_head) else {}
            I = sp.w1.master().shape[0]
            h1 = torch.empty((a.shape[0], I), dtype=torch.bfloat16, device=a.device)
            inter = nat.gemm(a, sp.w1.shadow(), bias=sp.b1.master(), act=nat.ACT_GELU, out2=h1, **wait)
            saved.update(h1=h1)
        elif sp.has_tail:
            inter = _flat2d(inputs[0])
        if sp.has_tail:
            wait = _ch_kwargs_wait(in_ch, mb_in) if (in_ch is not None and not sp.has_body) else {}
            ln_out = {}
            if out_ch is not None:
                ln_out = dict(y_ptr=out_ch.peer_act_ptr(mb_out), signal_flags=out_ch.peer_act_flags_ptr(mb_out))
            y2, z2, mean2, rstd2 = _gemm_ln(inter, sp.w2, sp.b2, a, sp.g2, sp.b2n, sp.eps, p_hid, rng,
                                            sp.rng_base + 3, ln_out, **wait)
            saved.update(z2=z2, mean2=mean2, rstd2=rstd2)
        if sp.has_body or sp.has_tail:
            saved.update(a=a, inter=inter)
        ctx.sp, ctx.saved, ctx.training = sp, saved, training
        ctx.in_ch, ctx.out_ch, ctx.mb, ctx.n_in = in_ch, out_ch, mb, n_in
        ctx.n_tensors = len(tensors)
        ctx.shape3 = shape3
        B, S, H = shape3
        if sp.has_tail:
            out = y2.view(B, S, H) if y2 is not None else _dummy_out(a)
            return out
        if sp.has_body:
            # the pass-through attention output is only returned when this span produced it;
            # a span that starts at Body lets the caller forward its own input tensor
            if sp.has_head:
                return inter.view(B, S, -1), a.view(B, S, H)
            return inter.view(B, S, -1)
        return y1.view(B, S, H) if y1 is not None else _dummy_out(x)

    @staticmethod
    def backward(ctx, *grads):
        sp, sv = ctx.sp, ctx.saved
        in_ch, out_ch = ctx.in_ch, ctx.out_ch
        mb_in, mb_out = ctx.mb if isinstance(ctx.mb, tuple) else (ctx.mb, ctx.mb)
        rng = sp.rng
        p_attn = sp.p_attn if ctx.training else 0.0
        p_hid = sp.p_hidden if ctx.training else 0.0
        B, S, H = ctx.shape3
        M = B * S
        d_a_extra = None  # gradient flowing into `a` (attention output) besides the FFN1 dgrad
        d_inter = None
        g_in: List[Optional[torch.Tensor]] = [None] * ctx.n_in
        if sp.has_tail:
            wait = {}
            if out_ch is not None:
                dy2 = out_ch.grad_view(mb_out, M, H)
                wait = dict(wait_flags=out_ch.grad_flags_ptr(mb_out), wait_epoch=out_ch.epoch_ptr,
                            wait_mult=out_ch.grad_wait_mult, error_flag=out_ch.error_ptr)
            else:
                dy2 = _flat2d(grads[0]).contiguous()
            dz2, dz2d = nat.layernorm_bwd(dy2, sv["z2"], sv["mean2"], sv["rstd2"], sp.g2.master(),
                                          None, None, dropout_p=p_hid, rng=rng,
                                          rng_stream=sp.rng_base + 3, **wait)
            g2 = dz2d if dz2d is not None else dz2
            _ln_pgrad(dy2, sv["z2"], sv["mean2"], sv["rstd2"], sp.g2, sp.b2n, g2, sp.b2)
            _wgrad(g2, sv["inter"], sp.w2, None)
            if sp.has_body:
                d_h1 = nat.gemm(g2, sp.w2.shadow(), b_mn=True, aux=sv["h1"], act=nat.ACT_DGELU_MUL_AUX)
            else:
                send = _grad_send_kwargs(in_ch, mb_in, which=0)
                d_inter = nat.gemm(g2, sp.w2.shadow(), b_mn=True, **send)
            d_a_extra = dz2
        elif sp.has_body:
            # span ends after Body: grads = (d_inter, d_attn_out)
            gi = _flat2d(grads[0]).contiguous()
            d_h1 = torch.empty_like(gi)
            nat.ext().dgelu_mul(gi.data_ptr(), sv["h1"].data_ptr(), d_h1.data_ptr(), gi.numel(),
                                torch.cuda.current_stream().cuda_stream)
            d_a_extra = None
            if sp.has_head and len(grads) > 1 and grads[1] is not None:
                d_a_extra = _flat2d(grads[1]).contiguous()
        if sp.has_body:
            _wgrad(d_h1, sv["a"], sp.w1, sp.b1)
            send = _grad_send_kwargs(in_ch, mb_in, which=0) if not sp.has_head else {}
            d_a = nat.gemm(d_h1, sp.w1.shadow(), b_mn=True, aux=d_a_extra,
                           add_aux=d_a_extra is not None, **send)
        elif sp.has_tail:
            d_a = d_a_extra  # tail-only span: grad of the attn_out input is dz2
        else:
            d_a = None
        if sp.has_head:
            wait = {}
            if not sp.has_body:
                if out_ch is not None:
                    dy1 = out_ch.grad_view(mb_out, M, H)
                    wait = dict(wait_flags=out_ch.grad_flags_ptr(mb_out), wait_epoch=out_ch.epoch_ptr,
                                wait_mult=out_ch.grad_wait_mult, error_flag=out_ch.error_ptr)
                else:
                    dy1 = _flat2d(grads[0]).contiguous()
            else:
                dy1 = d_a
            dz1, dz1d = nat.layernorm_bwd(dy1, sv["z1"], sv["mean1"], sv["rstd1"], sp.g1.master(),
                                          None, None, dropout_p=p_hid, rng=rng,
                                          rng_stream=sp.rng_base + 2, **wait)
            g1 = dz1d if dz1d is not None else dz1
            _ln_pgrad(dy1, sv["z1"], sv["mean1"], sv["rstd1"], sp.g1, sp.b1n, g1, sp.bo)
            _wgrad(g1, sv["ctxt"], sp.wo, None)
            dctx = nat.gemm(g1, sp.wo.shadow(), b_mn=True)
            dqkv = nat.attention_bwd(sv["qkv"], sv["mask2"], sv["ctxt"], sv["lse"], dctx, B, S,
                                     sp.heads, dropout_p=p_attn, rng=rng, rng_stream=sp.rng_base + 1)
            _wgrad(dqkv, sv["x"], sp.wqkv, sp.bqkv)
            send = _grad_send_kwargs(in_ch, mb_in, which=0)
            dx = nat.gemm(dqkv, sp.wqkv.shadow(), b_mn=True, aux=dz1, add_aux=True, **send)
            g_in[0] = None if dx is None else dx.view(B, S, H)
        elif sp.has_body:
            g_in[0] = None if d_a is None else d_a.view(B, S, H)
        else:  # tail only: inputs (inter, attn_out, mask)
            g_in[0] = None if d_inter is None else d_inter.view(B, S, -1)
            g_in[1] = d_a.view(B, S, H)
        ctx.saved = None
        return (None, None, None, None, None, None, *g_in, *([None] * (ctx.n_tensors - ctx.n_in)))


# Weight-gradient deferral ("B before W"): the input-gradient chain of a stage is on the critical
# path of the pipeline (the previous stage waits for it), the weight / bias gradients are not.
# With deferral on, backward only queues them; the engine flushes the queue right after the
# stage's last dgrad GEMM has published its tiles to the previous stage.
_DEFER_WGRAD = [False]
_WGRAD_QUEUE: list = []
# "immediate" mode: every weight gradient is launched at once, but on a side stream that forks
# from the current point of the main stream, so it overlaps the rest of the dgrad chain.
_WGRAD_STREAM: list = [None]
_WGRAD_KEEP: list = []


def set_wgrad_stream(stream) -> None:
    _WGRAD_STREAM[0] = stream


def wgrad_keepalive() -> list:
    return _WGRAD_KEEP


def set_wgrad_deferral(on: bool) -> None:
    _DEFER_WGRAD[0] = bool(on)
    if not on:
        flush_wgrads()


def flush_wgrads() -> list:
    """Launch every queued weight / bias gradient on the CURRENT stream; returns the queued
    tensors so that a caller that flushed onto a side stream can keep them alive until it joins."""
    q = list(_WGRAD_QUEUE)
    _WGRAD_QUEUE.clear()
    for item in q:
        _run_param_grad(item)
    return q


def _run_param_grad(item) -> None:
    if item[0] == "ln":
        _, dy, z, mean, rstd, gbank, bbank, x2, bias_bank = item
        nat.ln_param_grad(dy, z, mean, rstd, gbank.grad(), bbank.grad(), x2,
                          None if bias_bank is None else bias_bank.grad())
    else:
        g, act, wbank, bbank = item
        wbank.wgrad_target = True
        first = wbank.overwrite_first and wbank.fresh
        wbank.fresh = False
        nat.gemm(g, act, a_mn=True, b_mn=True, out=wbank.grad(), accumulate=not first)
        if bbank is not None:  # else: the column sum rides in the LayerNorm parameter-gradient kernel
            nat.colsum_(g, bbank.grad())


def _ln_pgrad(dy, z, mean, rstd, gbank: ParamBank, bbank: ParamBank, x2=None,
              bias_bank: Optional[ParamBank] = None) -> None:
    """LayerNorm gamma / beta gradients: like the weight gradients they are not needed by the
    previous stage, so they follow the same deferral / side-stream policy."""
    item = ("ln", dy, z, mean, rstd, gbank, bbank, x2, bias_bank)
    if _DEFER_WGRAD[0]:
        _WGRAD_QUEUE.append(item)
        return
    side = _WGRAD_STREAM[0]
    if side is not None:
        side.wait_stream(torch.cuda.current_stream(dy.device))
        with torch.cuda.stream(side):
            _run_param_grad(item)
        _WGRAD_KEEP.append((dy, z, mean, rstd, x2))
        return
    _run_param_grad(item)


def pending_wgrads() -> int:
    return len(_WGRAD_QUEUE)


def _wgrad(g: torch.Tensor, act: torch.Tensor, wbank: ParamBank,
           bbank: Optional[ParamBank]) -> None:
    """dW += g^T act (both operands MN-major over the token dimension), db += colsum(g) unless
    ``bbank`` is None (the column sum then rides in the LayerNorm parameter-gradient kernel)."""
    item = (g, act, wbank, bbank)
    if _DEFER_WGRAD[0]:
        _WGRAD_QUEUE.append(item)
        return
    side = _WGRAD_STREAM[0]
    if side is not None:
        side.wait_stream(torch.cuda.current_stream(g.device))
        with torch.cuda.stream(side):
            _run_param_grad(item)
        _WGRAD_KEEP.append((g, act))
        return
    _run_param_grad(item)


def _dummy_out(like: torch.Tensor) -> torch.Tensor:
    """Stand-in output for spans whose real output was written into the next stage's HBM."""
    return torch.zeros(1, dtype=torch.bfloat16, device=like.device)


def _gemm_ln(a, wbank, bbank, residual, gbank, betabank, eps, p_drop, rng, rng_stream,
             ln_out: dict, **wait):
    """Dense + bias + dropout + residual + LayerNorm (the reference's BertSelfOutput / BertOutput,
    scaelum/model/bert_layers.py:285-289,323-327).  ONE tcgen05 kernel with the LayerNorm in the
    GEMM epilogue whenever the row fits a thread-block cluster; with ``ln_out`` (peer pointer + panel
    flags of a fused stage boundary) that kernel stores y straight into the next stage's HBM.  Otherwise GEMM, then the standalone LayerNorm.
    Returns (y | None, z, mean, rstd)."""
    M, N = a.shape[0], wbank.master().shape[0]
    # both ends of a fused boundary evaluate the same predicate on the same (rows, cols): the
    # consumer's expected signal count per panel follows from it (parallel/p2p.py)
    if nat.gemm_ln_supported(M, N):
        return nat.gemm_ln(a, wbank.shadow(), gbank.master(), betabank.master(), eps=eps,
                           bias=bbank.master(), residual=residual, dropout_p=p_drop, rng=rng,
                           rng_stream=rng_stream, y_ld=N, **ln_out, **wait)
    z = nat.gemm(a, wbank.shadow(), bias=bbank.master(), aux=residual, add_aux=True,
                 dropout_p=p_drop, rng=rng, rng_stream=rng_stream, **wait)
    y, mean, rstd = _ln_fwd(z, gbank.master(), betabank.master(), eps, **ln_out)
    return y, z, mean, rstd


def _ln_fwd(z, gamma, beta, eps, y_ptr: int = 0, signal_flags: int = 0):
    M, H = z.shape
    mean = torch.empty(M, dtype=torch.float32, device=z.device)
    rstd = torch.empty(M, dtype=torch.float32, device=z.device)
    if y_ptr:
        y = None
        yp = y_ptr
    else:
        y = torch.empty_like(z)
        yp = y.data_ptr()
    nat.ext().layernorm_fwd(z=z.data_ptr(), y=yp, mean=mean.data_ptr(), rstd=rstd.data_ptr(),
                            gamma=gamma.data_ptr(), beta=beta.data_ptr(), M=M, H=H, eps=eps,
                            signal_flags=signal_flags,
                            stream=torch.cuda.current_stream().cuda_stream)
    return y, mean, rstd


def _grad_send_kwargs(in_ch, mb: int, which: int) -> dict:
    """dgrad GEMM epilogue writes the input gradient straight into the previous stage's HBM."""
    if in_ch is None:
        return {}
    return dict(out_ptr=in_ch.peer_grad_ptr(mb), out_ld=in_ch.grad_ld,
                signal_flags=in_ch.peer_grad_flags_ptr(mb))


# --------------------------------------------------------------------------------------------
# embeddings / pooler / classifier / loss
# --------------------------------------------------------------------------------------------
class EmbeddingsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, training: bool, out_ch, mb: int, input_ids, token_type_ids,
                attention_mask, *params):
        B, S = input_ids.shape
        H = emb.word.master().shape[1]
        dev = input_ids.device
        ids = input_ids.reshape(-1).contiguous()
        tts = token_type_ids.reshape(-1).contiguous()
        am = attention_mask.reshape(-1).contiguous()
        out = torch.empty((B * S, H), dtype=torch.bfloat16, device=dev)
        ext_mask = torch.empty((B * S,), dtype=torch.float32, device=dev)
        mean = torch.empty(B * S, dtype=torch.float32, device=dev)
        rstd = torch.empty(B * S, dtype=torch.float32, device=dev)
        p = emb.p_drop if training else 0.0
        nat.ext().embed_fwd(ids=ids.data_ptr(), tts=tts.data_ptr(), amask=am.data_ptr(),
                            word=emb.word.master().data_ptr(), pos=emb.pos.master().data_ptr(),
                            type=emb.type.master().data_ptr(), gamma=emb.gamma.master().data_ptr(),
                            beta=emb.beta.master().data_ptr(), out=out.data_ptr(),
                            ext_mask=ext_mask.data_ptr(), mean=mean.data_ptr(), rstd=rstd.data_ptr(),
                            B=B, S=S, H=H, eps=emb.eps, dropout_p=p,
                            rng_state=0 if emb.rng is None else emb.rng.ptr,
                            rng_stream=emb.rng_base, stream=torch.cuda.current_stream().cuda_stream)
        ctx.emb, ctx.training = emb, training
        ctx.saved = (ids, tts, mean, rstd)
        ctx.dims = (B, S, H)
        ctx.n_params = len(params)
        ctx.mark_non_differentiable(ext_mask)
        return out.view(B, S, H), ext_mask.view(B, 1, 1, S)

    @staticmethod
    def backward(ctx, dout, _dmask):
        emb = ctx.emb
        ids, tts, mean, rstd = ctx.saved
        B, S, H = ctx.dims
        p = emb.p_drop if ctx.training else 0.0
        dout = dout.reshape(B * S, H).contiguous()
        nat.ext().embed_bwd(dout=dout.data_ptr(), ids=ids.data_ptr(), tts=tts.data_ptr(),
                            word=emb.word.master().data_ptr(), pos=emb.pos.master().data_ptr(),
                            type=emb.type.master().data_ptr(), gamma=emb.gamma.master().data_ptr(),
                            mean=mean.data_ptr(), rstd=rstd.data_ptr(),
                            dword=emb.word.grad().data_ptr(), dpos=emb.pos.grad().data_ptr(),
                            dtype=emb.type.grad().data_ptr(), dgamma=emb.gamma.grad().data_ptr(),
                            dbeta=emb.beta.grad().data_ptr(), B=B, S=S, H=H,
                            type_rows=emb.type.master().shape[0], dropout_p=p,
                            rng_state=0 if emb.rng is None else emb.rng.ptr,
                            rng_stream=emb.rng_base, stream=torch.cuda.current_stream().cuda_stream)
        ctx.saved = None
        return (None, None, None, None, None, None, None, *([None] * ctx.n_params))


class EmbeddingParams:
    def __init__(self, mod):
        self.word = ParamBank([mod.word_embeddings.weight], False)
        self.pos = ParamBank([mod.position_embeddings.weight], False)
        self.type = ParamBank([mod.token_type_embeddings.weight], False)
        self.gamma = ParamBank([mod.LayerNorm.weight], False)
        self.beta = ParamBank([mod.LayerNorm.bias], False)
        self.banks = [self.word, self.pos, self.type, self.gamma, self.beta]
        self.eps = float(mod.LayerNorm.eps)
        self.p_drop = float(mod.dropout.p)
        self.rng: Optional[nat.RngState] = None
        self.rng_base = 0

    def all_params(self):
        return [p for b in self.banks for p in b.params]


class SmallLinearFn(torch.autograd.Function):
    """Pooler (first-token gather + dense + tanh) and classifier (input dropout + dense)."""

    @staticmethod
    def forward(ctx, lp, training: bool, x, *params):
        w, b = lp.w.master(), lp.b.master()
        N, K = w.shape
        if lp.first_token:
            B, S, H = x.shape
            M, ldx = B, S * H
            xs = x.contiguous()
        else:
            xs = x.contiguous()
            M, ldx = xs.shape[0], xs.shape[1]
        x_bf16 = xs.dtype == torch.bfloat16
        y = torch.empty((M, N), dtype=torch.float32, device=xs.device)
        p = lp.p_drop if training else 0.0
        nat.ext().small_linear_fwd(x=xs.data_ptr(), x_bf16=x_bf16, ldx=ldx, w=w.data_ptr(),
                                   b=b.data_ptr(), y=y.data_ptr(), M=M, N=N, K=K,
                                   act_tanh=1 if lp.act_tanh else 0, dropout_p=p,
                                   rng_state=0 if lp.rng is None else lp.rng.ptr,
                                   rng_stream=lp.rng_base, stream=torch.cuda.current_stream().cuda_stream)
        ctx.lp, ctx.training = lp, training
        ctx.saved = (xs, y, M, ldx, x_bf16)
        ctx.n_params = len(params)
        return y

    @staticmethod
    def backward(ctx, dy):
        lp = ctx.lp
        xs, y, M, ldx, x_bf16 = ctx.saved
        w = lp.w.master()
        N, K = w.shape
        dy = dy.contiguous().float()
        p = lp.p_drop if ctx.training else 0.0
        if lp.first_token:
            dx = torch.zeros_like(xs)
            lddx = ldx
        else:
            dx = torch.empty_like(xs)
            lddx = K
        nat.ext().small_linear_bwd(x=xs.data_ptr(), x_bf16=x_bf16, ldx=ldx, w=w.data_ptr(),
                                   y=y.data_ptr(), dy=dy.data_ptr(), dx=dx.data_ptr(),
                                   dx_bf16=x_bf16, lddx=lddx, dw=lp.w.grad().data_ptr(),
                                   db=lp.b.grad().data_ptr(), M=M, N=N, K=K,
                                   act_tanh=1 if lp.act_tanh else 0, dropout_p=p,
                                   rng_state=0 if lp.rng is None else lp.rng.ptr,
                                   rng_stream=lp.rng_base, stream=torch.cuda.current_stream().cuda_stream)
        ctx.saved = None
        return (None, None, dx, *([None] * ctx.n_params))


class PoolerFn(torch.autograd.Function):
    """BertPooler on the tensor-core GEMM: first-token rows are read through a strided TMA view
    (no gather copy), bias + tanh run in the epilogue; backward = two more GEMMs.

    Reference: scaelum/model/bert_layers.py:381-395.
    """

    @staticmethod
    def forward(ctx, lp, hidden, *params):
        B, S, H = hidden.shape
        hidden = hidden.contiguous()
        x0 = hidden.view(B, S * H)[:, :H]          # [B, H] view, row stride S*H
        y = nat.gemm(x0, lp.w.shadow(), bias=lp.b.master(), act=nat.ACT_TANH,
                     out_dtype=torch.float32)
        ctx.lp = lp
        ctx.saved = (hidden, y)
        ctx.n_params = len(params)
        return y

    @staticmethod
    def backward(ctx, dy):
        lp = ctx.lp
        hidden, y = ctx.saved
        B, S, H = hidden.shape
        x0 = hidden.view(B, S * H)[:, :H]
        g = (dy.float() * (1.0 - y * y))
        lp.b.grad().add_(g.sum(0))
        gb = g.to(torch.bfloat16)
        # dW[n,k] += sum_b g[b,n] x0[b,k]  (both operands MN-major over the batch dimension)
        nat.gemm(gb, x0, a_mn=True, b_mn=True, out=lp.w.grad(), accumulate=True)
        dhidden = torch.zeros_like(hidden)
        nat.gemm(gb, lp.w.shadow(), b_mn=True, out=dhidden.view(B, S * H)[:, :H])
        ctx.saved = None
        return (None, dhidden, *([None] * ctx.n_params))


class SmallLinearParams:
    def __init__(self, weight, bias, act_tanh: bool, first_token: bool, p_drop: float = 0.0):
        # the pooler (first_token) also runs on the GEMM kernel -> keep a bf16 shadow of its weight
        self.w = ParamBank([weight], need_shadow=first_token)
        self.b = ParamBank([bias], False)
        self.banks = [self.w, self.b]
        self.act_tanh = act_tanh
        self.first_token = first_token
        self.p_drop = p_drop
        self.rng: Optional[nat.RngState] = None
        self.rng_base = 0

    def all_params(self):
        return [p for b in self.banks for p in b.params]


class SoftmaxCrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        loss, dlogits = nat.softmax_ce(logits.contiguous().float(), labels.contiguous())
        ctx.save_for_backward(dlogits)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        (dlogits,) = ctx.saved_tensors
        return dlogits * dloss, None
