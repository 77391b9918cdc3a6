"""The reference's OWN layer classes (unmodified ``scaelum`` installed under baseline/_ref) as the
numerics oracle ("each kernel vs the reference layer class with identical weights"):

* CPU: state dicts interchange key by key; our fp32 eager path reproduces the reference's forward
  and backward to fp32 round-off for every registered BERT layer;
* GPU: the native sm_100a path at BERT-large geometry (H=1024, I=4096, 16 heads, 32 x 128 tokens)
  against scaelum/model/bert_layers.py:171-395 running in fp32 on the same GPU, per-tensor relative
  L2 tolerances; and with dropout ON, against the reference classes fed the EXACT masks the kernels
  used (regenerated on the host from the device RNG state).
"""
import copy
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scaelum")),
                                reason="reference not installed under baseline/_ref")


def _ref_layer_cls(name):
    for p in (os.path.join(ROOT, "baseline", "stubs"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from scaelum.registry import LAYER as REF_LAYER   # the unmodified reference package

    return REF_LAYER.get_module(name)


def _configs(hidden, heads, inter, p=0.0, vocab=1000):
    import skycomputing_b200 as sky  # noqa: F401
    from skycomputing_b200.models import BertConfig

    c = BertConfig(vocab, hidden_size=hidden, num_hidden_layers=1, num_attention_heads=heads,
                   intermediate_size=inter, max_position_embeddings=128,
                   hidden_dropout_prob=p, attention_probs_dropout_prob=p)
    return c


BLOCK = ["BertLayer_Head", "BertLayer_Body", "BertLayer_Tail"]


def _build_both(names, cfg, device):
    """[(reference layer, our layer)] with identical weights (reference initialises, we load)."""
    import skycomputing_b200 as sky

    pairs = []
    for n in names:
        if n == "BertTailForClassification":
            kw = dict(hidden_dropout_prob=cfg.hidden_dropout_prob, hidden_size=cfg.hidden_size,
                      num_classes=3)
        else:
            kw = dict(config=dict(cfg.__dict__))
        ref = _ref_layer_cls(n)(**copy.deepcopy(kw)).to(device)
        ours = sky.build_layer(n, **copy.deepcopy(kw)).to(device)
        missing = ours.load_state_dict(ref.state_dict())
        assert not missing.missing_keys and not missing.unexpected_keys
        assert list(ours.state_dict()) == list(ref.state_dict())     # same keys, same order
        pairs.append((ref, ours))
    return pairs


def _chain(layers, inputs):
    x = inputs
    for l in layers:
        x = l(*x) if isinstance(x, (tuple, list)) else l(x)
    return x if isinstance(x, (tuple, list)) else (x,)


def _rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_state_dicts_interchange_and_eager_path_equals_reference_cpu():
    from skycomputing_b200.models import set_backend

    torch.manual_seed(0)
    cfg = _configs(64, 4, 128)
    names = ["BertEmbeddings"] + BLOCK + ["BertPooler", "BertTailForClassification"]
    pairs = _build_both(names, cfg, "cpu")
    B, S = 3, 16
    ids = torch.randint(0, 1000, (B, S))
    tt = torch.randint(0, 2, (B, S))
    am = torch.ones(B, S, dtype=torch.long)
    am[1, 10:] = 0
    set_backend("torch")
    try:
        out_r = _chain([p[0] for p in pairs], (ids, tt, am))
        out_o = _chain([p[1] for p in pairs], (ids, tt, am))
        torch.testing.assert_close(out_o[0], out_r[0], rtol=1e-5, atol=1e-6)
        g = torch.randn_like(out_r[0])
        out_r[0].backward(g)
        out_o[0].backward(g)
    finally:
        set_backend("auto")
    for ref, ours in pairs:
        for (n, pr), (_, po) in zip(ref.named_parameters(), ours.named_parameters()):
            torch.testing.assert_close(po.grad, pr.grad, rtol=1e-4, atol=1e-6, msg=n)


def _block_inputs(B, S, H, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, S, H, device="cuda", generator=g)
    m = torch.ones(B, S, device="cuda")
    m[0, S - 30:] = 0
    m[B - 1, S // 2:] = 0
    ext = ((1.0 - m) * -10000.0).view(B, 1, 1, S)
    return x, ext


def _native_stage(ours_layers):
    from skycomputing_b200.builder import ModuleWrapper, SequentialWrapper

    return ModuleWrapper(rank=0, module=SequentialWrapper(*ours_layers), module_to_cuda=True,
                         cuda_device=0)


@pytest.mark.gpu
def test_native_block_matches_reference_classes_at_bert_large_geometry():
    """H=1024, I=4096, 16 heads, 32 x 128 tokens (the benchmark's shapes: the GEMM + LayerNorm
    epilogue kernel, 256-wide tiles, 4-CTA clusters are what runs here), two blocks."""
    from skycomputing_b200.models import set_backend

    torch.manual_seed(0)
    cfg = _configs(1024, 16, 4096, vocab=30522)
    pairs = _build_both(BLOCK * 2, cfg, "cuda")
    for ref, _ in pairs:                       # BERT-like scale instead of kaiming for the oracle
        for n, p in ref.named_parameters():
            if p.dim() == 2:
                torch.nn.init.normal_(p, std=0.02)
    for ref, ours in pairs:
        ours.load_state_dict(ref.state_dict())
    B, S, H = 32, 128, 1024
    x, ext = _block_inputs(B, S, H)
    xr = x.clone().requires_grad_(True)
    xn = x.clone().requires_grad_(True)
    cot = torch.randn(B, S, H, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    for ref, _ in pairs:
        ref.eval()                             # dropout p = 0 anyway; reference stays fp32
    out_r = _chain([p[0] for p in pairs], (xr, ext))[0]
    out_r.backward(cot)
    set_backend("native")
    try:
        stage = _native_stage([p[1] for p in pairs])
        stage.train()
        out_n = stage(xn, ext)[0]
        out_n.backward(cot.to(out_n.dtype))
    finally:
        set_backend("auto")
    # bf16 operands / fp32 accumulation vs an fp32 oracle: relative L2 error per tensor
    assert _rel_l2(out_n, out_r) < 1.5e-2
    assert _rel_l2(xn.grad, xr.grad) < 2.5e-2
    worst = {}
    for (ref, ours) in pairs:
        for (n, pr), (_, po) in zip(ref.named_parameters(), ours.named_parameters()):
            if n.endswith("key.bias"):
                _check_zero_gradient(ours, po)
                continue
            worst[n] = max(worst.get(n, 0.0), _rel_l2(po.grad, pr.grad))
    bad = {n: e for n, e in worst.items() if e > 3e-2}
    assert not bad, bad


def _check_zero_gradient(layer, key_bias):
    """The key bias has a mathematically ZERO gradient (softmax is invariant to a per-query shift
    of all scores): the reference leaves fp32 round-off there, the bf16 path leaves bf16 round-off;
    require it to be small against the query-bias gradient instead of comparing noise to noise."""
    q = dict(layer.named_parameters())["attention.self.query.bias"].grad
    assert float(key_bias.grad.abs().max()) < 0.1 * float(q.abs().max())


class _FixedMaskDropout(torch.nn.Module):
    """Stands in for nn.Dropout inside the REFERENCE classes: applies a given keep mask."""

    def __init__(self, keep, p):
        super().__init__()
        self.keep, self.p = keep, p

    def forward(self, x):
        return x * self.keep.to(x.dtype).view(x.shape) / (1.0 - self.p)


@pytest.mark.gpu
def test_exact_dropout_mask_parity_with_reference_classes():
    """Dropout ON (p = 0.1): the three masks of a block (attention probabilities, attention
    output, FFN output) are regenerated on the HOST from the device RNG state + the sites' stream
    ids and applied inside the reference's fp32 classes; forward and all gradients must then agree
    like in the dropout-free test (a wrong / shifted mask would give O(1) errors)."""
    from skycomputing_b200.models import default_rng, set_backend
    from skycomputing_b200.ops.dropout_ref import keep_mask_from_state

    torch.manual_seed(0)
    p = 0.1
    cfg = _configs(1024, 16, 4096, p=p, vocab=30522)
    pairs = _build_both(BLOCK, cfg, "cuda")
    for ref, _ in pairs:
        for n, prm in ref.named_parameters():
            if prm.dim() == 2:
                torch.nn.init.normal_(prm, std=0.02)
    for ref, ours in pairs:
        ours.load_state_dict(ref.state_dict())
    B, S, H, heads = 32, 128, 1024, 16
    x, ext = _block_inputs(B, S, H, seed=5)
    xr = x.clone().requires_grad_(True)
    xn = x.clone().requires_grad_(True)
    cot = torch.randn(B, S, H, device="cuda", generator=torch.Generator(device="cuda").manual_seed(6))
    set_backend("native")
    try:
        stage = _native_stage([pp[1] for pp in pairs])
        stage.train()
        out_n = stage(xn, ext)[0]
        out_n.backward(cot.to(out_n.dtype))
        span = stage.spans()[0]
        base = span.params.rng_base
        state = default_rng(torch.device("cuda", 0)).state
    finally:
        set_backend("auto")

    def mask(stream, shape):
        return torch.from_numpy(keep_mask_from_state(state, stream, shape, p)).cuda()

    head, _body, tail = (pp[0] for pp in pairs)
    head.attention.self.dropout = _FixedMaskDropout(mask(base + 1, (B, heads, S, S)), p)
    head.attention.output.dropout = _FixedMaskDropout(mask(base + 2, (B * S, H)), p)
    tail.output.dropout = _FixedMaskDropout(mask(base + 3, (B * S, H)), p)
    kept = float(head.attention.output.dropout.keep.float().mean())
    assert abs(kept - (1 - p)) < 5e-3
    out_r = _chain([pp[0] for pp in pairs], (xr, ext))[0]
    out_r.backward(cot)
    assert _rel_l2(out_n, out_r) < 1.5e-2, _rel_l2(out_n, out_r)
    assert _rel_l2(xn.grad, xr.grad) < 2.5e-2
    for (ref, ours) in pairs:
        for (n, pr), (_, po) in zip(ref.named_parameters(), ours.named_parameters()):
            if n.endswith("key.bias"):
                _check_zero_gradient(ours, po)
                continue
            assert _rel_l2(po.grad, pr.grad) < 3e-2, (n, _rel_l2(po.grad, pr.grad))
