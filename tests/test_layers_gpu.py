"""Native (sm_100a) layer spans vs the fp32 PyTorch oracle of the same layers: forward + all grads."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg(hidden=256, heads=4, inter=512, p=0.0):
    from skycomputing_b200.models import BertConfig

    c = BertConfig(1000, hidden_size=hidden, num_hidden_layers=2, num_attention_heads=heads,
                   intermediate_size=inter, max_position_embeddings=128,
                   hidden_dropout_prob=p, attention_probs_dropout_prob=p)
    return c


def _build(layer_types, cfg):
    from skycomputing_b200.builder import SequentialWrapper, build_layer

    layers = []
    for t in layer_types:
        if t == "BertTailForClassification":
            layers.append(build_layer(t, hidden_dropout_prob=cfg.hidden_dropout_prob,
                                      hidden_size=cfg.hidden_size, num_classes=3))
        else:
            layers.append(build_layer(t, config=cfg.__dict__))
    return SequentialWrapper(*layers)


def _max_rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-6)).item()


def _run_pair(layer_types, make_inputs, cfg, tol=4e-2, fused=True):
    """Build the stack twice with identical weights; oracle in fp32 eager, native through a
    ModuleWrapper (fused spans) or layer by layer."""
    from skycomputing_b200.builder import ModuleWrapper
    from skycomputing_b200.models import set_backend

    torch.manual_seed(0)
    ref = _build(layer_types, cfg).cuda()
    nat = copy.deepcopy(ref)
    ins_ref = make_inputs()
    ins_nat = [t.detach().clone().requires_grad_(t.requires_grad) if torch.is_tensor(t) else t
               for t in ins_ref]
    set_backend("torch")
    out_ref = ref(*ins_ref)
    out_ref = out_ref if isinstance(out_ref, (tuple, list)) else (out_ref,)
    set_backend("native")
    try:
        if fused:
            mw = ModuleWrapper(rank=0, module=nat, module_to_cuda=True, cuda_device=0)
            out_nat = mw(*ins_nat)
        else:
            out_nat = nat(*ins_nat)
            out_nat = out_nat if isinstance(out_nat, (tuple, list)) else (out_nat,)
        # compare every floating output that requires grad, then backprop the same cotangents
        g = torch.Generator(device="cuda").manual_seed(1)
        outs_r, outs_n, cots = [], [], []
        for o_r, o_n in zip(out_ref, out_nat):
            if torch.is_tensor(o_r) and o_r.is_floating_point() and o_r.requires_grad:
                assert _max_rel(o_n, o_r) < tol, f"forward mismatch {_max_rel(o_n, o_r)}"
                c = torch.randn(o_r.shape, device="cuda", generator=g)
                outs_r.append(o_r)
                outs_n.append(o_n)
                cots.append(c)
        torch.autograd.backward(outs_r, cots)
        torch.autograd.backward(outs_n, [c.to(o.dtype) for c, o in zip(cots, outs_n)])
    finally:
        set_backend("auto")
    # key biases have a mathematically zero gradient (softmax shift invariance): compare against
    # the overall gradient scale rather than each tensor's own (possibly ~0) magnitude
    gmax = max(p.grad.abs().max().item() for p in ref.parameters())
    for (n, p_r), (_, p_n) in zip(ref.named_parameters(), nat.named_parameters()):
        assert p_n.grad is not None, f"no grad for {n}"
        scale = max(p_r.grad.abs().max().item(), 0.05 * gmax)
        err = (p_n.grad.float() - p_r.grad.float()).abs().max().item() / scale
        assert err < tol, f"grad mismatch for {n}: {err}"
    for t_r, t_n in zip(ins_ref, ins_nat):
        if torch.is_tensor(t_r) and t_r.requires_grad:
            assert _max_rel(t_n.grad, t_r.grad) < tol, "input grad mismatch"


def _hidden_inputs(cfg, B=3, S=128, masked=True):
    def make():
        x = torch.randn(B, S, cfg.hidden_size, device="cuda", requires_grad=True)
        m = torch.ones(B, S, device="cuda")
        if masked:
            m[0, 90:] = 0
        ext = ((1.0 - m) * -10000.0).view(B, 1, 1, S)
        return [x, ext]
    return make


@pytest.mark.parametrize("fused", [True, False])
def test_full_block(fused):
    cfg = _cfg()
    _run_pair(["BertLayer_Head", "BertLayer_Body", "BertLayer_Tail"], _hidden_inputs(cfg), cfg,
              fused=fused)


def test_two_blocks_fused():
    cfg = _cfg()
    _run_pair(["BertLayer_Head", "BertLayer_Body", "BertLayer_Tail"] * 2, _hidden_inputs(cfg), cfg)


@pytest.mark.parametrize("types", [["BertLayer_Head"], ["BertLayer_Head", "BertLayer_Body"],
                                   ["BertLayer_Body", "BertLayer_Tail"]])
def test_partial_spans(types):
    cfg = _cfg()
    _run_pair(types, _hidden_inputs(cfg), cfg)


def test_tail_only_span():
    cfg = _cfg()

    def make():
        B, S = 2, 128
        inter = torch.randn(B, S, cfg.intermediate_size, device="cuda", requires_grad=True)
        a = torch.randn(B, S, cfg.hidden_size, device="cuda", requires_grad=True)
        ext = torch.zeros(B, 1, 1, S, device="cuda")
        return [inter, a, ext]

    _run_pair(["BertLayer_Tail"], make, cfg)


def test_embeddings_pooler_classifier():
    cfg = _cfg()

    def make():
        B, S = 4, 128
        ids = torch.randint(0, 1000, (B, S), device="cuda")
        tt = torch.randint(0, 2, (B, S), device="cuda")
        m = torch.ones(B, S, dtype=torch.long, device="cuda")
        m[1, 64:] = 0
        return [ids, tt, m]

    _run_pair(["BertEmbeddings", "BertLayer_Head", "BertLayer_Body", "BertLayer_Tail", "BertPooler",
               "BertTailForClassification"], make, cfg, tol=6e-2)


def test_training_step_decreases_loss_single_gpu():
    """ModuleWrapper + FusedSGD + native CE on one GPU with dropout ON: loss goes down."""
    from skycomputing_b200.builder import ModuleWrapper
    from skycomputing_b200.models import advance_rng, set_backend
    from skycomputing_b200.parallel import build_optimizer
    from skycomputing_b200.runner import build_loss

    cfg = _cfg(p=0.1)
    torch.manual_seed(0)
    set_backend("native")
    try:
        stack = _build(["BertEmbeddings", "BertLayer_Head", "BertLayer_Body", "BertLayer_Tail",
                        "BertPooler", "BertTailForClassification"], cfg)
        mw = ModuleWrapper(rank=0, module=stack, module_to_cuda=True, cuda_device=0)
        mw.train()
        opt = build_optimizer(mw, dict(optim_type="SGD", lr=0.05))
        loss_fn = build_loss(dict(type="CrossEntropyLoss"), torch.device("cuda"))
        B, S = 8, 128
        ids = torch.randint(0, 1000, (B, S), device="cuda")
        tt = torch.zeros(B, S, dtype=torch.long, device="cuda")
        m = torch.ones(B, S, dtype=torch.long, device="cuda")
        labels = torch.randint(0, 3, (B,), device="cuda")
        losses = []
        for _ in range(12):
            advance_rng()
            out = mw(ids, tt, m)
            loss = loss_fn(out[0], labels)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        assert losses[-1] < losses[0] * 0.8, losses
    finally:
        set_backend("auto")


def test_overwrite_first_gradients_match_zeroed_gradients():
    """FusedSGD leaves the wgrad-GEMM targets un-zeroed and the next step's first weight gradient
    overwrites them: after several steps the weights must equal those of the zero-every-step path."""
    from skycomputing_b200.builder import ModuleWrapper
    from skycomputing_b200.models import advance_rng, set_backend
    from skycomputing_b200.parallel import build_optimizer
    from skycomputing_b200.runner import build_loss

    cfg = _cfg(p=0.0)
    set_backend("native")
    try:
        finals = []
        for overwrite in (True, False):
            torch.manual_seed(0)
            stack = _build(["BertEmbeddings", "BertLayer_Head", "BertLayer_Body", "BertLayer_Tail",
                            "BertPooler", "BertTailForClassification"], cfg)
            mw = ModuleWrapper(rank=0, module=stack, module_to_cuda=True, cuda_device=0)
            mw.train()
            opt = build_optimizer(mw, dict(optim_type="SGD", lr=0.05))
            opt._overwrite_ok = overwrite
            loss_fn = build_loss(dict(type="CrossEntropyLoss"), torch.device("cuda"))
            g = torch.Generator(device="cuda").manual_seed(5)
            B, S = 8, 128
            ids = torch.randint(0, 1000, (B, S), device="cuda", generator=g)
            tt = torch.zeros(B, S, dtype=torch.long, device="cuda")
            m = torch.ones(B, S, dtype=torch.long, device="cuda")
            labels = torch.randint(0, 3, (B,), device="cuda", generator=g)
            for _ in range(5):
                advance_rng()
                loss_fn(mw(ids, tt, m)[0], labels).backward()
                opt.step()
            if overwrite:
                assert any(b.overwrite_first for b in opt.banks)
            finals.append(torch.cat([p.detach().float().flatten() for p in mw.parameters()]))
        diff = (finals[0] - finals[1]).abs().max().item()
        assert diff < 1e-4, diff
    finally:
        set_backend("auto")


def test_full_block_with_the_layernorm_fused_into_the_gemm(monkeypatch):
    """SKY_FUSE_LN=force: dense + bias + dropout + residual + LayerNorm of the attention output
    and of FFN2 run as ONE tcgen05 kernel each (2-CTA clusters at hidden 256) instead of the
    automatic policy's GEMM + standalone LayerNorm for launches this small."""
    from skycomputing_b200.ops import native as nat

    monkeypatch.setenv("SKY_FUSE_LN", "force")
    assert nat.gemm_ln_supported(3 * 128, 256)
    nat.enable_launch_counter()
    before = nat.launch_count()
    cfg = _cfg()
    _run_pair(["BertLayer_Head", "BertLayer_Body", "BertLayer_Tail"], _hidden_inputs(cfg), cfg)
    fused_launches = nat.launch_count() - before
    monkeypatch.setenv("SKY_FUSE_LN", "0")
    before = nat.launch_count()
    _run_pair(["BertLayer_Head", "BertLayer_Body", "BertLayer_Tail"], _hidden_inputs(cfg), cfg)
    assert (nat.launch_count() - before) - fused_launches == 2    # two LayerNorm launches gone


@pytest.mark.parametrize("S", [64, 256, 384])
def test_full_block_other_sequence_lengths_stay_native(S):
    """The reference's attention handles any S <= 512 (scaelum/model/bert_layers.py:254-275);
    here those lengths run the tiled tcgen05 attention kernels instead of an eager fallback."""
    from skycomputing_b200.ops import native as nat

    cfg = _cfg()
    cfg.max_position_embeddings = 512
    nat.enable_launch_counter()
    before = nat.launch_count()
    _run_pair(["BertLayer_Head", "BertLayer_Body", "BertLayer_Tail"],
              _hidden_inputs(cfg, B=2, S=S), cfg)
    assert nat.launch_count() - before >= 20      # the fused span ran (no eager per-layer path)
