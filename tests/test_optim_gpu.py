"""Fused multi-tensor optimizers (sgd_multi / adam_multi) vs torch.optim on the same parameters,
eager and replayed from a CUDA graph (the device-side step counter must advance per replay)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _stage():
    import skycomputing_b200 as sky
    from skycomputing_b200.models import BertConfig, set_backend

    set_backend("native")
    cfg = BertConfig(1000, hidden_size=256, num_hidden_layers=1, num_attention_heads=4,
                     intermediate_size=512, max_position_embeddings=128,
                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    mc = [dict(layer_type="BertLayer_Head", config=cfg.__dict__),
          dict(layer_type="BertLayer_Body", config=cfg.__dict__),
          dict(layer_type="BertLayer_Tail", config=cfg.__dict__)]
    torch.manual_seed(0)
    return sky.build_module_from_cfg(0, mc, dict(module_to_cuda=True, cuda_device=0))


@pytest.mark.parametrize("optim_type,kw", [
    ("Adam", dict(lr=1e-3)),
    ("Adam", dict(lr=2e-3, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.01)),
    ("AdamW", dict(lr=1e-3, weight_decay=0.05)),
    ("SGD", dict(lr=1e-2, momentum=0.9, weight_decay=1e-4)),
])
def test_fused_optimizer_matches_torch(optim_type, kw):
    import skycomputing_b200 as sky
    from skycomputing_b200.parallel.optim import FusedAdam, FusedSGD

    stage = _stage()
    opt = sky.build_optimizer(stage, dict(optim_type=optim_type, **kw))
    assert isinstance(opt, FusedAdam if optim_type.startswith("Adam") else FusedSGD)
    assert opt.graph_safe
    params = [p for p in stage.parameters()]
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in params]
    ref = getattr(torch.optim, optim_type)(ref_params, **kw)
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(4):
        for p, r in zip(params, ref_params):
            grad = torch.randn(p.shape, generator=g, device="cuda") * 0.1
            p.grad.copy_(grad)
            r.grad = grad.clone()
        opt.step()
        ref.step()
        for p, r in zip(params, ref_params):
            torch.testing.assert_close(p.detach(), r.detach(), rtol=2e-5, atol=2e-6)
            assert float(p.grad.abs().max()) == 0.0        # zeroed in the same launch
    # the bf16 compute shadows were refreshed by the kernel
    for bank in opt.banks:
        if bank.need_shadow:
            torch.testing.assert_close(bank._shadow.float(), bank.flat.bfloat16().float())


def test_fused_adam_replays_from_a_cuda_graph():
    import skycomputing_b200 as sky

    stage = _stage()
    opt = sky.build_optimizer(stage, dict(optim_type="Adam", lr=1e-3))
    params = [p for p in stage.parameters()]
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in params]
    ref = torch.optim.Adam(ref_params, lr=1e-3)
    grads = [torch.randn_like(p) * 0.1 for p in params]

    def one_step():
        for p, gr in zip(params, grads):
            p.grad.copy_(gr)
        opt.step()

    one_step()                                   # eager step 1 (also builds the descriptors)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        one_step()                               # captured, NOT executed
    for _ in range(3):
        graph.replay()                           # steps 2..4
    torch.cuda.synchronize()
    for _ in range(4):
        for r, gr in zip(ref_params, grads):
            r.grad = gr.clone()
        ref.step()
    for p, r in zip(params, ref_params):
        torch.testing.assert_close(p.detach(), r.detach(), rtol=2e-5, atol=2e-6)
    assert opt.state_dict()["step"] == 4


def test_unsupported_options_fall_back_to_torch_optim_and_are_not_graph_safe():
    import skycomputing_b200 as sky
    from skycomputing_b200.parallel.optim import TorchOptimizerAdapter

    stage = _stage()
    for cfg in (dict(optim_type="Adam", lr=1e-3, amsgrad=True),
                dict(optim_type="SGD", lr=1e-3, momentum=0.9, nesterov=True),
                dict(optim_type="RMSprop", lr=1e-3)):
        opt = sky.build_optimizer(stage, cfg)
        assert isinstance(opt, TorchOptimizerAdapter) and not opt.graph_safe
