"""Regression tests for the round-1 review findings (ADVICE.md / VERDICT.md): checkpoint load with
parameter-less layers, dtype-preserving state dicts, short / indivisible batches, the table-driven
hook protocol, graph-safety of optimizers, shape-aware fused-boundary negotiation."""
import os
import warnings

import pytest
import torch

import skycomputing_b200 as sky


# ---------------------------------------------------------------------------- hooks
def test_hook_table_and_dispatch():
    from skycomputing_b200.runner import hooks as H

    assert len(H.GENERIC_CALLBACKS) == 6 and len(H.MODE_CALLBACKS) == 8
    assert set(H.ALL_CALLBACKS) == {
        "before_run", "after_run", "before_epoch", "after_epoch", "before_iter", "after_iter",
        "before_train_epoch", "after_train_epoch", "before_val_epoch", "after_val_epoch",
        "before_train_iter", "after_train_iter", "before_val_iter", "after_val_iter"}
    seen = []

    class A(sky.Hook):
        def after_iter(self, runner):
            seen.append("after_iter")

    class B(sky.Hook):
        def before_train_epoch(self, runner):
            seen.append("before_train_epoch")

    assert A.overrides() == {"after_iter", "after_train_iter", "after_val_iter"}
    assert B.overrides() == {"before_train_epoch"}
    assert sky.Hook.overrides() == frozenset()
    a = A()
    a.fire(None, "after_train_iter")          # mode-specific falls through to the generic one
    a.fire(None, "before_run")                # default: no-op
    B().fire(None, "before_train_epoch")
    B().fire(None, "before_val_epoch")        # falls through to the (default) before_epoch
    assert seen == ["after_iter", "before_train_epoch"]
    with pytest.raises(AttributeError):
        a.fire(None, "after_train_step")      # a typo is an error, not a silent no-op

    class R:
        epoch, inner_iter, iter = 3, 4, 9
        data_loader = [0] * 5

    h = sky.Hook()
    assert h.every_n_epochs(R, 2) and not h.every_n_epochs(R, 3) and not h.every_n_epochs(R, None)
    assert h.every_n_inner_iters(R, 5) and h.every_n_iters(R, 10) and not h.every_n_iters(R, 0)
    assert h.end_of_epoch(R)


# ---------------------------------------------------------------------------- checkpoints
def _mlp_cfg():
    return [dict(layer_type="Linear", in_features=6, out_features=5), dict(layer_type="ReLU"),
            dict(layer_type="Linear", in_features=5, out_features=3)]


def _mlp_runner(tmp, steps, hooks):
    cfg = _mlp_cfg()
    wm = sky.WorkerManager(first_rank=0)
    wm.load_worker_pool_from_config([dict(name="w0", server_config={}, device=0,
                                          extra_config=dict(slowdown=0, mem_limit=-1,
                                                            timer_config=dict(root=tmp)))])
    wm = sky.Allocator(cfg, wm).even_allocate()
    model = sky.RpcModel(wm, this_rank=0)
    opt = sky.build_optimizer(model.local_stage, dict(optim_type="SGD", lr=0.1, momentum=0.9))
    runner = sky.Runner(model=model, parameter_server=sky.ParameterServer(cfg, lazy=True),
                        worker_manager=wm, optimizer=opt, max_epochs=1, max_iters=steps,
                        loss_cfg=dict(type="CrossEntropyLoss"), timer_cfg=dict(root=tmp),
                        logging_cfg=dict(mode="a", filename=os.path.join(tmp, "run.log")))
    for h in hooks:
        runner.register_hook(h)
    return runner, model


def _mlp_loader(n=8, batch=4):
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(n, 6, generator=g), torch.randint(0, 3, (n,), generator=g)
    return [(x[i:i + batch], y[i:i + batch]) for i in range(0, n, batch)]


def test_checkpoint_roundtrip_with_a_parameterless_layer(tmp_path):
    """ADVICE: save worked, load raised KeyError: 1 for Linear / ReLU / Linear."""
    tmp = str(tmp_path)
    runner, model = _mlp_runner(tmp, 2, [sky.CheckpointHook(save_path=tmp, save_interval=1)])
    runner.train(_mlp_loader())
    ckpt = os.path.join(tmp, "epoch_1.pth")
    keys = set(torch.load(ckpt).keys())
    assert keys == {"0.weight", "0.bias", "2.weight", "2.bias"}       # nothing for the ReLU
    trained = [p.detach().clone() for p in model.local_stage.parameters()]
    runner2, model2 = _mlp_runner(tmp, 0, [sky.CheckpointHook(load_checkpoint_from=ckpt)])
    runner2.train(_mlp_loader())
    for a, b in zip(trained, model2.local_stage.parameters()):
        assert torch.equal(a, b.detach())
    lazy = sky.ParameterServer(_mlp_cfg(), lazy=True)
    lazy.load_weights_from_file(ckpt)
    assert lazy.get_state_dict(1) == {} and set(lazy.get_state_dict(2)) == {"weight", "bias"}
    with pytest.raises(IndexError):
        lazy.get_state_dict(3)


def test_optimizer_restore_failure_is_reported_not_swallowed(tmp_path):
    tmp = str(tmp_path)
    runner, _ = _mlp_runner(tmp, 2, [sky.CheckpointHook(save_path=tmp, save_interval=1)])
    runner.train(_mlp_loader())
    ckpt = os.path.join(tmp, "epoch_1.pth")
    extra = sky.CheckpointHook._extra_path(ckpt, 0)
    st = torch.load(extra)
    st["optimizer"] = {"garbage": 1}
    torch.save(st, extra)
    runner2, _ = _mlp_runner(tmp, 0, [sky.CheckpointHook(load_checkpoint_from=ckpt,
                                                        resume_training_state=True)])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        runner2.train(_mlp_loader())
    assert any("NOT restored" in str(x.message) for x in w)
    runner3, _ = _mlp_runner(tmp, 0, [sky.CheckpointHook(load_checkpoint_from=ckpt,
                                                        resume_training_state=True,
                                                        strict_optimizer=True)])
    with pytest.raises(RuntimeError, match="NOT restored"):
        runner3.train(_mlp_loader())


def test_weights_to_cpu_keeps_integer_buffers():
    bn = torch.nn.BatchNorm1d(4)
    bn(torch.randn(8, 4))
    sd = sky.utils.weights_to_cpu(bn.state_dict())
    assert sd["num_batches_tracked"].dtype == torch.int64
    assert sd["running_mean"].dtype == torch.float32
    half = sky.utils.weights_to_cpu({"w": torch.ones(2, dtype=torch.bfloat16)})
    assert half["w"].dtype == torch.float32


# ---------------------------------------------------------------------------- batch shapes
def test_short_last_batch_is_skipped_and_indivisible_batch_is_an_error(tmp_path):
    tmp = str(tmp_path)
    runner, _ = _mlp_runner(tmp, 10, [])
    batches = _mlp_loader(n=10, batch=4)            # 4, 4, 2: the stock loader without drop_last
    assert [len(b[1]) for b in batches] == [4, 4, 2]
    runner.train(batches)
    assert runner.iter == 2                          # the short batch was skipped
    assert "skipping a batch of 2 samples" in open(os.path.join(tmp, "run.log")).read()
    eng = runner.engine
    with pytest.raises(ValueError, match="batch size 2"):
        eng.train_step([batches[2][0]], batches[2][1])
    runner2, _ = _mlp_runner(tmp, 10, [])
    runner2.micro_batches = 3
    with pytest.raises(ValueError, match="cannot be split into 3"):
        runner2.train(batches)


# ---------------------------------------------------------------------------- graph safety
def test_torch_optim_adapter_is_never_graph_captured():
    from skycomputing_b200.parallel.optim import TorchOptimizerAdapter

    lin = torch.nn.Linear(4, 4)
    opt = sky.build_optimizer(lin, dict(optim_type="Adam", lr=1e-3))
    assert isinstance(opt, TorchOptimizerAdapter) and opt.graph_safe is False


# ---------------------------------------------------------------------------- fused negotiation
def test_fused_boundary_support_evaluates_span_supports():
    from skycomputing_b200.models import BertConfig

    def stage(act):
        c = BertConfig(100, hidden_size=64, num_hidden_layers=1, num_attention_heads=1,
                       intermediate_size=128, max_position_embeddings=128, hidden_act=act)
        mc = [dict(layer_type="BertLayer_Head", config=c.__dict__),
              dict(layer_type="BertLayer_Body", config=c.__dict__),
              dict(layer_type="BertLayer_Tail", config=c.__dict__)]
        return sky.build_module_from_cfg(0, mc, dict(module_to_cuda=False))

    ok = stage("gelu")
    assert ok.fused_boundary_support() == (True, True)
    assert ok.fused_boundary_support((4, 128, 64)) == (True, True)
    assert ok.fused_boundary_support((4, 0, 0)) == (True, True)       # shape unknown: structural
    relu = stage("relu")
    assert relu.fused_boundary_support() == (True, True)              # structurally a whole block
    assert relu.fused_boundary_support((4, 128, 64)) == (False, False)


def test_fused_boundary_support_covers_the_cut_after_head():
    """Layer granularity (the reference's): a stage may end with BertLayer_Head alone (the K4
    GEMM + LayerNorm writes the peer slot) and the next one start with BertLayer_Body (FFN1 is the
    flag-gated consumer); a cut after BertLayer_Body (two tensors, 5x the bytes) is not fused."""
    from skycomputing_b200.models import BertConfig

    c = BertConfig(100, hidden_size=64, num_hidden_layers=1, num_attention_heads=1,
                   intermediate_size=128, max_position_embeddings=128)
    H, B, T = (dict(layer_type=f"BertLayer_{n}", config=c.__dict__) for n in ("Head", "Body", "Tail"))

    def support(layers):
        st = sky.build_module_from_cfg(0, layers, dict(module_to_cuda=False))
        return st.fused_boundary_support((4, 128, 64))

    assert support([H, B, T, H]) == (True, True)          # ... | Head   ends after Head
    assert support([B, T, H, B, T]) == (True, True)       # Body ...     starts after Head
    assert support([H, B]) == (True, False)               # ends after Body: not fused
    assert support([T, H, B, T]) == (False, True)         # starts after Body: not fused


def test_bench_plan_picker_matches_the_validated_plans():
    """bench.py's schedule choice (SchedulePlanner): the plans that were validated on 1 / 2 / 4 / 8
    GPUs, flag overrides, and slack for the allocator when a device is declared slow."""
    import argparse
    import importlib.util
    import pathlib

    spec = importlib.util.spec_from_file_location(
        "sky_bench", pathlib.Path(__file__).resolve().parents[1] / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    def plan(**kw):
        a = dict(gpus=1, layers=24, micro_batch=0, virtual_stages=0, slow_rank=-1, slowdown=0.0,
                 alloc="optimal")
        a.update(kw)
        return bench.pick_plan(argparse.Namespace(**a))

    assert plan(gpus=1)["schedule"] == "sequential" and plan(gpus=1)["virtual_stages"] == 1
    for n, v, m in ((2, 6, 2), (4, 6, 4), (8, 3, 8)):
        p = plan(gpus=n)
        assert (p["schedule"], p["virtual_stages"], p["micro_batches"], p["micro_batch"]) == \
            ("looped", v, m, 32), (n, p)
    assert plan(gpus=8, layers=160)["virtual_stages"] == 10
    assert plan(gpus=4, virtual_stages=1)["schedule"] == "1f1b"
    assert plan(gpus=4, micro_batch=16)["micro_batch"] == 16
    slow = plan(gpus=4, slow_rank=1, slowdown=1.0)
    assert slow["schedule"] == "looped" and slow["virtual_stages"] <= 2     # room to shed blocks
    assert plan(gpus=4, slow_rank=1, slowdown=1.0, alloc="even")["virtual_stages"] == 6
