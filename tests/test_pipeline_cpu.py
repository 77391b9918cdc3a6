"""Plumbing integration on CPU/gloo (BASELINE config 1 family): partitioned training equals
single-stage training, 1F1B equals sequential, hooks fire in order, checkpoints are
partition-independent, the stop flag is honoured, the launcher runs end to end."""
import os
import subprocess
import sys

import pytest
import torch

import skycomputing_b200 as sky
from tests._dist_helpers import run_distributed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model_cfg(layers=2, p=0.0):
    c = sky.BertConfig(100, hidden_size=32, num_hidden_layers=layers, num_attention_heads=4,
                       intermediate_size=64, max_position_embeddings=32, hidden_dropout_prob=p,
                       attention_probs_dropout_prob=p)
    enc = [dict(layer_type="BertLayer_Head", config=c.__dict__),
           dict(layer_type="BertLayer_Body", config=c.__dict__),
           dict(layer_type="BertLayer_Tail", config=c.__dict__)] * layers
    return ([dict(layer_type="BertEmbeddings", config=c.__dict__)] + enc
            + [dict(layer_type="BertPooler", config=c.__dict__),
               dict(layer_type="BertTailForClassification", hidden_dropout_prob=p, hidden_size=32,
                    num_classes=3)])


def _seed_layers(stage, first_layer_index):
    for off, layer in enumerate(stage.layers.children()):
        g = torch.Generator().manual_seed(100 + first_layer_index + off)
        with torch.no_grad():
            for p in layer.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)


def _train(rank, world, micro_batches, schedule, granularity, steps, tmp, ckpt_from, save_to):
    cfg = _model_cfg()
    wm = sky.WorkerManager(first_rank=0)
    wm.load_worker_pool_from_config([
        dict(name=f"w{i}", server_config={}, device=i,
             extra_config=dict(slowdown=0, mem_limit=-1, timer_config=dict(root=tmp)))
        for i in range(world)])
    wm = sky.Allocator(cfg, wm, granularity=granularity).even_allocate()
    model = sky.RpcModel(wm, this_rank=rank)
    _seed_layers(model.local_stage, model.local_module.layer_range[0])
    opt = sky.build_optimizer(model.local_stage, dict(optim_type="SGD", lr=0.1))
    ps = sky.ParameterServer(cfg, lazy=True) if rank == 0 else None
    runner = sky.Runner(model=model, parameter_server=ps, worker_manager=wm, optimizer=opt,
                        max_epochs=1, max_iters=steps, loss_cfg=dict(type="CrossEntropyLoss"),
                        timer_cfg=dict(root=tmp),
                        logging_cfg=dict(mode="a", filename=os.path.join(tmp, f"alloc{world}.log")),
                        micro_batches=micro_batches, schedule=schedule)
    order = []

    class Spy(sky.Hook):
        def before_run(self, r): order.append("before_run")
        def before_epoch(self, r): order.append("before_epoch")
        def before_iter(self, r): order.append("before_iter")
        def after_iter(self, r): order.append("after_iter")
        def after_epoch(self, r): order.append("after_epoch")
        def after_run(self, r): order.append("after_run")

    runner.register_hook(Spy())
    if ckpt_from or save_to:
        runner.register_hook(sky.CheckpointHook(load_checkpoint_from=ckpt_from,
                                                save_path=save_to, save_interval=1))
    dl = sky.build_dataloader_from_cfg(
        dataset_cfg=dict(type="SynthMNLIDataset", num_samples=8 * steps, max_seq_length=16,
                         vocab_size=100, seed=5),
        dataloader_cfg=dict(batch_size=8, shuffle=False))
    losses = []
    orig = runner.train_iteration

    def spy_iter(data, labels):
        out = orig(data, labels)
        if out is not None:
            losses.append(out)
        return out

    runner.train_iteration = spy_iter
    runner.train(dl)
    n_layers = [len(w.model_config) for w in wm.worker_pool]
    return dict(losses=losses, order=order, layers=n_layers, iters=runner.iter)


def test_partitioned_training_matches_single_stage(tmp_path):
    tmp = str(tmp_path)
    single = run_distributed(_train, 1, 1, "sequential", "layer", 4, tmp, None, None)[0]
    two = run_distributed(_train, 2, 1, "sequential", "layer", 4, tmp, None, None)
    three = run_distributed(_train, 3, 1, "sequential", "block", 4, tmp, None, None)
    l1 = single["losses"]
    l2 = [r["losses"] for r in two if r["losses"]][0]
    l3 = [r["losses"] for r in three if r["losses"]][0]
    assert len(l1) == 4 and l1[-1] < l1[0]
    assert l2 == pytest.approx(l1, rel=1e-5) and l3 == pytest.approx(l1, rel=1e-5)
    assert two[0]["layers"] == [5, 4]                     # cut inside a block is legal (layer gran.)
    assert single["order"] == (["before_run", "before_epoch"] + ["before_iter", "after_iter"] * 4
                               + ["after_epoch", "after_run"])
    assert single["iters"] == 4                           # max_iters is exact (no off-by-one)


def test_one_f_one_b_equals_sequential_microbatching(tmp_path):
    tmp = str(tmp_path)
    seq = run_distributed(_train, 2, 4, "sequential", "block", 3, tmp, None, None)
    f1b = run_distributed(_train, 2, 4, "1f1b", "block", 3, tmp, None, None)
    one = run_distributed(_train, 1, 4, "1f1b", "block", 3, tmp, None, None)
    ls = [r["losses"] for r in seq if r["losses"]][0]
    lf = [r["losses"] for r in f1b if r["losses"]][0]
    assert lf == pytest.approx(ls, rel=1e-5)
    assert one[0]["losses"] == pytest.approx(ls, rel=1e-5)


def test_checkpoint_is_partition_independent(tmp_path):
    tmp = str(tmp_path)
    save_dir = os.path.join(tmp, "ckpt")
    a = run_distributed(_train, 2, 1, "sequential", "block", 3, tmp, None, save_dir)
    ckpt = os.path.join(save_dir, "epoch_0.pth")   # saved in after_epoch of epoch 0
    files = os.listdir(save_dir)
    assert any(f.startswith("epoch_") and f.endswith(".pth") for f in files), files
    ckpt = os.path.join(save_dir, sorted(f for f in files if ".extra." not in f)[0])
    sd = torch.load(ckpt)
    assert any(k.startswith("0.word_embeddings") for k in sd) and any(k.startswith("8.") for k in sd)
    # resume the same weights under a DIFFERENT partition (3 stages) and on a single stage:
    b = run_distributed(_train, 3, 1, "sequential", "block", 2, tmp, ckpt, None)
    c = run_distributed(_train, 1, 1, "sequential", "block", 2, tmp, ckpt, None)
    lb = [r["losses"] for r in b if r["losses"]][0]
    lc = c[0]["losses"]
    assert lb == pytest.approx(lc, rel=1e-5)
    la = [r["losses"] for r in a if r["losses"]][0]
    assert abs(lc[0] - la[0]) > 1e-6                      # weights differ from the seeded init


def _stop_run(rank, world, tmp):
    cfg = _model_cfg(1)
    wm = sky.WorkerManager(first_rank=0)
    wm.load_worker_pool_from_config([dict(name=f"w{i}", server_config={}, device=i,
                                          extra_config=dict(timer_config=dict(root=tmp)))
                                     for i in range(world)])
    wm = sky.Allocator(cfg, wm).even_allocate()
    model = sky.RpcModel(wm, this_rank=rank)
    opt = sky.build_optimizer(model.local_stage, dict(optim_type="Adam", lr=1e-3))  # any torch.optim name
    runner = sky.Runner(model=model, parameter_server=None, worker_manager=wm, optimizer=opt,
                        max_epochs=5, max_iters=1000, loss_cfg=dict(type="CrossEntropyLoss"),
                        timer_cfg=dict(root=tmp), logging_cfg=None)
    runner.register_hook(sky.StopHook(root=tmp))
    runner.register_hook(sky.DistributedTimerHelperHook())

    class Trigger(sky.Hook):
        def after_iter(self, r):
            if r.iter == 2 and rank == 0:
                sky.StopHook.stop(tmp)

    runner.register_hook(Trigger())
    dl = sky.build_dataloader_from_cfg(
        dataset_cfg=dict(type="SynthMNLIDataset", num_samples=64, max_seq_length=16, vocab_size=100),
        dataloader_cfg=dict(batch_size=4))
    runner.train(dl)
    return runner.iter


def test_stop_flag_is_honoured_on_all_ranks(tmp_path):
    iters = run_distributed(_stop_run, 2, str(tmp_path))
    assert all(i > 1000 for i in iters)                    # pushed past max_iters on every rank
    assert not os.path.exists(os.path.join(str(tmp_path), "stop_flag.txt"))


def test_launcher_end_to_end_config1(tmp_path):
    """BASELINE config 1: 4-layer BERT, CORE_NUM=2, even, CPU/gloo, through the CLI."""
    env = dict(os.environ, TINY="1", LAYER_NUM="4", CORE_NUM="2", DEVICE="cpu", MAX_ITERS="3",
               PROJECT=str(tmp_path), ALLOCATE_TYPE="even")
    from tests._dist_helpers import free_port

    out = subprocess.run([sys.executable, "-m", "skycomputing_b200.launch", "-c",
                          os.path.join(ROOT, "experiment", "config.py"), "--spawn", "1", "-p",
                          str(free_port())], cwd=ROOT, env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    log = open(os.path.join(str(tmp_path), "logs", "2nodes_4layers", "even", "allocation.log")).read()
    assert "rank: 0, number of layers: 15" in log
    assert log.count("step time:") == 3 and "epoch: 0, iter: 2" in log


def test_launcher_dynamic_allocation_two_workers(tmp_path):
    env = dict(os.environ, TINY="1", LAYER_NUM="4", CORE_NUM="3", DEVICE="cpu", MAX_ITERS="2",
               PROJECT=str(tmp_path), ALLOCATE_TYPE="dynamic", MICRO_BATCHES="2", BATCH_SIZE="8")
    from tests._dist_helpers import free_port

    out = subprocess.run([sys.executable, "-m", "skycomputing_b200.launch", "-c",
                          os.path.join(ROOT, "experiment", "config.py"), "--spawn", "2", "-p",
                          str(free_port())], cwd=ROOT, env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    d = os.path.join(str(tmp_path), "logs", "3nodes_4layers", "dynamic")
    log = open(os.path.join(d, "allocation.log")).read()
    assert "dynamically allocated model layers" in log and log.count("number of layers") == 2
    assert os.path.exists(os.path.join(d, "metrics.jsonl"))


def test_launcher_reallocation_hook_from_config(tmp_path):
    """REALLOCATE_EVERY wires ReallocateHook through the launcher: the devices are re-benchmarked
    during training and the decision (keep / migrate) is logged on rank 0."""
    env = dict(os.environ, TINY="1", LAYER_NUM="4", CORE_NUM="3", DEVICE="cpu", MAX_ITERS="3",
               PROJECT=str(tmp_path), ALLOCATE_TYPE="dynamic", MICRO_BATCHES="2", BATCH_SIZE="8",
               REALLOCATE_EVERY="2")
    from tests._dist_helpers import free_port

    out = subprocess.run([sys.executable, "-m", "skycomputing_b200.launch", "-c",
                          os.path.join(ROOT, "experiment", "config.py"), "--spawn", "2", "-p",
                          str(free_port())], cwd=ROOT, env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    d = os.path.join(str(tmp_path), "logs", "3nodes_4layers", "dynamic")
    log = open(os.path.join(d, "allocation.log")).read()
    assert "reallocate:" in log, log


# ---------------------------------------------------------------------------------------------
# ReallocateHook: re-benchmark + re-allocate + migrate layers in the middle of a run
# ---------------------------------------------------------------------------------------------
class _FakeDeviceBench:
    """Device 1 is 3x slower than device 0 (keyed like DeviceBenchmarker: worker name by rank)."""

    def __init__(self, wm):
        self._wm = wm

    def benchmark(self):
        from skycomputing_b200.utils import generate_worker_name

        return {generate_worker_name(w.rank): dict(time=3.0 if w.device == 1 else 1.0, avai_mem=1e9)
                for w in self._wm.worker_pool}


class _FakeModelBench:
    def __init__(self, n):
        self._n = n

    def benchmark(self):
        return [1.0] * self._n, [1.0] * self._n


def _train_with_reallocation(rank, world, use_hook, steps, tmp):
    cfg = _model_cfg(layers=4)
    wm = sky.WorkerManager(first_rank=0)
    wm.load_worker_pool_from_config([
        dict(name=f"w{i}", server_config={}, device=i,
             extra_config=dict(slowdown=0, mem_limit=-1, timer_config=dict(root=tmp)))
        for i in range(world)])
    wm = sky.Allocator(cfg, wm, granularity="block").even_allocate()
    model = sky.RpcModel(wm, this_rank=rank)
    _seed_layers(model.local_stage, model.local_module.layer_range[0])
    optim_cfg = dict(optim_type="SGD", lr=0.1)
    opt = sky.build_optimizer(model.local_stage, dict(optim_cfg))
    runner = sky.Runner(model=model, parameter_server=None, worker_manager=wm, optimizer=opt,
                        max_epochs=1, max_iters=steps, loss_cfg=dict(type="CrossEntropyLoss"),
                        timer_cfg=dict(root=tmp), logging_cfg=None, micro_batches=2,
                        schedule="1f1b")
    hook = None
    if use_hook:
        def factory(pool):
            return sky.Allocator(cfg, pool, _FakeModelBench(len(cfg)), _FakeDeviceBench(pool),
                                 granularity="block", solver="exact")

        hook = sky.ReallocateHook(interval=2, allocator_factory=factory, optimizer_cfg=optim_cfg,
                                  allocate_type="dynamic", min_gain=0.05)
        runner.register_hook(hook)
    dl = sky.build_dataloader_from_cfg(
        dataset_cfg=dict(type="SynthMNLIDataset", num_samples=8 * steps, max_seq_length=16,
                         vocab_size=100, seed=5),
        dataloader_cfg=dict(batch_size=8, shuffle=False))
    losses = []
    orig = runner.train_iteration

    def spy_iter(data, labels):
        out = orig(data, labels)
        if out is not None:
            losses.append(out)
        return out

    runner.train_iteration = spy_iter
    runner.train(dl)
    sums = {}
    mod = runner.model.local_module
    b, _e = mod.layer_range
    for off, sd in enumerate(mod.get_state_dict()):
        sums[b + off] = float(sum(v.double().abs().sum() for v in sd.values()))
    return dict(losses=losses, sums=sums, range=tuple(mod.layer_range),
                migrations=0 if hook is None else hook.migrations,
                decision=None if hook is None else hook.last_decision)


def test_reallocate_hook_migrates_layers_without_changing_the_training_result(tmp_path):
    tmp = str(tmp_path)
    plain = run_distributed(_train_with_reallocation, 2, False, 5, tmp)
    moved = run_distributed(_train_with_reallocation, 2, True, 5, tmp)
    # even split of the 4 block units (emb rides with the first, pooler + classifier with the last
    # block), then the 3x slower device 1 sheds a block to device 0
    assert plain[0]["range"] != moved[0]["range"]
    assert moved[0]["migrations"] == 1 and moved[1]["migrations"] == 1   # iter 2 moves, iter 4 keeps
    n0 = moved[0]["range"][1] - moved[0]["range"][0]
    n1 = moved[1]["range"][1] - moved[1]["range"][0]
    assert n0 > n1
    assert moved[0]["decision"]["gain"] >= 0.0
    # training is partition independent: same losses, same final weights layer by layer
    lp = [r["losses"] for r in plain if r["losses"]][0]
    lm = [r["losses"] for r in moved if r["losses"]][0]
    assert lm == pytest.approx(lp, rel=1e-5)
    sp, sm = {}, {}
    for r in plain:
        sp.update(r["sums"])
    for r in moved:
        sm.update(r["sums"])
    assert sorted(sp) == sorted(sm) == list(range(15))
    for k in sp:
        assert sm[k] == pytest.approx(sp[k], rel=1e-6), k


# ---------------------------------------------------------------------------------------------
# Looped (virtual-stage) pipeline: v chunks per rank, ring of ranks
# ---------------------------------------------------------------------------------------------
def _train_looped(rank, world, virtual_stages, micro_batches, steps, tmp, alloc, save_to=None,
                  layers=6):
    cfg = _model_cfg(layers=layers)
    wm = sky.WorkerManager(first_rank=0)
    wm.load_worker_pool_from_config([
        dict(name=f"w{i}", server_config={}, device=i,
             extra_config=dict(slowdown=0, mem_limit=-1, timer_config=dict(root=tmp)))
        for i in range(world)])
    if alloc == "even":
        allocator = sky.Allocator(cfg, wm, granularity="block")
    else:
        allocator = sky.Allocator(cfg, wm, _FakeModelBench(len(cfg)), _FakeDeviceBench(wm),
                                  granularity="block", solver="exact")
    wm = allocator.allocate(alloc, virtual_stages=virtual_stages)
    model = sky.RpcModel(wm, this_rank=rank)
    for mod in model.model:
        if mod.is_local:
            _seed_layers(mod.module, mod.layer_range[0])
    opt = sky.build_optimizer(model.optim_module, dict(optim_type="SGD", lr=0.1))
    ps = sky.ParameterServer(cfg, lazy=True) if (rank == 0 and save_to) else None
    runner = sky.Runner(model=model, parameter_server=ps, worker_manager=wm, optimizer=opt,
                        max_epochs=1, max_iters=steps, loss_cfg=dict(type="CrossEntropyLoss"),
                        timer_cfg=dict(root=tmp), logging_cfg=None, micro_batches=micro_batches,
                        schedule="1f1b" if micro_batches > 1 else "sequential")
    if save_to:
        runner.register_hook(sky.CheckpointHook(save_path=save_to, save_interval=1,
                                                save_optimizer=False))
    dl = sky.build_dataloader_from_cfg(
        dataset_cfg=dict(type="SynthMNLIDataset", num_samples=8 * steps, max_seq_length=16,
                         vocab_size=100, seed=5),
        dataloader_cfg=dict(batch_size=8, shuffle=False))
    losses = []
    orig = runner.train_iteration

    def spy_iter(data, labels):
        out = orig(data, labels)
        if out is not None:
            losses.append(out)
        return out

    runner.train_iteration = spy_iter
    runner.train(dl)
    sums = {}
    for mod in runner.model.model:
        if mod.is_local:
            b, _e = mod.layer_range
            for off, sd in enumerate(mod.get_state_dict()):
                sums[b + off] = float(sum(v.double().abs().sum() for v in sd.values()))
    return dict(losses=losses, sums=sums, schedule=runner.engine.schedule,
                chunks=[list(w.chunks) if w.chunks else None for w in wm.worker_pool])


def test_looped_pipeline_equals_plain_pipeline(tmp_path):
    """v chunks per rank on a ring of ranks: same losses and the same final weights as the plain
    one-span-per-rank pipeline (which equals single-stage training, tested above)."""
    tmp = str(tmp_path)
    plain = run_distributed(_train_looped, 2, 1, 2, 4, tmp, "even")
    loop2 = run_distributed(_train_looped, 2, 2, 2, 4, tmp, "even")
    loop3 = run_distributed(_train_looped, 3, 2, 4, 4, tmp, "even")
    assert plain[0]["schedule"] != "looped" and loop2[0]["schedule"] == "looped"
    assert loop2[0]["chunks"] == [[(0, 7), (13, 16)], [(7, 13), (16, 21)]]
    lp = [r["losses"] for r in plain if r["losses"]][0]
    l2 = [r["losses"] for r in loop2 if r["losses"]][0]
    l3 = [r["losses"] for r in loop3 if r["losses"]][0]
    assert len(lp) == 4 and l2 == pytest.approx(lp, rel=1e-5)
    # 4 micro-batches of 2 vs 2 micro-batches of 4: the same mean loss per step
    assert l3 == pytest.approx(lp, rel=1e-4)
    sp, s2 = {}, {}
    for r in plain:
        sp.update(r["sums"])
    for r in loop2:
        s2.update(r["sums"])
    assert sorted(sp) == sorted(s2) == list(range(21))
    for k in sp:
        assert s2[k] == pytest.approx(sp[k], rel=1e-6), k


def test_looped_allocation_respects_device_speeds(tmp_path):
    """Exact solver over v x D virtual devices: the 3x slower device gets the lighter chunks."""
    out = run_distributed(_train_looped, 2, 2, 2, 2, str(tmp_path), "optimal")
    chunks = out[0]["chunks"]
    n0 = sum(e - b for b, e in chunks[0])
    n1 = sum(e - b for b, e in chunks[1])
    assert n0 > n1 and n0 + n1 == 21
    assert [r["losses"] for r in out if r["losses"]][0][-1] > 0


def test_launcher_looped_pipeline_from_config(tmp_path):
    env = dict(os.environ, TINY="1", LAYER_NUM="4", CORE_NUM="3", DEVICE="cpu", MAX_ITERS="2",
               PROJECT=str(tmp_path), ALLOCATE_TYPE="even", MICRO_BATCHES="2", BATCH_SIZE="8",
               VIRTUAL_STAGES="2")
    from tests._dist_helpers import free_port

    out = subprocess.run([sys.executable, "-m", "skycomputing_b200.launch", "-c",
                          os.path.join(ROOT, "experiment", "config.py"), "--spawn", "2", "-p",
                          str(free_port())], cwd=ROOT, env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    d = os.path.join(str(tmp_path), "logs", "3nodes_4layers", "even")
    log = open(os.path.join(d, "allocation.log")).read()
    assert "runs layer spans [(0, 4), (7, 10)]" in log and log.count("step time") == 2


def test_checkpoint_of_a_looped_pipeline_equals_plain_checkpoint(tmp_path):
    """Layer-indexed checkpoints do not care that a rank owns several non-adjacent chunks."""
    tmp = str(tmp_path)
    a_dir, b_dir = os.path.join(tmp, "looped"), os.path.join(tmp, "plain")
    run_distributed(_train_looped, 2, 2, 2, 3, tmp, "even", a_dir)
    run_distributed(_train_looped, 2, 1, 2, 3, tmp, "even", b_dir)
    fa = [f for f in os.listdir(a_dir) if f.endswith(".pth") and ".extra." not in f]
    fb = [f for f in os.listdir(b_dir) if f.endswith(".pth") and ".extra." not in f]
    assert fa and fa == fb
    sa, sb = torch.load(os.path.join(a_dir, fa[0])), torch.load(os.path.join(b_dir, fb[0]))
    assert sorted(sa) == sorted(sb) and any(k.startswith("14.") for k in sa)
    for k in sa:
        assert torch.allclose(sa[k].float(), sb[k].float(), rtol=1e-5, atol=1e-7), k
