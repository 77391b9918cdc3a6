"""CPU unit tests of the L0/L2/L4 plumbing: config, registry, builders, workers, logger, timers,
stimulator, estimator/benchmarkers (goldens derived from the reference)."""
import os
import textwrap

import numpy as np
import pytest
import torch

import skycomputing_b200 as sky
from skycomputing_b200 import _core


# ---------------------------------------------------------------- config
def test_config_attribute_access_and_base(tmp_path):
    base = tmp_path / "base.py"
    base.write_text(textwrap.dedent("""
        import os
        a = 1
        nested = dict(x=1)
        def helper():
            return 3
        _private = 5
    """))
    child = tmp_path / "child.py"
    child.write_text("base = 'base.py'\na = 2\nb = [1, 2]\n")
    cfg = sky.load_config(str(child))
    assert cfg.a == 2 and cfg["b"] == [1, 2] and cfg.nested == {"x": 1}
    assert "helper" not in cfg and "_private" not in cfg and "os" not in cfg and "base" not in cfg
    with pytest.raises(KeyError):
        cfg["missing"]
    cfg.z = 9
    assert cfg["z"] == 9
    assert sky.Config.from_dict({"k": 1}).k == 1


def test_example_config_loads():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["LAYER_NUM"] = "4"
    os.environ["CORE_NUM"] = "3"
    try:
        cfg = sky.load_config(os.path.join(root, "experiment", "config.py"))
    finally:
        os.environ.pop("LAYER_NUM"), os.environ.pop("CORE_NUM")
    assert len(cfg.model_config) == 3 * 4 + 3
    assert len(cfg.worker_config) == 2
    for key in ("model_config", "rpc_config", "data_config", "logging_config", "worker_config",
                "allocator_config", "train_config"):
        assert key in cfg


# ---------------------------------------------------------------- registry / builders
def test_registry_semantics():
    reg = sky.Registry("t")

    @reg.register_module
    class Foo:
        pass

    assert Foo is not None and reg.get_module("Foo") is Foo  # decorator returns the class
    with pytest.raises(AssertionError):
        reg.register_module(Foo)
    assert reg.get_module("Conv2d") is torch.nn.Conv2d        # torch.nn fallback
    with pytest.raises(NameError):
        reg.get_module("NoSuchLayer")
    with pytest.raises(NameError):
        reg.get_module("Conv2d", include_torch=False)


def test_registered_names_match_reference_surface():
    for n in ["BertEmbeddings", "BertLayer_Head", "BertLayer_Body", "BertLayer_Tail", "BertPooler",
              "BertTailForClassification", "BasicBlock", "BottleNeck", "ResLayer", "ResTail",
              "ResHead", "ResNet"]:
        assert n in sky.LAYER
    for n in ["GlueDataset", "RandomMlpDataset", "CIFAR10Dataset", "SynthMNLIDataset"]:
        assert n in sky.DATASET
    for n in ["CheckpointHook", "StopHook", "DistributedTimerHelperHook"]:
        assert n in sky.HOOKS
    for n in ["RandomTensorGenerator", "DataloaderGenerator"]:
        assert n in sky.DATA_GENERATOR
    assert sky.__version__


def test_sequential_wrapper_splats_tuples():
    class A(torch.nn.Module):
        def forward(self, x):
            return x, x + 1

    class B(torch.nn.Module):
        def forward(self, x, y):
            return x * y

    out = sky.SequentialWrapper(A(), B())(torch.tensor(2.0))
    assert out.item() == 6.0


def test_build_module_from_cfg_and_module_wrapper_cpu(tmp_path):
    cfg = [dict(layer_type="Linear", in_features=8, out_features=8)] * 2
    extra = dict(slowdown=1, mem_limit=123, timer_config=dict(root=str(tmp_path)),
                 logging_config=dict(mode="a", filename=str(tmp_path / "node.log")))
    mw = sky.build_module_from_cfg(3, cfg, extra)
    assert "record_forward_time" not in extra           # argument dict is not mutated
    x = torch.randn(4, 8, requires_grad=True)
    (y,) = mw(x)
    y.sum().backward()
    assert x.grad is not None
    assert len(mw.forward_time) == 1 and mw.forward_time[0] > 0
    assert len(mw.backward_time) == 1
    assert mw.detect_mem() == 123
    mw.flush_logs()
    text = (tmp_path / "node.log").read_text()
    assert "forward time on rank 3" in text and "backward time on rank 3" in text
    mw2 = sky.build_module_from_cfg(0, cfg, dict(mem_limit=-1))
    assert mw2.detect_mem() > 0                          # psutil path


def test_dataloader_and_generators():
    dl = sky.build_dataloader_from_cfg(
        dataset_cfg=dict(type="SynthMNLIDataset", num_samples=10, max_seq_length=16, vocab_size=100),
        dataloader_cfg=dict(batch_size=5))
    (ids, tt, m), y = next(iter(dl))
    assert ids.shape == (5, 16) and tt.shape == (5, 16) and m.shape == (5, 16) and y.shape == (5,)
    assert ids.dtype == torch.long and int(y.max()) < 3
    g = sky.build_data_generator("RandomTensorGenerator", generator_cfg=dict(size=(2, 3)))
    assert g.generate().shape == (2, 3)
    g2 = sky.build_data_generator("DataloaderGenerator", generator_cfg=dict(
        dataset_cfg=dict(type="SynthMNLIDataset", num_samples=4, max_seq_length=8, vocab_size=50),
        dataloader_cfg=dict(batch_size=2)))
    assert len(g2.generate()) == 3
    ref = sky.SynthMNLIDataset(4, 8, 50, reference_order=True)
    nat = sky.SynthMNLIDataset(4, 8, 50, reference_order=False)
    assert torch.equal(ref[0][0][1], nat[0][0][2])       # mask position differs by convention


# ---------------------------------------------------------------- workers
def test_worker_manager_ranks_and_roundtrip():
    wm = sky.WorkerManager()
    wm.load_worker_pool_from_config([dict(name=f"w{i}", server_config={}, extra_config={}) for i in range(3)])
    assert [w.rank for w in wm.worker_pool] == [1, 2, 3]  # rank 0 reserved (reference numbering)
    wm.worker_pool[0].order, wm.worker_pool[1].order, wm.worker_pool[2].order = 3, 1, 2
    wm.reset_rank_by_order()
    assert [w.name for w in wm.worker_pool] == ["w1", "w2", "w0"]
    assert [w.rank for w in wm.worker_pool] == [1, 2, 3]
    wid = wm.worker_pool[1].id
    assert wm.get_by_id(wid).name == "w2"
    assert wm.get_by_id("nope", allow_not_found=True) is None
    with pytest.raises(LookupError):
        wm.get_by_id("nope")
    wm.remove_worker_by_id(wid)
    assert wm.size == 2 and [w.rank for w in wm.worker_pool] == [1, 2]
    wm.add_worker(None, dict(name="new", server_config={}, extra_config={}))
    assert wm.worker_pool[-1].rank == 3
    clone = sky.WorkerManager.deserialize(wm.serialize())
    assert [w.name for w in clone.worker_pool] == [w.name for w in wm.worker_pool]
    spmd = sky.WorkerManager(first_rank=0)
    spmd.load_worker_pool_from_config([dict(name="a", server_config={}, extra_config={})] * 2)
    assert [w.rank for w in spmd.worker_pool] == [0, 1] and [w.device for w in spmd.worker_pool] == [0, 1]


# ---------------------------------------------------------------- logger / timers / stimulator
def test_logger_and_distributed_timer(tmp_path):
    lg = sky.Logger(str(tmp_path / "sub" / "a.log"), mode="a")
    lg.info("hello")
    line = (tmp_path / "sub" / "a.log").read_text()
    assert line.startswith("INFO - ") and line.rstrip().endswith(" - hello")
    t = sky.DistributedTimer(root=str(tmp_path))
    t.add_timestamp()
    t.add_timestamp()
    assert t.get_prev_interval() >= 0
    assert (tmp_path / "dist_timer.txt").read_text().startswith("timestamp: ")
    t.clean_prev_file()
    assert not (tmp_path / "dist_timer.txt").exists()


def test_stimulator_matches_numpy_reference_streams():
    s = sky.Stimulator(8)
    assert np.array_equal(s.m_slowdown, 2 * np.random.default_rng(seed=22).random((9,)) + 1)
    assert np.array_equal(s.c_slowdown, np.random.default_rng(seed=32).random((9,)) + 1)
    assert np.array_equal(s.n_slowdown, s.c_slowdown)    # the reference shares seed 32
    assert 1 <= s.memory_slowdown(3) < 3 and 1 <= s.compute_slowdown(3) < 2
    for seed in (0, 1, 35, 2 ** 40 + 7):
        assert np.array_equal(np.array(_core.numpy_default_rng_random(seed, 5)),
                              np.random.default_rng(seed=seed).random(5))


# ---------------------------------------------------------------- estimator / benchmarkers
def _tiny_bert_cfg(layers=2):
    c = sky.BertConfig(100, hidden_size=32, num_hidden_layers=layers, num_attention_heads=4,
                       intermediate_size=64, max_position_embeddings=32)
    enc = [dict(layer_type="BertLayer_Head", config=c.__dict__),
           dict(layer_type="BertLayer_Body", config=c.__dict__),
           dict(layer_type="BertLayer_Tail", config=c.__dict__)] * layers
    return ([dict(layer_type="BertEmbeddings", config=c.__dict__)] + enc
            + [dict(layer_type="BertPooler", config=c.__dict__),
               dict(layer_type="BertTailForClassification", hidden_dropout_prob=0.1, hidden_size=32,
                    num_classes=3)])


def test_model_benchmarker_generic_dedup():
    cfg = _tiny_bert_cfg(3)
    gen = sky.build_data_generator("DataloaderGenerator", generator_cfg=dict(
        dataset_cfg=dict(type="SynthMNLIDataset", num_samples=4, max_seq_length=16, vocab_size=100),
        dataloader_cfg=dict(batch_size=4)))
    flops, mem = sky.ModelBenchmarker(cfg, gen, device="cpu").benchmark()
    assert len(flops) == len(cfg) == len(mem)
    assert flops[1] == flops[4] == flops[7] and flops[2] == flops[5] and flops[3] == flops[6]
    B, S, H, I = 4, 16, 32, 64
    assert flops[2] == 2 * B * S * H * I                 # FFN1 GEMM
    assert flops[3] == 2 * B * S * H * I                 # FFN2 GEMM
    assert flops[1] == 4 * 2 * B * S * H * H + 2 * 2 * B * S * S * H   # QKV+out proj + QK^T + PV
    assert all(m > 0 for m in mem)


def test_analytic_model_benchmark_matches_measured():
    cfg = _tiny_bert_cfg(2)
    gen = sky.build_data_generator("DataloaderGenerator", generator_cfg=dict(
        dataset_cfg=dict(type="SynthMNLIDataset", num_samples=4, max_seq_length=16, vocab_size=100),
        dataloader_cfg=dict(batch_size=4)))
    f_meas, m_meas = sky.ModelBenchmarker(cfg, gen, device="cpu").benchmark()
    f_ana, m_ana = sky.ModelBenchmarker(cfg, gen, device="cpu", analytic=True).benchmark()
    assert f_ana == pytest.approx(f_meas, rel=1e-6, abs=1e-6)
    assert m_ana == pytest.approx(m_meas, rel=0.05)


def test_device_benchmarker_single_process_with_slowdown_and_stimulate(monkeypatch):
    wm = sky.WorkerManager()
    wm.load_worker_pool_from_config([
        dict(name="fast", server_config={}, extra_config=dict(slowdown=0, mem_limit=1000)),
        dict(name="slow", server_config={}, extra_config=dict(slowdown=3, mem_limit=500))])
    # big enough (and warmed up) that host timing noise cannot hide the 4x throttle
    gen = sky.build_data_generator("RandomTensorGenerator", generator_cfg=dict(size=(256, 512)))
    db = sky.DeviceBenchmarker(wm, gen, model_config=[dict(layer_type="Linear", in_features=512,
                                                           out_features=512)] * 6, iterations=10,
                               warmup=3)
    res = db.benchmark()
    assert list(res) == ["worker1", "worker2"]
    assert res["worker2"]["time"] > 2.0 * res["worker1"]["time"]
    assert res["worker1"]["avai_mem"] == 1000 and res["worker2"]["avai_mem"] == 500
    monkeypatch.setenv("STIMULATE", "1")
    db2 = sky.DeviceBenchmarker(wm, gen, model_config=[dict(layer_type="Linear", in_features=512,
                                                            out_features=512)], iterations=2)
    res2 = db2.benchmark()
    s = sky.Stimulator(2)
    assert res2["worker1"]["avai_mem"] == pytest.approx(1000 / s.memory_slowdown(1))


def test_parameter_server_layer_indexed_roundtrip(tmp_path):
    cfg = _tiny_bert_cfg(1)
    ps = sky.ParameterServer(cfg)
    sd0 = {k: v.clone() + 1 for k, v in ps.get_state_dict(2).items()}
    ps.update_weights(sd0, 2)
    f = str(tmp_path / "ckpt.pth")
    ps.save_weights_to_file(f)
    ps2 = sky.ParameterServer(cfg)
    ps2.load_weights_from_file(f)
    for k, v in ps2.get_state_dict(2).items():
        assert torch.equal(v, sd0[k])
    lazy = sky.ParameterServer(cfg, lazy=True)
    lazy.load_weights_from_file(f)
    assert set(lazy.full_state_dict()) == set(ps.module_list.state_dict())


def test_reference_named_helpers_exist_and_work(tmp_path):
    """BackwardSlowdown{Function,Module}, call_method /
    remote_method / parameter_rrefs, bias_gelu_training, glue.file_utils."""
    import torch

    import skycomputing_b200 as sky
    from skycomputing_b200.builder import BackwardSlowdownFunction, BackwardSlowdownModule
    from skycomputing_b200.dataset.glue import file_utils as fu
    from skycomputing_b200.models.bert_layers import bias_gelu_training
    from skycomputing_b200.utils import OwnerRef, call_method, parameter_rrefs, remote_method

    # backward probe: identity, one gradient per input, throttles through the shared timer file
    timer = sky.DistributedTimer(root=str(tmp_path))
    x = torch.ones(4, requires_grad=True)
    y = BackwardSlowdownModule(0, 0.0, timer, None, True)(x)
    timer.add_timestamp()
    (y * 2).sum().backward()
    assert torch.equal(x.grad, torch.full((4,), 2.0))
    assert BackwardSlowdownFunction.apply(x, 0, 0.0, None, None, False).shape == x.shape

    lin = torch.nn.Linear(3, 2)
    refs = parameter_rrefs(lin)
    assert len(refs) == 2 and all(isinstance(r, OwnerRef) and r.is_owner() for r in refs)
    assert remote_method(lambda p: tuple(p.shape), refs[0]) == (2, 3)
    assert call_method(lambda p, k: p.numel() * k, refs[1], 3) == 6
    assert refs[0].to_here() is refs[0].local_value()

    out = bias_gelu_training(torch.zeros(3), torch.tensor([-1.0, 0.0, 1.0]))
    assert torch.allclose(out, torch.nn.functional.gelu(torch.tensor([-1.0, 0.0, 1.0])))

    # file cache: local path wins, cached URL copy wins, offline fetch fails with a usable message
    p = tmp_path / "words.txt"
    p.write_text("a\nb \n")
    assert fu.cached_path(str(p)) == str(p)
    assert fu.read_set_from_file(str(p)) == {"a", "b"}
    url = "https://example.com/bert/vocab.txt"
    cached = tmp_path / fu.url_to_filename(url)
    cached.write_text("cached")
    assert fu.cached_path(url, cache_dir=str(tmp_path)) == str(cached)
    assert fu.get_file_extension("X/Y.TSV") == ".tsv" and fu.get_file_extension("a.b", dot=False) == "b"
    assert fu.split_s3_path("s3://bucket/some/key") == ("bucket", "some/key")
    assert fu.url_to_filename(url, "etag") != fu.url_to_filename(url)
    import pytest
    with pytest.raises(FileNotFoundError):
        fu.cached_path(str(tmp_path / "missing.bin"))


def test_trace_tools_on_committed_timelines():
    """tools/render_trace.py and tools/predict_schedule.py run on the committed 8-GPU traces and the
    step-time model stays within 3 % of the measured span."""
    import re
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "profiles", "trace_n8_mb32")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "render_trace.py"), "--dir", d,
                          "--n", "8", "--res", "400"], capture_output=True, text=True, cwd=root)
    assert out.returncode == 0, out.stderr
    assert out.stdout.count("\ns") >= 7 and "F " in out.stdout
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "predict_schedule.py"), "--dir", d],
                         capture_output=True, text=True, cwd=root)
    assert out.returncode == 0, out.stderr
    measured = float(re.search(r"measured step span\s+([0-9.]+)", out.stdout).group(1))
    model = float(re.search(r"plain 1F1B model\s+([0-9.]+)", out.stdout).group(1))
    assert abs(model - measured) / measured < 0.03
