"""Reference arm of bench.py: runs the UNMODIFIED hpcaitech/SkyComputing (``scaelum``) package
installed under baseline/_ref through its own public API and stock code path - TensorPipe RPC,
``module_to_cuda=True, output_to_cpu=True`` CPU staging, dist_autograd, DistributedOptimizer(SGD),
``Runner.train`` - on synthetic MNLI-shaped data and random-init BERT-large weights.

Nothing from skycomputing_b200 is imported here.  One process per GPU (torchrun for N > 1): rank 0
is the reference's master AND hosts the first stage as a ``LocalModule`` (worker rank 0, the
reference's own mechanism, scaelum/model/rpc_model.py:28-33), ranks 1..N-1 serve their stages as
``RemoteModule`` targets.  Timing: a ``scaelum.runner.Hook`` records, around every iteration,
``torch.cuda.synchronize()`` + wall clock on rank 0 (the reference's loop is fully synchronous:
the iteration returns only after the last RPC of ``optimizer.step`` completed on every stage, so
rank-0 time IS the max over ranks) and a CUDA-event pair on rank 0's device.
"""
from __future__ import annotations

import json
import os
import os.path as osp
import shutil
import sys
import time

HERE = osp.dirname(osp.abspath(__file__))


def _paths():
    ref = osp.join(HERE, "_ref")
    stubs = osp.join(HERE, "stubs")
    if not osp.isdir(osp.join(ref, "scaelum")):
        return None
    for p in (ref, stubs):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, ref)
    sys.path.insert(0, stubs)
    return ref


def unavailable(reason: str) -> dict:
    return {"impl": "reference", "unavailable": reason}


def run(n_gpus: int, steps: int, warmup: int, batch_per_gpu: int = 32, seq_len: int = 128,
        layer_num: int = 24) -> dict:
    if _paths() is None:
        return unavailable("baseline/_ref/scaelum is missing (reference not installed)")
    import torch

    if not torch.cuda.is_available():
        return unavailable("no CUDA device")
    import torch.distributed.rpc as rpc
    from torch import optim
    from torch.distributed.optim import DistributedOptimizer
    from torch.utils.data import Dataset

    import scaelum  # noqa: F401  (the unmodified reference)
    from scaelum import Runner, RpcModel
    from scaelum.builder import build_dataloader_from_cfg
    from scaelum.dynamics import Allocator, ParameterServer, WorkerManager
    from scaelum.model.bert import BertConfig
    from scaelum.registry import DATASET, HOOKS
    from scaelum.runner import Hook

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    assert world == n_gpus, f"--gpus {n_gpus} but WORLD_SIZE={world}"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    torch.cuda.set_device(local_rank)
    global_batch = batch_per_gpu * n_gpus

    # -- synthetic MNLI-shaped dataset registered into the reference's own registry ------------
    if "SynthMNLIForReference" not in DATASET._registry:
        class SynthMNLIForReference(Dataset):
            def __init__(self, num, seq, vocab, seed=0):
                g = torch.Generator().manual_seed(seed)
                self.ids = torch.randint(1000, vocab, (num, seq), generator=g)
                self.mask = torch.ones(num, seq, dtype=torch.long)
                self.seg = torch.zeros(num, seq, dtype=torch.long)
                self.seg[:, seq // 2:] = 1
                self.labels = torch.randint(0, 3, (num,), generator=g)

            def __len__(self):
                return self.ids.size(0)

            def __getitem__(self, idx):
                # GlueDataset item order (scaelum/dataset/bert_dataset.py:34-37)
                return (self.ids[idx], self.mask[idx], self.seg[idx]), self.labels[idx]

        DATASET._registry["SynthMNLIForReference"] = SynthMNLIForReference

    log_root = f"/tmp/sky_reference_arm/{n_gpus}gpus"
    if rank == 0:
        shutil.rmtree(log_root, ignore_errors=True)
        os.makedirs(log_root, exist_ok=True)

    print(f"[reference] init_rpc rank {rank}/{world}", flush=True)
    rpc.init_rpc(name=f"worker{rank}", rank=rank, world_size=world,
                 backend=rpc.BackendType.TENSORPIPE,
                 rpc_backend_options=rpc.TensorPipeRpcBackendOptions(
                     num_worker_threads=8, rpc_timeout=1200,
                     init_method=f"tcp://{os.environ['MASTER_ADDR']}:{os.environ['MASTER_PORT']}"))
    result = None
    try:
        if rank == 0:
            cfg = BertConfig(30522, hidden_size=1024, num_hidden_layers=24,
                             num_attention_heads=16, intermediate_size=4096)
            encoder = [dict(layer_type="BertLayer_Head", config=cfg.__dict__),
                       dict(layer_type="BertLayer_Body", config=cfg.__dict__),
                       dict(layer_type="BertLayer_Tail", config=cfg.__dict__)] * layer_num
            model_config = ([dict(layer_type="BertEmbeddings", config=cfg.__dict__)] + encoder
                            + [dict(layer_type="BertPooler", config=cfg.__dict__),
                               dict(layer_type="BertTailForClassification",
                                    hidden_dropout_prob=cfg.hidden_dropout_prob,
                                    hidden_size=cfg.hidden_size, num_classes=3)])
            worker_config = [dict(
                name=f"gpu-{i}", server_config=dict(host="localhost", port="8001"),
                extra_config=dict(slowdown=0, logging_config=dict(
                    mode="a", filename=osp.join(log_root, f"node-{i}-train.log")),
                    mem_limit=-1, cuda_device=i, module_to_cuda=True, output_to_cpu=True,
                    timer_config=dict(root=log_root))) for i in range(n_gpus)]
            wm = WorkerManager()
            wm.load_worker_pool_from_config(worker_config)
            for i, w in enumerate(wm.worker_pool):
                w.rank = i  # worker 0 lives in the master process (LocalModule), others remote
            data_loader = build_dataloader_from_cfg(
                dataset_cfg=dict(type="SynthMNLIForReference",
                                 num=global_batch * (steps + warmup + 1), seq=seq_len, vocab=30522),
                dataloader_cfg=dict(batch_size=global_batch, shuffle=False, num_workers=0,
                                    drop_last=True))
            allocator = Allocator(model_cfg=model_config, worker_manager=wm,
                                  model_benchmarker=None, device_benchmarker=None)
            wm = allocator.even_allocate()
            model = RpcModel(worker_manager=wm)
            dist_optim = DistributedOptimizer(optim.SGD, model.parameter_rrefs(), lr=0.001)

            class TimingHook(Hook):
                def __init__(self):
                    self.t0 = []
                    self.t1 = []
                    self.ev = []

                def before_iter(self, runner):
                    torch.cuda.synchronize()
                    e0 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                    self.ev.append([e0, None])
                    self.t0.append(time.time())

                def after_iter(self, runner):
                    torch.cuda.synchronize()
                    self.t1.append(time.time())
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record()
                    self.ev[-1][1] = e1

            hook = TimingHook()
            runner = Runner(model=model, parameter_server=None, worker_manager=wm,
                            optimizer=dist_optim, max_epochs=1, max_iters=warmup + steps - 1,
                            loss_cfg=dict(type="CrossEntropyLoss"),
                            timer_cfg=dict(root=log_root),
                            logging_cfg=dict(mode="a", filename=osp.join(log_root, "allocation.log")))
            runner.register_hook(hook)
            runner.train(data_loader)
            torch.cuda.synchronize()
            n = min(len(hook.t0), len(hook.t1))
            timed = list(range(warmup, min(n, warmup + steps)))
            wall = sum(hook.t1[i] - hook.t0[i] for i in timed)
            total = hook.t1[timed[-1]] - hook.t0[timed[0]]  # includes data loading between iters
            k = len(timed)
            result = {
                "impl": "reference",
                "metric": "BERT-large training throughput (sequences/s, whole job)",
                "value": global_batch * k / total, "unit": "sequences/s", "n_gpus": n_gpus,
                "steps": k, "warmup": warmup, "ms_per_step": 1e3 * total / k,
                "ms_per_step_compute_only": 1e3 * wall / k,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
                "data": "synthetic MNLI-shaped (random ids, seq 128), random-init weights",
                "config": {"model": f"BERT-large L={layer_num} H=1024 A=16 I=4096",
                           "global_batch": global_batch, "seq_len": seq_len,
                           "parallelism": f"pp{n_gpus} (reference RPC pipeline, one batch in flight)",
                           "allocator": "even", "optimizer": "SGD lr=1e-3",
                           "transport": "torch.distributed.rpc TensorPipe, CPU-staged"},
                "e2e": {"value": global_batch * k / total, "unit": "sequences/s",
                        "h2d_bytes_per_step": 3 * global_batch * seq_len * 8,
                        "d2h_bytes_per_step": global_batch * 3 * 4},
                "timing": "rank-0 wall clock around synchronous iterations (cuda.synchronize on "
                          "both sides); the reference's loop returns only after every stage "
                          "finished, so this is the max over ranks",
            }
    finally:
        rpc.shutdown()
    return result if result is not None else {"impl": "reference", "rank": rank}


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    out = run(a.gpus, a.steps, a.warmup)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(out))
