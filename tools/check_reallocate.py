"""ReallocateHook on real GPUs (run under torchrun, >= 2 ranks): a pipeline with an even split
trains, GPU 1 is throttled by the device-side slowdown, the hook re-benchmarks every 3 iterations,
the exact allocator moves blocks to the fast GPU, layers migrate, and Runner.rebuild swaps the
fused-boundary engine (peer regions, flags, whole-step CUDA graph) in the middle of the run.
Prints one line: REALLOC {json}."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skycomputing_b200 as sky  # noqa: E402
from skycomputing_b200.models import BertConfig, set_backend  # noqa: E402

LAYERS, BATCH, MICRO = 8, 16, 2


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    set_backend("native")
    torch.manual_seed(3)
    cfg = BertConfig(30522, hidden_size=1024, num_hidden_layers=LAYERS, num_attention_heads=16,
                     intermediate_size=4096, hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    enc = [dict(layer_type="BertLayer_Head", config=cfg.__dict__),
           dict(layer_type="BertLayer_Body", config=cfg.__dict__),
           dict(layer_type="BertLayer_Tail", config=cfg.__dict__)] * LAYERS
    model_config = ([dict(layer_type="BertEmbeddings", config=cfg.__dict__)] + enc
                    + [dict(layer_type="BertPooler", config=cfg.__dict__),
                       dict(layer_type="BertTailForClassification", hidden_dropout_prob=0.0,
                            hidden_size=1024, num_classes=3)])
    root = "/tmp/sky_realloc"
    os.makedirs(root, exist_ok=True)
    workers = [dict(name=f"gpu-{i}", server_config={}, device=i,
                    extra_config=dict(module_to_cuda=True, cuda_device=local, mem_limit=-1,
                                      slowdown=2.0 if i == 1 else 0,      # GPU 1 is 3x slower
                                      timer_config=dict(root=root)))
               for i in range(world)]
    wm = sky.WorkerManager(first_rank=0)
    wm.load_worker_pool_from_config(workers)
    wm = sky.Allocator(model_config, wm, granularity="block").even_allocate()
    layers_before = [sum(len(w.model_config) for w in wm.worker_pool if w.device == d)
                     for d in range(world)]
    model = sky.RpcModel(wm, this_rank=rank)
    optim_cfg = dict(optim_type="SGD", lr=1e-3)
    opt = sky.build_optimizer(model.optim_module, dict(optim_cfg))
    runner = sky.Runner(model=model, parameter_server=None, worker_manager=wm, optimizer=opt,
                        max_epochs=1, max_iters=9, loss_cfg=dict(type="CrossEntropyLoss"),
                        timer_cfg=dict(root=root), logging_cfg=None, micro_batches=MICRO,
                        schedule="1f1b", boundary="fused", device=dev)

    def factory(pool):
        gen = sky.build_data_generator("DataloaderGenerator", generator_cfg=dict(
            dataset_cfg=dict(type="SynthMNLIDataset", num_samples=BATCH // MICRO, max_seq_length=128),
            dataloader_cfg=dict(batch_size=BATCH // MICRO)))
        mb = sky.ModelBenchmarker(model_config, gen, device="cpu", analytic=True)
        db = sky.DeviceBenchmarker(pool, None, model_config=[], iterations=10, warmup=2,
                                   proxy="bert_block",
                                   block_shape=dict(tokens=BATCH // MICRO * 128, hidden=1024,
                                                    intermediate=4096))
        return sky.Allocator(model_config, pool, mb, db, granularity="block", solver="exact")

    hook = sky.ReallocateHook(interval=3, allocator_factory=factory, optimizer_cfg=optim_cfg,
                              allocate_type="optimal", min_gain=0.05)
    runner.register_hook(hook)
    model.train(True)
    dl = sky.build_dataloader_from_cfg(
        dataset_cfg=dict(type="SynthMNLIDataset", num_samples=BATCH * 9, max_seq_length=128, seed=5),
        dataloader_cfg=dict(batch_size=BATCH, shuffle=False, drop_last=True))
    losses = []
    orig = runner.train_iteration

    def spy(data, labels):
        out = orig(data, labels)
        if out is not None:
            losses.append(round(out, 4))
        return out

    runner.train_iteration = spy
    runner.train(dl)
    torch.cuda.synchronize()
    eng = runner.engine
    info = dict(rank=rank, migrations=hook.migrations, losses=losses,
                # layers on THIS device (the pool is in pipeline order, which the exact allocator
                # may permute: look the device up, do not index by rank)
                layers=sum(len(w.model_config) for w in runner.worker_manager.worker_pool
                           if w.device == rank),
                fused=bool(eng.in_fused or eng.out_fused), graph=eng._graph is not None,
                err=eng.fused.error_code() if eng.fused is not None else 0,
                decision=hook.last_decision)
    gathered = [None] * world
    dist.all_gather_object(gathered, info)
    if rank == 0:
        last = [g for g in gathered if g["losses"]][0]
        print("REALLOC " + json.dumps(dict(
            migrations=gathered[0]["migrations"], layers_before=layers_before,
            layers_after=[g["layers"] for g in gathered], losses=last["losses"],
            fused_after=all(g["fused"] for g in gathered),
            graph_after=all(g["graph"] for g in gathered),
            err_any=max(g["err"] for g in gathered), decision=gathered[0]["decision"])), flush=True)
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
