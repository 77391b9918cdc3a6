"""B1 arm: "the reference's NCCL(+cuBLAS) build" (BASELINE.json / SURVEY §6).

The reference itself has no NCCL path, so this harness assembles the strongest baseline its own
code allows without touching it:

* the layer list, the layer CLASSES (``scaelum.registry.LAYER``: ATen -> cuBLAS / cuDNN kernels)
  and the ``even`` split are the reference's own (``scaelum.dynamics.Allocator.even_allocate``);
* stage boundaries use ``torch.distributed`` NCCL ``send / recv`` device-to-device instead of the
  CPU-staged TensorPipe hop, backward through ``torch.autograd`` with the received gradient, one
  ``torch.optim.SGD`` per stage;
* precision: ``fp32`` (the reference's) or ``bf16`` = ``torch.autocast`` around the forward, i.e.
  what a user gets by switching the reference to mixed precision - the same-precision comparison
  for our bf16 kernels;
* schedule: one batch in flight like the reference (``micro_batches=1``) or GPipe-style
  micro-batching (all forwards, then all backwards) to give the baseline pipelining too.

Nothing from skycomputing_b200 is imported.  Times are CUDA events on every rank, max over ranks.
``bench.py --impl b1 [--b1-dtype fp32|bf16] [--b1-micro-batches m]`` prints the JSON line;
``bench.py`` (our arm) calls ``run_in_process`` after its own measurement for the ``vs_b1`` block.
"""
from __future__ import annotations

import os
import os.path as osp
import sys

HERE = osp.dirname(osp.abspath(__file__))


def _paths() -> bool:
    ref, stubs = osp.join(HERE, "_ref"), osp.join(HERE, "stubs")
    if not osp.isdir(osp.join(ref, "scaelum")):
        return False
    for p in (ref, stubs):
        if p not in sys.path:
            sys.path.insert(0, p)
    return True


def _layer_config(layer_num: int):
    from scaelum.model.bert import BertConfig

    cfg = BertConfig(30522, hidden_size=1024, num_hidden_layers=layer_num, num_attention_heads=16,
                     intermediate_size=4096)
    d = cfg.__dict__
    enc = [dict(layer_type="BertLayer_Head", config=d), dict(layer_type="BertLayer_Body", config=d),
           dict(layer_type="BertLayer_Tail", config=d)] * layer_num
    return ([dict(layer_type="BertEmbeddings", config=d)] + enc
            + [dict(layer_type="BertPooler", config=d),
               dict(layer_type="BertTailForClassification", hidden_dropout_prob=0.1,
                    hidden_size=1024, num_classes=3)])


def _even_spans(model_config, n_stages: int):
    """The reference's own even split (scaelum/dynamics/allocator.py:259-280)."""
    from scaelum.dynamics import Allocator, WorkerManager

    wm = WorkerManager()
    wm.load_worker_pool_from_config([dict(name=f"w{i}", server_config={}, extra_config={})
                                     for i in range(n_stages)])
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        wm = Allocator(model_config, wm, None, None).even_allocate()
    return [w.model_config for w in wm.worker_pool]


def run_in_process(n_gpus: int, steps: int, warmup: int, dtype: str = "bf16",
                   micro_batches: int = 1, batch_per_gpu: int = 32, seq_len: int = 128,
                   layer_num: int = 24, group=None) -> dict:
    """Needs an initialised NCCL process group of ``n_gpus`` ranks (one per GPU)."""
    if not _paths():
        return {"impl": "b1", "unavailable": "baseline/_ref/scaelum is missing"}
    import torch
    import torch.distributed as dist
    import torch.nn as nn

    from scaelum.builder import SequentialWrapper, build_layer

    rank = dist.get_rank(group) if n_gpus > 1 else 0
    device = torch.device("cuda", torch.cuda.current_device())
    m = max(1, micro_batches)
    global_batch = batch_per_gpu * n_gpus
    assert global_batch % m == 0
    mb = global_batch // m
    first, last = rank == 0, rank == n_gpus - 1
    torch.manual_seed(1234 + rank)
    spans = _even_spans(_layer_config(layer_num), n_gpus)
    layers = []
    for lc in spans[rank]:
        lc = dict(lc)
        layers.append(build_layer(lc.pop("layer_type"), **lc))
    stage = SequentialWrapper(*layers).to(device)
    stage.train()
    opt = torch.optim.SGD(stage.parameters(), lr=1e-3)
    loss_fn = nn.CrossEntropyLoss()
    autocast = dtype == "bf16"
    act_dtype = torch.bfloat16 if autocast else torch.float32

    g = torch.Generator().manual_seed(0)
    batches = []
    for _ in range(4):
        ids = torch.randint(1000, 30522, (global_batch, seq_len), generator=g)
        batches.append((ids.pin_memory(), torch.zeros_like(ids).pin_memory(),
                        torch.ones_like(ids).pin_memory(),
                        torch.randint(0, 3, (global_batch,), generator=g).pin_memory()))
    def fwd(args):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = stage(*args)
        return out if isinstance(out, (tuple, list)) else (out,)

    def diff_flags(n):
        # every floating boundary tensor carries a gradient except a trailing attention mask
        return [i < n - 1 or n == 1 for i in range(n)]

    # ---- shape propagation: the reference's even split may cut INSIDE a transformer block (after
    # BertLayer_Head: 2 tensors, after BertLayer_Body: 3 tensors of different widths), so every
    # stage learns the layout of its inputs from a dry run of the stage in front of it
    in_meta = None
    for r in range(n_gpus):
        box = [None]
        if rank == r:
            if first:
                probe = tuple(t[:mb].to(device) for t in batches[0][:3])
            else:
                probe = tuple(torch.zeros(shape, device=device, dtype=act_dtype) for shape in in_meta)
            with torch.no_grad():
                outs = fwd(probe)
            box = [[tuple(o.shape) for o in outs]]
            del outs, probe
        if n_gpus > 1:
            dist.broadcast_object_list(box, src=r, group=group)
        if rank == r + 1:
            in_meta = box[0]
    torch.cuda.empty_cache()

    def step(i: int):
        ids, tt, am, labels = (t.to(device, non_blocking=True) for t in batches[i % 4])
        saved = []
        loss_val = None
        for j in range(m):                                       # ---- forwards
            sl = slice(j * mb, (j + 1) * mb)
            if first:
                args = (ids[sl], tt[sl], am[sl])
            else:
                args = []
                for shape, d in zip(in_meta, diff_flags(len(in_meta))):
                    t = torch.empty(shape, device=device, dtype=act_dtype)
                    dist.recv(t, rank - 1, group=group)
                    args.append(t.requires_grad_(True) if d else t)
                args = tuple(args)
            outs = fwd(args)
            if last:
                loss = loss_fn(outs[0].float(), labels[sl]) / m
                saved.append((args, (loss,)))
            else:
                # boundary tensors travel in the compute dtype (bf16 under autocast, where
                # LayerNorm itself returns fp32): a differentiable cast keeps autograd intact
                ys = tuple(o.to(act_dtype) for o in outs)
                for y in ys:
                    dist.send(y.detach().contiguous(), rank + 1, group=group)
                saved.append((args, ys))
        for j in range(m):                                       # ---- backwards
            args, outs = saved[j]
            if last:
                outs[0].backward()
                loss_val = outs[0].detach() if loss_val is None else loss_val + outs[0].detach()
            else:
                ts, gs = [], []
                for y, d in zip(outs, diff_flags(len(outs))):
                    if d:
                        gout = torch.empty_like(y)
                        dist.recv(gout, rank + 1, group=group)
                        if y.requires_grad:
                            ts.append(y)
                            gs.append(gout)
                torch.autograd.backward(ts, gs)
            if not first:
                for a, d in zip(args, diff_flags(len(args))):
                    if d:
                        g_in = a.grad if a.grad is not None else torch.zeros_like(a)
                        dist.send(g_in.contiguous(), rank - 1, group=group)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss_val

    def sync():
        torch.cuda.synchronize()
        if n_gpus > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize()

    for i in range(max(warmup, 3)):
        step(i)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = None
    for i in range(steps):
        loss = step(i)
    if loss is not None:
        loss = float(loss.item())                                # D2H read of the result
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if n_gpus > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    ms = float(t.item())
    del stage, opt
    torch.cuda.empty_cache()
    return {
        "impl": "b1", "metric": "BERT-large training throughput (sequences/s, whole job)",
        "value": global_batch * steps / (ms * 1e-3), "unit": "sequences/s", "n_gpus": n_gpus,
        "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "weak", "dtype": dtype,
        "data": "synthetic MNLI-shaped (random ids, seq 128), random-init weights",
        "config": {"model": f"BERT-large L={layer_num} H=1024 A=16 I=4096", "global_batch": global_batch,
                   "seq_len": seq_len, "parallelism": f"pp{n_gpus}", "micro_batches": m,
                   "schedule": "one batch in flight" if m == 1 else "gpipe (all F, then all B)",
                   "layers": "reference classes (scaelum.registry.LAYER), ATen / cuBLAS kernels",
                   "allocator": "reference even_allocate", "transport": "NCCL send/recv",
                   "optimizer": "torch.optim.SGD lr=1e-3",
                   "precision": "torch.autocast(bfloat16)" if autocast else "fp32 (torch default)"},
        "final_loss": loss,
    }


def run(n_gpus: int, steps: int, warmup: int, dtype: str = "bf16", micro_batches: int = 1,
        layer_num: int = 24) -> dict:
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        return {"impl": "b1", "unavailable": "no CUDA device"}
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    assert world == n_gpus, f"--gpus {n_gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    if n_gpus > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))
    try:
        out = run_in_process(n_gpus, steps, warmup, dtype, micro_batches, layer_num=layer_num)
    finally:
        if n_gpus > 1:
            dist.destroy_process_group()
    return out
