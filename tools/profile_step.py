"""Eager (no CUDA graph) single-GPU BERT training steps for a kernel launch list under ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skycomputing_b200 as sky  # noqa: E402
from skycomputing_b200.models import BertConfig, advance_rng, set_backend  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
set_backend("native")
torch.manual_seed(0)
cfg = BertConfig.bert_large()
enc = [dict(layer_type="BertLayer_Head", config=cfg.__dict__),
       dict(layer_type="BertLayer_Body", config=cfg.__dict__),
       dict(layer_type="BertLayer_Tail", config=cfg.__dict__)] * L
mc = ([dict(layer_type="BertEmbeddings", config=cfg.__dict__)] + enc
      + [dict(layer_type="BertPooler", config=cfg.__dict__),
         dict(layer_type="BertTailForClassification", hidden_dropout_prob=0.1, hidden_size=1024,
              num_classes=3)])
stage = sky.build_module_from_cfg(0, mc, dict(module_to_cuda=True, cuda_device=0))
stage._record_forward_time = False
stage.train()
opt = sky.build_optimizer(stage, dict(optim_type="SGD", lr=1e-3))
loss_fn = sky.build_loss(dict(type="CrossEntropyLoss"), torch.device("cuda", 0))
ids = torch.randint(1000, 30522, (B, 128), device="cuda")
tt = torch.zeros(B, 128, dtype=torch.long, device="cuda")
mask = torch.ones(B, 128, dtype=torch.long, device="cuda")
labels = torch.randint(0, 3, (B,), device="cuda")
for _ in range(steps):
    advance_rng()
    loss = loss_fn(stage(ids, tt, mask)[0], labels)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("done", float(loss))
