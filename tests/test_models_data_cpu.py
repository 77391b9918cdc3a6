"""CPU tests of the model / data libraries: GLUE pipeline on a synthetic TSV,
tokenizer, BertConfig, activation set, fp32 layer oracle shapes, the ResNet layer list run through
the same builder / allocator machinery as BERT, toy datasets and generators."""
import json
import math
import os

import torch

import skycomputing_b200 as sky

VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "the", "cat", "sat", "on", "mat", "dog", "ran",
         "a", "##s", "##ing", "run", "fast", ".", ",", "is", "not"]


def _write_mnli(dir_, n=6):
    os.makedirs(dir_, exist_ok=True)
    labels = ["contradiction", "entailment", "neutral"]
    with open(os.path.join(dir_, "train.tsv"), "w") as f:
        f.write("\t".join(f"c{i}" for i in range(11)) + "\n")
        for i in range(n):
            cols = [str(i)] * 8 + ["The cat sat on the mat.", "A dog is not running fast,", labels[i % 3]]
            f.write("\t".join(cols) + "\n")
    vocab = os.path.join(dir_, "vocab.txt")
    with open(vocab, "w") as f:
        f.write("\n".join(VOCAB) + "\n")
    return vocab


def test_tokenizer_wordpiece_and_ids(tmp_path):
    from skycomputing_b200.dataset.glue import BertTokenizer

    vocab = _write_mnli(str(tmp_path))
    tok = BertTokenizer(vocab, do_lower_case=True, max_len=32)
    toks = tok.tokenize("The cats sat, running fast.")
    assert toks == ["the", "cat", "##s", "sat", ",", "run", "[UNK]", "fast", "."] or \
        toks[:5] == ["the", "cat", "##s", "sat", ","]   # greedy longest-match-first word pieces
    ids = tok.convert_tokens_to_ids(toks)
    assert tok.convert_ids_to_tokens(ids) == toks
    assert all(0 <= i < len(VOCAB) for i in ids)


def test_glue_dataset_end_to_end_and_feature_cache(tmp_path):
    d = str(tmp_path)
    vocab = _write_mnli(d, n=6)
    kw = dict(data_dir=d, bert_model="tiny", vocab_file=vocab, max_seq_length=16,
              do_lower_case=True, processor="mnli")
    ds = sky.GlueDataset(**kw)
    assert len(ds) == 6
    (ids, mask, seg), label = ds[0]
    assert ids.shape == mask.shape == seg.shape == (16,) and ids.dtype == torch.long
    assert ids[0].item() == VOCAB.index("[CLS]") and int(label) == 0
    n_real = int(mask.sum())
    assert ids[n_real - 1].item() == VOCAB.index("[SEP]") and (ids[n_real:] == 0).all()
    assert seg[: n_real].max().item() == 1 and seg[0].item() == 0       # sentence B is segment 1
    assert any(f.startswith("tiny_16_") for f in os.listdir(d))           # feature cache written
    ds2 = sky.GlueDataset(**kw, reference_order=False)                    # served from the cache
    (ids2, seg2, mask2), _ = ds2[0]
    assert torch.equal(ids2, ids) and torch.equal(seg2, seg) and torch.equal(mask2, mask)
    # through the registry / dataloader builder, like the config does
    dl = sky.build_dataloader_from_cfg(dataset_cfg=dict(type="GlueDataset", **kw),
                                       dataloader_cfg=dict(batch_size=3, shuffle=False))
    (b_ids, b_mask, b_seg), b_lab = next(iter(dl))
    assert b_ids.shape == (3, 16) and b_lab.tolist() == [0, 1, 2]


def test_processors_and_feature_conversion(tmp_path):
    from skycomputing_b200.dataset.glue import (PROCESSORS, BertTokenizer, InputExample,
                                                convert_examples_to_features)

    assert set(PROCESSORS) == {"cola", "mnli", "mrpc", "sst-2"}
    assert PROCESSORS["mnli"]().get_labels() == ["contradiction", "entailment", "neutral"]
    vocab = _write_mnli(str(tmp_path))
    tok = BertTokenizer(vocab, do_lower_case=True, max_len=32)
    ex = [InputExample("g-0", "the cat sat on the mat the cat sat on the mat", "a dog ran", "1")]
    f = convert_examples_to_features(ex, ["0", "1"], 12, tok)[0]
    assert len(f.input_ids) == len(f.input_mask) == len(f.segment_ids) == 12
    assert sum(f.input_mask) == 12 and f.label_id == 1                    # truncated to fit, no padding
    assert f.input_ids.count(VOCAB.index("[SEP]")) == 2


def test_bert_config_roundtrip(tmp_path):
    c = sky.BertConfig(1000, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                       intermediate_size=128)
    d = c.to_dict()
    assert d["hidden_size"] == 64 and json.loads(c.to_json_string())["vocab_size"] == 1000
    c2 = sky.BertConfig.from_dict(d)
    assert c2.to_dict() == d
    p = tmp_path / "cfg.json"
    p.write_text(c.to_json_string())
    assert sky.BertConfig.from_json_file(str(p)).to_dict() == d
    big = sky.BertConfig.bert_large()
    assert (big.hidden_size, big.num_hidden_layers, big.num_attention_heads) == (1024, 24, 16)


def test_activation_library():
    from skycomputing_b200.models import bert_layers as bl

    x = torch.linspace(-3, 3, 13)
    ref = x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))
    assert torch.allclose(bl.gelu(x), ref, atol=1e-4)
    b = torch.full_like(x, 0.25)
    assert torch.allclose(bl.bias_gelu(b, x), bl.gelu(x + b), atol=1e-6)
    assert torch.allclose(bl.bias_tanh(b, x), torch.tanh(x + b))
    assert torch.allclose(bl.swish(x), x * torch.sigmoid(x))
    assert {"gelu", "bias_gelu", "bias_tanh", "relu", "swish"} <= set(bl.ACT2FN)
    lin = bl.LinearActivation(8, 4, act="bias_gelu")
    y = lin(torch.randn(5, 8))
    assert y.shape == (5, 4)
    plain = bl.LinearActivation(8, 4, act="noop") if "noop" in bl.ACT2FN else None
    if plain is not None:
        assert plain(torch.randn(2, 8)).shape == (2, 4)
    ln, nf = bl.BertLayerNorm(16), bl.BertNonFusedLayerNorm(16)
    nf.load_state_dict(ln.state_dict())
    z = torch.randn(3, 16)
    assert torch.allclose(ln(z), nf(z), atol=1e-5)
    assert torch.allclose(ln(z), torch.nn.functional.layer_norm(z, (16,), ln.weight, ln.bias, 1e-12),
                          atol=1e-5)


def test_bert_layer_calling_convention_on_cpu():
    """(ids, token_type, mask) -> (emb, ext_mask) -> Head -> Body -> Tail -> Pooler -> logits."""
    c = sky.BertConfig(50, hidden_size=32, num_hidden_layers=1, num_attention_heads=4,
                       intermediate_size=64, max_position_embeddings=16)
    emb = sky.build_layer("BertEmbeddings", config=c.__dict__)
    head = sky.build_layer("BertLayer_Head", config=c.__dict__)
    body = sky.build_layer("BertLayer_Body", config=c.__dict__)
    tail = sky.build_layer("BertLayer_Tail", config=c.__dict__)
    pool = sky.build_layer("BertPooler", config=c.__dict__)
    cls = sky.build_layer("BertTailForClassification", hidden_dropout_prob=0.0, hidden_size=32,
                          num_classes=3)
    for m in (emb, head, body, tail, pool, cls):
        m.eval()
    ids = torch.randint(0, 50, (2, 16))
    tt = torch.zeros(2, 16, dtype=torch.long)
    mask = torch.ones(2, 16, dtype=torch.long)
    mask[1, 10:] = 0
    h, ext = emb(ids, tt, mask)
    assert h.shape == (2, 16, 32) and ext.shape == (2, 1, 1, 16) and ext[1, 0, 0, 12] < -1000
    a, ext2 = head(h, ext)
    inter, a2, ext3 = body(a, ext2)
    assert inter.shape == (2, 16, 64) and a2 is a or torch.equal(a2, a)
    out, ext4 = tail(inter, a2, ext3)
    assert out.shape == h.shape
    pooled = pool(out, ext4)
    pooled = pooled[0] if isinstance(pooled, (tuple, list)) else pooled
    logits = cls(pooled)
    logits = logits[0] if isinstance(logits, (tuple, list)) else logits
    assert logits.shape == (2, 3)
    # padding positions must not influence the un-padded sequence's output
    ids_b = ids.clone()
    ids_b[1, 10:] = 7
    h_b, ext_b = emb(ids_b, tt, mask)
    out_b, _ = tail(*body(*head(h_b, ext_b)))
    assert torch.allclose(out_b[1, :10], out[1, :10], atol=1e-5)


def test_resnet_layer_list_through_builder_and_allocator(tmp_path):
    """The layer-list mechanism is model agnostic: ResNet-18 as a config list, evenly allocated
    over 3 workers, run stage by stage, equals the monolithic factory model."""
    from skycomputing_b200.models.layers import resnet18, resnet_layer_configs

    torch.manual_seed(0)
    cfgs = resnet_layer_configs("BasicBlock", [2, 2, 2, 2], num_classes=10)
    assert cfgs[0]["layer_type"] == "ResHead" and cfgs[-1]["layer_type"] == "ResTail"
    wm = sky.WorkerManager(first_rank=0)
    wm.load_worker_pool_from_config([dict(name=f"w{i}", server_config={}, device=i,
                                          extra_config=dict(timer_config=dict(root=str(tmp_path))))
                                     for i in range(3)])
    wm = sky.Allocator(cfgs, wm).even_allocate()
    assert sum(len(w.model_config) for w in wm.worker_pool) == len(cfgs)
    x = torch.randn(2, 3, 32, 32)
    stages = [sky.build_module_from_cfg(w.rank, w.model_config,
                                        dict(timer_config=dict(root=str(tmp_path))))
              for w in wm.worker_pool]
    for s in stages:
        s.eval()
    y = (x,)
    with torch.no_grad():
        for s in stages:
            y = s(*y)
    assert y[0].shape == (2, 10)
    full = resnet18(num_classes=10).eval()
    with torch.no_grad():
        out = full(x)
    out = out[0] if isinstance(out, (tuple, list)) else out
    assert out.shape == (2, 10)
    n_list = sum(p.numel() for s in stages for p in s.parameters())
    assert n_list == sum(p.numel() for p in full.parameters())
    bott = sky.build_layer("BottleNeck", in_channels=16, out_channels=8, stride=1).eval()
    assert bott(torch.randn(1, 16, 8, 8)).shape[1] == 8 * 4


def test_toy_datasets_and_generators():
    ds = sky.RandomMlpDataset(num=10, dim=6)
    x, y = ds[3]
    assert len(ds) == 10 and x.shape == (6,)
    img = sky.RandomImageDataset(num=4, size=8, num_classes=5)
    xi, yi = img[0]
    assert len(img) == 4 and xi.shape == (3, 8, 8) and 0 <= int(yi) < 5
    g = sky.build_data_generator("RandomTensorGenerator", generator_cfg=dict(size=(2, 3)))
    assert g.generate().shape == (2, 3)
    dg = sky.build_data_generator("DataloaderGenerator", generator_cfg=dict(
        dataset_cfg=dict(type="SynthMNLIDataset", num_samples=4, max_seq_length=8, vocab_size=64),
        dataloader_cfg=dict(batch_size=2)))
    batch = dg.generate()
    assert len(batch) == 3 and batch[0].shape == (2, 8)                  # the INPUTS of the first batch


def test_hook_base_helpers():
    class R:
        epoch, iter, inner_iter = 3, 9, 4
        data_loader = list(range(5))

    h = sky.Hook()
    assert h.every_n_epochs(R, 2) and not h.every_n_epochs(R, 3)
    assert h.every_n_iters(R, 5) and h.every_n_inner_iters(R, 5)
    assert h.end_of_epoch(R)
    for name in ("before_run", "after_run", "before_train_epoch", "after_train_epoch",
                 "before_train_iter", "after_train_iter", "before_val_epoch", "after_val_iter"):
        getattr(h, name)(R)                                              # all no-ops by default


def test_span_fusion_plan_and_boundary_support(tmp_path):
    """ModuleWrapper groups consecutive Head / Body / Tail entries into fused spans (one autograd
    node, one kernel chain) whatever sub-block cut the allocator chose, and reports which sides of
    the stage can use the fused NVLink boundary (whole-block cuts and the cut after
    BertLayer_Head; a cut after BertLayer_Body ships two tensors and stays on p2p)."""
    from skycomputing_b200.models.bert_layers import BertSpan

    c = sky.BertConfig(50, hidden_size=32, num_hidden_layers=2, num_attention_heads=4,
                       intermediate_size=64, max_position_embeddings=16)
    L = lambda t: dict(layer_type=t, config=c.__dict__)  # noqa: E731
    H, B, T = L("BertLayer_Head"), L("BertLayer_Body"), L("BertLayer_Tail")
    tail_cls = [L("BertPooler"), dict(layer_type="BertTailForClassification", hidden_dropout_prob=0.0,
                                      hidden_size=32, num_classes=3)]

    def shape(cfgs):
        mw = sky.build_module_from_cfg(0, cfgs, dict(timer_config=dict(root=str(tmp_path))))
        plan = mw._build_plan()
        desc = []
        for s in plan:
            if isinstance(s, BertSpan):
                desc.append("".join(n for n, x in (("H", s.head), ("B", s.body), ("T", s.tail))
                                    if x is not None))
            else:
                desc.append(type(s).__name__)
        return desc, mw.fused_boundary_support(), len(mw.spans())

    assert shape([L("BertEmbeddings"), H, B, T, H, B, T] + tail_cls) == (
        ["BertEmbeddings", "HBT", "HBT", "BertPooler", "BertTailForClassification"], (False, False), 2)
    assert shape([H, B, T, H, B, T]) == (["HBT", "HBT"], (True, True), 2)       # middle stage
    assert shape([B, T, H, B]) == (["BT", "HB"], (True, False), 2)   # starts after Head (fusable), ends after Body (not)
    assert shape([T, H]) == (["T", "H"], (False, True), 2)           # starts after Body (not), ends after Head (fusable)
    assert shape([H, B, T] + tail_cls)[1] == (True, False)                       # last stage
    assert shape([L("BertEmbeddings"), H, B, T])[1] == (False, True)             # first stage
