"""Host-side scheduling of the parameter-gradient kernels (ops/functions.py), exercised on CPU with
the kernel wrappers replaced by recorders: deferral queue order, overwrite-first stores, bias column
sums riding in the LayerNorm parameter-gradient launch."""
import torch

from skycomputing_b200.ops import functions as F


class _Recorder:
    def __init__(self):
        self.calls = []

    def gemm(self, a, b, **kw):
        self.calls.append(("gemm", kw.get("accumulate"), kw["out"].data_ptr()))

    def colsum_(self, x, out):
        self.calls.append(("colsum", out.data_ptr()))

    def ln_param_grad(self, dy, z, mean, rstd, dgamma, dbeta, x2=None, out2=None):
        self.calls.append(("ln", dgamma.data_ptr(), None if out2 is None else out2.data_ptr()))


def _bank(*shape):
    return F.ParamBank([torch.nn.Parameter(torch.zeros(*shape))], need_shadow=False)


def test_deferred_queue_keeps_program_order_and_flushes_once(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(F, "nat", rec)
    w, b, g, bn, bo = _bank(4, 3), _bank(4), _bank(3), _bank(3), _bank(3)
    x = torch.zeros(5, 3)
    F.set_wgrad_deferral(True)
    try:
        F._wgrad(torch.zeros(5, 4), x, w, b)
        F._ln_pgrad(x, x, torch.zeros(5), torch.zeros(5), g, bn, x, bo)
        F._wgrad(torch.zeros(5, 4), x, w, None)          # bias handled elsewhere: no colsum
        assert rec.calls == [] and F.pending_wgrads() == 3
        items = F.flush_wgrads()
        assert len(items) == 3 and F.pending_wgrads() == 0
        kinds = [c[0] for c in rec.calls]
        assert kinds == ["gemm", "colsum", "ln", "gemm"]
        assert rec.calls[2] == ("ln", g.grad().data_ptr(), bo.grad().data_ptr())
        assert F.flush_wgrads() == []
    finally:
        F.set_wgrad_deferral(False)


def test_overwrite_first_store_then_accumulate(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(F, "nat", rec)
    w, b = _bank(4, 3), _bank(4)
    x, g = torch.zeros(5, 3), torch.zeros(5, 4)
    F._wgrad(g, x, w, b)                                  # first backward ever: accumulate into zeros
    assert w.wgrad_target and rec.calls[0][:2] == ("gemm", True)
    w.overwrite_first = True                              # what FusedSGD decides after that step
    w.fresh = True                                        # ... and sets after every optimizer step
    rec.calls.clear()
    F._wgrad(g, x, w, b)                                  # micro-batch 0 of the next step: store
    F._wgrad(g, x, w, b)                                  # micro-batch 1: accumulate
    assert [c[1] for c in rec.calls if c[0] == "gemm"] == [False, True]
    assert not w.fresh
    b2 = _bank(4, 3)                                      # a bank the optimizer did not mark
    b2.fresh = True
    rec.calls.clear()
    F._wgrad(g, x, b2, None)
    assert rec.calls == [("gemm", True, b2.grad().data_ptr())]


def test_fused_sgd_marks_only_wgrad_targets(monkeypatch):
    """FusedSGD.step sets `fresh` on banks it decided to leave un-zeroed, and only on those."""
    from skycomputing_b200.parallel import optim

    class FakeExt:
        def pack_sgd_descriptors(self, descs):
            self.descs = descs
            return bytes(48 * len(descs))

        def sgd_multi(self, **kw):
            self.stepped = kw

    class FakeNat:
        def __init__(self):
            self._e = FakeExt()

        def ext(self):
            return self._e

    w, b = _bank(4, 3), _bank(4)
    w.wgrad_target = True
    opt = optim.FusedSGD.__new__(optim.FusedSGD)
    opt.lr, opt.momentum, opt.weight_decay = 0.1, 0.0, 0.0
    opt._nat = FakeNat()
    opt.banks, opt._mom = [w, b], [None, None]
    opt._rest_opt, opt._desc_dev, opt._max_numel = None, None, 0
    opt._overwrite_ok = True
    opt.param_groups = [dict(lr=0.1, momentum=0.0, weight_decay=0.0)]
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: type("S", (), {"cuda_stream": 0})())
    opt.step()
    assert [d[5] for d in opt._nat.ext().descs] == [1, 0]
    assert w.fresh and not b.fresh
    opt.zero_grad()
    assert not w.fresh
