"""pthflops is not installable offline; same contract (count_ops -> (flops, None)) on top of
torch.utils.flop_counter (used only by the reference's ModelBenchmarker)."""
import torch
from torch.utils.flop_counter import FlopCounterMode


def count_ops(model, data, print_readable=False, **kw):
    data = tuple(d.detach() if isinstance(d, torch.Tensor) else d
                 for d in (data if isinstance(data, (tuple, list)) else (data,)))
    with FlopCounterMode(display=False) as fc:
        model(*data)
    return fc.get_total_flops(), None
