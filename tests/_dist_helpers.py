"""Helpers for multi-process CPU tests (gloo, world_size > 1, no cluster needed)."""
import os
import socket
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn, args, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.manual_seed(0)
    torch.set_num_threads(2)
    try:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        ret[rank] = fn(rank, world, *args)
    except Exception:
        ret[rank] = "ERROR\n" + traceback.format_exc()
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def run_distributed(fn, world: int, *args):
    """Run fn(rank, world, *args) on `world` gloo ranks; returns the list of return values."""
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, free_port(), fn, args, ret), nprocs=world, join=True)
    out = [ret.get(r) for r in range(world)]
    errs = [f"[rank {r}] {o}" for r, o in enumerate(out) if isinstance(o, str) and o.startswith("ERROR")]
    if errs:
        # the first failing rank usually makes its peers fail with "connection closed": show all
        raise AssertionError("\n".join(errs))
    return out
