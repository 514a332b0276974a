"""Multi-rank test plumbing: one process per rank.  On a box with at least `ws` GPUs every rank
gets its own device and the process group is NCCL; on a 1-GPU box (the driver's GPU test tier) the
ranks share cuda:0 and the group is gloo, which all-reduces / broadcasts CUDA tensors through host
memory -- the optimizers' own kernels and the payload layout are the same, only the transport differs."""
import os
import socket

import torch
import torch.multiprocessing as mp


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def init_rank(rank, ws, port):
    """Returns (device string, backend)."""
    import torch.distributed as dist
    ndev = torch.cuda.device_count()
    idx = rank % ndev
    backend = "nccl" if ndev >= ws else "gloo"
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(idx)
    dist.init_process_group(backend, rank=rank, world_size=ws)
    return f"cuda:{idx}", backend


def run_ranks(target, ws, args=(), timeout=900):
    """Spawn `ws` processes running target(rank, ws, port, *args, q); returns {rank: result}."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=target, args=(r, ws, port) + tuple(args) + (q,)) for r in range(ws)]
    for p in procs:
        p.start()
    try:
        outs = dict(q.get(timeout=timeout) for _ in range(ws))
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.terminate()
    for r, o in outs.items():
        if isinstance(o, dict) and "error" in o:
            raise AssertionError(f"rank {r} failed:\n{o['error']}")
    return outs


def guarded(fn):
    """Decorator for workers: report exceptions through the queue instead of hanging the parent."""
    import functools
    import traceback

    @functools.wraps(fn)
    def wrapper(rank, ws, port, *rest):
        q = rest[-1]
        try:
            return fn(rank, ws, port, *rest)
        except Exception:  # noqa: BLE001
            q.put((rank, {"error": traceback.format_exc()}))
    return wrapper
