"""GPU (torchrun, N>=2): cost of torch.distributed all-reduce at the hot path's payload sizes and a
phase breakdown of one APTS_D outer iteration."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for n in (217356, 2719106):
    t = torch.randn(n, device=dev)
    for _ in range(10): dist.all_reduce(t, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): dist.all_reduce(t, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 200
    # with a dependent tiny kernel between (like the real path)
    t0 = time.perf_counter()
    for _ in range(200):
        dist.all_reduce(t, op=dist.ReduceOp.AVG); x = t[n - 1].clone(); float(x)
    dt2 = (time.perf_counter() - t0) / 200
    if rank == 0: print(f"n={n}: all_reduce {dt*1e6:.1f} us back-to-back, {dt2*1e6:.1f} us with a host read after each")
from bench import build_optimizer, PER_RANK_BATCH
from dd4ml_b200.workloads import mnist_shape_batch
model, opt = build_optimizer(dev, world)
X, Y = mnist_shape_batch(PER_RANK_BATCH, 1234 + rank, device=dev)
for _ in range(5): opt.step(X, Y)
acc = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(*a, **k); torch.cuda.synchronize()
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; acc[name + "_n"] = acc.get(name + "_n", 0) + 1
        return r
    return w
opt.glob_closure_main = timed("glob_closure", opt.glob_closure_main)
opt.loc_steps = timed("loc_steps", opt.loc_steps)
opt.glob_steps = timed("glob_steps", opt.glob_steps)
opt._allreduce_mean = timed("allreduce_mean", opt._allreduce_mean)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
N = 20
for _ in range(N): opt.step(X, Y)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
if rank == 0:
    print(f"outer iteration {tot/N*1e3:.2f} ms (with sync instrumentation)")
    for k in sorted(acc):
        if not k.endswith("_n"): print(f"  {k:16s} {acc[k]/N*1e3:7.3f} ms/iter  ({acc[k+'_n']/N:.1f} calls/iter, {acc[k]/acc[k+'_n']*1e6:.0f} us/call)")
dist.barrier(); dist.destroy_process_group()
