"""GPU diagnostic: what the end-to-end feed adds to an APTS_D outer iteration (bench workload, N = 1).
Variants: resident batch / + loss read-back and stream sync per step / pinned-host feed without the
per-step sync / the full e2e step of bench.py.  CUDA events around 30 steps each (L2 flush inside)."""
import argparse, itertools, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dd4ml_b200.dataloaders import HostBatchPrefetcher

dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False
flush = torch.empty(bench.L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
job = bench.build_job("ffnn_apts_d", dev, 1, 0)
res = job.upload(job.host_batch)
host_loss = torch.empty((), dtype=torch.float64).pin_memory()
feed = iter(HostBatchPrefetcher(itertools.repeat(job.host_batch), dev))


def v_resident(i): job.step(res)
def v_resident_sync(i):
    loss = job.step(res); host_loss.copy_(loss.detach().double(), non_blocking=True); torch.cuda.current_stream().synchronize()
def v_feed_nosync(i): job.step(next(feed))
def v_e2e(i):
    loss = job.step(next(feed)); host_loss.copy_(loss.detach().double(), non_blocking=True); torch.cuda.current_stream().synchronize()
def v_upload_sync(i):
    loss = job.step(job.upload(job.host_batch)); host_loss.copy_(loss.detach().double(), non_blocking=True); torch.cuda.current_stream().synchronize()


for name, fn in (("resident", v_resident), ("resident+readback+sync", v_resident_sync), ("feed, no sync", v_feed_nosync),
                 ("e2e (feed+readback+sync)", v_e2e), ("blocking upload+readback+sync", v_upload_sync), ("resident", v_resident)):
    for i in range(5):
        fn(i)
    ev0 = job.opt.grad_evals if hasattr(job.opt, "grad_evals") else 0
    ms, wall = bench.timed_region(fn, 30, 1, dev, flush)
    print(f"{name:32s} {ms / 30:.3f} ms/step  (wall {1e3 * wall / 30:.3f})")
