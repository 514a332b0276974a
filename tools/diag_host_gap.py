"""Diagnostics (GPU): how long the GPU waits for the host after every report read of the bench
workload (time from the return of the stream synchronize to the next kernel/graph launch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_optimizer, PER_RANK_BATCH
from dd4ml_b200 import _device, _lib, graphs
from dd4ml_b200.workloads import mnist_shape_batch

dev = torch.device("cuda", 0)
model, opt = build_optimizer(dev, 1)
X, Y = mnist_shape_batch(PER_RANK_BATCH, 1234, device=dev)
for _ in range(5):
    opt.step(X, Y)

state = {"t_sync": None, "gaps": [], "waits": []}
orig_read = _device.DeviceState.read
def read(self, full=False):
    t0 = time.perf_counter()
    r = orig_read(self, full)
    state["waits"].append(time.perf_counter() - t0)       # host blocked on the GPU
    state["t_sync"] = time.perf_counter()
    return r
_device.DeviceState.read = read
lib = _lib.get()
orig_call = lib._call
def call(name, launches, *a):
    if state["t_sync"] is not None and launches:
        state["gaps"].append(time.perf_counter() - state["t_sync"]); state["t_sync"] = None
    return orig_call(name, launches, *a)
lib._call = call
lib.__dict__ = {k: v for k, v in lib.__dict__.items() if not callable(v) or k in ("_call",)} | {"cdll": lib.cdll, "layout": lib.layout}
orig_replay = graphs.GraphedEval.__call__
def replay(self, *a, **k):
    if state["t_sync"] is not None:
        state["gaps"].append(time.perf_counter() - state["t_sync"]); state["t_sync"] = None
    return orig_replay(self, *a, **k)
graphs.GraphedEval.__call__ = replay

torch.cuda.synchronize()
t0 = time.perf_counter()
N = 30
for _ in range(N):
    opt.step(X, Y)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
g, w = state["gaps"], state["waits"]
print(f"{N} outer iterations: {wall / N * 1e3:.2f} ms each; {len(w) / N:.1f} reads/iteration; "
      f"host blocked on GPU {sum(w) / N * 1e3:.2f} ms/iteration; GPU idle after reads (host gap) "
      f"{sum(g) / N * 1e3:.3f} ms/iteration = {len(g) / N:.1f} gaps x {sum(g) / max(len(g), 1) * 1e6:.0f} us")

# implicit synchronisations (tensor.item(), float(tensor), nonzero ...) that the sync counter of
# DeviceState would not see: torch reports them in sync-debug mode
import warnings
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as wlist:
    warnings.simplefilter("always")
    for _ in range(5):
        opt.step(X, Y)
torch.cuda.set_sync_debug_mode("default")
msgs = [str(w.message)[:100] for w in wlist if "synchroniz" in str(w.message).lower()]
print(f"implicit synchronisations in 5 outer iterations: {len(msgs)}", set(msgs))
