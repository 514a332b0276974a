"""Diagnostics (GPU): cProfile of the host side of the bench workload's outer iteration."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_optimizer, PER_RANK_BATCH
from dd4ml_b200.workloads import mnist_shape_batch
dev = torch.device("cuda", 0)
model, opt = build_optimizer(dev, 1)
X, Y = mnist_shape_batch(PER_RANK_BATCH, 1234, device=dev)
for _ in range(5):
    opt.step(X, Y)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    opt.step(X, Y)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(int(os.environ.get("TOP", 35)))
