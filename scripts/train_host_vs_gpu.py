"""Is the flow-level training step host-bound?  Host time of the eager step() call (no sync inside) against the CUDA-event time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cwfa_b200
from cwfa_b200.training import FlowLevelTrainer

dev = "cuda:0"
S, D = 512, 96
model = cwfa_b200.CWFAModel(n_depths=D, volume_side_size=S, INN_max_down_steps=2, seed=0).to(dev)
tr = FlowLevelTrainer(model, 0, precision="bf16")
g = torch.Generator().manual_seed(1)
mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
gt, views, mv, vin = mk(1, D, S, S), mk(1, 29, S, S), mk(1, D // 2, S, S, sc=0.1), mk(1, D // 2, S, S)
for _ in range(3):
    tr.step(gt, views, mv, vin)
torch.cuda.synchronize()
n = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    tr.step(gt, views, mv, vin)
e1.record()
t_host = (time.perf_counter() - t0) / n * 1e3
torch.cuda.synchronize()
print(f"eager step: host {t_host:.1f} ms per call (returns before the GPU is done), GPU timeline {e0.elapsed_time(e1) / n:.1f} ms per step")

# the same step as ONE CUDA-graph replay (forward + backward + Lion captured after eager warm-up): pure GPU time
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        tr.step(gt, views, mv, vin)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    parts = tr.step(gt, views, mv, vin)
torch.cuda.synchronize()
for _ in range(2):
    graph.replay()
torch.cuda.synchronize()
l0 = float(parts["loss"])
e0.record()
for _ in range(n):
    graph.replay()
e1.record()
torch.cuda.synchronize()
print(f"graphed step: {e0.elapsed_time(e1) / n:.1f} ms per step; loss {l0:.6f} -> {float(parts['loss']):.6f} (keeps training under replay)")
