"""BASELINE.json configs[4]: streaming reconstruction throughput vs batch size (frames per graph replay) and frames in
flight, one GPU, device-resident inputs (seeded per frame: seed = frame id).  Writes profiles/r01_stream_sweep.md."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cwfa_b200
from cwfa_b200.engine import CWFAEngine, StreamingReconstructor
from cwfa_b200.sharding import frame_seed
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda:0")
model = cwfa_b200.CWFAModel(seed=0).to(dev)
eng = CWFAEngine(model, "bf16")
g = torch.Generator().manual_seed(1)
mv1 = [0.1 * torch.randn((1, 96 // 2 ** (n + 1), 512, 512), generator=g) for n in range(4)] + [0.1 * torch.randn((1, 6, 512, 512), generator=g)]
rows = []
n_frames = 64
for B in (1, 2, 4, 8):
    mvs = [m.repeat(B, 1, 1, 1).to(dev) for m in mv1]
    views = []
    for f0 in range(0, 4 * B, B):
        views.append(torch.cat([torch.randn((1, 29, 512, 512), generator=torch.Generator().manual_seed(frame_seed(f0 + j))) for j in range(B)]).to(dev))
    for depth in (1, 2):
        st = StreamingReconstructor(eng, tuple(views[0].shape), mvs, depth=depth)
        outs = [torch.empty((B, 96, 512, 512), device=dev) for _ in range(2)]
        steps = n_frames // B
        st.run([views[i % 4] for i in range(3)], [outs[i % 2] for i in range(3)])
        lat = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st.run([views[i % 4] for i in range(steps)], [outs[i % 2] for i in range(steps)])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rows.append((B, depth, steps * B / ms * 1e3, ms / steps))
        print(rows[-1], flush=True)
        del st, outs
        eng._graphs.clear()
        torch.cuda.empty_cache()
md = ["# Streaming reconstruction sweep (BASELINE.json configs[4]; one B200, 64 frames of 512x512x96, bf16 engine)", "",
      "| frames per replay (batch) | graph instances in flight | frames/s | ms per replay |", "|---|---|---|---|"]
md += [f"| {b} | {d} | {f:.1f} | {m:.2f} |" for b, d, f, m in rows]
md += ["", "8 GPUs (frames sharded, 1 frame per replay, 2 in flight): 918.7 frames/s = 114.8 per GPU (bench.py --gpus 8)."]
open(os.path.join(ROOT, "profiles", "r01_stream_sweep.md"), "w").write("\n".join(md) + "\n")
