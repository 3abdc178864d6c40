"""BASELINE.json configs[4] / SURVEY 8d config 5: streaming reconstruction of the frames of ONE rank's shard of 1024 synthetic
frames (seed = frame id), batch in {1,2,4,8,16} frames per graph replay: frames/s and p50 / p99 in-pipeline latency
(input copy issued -> output written, CUDA events).  Under torchrun every rank streams its own contiguous shard.
Writes gpurun_out/stream_latency.json (rank 0)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cwfa_b200
from cwfa_b200.engine import CWFAEngine, StreamingReconstructor
from cwfa_b200.sharding import frame_seed, frame_shard

TOTAL = 1024
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
shards = int(os.environ.get("CWFA_SHARDS", "8" if world == 1 else str(world)))     # 1 GPU: stream the first of 8 shards (128 frames)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
a, b = frame_shard(TOTAL, rank if world > 1 else 0, shards)
model = cwfa_b200.CWFAModel(seed=0).to(dev)
eng = CWFAEngine(model, "bf16")
g = torch.Generator().manual_seed(1)
mv1 = [0.1 * torch.randn((1, 96 // 2 ** (n + 1), 512, 512), generator=g) for n in range(4)] + [0.1 * torch.randn((1, 6, 512, 512), generator=g)]
frames = [torch.randn((1, 29, 512, 512), generator=torch.Generator().manual_seed(frame_seed(f))).to(dev) for f in range(a, b)]
rows = []
for B in (1, 2, 4, 8, 16):
    mvs = [m.repeat(B, 1, 1, 1).to(dev) for m in mv1]
    batches = [torch.cat(frames[i:i + B]) for i in range(0, len(frames) - B + 1, B)]
    st = StreamingReconstructor(eng, tuple(batches[0].shape), mvs, depth=2)
    outs = [torch.empty((B, 96, 512, 512), device=dev) for _ in range(3)]
    st.run(batches[:3], outs[:3])                                   # warm-up
    lat = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    st.run(batches, [outs[i % 3] for i in range(len(batches))], latency_events=lat)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ls = sorted(s.elapsed_time(e) for s, e in lat)
    pct = lambda q: ls[min(len(ls) - 1, int(round(q * (len(ls) - 1))))]
    fps = len(batches) * B / ms * 1e3
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        fps = world * len(batches) * B / float(t) * 1e3
    rows.append(dict(batch=B, frames=len(batches) * B, n_gpus=world, frames_per_s=fps, latency_ms_p50=pct(0.5), latency_ms_p99=pct(0.99),
                     latency_ms_max=ls[-1]))
    if rank == 0:
        print(json.dumps(rows[-1]), flush=True)
    del st, outs, batches, mvs
    eng._graphs.clear()
    torch.cuda.empty_cache()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(dict(total_frames=TOTAL, shard=[a, b], rows=rows), open(f"gpurun_out/stream_latency_{world}gpu.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
