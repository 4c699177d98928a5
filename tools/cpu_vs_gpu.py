"""Is the step GPU-bound or launch-bound? Host time to enqueue one all-heads step vs its device time."""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
from l4p_b200 import weights
from l4p_b200.config import load_model
dev = torch.device("cuda")
lit = load_model(device=dev, max_queries=bench.NQ + 1)
model = lit.l4p_model
weights.fill_module_fast_(model, seed=0)
batch = {k: v.to(dev) for k, v in bench.synth_batch(1).items()}
with torch.no_grad():
    for _ in range(3):
        bench.run_clip(model, batch, 0)
    torch.cuda.synchronize()
    host, devt = [], []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        bench.run_clip(model, batch, 0)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        host.append((t1 - t0) * 1e3); devt.append(e0.elapsed_time(e1))
        print(f"host enqueue {host[-1]:.2f} ms, device {devt[-1]:.2f} ms, host wait after enqueue {(t2 - t1) * 1e3:.2f} ms")
    # encoder only
    enc = model.video_encoder if hasattr(model, "video_encoder") else None
    rgb = batch["rgb_b3thw"]
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); model.encode_features({"rgb_b3thw": rgb}) if hasattr(model, "encode_features") else None; e1.record()
        t1 = time.perf_counter(); torch.cuda.synchronize()
        print(f"encoder: host enqueue {(t1 - t0) * 1e3:.2f} ms, device {e0.elapsed_time(e1):.2f} ms")
