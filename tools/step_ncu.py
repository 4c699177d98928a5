"""One all-heads step inside a cudaProfiler range (for `ncu --profile-from-start off -k regex:...`)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from l4p_b200 import weights
from l4p_b200.config import load_model
dev = torch.device("cuda")
lit = load_model(device=dev, max_queries=bench.NQ + 1)
model = lit.l4p_model
weights.fill_module_fast_(model, seed=0)
model.parallel_heads = False
tasks = sys.argv[1].split(",") if len(sys.argv) > 1 else bench.TASKS
batch = {k: v.to(dev) for k, v in bench.synth_batch(1).items()}
with torch.no_grad():
    for _ in range(2):
        model.forward(batch, tasks)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model.forward(batch, tasks)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
