"""Where does the end-to-end leg of bench.py lose time against the device-resident leg? Times the same step with (a) resident
inputs, (b) + H2D from pinned memory through predict_step, (c) + D2H of the packed outputs on a copy stream, (d) both, and the
host-side enqueue time of one step."""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from l4p_b200 import weights  # noqa: E402
from l4p_b200.config import load_model  # noqa: E402

dev = torch.device("cuda", 0)
lit = load_model(device=dev, max_queries=bench.NQ + 1)
weights.fill_module_fast_(lit.l4p_model, seed=0)
host = {k: v.pin_memory() for k, v in bench.synth_batch(1).items()}
resident = {k: v.to(dev) for k, v in host.items()}
copy_stream = torch.cuda.Stream(device=dev)


def step(b):
    out = lit.predict_step(b, 0)
    return bench.pack_outputs(out, bench.OUT_KEYS, 1)


def run(name, h2d, d2h, steps=10):
    with torch.no_grad():
        r = step(dict(resident))
        pinned = [torch.empty(r.shape, dtype=torch.float32).pin_memory() for _ in range(2)]
        for _ in range(3):
            step(dict(resident))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            r = step(dict(host) if h2d else dict(resident))
            if d2h:
                done = torch.cuda.Event()
                done.record(main)
                copy_stream.wait_event(done)
                with torch.cuda.stream(copy_stream):
                    pinned[i % 2].copy_(r, non_blocking=True)
                    r.record_stream(copy_stream)
        main.wait_stream(copy_stream)
        e1.record()
        t_enq = time.perf_counter() - t0
        torch.cuda.synchronize()
    print(f"{name:28s}: {e0.elapsed_time(e1) / steps:7.2f} ms / step on the device, host enqueue {1e3 * t_enq / steps:6.2f} ms / step")


run("resident", False, False)
run("+ H2D (pinned, predict_step)", True, False)
run("+ D2H (copy stream)", False, True)
run("+ H2D + D2H", True, True)
run("resident (again)", False, False)
