"""BASELINE.json configs[4]: bf16 vs fp16 encoder sweep at batch 8 (8 x 16x224x224 windows through the 40-block
ViT-giant encoder) + the attention kernel at the same batch. CUDA-event timed, weights/activations exceed L2.
  python tools/encoder_sweep.py [B]"""
import sys
from functools import partial

import torch

sys.path.insert(0, ".")
from l4p_b200 import ops, weights  # noqa: E402
from l4p_b200.models.videomae import VideoMAEEncoder  # noqa: E402
from tests.util import synth_rgb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ENC_FLOPS_PER_WINDOW = 5085.58e9
ATT_FLOPS = 4 * 2048 * 2048 * 88 * 16
dev = torch.device("cuda")
rgb = synth_rgb(B, 16).to(dev)
for dt in (torch.float16, torch.bfloat16):
    enc = VideoMAEEncoder(img_size=224, patch_size=14, embed_dim=1408, depth=40, num_heads=16, mlp_ratio=48 / 11,
                          qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), init_values=0.0,
                          tubelet_size=2, all_frames=16).to(dev)
    weights.fill_module_fast_(enc, seed=0)
    enc.compute_dtype = dt
    with torch.no_grad():
        for _ in range(3):
            enc(rgb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            enc(rgb)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"encoder B={B} {str(dt).split('.')[-1]}: {ms:.2f} ms/pass, {16 * B / ms * 1e3:.0f} frames/s, {ENC_FLOPS_PER_WINDOW * B / ms / 1e9:.0f} TFLOP/s (algorithmic)")
    # the same pass replayed from a CUDA graph (no Python enqueue between the 280+ kernels)
    from l4p_b200.graph import StepGraph  # noqa: E402
    l0 = ops.LAUNCHES
    g = StepGraph(lambda b: {"last": enc(b["rgb"])[40]}, {"rgb": rgb}, dev, warmup=1)
    with torch.no_grad():
        for _ in range(2):
            g({"rgb": rgb})
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            g({"rgb": rgb})
        e1.record()
        torch.cuda.synchronize()
    msg = e0.elapsed_time(e1) / 5
    print(f"encoder B={B} {str(dt).split('.')[-1]} CUDA graph ({g.launches} launches per pass): {msg:.2f} ms/pass, "
          f"{ENC_FLOPS_PER_WINDOW * B / msg / 1e9:.0f} TFLOP/s (algorithmic)")
    del g
    q = torch.randn(B, 16, 2048, 96, device=dev, dtype=dt); q[..., 88:] = 0
    k = torch.randn_like(q); k[..., 88:] = 0
    vt = torch.randn(B, 16, 96, 2048, device=dev, dtype=dt); vt[:, :, 88:] = 0
    o = torch.empty(B * 2048, 1408, device=dev, dtype=dt)
    for _ in range(3):
        ops.attention(q, k, vt, o, 88, 88 ** -0.5)
    e0.record()
    for _ in range(20):
        ops.attention(q, k, vt, o, 88, 88 ** -0.5)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"attention B={B} {str(dt).split('.')[-1]}: {us:.1f} us/launch, {ATT_FLOPS * B / us / 1e6:.0f} TFLOP/s (algorithmic, d=88)")
    del enc
