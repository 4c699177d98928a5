"""Grouped DPT decoders (dpt.forward_grouped: the flow / depth / motion-mask heads of the shipped config as one launch
sequence with grouped weights) against the same adapters run one by one, at full size."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_grouped_dpt_equals_per_head():
    from l4p_b200 import weights
    from l4p_b200.models.task_heads.dpt import DPTOutputAdapter_fix, forward_grouped

    kw = dict(hooks=(14, 21, 28, 36), layer_dims=(256, 512, 1024, 1024), feature_dim=256, last_dim=128)
    ads = [DPTOutputAdapter_fix(num_channels=c, **kw).cuda() for c in (2, 1, 1)]
    for i, a in enumerate(ads):
        weights.fill_module_fast_(a, seed=10 + i)
    g = torch.Generator(device="cuda").manual_seed(3)
    taps = [torch.randn(2048, 1408, device="cuda", generator=g).half() for _ in range(4)]
    exp = [False, True, False]
    with torch.no_grad():
        single = [a(taps, 1, (16, 224, 224), exp_out=e) for a, e in zip(ads, exp)]
        grouped = forward_grouped(ads, taps, 1, (16, 224, 224), exp_outs=exp)
    torch.cuda.synchronize()
    for s, gr in zip(single, grouped):
        assert s.shape == gr.shape
        # split-K summation order (fp32 atomics) differs between a 1-head and a 3-head launch: round-off level
        err = (s - gr).norm() / s.norm()
        assert err < 2e-3, float(err)
