"""Packed weight arena (l4p_b200/arena.py) on CPU tensors with the kernels' CPU stand-in: packing and
releasing the fp32 masters must not change any output; all operands end up inside one buffer."""
import torch

from tests import emu
from tests.test_host_emulated import IMG, _tiny_model, rnd


def test_pack_and_release_keeps_outputs(monkeypatch):
    from l4p_b200.arena import pack_weights

    emu.install(monkeypatch)
    tasks = ["depth", "flow_2d_backward"]
    model = _tiny_model(tasks)
    data = dict(rgb_b3thw=rnd((1, 3, 8, 56, 56), 16), intrinsics_b44t=None, img_info=IMG)
    want = model.forward(dict(data), tasks)
    info = pack_weights(model, torch.device("cpu"), release_masters=True)
    lo, hi = info.buffer.data_ptr(), info.buffer.data_ptr() + info.total_bytes
    n = 0
    for m in model.modules():
        pk = getattr(m, "_packed", None)
        if isinstance(pk, dict):
            stack = [pk]
            while stack:
                x = stack.pop()
                if torch.is_tensor(x):
                    assert lo <= x.data_ptr() < hi, "an operand was left outside the arena"
                    n += 1
                elif isinstance(x, dict):
                    stack.extend(x.values())
                elif isinstance(x, (list, tuple)):
                    stack.extend(x)
    assert n >= info.tensors > 20 and info.masters_released_bytes > 0
    got = model.forward(dict(data), tasks)
    for k in ("depth_est_b1thw", "flow_2d_backward_est_b2thw"):
        # same operands from a different address: the CPU stand-in's convolutions may take another vector path (round-off);
        # on the device the kernels are address-agnostic and tests/test_ckpt_gpu.py asserts bit equality on the encoder
        assert (got[k] - want[k]).abs().max() <= 1e-5 * want[k].abs().max(), k
