"""N3: GPU video preprocessing for the hot path - the rgb branch of the reference's dataset pipeline
(`l4p/data/l4p_dataset_mini.py:543-587`): temporal mirror padding, spatial resize, crop, ImageNet normalisation,
fused into ONE gather kernel (`l4p_preprocess_rgb`) that reads the decoder's uint8 frames and writes `rgb_b3thw`.

Only inference-time behaviour is mirrored (`center_crop=True`, `start_crop_time=True`: deterministic crop origin,
`demo/demo.py`); the training-time random crops, flow / depth / track keys and augmentations are out of scope."""
from __future__ import annotations

import ctypes as C
from math import ceil
from typing import Dict, Optional, Tuple

import torch

from . import lib as _l
from .ops import _count, _dev_init, _stream

INPUT_MEAN = (0.485, 0.456, 0.406)   # l4p_dataset_mini.py:103
INPUT_STD = (0.229, 0.224, 0.225)    # l4p_dataset_mini.py:104
DEFAULT_SAMPLE_SIZE = (16, 224, 224)


def plan(T0: int, H0: int, W0: int, resize_size: Optional[Tuple[int, int]], crop_size: Optional[Tuple[int, int, int]],
         length_multiply_of: int = 8) -> Dict[str, int]:
    """Shapes and crop origin the reference's `__getitem__` arrives at for a T0 x H0 x W0 video
    (l4p_dataset_mini.py:547-567, 292-318 with center_crop and start_crop_time)."""
    if crop_size is None:
        T_new = ceil(max(T0, DEFAULT_SAMPLE_SIZE[0]) / length_multiply_of) * length_multiply_of
        crop_size = (T_new,) + DEFAULT_SAMPLE_SIZE[1:]
    To, Hc, Wc = crop_size
    T_pad = T0
    if T0 == 1:
        T_pad = To                      # repeat_single_frame
    else:
        while T_pad < To:               # mirror_and_pad: T -> 2T - 1
            T_pad = 2 * T_pad - 1
    Hs, Ws = (H0, W0) if resize_size is None else resize_size
    if Hs < Hc or Ws < Wc or T_pad < To:
        raise AssertionError(f"Cropping Error: diff_shape {[T_pad - To, Hs - Hc, Ws - Wc]}")
    return dict(To=To, Hc=Hc, Wc=Wc, Hs=Hs, Ws=Ws, t0=0, i0=int((Hs - Hc) * 0.5), j0=int((Ws - Wc) * 0.5), T_pad=T_pad)


def preprocess_video(frames_u8: torch.Tensor, resize_size: Optional[Tuple[int, int]] = (224, 224),
                     crop_size: Optional[Tuple[int, int, int]] = None, length_multiply_of: int = 8) -> Dict[str, torch.Tensor]:
    """frames_u8: CUDA uint8 [T0, H0, W0, 3] (decoder layout). Returns the reference's batch entries for the rgb key:
    `rgb_b3thw` fp32 [1,3,T,Hc,Wc] (normalised), `rgb_mean_b3111`, `rgb_std_b3111`, `ori_video_len`."""
    if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
        raise _l.L4PError(f"preprocess_video: expected uint8 [T,H,W,3], got {frames_u8.dtype} {tuple(frames_u8.shape)}")
    _dev_init(frames_u8)
    frames_u8 = frames_u8.contiguous()
    T0, H0, W0, _ = frames_u8.shape
    pl = plan(T0, H0, W0, resize_size, crop_size, length_multiply_of)
    out = torch.empty(1, 3, pl["To"], pl["Hc"], pl["Wc"], device=frames_u8.device, dtype=torch.float32)
    mean = (C.c_float * 3)(*INPUT_MEAN)
    std = (C.c_float * 3)(*INPUT_STD)
    _l.check(_l.load().l4p_preprocess_rgb(frames_u8.data_ptr(), out.data_ptr(), T0, H0, W0, pl["Hs"], pl["Ws"], pl["t0"],
                                          pl["i0"], pl["j0"], pl["To"], pl["Hc"], pl["Wc"], C.cast(mean, C.c_void_p),
                                          C.cast(std, C.c_void_p), _stream()), "l4p_preprocess_rgb")
    _count()
    dev = frames_u8.device
    return {"rgb_b3thw": out,
            "rgb_mean_b3111": torch.tensor(INPUT_MEAN, device=dev).view(1, 3, 1, 1, 1),
            "rgb_std_b3111": torch.tensor(INPUT_STD, device=dev).view(1, 3, 1, 1, 1),
            "ori_video_len": T0}
