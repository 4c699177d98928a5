"""TEST INFRASTRUCTURE ONLY (see oracle/l4p_oracle.py): CPU restatement of the rgb branch of the reference's dataset
pipeline, `l4p/data/l4p_dataset_mini.py`, with the same torch calls the reference makes.

Pinned: bit-exact against the UNMODIFIED reference `L4PDataset.__getitem__` (imported through oracle/ref_loader.py with a
kornia stub; the rgb key never touches it) on seeded uint8 clips - committed fixture tests/golden/golden_preprocess.pt
(tests/golden/make_golden_preprocess.py) and a live run when /root/reference is present (tests/test_oracle_golden.py)."""
from __future__ import annotations

from math import ceil
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

MEAN = torch.tensor([0.485, 0.456, 0.406])   # l4p_dataset_mini.py:103
STD = torch.tensor([0.229, 0.224, 0.225])    # l4p_dataset_mini.py:104


def mirror_and_pad(rgb_3thw: torch.Tensor) -> torch.Tensor:
    """l4p_dataset_mini.py:126-190 (rgb key): cat[x, flip_t(x)[1:]]."""
    return torch.cat([rgb_3thw, torch.flip(rgb_3thw, dims=[1])[:, 1:]], dim=1)


def preprocess(frames_u8_thw3: torch.Tensor, resize_size: Optional[Tuple[int, int]] = (224, 224),
               crop_size: Optional[Tuple[int, int, int]] = None, multiple: int = 8) -> torch.Tensor:
    """`__getitem__` (l4p_dataset_mini.py:543-587) for the rgb key with center_crop / start_crop_time.
    frames uint8 [T,H,W,3] -> normalised rgb_b3thw fp32 [1,3,T',Hc,Wc]."""
    x = frames_u8_thw3.permute(3, 0, 1, 2).float() / 255.0            # [3,T,H,W] in [0,1]
    T = x.shape[1]
    if crop_size is None:
        crop_size = (ceil(max(T, 16) / multiple) * multiple, 224, 224)  # :550-552
    if T == 1:
        x = x.repeat(1, crop_size[0], 1, 1)                            # repeat_single_frame (:192-235)
    else:
        while x.shape[1] < crop_size[0]:                               # :556-559
            x = mirror_and_pad(x)
    if resize_size is not None and (resize_size[0] != x.shape[2] or resize_size[1] != x.shape[3]):
        x = F.interpolate(x[None], (x.shape[1], resize_size[0], resize_size[1]), mode="trilinear")[0]   # :266
    _, Tn, H, W = x.shape
    i0, j0 = int((H - crop_size[1]) * 0.5), int((W - crop_size[2]) * 0.5)                             # :312-314
    x = x[:, 0:crop_size[0], i0:i0 + crop_size[1], j0:j0 + crop_size[2]]                              # :330, t0 = 0 (:310)
    x = (x - MEAN[:, None, None, None]) / STD[:, None, None, None]                                    # :576-580
    return x[None].contiguous()
