"""Golden vectors for the GPU-preprocessing row: the rgb branch of the UNMODIFIED reference dataset
pipeline (`L4PDataset.__getitem__`, l4p/data/l4p_dataset_mini.py:543-587, imported through oracle/ref_loader.py) on seeded
uint8 clips. Run in the build container:  python tests/golden/make_golden_preprocess.py
Only strided sub-samples and float64 checksums of the outputs are stored; inputs are regenerated from the seeds."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_loader  # noqa: E402

CASES = [(10, 135, 240, (256, 320), None), (24, 224, 224, (224, 224), None), (1, 97, 131, (240, 300), (16, 224, 224)),
         (5, 300, 300, None, (24, 224, 224)), (40, 180, 320, (224, 398), (32, 224, 224))]


def frames_for(case):
    T0, H0, W0 = case[:3]
    g = torch.Generator().manual_seed(T0 * 1000 + H0)
    return torch.randint(0, 256, (T0, H0, W0, 3), generator=g, dtype=torch.uint8)


def reference_item(frames, resize, crop):
    ref_loader.load()
    from l4p.data.l4p_dataset_mini import L4PData, L4PDataset

    class OneClip(L4PDataset):
        def __len__(self):
            return 1

        def getitem_helper(self, index):
            return L4PData(rgb_b3thw=frames.permute(3, 0, 1, 2).float() / 255.0, seq_name="clip", dataset_name="synthetic")

    return OneClip(crop_size=crop, center_crop=True, start_crop_time=True, resize_size=resize)[0]


def main():
    g = {}
    for i, case in enumerate(CASES):
        item = reference_item(frames_for(case), case[3], case[4])
        x = item["rgb_b3thw"]
        g[f"{i}/shape"] = torch.tensor(x.shape)
        g[f"{i}/sub"] = x[:, ::5, ::37, ::41].clone()
        g[f"{i}/sum"] = x.double().sum()
        g[f"{i}/ori_video_len"] = torch.tensor(item["ori_video_len"])
    torch.save(g, Path(__file__).resolve().parent / "golden_preprocess.pt")
    print("wrote golden_preprocess.pt", len(g), "entries")


if __name__ == "__main__":
    main()
