"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference (/root/reference,
imported through oracle/ref_loader.py) on seeded synthetic weights and inputs. Run in the build container:

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors; these fixtures are the pins for oracle/l4p_oracle.py.
Inputs are regenerated from seeds by the tests, only reference OUTPUTS (small) are stored.
"""
import json
import sys
from functools import partial
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from l4p_b200 import weights  # noqa: E402  (deterministic generator keyed on state-dict names)
from oracle import ref_loader  # noqa: E402

OUT = Path(__file__).resolve().parent
ref_loader.load()
from l4p.models.l4p_videomae import VideoMAEEncoder  # noqa: E402
from l4p.models.task_heads.dense_heads import VideoMAEDepthDPTHead, VideoMAEFlowDPTHead, VideoMAETraj3DDPTHead  # noqa: E402
from l4p.models.task_heads.sparse_heads import VideoMAETrack2DSamHead  # noqa: E402
from l4p.models.aligner import LstSqAffineAligner  # noqa: E402
from l4p.utils import geometry_utils as RG  # noqa: E402
from l4p.models.VideoMAEv2.models.modeling_finetune import Block, get_sinusoid_encoding_table  # noqa: E402

TINY = dict(img=56, T=4, dim=64, depth=3, heads=4)


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


@torch.no_grad()
def main():
    g = {}
    # ---- (a) tiny encoder -------------------------------------------------------------------------
    enc = VideoMAEEncoder(img_size=TINY["img"], patch_size=14, embed_dim=TINY["dim"], depth=TINY["depth"],
                          num_heads=TINY["heads"], mlp_ratio=4, qkv_bias=True,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), init_values=0.0, tubelet_size=2,
                          all_frames=TINY["T"]).eval()
    weights.fill_module_(enc, seed=11)
    rgb = rnd((1, 3, TINY["T"], TINY["img"], TINY["img"]), 12)
    feats = enc(rgb)
    g["enc_feats"] = torch.stack(feats)
    # ---- (b) tiny DPT heads on those features, incl. 3-window stitching -----------------------------
    hooks = [1, 2, 3, 3]
    depth = VideoMAEDepthDPTHead("depth", depth=3, embed_dim=TINY["dim"], depth_fn="exp", hooks_idx=hooks,
                                 align_window_overlap_fn="inverse").eval()
    weights.fill_module_(depth, seed=13)
    g["depth_single"] = depth.forward(feats, img_info=(4, 56, 56))["depth_est_b1thw"]
    flow = VideoMAEFlowDPTHead("flow_2d_backward", out_nchan=2, depth=3, embed_dim=TINY["dim"], hooks_idx=hooks).eval()
    weights.fill_module_(flow, seed=14)
    cam = VideoMAETraj3DDPTHead("traj3d", depth=3, embed_dim=TINY["dim"], hooks_idx=hooks, output_size=(4, 4, 4),
                                use_intrinsics=False, fixed_intrinsics=True).eval()
    weights.fill_module_(cam, seed=15)
    g["cam_rays"] = cam.task_head(feats, (4, 56, 56))
    rgb_long = rnd((1, 3, 8, 56, 56), 16)
    starts = torch.arange(0, 8 - 4 + 1, 2)
    feats2d = [enc(rgb_long[:, :, s:s + 4]) for s in starts]
    intr = torch.eye(4)[None, :, :, None].repeat(1, 1, 1, 8)
    g["depth_windowed"] = depth.forward_windowed(feats2d, img_info=(4, 56, 56), time_strides=starts, intrinsics_b44t=intr)["depth_est_b1thw"]
    g["flow_windowed"] = flow.forward_windowed(feats2d, img_info=(4, 56, 56), time_strides=starts, intrinsics_b44t=intr)["flow_2d_backward_est_b2thw"]
    # ---- (c) tiny track head: single window and 3-window memory tracker ------------------------------
    trk = VideoMAETrack2DSamHead(task_name="track_2d", prompt_embed_dim=TINY["dim"], image_size=(4, 56, 56),
                                 estimate_vis=True, estimate_depth=True, sam_head_depth=2, num_point_embeddings=2,
                                 modify_pointlabels_for_windowing=True, prompt_using_features=True, attend_to_past=True,
                                 estimation_directions=[1], depth_fn="exp", vis_fn="linear").eval()
    weights.fill_module_(trk, seed=17)
    q = torch.tensor([[[0.5, 10.5, 12.5], [1.5, 40.5, 30.5], [0.5, 28.0, 28.0], [5.5, 20.5, 44.5]]])
    lab = torch.ones(1, 4)
    o = trk.forward(feats, q[:, :3], lab[:, :3])
    for k in ("track_2d_traj_est_bn2t", "track_2d_vis_est_bn1t", "track_2d_depth_est_bn1t", "track_2d_prompt_features_bnc",
              "track_2d_enc_features_with_track_history_bnpc"):
        g["trk_single/" + k] = o[k]
    ow = trk.forward_windowed(feats2d, q, lab, time_strides=starts)
    for k, v in ow.items():
        g["trk_windowed/" + k] = v
    # ---- (d) geometry / aligner known answers ----------------------------------------------------------
    T = 6
    K = torch.eye(4)[None, :, :, None].repeat(1, 1, 1, T)
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2] = 0.9, 1.1, 0.5, 0.5  # normalised intrinsics
    gen = torch.Generator().manual_seed(21)
    ext = torch.zeros(1, 4, 4, T)
    for t in range(T):
        a = torch.randn(3, 3, generator=gen)
        qm, _ = torch.linalg.qr(a)
        if torch.linalg.det(qm) < 0:
            qm[:, 0] = -qm[:, 0]
        if t == 0:
            qm = torch.eye(3)
        ext[0, :3, :3, t] = qm
        ext[0, :3, 3, t] = torch.randn(3, generator=gen) * (0.0 if t == 0 else 0.5)
        ext[0, 3, 3, t] = 1
    rays, _ = RG.get_rays_plucker(K, ext, (16, 16))
    g["geo_rays"] = rays
    ext_rec, ctr = RG.rays_to_cameras(rays, K)
    g["geo_ext_from_rays"], g["geo_centers"] = ext_rec, ctr
    ext2, _, kest = RG.rays_to_cameras_and_intrinsics(rays, reproj_threshold=0.2, output_size=(224, 224), fixed_intrinsics=True)
    g["geo_ext_fixed_k"], g["geo_kest"] = ext2, kest
    g["geo_K"], g["geo_ext"] = K, ext
    x = torch.rand(1, 1, 8, 20, 20, generator=gen) + 0.5
    al = LstSqAffineAligner(pre_post_fn="inverse")
    y = 1.0 / (2.0 * (1.0 / x) + 0.3)
    al.solve(x, y, None, None)
    g["affine_sol"] = al.sol
    g["affine_apply"] = al.apply(x)
    # ---- (e) full-size pieces: one ViT-giant block, position table, state-dict manifest ------------------
    blk = Block(dim=1408, num_heads=16, mlp_ratio=48 / 11, qkv_bias=True, init_values=0.0,
                norm_layer=partial(torch.nn.LayerNorm, eps=1e-6)).eval()
    weights.fill_module_(blk, seed=31)
    xin = rnd((1, 2048, 1408), 32)
    g["block_out_sub"] = blk(xin)[:, ::97, ::13].clone()
    tab = get_sinusoid_encoding_table(2048, 1408)
    g["pos_sub"] = tab[:, ::31, ::7].clone()
    g["pos_sum"] = tab.double().sum()
    torch.save({k: v.clone() for k, v in g.items()}, OUT / "golden_small.pt")

    import yaml
    cfg = yaml.safe_load(open(Path(ref_loader.REFERENCE_ROOT) / "configs" / "model.yaml"))
    model = ref_loader.instantiate(cfg)
    man = {k: list(v.shape) for k, v in model.state_dict().items()}
    json.dump(man, open(OUT / "state_dict_manifest.json", "w"))
    print("wrote", OUT / "golden_small.pt", len(g), "tensors;", len(man), "state-dict keys;",
          sum(p.numel() for p in model.parameters()), "params")


if __name__ == "__main__":
    main()
