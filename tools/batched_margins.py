import sys, torch
sys.path.insert(0, ".")
from tests.util import grid_queries, rel_l2, synth_intrinsics, synth_rgb
from l4p_b200 import weights
from l4p_b200.config import load_model
lit = load_model(device=torch.device("cuda"), max_queries=17, compute_dtype=torch.float16)
model = lit.l4p_model
weights.fill_module_fast_(model, seed=0)
tasks = ["flow_2d_backward", "track_2d", "depth", "dyn_mask", "camray"]
rgb = synth_rgb(2, 16, seed=3); intr = synth_intrinsics(2, 16)
q = grid_queries(4).repeat(2, 1, 1); q[1, :, 1:] = 224.0 - q[1, :, 1:]
lab = torch.ones(2, q.shape[1])
batch = dict(rgb_b3thw=rgb.cuda(), intrinsics_b44t=intr.cuda(), track_2d_pointquerries_bn3=q.cuda(), track_2d_pointlabels_bn=lab.cuda())
keys = ["depth_est_b1thw", "flow_2d_backward_est_b2thw", "dyn_mask_est_b1thw", "track_2d_traj_est_bn2t", "track_2d_vis_est_bn1t", "track_2d_depth_est_bn1t"]
with torch.no_grad():
    for rep in range(3):
        both = {k: v.float().cpu() for k, v in model.forward(batch, tasks).items() if k in keys}
        single = []
        for b in range(2):
            one = model.forward({k: v[b:b + 1].contiguous() for k, v in batch.items()}, tasks)
            single.append({k: one[k].float().cpu() for k in keys})
        out = []
        for k in keys:
            ref = torch.cat([s[k] for s in single], dim=0)
            out.append(f"{k.split('_est')[0]}: {rel_l2(both[k], ref):.2e}/{(both[k]-ref).abs().max().item():.3g}")
        print(" | ".join(out))
