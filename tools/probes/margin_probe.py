"""How far is the kernels' decision variable d = logit_fg - logit_bg (up-sampled, at every output pixel) from the reference's?
The margin screen of the synthetic episodes (pemp_b200/episode_screen.json) is only as good as its threshold exceeds this
deviation.  Prints per workload: max / percentiles of |d_ours - d_ref| over all pixels of N raw (unscreened) episodes, the
flipped pixels with the reference margin at each, and the smallest threshold that would have screened every flip out.
    python tools/probes/margin_probe.py [--episodes 64]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import restate as O  # noqa: E402
from pemp_b200 import episodes as E, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=64)
a = ap.parse_args()
torch.set_num_threads(os.cpu_count() or 1)
for name, spec, stages in (("pemp S=1", E.EpisodeSpec(shot=1, stages=2), (1, 2)), ("pemp S=5", E.EpisodeSpec(shot=5, stages=2), (1, 2))):
    S, Q, c, h, w = spec.shot, spec.query, spec.channels, spec.h, spec.w
    dev_max, flips, devs = 0.0, [], []
    for i in range(a.episodes):
        b = E.make_batch(spec, [i])
        for st in stages:
            ctr = E.make_ctr(spec, st)
            want = O.pemp_head(b[f"feats{st}"], b["sup_mask"], ctr, 1, S, Q)
            f5 = b[f"feats{st}"].cuda().view(1, S + Q, c, h, w)
            low = ops.mask_nearest(b["sup_mask"].cuda().view(S, 2, spec.H, spec.W), h, w).view(S, 2, h * w)
            fgp, bgp, _ = ops.meta_proto_attn(f5[:, :S], ctr.cuda(), low[:, 0], low[:, 1], 1, S)
            pred = ops.cosine_match(f5[:, S:], fgp, bgp)["pred"].view(Q, 2, h, w)
            lg = ops.upsample_argmax(pred, (spec.H, spec.W), want_logits=True, want_mask8=False)["logits"].cpu()
            d_ref = (want["logits"][:, 1] - want["logits"][:, 0]).double()
            d_our = (lg[:, 1] - lg[:, 0]).double()
            dev = (d_our - d_ref).abs()
            devs.append(float(dev.max()))
            bad = (d_our > 0) != (d_ref > 0)
            for m in d_ref[bad].abs().tolist():
                flips.append(m)
    devs = np.array(devs)
    print(json.dumps({"workload": name, "heads_checked": len(devs), "pixels_per_head": spec.H * spec.W,
                      "max_abs_dev_of_decision_variable": float(devs.max()), "median_of_per_head_max": float(np.median(devs)),
                      "p90_of_per_head_max": float(np.percentile(devs, 90)), "flipped_pixels": len(flips),
                      "reference_margin_at_flips": sorted(flips)}), flush=True)
