"""Run ONE op of the library a few times at its BASELINE shape (the short command to put under ncu).
    python tools/prof_op.py --op K6|K7|K11|K9|K9x3|K2|K3|K2s1|K2b|K3b [--iters 3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pemp_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--op", required=True)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--B", type=int, default=64)
a = ap.parse_args()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B, S, c, h, H = a.B, 5, 512, 51, 401
hw = h * h
if a.op in ("K6", "K7", "K2", "K3", "K2s1", "K4", "K4h", "K2b", "K3b"):
    if a.op == "K2s1":
        S = 1
    sup = torch.randn(B * S, c, hw, device=dev, generator=g) * 0.5
    qry = torch.randn(B, c, hw, device=dev, generator=g) * 0.5
    ctr = torch.rand(c, 6, device=dev, generator=g)
    fgfull = torch.zeros(B * S, 1, H, H, device=dev)
    fgfull[:, :, 100:300, 80:330] = 1
    sup_mask = torch.cat((fgfull, 1 - fgfull), 1)
    low = ops.mask_nearest(sup_mask, h, h).view(B * S, 2, hw)
    fgp, bgp, _ = ops.meta_proto_attn(sup, ctr, low[:, 0], low[:, 1], B, S)
    pred = ops.cosine_match(qry, fgp, bgp)["pred"].view(B, 2, h, h)
    if a.op in ("K2b", "K3b"):          # backward kernels of the training path (K12)
        _, _, saved = ops.meta_proto_attn_train(sup, ctr, low[:, 0], low[:, 1], B, S)
        gfp, gbp = torch.randn(B, c, 3, device=dev, generator=g), torch.randn(B, c, 3, device=dev, generator=g)
        gpred = torch.randn(B, 2, hw, device=dev, generator=g)
    fn = {"K2b": lambda: ops.meta_proto_attn_bwd(saved, gfp, gbp, B, S),
          "K3b": lambda: ops.cosine_match_bwd(qry, fgp, bgp, gpred),
          "K6": lambda: ops.map_pool_fullres(sup.view(B * S, c, h, h), sup_mask, B, S),
          "K7": lambda: ops.panet_align(qry.view(B, c, h, h), pred, sup.view(B * S, c, h, h), fgfull, 1),
          "K2": lambda: ops.meta_proto_attn(sup, ctr, low[:, 0], low[:, 1], B, S),
          "K2s1": lambda: ops.meta_proto_attn(sup, ctr, low[:, 0], low[:, 1], B, S),
          "K3": lambda: ops.cosine_match(qry, fgp, bgp),
          "K4": lambda: ops.upsample_argmax(pred, (H, H), want_mask8=True),
          "K4h": lambda: ops.upsample_argmax_hist(pred, (H, H), (torch.rand(B, H, H, device=dev) > 0.5).to(torch.uint8),
                                                  torch.ones(B, dtype=torch.int64, device=dev),
                                                  torch.zeros(21, 3, dtype=torch.int64, device=dev))}[a.op]
elif a.op == "K11":
    Nc, cc, hc = 16 * 6, 256, 101
    xc = torch.randn(Nc, cc, hc, hc, device=dev, generator=g)
    mc = (torch.rand(Nc, 1, hc, hc, device=dev, generator=g) > 0.7).float()
    wc, bc = torch.randn(2, 2 * cc, device=dev, generator=g), torch.randn(2, device=dev, generator=g)
    fn = lambda: ops.comm_module(xc, mc, wc, bc, 6, 1)
elif a.op in ("K9", "K9x3"):
    Bp, Sp, Cp, sp = 1, 5, 2048, 60
    q4 = torch.relu(torch.randn(Bp, Cp, sp, sp, device=dev, generator=g))
    s4 = torch.relu(torch.randn(Sp, Bp, Cp, sp, sp, device=dev, generator=g))
    sm = (torch.rand(Sp, Bp, sp, sp, device=dev, generator=g) > 0.5).float()
    prec = ops.PRIOR_BF16 if a.op == "K9" else ops.PRIOR_BF16X3
    fn = lambda: ops.prior_mask(q4, s4, sm, precision=prec)
else:
    raise SystemExit("unknown op")
for _ in range(a.iters):
    fn()
torch.cuda.synchronize()
print("ok", a.op)
