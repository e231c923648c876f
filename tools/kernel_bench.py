"""Per-kernel device timing at the BASELINE shapes (CUDA events, inputs > L2).  Developer tool:
    python tools/kernel_bench.py [--B 64] [--S 5]
Prints one JSON line per kernel: ms, algorithmic GB, GB/s, fraction of the measured HBM peak."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pemp_b200 import ops  # noqa: E402


def peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--S", type=int, default=5)
    ap.add_argument("--c", type=int, default=512)
    ap.add_argument("--hw", type=int, default=51)
    ap.add_argument("--H", type=int, default=401)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    B, S, c, h, H = a.B, a.S, a.c, a.hw, a.H
    hw = h * h
    pk, how = peak()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    sup = torch.randn(B * S, c, hw, device=dev, generator=g) * 0.5
    qry = torch.randn(B, c, hw, device=dev, generator=g) * 0.5
    ctr = torch.rand(c, 6, device=dev, generator=g)
    fgfull = torch.zeros(B * S, 1, H, H, device=dev)
    fgfull[:, :, 100:300, 80:330] = 1
    sup_mask = torch.cat((fgfull, 1 - fgfull), 1)
    low = ops.mask_nearest(sup_mask, h, h).view(B * S, 2, hw)
    fgp, bgp, _ = ops.meta_proto_attn(sup, ctr, low[:, 0], low[:, 1], B, S)
    pred = ops.cosine_match(qry, fgp, bgp)["pred"].view(B, 2, h, h)
    m8 = ops.upsample_argmax(pred, (H, H))["mask8"]
    ref = (torch.rand(B, H, H, device=dev) > 0.5).to(torch.uint8)
    cls = torch.randint(1, 6, (B,), device=dev)
    stat = torch.zeros(21, 3, dtype=torch.int64, device=dev)
    feats_all = torch.cat((sup.view(B, S, c, h, h), qry.view(B, 1, c, h, h)), dim=1).view(B * (S + 1), c, h, h).contiguous()
    rows = [
        ("K0 mask_nearest", lambda: ops.mask_nearest(sup_mask, h, h), B * S * 2 * hw * 4 * 2),
        ("K2 meta_proto_attn", lambda: ops.meta_proto_attn(sup, ctr, low[:, 0], low[:, 1], B, S),
         B * (S * (c * hw + 2 * hw) * 4 + 2 * c * 6 * 4)),
        ("K1 map_pool_lowres", lambda: ops.map_pool_lowres(sup, low[:, 0], low[:, 1], B, S),
         B * (S * (c * hw + 2 * hw) * 4 + 2 * c * 4)),
        ("K3 cosine_match", lambda: ops.cosine_match(qry, fgp, bgp), B * (c * hw * 4 + c * 6 * 4 + 2 * hw * 4)),
        ("K4 upsample_argmax", lambda: ops.upsample_argmax(pred, (H, H)), B * (2 * hw * 4 + H * H)),
        ("K10 iou_hist", lambda: ops.iou_hist(m8, ref, cls, stat), B * 2 * H * H),
        ("K4+K10 upsample_argmax_hist", lambda: ops.upsample_argmax_hist(pred, (H, H), ref.view(B, H, H), cls, stat),
         B * (2 * hw * 4 + 2 * H * H)),
        ("K6 map_pool_fullres", lambda: ops.map_pool_fullres(sup.view(B * S, c, h, h), sup_mask, B, S),
         B * (S * (c * hw * 4 + 2 * H * H * 4) + 2 * c * 4)),
        # PANet alignLoss (panet.py:158-194): query pooling + S cosine passes over the support maps + fused up-sample / CE
        ("K7 panet_align", lambda: ops.panet_align(qry.view(B, c, h, h), pred, sup.view(B * S, c, h, h), fgfull, 1),
         B * ((1 + S) * c * hw * 4 + 2 * hw * 4 + S * H * H * 4)),
        ("K8 weighted_gap", lambda: ops.weighted_gap(sup.view(B * S, c, h, h), low[:, 0].reshape(B * S, 1, h, h)),
         B * S * (c * hw + hw) * 4),
        # CaNet dense-comparison input (canet.py:172-180): K0 + K1 (fg only) over the supports, then one read of the query maps
        # and one write of the 2c-channel tensor
        ("K15 canet_map_tile", lambda: ops.canet_map_tile(feats_all, sup_mask.view(B, S, 2, H, H), B, S, 1),
         B * (S * (c * hw + hw) * 4 + S * 2 * hw * 8 + c * hw * 4 + 2 * c * hw * 4)),   # K0 samples hw of the H*W mask pixels
    ]
    if not a.only or "K11" in a.only:
        # ResNetCM.comm call site 2 (backbones.py:235): x2 [B*6, 256, 101, 101], stride 1
        Nc, cc, hc = 16 * 6, 256, 101
        xc = torch.randn(Nc, cc, hc, hc, device=dev, generator=g)
        mc = (torch.rand(Nc, 1, hc, hc, device=dev, generator=g) > 0.7).float()
        wc, bc = torch.randn(2, 2 * cc, device=dev, generator=g), torch.randn(2, device=dev, generator=g)
        ms = timeit(lambda: ops.comm_module(xc, mc, wc, bc, 6, 1))
        nb = Nc * (cc * hc * hc + 2 * hc * hc + 2 * hc * hc) * 4
        print(json.dumps({"kernel": "K11 comm_module (maxpool + pool + linear + expand)", "ms": round(ms, 4), "alg_GB": round(nb / 1e9, 4),
                          "GBps": round(nb / ms / 1e6, 1), "frac_of_hbm_peak": round(nb / ms / 1e6 / pk, 3), "peak": how, "N": Nc, "c": cc, "hw": hc * hc}))
    if not a.only or "K9" in a.only:
        Bp, Sp, Cp, spx = 1, 5, 2048, 60
        q4 = torch.relu(torch.randn(Bp, Cp, spx, spx, device=dev, generator=g))
        s4 = torch.relu(torch.randn(Sp, Bp, Cp, spx, spx, device=dev, generator=g))
        sm = (torch.rand(Sp, Bp, spx, spx, device=dev, generator=g) > 0.5).float()
        flop = 2.0 * Cp * (Sp * spx * spx) * (spx * spx) * Bp
        for prec, tag in ((0, "bf16"), (2, "bf16x3"), (1, "fp32")):
            ms = timeit(lambda: ops.prior_mask(q4, s4, sm, precision=prec), iters=5, warm=2)
            print(json.dumps({"kernel": f"K9 prior_mask {tag} (whole op incl. pre-pass + tail)", "ms": round(ms, 4),
                              "alg_GFLOP": round(flop / 1e9, 1), "TFLOPs": round(flop / ms / 1e9, 1), "S": Sp, "C": Cp, "hw": spx * spx}))
    for name, fn, nbytes in rows:
        if a.only and a.only not in name:
            continue
        ms = timeit(fn)
        gbs = nbytes / ms / 1e6
        print(json.dumps({"kernel": name, "ms": round(ms, 4), "alg_GB": round(nbytes / 1e9, 4), "GBps": round(gbs, 1),
                          "frac_of_hbm_peak": round(gbs / pk, 3), "peak": how, "B": B, "S": S}))


if __name__ == "__main__":
    main()
