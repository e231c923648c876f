"""Per-kernel device timing at the BASELINE shapes (CUDA events, inputs > L2).
    python tools/kernel_bench.py [--B 64] [--S 5]
prints one JSON line per kernel: ms, algorithmic GB (or GFLOP), GB/s (TFLOP/s), fraction of the measured peak.
`run(...)` returns the same rows as dicts: `bench.py` attaches them to its JSON line as `roofline_extra`, so the kernels that
are not on the headline workload (K1, K6, K7, K8, K9, K11, K12, K15) are measured in the driver's own run as well."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pemp_b200 import ops  # noqa: E402


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1650.0, "fallback (B200_PROFILING.md)"


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


def run(B=64, S=5, c=512, h=51, H=401, only="", k9_batch=2, emit=None):
    hw = h * h
    pk, tpk, how = peaks()
    out = []

    def add(row):
        out.append(row)
        if emit:
            emit(row)

    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    sup = torch.randn(B * S, c, hw, device=dev, generator=g) * 0.5
    qry = torch.randn(B, c, hw, device=dev, generator=g) * 0.5
    ctr = torch.rand(c, 6, device=dev, generator=g)
    fgfull = torch.zeros(B * S, 1, H, H, device=dev)
    fgfull[:, :, 100:300, 80:330] = 1
    sup_mask = torch.cat((fgfull, 1 - fgfull), 1)
    lab8 = fgfull[:, 0].to(torch.uint8)
    low = ops.mask_nearest(sup_mask, h, h).view(B * S, 2, hw)
    fgp, bgp, _ = ops.meta_proto_attn(sup, ctr, low[:, 0], low[:, 1], B, S)
    fg1, bg1 = ops.map_pool_lowres(sup, low[:, 0], low[:, 1], B, S)
    pred = ops.cosine_match(qry, fgp, bgp)["pred"].view(B, 2, h, h)
    m8 = ops.upsample_argmax(pred, (H, H))["mask8"]
    ref = (torch.rand(B, H, H, device=dev) > 0.5).to(torch.uint8)
    cls = torch.randint(1, 6, (B,), device=dev)
    stat = torch.zeros(21, 3, dtype=torch.int64, device=dev)
    feats_all = torch.cat((sup.view(B, S, c, h, h), qry.view(B, 1, c, h, h)), dim=1).view(B * (S + 1), c, h, h).contiguous()
    # training path (K12): the forward-for-training state, upstream gradients of the prototypes / of the prediction
    _, _, saved = ops.meta_proto_attn_train(sup, ctr, low[:, 0], low[:, 1], B, S)
    gfp, gbp = torch.randn(B, c, 3, device=dev, generator=g), torch.randn(B, c, 3, device=dev, generator=g)
    gpred = torch.randn(B, 2, hw, device=dev, generator=g)
    rows = [
        # backward kernels: read the feature map once, write its gradient once
        ("K12 meta_proto_attn_bwd (TMA + 3xTF32 mma.sync)", lambda: ops.meta_proto_attn_bwd(saved, gfp, gbp, B, S),
         B * S * (2 * c * hw + 2 * hw) * 4),
        ("K12 cosine_match_bwd", lambda: ops.cosine_match_bwd(qry, fgp, bgp, gpred), B * (2 * c * hw + 2 * hw) * 4),
        ("K0 mask_nearest", lambda: ops.mask_nearest(sup_mask, h, h), B * S * 2 * hw * 4 * 2),
        ("K2 meta_proto_attn", lambda: ops.meta_proto_attn(sup, ctr, low[:, 0], low[:, 1], B, S),
         B * (S * (c * hw + 2 * hw) * 4 + 2 * c * 6 * 4)),
        ("K1 map_pool_lowres", lambda: ops.map_pool_lowres(sup, low[:, 0], low[:, 1], B, S),
         B * (S * (c * hw + 2 * hw) * 4 + 2 * c * 4)),
        ("K3 cosine_match P=3", lambda: ops.cosine_match(qry, fgp, bgp), B * (c * hw * 4 + c * 6 * 4 + 2 * hw * 4)),
        # single-prototype matching over the S support maps of every episode (Baseline / PANet / alignLoss shape)
        ("K3 cosine_match P=1", lambda: ops.cosine_match(sup, fg1, bg1), B * S * (c * hw * 4 + 2 * hw * 4) + B * c * 2 * 4),
        ("K4 upsample_argmax", lambda: ops.upsample_argmax(pred, (H, H)), B * (2 * hw * 4 + H * H)),
        ("K10 iou_hist", lambda: ops.iou_hist(m8, ref, cls, stat), B * 2 * H * H),
        ("K4+K10 upsample_argmax_hist", lambda: ops.upsample_argmax_hist(pred, (H, H), ref.view(B, H, H), cls, stat),
         B * (2 * hw * 4 + 2 * H * H)),
        ("K6 map_pool_fullres", lambda: ops.map_pool_fullres(sup.view(B * S, c, h, h), sup_mask, B, S),
         B * (S * (c * hw * 4 + 2 * H * H * 4) + 2 * c * 4)),
        ("K6 map_pool_fullres (uint8 label maps)", lambda: ops.map_pool_fullres(sup.view(B * S, c, h, h), lab8, B, S),
         B * (S * (c * hw * 4 + H * H) + 2 * c * 4)),
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
    for name, fn, nbytes in rows:
        if only and not any(tok in name for tok in only.split(",")):
            continue
        ms = timeit(fn)
        gbs = nbytes / ms / 1e6
        add({"kernel": name, "bound": "hbm", "ms": round(ms, 4), "alg_GB": round(nbytes / 1e9, 4), "achieved": round(gbs, 1), "unit": "GB/s",
             "peak": pk, "frac": round(gbs / pk, 3), "peak_source": how, "B": B, "S": S})
    del sup, qry, sup_mask, fgfull, feats_all
    if not only or "K11" in only:
        # ResNetCM.comm call site 2 (backbones.py:235): x2 [B*6, 256, 101, 101], stride 1
        Nc, cc, hc = 16 * 6, 256, 101
        xc = torch.randn(Nc, cc, hc, hc, device=dev, generator=g)
        mc = (torch.rand(Nc, 1, hc, hc, device=dev, generator=g) > 0.7).float()
        wc, bc = torch.randn(2, 2 * cc, device=dev, generator=g), torch.randn(2, device=dev, generator=g)
        ms = timeit(lambda: ops.comm_module(xc, mc, wc, bc, 6, 1))
        nb = Nc * (cc * hc * hc + 2 * hc * hc + 2 * hc * hc) * 4
        add({"kernel": "K11 comm_module (maxpool + pool + linear + expand)", "bound": "hbm", "ms": round(ms, 4), "alg_GB": round(nb / 1e9, 4),
             "achieved": round(nb / ms / 1e6, 1), "unit": "GB/s", "peak": pk, "frac": round(nb / ms / 1e6 / pk, 3), "peak_source": how,
             "N": Nc, "c": cc, "hw": hc * hc})
        del xc, mc
    if not only or "K9" in only:
        # BASELINE config 5: PFENet 5-shot prior at 473 x 473 -> 60 x 60, C = 2048; k9_batch episodes per call (inputs > L2)
        Bp, Sp, Cp, spx = k9_batch, 5, 2048, 60
        q4 = torch.relu(torch.randn(Bp, Cp, spx, spx, device=dev, generator=g))
        s4 = torch.relu(torch.randn(Sp, Bp, Cp, spx, spx, device=dev, generator=g))
        sm = (torch.rand(Sp, Bp, spx, spx, device=dev, generator=g) > 0.5).float()
        flop = 2.0 * Cp * (Sp * spx * spx) * (spx * spx) * Bp
        for prec, tag, mult in ((0, "bf16", 1), (2, "bf16x3", 3)):
            ms = timeit(lambda: ops.prior_mask(q4, s4, sm, precision=prec), iters=5, warm=2)
            tf = flop / ms / 1e9
            add({"kernel": f"K9 prior_mask {tag} (whole op: pre-pass + tcgen05 GEMM + tail)", "bound": "tensor", "ms": round(ms, 4),
                 "alg_GFLOP": round(flop / 1e9, 1), "achieved": round(tf, 1), "unit": "TFLOP/s", "peak": tpk, "frac": round(tf / tpk, 3),
                 "executed_TFLOPs": round(tf * mult, 1), "tensor_pipe_frac_executed": round(tf * mult / tpk, 3),
                 "peak_source": how, "B": Bp, "S": Sp, "C": Cp, "hw": spx * spx})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--S", type=int, default=5)
    ap.add_argument("--c", type=int, default=512)
    ap.add_argument("--hw", type=int, default=51)
    ap.add_argument("--H", type=int, default=401)
    ap.add_argument("--only", default="")
    ap.add_argument("--k9-batch", type=int, default=2)
    ap.add_argument("--fp32-anchor", action="store_true", help="also time the CUDA-core fp32 anchor of K9 (12 ms)")
    a = ap.parse_args()
    run(a.B, a.S, a.c, a.hw, a.H, a.only, a.k9_batch, emit=lambda r: print(json.dumps(r), flush=True))
    if a.fp32_anchor:
        dev = "cuda"
        q4 = torch.relu(torch.randn(1, 2048, 60, 60, device=dev))
        s4 = torch.relu(torch.randn(5, 1, 2048, 60, 60, device=dev))
        sm = (torch.rand(5, 1, 60, 60, device=dev) > 0.5).float()
        ms = timeit(lambda: ops.prior_mask(q4, s4, sm, precision=1), iters=3, warm=1)
        print(json.dumps({"kernel": "K9 prior_mask fp32 CUDA-core anchor (tests only)", "ms": round(ms, 3),
                          "TFLOPs": round(2.0 * 2048 * 5 * 3600 * 3600 / ms / 1e9, 1)}))


if __name__ == "__main__":
    main()
