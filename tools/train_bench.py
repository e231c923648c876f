"""Times the training-path kernels (K2 forward-for-training, K2 backward, K3 backward) at the PEMP Stage-1 training shape
and the same step through stock PyTorch autograd over the reference's ops.  One JSON line per row."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pemp_b200 import autograd as A, ops   # noqa: E402


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def torch_head(f5, ctr, fg, bg, S, P, scalar=20.0):
    """The reference's ops (pemp_stage1.py:202-261) in stock PyTorch on the GPU."""
    B, _, c, h, w = f5.shape
    sup = f5[:, :S].reshape(B * S, c, h * w)
    qry = f5[:, S:].reshape(-1, c, h * w)
    diff = sup[:, :, None, :] - ctr.view(1, c, 2 * P, 1)
    attn = torch.softmax(-(diff ** 2).sum(1).view(B * S, 2, P, h * w), dim=2) * torch.stack((fg, bg), 1)[:, :, None, :]
    attn = attn.view(B * S, 1, 2 * P, h * w)
    cen = ((sup[:, :, None, :] * attn).sum(3) / (attn.sum(3) + 1e-6)).view(B, S, c, 2, P).mean(1)      # [B, c, 2, P]
    qn = torch.nn.functional.normalize(qry, dim=1, eps=1e-8)
    pn = torch.nn.functional.normalize(cen, dim=1, eps=1e-8)
    Q = qry.shape[0] // B
    sim = torch.einsum("bqcx,bcgp->bqgpx", qn.view(B, Q, c, h * w), pn) * scalar
    pred = sim.max(dim=3).values.flip(2)                       # channel 0 = background
    return pred.reshape(B * Q, 2, h, w)


def main():
    B, S, Q, c, h, P = int(os.environ.get("TB_B", 16)), 5, 1, 512, 51, 3
    dev = "cuda"
    hw = h * h
    peak = 6546.2
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    g = torch.Generator(device=dev).manual_seed(1)
    feats = torch.randn(B, S + Q, c, h, h, device=dev, generator=g)
    ctr = torch.randn(c, 2 * P, device=dev, generator=g) * 0.5
    fg = (torch.rand(B * S, hw, device=dev, generator=g) > 0.6).float()
    bg = 1 - fg
    gf = torch.randn(B, c, P, device=dev, generator=g)
    gb = torch.randn(B, c, P, device=dev, generator=g)
    gp = torch.randn(B * Q, 2, hw, device=dev, generator=g)
    fgp, bgp, saved = ops.meta_proto_attn_train(feats[:, :S], ctr, fg, bg, B, S)
    rows = [
        ("K2 forward (training: + per-shot centres)", lambda: ops.meta_proto_attn_train(feats[:, :S], ctr, fg, bg, B, S),
         B * S * (c * hw + 2 * hw) * 4),
        ("K2 backward", lambda: ops.meta_proto_attn_bwd(saved, gf, gb, B, S), B * S * (2 * c * hw + 2 * hw) * 4),
        ("K3 backward", lambda: ops.cosine_match_bwd(feats[:, S:], fgp, bgp, gp), B * Q * (2 * c * hw + 2 * hw) * 4),
    ]
    for name, fn, nbytes in rows:
        ms = timeit(fn)
        print(json.dumps({"kernel": name, "ms": round(ms, 4), "alg_GB": round(nbytes / 1e9, 4), "GBps": round(nbytes / ms / 1e6, 1),
                          "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 3), "B": B, "S": S}))
    target = torch.randint(0, 2, (B * Q, 401, 401), device=dev, generator=g)
    low = torch.stack((fg, bg), 1)

    # CELossDT weights (K14) on object-like masks, against the reference's host path (boundary -> host, scipy EDT, back)
    blob = torch.zeros(B * Q, 401, 401, dtype=torch.int64, device=dev)
    for n in range(B * Q):
        blob[n, 40 + n % 50: 250 + n % 90, 60 + n % 70: 300 + n % 40] = 1
    ms = timeit(lambda: ops.boundary_weight(blob, 5.0))
    import time
    from oracle import restate as O     # the CPU restatement of the reference's host path: baseline only
    host = blob[:8].cpu()
    t0 = time.perf_counter()
    O.boundary_weight(host, 5.0)
    cpu_ms = (time.perf_counter() - t0) * 1e3 * (B * Q) / 8
    print(json.dumps({"kernel": "K14 boundary_weight (CELossDT)", "ms": round(ms, 4), "images": B * Q, "HxW": "401x401",
                      "host_scipy_ms_same_batch": round(cpu_ms, 1), "note": "host time excludes the two PCIe copies of the reference"}))

    def ours():
        f = feats.view(B * (S + Q), c, h, h).detach().requires_grad_(True)
        cc = ctr.detach().requires_grad_(True)
        loss, _ = A.pemp_head_loss(f, low, cc, B, S, Q, target)
        loss.backward()

    def stock():
        f = feats.detach().requires_grad_(True)
        cc = ctr.detach().requires_grad_(True)
        pred = torch_head(f, cc, fg, bg, S, P)
        lg = torch.nn.functional.interpolate(pred, size=(401, 401), mode="bilinear", align_corners=True)
        torch.nn.functional.cross_entropy(lg, target, ignore_index=255).backward()

    torch.cuda.reset_peak_memory_stats()
    t_ours = timeit(ours, n=10)
    m_ours = torch.cuda.max_memory_allocated()
    torch.cuda.reset_peak_memory_stats()
    t_stock = timeit(stock, n=5, warm=2)
    m_stock = torch.cuda.max_memory_allocated()
    print(json.dumps({"step": "head forward + loss + backward", "episodes": B, "ours_ms": round(t_ours, 3), "stock_pytorch_ms": round(t_stock, 3),
                      "speedup": round(t_stock / t_ours, 2), "ours_peak_GB": round(m_ours / 1e9, 2), "stock_peak_GB": round(m_stock / 1e9, 2)}))


if __name__ == "__main__":
    main()
