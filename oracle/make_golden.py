"""Generate `tests/golden/*.npz` by running the UNMODIFIED reference — TEST INFRASTRUCTURE ONLY.

Run in the build container (where `/root/reference` exists):

    python -m oracle.make_golden

Every fixture stores the inputs (or the seed that regenerates them through
`pemp_b200.episodes`) and the outputs of the reference's own code, imported via
`oracle/ref_import.py`:

  pemp_small_*      `PEMPStage1.forward` / `PEMPStage2.forward` head part (`pemp_stage1.py:144-163`,
                    `pemp_stage2.py:144-162`) incl. `mpm`, `compute_similarity`, response map, `adaptive_p`
  pemp_full_*       same at the BASELINE shape (c=512, 51x51, 401x401); inputs regenerated from the seed
  baseline_*, panet_*   `Baseline.forward` / `PANet.forward` head + `alignLoss` (`baseline.py:97-118`, `panet.py:96-194`)
  pfenet_*          `Weighted_GAP` (`pfenet.py:15-20`) and the prior block (`pfenet.py:201-231`)
  comm_*            `ResNetCM.comm` / `VGG16CM.comm` (`backbones.py:208-222, 469-479`)
  metric_*          `FewShotMetric` (`core/metrics.py`) on random masks and on the two episodes the
                    reference ships under `http/static/1005_pascal_1shot_pemp_stage2_s0/` (the only
                    known-answer vectors in the reference; `data.json:"acc"` is their Dice score)
"""
import json
import os

import numpy as np
import torch

from oracle import ref_import as R
from pemp_b200 import episodes as E

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def _save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def _pemp_case(name, spec, B, ctr_stage, store_inputs, ret_ind=True, out_shape=None, first=0, screened=False):
    indices = list(range(first, first + B))
    if screened:
        # margin screen (SURVEY 7 hard part 2): the first B episodes >= first whose smallest |fg - bg| logit gap through the
        # reference head is >= screen.THRESHOLD, so that masks and counts of the fixture can be compared bit for bit
        from oracle import screen
        indices, i = [], first
        while len(indices) < B:
            if screen.episode_margin("pemp_stage2", spec, i)[0] >= screen.THRESHOLD:
                indices.append(i)
            i += 1
    batch = E.make_batch(spec, indices)
    S, Q = spec.shot, spec.query
    si, qi = R.dummy_images(B, S, Q, spec.H, spec.W)
    arrays = {"B": B, "first": first, "indices": np.array(indices), "spec": json.dumps(spec.__dict__)}
    margins = []
    for stage, model in ((1, "pemp_stage1"), (2, "pemp_stage2")):
        feats = batch[f"feats{stage}"]
        ctr = E.make_ctr(spec, stage) if ctr_stage else None
        net = R.head_only(model, feats, ctr)
        with torch.no_grad():
            if stage == 1:
                res = net(si, batch["sup_mask"], qi, out_shape, ret_ind and ctr is not None)
            else:
                prior = torch.zeros(B * Q, 1, spec.H, spec.W, dtype=torch.int64)    # encoder is stubbed
                res = net(si, batch["sup_mask"], qi, prior, out_shape, ret_ind and ctr is not None)
        logits, response = res if isinstance(res, tuple) else (res, None)
        margins.append(float((logits[:, 1] - logits[:, 0]).abs().min()))
        # low-res prediction and prototypes through the reference's own `mpm`
        c, h, w = spec.channels, spec.h, spec.w
        f5 = feats.view(B, S + Q, c, h, w)
        low = torch.nn.functional.interpolate(batch["sup_mask"].view(B * S, 2, spec.H, spec.W), (h, w), mode="nearest")
        with torch.no_grad():
            pred = net.mpm(f5[:, :S], f5[:, S:], low[:, 0], low[:, 1], False)
        arrays[f"s{stage}_pred_lowres"] = _np(pred)
        arrays[f"s{stage}_mask"] = np.packbits(_np(logits.argmax(1)).astype(np.uint8))
        arrays[f"s{stage}_mask_shape"] = np.array(logits.argmax(1).shape)
        if store_inputs:
            arrays[f"s{stage}_feats"] = _np(feats)
            arrays[f"s{stage}_logits"] = _np(logits)
            if ctr is not None:
                arrays[f"s{stage}_ctr"] = _np(ctr)
        if response is not None and store_inputs:
            arrays[f"s{stage}_response"] = _np(response).astype(np.uint8)
        if stage == 2 and ctr is not None:
            arrays["s2_adaptive_p"] = _np(net.adaptive_p)
    arrays["min_margin"] = min(margins)          # of the reference's own logits (both stages)
    if store_inputs:
        arrays["sup_fg"] = np.packbits(_np(batch["sup_mask"][:, :, 0]).astype(np.uint8))
        arrays["sup_bg"] = np.packbits(_np(batch["sup_mask"][:, :, 1]).astype(np.uint8))
    # metric on the stage-2 mask vs the synthetic ground truth, through the reference FewShotMetric
    if (out_shape or (spec.H, spec.W)) == (spec.out_h, spec.out_w):
        fm = R.few_shot_metric(spec.classes)
        fm.update(_np(logits.argmax(1)), _np(batch["qry_msk"]), batch["cls"])
        arrays["stat"] = fm.stat.astype(np.int64)
    _save(name, **arrays)


def pemp_cases():
    small = E.EpisodeSpec(shot=2, query=1, channels=32, h=13, w=13, H=97, W=97, out_h=80, out_w=120)
    _pemp_case("pemp_small_ctr", small, B=2, ctr_stage=True, store_inputs=True, out_shape=(80, 120))
    _pemp_case("pemp_small_map", small, B=2, ctr_stage=False, store_inputs=True, out_shape=(80, 120))
    five = E.EpisodeSpec(shot=5, query=1, channels=40, h=11, w=15, H=81, W=113, out_h=81, out_w=113)
    _pemp_case("pemp_small_5shot", five, B=1, ctr_stage=True, store_inputs=True)
    full = E.EpisodeSpec(shot=5)
    _pemp_case("pemp_full_5shot", full, B=1, ctr_stage=True, store_inputs=False, ret_ind=False, screened=True)
    full1 = E.EpisodeSpec(shot=1)
    _pemp_case("pemp_full_1shot", full1, B=2, ctr_stage=True, store_inputs=False, ret_ind=False, first=7, screened=True)


def pemp_masks_noncomplementary():
    """fg/bg masks with a band where both are 0 (the '255 boundary' case of `pemp_stage2.py:124-126`)
    and soft (non-binary) values, straight through `PEMPStage1.mpm`."""
    torch.manual_seed(11)
    B, S, Q, c, h, w = 1, 3, 1, 24, 9, 12
    sup = torch.randn(B, S, c, h, w)
    qry = torch.randn(B, Q, c, h, w)
    fg = (torch.rand(B * S, h, w) > 0.6).float()
    bg = ((1 - fg) * (torch.rand(B * S, h, w) > 0.2)).float()
    fg[0] *= torch.rand(h, w)                  # soft weights on one shot
    ctr = torch.rand(c, 6)
    out = {}
    for tag, cc in (("ctr", ctr), ("map", None)):
        net = R.head_only("pemp_stage1", None, cc)
        with torch.no_grad():
            res = net.mpm(sup, qry, fg, bg, cc is not None)
        pred, resp = res if isinstance(res, tuple) else (res, None)
        out[f"{tag}_pred"] = _np(pred)
        if resp is not None:
            out[f"{tag}_response"] = _np(resp).astype(np.uint8)
    _save("pemp_masks_general", sup=_np(sup), qry=_np(qry), fg=_np(fg), bg=_np(bg), ctr=_np(ctr), **out)


def baseline_panet_cases():
    for name, model, B, S, Q in (("baseline_b2s1", "baseline", 2, 1, 1), ("baseline_b1s3", "baseline", 1, 3, 1),
                                 ("panet_b2s1", "panet", 2, 1, 1), ("panet_b1s3q2", "panet", 1, 3, 2)):
        spec = E.EpisodeSpec(shot=S, query=Q, channels=32, h=13, w=13, H=97, W=97, out_h=97, out_w=97, stages=1)
        batch = E.make_batch(spec, range(B), base_seed=4321)
        si, qi = R.dummy_images(B, S, Q, spec.H, spec.W)
        net = R.head_only(model, batch["feats1"])
        with torch.no_grad():
            res = net(si, batch["sup_mask"], qi, (90, 75))
        logits, loss = res if isinstance(res, tuple) else (res, None)
        arrays = dict(B=B, S=S, Q=Q, feats=_np(batch["feats1"]),
                      sup_fg=np.packbits(_np(batch["sup_mask"][:, :, 0]).astype(np.uint8)),
                      mask_shape=np.array(batch["sup_mask"][:, :, 0].shape), logits=_np(logits))
        if loss is not None:
            arrays["align_loss"] = _np(loss)
        _save(name, **arrays)


def pfenet_cases():
    torch.manual_seed(5)
    # (H-1) % 8 == 0 -> the mask resize is the stride-8 pick; H = 100 -> fractional mask values
    for name, H, sp, C, S, B in (("pfenet_prior_97", 97, 13, 64, 3, 2), ("pfenet_prior_100", 100, 13, 48, 2, 1)):
        q4 = torch.relu(torch.randn(B, C, sp, sp))
        s4 = [torch.relu(torch.randn(B, C, sp, sp)) for _ in range(S)]
        masks = []
        for _ in range(S):
            m = torch.zeros(B, 1, H, H)
            for b in range(B):
                y0, y1, x0, x1 = E._rect(torch.Generator().manual_seed(int(torch.randint(0, 10**6, ()).item())), H, H)
                m[b, 0, y0:y1, x0:x1] = 1
            masks.append(m)
        prior = R.pfenet_prior(q4, s4, masks, (sp, sp), (sp, sp))
        _save(name, q4=_np(q4), s4=_np(torch.stack(s4)), masks=np.packbits(_np(torch.stack(masks)).astype(np.uint8)),
              masks_shape=np.array(torch.stack(masks).shape), prior=_np(prior))
    sf = torch.randn(3, 40, 13, 13)
    mk = (torch.rand(3, 1, 13, 13) > 0.5).float()
    mk[2] = 0                                   # empty mask: eps keeps it finite
    _save("pfenet_weighted_gap", supp_feat=_np(sf), mask=_np(mk), out=_np(R.weighted_gap(sf, mk)))


def metric_cases():
    from PIL import Image
    root = os.path.join(R.REF_ROOT, "http", "static", "1005_pascal_1shot_pemp_stage2_s0")
    arrays = {}
    for ep, cls in (("000_01", 1), ("001_03", 3)):
        d = os.path.join(root, ep)
        files = sorted(os.listdir(d))
        pred = np.array(Image.open(os.path.join(d, next(f for f in files if "_qry_pred_" in f))).convert("L"))
        msk = np.array(Image.open(os.path.join(d, next(f for f in files if "_qry_msk_" in f))).convert("L"))
        acc = json.load(open(os.path.join(d, "data.json")))
        # PNGs hold 0/255 (`core/base_trainer.py:356-363`); the metric sees labels 0/1
        pred01, msk01 = (pred > 127).astype(np.uint8), (msk > 127).astype(np.uint8)
        fm = R.few_shot_metric(20)
        fm.update(pred01[None], msk01[None], [cls])
        arrays[f"{ep}_pred"] = np.packbits(pred01)
        arrays[f"{ep}_msk"] = np.packbits(msk01)
        arrays[f"{ep}_shape"] = np.array(pred01.shape)
        arrays[f"{ep}_cls"] = cls
        arrays[f"{ep}_stat"] = fm.stat.astype(np.int64)
        arrays[f"{ep}_acc_json"] = json.dumps(acc)
    _save("metric_known_answers", **arrays)

    rng = np.random.RandomState(5678)               # the reference's test-sampler seed, `datasets.py:29`
    N, H, W, C = 6, 57, 83, 20
    pred = rng.randint(0, 2, (N, H, W)).astype(np.uint8)
    ref = rng.choice([0, 1, 255], size=(N, 1, H, W), p=[0.55, 0.4, 0.05]).astype(np.uint8)
    cls = rng.randint(1, C + 1, N)
    fm = R.few_shot_metric(C)
    fm.update(pred, ref.reshape(N, H, W), cls)
    labels = sorted(set(int(x) for x in cls))
    mi, mm = fm.mIoU(labels)
    bi, bm = fm.mIoU(labels, binary=True)
    _save("metric_random", pred=pred, ref=ref, cls=cls, stat=fm.stat.astype(np.int64), labels=np.array(labels),
          miou=mi, miou_mean=mm, biou=bi, biou_mean=bm)


def comm_cases():
    """`ResNetCM.comm` / `VGG16CM.comm` (backbones.py:208-222, 469-479), the "next" row of SURVEY 8f."""
    g = torch.Generator().manual_seed(77)
    for name, variant, (B, spq, c, h, Hm, stride, n) in (("comm_resnet_s2", "resnet", (2, 3, 24, 21, 41, 2, 2)),
                                                        ("comm_resnet_s1", "resnet", (1, 2, 16, 26, 26, 1, 2)),
                                                        ("comm_vgg_s2", "vgg", (2, 2, 32, 13, 26, 2, 3))):
        N = B * spq
        x = torch.randn(N, c, h, h, generator=g)
        mask = (torch.rand(N, 1, Hm, Hm, generator=g) > 0.7).float()
        mask[0] = 0                                                   # an image without any foreground
        weight, bias = torch.randn(n, 2 * c, generator=g) * 0.1, torch.randn(n, generator=g)
        feat, pooled = R.comm(variant, x, mask, weight, bias, spq, stride)
        _save(name, x=_np(x), mask=_np(mask), weight=_np(weight), bias=_np(bias), spq=spq, stride=stride,
              feat=_np(feat), pooled=_np(pooled))


def cedt_cases():
    """`CELossDT` (core/losses.py:17-43), SURVEY 8f row 4: the reference's own class on CPU tensors."""
    g = torch.Generator().manual_seed(99)
    for name, (N, H, W, sigma) in (("cedt_a", (3, 41, 57, 5.0)), ("cedt_b", (2, 64, 64, 2.0))):
        target = torch.zeros(N, H, W, dtype=torch.int64)
        for n in range(N):
            for _ in range(2):
                y0, x0 = int(torch.randint(0, H - 8, (1,), generator=g)), int(torch.randint(0, W - 8, (1,), generator=g))
                hh, ww = int(torch.randint(3, H // 2, (1,), generator=g)), int(torch.randint(3, W // 2, (1,), generator=g))
                target[n, y0:y0 + hh, x0:x0 + ww] = 1
        target[0, :, :3][target[0, :, :3] == 0] = 255                 # an ignore band
        target[N - 1] = 0                                             # a plane without any foreground (no boundary)
        target[N - 1, :2] = 255
        inputs = torch.randn(N, 2, H, W, generator=g) * 3
        loss, weight = R.ce_loss_dt(inputs, target, sigma)
        _save(name, inputs=_np(inputs), target=_np(target), sigma=sigma, loss=_np(loss), weight=_np(weight))


def canet_cases():
    """CaNet's masked average pooling + tiling + concatenation (networks/canet.py:172-180), SURVEY 8f row 4."""
    g = torch.Generator().manual_seed(123)
    for name, (B, S, Q, c, h, H) in (("canet_b2s2", (2, 2, 1, 24, 13, 97)), ("canet_b1s1q2", (1, 1, 2, 16, 9, 40))):
        feats = torch.randn(B, S + Q, c, h, h, generator=g)
        fg = (torch.rand(B, S, 1, H, H, generator=g) > 0.6).float()
        fg[0, 0] = 0                                                  # a shot without foreground
        sup_mask = torch.cat((fg, 1 - fg), dim=2)
        out = R.canet_map_tile(feats, sup_mask, B, S, Q)
        _save(name, features=_np(feats), sup_mask=_np(sup_mask), S=S, Q=Q, out=_np(out))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    pemp_cases()
    pemp_masks_noncomplementary()
    baseline_panet_cases()
    pfenet_cases()
    metric_cases()
    comm_cases()
    cedt_cases()
    canet_cases()


if __name__ == "__main__":
    main()
