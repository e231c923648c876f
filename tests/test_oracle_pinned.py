"""Pin `oracle/restate.py` to the reference: every fixture in tests/golden/ was produced by the
unmodified reference code (see `oracle/make_golden.py`).  CPU only."""
import json

import numpy as np
import pytest
import torch

from conftest import golden, nrel, unpack_bits
from oracle import restate as O
from pemp_b200 import episodes as E


def _spec(g):
    return E.EpisodeSpec(**json.loads(str(g["spec"])))


def _sup_mask(g, spec, B):
    shape = (B, spec.shot, spec.H, spec.W)
    fg = torch.from_numpy(unpack_bits(g["sup_fg"], shape).astype(np.float32))
    bg = torch.from_numpy(unpack_bits(g["sup_bg"], shape).astype(np.float32))
    return torch.stack((fg, bg), dim=2)


@pytest.mark.parametrize("name,out_shape", [("pemp_small_ctr", (80, 120)), ("pemp_small_map", (80, 120)),
                                            ("pemp_small_5shot", None)])
def test_pemp_head_small_bit_exact(name, out_shape):
    g = golden(name)
    spec, B = _spec(g), int(g["B"])
    sup_mask = _sup_mask(g, spec, B)
    # the generator must reproduce the stored inputs (fixtures with seeds-only rely on it)
    batch = E.make_batch(spec, [int(i) for i in g["indices"]])
    assert torch.equal(batch["sup_mask"], sup_mask)
    for stage in (1, 2):
        feats = torch.from_numpy(g[f"s{stage}_feats"])
        assert torch.equal(batch[f"feats{stage}"], feats)
        ctr = torch.from_numpy(g[f"s{stage}_ctr"]) if f"s{stage}_ctr" in g else None
        out = O.pemp_head(feats, sup_mask, ctr, B, spec.shot, spec.query, out_shape, ret_ind=ctr is not None)
        assert np.array_equal(out["pred_lowres"].numpy(), g[f"s{stage}_pred_lowres"])
        assert np.array_equal(out["logits"].numpy(), g[f"s{stage}_logits"])
        mask = unpack_bits(g[f"s{stage}_mask"], g[f"s{stage}_mask_shape"])
        assert np.array_equal(O.argmax2(out["logits"]).numpy(), mask)
        if ctr is not None:
            assert np.array_equal(out["response"].numpy(), g[f"s{stage}_response"])
    if "s2_adaptive_p" in g:
        assert np.array_equal(out["adaptive_p"].numpy(), g["s2_adaptive_p"])
    if "stat" in g:
        stat = O.few_shot_stat(O.argmax2(out["logits"]).numpy(), batch["qry_msk"].numpy(), batch["cls"].numpy(), spec.classes)
        assert np.array_equal(stat, g["stat"])


@pytest.mark.parametrize("name", ["pemp_full_5shot", "pemp_full_1shot"])
def test_pemp_head_full_size(name):
    """BASELINE shape (c=512, 51x51 -> 401x401); inputs come from the seeded generator."""
    g = golden(name)
    spec, B = _spec(g), int(g["B"])
    batch = E.make_batch(spec, [int(i) for i in g["indices"]])       # margin-screened when the reference produced the fixture
    margins = []
    for stage in (1, 2):
        out = O.pemp_head(batch[f"feats{stage}"], batch["sup_mask"], E.make_ctr(spec, stage), B, spec.shot, spec.query)
        margins.append(float((out["logits"][:, 1] - out["logits"][:, 0]).abs().min()))
        assert np.array_equal(out["pred_lowres"].numpy(), g[f"s{stage}_pred_lowres"])
        mask = unpack_bits(g[f"s{stage}_mask"], g[f"s{stage}_mask_shape"])
        assert np.array_equal(O.argmax2(out["logits"]).numpy(), mask)
    assert np.array_equal(out["adaptive_p"].numpy(), g["s2_adaptive_p"])
    stat = O.few_shot_stat(O.argmax2(out["logits"]).numpy(), batch["qry_msk"].numpy(), batch["cls"].numpy(), spec.classes)
    assert np.array_equal(stat, g["stat"])
    assert min(margins) == float(g["min_margin"]) >= 2e-5         # the screen the reference run applied holds for the restatement


def test_pemp_general_masks():
    g = golden("pemp_masks_general")
    sup, qry = torch.from_numpy(g["sup"]), torch.from_numpy(g["qry"])
    fg, bg, ctr = (torch.from_numpy(g[k]) for k in ("fg", "bg", "ctr"))
    B, S, c, h, w = sup.shape
    s = sup.reshape(B * S, c, h * w)
    q = qry.reshape(-1, c, h * w)
    fgp, bgp, _ = O.meta_proto_attention(s, fg.view(B * S, -1), bg.view(B * S, -1), ctr, B, S, 3)
    pred, idx = O.reduce_over_protos(O.cosine_match(q, fgp, bgp))
    assert np.array_equal(pred.view(-1, 2, h, w).numpy(), g["ctr_pred"])
    assert np.array_equal(O.response_map(pred, idx).view(-1, h, w).numpy(), g["ctr_response"])
    f0, b0 = O.map_pool_lowres(s, fg.view(B * S, -1), bg.view(B * S, -1), B, S)
    assert np.array_equal(O.cosine_match(q, f0, b0)[:, :, 0].view(-1, 2, h, w).numpy(), g["map_pred"])


@pytest.mark.parametrize("name", ["baseline_b2s1", "baseline_b1s3", "panet_b2s1", "panet_b1s3q2"])
def test_baseline_panet(name):
    g = golden(name)
    B, S, Q = int(g["B"]), int(g["S"]), int(g["Q"])
    feats = torch.from_numpy(g["feats"])
    fg = torch.from_numpy(unpack_bits(g["sup_fg"], g["mask_shape"]).astype(np.float32))
    sup_mask = torch.stack((fg, 1 - fg), dim=2)
    fn = O.panet_head if name.startswith("panet") else O.baseline_head
    out = fn(feats, sup_mask, B, S, Q, (90, 75))
    assert np.array_equal(out["logits"].numpy(), g["logits"])
    if "align_loss" in g:
        assert nrel(out["align_loss"].numpy(), g["align_loss"]) < 1e-6


@pytest.mark.parametrize("name", ["pfenet_prior_97", "pfenet_prior_100"])
def test_pfenet_prior(name):
    g = golden(name)
    q4, s4 = torch.from_numpy(g["q4"]), torch.from_numpy(g["s4"])
    masks = torch.from_numpy(unpack_bits(g["masks"], g["masks_shape"]).astype(np.float32))
    prior = O.pfenet_prior(q4, list(s4), list(masks))
    assert np.array_equal(prior.numpy(), g["prior"])


def test_weighted_gap():
    g = golden("pfenet_weighted_gap")
    out = O.weighted_gap(torch.from_numpy(g["supp_feat"]), torch.from_numpy(g["mask"]))
    assert nrel(out.numpy(), g["out"]) < 1e-6


@pytest.mark.parametrize("name", ["comm_resnet_s2", "comm_resnet_s1", "comm_vgg_s2"])
def test_comm_module_bit_exact(name):
    """`ResNetCM.comm` / `VGG16CM.comm` (backbones.py:208-222, 469-479) through the restatement."""
    g = golden(name)
    feat, pooled = O.comm_module(*(torch.from_numpy(g[k]) for k in ("x", "mask", "weight", "bias")), int(g["spq"]), int(g["stride"]))
    assert np.array_equal(pooled.numpy(), g["pooled"])
    assert np.array_equal(feat.numpy(), g["feat"])


@pytest.mark.parametrize("name", ["cedt_a", "cedt_b"])
def test_ce_loss_dt_bit_exact(name):
    """`CELossDT` (core/losses.py:17-43): boundary map, exact distance transform, weights and the weighted loss."""
    g = golden(name)
    loss, weight = O.ce_loss_dt(torch.from_numpy(g["inputs"]), torch.from_numpy(g["target"]), float(g["sigma"]))
    assert np.array_equal(weight.numpy(), g["weight"])
    assert float(loss) == float(g["loss"])
    # the plane without foreground gets scipy's distances to its virtual zero at (-1, 0): pixel (0, 0) is at distance 1
    assert float(weight[-1, 0, 0]) == float(np.float32(np.exp(-1.0 / float(g["sigma"]) ** 2) + 1.0))


def test_metric_known_answers():
    """The two episodes the reference ships (`http/static/.../{000_01,001_03}`): expected rows are the
    numbers in SURVEY 4 / BASELINE.md, and the Dice must round to `data.json:"acc"`."""
    g = golden("metric_known_answers")
    expected = {"000_01": ((109614, 1824, 1916), (53146, 1916, 1824), 0.966),
                "001_03": ((214817, 24, 2273), (11386, 2273, 24), 0.908)}
    for ep, (bg_row, fg_row, acc) in expected.items():
        shape = g[f"{ep}_shape"]
        pred, msk = unpack_bits(g[f"{ep}_pred"], shape), unpack_bits(g[f"{ep}_msk"], shape)
        cls = int(g[f"{ep}_cls"])
        stat = O.few_shot_stat(pred[None], msk[None], [cls], 20)
        assert np.array_equal(stat, g[f"{ep}_stat"])
        assert tuple(stat[0]) == bg_row and tuple(stat[cls]) == fg_row
        tp, fp, fn = stat[cls]
        dice = 2 * tp / (2 * tp + fp + fn)
        assert round(dice, 3) == acc
        assert str(acc) in str(g[f"{ep}_acc_json"])


def test_metric_random():
    g = golden("metric_random")
    stat = O.few_shot_stat(g["pred"], g["ref"], g["cls"], 20)
    assert np.array_equal(stat, g["stat"])
    mi, mm = O.miou(stat, g["labels"])
    bi, bm = O.miou(stat, g["labels"], binary=True)
    assert np.array_equal(mi, g["miou"]) and mm == float(g["miou_mean"])
    assert np.array_equal(bi, g["biou"]) and bm == float(g["biou_mean"])


def test_nearest_and_bilinear_match_aten():
    """The two resampling formulas are ATen's, not the textbook ones; pin them to the installed torch."""
    import torch.nn.functional as F
    torch.manual_seed(3)
    for (H, W, h, w) in ((401, 401, 51, 51), (333, 500, 42, 63), (97, 97, 13, 13)):
        x = torch.rand(2, 2, H, W)
        assert torch.equal(O.mask_nearest(x, h, w), F.interpolate(x, (h, w), mode="nearest"))
        y = torch.randn(2, 2, h, w) * 20
        assert torch.equal(O.bilinear_upsample(y, H, W), F.interpolate(y, (H, W), mode="bilinear", align_corners=True))
    q, p = torch.randn(4, 64, 100), torch.randn(4, 64)
    assert torch.equal(O.cosine_to(q, p), F.cosine_similarity(q, p[:, :, None], dim=1))


@pytest.mark.parametrize("name", ["canet_b2s2", "canet_b1s1q2"])
def test_canet_map_tile_bit_exact(name):
    """networks/canet.py:172-180 through the restatement."""
    g = golden(name)
    f = torch.from_numpy(g["features"])
    out = O.canet_map_tile(f, torch.from_numpy(g["sup_mask"]), f.shape[0], int(g["S"]), int(g["Q"]))
    assert np.array_equal(out.numpy(), g["out"])


def test_episode_screen_table_matches_the_oracle():
    """`pemp_b200/episode_screen.json` (read by the product's episode generator) is what `oracle/screen.py` derives from the
    reference restatement: re-derive a sample of accepted and rejected indices of the headline workload."""
    from oracle import screen
    from pemp_b200 import episodes as E
    spec = E.EpisodeSpec(shot=5, stages=2)
    entry = E.screen_table()[E.screen_key("pemp_stage2", spec)]
    rejected = set(entry["rejected"])
    sample = sorted(rejected)[:3] + [i for i in range(entry["candidates"]) if i not in rejected][:3]
    for i in sample:
        m, _ = screen.episode_margin("pemp_stage2", spec, i)
        assert (m < screen.THRESHOLD) == (i in rejected), (i, m)


def test_reference_runner_equals_the_port():
    """`oracle/ref_run.py` drives the reference's own classes the way its evaluator does; masks, counts and margins equal the
    restatement's bit for bit (so parity blocks mean the same whichever of the two is available)."""
    from oracle import ref_import as R, ref_run
    from pemp_b200 import episodes as E
    if not R.available():
        pytest.skip("reference tree not present on this machine")
    small = dict(channels=64, h=13, w=13, H=97, W=97, out_h=90, out_w=75)
    for wl, spec in (("pemp_stage2", E.EpisodeSpec(shot=2, **small)), ("pemp_stage1", E.EpisodeSpec(shot=1, stages=1, **small)),
                     ("baseline", E.EpisodeSpec(shot=1, stages=1, **small)),
                     ("panet", E.EpisodeSpec(shot=3, stages=1, classes=80, cls_hi=80, **small))):
        b = E.make_batch(spec, [5])
        r = ref_run.runner(wl, spec)
        assert r.kind == "reference"
        a = r.episode(b)
        r.kind = "port"
        p = r.episode(b)
        assert torch.equal(a["mask"], p["mask"]) and np.array_equal(a["stat"], p["stat"]) and a["margin"] == p["margin"]
        assert a.get("align_loss") == p.get("align_loss")
