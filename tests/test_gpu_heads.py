"""The host-side mirror of the reference API (`pemp_b200.heads`, `.metrics`, `.evaluator`) on the GPU,
checked against fixtures the unmodified reference produced (tests/golden) and the oracle."""
import json

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import golden, nrel, unpack_bits
from oracle import restate as O
from pemp_b200 import episodes as E

pytestmark = pytest.mark.gpu
TOL = 1e-5


class _Stub(nn.Module):
    """Stands in for a reference model: `encoder` returns pre-computed features, `ctr` is the parameter."""

    def __init__(self, features, ctr=None):
        super().__init__()
        self.features = features
        self.ctr = None if ctr is None else nn.Parameter(ctr)

    def encoder(self, _x):
        return self.features


def _case(name):
    g = golden(name)
    spec = E.EpisodeSpec(**json.loads(str(g["spec"])))
    B = int(g["B"])
    shape = (B, spec.shot, spec.H, spec.W)
    fg = torch.from_numpy(unpack_bits(g["sup_fg"], shape).astype(np.float32))
    bg = torch.from_numpy(unpack_bits(g["sup_bg"], shape).astype(np.float32))
    return g, spec, B, torch.stack((fg, bg), dim=2).cuda()


@pytest.mark.parametrize("name,out_shape", [("pemp_small_ctr", (80, 120)), ("pemp_small_map", (80, 120)), ("pemp_small_5shot", None)])
def test_forward_dropins_against_reference_fixtures(name, out_shape):
    """`pemp_stage1_forward` / `pemp_stage2_forward` bound on a stub model == reference forward (B > 1 goes
    through the in-place episode-stride path)."""
    from pemp_b200 import heads
    g, spec, B, sup_mask = _case(name)
    S, Q = spec.shot, spec.query
    sup_img = torch.zeros(B, S, 1, spec.H, spec.W, device="cuda")
    qry_img = torch.zeros(B, Q, 1, spec.H, spec.W, device="cuda")
    for stage in (1, 2):
        ctr = torch.from_numpy(g[f"s{stage}_ctr"]).cuda() if f"s{stage}_ctr" in g else None
        net = _Stub(torch.from_numpy(g[f"s{stage}_feats"]).cuda(), ctr)
        with torch.no_grad():
            if stage == 1:
                res = heads.pemp_stage1_forward(net, sup_img, sup_mask, qry_img, out_shape, ret_ind=ctr is not None)
            else:
                prior = torch.zeros(B * Q, 1, spec.H, spec.W, dtype=torch.int64, device="cuda")
                res = heads.pemp_stage2_forward(net, sup_img, sup_mask, qry_img, prior, out_shape, ret_ind=ctr is not None)
        logits, response = res if isinstance(res, tuple) else (res, None)
        assert logits.shape == g[f"s{stage}_logits"].shape
        assert nrel(logits.cpu().numpy(), g[f"s{stage}_logits"]) < TOL
        if response is not None:
            assert response.dtype == torch.int64
            margin = np.abs(g[f"s{stage}_logits"][:, 1] - g[f"s{stage}_logits"][:, 0])
            same = response.cpu().numpy() == g[f"s{stage}_response"]
            assert same.mean() > 0.995 and same[margin > 1e-3].mean() > 0.999
        if stage == 2 and ctr is not None:
            assert nrel(net.adaptive_p.cpu().numpy(), g["s2_adaptive_p"]) < TOL


def test_compute_similarity_shapes_match_reference():
    from pemp_b200 import heads
    torch.manual_seed(0)
    q = torch.randn(2, 32, 1, 7, 9)
    fg, bg = torch.randn(2, 32, 3), torch.randn(2, 32, 3)
    out = heads.compute_similarity(None, fg.cuda(), bg.cuda(), q.cuda())
    assert out.shape == (2, 2, 3, 7, 9)
    ref = O.cosine_match(q.view(2, 32, 63), fg, bg).view(2, 2, 3, 7, 9)
    assert nrel(out.cpu(), ref) < TOL
    out2 = heads.compute_similarity(None, fg[:, :, 0].cuda().contiguous(), bg[:, :, 0].cuda().contiguous(), q[:, :, 0].cuda())
    assert out2.shape == (2, 2, 7, 9)
    # PANet expansion: 4 support maps against 2 prototype sets (panet.py:145-149)
    s = torch.randn(4, 32, 7, 9)
    out3 = heads.compute_similarity(None, fg[:, :, 0].cuda().contiguous(), bg[:, :, 0].cuda().contiguous(), s.cuda())
    ref3 = O.cosine_match(s.view(4, 32, 63), fg[:, :, 0], bg[:, :, 0])[:, :, 0].view(4, 2, 7, 9)
    assert nrel(out3.cpu(), ref3) < TOL


@pytest.mark.parametrize("name", ["baseline_b2s1", "baseline_b1s3", "panet_b2s1", "panet_b1s3q2"])
def test_baseline_panet_forward_dropins(name):
    from pemp_b200 import heads
    g = golden(name)
    B, S, Q = int(g["B"]), int(g["S"]), int(g["Q"])
    fg = torch.from_numpy(unpack_bits(g["sup_fg"], g["mask_shape"]).astype(np.float32))
    sup_mask = torch.stack((fg, 1 - fg), dim=2).cuda()
    H, W = sup_mask.shape[-2:]
    net = _Stub(torch.from_numpy(g["feats"]).cuda())
    sup_img, qry_img = torch.zeros(B, S, 1, H, W, device="cuda"), torch.zeros(B, Q, 1, H, W, device="cuda")
    if name.startswith("panet"):
        logits, loss = heads.panet_forward(net, sup_img, sup_mask, qry_img, (90, 75))
        assert loss.dim() == 0 and abs(float(loss) - float(g["align_loss"])) < 1e-5 * max(1.0, float(g["align_loss"]))
    else:
        logits = heads.baseline_forward(net, sup_img, sup_mask, qry_img, (90, 75))
    assert nrel(logits.cpu().numpy(), g["logits"]) < TOL


def test_panet_align_beyond_reference_shapes():
    """B > 1 with S > 1 raises inside the reference (`.view` of an expanded tensor); the oracle states the intended
    b-major semantics and the kernel follows it."""
    from pemp_b200 import ops
    torch.manual_seed(1)
    B, S, Q, c, h, w, H, W = 2, 3, 2, 24, 9, 9, 65, 65
    qf, sf = torch.randn(B * Q, c, h, w), torch.randn(B * S, c, h, w)
    pred = torch.randn(B * Q, 2, h, w)
    m = (torch.rand(B * S, 1, H, W) > 0.5).float()
    want = O.panet_align_loss(qf, pred, sf, m, Q)
    got = ops.panet_align(qf.cuda(), pred.cuda(), sf.cuda(), m.cuda(), Q)
    assert abs(float(got) - float(want)) < 1e-5 * max(1.0, float(want))


def test_pfenet_heads():
    from pemp_b200 import heads
    g = golden("pfenet_prior_97")
    q4, s4 = torch.from_numpy(g["q4"]).cuda(), torch.from_numpy(g["s4"]).cuda()
    masks = torch.from_numpy(unpack_bits(g["masks"], g["masks_shape"]).astype(np.float32)).cuda()
    prior = heads.prior_mask(q4, list(s4), list(masks))
    assert prior.shape == g["prior"].shape
    assert np.abs(prior.cpu().numpy() - g["prior"]).max() < 2e-4      # normalised map: error amplified by 1/(max-min)
    gg = golden("pfenet_weighted_gap")
    out = heads.Weighted_GAP(torch.from_numpy(gg["supp_feat"]).cuda(), torch.from_numpy(gg["mask"]).cuda())
    assert nrel(out.cpu().numpy(), gg["out"]) < TOL


def test_few_shot_metric_interface():
    from pemp_b200.metrics import FewShotMetric
    g = golden("metric_random")
    fm = FewShotMetric(20)
    fm.update(g["pred"], g["ref"], g["cls"])                   # NumPy in, like the reference's test_step output
    assert fm.stat.dtype == np.float64 and np.array_equal(fm.stat, g["stat"].astype(np.float64))
    mi, mm = fm.mIoU(g["labels"])
    bi, bm = fm.mIoU(g["labels"], binary=True)
    assert np.array_equal(mi, g["miou"]) and mm == float(g["miou_mean"])
    assert np.array_equal(bi, g["biou"]) and bm == float(g["biou_mean"])
    fm2 = FewShotMetric(20)
    fm2.update(torch.from_numpy(g["pred"]).cuda().long(), torch.from_numpy(g["ref"]).cuda(), torch.from_numpy(g["cls"]).cuda())
    assert np.array_equal(fm2.stat, fm.stat)
    ka = golden("metric_known_answers")
    fm3 = FewShotMetric(20)
    for ep in ("000_01", "001_03"):
        shape = ka[f"{ep}_shape"]
        fm3.update(unpack_bits(ka[f"{ep}_pred"], shape)[None], unpack_bits(ka[f"{ep}_msk"], shape)[None], [int(ka[f"{ep}_cls"])])
    assert np.array_equal(fm3.stat, (ka["000_01_stat"] + ka["001_03_stat"]).astype(np.float64))


def _pipeline_vs_oracle(spec, B, batch):
    from pemp_b200.evaluator import PEMPStage2Pipeline
    S, Q = spec.shot, spec.query
    ctr1, ctr2 = E.make_ctr(spec, 1), E.make_ctr(spec, 2)
    want = O.stage2_episode_batch(batch["feats1"], batch["feats2"], batch["sup_mask"], ctr1, ctr2, B, S, Q,
                                  batch["qry_msk"].numpy(), batch["cls"].numpy(), spec.classes)
    pipe = PEMPStage2Pipeline(ctr1.cuda(), ctr2.cuda(), spec.classes)
    c, h, w = spec.channels, spec.h, spec.w
    f1 = batch["feats1"].cuda().view(B, S + Q, c, h, w)
    f2 = batch["feats2"].cuda().view(B, S + Q, c, h, w)
    stat = torch.zeros(spec.classes + 1, 3, dtype=torch.int64, device="cuda")
    prior, mask = pipe.step(f1[:, :S], f1[:, S:], f2[:, :S], f2[:, S:], batch["sup_mask"].cuda(), batch["qry_msk"].cuda(),
                            batch["cls"].cuda(), stat)
    return want, prior.cpu().long(), mask.cpu().long(), stat.cpu().numpy()


def test_stage2_pipeline_against_oracle_batch():
    """Evaluator glue (entry/pemp_stage2.py:58-65) for a batch of margin-screened episodes: prior masks, masks and the count
    table are bit-identical to the oracle's - no tolerance, no conditional."""
    from conftest import screened_episodes
    spec = E.EpisodeSpec(shot=2, channels=64, h=13, w=13, H=97, W=97, out_h=90, out_w=75)
    B = 3
    batch, idx, rejected = screened_episodes("pemp_stage2", spec, B, start=10)
    want, prior, mask, stat = _pipeline_vs_oracle(spec, B, batch)
    print(json.dumps({"case": "stage-2 pipeline, small spec", "episodes": idx, "rejected_by_margin_screen": rejected}))
    assert int((prior != want["prior"]).sum()) == 0
    assert int((mask != want["mask"]).sum()) == 0
    assert np.array_equal(stat, want["stat"])


def test_stage2_bench_batch_of_64_episodes_equals_the_oracle():
    """The headline batch itself (BASELINE config 3: 64 five-shot episodes, c = 512, 51 x 51, 401 x 401 - the episodes rank 0 of
    `bench.py` times) against `O.stage2_episode_batch`, episode by episode on the CPU (the reference's test batch size is 1):
    every prior mask, every mask and the accumulated count table bit-identical.  The episode set is the margin-screened one
    (`pemp_b200/episode_screen.json`, rejection rate printed); the UNSCREENED flip count of the first 16 raw episodes is printed
    as well (SURVEY 7 hard part 2, tier T3) and is not part of the gate."""
    from pemp_b200.evaluator import PEMPStage2Pipeline
    spec = E.EpisodeSpec(shot=5, stages=2)
    B, S, Q, c, h, w = 64, spec.shot, spec.query, spec.channels, spec.h, spec.w
    idx = E.screened_indices("pemp_stage2", spec, B)
    batch = E.make_batch(spec, idx)
    ctr1, ctr2 = E.make_ctr(spec, 1), E.make_ctr(spec, 2)
    pipe = PEMPStage2Pipeline(ctr1.cuda(), ctr2.cuda(), spec.classes)

    def gpu(b, n):
        f1 = b["feats1"].cuda().view(n, S + Q, c, h, w)
        f2 = b["feats2"].cuda().view(n, S + Q, c, h, w)
        stat = torch.zeros(spec.classes + 1, 3, dtype=torch.int64, device="cuda")
        prior, mask = pipe.step(f1[:, :S], f1[:, S:], f2[:, :S], f2[:, S:], b["sup_mask"].cuda(), b["qry_msk"].cuda(), b["cls"].cuda(), stat)
        return prior.cpu().long(), mask.cpu().long(), stat.cpu().numpy()

    def oracle(b, n):
        per = S + Q
        priors, masks, stat = [], [], np.zeros((spec.classes + 1, 3), np.int64)
        for j in range(n):
            r = O.stage2_episode_batch(b["feats1"][j * per:(j + 1) * per], b["feats2"][j * per:(j + 1) * per], b["sup_mask"][j:j + 1],
                                       ctr1, ctr2, 1, S, Q, b["qry_msk"][j:j + 1].numpy(), b["cls"][j:j + 1].numpy(), spec.classes)
            priors.append(r["prior"]); masks.append(r["mask"]); stat += r["stat"]
        return torch.cat(priors), torch.cat(masks), stat

    prior, mask, stat = gpu(batch, B)
    w_prior, w_mask, w_stat = oracle(batch, B)
    raw = E.make_batch(spec, range(16))                              # unscreened: reported, not gated
    r_prior, r_mask, _ = gpu(raw, 16)
    o_prior, o_mask, _ = oracle(raw, 16)
    print(json.dumps({"case": "stage2_5shot bench batch", "episodes": B, "screen": E.screen_stats("pemp_stage2", spec),
                      "flips_screened": [int((prior != w_prior).sum()), int((mask != w_mask).sum())],
                      "flips_unscreened_first_16_raw_episodes": [int((r_prior != o_prior).sum()), int((r_mask != o_mask).sum())],
                      "pixels_per_episode": int(mask[0].numel())}))
    assert int((prior != w_prior).sum()) == 0
    assert int((mask != w_mask).sum()) == 0
    assert np.array_equal(stat, w_stat)


def test_pipeline_step_takes_label_maps_or_float_masks():
    """`PEMPStage2Pipeline.step` fed with the uint8 label maps (what a host-side caller ships: an eighth of the bytes) equals the
    step fed with the loader's float `sup_mask`, bit for bit - including shots with a 255 boundary band."""
    from pemp_b200.evaluator import PEMPStage2Pipeline
    spec = E.EpisodeSpec(shot=3, channels=64, h=13, w=13, H=97, W=97, out_h=90, out_w=75)
    B, S, Q, c, h, w = 2, spec.shot, spec.query, spec.channels, spec.h, spec.w
    batch = {k: v.cuda() for k, v in E.make_batch(spec, range(B)).items()}
    labels = (batch["sup_mask"][:, :, 0] > 0.5).to(torch.uint8)
    labels[:, 0, 40:44] = 255                                         # a boundary band: neither foreground nor background
    sup_mask = torch.stack(((labels == 1).float(), (labels == 0).float()), dim=2)
    pipe = PEMPStage2Pipeline(E.make_ctr(spec, 1).cuda(), E.make_ctr(spec, 2).cuda(), spec.classes)
    f1 = batch["feats1"].view(B, S + Q, c, h, w)
    f2 = batch["feats2"].view(B, S + Q, c, h, w)
    outs = []
    for masks in (sup_mask, labels):
        stat = torch.zeros(spec.classes + 1, 3, dtype=torch.int64, device="cuda")
        prior, mask = pipe.step(f1[:, :S], f1[:, S:], f2[:, :S], f2[:, S:], masks, batch["qry_msk"], batch["cls"], stat)
        outs.append((prior, mask, stat))
    assert all(torch.equal(a, b) for a, b in zip(*outs))


def test_episode_view_equals_dense_copy():
    from pemp_b200 import ops
    torch.manual_seed(2)
    B, S, Q, c, h, w = 3, 2, 1, 32, 9, 11
    feats = torch.randn(B, S + Q, c, h, w, device="cuda")
    fg = (torch.rand(B * S, h * w, device="cuda") > 0.5).float()
    ctr = torch.rand(c, 6, device="cuda")
    a = ops.meta_proto_attn(feats[:, :S], ctr, fg, 1 - fg, B, S)
    b = ops.meta_proto_attn(feats[:, :S].reshape(B * S, c, h * w), ctr, fg, 1 - fg, B, S)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    p1 = ops.cosine_match(feats[:, S:], a[0], a[1])["pred"]
    p2 = ops.cosine_match(feats[:, S:].reshape(B * Q, c, h * w), a[0], a[1])["pred"]
    assert torch.equal(p1, p2)
    m1 = ops.map_pool_lowres(feats[:, :S], fg, 1 - fg, B, S)
    m2 = ops.map_pool_lowres(feats[:, :S].reshape(B * S, c, h * w), fg, 1 - fg, B, S)
    assert torch.equal(m1[0], m2[0]) and torch.equal(m1[1], m2[1])


def test_bench_size_properties_of_the_stage2_pipeline():
    """BASELINE size (64 five-shot episodes per step, c = 512, 51 x 51, 401 x 401) through properties that do not need the
    oracle: (1) the batch holds every (margin-screened) episode twice (b and b + 32), in different CTAs / tile ranges of the
    persistent kernels - prototypes agree to fp32 summation-order noise, prior masks, masks and counts of the two copies are
    identical; (2) the count table of the full batch equals the sum of the tables of its two halves run separately
    (different launch shapes), bit for bit; (3) rows 1.. of the table only contain the classes present."""
    from pemp_b200 import ops
    from pemp_b200.evaluator import PEMPStage2Pipeline
    spec = E.EpisodeSpec()
    B, S, Q, c, h, w = 64, spec.shot, spec.query, spec.channels, spec.h, spec.w
    half = {k: v.cuda() for k, v in E.make_batch(spec, E.screened_indices("pemp_stage2", spec, B // 2, start=64)).items()}
    rep = lambda t: torch.cat((t, t), dim=0).contiguous()
    f1 = rep(half["feats1"].view(B // 2, S + Q, c, h, w))
    f2 = rep(half["feats2"].view(B // 2, S + Q, c, h, w))
    sup_mask, qry_msk, cls = rep(half["sup_mask"]), rep(half["qry_msk"]), rep(half["cls"])
    ctr1, ctr2 = E.make_ctr(spec, 1).cuda(), E.make_ctr(spec, 2).cuda()
    pipe = PEMPStage2Pipeline(ctr1, ctr2, spec.classes)

    def run(sl):
        stat = torch.zeros(spec.classes + 1, 3, dtype=torch.int64, device="cuda")
        prior, mask = pipe.step(f1[sl, :S], f1[sl, S:], f2[sl, :S], f2[sl, S:], sup_mask[sl], qry_msk[sl], cls[sl], stat)
        return prior, mask, stat

    prior, mask, stat = run(slice(0, B))
    # (1) the two copies of every episode
    low = ops.mask_nearest(sup_mask.view(B * S, 2, spec.H, spec.W), h, w).view(B * S, 2, h * w)
    fgp, bgp, _ = ops.meta_proto_attn(f2[:, :S], ctr2, low[:, 0], low[:, 1], B, S)
    assert nrel(fgp[:32].cpu(), fgp[32:].cpu()) < 2e-6 and nrel(bgp[:32].cpu(), bgp[32:].cpu()) < 2e-6
    assert torch.equal(mask[:32], mask[32:]) and torch.equal(prior[:32], prior[32:])
    # (2) additivity across launch shapes
    _, m_a, s_a = run(slice(0, 32))
    _, m_b, s_b = run(slice(32, 64))
    assert torch.equal(torch.cat((m_a, m_b)), mask)
    assert torch.equal(s_a + s_b, stat) and torch.equal(s_a, s_b)
    # the table is exactly what the masks say (recount on the host with the reference's NumPy formulas)
    ref = O.few_shot_stat(mask.cpu().numpy(), qry_msk.cpu().numpy(), cls.cpu().numpy(), spec.classes)
    assert np.array_equal(stat.cpu().numpy(), ref)
    # (3) only the classes present have counts
    present = set(int(v) for v in cls.tolist())
    rows = stat.cpu().numpy()
    assert all((rows[k] == 0).all() for k in range(1, spec.classes + 1) if k not in present)


@pytest.mark.parametrize("B,c", [(3, 64), (8, 512)])
def test_graphed_step_replays_the_eager_step(B, c):
    """SURVEY 8f row 1 ("CUDA-graph the head"): one graph launch replays the nine kernels of `step` on static buffers -
    same masks and counts as the eager step, also after the buffers were refilled, and `stat` keeps accumulating."""
    from pemp_b200 import ops
    from pemp_b200.evaluator import PEMPStage2Pipeline
    spec = E.EpisodeSpec(channels=c)
    S, Q, h, w = spec.shot, spec.query, spec.h, spec.w
    ctr1, ctr2 = E.make_ctr(spec, 1).cuda(), E.make_ctr(spec, 2).cuda()
    pipe = PEMPStage2Pipeline(ctr1, ctr2, spec.classes)
    batches = [E.device_batch(spec, B, "cuda", seed=900 + k) for k in range(3)]

    def args_of(b, stat):
        f1 = b["feats1"].view(B, S + Q, c, h, w)
        f2 = b["feats2"].view(B, S + Q, c, h, w)
        return f1[:, :S], f1[:, S:], f2[:, :S], f2[:, S:], b["sup_mask"], b["qry_msk"], b["cls"], stat

    static = {k: v.clone() for k, v in batches[0].items() if torch.is_tensor(v)}
    g_stat = torch.zeros(spec.classes + 1, 3, dtype=torch.int64, device="cuda")
    g_stat[0, 0] = 7                                  # pre-existing counts survive the capture
    graphed = pipe.capture(*args_of(static, g_stat))
    per_step = 9 if c == 512 else 11                  # the generic K2 (c = 64) has a separate finalize launch per head
    assert graphed.launches == per_step
    assert int(g_stat[0, 0]) == 7 and int(g_stat.sum()) == 7      # neither warm-up nor capture counted anything
    e_stat = torch.zeros_like(g_stat)
    e_stat[0, 0] = 7
    for b in batches:
        for k in ("feats1", "feats2", "sup_mask", "qry_msk", "cls"):
            static[k].copy_(b[k])
        n0 = ops.launch_count()
        prior, mask = graphed.replay()
        assert ops.launch_count() - n0 == per_step
        e_prior, e_mask = pipe.step(*args_of(b, e_stat))
        assert torch.equal(prior, e_prior) and torch.equal(mask, e_mask)
        assert torch.equal(g_stat, e_stat)
    assert int(g_stat.sum()) > 7


@pytest.mark.parametrize("name", ["canet_b2s2", "canet_b1s1q2"])
def test_canet_map_tile_against_reference_fixtures(name):
    """fixtures produced by the reference's own lines networks/canet.py:172-180.  The query half is a copy
    (bit exact); the tiled prototype is K1 (fp32 sums: 1e-5 norm-wise)."""
    from conftest import golden
    from pemp_b200 import ops
    g = golden(name)
    f = torch.from_numpy(g["features"])
    B, SQ, c, h, w = f.shape
    S, Q = int(g["S"]), int(g["Q"])
    out = ops.canet_map_tile(f.cuda().view(B * SQ, c, h, w), torch.from_numpy(g["sup_mask"]).cuda(), B, S, Q).cpu()
    want = torch.from_numpy(g["out"])
    assert torch.equal(out[:, :c], want[:, :c])
    assert nrel(out[:, c:], want[:, c:]) < 1e-5
    assert torch.equal(out[:, c:], out[:, c:, :1, :1].expand(-1, -1, h, w))          # constant over h x w


def test_canet_map_tile_bench_shape_against_oracle():
    from pemp_b200 import ops
    B, S, Q, c, h, H = 4, 5, 1, 256, 41, 321
    g = torch.Generator().manual_seed(8)
    f = torch.randn(B, S + Q, c, h, h, generator=g)
    fg = (torch.rand(B, S, 1, H, H, generator=g) > 0.5).float()
    sup_mask = torch.cat((fg, 1 - fg), dim=2)
    want = O.canet_map_tile(f, sup_mask, B, S, Q)
    out = ops.canet_map_tile(f.cuda().view(B * (S + Q), c, h, h), sup_mask.cuda(), B, S, Q).cpu()
    assert torch.equal(out[:, :c], want[:, :c]) and nrel(out[:, c:], want[:, c:]) < 1e-5


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ops_follow_the_device_of_their_operands():
    """raw pointers of one GPU must never be launched on another GPU's stream.  A call whose tensors live on
    cuda:1 while cuda:0 is current runs under cuda:1 (same result as on cuda:0); operands on two devices are refused."""
    from pemp_b200 import ops
    from pemp_b200.metrics import FewShotMetric
    torch.manual_seed(4)
    B, S, c, h, w = 2, 2, 64, 9, 9
    f = torch.randn(B * S, c, h * w)
    fg = (torch.rand(B * S, h * w) > 0.5).float()
    ctr = torch.rand(c, 6)
    assert torch.cuda.current_device() == 0
    a = ops.meta_proto_attn(f.cuda(0), ctr.cuda(0), fg.cuda(0), (1 - fg).cuda(0), B, S)
    b = ops.meta_proto_attn(f.cuda(1), ctr.cuda(1), fg.cuda(1), (1 - fg).cuda(1), B, S)
    assert b[0].device.index == 1 and all(torch.equal(x.cpu(), y.cpu()) for x, y in zip(a, b))
    assert torch.cuda.current_device() == 0
    with pytest.raises(ValueError, match="different devices"):
        ops.meta_proto_attn(f.cuda(0), ctr.cuda(1), fg.cuda(0), (1 - fg).cuda(0), B, S)
    fm = FewShotMetric(20, device="cuda:1")
    pred = (torch.rand(3, 33, 33) > 0.5).to(torch.uint8)
    ref = (torch.rand(3, 33, 33) > 0.5).to(torch.uint8)
    fm.update(pred.numpy(), ref.numpy(), [1, 2, 3])
    assert np.array_equal(fm.stat, O.few_shot_stat(pred.numpy(), ref.numpy(), [1, 2, 3], 20).astype(np.float64))
