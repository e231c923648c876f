"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, host logic
(episode generator, sharding, argument validation, drop-in binding) and the world-size-2 gloo path."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from pemp_b200 import _cabi, dist as pdist, episodes as E


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pemp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pemp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 20
    lib = _cabi.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pemp_b200.h but not exported"
    assert sorted(_cabi.SIGNATURES) == names          # the ctypes table covers the header one to one
    assert lib.pemp_abi_version() == 1
    assert "WORKSPACE" in _cabi.strerror(-3) and _cabi.strerror(0) == "ok"


def test_workspace_queries_are_pure_host_functions():
    lib = _cabi.lib()
    assert lib.pemp_meta_proto_attn_workspace_bytes(64, 5, 512, 2601, 3) > 0
    assert lib.pemp_meta_proto_attn_workspace_bytes(64, 5, 512, 2601, 5) == 0      # p out of range
    assert lib.pemp_map_pool_workspace_bytes(0, 1, 1, 1) == 0
    assert lib.pemp_panet_align_workspace_bytes(1, 5, 1, 512, 51, 51, 401, 401) > 0


def test_no_cpu_fallback():
    from pemp_b200 import ops
    with pytest.raises(ValueError, match="CUDA"):
        ops.mask_nearest(torch.zeros(1, 2, 8, 8), 4, 4)
    with pytest.raises(ValueError, match="CUDA"):
        ops.cosine_match(torch.zeros(2, 8, 10), torch.zeros(2, 8), torch.zeros(2, 8))
    if not torch.cuda.is_available():
        from pemp_b200.metrics import FewShotMetric
        with pytest.raises(RuntimeError):
            FewShotMetric(20)


def test_training_path_has_no_cpu_fallback_and_plans_on_the_host():
    """K12-K15 host logic: CPU tensors are refused before the library is touched; the workspace planners are pure host code
    (the backward planner picks the image split from the wave count - more CTAs per image when there are few images)."""
    from pemp_b200 import autograd as A, ops
    f = torch.zeros(1, 2, 16, 4, 4)
    with pytest.raises(ValueError, match="CUDA"):
        A.meta_proto_attn(f[:, :1], torch.zeros(16, 6), torch.zeros(1, 16), torch.zeros(1, 16))
    with pytest.raises(ValueError, match="CUDA"):
        A.cosine_match(f[:, 1:], torch.zeros(1, 16, 3), torch.zeros(1, 16, 3))
    with pytest.raises(ValueError, match="CUDA"):
        A.upsample_ce(torch.zeros(1, 2, 4, 4), torch.zeros(1, 9, 9, dtype=torch.int64))
    with pytest.raises(ValueError, match="CUDA"):
        ops.boundary_weight(torch.zeros(1, 9, 9, dtype=torch.int64), 5.0)
    with pytest.raises(ValueError):
        A.pemp_head_loss(torch.zeros(2, 16, 4, 4), torch.zeros(1, 2, 16), torch.zeros(16, 6), 1, 1, 1,
                         torch.zeros(1, 9, 9, dtype=torch.int64), out_shape=(5, 5))
    lib = _cabi.lib()
    few = lib.pemp_meta_proto_attn_bwd_workspace_bytes(1, 1, 512, 2601, 3)
    many = lib.pemp_meta_proto_attn_bwd_workspace_bytes(64, 5, 512, 2601, 3)
    assert few > 0 and many > few
    per_image_few = few - 512 * 6 * 4            # minus the coefficient table: partials of one image
    assert per_image_few > (many / 320) * 2      # one image alone is split over more CTAs than each of 320 images
    assert lib.pemp_meta_proto_attn_bwd_workspace_bytes(1, 1, 512, 2601, 5) == 0
    # tensor-path shapes (p = 3, c in {128, 256, 512, 1024}, hw >= 32) also carry the image tables [N, c, 12] and the per-image
    # partial sums [N, (c + 1) 6]; any other shape plans the CUDA-core kernel only
    N, c, hw = 10, 512, 2601
    with_tables = lib.pemp_meta_proto_attn_bwd_workspace_bytes(2, 5, c, hw, 3)
    assert with_tables >= N * c * 12 * 4 + N * (c + 1) * 6 * 4 + N * c * 6 * 4
    small_map = lib.pemp_meta_proto_attn_bwd_workspace_bytes(2, 5, c, 25, 3)       # hw < 32: no tensor map
    other_p = lib.pemp_meta_proto_attn_bwd_workspace_bytes(2, 5, c, hw, 2)
    assert 0 < small_map < N * c * 12 * 4 + lib.pemp_meta_proto_attn_bwd_workspace_bytes(2, 5, c, 25, 3) and other_p > 0
    assert lib.pemp_debug_bwd_path(0) in (0, 1)                                     # the switch exists and reports its previous value
    assert lib.pemp_cosine_match_bwd_workspace_bytes(64, 64, 512, 2601, 3) > 0
    assert lib.pemp_upsample_ce_workspace_bytes(64, 51, 51, 401, 401) >= 64 * 401 * 401 * 4
    assert lib.pemp_boundary_weight_workspace_bytes(64, 401, 401) >= 64 * 401 * 401 * 5
    assert lib.pemp_boundary_weight_workspace_bytes(0, 401, 401) == 0


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pemp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle\.|oracle/", src, flags=re.M), f"{f} uses the oracle"


def test_episode_generator_is_index_deterministic():
    spec = E.EpisodeSpec(shot=2, channels=16, h=9, w=9, H=65, W=65, out_h=50, out_w=70)
    a = E.make_batch(spec, [3, 4, 5])
    b = E.make_batch(spec, [5])
    assert torch.equal(a["feats1"][2 * 3:], b["feats1"]) and torch.equal(a["qry_msk"][2], b["qry_msk"][0])
    assert a["sup_mask"].shape == (3, 2, 2, 65, 65) and a["qry_msk"].dtype == torch.uint8
    assert set(np.unique(a["qry_msk"].numpy())) <= {0, 1, 255}
    assert torch.equal(a["sup_mask"][:, :, 0] + a["sup_mask"][:, :, 1], torch.ones(3, 2, 65, 65))
    assert a["cls"].min() >= spec.cls_lo and a["cls"].max() <= spec.cls_hi


def test_sharding_covers_each_episode_once():
    for world in (1, 2, 4, 8):
        got = sorted(i for r in range(world) for i in pdist.shard_indices(37, r, world, first=5))
        assert got == list(range(5, 42))


def test_miou_matches_reference_formula():
    from pemp_b200.metrics import miou_from_stat
    from oracle import restate as O
    from conftest import golden
    g = golden("metric_random")
    mi, mm = miou_from_stat(g["stat"], g["labels"])
    bi, bm = miou_from_stat(g["stat"], g["labels"], binary=True)
    assert np.array_equal(mi, g["miou"]) and mm == float(g["miou_mean"])
    assert np.array_equal(bi, g["biou"]) and bm == float(g["biou_mean"])
    assert np.array_equal(O.miou(g["stat"], g["labels"])[0], mi)


def _unwrapped(fn):
    while hasattr(fn, "__wrapped__"):
        fn = fn.__wrapped__
    return fn


def test_dropin_binds_reference_classes():
    """Where the reference tree is present (the build container, or the staged copy on the GPU box), patch() must rebind its
    methods to ours - wrapped in the reference module's own Sacred `capture`, so injected config reaches them - and unpatch()
    restore them."""
    from oracle import ref_import as R
    if not R.available():
        pytest.skip("reference tree not present on this machine")
    from pemp_b200 import dropin, heads
    s1 = R.module("networks.pemp_stage1").PEMPStage1
    pa = R.module("networks.panet").PANet
    pf = R.module("networks.pfenet")
    cm = R.module("core.metrics")
    before = (s1.mpm, s1.forward, pa.alignLoss, pf.Weighted_GAP, cm.FewShotMetric, pf.PFENet.forward)
    dropin.patch()
    try:
        assert _unwrapped(s1.mpm) is heads.mpm and _unwrapped(s1.forward) is heads.pemp_stage1_forward
        assert _unwrapped(pa.alignLoss) is heads.alignLoss and pf.Weighted_GAP is heads.Weighted_GAP
        assert pf.prior_mask is heads.prior_mask
        assert cm.FewShotMetric.__module__ == "pemp_b200.metrics"
        net = R.head_only("pemp_stage1", torch.zeros(2, 8, 4, 4), torch.rand(8, 6))
        with pytest.raises(ValueError, match="CUDA"):         # our code ran; CPU tensors are refused
            net(torch.zeros(1, 1, 1, 9, 9), torch.zeros(1, 1, 2, 9, 9), torch.zeros(1, 1, 1, 9, 9))
        # PFENet.forward is now the reference's own source with its inline prior block replaced by one call
        fwd = pf.PFENet.forward
        assert "pemp_b200 splice" in fwd.__code__.co_filename
        assert "_pemp_prior_mask" in fwd.__code__.co_names and "bmm" not in fwd.__code__.co_names
        assert {"layer4", "down_query", "init_merge", "res2", "cls"} <= set(fwd.__code__.co_names)      # the rest is untouched
    finally:
        dropin.unpatch()
    assert (s1.mpm, s1.forward, pa.alignLoss, pf.Weighted_GAP, cm.FewShotMetric, pf.PFENet.forward) == before
    assert not hasattr(pf, "prior_mask")
    assert "bmm" in pf.PFENet.forward.__code__.co_names


def test_dropin_receives_the_sacred_config():
    """`net.dist_scalar` is injected by Sacred's `capture` by parameter name (pemp_stage1.py:232-234): a patched model must
    evaluate with the configured value, not a hard-coded 20."""
    from oracle import ref_import as R
    if not R.available():
        pytest.skip("reference tree not present on this machine")
    from pemp_b200 import dropin, heads, ops
    mod = R.module("networks.pemp_stage1")
    seen = {}
    real = ops.cosine_match

    def spy(qry, fg, bg, scalar=20.0, **kw):
        seen["scalar"] = scalar
        raise RuntimeError("stop here")
    old_cfg = mod.net_ingredient.cfg.get("dist_scalar")
    dropin.patch()
    ops.cosine_match = spy
    try:
        mod.net_ingredient.cfg["dist_scalar"] = 7
        with pytest.raises(RuntimeError, match="stop here"):
            mod.PEMPStage1.compute_similarity(None, torch.zeros(1, 4), torch.zeros(1, 4), torch.zeros(1, 4, 2, 2))
        assert seen["scalar"] == 7
        assert heads._scalar(None, None) == 20 and heads._scalar(None, 3) == 3
    finally:
        ops.cosine_match = real
        mod.net_ingredient.cfg["dist_scalar"] = old_cfg
        dropin.unpatch()


def test_episode_screen_table_is_used_and_never_silently_bypassed():
    spec = E.EpisodeSpec(shot=5, stages=2)
    acc = E.screened_indices("pemp_stage2", spec)
    st = E.screen_stats("pemp_stage2", spec)
    assert st["threshold"] == 2e-5 and st["candidates"] >= 512 and len(acc) == st["candidates"] - st["rejected"]
    assert len(acc) >= 512                                   # 8 ranks x 64 disjoint screened episodes
    shards = [E.screened_indices("pemp_stage2", spec, 64, start=r, step=8) for r in range(8)]
    flat = [i for sh in shards for i in sh]
    assert len(set(flat)) == 512 and set(flat) <= set(acc)
    with pytest.raises(KeyError):
        E.screened_indices("pemp_stage2", E.EpisodeSpec(shot=4, stages=2), 8)      # no table for this stream: loud
    with pytest.raises(ValueError):
        E.screened_indices("pemp_stage2", spec, 10 ** 6)


def test_bench_reference_arm_runs_the_reference_code(tmp_path):
    """`bench.py --impl reference` prints one JSON line produced by the reference's own classes (kind "reference") where its
    files are present, on the named workload."""
    import json
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "stage1_1shot",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    out = json.loads(lines[0])
    from oracle import ref_import as R
    assert out["impl"] == "reference" and out["unit"] == "episodes/s" and out["value"] > 0
    assert out["cpu_baseline"]["kind"] == ("reference" if R.available() else "port")
    assert out["e2e"]["h2d_bytes_per_step"] == 0 and "stage1_1shot" in out["config"]["workload"]


_GLOO_WORKER = r"""
import os, sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])
from pemp_b200 import dist as pdist, episodes as E
from oracle import restate as O
rank, _, world = pdist.init(backend="gloo")
spec = E.EpisodeSpec(shot=1, channels=8, h=5, w=5, H=33, W=33, out_h=33, out_w=33, stages=1)
n = 6
stat = np.zeros((spec.classes + 1, 3), np.int64)
for i in pdist.shard_indices(n, rank, world):
    ep = E.make_episode(spec, i)
    pred = (torch.rand(1, 33, 33, generator=torch.Generator().manual_seed(i)) > 0.5).numpy().astype(np.uint8)
    O.few_shot_stat(pred, ep["qry_msk"].numpy(), [ep["cls"]], spec.classes, stat)
t = torch.from_numpy(stat)
pdist.all_reduce_stat(t)
elapsed = pdist.max_over_ranks(float(rank + 1), "cpu")
pdist.barrier()
if rank == 0:
    np.save(sys.argv[2], t.numpy())
    assert elapsed == float(world)
"""


def test_gloo_world2_stat_allreduce_equals_single_rank(tmp_path):
    """Sharded counts + all-reduce == single-process counts, bit for bit (CPU emulation of the N>1 path)."""
    from oracle import restate as O
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    out = tmp_path / "stat.npy"
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script), ROOT, str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    spec = E.EpisodeSpec(shot=1, channels=8, h=5, w=5, H=33, W=33, out_h=33, out_w=33, stages=1)
    stat = np.zeros((spec.classes + 1, 3), np.int64)
    for i in range(6):
        ep = E.make_episode(spec, i)
        pred = (torch.rand(1, 33, 33, generator=torch.Generator().manual_seed(i)) > 0.5).numpy().astype(np.uint8)
        O.few_shot_stat(pred, ep["qry_msk"].numpy(), [ep["cls"]], spec.classes, stat)
    assert np.array_equal(np.load(out), stat)
