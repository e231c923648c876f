#!/usr/bin/env python
"""Benchmark of the prototype-matching head (BASELINE.json metric: episodes/s of the prototype head; HBM GB/s and
tensor-pipe fraction of the dominant kernels).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one batch of synthetic episodes through the whole head path of the workload (BASELINE.json configs):
  stage2_5shot  (default, the configuration north_star's target is quoted on; config 3)  PEMP Stage-2 ResNet-50 5-shot head:
                K0 nearest masks, stage-1 head (K2 meta-prototype attention, K3 cosine matching, K4 up-sample + argmax ->
                prior), stage-2 head (K2, K3, K4) and the K10 IoU counts; c = 512, 51x51 features, 401x401 masks
                (entry/pemp_stage2.py:58-65).
  stage1_1shot  (config 2)  PEMP Stage-1 1-shot head (K0, K2, K3, K4 + K10).
  baseline_1shot (config 1) Baseline ResNet-50 1-shot head: K6 full-resolution MAP, K3, K4 + K10 (baseline.py:97-118).
  panet_5shot_coco (config 4) PANet 5-shot head with the prototype-alignment reverse pass (K6, K3, K4 + K10, K7;
                panet.py:96-194), COCO-20i-shaped labels: 80 classes.
  pfenet_5shot  (config 5)  PFENet 5-shot prior-mask stress test: mask resize, K8 Weighted_GAP per shot, K9 prior contraction
                3600 x 3600 x 2048 per shot on tcgen05 (pfenet.py:191-231).
Episodes are independent: N ranks each process `--batch` episodes per step (weak scaling); the only collective is ONE
all-reduce of the (C+1) x 3 int64 count table per evaluation round (= the K timed steps), inside the timed region.

The line printed by rank 0 follows the driver contract: `value` = whole-job episodes/s with inputs resident in HBM,
`e2e` = the same through the host-facing call with pinned HOST buffers (H2D of every input and D2H of the result inside the
timed region), `roofline` for the dominant kernel timed with CUDA events inside the timed region, `cpu_baseline` = the
reference's own code (staged copy, `oracle/_ref`) or its port timed on this box's cores (N = 1 only), `parity` = the GPU
results of the timed batch against that same reference run (masks, counts), `roofline_extra` (headline workload, N = 1) =
the kernels of the other configs timed in the same process.  `--impl reference` times the CPU path alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

_JSON_OUT = sys.stdout

WORKLOADS = {
    "stage2_5shot": dict(kind="pemp", screen="pemp_stage2", shot=5, stages=2, batch=64, config=3,
                         desc="PEMP Stage-2 ResNet-50 5-shot prototype head (stage-1 head -> prior -> stage-2 head -> IoU)"),
    "stage1_1shot": dict(kind="pemp", screen="pemp_stage1", shot=1, stages=1, batch=64, config=2,
                         desc="PEMP Stage-1 ResNet-50 1-shot prototype head"),
    "baseline_1shot": dict(kind="baseline", screen="baseline", shot=1, stages=1, batch=64, config=1, align=False,
                           desc="Baseline ResNet-50 1-shot head (full-resolution masked average pooling + cosine matching + IoU)"),
    "panet_5shot_coco": dict(kind="baseline", screen="panet", shot=5, stages=1, batch=64, config=4, align=True, classes=80,
                             desc="PANet 5-shot head with prototype-alignment reverse pass, COCO-20i-shaped (80 classes)"),
    "pfenet_5shot": dict(kind="pfenet", shot=5, batch=8, config=5, C=2048, sp=60, image=473, c_mid=256,
                         desc="PFENet ResNet-50 5-shot prior-mask stress test (3600 x 3600 x 2048 contraction per shot) + Weighted_GAP"),
}
MARGIN = 2e-5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="stage2_5shot", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="episodes per GPU per step (default: per workload)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--sustained-steps", type=int, default=200)
    ap.add_argument("--prior-precision", default="bf16x3", choices=["bf16", "bf16x3"], help="pfenet_5shot: K9 path")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay of the step")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skips the CPU arm AND the parity block that rides on it")
    ap.add_argument("--no-extra", action="store_true", help="skip roofline_extra (kernels of the other configs)")
    ap.add_argument("--cpu-seconds", type=float, default=25.0)
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1650.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def window(self, t0, t1):
        return [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 7]

    def stop(self, t0, t1, extra=None):
        """Samples taken inside [t0, t1] (the timed region); `extra` = (t0, t1) of the sustained loop (the same step repeated
        >= 200 times right after it), used when the timed region was shorter than the sampling period."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows, where = self.window(t0, t1), "timed region"
        if len(rows) < 2 and extra is not None:
            rows, where = self.window(t0, extra[1]), "timed region + the sustained loop that follows it (same step, same load)"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "sampled_during": where, "reasons": reasons}


# ------------------------------------------------------------------------------------------------ workloads
def pick_indices(screen, spec, B, rank, world):
    """B screened episode indices for this rank (disjoint over ranks while the table lasts, else the accepted set repeated)."""
    from pemp_b200 import episodes as E
    acc = E.screened_indices(screen, spec)
    if len(acc) >= B * world:
        return acc[rank::world][:B], False
    return [acc[(rank * B + i) % len(acc)] for i in range(B)], True


class PempWorkload:
    """stage2_5shot / stage1_1shot."""

    def __init__(self, args, w, rank, world, dev):
        from pemp_b200 import episodes as E
        from pemp_b200.evaluator import PEMPStage2Pipeline
        self.w, self.dev = w, dev
        self.spec = spec = E.EpisodeSpec(shot=w["shot"], stages=w["stages"])
        self.B = B = args.batch or w["batch"]
        self.S, self.Q, self.c, self.h, self.wd = spec.shot, spec.query, spec.channels, spec.h, spec.w
        self.indices, self.repeated = pick_indices(w["screen"], spec, B, rank, world)
        self.host = E.make_batch(spec, self.indices)                   # CPU tensors, the reference's batch layout
        self.ctr1, self.ctr2 = E.make_ctr(spec, 1).to(dev), E.make_ctr(spec, 2).to(dev)
        self.pipe = PEMPStage2Pipeline(self.ctr1, self.ctr2, spec.classes)
        self.batch = {k: v.to(dev) for k, v in self.host.items()}
        self.stages = w["stages"]
        S, c, hw = self.S, self.c, self.h * self.wd
        self.dominant = {"kernel": "mpa_tma_kernel (+ finalize; K2 meta_proto_attn)", "bound": "hbm",
                         "per_launch": B * (S * (c * hw + 2 * hw) * 4 + 2 * c * 2 * spec.protos * 4), "launches_per_step": self.stages,
                         "traffic_key": f"mpa_tma_kernel:B{B}:S{S}:c{c}:hw{hw}"}
        k3 = self.Q * c * hw * 4 + c * 2 * spec.protos * 4 + self.Q * 2 * hw * 4
        k4 = 2 * hw * 4 + spec.H * spec.W
        self.alg_bytes_per_episode = self.stages * (self.dominant["per_launch"] / B + k3 + k4) + 2 * spec.out_h * spec.out_w
        self.result_bytes = (spec.classes + 1) * 3 * 8

    def new_result(self):
        return torch.zeros(self.spec.classes + 1, 3, dtype=torch.int64, device=self.dev)

    def _run(self, b, st, timer, n):
        S, Q, c, h, wd = self.S, self.Q, self.c, self.h, self.wd
        fa = b["feats1"].view(n, S + Q, c, h, wd)
        masks = b["labels"] if "labels" in b else b["sup_mask"]
        if self.stages == 2:
            fb = b["feats2"].view(n, S + Q, c, h, wd)
            return self.pipe.step(fa[:, :S], fa[:, S:], fb[:, :S], fb[:, S:], masks, b["qry_msk"], b["cls"], st, timer)
        from pemp_b200 import ops
        H, W = masks.shape[-2:]
        if masks.dtype == torch.uint8:
            low = ops.mask_nearest_labels(masks.view(n * S, H, W), h, wd).view(n * S, 2, h * wd)
        else:
            low = ops.mask_nearest(masks.view(n * S, 2, H, W), h, wd).view(n * S, 2, h * wd)
        return None, self.pipe._head(b["feats1"], low, self.ctr1, n, S, Q, tuple(b["qry_msk"].shape[-2:]), timer,
                                     hist=(b["qry_msk"], b["cls"], st))      # the stage-1 model: its own `ctr`

    def step(self, st, timer=None):
        return self._run(self.batch, st, timer, self.B)

    # ---- e2e: what a host-side caller hands over (label map instead of the expanded float masks, see ops.mask_nearest_labels)
    def host_chunks(self, chunk):
        h = dict(self.host)
        h["labels"] = (h.pop("sup_mask")[:, :, 0] > 0.5).to(torch.uint8)          # synthetic masks are binary and complementary
        per = self.S + self.Q
        out = []
        for i in range(0, self.B, chunk):
            out.append({k: (v[i * per:(i + chunk) * per] if k.startswith("feats") else v[i:i + chunk]).contiguous().pin_memory()
                        for k, v in h.items()})
        return out

    def step_chunk(self, b, st, n):
        return self._run(b, st, None, n)

    e2e_chunk = 8
    e2e_path = ("pinned host features / uint8 label maps -> double-buffered cudaMemcpyAsync (8-episode chunks) -> head kernels "
                "-> count table D2H")

    # ---- parity against the reference run of the same episodes
    def cpu_runner(self):
        from oracle import ref_run
        return ref_run.runner(self.w["screen"], self.spec)

    def host_episode(self, j):
        per = self.S + self.Q
        return {k: (v[j * per:(j + 1) * per] if k.startswith("feats") else v[j:j + 1]) for k, v in self.host.items()}

    def parity(self, refs):
        """refs: {episode position j -> reference outputs}.  GPU step on exactly those episodes vs the reference."""
        js = sorted(refs)
        per = self.S + self.Q
        sel = torch.tensor(js)
        rows = torch.cat([torch.arange(j * per, (j + 1) * per) for j in js])
        b = {k: (v[rows] if k.startswith("feats") else v[sel]).to(self.dev) for k, v in self.host.items()}
        st = self.new_result()
        prior, mask = self._run(b, st, None, len(js))
        torch.cuda.synchronize()
        want_mask = torch.cat([refs[j]["mask"] for j in js]).to(torch.uint8)
        flips = int((mask.cpu() != want_mask).sum())
        flips_prior = None
        if prior is not None:
            flips_prior = int((prior.cpu() != torch.cat([refs[j]["prior"] for j in js]).to(torch.uint8)).sum())
        want_stat = sum(refs[j]["stat"] for j in js)
        margins = [refs[j]["margin"] for j in js]
        return {"episodes_checked": len(js), "mask_pixels": int(want_mask.numel()), "mask_flips": flips, "prior_flips": flips_prior,
                "stat_equal": bool((st.cpu().numpy() == want_stat).all()), "min_reference_margin": min(margins),
                "episodes_below_margin": int(sum(m < MARGIN for m in margins)),
                "ok": flips == 0 and not flips_prior and bool((st.cpu().numpy() == want_stat).all())}


class BaselineWorkload:
    """baseline_1shot / panet_5shot_coco."""

    def __init__(self, args, w, rank, world, dev):
        from pemp_b200 import episodes as E
        self.w, self.dev = w, dev
        classes = w.get("classes", 20)
        self.spec = spec = E.EpisodeSpec(shot=w["shot"], stages=1, classes=classes, cls_lo=1, cls_hi=80 if classes == 80 else 5)
        self.B = B = args.batch or w["batch"]
        self.S, self.Q, self.c, self.h, self.wd = spec.shot, spec.query, spec.channels, spec.h, spec.w
        self.align = w["align"]
        self.indices, self.repeated = pick_indices(w["screen"], spec, B, rank, world)
        uniq = sorted(set(self.indices))
        made = E.make_batch(spec, uniq)
        per = self.S + self.Q
        pos = [uniq.index(i) for i in self.indices]
        rows = torch.cat([torch.arange(p * per, (p + 1) * per) for p in pos])
        sel = torch.tensor(pos)
        self.host = {k: (v[rows] if k.startswith("feats") else v[sel]) for k, v in made.items()}
        self.batch = {k: v.to(dev) for k, v in self.host.items()}
        S, Q, c, hw, HW = self.S, self.Q, self.c, self.h * self.wd, spec.H * spec.W
        k6 = S * (c * hw * 4 + 2 * HW * 4) + 2 * c * 4
        k7 = (Q + S) * c * hw * 4 + 2 * hw * 4 + S * HW * 4
        k3 = Q * c * hw * 4 + c * 2 * 4 + Q * 2 * hw * 4
        if self.align:
            self.dominant = {"kernel": "K7 panet_align (arg-max masks + pool_tma_kernel + cosine_tma_kernel<2> + upsample_ce_band_kernel)",
                             "bound": "hbm", "per_launch": B * k7, "launches_per_step": 1, "traffic_key": None}
        else:
            self.dominant = {"kernel": "K6 map_pool_fullres (adjoint resampler + pool_tma_kernel + finalize)", "bound": "hbm",
                             "per_launch": B * k6, "launches_per_step": 1, "traffic_key": None}
        self.alg_bytes_per_episode = k6 + k3 + 2 * hw * 4 + 2 * spec.out_h * spec.out_w + (k7 if self.align else 0)
        self.result_bytes = (classes + 1) * 3 * 8 + (4 if self.align else 0)

    def new_result(self):
        return torch.zeros(self.spec.classes + 1, 3, dtype=torch.int64, device=self.dev)

    def _run(self, b, st, timer, n):
        from pemp_b200 import ops
        S, Q, c, h, wd = self.S, self.Q, self.c, self.h, self.wd
        f5 = b["feats1"].view(n, S + Q, c, h, wd)
        if "labels" in b:                                     # the uint8 label map a host-side caller ships (e2e)
            H, W = b["labels"].shape[-2:]
            mask = b["labels"].view(n * S, H, W)
            mask_fg = mask
        else:                                                 # the float masks the loader emits (drop-in tensor contract)
            H, W = b["sup_mask"].shape[-2:]
            mask = b["sup_mask"].view(n * S, 2, H, W)
            mask_fg = mask[:, 0:1]
        k6 = lambda: ops.map_pool_fullres(f5[:, :S], mask, n, S)
        fgp, bgp = timer.bracket(k6) if (timer is not None and not self.align) else k6()
        pred = ops.cosine_match(f5[:, S:], fgp, bgp, 20.0)["pred"].view(n * Q, 2, h, wd)
        out_hw = tuple(b["qry_msk"].shape[-2:])
        m8 = ops.upsample_argmax_hist(pred, out_hw, b["qry_msk"].view(n * Q, *out_hw), b["cls"], st)
        loss = None
        if self.align:
            k7 = lambda: ops.panet_align(f5[:, S:], pred, f5[:, :S], mask_fg, Q)
            loss = timer.bracket(k7) if timer is not None else k7()
        return loss, m8

    def step(self, st, timer=None):
        return self._run(self.batch, st, timer, self.B)

    def host_chunks(self, chunk):
        h = dict(self.host)
        h["labels"] = (h.pop("sup_mask")[:, :, 0] > 0.5).to(torch.uint8)          # synthetic masks are binary and complementary
        per = self.S + self.Q
        return [{k: (v[i * per:(i + chunk) * per] if k.startswith("feats") else v[i:i + chunk]).contiguous().pin_memory()
                 for k, v in h.items()} for i in range(0, self.B, chunk)]

    def step_chunk(self, b, st, n):
        return self._run(b, st, None, n)

    e2e_chunk = 8
    e2e_path = ("pinned host features / uint8 label maps (K6 and K7 form the float planes on the fly) -> double-buffered "
                "cudaMemcpyAsync -> head kernels -> count table D2H")

    def cpu_runner(self):
        from oracle import ref_run
        return ref_run.runner(self.w["screen"], self.spec)

    def host_episode(self, j):
        per = self.S + self.Q
        return {k: (v[j * per:(j + 1) * per] if k.startswith("feats") else v[j:j + 1]) for k, v in self.host.items()}

    def parity(self, refs):
        js = sorted(refs)
        per = self.S + self.Q
        out = {"episodes_checked": len(js), "mask_flips": 0, "mask_pixels": 0, "stat_equal": True, "align_loss_rel_err": None,
               "min_reference_margin": min(refs[j]["margin"] for j in js),
               "episodes_below_margin": int(sum(refs[j]["margin"] < MARGIN for j in js))}
        for j in js:              # the reference's alignLoss is per call (test batch size 1): compare episode by episode
            b = {k: (v[j * per:(j + 1) * per] if k.startswith("feats") else v[j:j + 1]).to(self.dev) for k, v in self.host.items()}
            st = self.new_result()
            loss, m8 = self._run(b, st, None, 1)
            torch.cuda.synchronize()
            out["mask_flips"] += int((m8.cpu() != refs[j]["mask"].to(torch.uint8)).sum())
            out["mask_pixels"] += int(m8.numel())
            out["stat_equal"] = out["stat_equal"] and bool((st.cpu().numpy() == refs[j]["stat"]).all())
            if loss is not None:
                rel = abs(float(loss) - refs[j]["align_loss"]) / max(1.0, abs(refs[j]["align_loss"]))
                out["align_loss_rel_err"] = max(out["align_loss_rel_err"] or 0.0, rel)
        out["ok"] = out["mask_flips"] == 0 and out["stat_equal"] and (out["align_loss_rel_err"] is None or out["align_loss_rel_err"] < 1e-5)
        return out


class PfenetWorkload:
    """pfenet_5shot: the prior block + Weighted_GAP of one PFENet forward per episode."""

    def __init__(self, args, w, rank, world, dev):
        from pemp_b200 import episodes as E, ops
        self.w, self.dev = w, dev
        self.B = B = args.batch or w["batch"]
        self.S, self.C, self.sp, self.img, self.cm = w["shot"], w["C"], w["sp"], w["image"], w["c_mid"]
        self.precision = ops.PRIOR_BF16 if args.prior_precision == "bf16" else ops.PRIOR_BF16X3
        S, C, sp = self.S, self.C, self.sp
        g = torch.Generator(device=dev).manual_seed(E.REFERENCE_SEED + 17 * rank)
        cpu = torch.Generator().manual_seed(E.REFERENCE_SEED + 17 * rank)
        # q4, s4 = relu(N(0,1)) [2048, sp, sp]; binary rectangle masks at image size; down_supp output for Weighted_GAP
        self.batch = {
            "q4": torch.relu(torch.randn(B, C, sp, sp, device=dev, generator=g)),
            "s4": torch.relu(torch.randn(S, B, C, sp, sp, device=dev, generator=g)),
            "supp_feat": torch.relu(torch.randn(S, B, self.cm, sp, sp, device=dev, generator=g)),
        }
        masks = torch.zeros(S, B, 1, self.img, self.img)
        for s in range(S):
            for b in range(B):
                y0, y1, x0, x1 = E._rect(cpu, self.img, self.img)
                masks[s, b, 0, y0:y1, x0:x1] = 1.0
        self.batch["masks"] = masks.to(dev)
        hw = sp * sp
        self.flops_per_episode = 2.0 * C * (S * hw) * hw
        self.dominant = {"kernel": f"prior_tc_kernel (K9 pemp_prior_mask, {args.prior_precision}: pre-pass + tcgen05 GEMM + tail)",
                         "bound": "tensor", "per_launch": B * self.flops_per_episode, "launches_per_step": 1, "traffic_key": None,
                         "executed_multiplier": 3 if self.precision == ops.PRIOR_BF16X3 else 1}
        self.alg_bytes_per_episode = (1 + S) * C * hw * 4 + S * (self.cm * hw * 4 + self.img * self.img * 4)
        self.result_bytes = B * hw * 4 + S * B * self.cm * 4
        self.repeated, self.indices = False, list(range(B))
        self.spec = None

    def new_result(self):
        return None

    def _run(self, b, timer):
        from pemp_b200 import ops
        S, sp = self.S, self.sp
        n = b["q4"].shape[0]
        small = ops.bilinear_resize(b["masks"], (sp, sp))                                   # [S, n, 1, sp, sp]   pfenet.py:191,205
        gap = ops.weighted_gap(b["supp_feat"].view(S * n, self.cm, sp, sp), small.view(S * n, 1, sp, sp))     # pfenet.py:197-198
        k9 = lambda: ops.prior_mask(b["q4"], b["s4"], small[:, :, 0], self.precision)
        prior = timer.bracket(k9) if timer is not None else k9()
        return prior, gap.view(S, n, self.cm, 1, 1)

    def step(self, st, timer=None):
        return self._run(self.batch, timer)

    e2e_chunk = 2
    e2e_path = "pinned host layer-4 / down_supp features and masks -> double-buffered cudaMemcpyAsync (2-episode chunks) -> kernels -> prior maps + pooled vectors D2H"

    def host_chunks(self, chunk):
        out = []
        for i in range(0, self.B, chunk):
            out.append({"q4": self.batch["q4"][i:i + chunk].cpu().pin_memory(),
                        "s4": self.batch["s4"][:, i:i + chunk].contiguous().cpu().pin_memory(),
                        "supp_feat": self.batch["supp_feat"][:, i:i + chunk].contiguous().cpu().pin_memory(),
                        "masks": self.batch["masks"][:, i:i + chunk].contiguous().cpu().pin_memory()})
        return out

    def step_chunk(self, b, st, n):
        return self._run(b, None)

    def cpu_runner(self):
        from oracle import ref_run
        return ref_run.runner("pfenet")

    def host_episode(self, j):
        return {"q4": self.batch["q4"][j:j + 1].cpu(), "s4": self.batch["s4"][:, j:j + 1].cpu(),
                "supp_feat": self.batch["supp_feat"][:, j:j + 1].cpu(), "masks": self.batch["masks"][:, j:j + 1].cpu()}

    def parity(self, refs):
        from oracle import restate as O
        from pemp_b200 import ops
        js = sorted(refs)
        tol = 1e-5 if self.precision == ops.PRIOR_BF16X3 else 3e-3
        out = {"episodes_checked": len(js), "rowmax_tolerance": tol, "rowmax_nrel": 0.0, "prior_max_abs_err": 0.0, "prior_bound": 0.0,
               "gap_nrel": 0.0}
        for j in js:
            b = {k: (v[j:j + 1] if k == "q4" else v[:, j:j + 1]).contiguous() for k, v in self.batch.items()}
            small = ops.bilinear_resize(b["masks"], (self.sp, self.sp))
            prior, rowmax = ops.prior_mask(b["q4"], b["s4"], small[:, :, 0], self.precision, want_rowmax=True)
            gap = ops.weighted_gap(b["supp_feat"].view(self.S, self.cm, self.sp, self.sp), small.view(self.S, 1, self.sp, self.sp))
            torch.cuda.synchronize()
            h = self.host_episode(j)
            want_rm = torch.stack([O.pfenet_rowmax(h["q4"], h["s4"][s], small[s].cpu()) for s in range(self.S)])
            amp = float(1.0 / (want_rm.max(dim=2).values - want_rm.min(dim=2).values).min())
            out["rowmax_nrel"] = max(out["rowmax_nrel"], float((rowmax.cpu() - want_rm).abs().max() / want_rm.abs().max()))
            out["prior_max_abs_err"] = max(out["prior_max_abs_err"], float((prior.cpu() - refs[j]["prior"]).abs().max()))
            out["prior_bound"] = max(out["prior_bound"], 2 * tol * max(1.0, amp))
            want_gap = refs[j]["gap"].reshape(self.S, self.cm)
            out["gap_nrel"] = max(out["gap_nrel"], float((gap.cpu().reshape(self.S, self.cm) - want_gap).abs().max() / want_gap.abs().max()))
        out["ok"] = out["rowmax_nrel"] < tol and out["prior_max_abs_err"] < out["prior_bound"] and out["gap_nrel"] < 1e-5
        return out


def make_workload(args, rank, world, dev):
    w = WORKLOADS[args.workload]
    cls = {"pemp": PempWorkload, "baseline": BaselineWorkload, "pfenet": PfenetWorkload}[w["kind"]]
    return cls(args, w, rank, world, dev), w


def config_of(args, wl, w, world):
    cfg = {"workload": f"{args.workload} (BASELINE.json configs[{w['config']}]): {w['desc']}", "shot": w["shot"],
           "episodes_per_gpu_per_step": wl.B, "global_episodes_per_step": wl.B * world, "parallelism": f"episode-sharded dp{world}",
           "l2_policy": "inputs larger than L2 (>= 1 GB per step)"}
    if wl.spec is not None:
        s = wl.spec
        cfg.update({"query": s.query, "channels": s.channels, "feature_hw": [s.h, s.w], "image_hw": [s.H, s.W], "protos": s.protos,
                    "classes": s.classes})
        from pemp_b200 import episodes as E
        cfg["episodes"] = dict(E.screen_stats(w["screen"], s), note="margin-screened synthetic episodes (min reference |fg-bg| >= the threshold)",
                               repeated_to_fill_batch=wl.repeated)
    else:
        cfg.update({"channels": w["C"], "feature_hw": [w["sp"], w["sp"]], "image_hw": [w["image"], w["image"]],
                    "prior_precision": args.prior_precision})
    return cfg


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_arm(wl, budget_s, threads, max_episodes, warm=1):
    """The reference's own code (or its port) on the host cores, one episode per call (the reference's test batch size,
    `data_kits/datasets.py:23`), on the first episodes of the timed batch.  -> (rate, n, kind, {position: outputs})"""
    torch.set_num_threads(threads)
    runner = wl.cpu_runner()
    refs, times = {}, []
    t_start = time.perf_counter()
    j = 0
    while j < min(max_episodes, wl.B):
        b = wl.host_episode(j)
        t0 = time.perf_counter()
        with torch.no_grad():
            r = runner.episode(b)
        dt = time.perf_counter() - t0
        if j >= warm or max_episodes <= warm:
            times.append(dt)
        refs[j] = r
        j += 1
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    times.sort()
    return 1.0 / times[len(times) // 2], len(times), runner.kind, refs


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    dev = torch.device("cpu")
    args.batch = args.batch or {"pemp": 12, "baseline": 4, "pfenet": 3}[WORKLOADS[args.workload]["kind"]]
    wl, w = make_workload(args, 0, 1, dev)
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, wl.B - 1))
    torch.set_num_threads(threads)
    runner = wl.cpu_runner()
    with torch.no_grad():
        runner.episode(wl.host_episode(0))                                 # warm-up
        eps = [wl.host_episode(1 + i) for i in range(steps)]
        t0 = time.perf_counter()
        for b in eps:
            runner.episode(b)
        dt = time.perf_counter() - t0
    value = steps / dt
    cfg = config_of(args, wl, w, 1)
    cfg["episodes_per_gpu_per_step"] = cfg["global_episodes_per_step"] = 1
    cfg["parallelism"] = "single process, CPU"
    sample = f"{steps} steps of 1 episode each (the reference's test batch size), inputs resident in host memory"
    what = ("the reference's own code (unmodified files staged under oracle/_ref/reference, stub encoder returning the synthetic "
            "features) on torch CPU") if runner.kind == "reference" else \
           "oracle/restate.py (bit-exact CPU restatement of the reference head; reference files not staged on this box)"
    print(json.dumps({
        "impl": "reference", "metric": "episodes/sec of prototype head", "value": value, "unit": "episodes/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "episodes/s", "cores": threads, "kind": runner.kind, "sample": sample},
        "e2e": {"value": value, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": what + f" with all {threads} host threads"}), file=_JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    from pemp_b200 import _cabi, dist as pdist, ops
    from pemp_b200.evaluator import KernelTimer

    rank, local_rank, world = pdist.init()
    binding = pdist.bind_host_to_gpu(local_rank)
    sampler = ClockSampler(local_rank) if rank == 0 else None      # started early: nvidia-smi needs ~0.5 s to emit
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _cabi.check(_cabi.lib().pemp_check_device(), "pemp_check_device")
    wl, w = make_workload(args, rank, world, dev)
    B = wl.B
    peaks, how = measured_peaks()

    ar_events = []

    def loop(n, st, timer=None):
        for _ in range(n):
            wl.step(st, timer)
        if st is not None:                     # once per evaluation round, inside the timed region
            if world > 1:
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                pdist.all_reduce_stat(st)
                a1.record()
                ar_events.append((a0, a1))
            else:
                pdist.all_reduce_stat(st)

    # ---------------- value: device-resident inputs ------------------------------------------------------
    import gc
    scratch = wl.new_result()
    loop(max(args.warmup, 3), scratch, KernelTimer())       # warm-up through the same code path (event pool, allocator, NCCL)
    torch.cuda.synchronize()
    gc.collect()
    gc.disable()                                              # no collector pause inside a 17 ms timed region
    pdist.barrier()
    timer = KernelTimer()
    stat = wl.new_result()
    launches0 = ops.launch_count()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    loop(args.steps, stat, timer)
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    launches = ops.launch_count() - launches0
    gc.enable()
    pdist.barrier()
    ms_total = pdist.max_over_ranks(ev0.elapsed_time(ev1), dev)
    ms_mine = ev0.elapsed_time(ev1)
    allreduce_ms = ar_events[-1][0].elapsed_time(ar_events[-1][1]) if ar_events else None     # includes waiting for the slowest rank
    ms_per_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total / 1e3)

    # multi-rank result: the all-reduced table must equal the sum of the per-rank tables (bit for bit)
    allreduce_ok = None
    if stat is not None and world > 1:
        mine = wl.new_result()
        for _ in range(1):
            wl.step(mine)
        summed = mine.clone()
        pdist.all_reduce_stat(summed)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        torch.distributed.all_gather(gathered, mine)
        allreduce_ok = bool(torch.equal(summed, torch.stack(gathered).sum(dim=0))) and bool(torch.equal(summed * args.steps, stat))

    # ---------------- sustained: the same step >= 200 times back to back ---------------------------------
    sus_n = max(args.sustained_steps, args.steps, int(400.0 / ms_per_step))      # >= 0.4 s: several clock samples under load
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    pdist.barrier()
    e0 = time.time()
    s0.record()
    loop(sus_n, scratch)
    s1.record()
    torch.cuda.synchronize()
    e1 = time.time()
    pdist.barrier()
    ms_sus = pdist.max_over_ranks(s0.elapsed_time(s1), dev)
    sustained = {"value": B * world * sus_n / (ms_sus / 1e3), "unit": "episodes/s", "steps": sus_n, "ms_per_step": ms_sus / sus_n}
    clocks = sampler.stop(t_wall0, t_wall1, (e0, e1)) if sampler else None

    # roofline of the dominant kernel: algorithmic bytes (flops) per launch / mean launch duration inside the timed region
    dom = wl.dominant
    k_ms = timer.mean_ms()
    if dom["bound"] == "hbm":
        achieved, peak, unit = dom["per_launch"] / (k_ms / 1e3) / 1e9, peaks["hbm_gbs"], "GB/s"
    else:
        achieved, peak, unit = dom["per_launch"] / (k_ms / 1e3) / 1e12, peaks["bf16_tflops"], "TFLOP/s"
    traffic = None      # dram read+write bytes per launch from the committed ncu --set full capture of this shape
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath) and dom["traffic_key"]:
        traffic = json.load(open(tpath)).get(dom["traffic_key"])
    roofline = {"kernel": dom["kernel"], "bound": dom["bound"], "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                "traffic": traffic, "peak_source": how, "launch_ms": k_ms, "launches_timed": timer.count(),
                "share_of_step": k_ms * dom["launches_per_step"] / ms_per_step}
    if dom["bound"] == "hbm":
        roofline["algorithmic_bytes_per_launch"] = dom["per_launch"]
    else:
        roofline["algorithmic_flops_per_launch"] = dom["per_launch"]
        roofline["peak_note"] = "bf16 dense burst peak; sustained peak %.1f -> frac %.3f" % (
            peaks.get("bf16_tflops_sustained", peak), achieved / peaks.get("bf16_tflops_sustained", peak))
        roofline["executed_TFLOPs"] = achieved * dom["executed_multiplier"]
        roofline["tensor_pipe_frac_executed"] = achieved * dom["executed_multiplier"] / peak

    # ---------------- the same step replayed from a CUDA graph (PEMPStage2Pipeline.capture) ---------------
    graphed = None
    if args.workload == "stage2_5shot" and not args.no_graph:
        S, Q, c, h, wd = wl.S, wl.Q, wl.c, wl.h, wl.wd
        f1 = wl.batch["feats1"].view(B, S + Q, c, h, wd)
        f2 = wl.batch["feats2"].view(B, S + Q, c, h, wd)
        gstat = torch.zeros_like(stat)
        g = wl.pipe.capture(f1[:, :S], f1[:, S:], f2[:, :S], f2[:, S:], wl.batch["sup_mask"], wl.batch["qry_msk"], wl.batch["cls"], gstat)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        pdist.barrier()
        gstat.zero_()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            g.replay()
        pdist.all_reduce_stat(gstat)
        g1.record()
        torch.cuda.synchronize()
        pdist.barrier()
        ms_g = pdist.max_over_ranks(g0.elapsed_time(g1), dev)
        graphed = {"value": B * world * args.steps / (ms_g / 1e3), "unit": "episodes/s", "ms_per_step": ms_g / args.steps,
                   "kernels_per_graph": g.launches, "same_counts_as_eager": bool(torch.equal(gstat, stat))}
        del g

    # ---------------- e2e: pinned host inputs, H2D + D2H inside the timed region -------------------------
    e2e = None
    if not args.no_e2e:
        chunk = wl.e2e_chunk if B % wl.e2e_chunk == 0 else B
        chunks = wl.host_chunks(chunk)
        copy_stream = torch.cuda.Stream(device=dev)
        bufs = [{k: torch.empty_like(v, device=dev) for k, v in chunks[0].items()} for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        h2d_bytes = sum(v.numel() * v.element_size() for ch in chunks for v in ch.values())
        host_out = torch.zeros(max(wl.result_bytes, 8), dtype=torch.uint8).pin_memory()

        def e2e_step():
            st = wl.new_result()
            cur = torch.cuda.current_stream()
            outs = []
            for ci, hc in enumerate(chunks):
                b = bufs[ci % 2]
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[ci % 2])
                    for k, v in hc.items():
                        b[k].copy_(v, non_blocking=True)
                    ready[ci % 2].record(copy_stream)
                cur.wait_event(ready[ci % 2])
                outs.append(wl.step_chunk(b, st, chunk))
                free[ci % 2].record(cur)
            if st is not None:
                pdist.all_reduce_stat(st)
                flat = st.view(torch.uint8).reshape(-1)
            else:                           # pfenet: the prior maps and pooled vectors are the result
                flat = torch.cat([torch.cat((p.reshape(-1), g.reshape(-1))) for p, g in outs]).view(torch.uint8).reshape(-1)
            host_out[:flat.numel()].copy_(flat, non_blocking=True)
            cur.synchronize()
            return flat.numel()

        for f in free:
            f.record(torch.cuda.current_stream())
        d2h = e2e_step()
        torch.cuda.synchronize()
        pdist.barrier()
        a, bq = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.e2e_steps):
            e2e_step()
        bq.record()
        torch.cuda.synchronize()
        ms_e2e = pdist.max_over_ranks(a.elapsed_time(bq), dev)
        e2e = {"value": B * world * args.e2e_steps / (ms_e2e / 1e3), "unit": "episodes/s", "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "ms_per_step": ms_e2e / args.e2e_steps,
               "h2d_GBps_per_gpu": h2d_bytes / (ms_e2e / args.e2e_steps) / 1e6,
               "h2d_GBps_aggregate": h2d_bytes * world / (ms_e2e / args.e2e_steps) / 1e6,
               "limiter": ("host->device link: %.1f MB of inputs per episode cross PCIe (device time of the step is %.1f%% of the "
                           "e2e step)" % (h2d_bytes / B / 1e6, 100 * ms_per_step / (ms_e2e / args.e2e_steps))) + (
                   "" if world == 1 else "; %d ranks share the host side of the box (one NUMA node, shared PCIe uplinks): %.1f GB/s per "
                   "GPU against ~55 GB/s for one GPU alone" % (world, h2d_bytes / (ms_e2e / args.e2e_steps) / 1e6)),
               "host_binding": binding, "path": wl.e2e_path}
        del chunks, bufs

    # ---------------- CPU arm + parity of the timed batch against it (rank 0) ----------------------------
    cpu, parity = None, None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        kind = w["kind"]
        cap = {"pemp": B, "baseline": 6, "pfenet": 2}[kind] if world == 1 else {"pemp": 8, "baseline": 2, "pfenet": 1}[kind]
        budget = args.cpu_seconds if world == 1 else min(args.cpu_seconds, 10.0)
        rate, n, ckind, refs = cpu_arm(wl, budget, threads, cap)
        if world == 1:
            cpu = {"value": rate, "unit": "episodes/s", "cores": threads, "kind": ckind,
                   "sample": f"median of {n} single-episode calls (reference test batch size 1) on the first episodes of the timed batch, "
                             f"after 1 warm-up; {'the reference files staged under oracle/_ref' if ckind == 'reference' else 'oracle/restate.py'}"}
        parity = wl.parity(refs)
        parity["checker"] = ckind
        if allreduce_ok is not None:
            parity["allreduced_table_equals_sum_of_rank_tables"] = allreduce_ok
    elif rank == 0 and allreduce_ok is not None:
        parity = {"allreduced_table_equals_sum_of_rank_tables": allreduce_ok}

    # ---------------- kernels of the other configs, timed in this process (headline, N = 1) --------------
    extra = None
    if rank == 0 and world == 1 and args.workload == "stage2_5shot" and not args.no_extra:
        del wl.batch
        torch.cuda.empty_cache()
        from tools import kernel_bench
        extra = kernel_bench.run(only="K1 ,K3,K6,K7,K8,K9,K11,K12")

    if rank == 0:
        hbm_eps = peaks["hbm_gbs"] * 1e3 / (wl.alg_bytes_per_episode / 1e6)
        out = {"metric": "episodes/sec of prototype head", "value": value, "unit": "episodes/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32" if w["kind"] != "pfenet" else f"f32 in / {args.prior_precision} tensor-core products, f32 accumulate",
               "data": "synthetic", "config": config_of(args, wl, w, world), "clocks": clocks, "e2e": e2e, "sustained": sustained,
               "graphed": graphed, "gpu_launches": launches, "allreduce": None if allreduce_ms is None else {
                   "ms_on_rank0": allreduce_ms, "rank0_timed_region_ms": ms_mine, "max_over_ranks_ms": ms_total,
                   "note": "one all-reduce of the count table per evaluation round, issued after the K steps; its time on a rank "
                           "includes waiting for the slowest rank"}, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
               "episode_roofline": {"algorithmic_MB_per_episode": wl.alg_bytes_per_episode / 1e6, "episodes_per_s_at_hbm_peak": hbm_eps,
                                    "frac": (value / world) / hbm_eps, "note": "whole step against the HBM roofline of its algorithmic bytes"},
               "roofline_extra": extra}
        if w["kind"] == "pfenet":
            tf = wl.flops_per_episode * value / world / 1e12
            out["episode_roofline"] = {"algorithmic_GFLOP_per_episode": wl.flops_per_episode / 1e9, "achieved_TFLOPs_whole_step": tf,
                                       "frac_of_bf16_peak": tf / peaks["bf16_tflops"],
                                       "note": "whole step (mask resize + K8 + K9) against the bf16 tensor peak on algorithmic FLOPs"}
        print(json.dumps(out), file=_JSON_OUT, flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly one JSON line.  Native libraries write there too (NCCL prints its "NCCL version ..." banner on
    # stdout at every debug level the image or the caller may set), so file descriptor 1 is pointed at stderr for the whole
    # run and the JSON line goes out through a private copy of the original stdout.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
