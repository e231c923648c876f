"""Episode runners for the CPU arm of `bench.py` and for parity blocks — TEST INFRASTRUCTURE ONLY.

`runner(workload, spec)` returns an object with `.kind` and `.episode(batch)`:
  * kind "reference": the reference's OWN code (`oracle/ref_import.py`: `/root/reference` in the build container, the staged
    unmodified copy `oracle/_ref/reference/` on the GPU box) driven exactly as its evaluator does
    (`entry/pemp_stage2.py:58-65`, `entry/baseline.py:46-52`, `entry/panet.py:51-57`, `networks/pfenet.py:195-231`,
    `core/metrics.py:9-23`) with a stub encoder that returns the synthetic features;
  * kind "port": `oracle/restate.py`, used only where the reference files are absent.
`.episode(batch)` takes a one-episode CPU batch (`pemp_b200.episodes.make_batch(spec, [i])`, or the PFENet dict of `bench.py`)
and returns dict(mask [Q,H',W'] int64, stat [(C+1),3] int64, margin (smallest |fg - bg| logit gap), + workload extras).
Nothing under `pemp_b200/` imports this module.
"""
import numpy as np
import torch

from oracle import ref_import as R
from oracle import restate as O
from pemp_b200 import episodes as E


def _margin(logits):
    return float((logits[:, 1] - logits[:, 0]).abs().min())


class _Base:
    def __init__(self, workload, spec):
        self.workload, self.spec = workload, spec
        self.kind = "reference" if R.available() else "port"
        if workload.startswith("pemp"):
            self.ctr1, self.ctr2 = E.make_ctr(spec, 1), E.make_ctr(spec, 2)


class PempRunner(_Base):
    """workload 'pemp_stage2': stage-1 head -> prior -> stage-2 head -> mask -> counts; 'pemp_stage1': one head."""

    def episode(self, b):
        spec = self.spec
        S, Q = spec.shot, spec.query
        out_shape = (spec.out_h, spec.out_w)
        two = self.workload == "pemp_stage2"
        if self.kind == "port":
            if two:
                r = O.stage2_episode_batch(b["feats1"], b["feats2"], b["sup_mask"], self.ctr1, self.ctr2, 1, S, Q,
                                           b["qry_msk"].numpy(), b["cls"].numpy(), spec.classes)
                return {"prior": r["prior"], "mask": r["mask"], "stat": r["stat"],
                        "margin": min(_margin(r["stage1"]["logits"]), _margin(r["stage2"]["logits"]))}
            s1 = O.pemp_head(b["feats1"], b["sup_mask"], self.ctr1, 1, S, Q, out_shape=out_shape)
            mask = O.argmax2(s1["logits"])
            return {"mask": mask, "margin": _margin(s1["logits"]),
                    "stat": O.few_shot_stat(mask.numpy(), b["qry_msk"].numpy(), b["cls"].numpy(), spec.classes)}
        si, qi = R.dummy_images(1, S, Q, spec.H, spec.W)
        fm = R.few_shot_metric(spec.classes)
        with torch.no_grad():
            net1 = R.head_only("pemp_stage1", b["feats1"], self.ctr1)
            if not two:
                logits = net1(si, b["sup_mask"], qi, out_shape)
                mask = logits.argmax(dim=1)
                fm.update(mask.numpy(), b["qry_msk"].numpy(), b["cls"])
                return {"mask": mask, "stat": fm.stat.astype(np.int64), "margin": _margin(logits)}
            l1 = net1(si, b["sup_mask"], qi)                                   # entry/pemp_stage2.py:59
            prior = l1.argmax(dim=1, keepdim=True)                             # :60
            net2 = R.head_only("pemp_stage2", b["feats2"], self.ctr2)
            l2 = net2(si, b["sup_mask"], qi, prior, out_shape)                 # :62
            mask = l2.argmax(dim=1)                                            # :64
        fm.update(mask.numpy(), b["qry_msk"].numpy(), b["cls"])                 # core/base_trainer.py:82
        return {"prior": prior[:, 0], "mask": mask, "stat": fm.stat.astype(np.int64), "margin": min(_margin(l1), _margin(l2))}


class BaselineRunner(_Base):
    """workload 'baseline' (`Baseline.forward`) / 'panet' (`PANet.forward` incl. `alignLoss`)."""

    def episode(self, b):
        spec = self.spec
        S, Q = spec.shot, spec.query
        out_shape = (spec.out_h, spec.out_w)
        if self.kind == "port":
            head = O.panet_head if self.workload == "panet" else O.baseline_head
            r = head(b["feats1"], b["sup_mask"], 1, S, Q, out_shape=out_shape)
            mask = O.argmax2(r["logits"])
            out = {"mask": mask, "stat": O.few_shot_stat(mask.numpy(), b["qry_msk"].numpy(), b["cls"].numpy(), spec.classes),
                   "margin": _margin(r["logits"])}
            if self.workload == "panet":
                out["align_loss"] = float(r["align_loss"])
                out["margin"] = min(out["margin"], _margin(r["pred_lowres"]))
            return out
        si, qi = R.dummy_images(1, S, Q, spec.H, spec.W)
        net = R.head_only(self.workload, b["feats1"])
        fm = R.few_shot_metric(spec.classes)
        seen = {}
        if self.workload == "panet":            # alignLoss thresholds the LOW-RES prediction (panet.py:164): note its margin
            inner = net.alignLoss

            def spy(qry_fts, pred, *a, **k):
                seen["pred_lowres"] = pred
                return inner(qry_fts, pred, *a, **k)
            net.alignLoss = spy
        with torch.no_grad():
            res = net(si, b["sup_mask"], qi, out_shape)
        logits, loss = res if isinstance(res, tuple) else (res, None)
        mask = logits.argmax(dim=1)
        fm.update(mask.numpy(), b["qry_msk"].numpy(), b["cls"])
        out = {"mask": mask, "stat": fm.stat.astype(np.int64), "margin": _margin(logits)}
        if loss is not None:
            out["align_loss"] = float(loss)
            out["margin"] = min(out["margin"], _margin(seen["pred_lowres"]))
        return out


class PfenetRunner(_Base):
    """workload 'pfenet': the prior block (pfenet.py:201-231) + `Weighted_GAP` of every shot (pfenet.py:197-198).
    batch: q4 [1,C,sp,sp], s4 [S,1,C,sp,sp], masks [S,1,1,H,W], supp_feat [S,1,c,sp,sp]."""

    def episode(self, b):
        S = b["s4"].shape[0]
        sp = b["q4"].shape[-1]
        with torch.no_grad():
            if self.kind == "port":
                prior = O.pfenet_prior(b["q4"], list(b["s4"]), list(b["masks"]))
                small = [O.bilinear_upsample(m, sp, sp) for m in b["masks"]]
                gap = torch.stack([O.weighted_gap(b["supp_feat"][s], small[s]) for s in range(S)])
            else:
                prior = R.pfenet_prior(b["q4"], list(b["s4"]), list(b["masks"]), (sp, sp), (sp, sp))
                small = [torch.nn.functional.interpolate(m, size=(sp, sp), mode="bilinear", align_corners=True) for m in b["masks"]]
                gap = torch.stack([R.weighted_gap(b["supp_feat"][s], small[s]) for s in range(S)])
        return {"prior": prior, "gap": gap}


def runner(workload, spec=None):
    if workload.startswith("pemp"):
        return PempRunner(workload, spec)
    if workload in ("baseline", "panet"):
        return BaselineRunner(workload, spec)
    if workload == "pfenet":
        return PfenetRunner(workload, spec)
    raise ValueError(workload)
