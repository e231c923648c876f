"""Install the B200 head behind the reference's own classes.

    import pemp_b200.dropin as dropin
    dropin.patch()            # after `sys.path` contains the reference root
    ...                       # reference code runs unchanged: entry/*.py, core/base_trainer.py
    dropin.unpatch()

What is rebound:
  networks.pemp_stage1.PEMPStage1 : forward, mpm, compute_similarity
  networks.pemp_stage2.PEMPStage2 : forward, mpm, compute_similarity
  networks.baseline.Baseline      : forward, compute_similarity
  networks.panet.PANet            : forward, compute_similarity, alignLoss
  networks.pfenet                 : Weighted_GAP (module function) + `prior_mask` added to the module
  core.metrics                    : FewShotMetric (and the name imported into core.base_trainer)
  networks.backbones              : ResNetCM.comm, VGG16CM.comm (only with `patch(comm=True)`: the "next" row of SURVEY 8f)
The encoders (`self.encoder`) are otherwise untouched: the backbone stays on stock PyTorch.
"""
import importlib
import sys

from . import heads, metrics

_saved = []


def _set(obj, name, value):
    _saved.append((obj, name, getattr(obj, name, _MISSING)))
    setattr(obj, name, value)


_MISSING = object()


def _maybe(module_name):
    try:
        return importlib.import_module(module_name)
    except Exception:
        return None


def patch(models=("pemp_stage1", "pemp_stage2", "baseline", "panet", "pfenet"), metric=True, comm=False):
    if _saved:
        return
    if comm:
        bb = _maybe("networks.backbones")
        for cls_name in ("ResNetCM", "VGG16CM"):
            if bb is not None and hasattr(bb, cls_name):
                _set(getattr(bb, cls_name), "comm", heads.comm)
    table = {
        "pemp_stage1": ("PEMPStage1", {"forward": heads.pemp_stage1_forward, "mpm": heads.mpm,
                                       "compute_similarity": heads.compute_similarity}),
        "pemp_stage2": ("PEMPStage2", {"forward": heads.pemp_stage2_forward, "mpm": heads.mpm,
                                       "compute_similarity": heads.compute_similarity}),
        "baseline": ("Baseline", {"forward": heads.baseline_forward, "compute_similarity": heads.compute_similarity}),
        "panet": ("PANet", {"forward": heads.panet_forward, "compute_similarity": heads.compute_similarity,
                            "alignLoss": heads.alignLoss}),
    }
    for name in models:
        mod = _maybe(f"networks.{name}")
        if mod is None:
            continue
        if name == "pfenet":
            _set(mod, "Weighted_GAP", heads.Weighted_GAP)
            _set(mod, "prior_mask", heads.prior_mask)
            continue
        cls_name, methods = table[name]
        cls = getattr(mod, cls_name)
        for attr, fn in methods.items():
            _set(cls, attr, fn)
    if metric:
        cm = _maybe("core.metrics")
        if cm is not None:
            _set(cm, "FewShotMetric", metrics.FewShotMetric)
        bt = sys.modules.get("core.base_trainer")
        if bt is not None and hasattr(bt, "FewShotMetric"):
            _set(bt, "FewShotMetric", metrics.FewShotMetric)


def unpatch():
    while _saved:
        obj, name, old = _saved.pop()
        if old is _MISSING:
            delattr(obj, name)
        else:
            setattr(obj, name, old)
