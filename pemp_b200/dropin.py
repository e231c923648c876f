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
  networks.pfenet                 : Weighted_GAP (module function), `prior_mask` added to the module, and PFENet.forward
                                    re-compiled from the reference's OWN source with its inline prior block
                                    (pfenet.py:201-229) replaced by one `prior_mask(...)` call - nothing of the reference is
                                    copied into this package; if the block's marker lines are not found `patch()` raises
  core.metrics                    : FewShotMetric (and the name imported into core.base_trainer)
  networks.backbones              : ResNetCM.comm, VGG16CM.comm (only with `patch(comm=True)`: the "next" row of SURVEY 8f)
The encoders (`self.encoder`) are otherwise untouched: the backbone stays on stock PyTorch.
"""
import importlib
import inspect
import sys
import textwrap

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


_PRIOR_FIRST = "corr_query_mask_list = []"                                            # pfenet.py:201
_PRIOR_LAST = "corr_query_mask = torch.cat(corr_query_mask_list, 1).mean(1).unsqueeze(1)"   # pfenet.py:229


def splice_pfenet_forward(mod, precision=None):
    """-> a `PFENet.forward` compiled from the reference's own source text (read at patch time, never stored here) in which the
    inline prior block - everything from `corr_query_mask_list = []` through the shot mean (pfenet.py:201-229) - is replaced
    by `corr_query_mask = prior_mask(query_feat_4, final_supp_list, mask_list, out_hw=<feat-3 size>)`.  The rest of the
    forward (backbone, pyramid, classifier) stays the reference's stock PyTorch code, line for line."""
    src = textwrap.dedent(inspect.getsource(mod.PFENet.forward))
    lines = src.split("\n")
    first = [i for i, ln in enumerate(lines) if ln.strip() == _PRIOR_FIRST]
    last = [i for i, ln in enumerate(lines) if ln.strip() == _PRIOR_LAST]
    if len(first) != 1 or len(last) != 1 or last[0] <= first[0]:
        raise RuntimeError("pemp_b200.dropin: the prior block of networks/pfenet.py (lines 201-229 of the reference) was not "
                           "found in PFENet.forward; this reference version cannot be spliced automatically")
    indent = lines[first[0]][:len(lines[first[0]]) - len(lines[first[0]].lstrip())]
    call = (f"{indent}corr_query_mask = _pemp_prior_mask(query_feat_4, final_supp_list, mask_list, "
            f"out_hw=(query_feat_3.size(2), query_feat_3.size(3)))")
    new_src = "\n".join(lines[:first[0]] + [call] + lines[last[0] + 1:])
    # drop decorators of the original definition (e.g. @net_ingredient.capture): the caller re-applies capture
    new_src = new_src[new_src.index("def forward"):]
    # compiled against the module's OWN globals (not a copy): `Weighted_GAP`, `F`, `nn`, ... resolve exactly as in the original
    # forward, including later rebinding; the one new name, `_pemp_prior_mask`, is installed on the module by `patch()`
    hook = heads.prior_mask if precision is None else (
        lambda q, s_, m, out_hw=None: heads.prior_mask(q, s_, m, precision=precision, out_hw=out_hw))
    _set(mod, "_pemp_prior_mask", hook)
    local = {}
    exec(compile(new_src, f"<pemp_b200 splice of {getattr(mod, '__file__', 'networks/pfenet.py')}>", "exec"), vars(mod), local)
    fn = local["forward"]
    fn.__qualname__ = "PFENet.forward"
    return fn


def _capture(mod, fn):
    """Re-apply the reference module's Sacred ingredient to a replacement so that config values the reference injects by
    parameter name (`dist_scalar`, `protos`; pemp_stage1.py:165-167, 232-234) reach it exactly as they reach the original."""
    ing = getattr(mod, "net_ingredient", None)
    if ing is None or not hasattr(ing, "capture"):
        return fn
    try:
        return ing.capture(fn)
    except Exception:
        return fn


def patch(models=("pemp_stage1", "pemp_stage2", "baseline", "panet", "pfenet"), metric=True, comm=False, prior_precision=None):
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
            if hasattr(mod, "PFENet"):
                _set(mod.PFENet, "forward", splice_pfenet_forward(mod, prior_precision))
            continue
        cls_name, methods = table[name]
        cls = getattr(mod, cls_name)
        for attr, fn in methods.items():
            _set(cls, attr, _capture(mod, fn))
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
