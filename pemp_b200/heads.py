"""Host-side mirror of the reference's model-facing prototype-head API.

Function names, argument order and return shapes follow the reference methods they replace, so they can be
bound onto the reference classes unchanged (`pemp_b200.dropin.patch`) or used through the small stand-alone
classes at the bottom.  Sacred-injected config values (`dist_scalar=20`, `protos=3`; `pemp_stage1.py:21-29`)
are keyword arguments here (injected by the reference's own Sacred ingredient once `dropin.patch()` re-applies its `capture`).
Under `torch.no_grad()` (`core/base_trainer.py:69`) the forward-only kernels run; when autograd is recording and an input
requires grad every function takes a differentiable path (`pemp_b200.autograd`), so a patched model trains correctly.
CUDA tensors only.
"""
import torch
import torch.nn as nn

from . import ops

DIST_SCALAR = 20      # `net.dist_scalar`, pemp_stage1.py:23 / baseline.py:22 / panet.py:21


# ------------------------------------------------------------------------------------------------------
# PEMP stage 1 / stage 2
# ------------------------------------------------------------------------------------------------------
def _wants_grad(*tensors):
    """True inside a training step: the reference calls the same methods under autograd (entry/pemp_stage1.py:57-65,
    entry/panet.py:108-115), so the drop-ins then take the differentiable path (`pemp_b200.autograd`: the same forward
    kernels plus hand-written backward kernels) instead of the forward-only one."""
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def _scalar(self, dist_scalar):
    """`dist_scalar` precedence: explicit argument (that is where Sacred's `net_ingredient.capture` injects `net.dist_scalar`
    once `dropin.patch()` has wrapped these functions with it), then a `dist_scalar` attribute of the module, then the
    reference's default of 20."""
    if dist_scalar is not None:
        return dist_scalar
    return getattr(self, "dist_scalar", DIST_SCALAR) if self is not None else DIST_SCALAR


def _upsample(pred, out_shape, differentiable):
    if differentiable:      # stock differentiable op, as in the reference; `autograd.upsample_ce` fuses it with the loss
        return torch.nn.functional.interpolate(pred, size=tuple(out_shape), mode="bilinear", align_corners=True)
    return ops.upsample_argmax(pred, out_shape, want_logits=True, want_mask8=False)["logits"]


def compute_similarity(self, fg_proto, bg_proto, qry_fts, dist_scalar=None):
    """`compute_similarity` of all four reference models (pemp_stage1.py:233-261, pemp_stage2.py:205-233,
    baseline.py:121-149, panet.py:122-156).

    fg_proto / bg_proto [B, c] with qry_fts [N, c, h, w]      -> [N, 2, h, w]
    fg_proto / bg_proto [B, c, p] with qry_fts [N, c, 1, h, w] -> [N, 2, p, h, w]
    Channel 0 is background, 1 foreground.  N may be a multiple of B (prototypes are expanded b-major,
    panet.py:145-149)."""
    if qry_fts.dim() == 5:                      # PEMP passes [N, c, 1, h, w] (pemp_stage1.py:196)
        N, c, _, h, w = qry_fts.shape
    else:
        N, c, h, w = qry_fts.shape
    if _wants_grad(fg_proto, bg_proto, qry_fts):        # training: same forward kernel + pemp_cosine_sim_bwd
        from . import autograd as A
        sim = A.cosine_sim(qry_fts.reshape(N, c, h * w), fg_proto, bg_proto, _scalar(self, dist_scalar))
    else:
        sim = ops.cosine_match(qry_fts.reshape(N, c, h * w), fg_proto, bg_proto, _scalar(self, dist_scalar), want_sim=True,
                               want_pred=False)["sim"]
    if fg_proto.dim() == 2:
        return sim.view(N, 2, h, w)
    return sim.view(N, 2, fg_proto.shape[2], h, w)


def mpm(self, sup_fts, qry_fts, sup_fg, sup_bg, ret_ind, protos=None, dist_scalar=None):
    """`PEMPStage1.mpm` / `PEMPStage2.mpm` (pemp_stage1.py:166-230, pemp_stage2.py:165-202).

    sup_fts [B, S, c, h, w]; qry_fts [B, Q, c, h, w]; sup_fg / sup_bg [BS, h, w]
    -> pred [BQ, 2, h, w]  or  (pred, response [BQ, h, w] int64) when `ret_ind` and the model has `ctr`.
    Stage 2's side effect `self.adaptive_p [B, c, 2p]` (pemp_stage2.py:185) is kept."""
    B, S, c, h, w = sup_fts.shape
    Q = qry_fts.shape[1]
    hw = h * w
    sup, qry = sup_fts, qry_fts                     # 5-D episode views are read in place (episode stride)
    fg, bg = sup_fg.reshape(B * S, hw), sup_bg.reshape(B * S, hw)
    scalar = _scalar(self, dist_scalar)
    ctr = getattr(self, "ctr", None)
    if ctr is not None and protos is not None and protos * 2 != ctr.shape[1]:
        raise ValueError(f"protos={protos} does not match ctr of shape {tuple(ctr.shape)}")
    if _wants_grad(sup_fts, qry_fts, ctr):
        from . import autograd as A
        if ctr is not None:
            fg_proto, bg_proto = A.meta_proto_attn(sup, ctr, fg, bg, eps=1e-6)
            if getattr(self, "_pemp_keep_adaptive", False):
                p = ctr.shape[1] // 2
                self.adaptive_p = torch.cat((fg_proto, bg_proto), dim=2).detach().view(B, c, 2 * p)
        else:
            fg_proto, bg_proto = A.map_pool_lowres(sup, fg, bg, eps=1e-5)
        pred = A.cosine_match(qry, fg_proto, bg_proto, scalar)
        if ret_ind and ctr is not None:
            with torch.no_grad():
                resp = ops.cosine_match(qry, fg_proto.detach(), bg_proto.detach(), scalar, want_pred=False, want_response=True)
            return pred, resp["response"].view(B * Q, h, w)
        return pred
    if ctr is not None:
        fg_proto, bg_proto, adaptive = ops.meta_proto_attn(sup, ctr.detach(), fg, bg, B, S, eps=1e-6)
        if getattr(self, "_pemp_keep_adaptive", False):
            self.adaptive_p = adaptive
        out = ops.cosine_match(qry, fg_proto, bg_proto, scalar, want_pred=True, want_response=bool(ret_ind))
        pred = out["pred"].view(B * Q, 2, h, w)
        if ret_ind:
            return pred, out["response"].view(B * Q, h, w)
        return pred
    fg_proto, bg_proto = ops.map_pool_lowres(sup, fg, bg, B, S, eps=1e-5)
    return ops.cosine_match(qry, fg_proto, bg_proto, scalar)["pred"].view(B * Q, 2, h, w)


def pemp_head(self, features, sup_mask, B, S, Q, out_shape=None, ret_ind=False, dist_scalar=None):
    """Everything `PEMPStage1.forward` / `PEMPStage2.forward` do after the encoder call
    (pemp_stage1.py:141-163, pemp_stage2.py:140-162): split, nearest mask down-sampling, `mpm`, bilinear
    up-sampling.  features [B(S+Q), c, h, w]; sup_mask [B, S, 2, H, W]."""
    _, c, h, w = features.shape
    H, W = sup_mask.shape[-2:]
    feats = features.view(B, S + Q, c, h, w)
    low = ops.mask_nearest(sup_mask.reshape(B * S, 2, H, W), h, w)            # [BS, 2, h, w]
    pred = mpm(self, feats[:, :S], feats[:, S:], low[:, 0], low[:, 1], ret_ind, dist_scalar=dist_scalar)
    if out_shape is None:
        out_shape = (H, W)
    if ret_ind and isinstance(pred, tuple):
        pred, response = pred
        return _upsample(pred, out_shape, pred.requires_grad), ops.nearest_resize_labels(response, out_shape)
    return _upsample(pred, out_shape, pred.requires_grad)


def pemp_stage1_forward(self, sup_img, sup_mask, qry_img, out_shape=None, ret_ind=False, dist_scalar=None):
    """Drop-in `PEMPStage1.forward` (pemp_stage1.py:112-163): stock encoder, B200 head."""
    B, S, channel, H, W = sup_img.size()
    Q = qry_img.size(1)
    img_cat = torch.cat((sup_img, qry_img), dim=1).view(B * (S + Q), channel, H, W)
    features = self.encoder(img_cat)
    return pemp_head(self, features, sup_mask, B, S, Q, out_shape, ret_ind, dist_scalar)


def pemp_stage2_forward(self, sup_img, sup_mask, qry_img, qry_prior, out_shape=None, ret_ind=False, dist_scalar=None):
    """Drop-in `PEMPStage2.forward` (pemp_stage2.py:104-162): 4-channel input assembly and the ResNetCM
    encoder stay on PyTorch, the head runs on the B200 kernels."""
    B, S, channel, H, W = sup_img.size()
    Q = qry_img.size(1)
    img_cat = torch.cat((sup_img, qry_img), dim=1).view(B * (S + Q), channel, H, W)
    sup_prior = sup_mask[:, :, :1]
    qry_prior = qry_prior.view(B, Q, *qry_prior.shape[-3:])
    prior_cat = torch.cat((sup_prior, qry_prior.float()), dim=1).view(B * (S + Q), 1, H, W)
    inputs = torch.cat((img_cat, prior_cat), dim=1)
    features = self.encoder((inputs, prior_cat))
    self._pemp_keep_adaptive = True
    return pemp_head(self, features, sup_mask, B, S, Q, out_shape, ret_ind, dist_scalar)


# ------------------------------------------------------------------------------------------------------
# Baseline / PANet
# ------------------------------------------------------------------------------------------------------
def baseline_head(self, features, sup_mask, B, S, Q, out_shape=None, with_align=False, dist_scalar=None):
    """`Baseline.forward` / `PANet.forward` after the encoder (baseline.py:97-118, panet.py:96-119)."""
    _, c, h, w = features.shape
    H, W = sup_mask.shape[-2:]
    feats = features.view(B, S + Q, c, h, w)
    sup_fts, qry_fts = feats[:, :S], feats[:, S:]   # episode views, read in place
    mask = sup_mask.reshape(B * S, 2, H, W)
    scalar = _scalar(self, dist_scalar)
    if _wants_grad(features):
        from . import autograd as A
        fg_proto, bg_proto = A.map_pool_fullres(sup_fts, mask, eps=1e-5)
        pred = A.cosine_match(qry_fts, fg_proto, bg_proto, scalar)
    else:
        fg_proto, bg_proto = ops.map_pool_fullres(sup_fts, mask, B, S, eps=1e-5)
        pred = ops.cosine_match(qry_fts, fg_proto, bg_proto, scalar)["pred"].view(B * Q, 2, h, w)
    if out_shape is None:
        out_shape = (H, W)
    output = _upsample(pred, out_shape, pred.requires_grad)
    if with_align:
        return output, alignLoss(self, qry_fts, pred, sup_fts, mask[:, 0:1], Q, scalar)
    return output


def baseline_forward(self, sup_img, sup_mask, qry_img, out_shape=None, dist_scalar=None):
    B, S, C, H, W = sup_img.size()
    Q = qry_img.size(1)
    features = self.encoder(torch.cat((sup_img, qry_img), dim=1).view(B * (S + Q), C, H, W))
    return baseline_head(self, features, sup_mask, B, S, Q, out_shape, with_align=False, dist_scalar=dist_scalar)


def panet_forward(self, sup_img, sup_mask, qry_img, out_shape=None, dist_scalar=None):
    B, S, C, H, W = sup_img.size()
    Q = qry_img.size(1)
    features = self.encoder(torch.cat((sup_img, qry_img), dim=1).view(B * (S + Q), C, H, W))
    return baseline_head(self, features, sup_mask, B, S, Q, out_shape, with_align=True, dist_scalar=dist_scalar)


def alignLoss(self, qry_fts, pred, sup_fts, sup_mask_fg, Q, dist_scalar=None):
    """`PANet.alignLoss` (panet.py:158-194) -> 0-dim tensor.  Unlike the reference (whose `.view` on an
    expanded tensor raises for B > 1 with S > 1) any B, S, Q combination works."""
    scalar = _scalar(self, dist_scalar)
    if _wants_grad(qry_fts, sup_fts):
        from . import autograd as A
        B = qry_fts.shape[0] if qry_fts.dim() == 5 else qry_fts.shape[0] // Q
        q5 = qry_fts if qry_fts.dim() == 5 else qry_fts.view(B, Q, *qry_fts.shape[1:])
        s5 = sup_fts if sup_fts.dim() == 5 else sup_fts.view(B, -1, *sup_fts.shape[1:])
        H, W = sup_mask_fg.shape[-2:]
        return A.panet_align_loss(q5, pred.detach(), s5, sup_mask_fg.reshape(-1, H, W), scalar)
    return ops.panet_align(qry_fts, pred, sup_fts, sup_mask_fg, Q, scalar)


# ------------------------------------------------------------------------------------------------------
# PFENet
# ------------------------------------------------------------------------------------------------------
def Weighted_GAP(supp_feat, mask):
    """`networks.pfenet.Weighted_GAP` (pfenet.py:15-20): [B,c,h,w], [B,1,h,w] -> [B,c,1,1].
    The reference trains through it (`down_supp` feeds it, pfenet.py:197-198): under autograd the differentiable K8
    (`autograd.weighted_gap`: K1's backward kernel) runs; a mask that itself requires grad - never the case in the
    reference, masks are labels - takes the stock expression so that no gradient is ever silently dropped."""
    if _wants_grad(supp_feat, mask):
        if mask.requires_grad:
            area = mask.sum(dim=(2, 3), keepdim=True) + 0.0005
            return (supp_feat * mask).sum(dim=(2, 3), keepdim=True) / area
        from . import autograd as A
        return A.weighted_gap(supp_feat, mask)
    return ops.weighted_gap(supp_feat, mask)


def prior_mask(query_feat_4, final_supp_list, mask_list, precision=None, out_hw=None):
    """The prior block of `PFENet.forward` (pfenet.py:201-231) as one call.

    query_feat_4 [B, C, sp, sp]; final_supp_list: S tensors [B, C, sp, sp]; mask_list: S binary masks
    [B, 1, H, W] -> corr_query_mask [B, 1, sp, sp] (bilinearly resized to `out_hw`, the feat-3 size of pfenet.py:224-225,
    when that differs - it does not for the reference's dilated ResNet).
    precision: `ops.PRIOR_BF16X3` (default; tcgen05 tensor cores, three bf16 products of a hi/lo split = fp32-grade cosines),
    `ops.PRIOR_BF16` (single product, 3x faster, cosines to 3e-3) or `ops.PRIOR_FP32` (CUDA-core anchor used by the tests).
    The block runs under `torch.no_grad()` in the reference (its operands come out of the frozen backbone), so it is
    forward-only here as well."""
    if precision is None:
        precision = ops.PRIOR_DEFAULT
    sp_h, sp_w = query_feat_4.shape[-2:]
    with torch.no_grad():
        s4 = torch.stack(list(final_supp_list), dim=0)
        masks = torch.stack(list(mask_list), dim=0)                                # [S, B, 1, H, W]
        small = ops.bilinear_resize(masks, s4.shape[-2:])[:, :, 0]                 # [S, B, sp, sp]
        prior = ops.prior_mask(query_feat_4, s4, small, precision).view(query_feat_4.shape[0], 1, sp_h, sp_w)
        if out_hw is not None and tuple(out_hw) != (sp_h, sp_w):
            prior = ops.bilinear_resize(prior, out_hw)
    return prior


# ------------------------------------------------------------------------------------------------------
# Stand-alone modules (same parameters / state-dict keys as the reference heads)
# ------------------------------------------------------------------------------------------------------
class PEMPHead(nn.Module):
    """Head of `PEMPStage1` / `PEMPStage2`: the only parameter is `ctr [c, 2p]` (state-dict key `ctr`,
    pemp_stage1.py:104-107), so reference checkpoints load."""

    def __init__(self, out_channels=512, protos=3, dist_scalar=DIST_SCALAR, keep_adaptive=False):
        super().__init__()
        self.dist_scalar = dist_scalar
        self.ctr = nn.Parameter(torch.rand(out_channels, protos * 2), requires_grad=True) if protos > 0 else None
        self._pemp_keep_adaptive = keep_adaptive

    mpm = mpm
    compute_similarity = compute_similarity

    def forward(self, features, sup_mask, B, S, Q, out_shape=None, ret_ind=False):
        return pemp_head(self, features, sup_mask, B, S, Q, out_shape, ret_ind)


class BaselineHead(nn.Module):
    def __init__(self, dist_scalar=DIST_SCALAR, align=False):
        super().__init__()
        self.dist_scalar = dist_scalar
        self.align = align

    compute_similarity = compute_similarity
    alignLoss = alignLoss

    def forward(self, features, sup_mask, B, S, Q, out_shape=None):
        return baseline_head(self, features, sup_mask, B, S, Q, out_shape, with_align=self.align)


# ----------------------------------------------------------------------------------------------
# "next" row: communication module of the Stage-2 backbones
# ----------------------------------------------------------------------------------------------

def comm(self, x, mask, linear, stride=2):
    """Drop-in for `ResNetCM.comm` / `VGG16CM.comm` (backbones.py:208-222, 469-479): same arguments (`linear` is the
    nn.Linear the backbone passes), same returns `(feat [N, n, h, w], pooled mask [N, 1, h, w])`.
    Stage-2 training back-propagates through `comm` into `x` and the linear layer: K11 is forward-only, so under autograd
    the same quantities are formed with stock differentiable ops (no gradient is silently dropped); evaluation
    (`torch.no_grad()`, core/base_trainer.py:69) runs the fused kernel."""
    if _wants_grad(x, mask, linear.weight, linear.bias):
        pooled = torch.nn.functional.max_pool2d(mask, 3, stride, 1)
        N, c, h, w = x.shape
        spq = self.spq
        masked = (x * pooled).flatten(2)
        stats = torch.cat((masked.mean(dim=2), masked.max(dim=2)[0]), dim=1).view(N // spq, spq, 2 * c).mean(dim=1)
        feat = linear(stats)
        return feat[:, None, :, None, None].expand(-1, spq, -1, h, w).reshape(N, -1, h, w), pooled
    return ops.comm_module(x, mask, linear.weight, linear.bias, self.spq, stride)
