"""Training path of the PEMP head: `torch.autograd.Function`s over the K2 / K3 kernels and their
hand-written backward kernels (`csrc/train.cu`), so that `loss.backward()` of `entry/pemp_stage1.py:57-65` runs on them.

    fg, bg = meta_proto_attn(sup_fts, ctr, sup_fg, sup_bg)          # pemp_stage1.py:202-213, grads to sup_fts and ctr
    pred   = cosine_match(qry_fts, fg, bg, dist_scalar)             # pemp_stage1.py:214-215,233-261, grads to all three
    loss   = F.cross_entropy(F.interpolate(pred, size, mode="bilinear", align_corners=True), target, ignore_index=255)

Masks get no gradient (they are labels).  `upsample_ce` is the up-sampling + cross entropy of the last line as one op
(K13: the full-size logits are never stored; the gradient comes out of the forward pass).
There is no CPU path: CUDA float32 tensors only.
"""
import torch

from . import ops


class _MetaProtoAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sup_fts, ctr, fg, bg, eps):
        B, S = sup_fts.shape[:2]
        fgp, bgp, saved = ops.meta_proto_attn_train(sup_fts, ctr, fg, bg, B, S, eps)
        fts, ep, ctr_c, fg_c, bg_c, centre, den = saved
        ctx.save_for_backward(fts, ctr_c, fg_c, bg_c, centre, den)     # version-checked, released after backward
        ctx.ep = ep
        ctx.shape = tuple(sup_fts.shape)
        return fgp, bgp

    @staticmethod
    def backward(ctx, g_fg, g_bg):
        B, S = ctx.shape[:2]
        fts, ctr, fg, bg, centre, den = ctx.saved_tensors
        zero = lambda g, like: torch.zeros(like, dtype=torch.float32, device=fts.device) if g is None else g.contiguous()
        p = ctr.shape[1] // 2
        like = (B, ctx.shape[2], p)
        d_fts, d_ctr = ops.meta_proto_attn_bwd((fts, ctx.ep, ctr, fg, bg, centre, den), zero(g_fg, like), zero(g_bg, like), B, S)
        return d_fts.view(ctx.shape), d_ctr, None, None, None


class _CosineMatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qry_fts, fg_proto, bg_proto, scalar):
        out = ops.cosine_match(qry_fts, fg_proto, bg_proto, scalar)["pred"]
        ctx.save_for_backward(qry_fts, fg_proto, bg_proto)
        ctx.scalar = scalar
        return out

    @staticmethod
    def backward(ctx, g_pred):
        qry, fgp, bgp = ctx.saved_tensors
        d_qry, d_fg, d_bg = ops.cosine_match_bwd(qry, fgp, bgp, g_pred.contiguous(), ctx.scalar)
        return d_qry.view(qry.shape), d_fg, d_bg, None


class _CosineSim(torch.autograd.Function):
    """The per-prototype similarity maps of `compute_similarity` (pemp_stage1.py:233-261) with their backward kernel
    (`pemp_cosine_sim_bwd`: the gradient tile is a rank-2P combination of the normalised prototypes)."""

    @staticmethod
    def forward(ctx, qry_fts, fg_proto, bg_proto, scalar):
        out = ops.cosine_match(qry_fts, fg_proto, bg_proto, scalar, want_sim=True, want_pred=False)["sim"]
        ctx.save_for_backward(qry_fts, fg_proto, bg_proto)
        ctx.scalar = scalar
        return out

    @staticmethod
    def backward(ctx, g_sim):
        qry, fgp, bgp = ctx.saved_tensors
        d_qry, d_fg, d_bg = ops.cosine_match_bwd(qry, fgp, bgp, g_sim.contiguous(), ctx.scalar, dense=True)
        return d_qry.view(qry.shape), d_fg, d_bg, None


def cosine_sim(qry_fts, fg_proto, bg_proto, scalar=20.0):
    """qry_fts [N, c, hw]; prototypes [Bp, c] or [Bp, c, P] -> sim [N, 2, P, hw] (channel 0 background), differentiable in
    all three."""
    return _CosineSim.apply(qry_fts, fg_proto, bg_proto, scalar)


class _MapPoolLowres(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sup_fts, fg, bg, eps):
        B, S, c = sup_fts.shape[:3]
        fgp, bgp = ops.map_pool_lowres(sup_fts, fg, bg, B, S, eps)
        ctx.save_for_backward(fg, bg)
        ctx.shape, ctx.eps = tuple(sup_fts.shape), eps
        return fgp, bgp

    @staticmethod
    def backward(ctx, g_fg, g_bg):
        fg, bg = ctx.saved_tensors
        B, S, c = ctx.shape[:3]
        z = lambda g: torch.zeros(B, c, dtype=torch.float32, device=fg.device) if g is None else g.contiguous()
        d = ops.map_pool_lowres_bwd(fg, bg, z(g_fg), z(g_bg), B, S, c, ctx.eps)
        return d.view(ctx.shape), None, None, None


class _WeightedGAP(torch.autograd.Function):
    """PFENet's `Weighted_GAP` (pfenet.py:15-20) for the training path: forward K8, backward = the K1 backward kernel with
    one mask per image (d f = g (x) m / (sum m + 5e-4)); the mask is a label and gets no gradient."""

    @staticmethod
    def forward(ctx, supp_feat, mask):
        out = ops.weighted_gap(supp_feat, mask)
        ctx.save_for_backward(mask)
        ctx.shape = tuple(supp_feat.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        B, c, h, w = ctx.shape
        d = ops.map_pool_lowres_bwd(mask.reshape(B, h * w), None, g.reshape(B, c).contiguous(), None, B, 1, c, 5e-4)
        return d.view(ctx.shape), None


def weighted_gap(supp_feat, mask):
    """`Weighted_GAP(supp_feat [B,c,h,w], mask [B,1,h,w]) -> [B,c,1,1]`, differentiable in supp_feat."""
    return _WeightedGAP.apply(supp_feat, mask)


def map_pool_lowres(sup_fts, sup_fg, sup_bg, eps=1e-5):
    """Masked average pooling at feature resolution (`pemp_stage1.py:223-227`, the baseline head's prototypes):
    sup_fts [B, S, c, h, w], masks [B*S, h*w] -> fg_proto, bg_proto [B, c]; differentiable in sup_fts."""
    if sup_fts.dim() != 5:
        raise ValueError("sup_fts must be [B, S, c, h, w]")
    return _MapPoolLowres.apply(sup_fts, sup_fg, sup_bg, eps)


class _MapPoolFullres(torch.autograd.Function):
    """K6 forward; backward = the K1 backward kernel on the adjoint weight maps U^T mask (pooling the up-sampled features
    with the mask equals pooling the features with U^T mask, and sum(U^T mask) = sum(mask))."""

    @staticmethod
    def forward(ctx, sup_fts, sup_mask, eps):
        B, S = sup_fts.shape[:2]
        fgp, bgp = ops.map_pool_fullres(sup_fts, sup_mask, B, S, eps)
        ctx.save_for_backward(sup_mask)
        ctx.shape, ctx.eps = tuple(sup_fts.shape), eps
        return fgp, bgp

    @staticmethod
    def backward(ctx, g_fg, g_bg):
        (sup_mask,) = ctx.saved_tensors
        B, S, c, h, w = ctx.shape
        wt, _ = ops.bilinear_adjoint(sup_mask, (h, w), want_sum=False)                 # [BS, 2, h, w]
        wt = wt.view(B * S, 2, h * w)
        z = lambda g: torch.zeros(B, c, dtype=torch.float32, device=wt.device) if g is None else g.contiguous()
        d = ops.map_pool_lowres_bwd(wt[:, 0], wt[:, 1], z(g_fg), z(g_bg), B, S, c, ctx.eps)
        return d.view(ctx.shape), None, None


def map_pool_fullres(sup_fts, sup_mask, eps=1e-5):
    """Masked average pooling at mask resolution (`baseline.py:100-110`, `panet.py:99-109`): sup_fts [B, S, c, h, w],
    sup_mask [B*S, 2, H, W] -> fg_proto, bg_proto [B, c]; differentiable in sup_fts."""
    if sup_fts.dim() != 5:
        raise ValueError("sup_fts must be [B, S, c, h, w]")
    return _MapPoolFullres.apply(sup_fts, sup_mask, eps)


def panet_align_loss(qry_fts, pred, sup_fts, sup_mask_fg, scalar=20.0):
    """PANet's prototype-alignment loss (`panet.py:158-194`) from the differentiable pieces: query prototypes pooled with
    the arg-max masks of `pred` (no gradient through the arg-max), every support map matched against them, cross entropy
    against the support foreground mask.  qry_fts [B, Q, c, h, w], pred [B*Q, 2, h, w], sup_fts [B, S, c, h, w],
    sup_mask_fg [B*S, H, W] -> 0-dim loss; differentiable in qry_fts and sup_fts."""
    B, Q = qry_fts.shape[:2]
    with torch.no_grad():
        fgm = (pred[:, 1] > pred[:, 0]).float().reshape(B * Q, -1)
        bgm = 1.0 - fgm
        target = sup_mask_fg.long()
    fgp, bgp = map_pool_lowres(qry_fts, fgm, bgm, 1e-5)
    rev = cosine_match(sup_fts, fgp, bgp, scalar)
    return upsample_ce(rev, target)


class _UpsampleCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, weight):
        loss, d_pred = ops.upsample_ce(pred, target, want_grad=True, weight=weight)
        ctx.save_for_backward(d_pred)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (d_pred,) = ctx.saved_tensors
        return d_pred * g, None, None


class _PempHead(torch.autograd.Function):
    """K2 -> K3 on the encoder output [B, S+Q, c, h, w] as ONE autograd node: both backward kernels write straight into
    the two halves of one gradient tensor (separate nodes would make autograd zero-fill, copy and add two full-size
    tensors for the `[:, :S]` / `[:, S:]` slices - three extra passes over the largest tensor of the step)."""

    @staticmethod
    def forward(ctx, f5, ctr, fg, bg, S, scalar, eps):
        B = f5.shape[0]
        sup, qry = f5[:, :S], f5[:, S:]
        fgp, bgp, saved = ops.meta_proto_attn_train(sup, ctr, fg, bg, B, S, eps)
        pred = ops.cosine_match(qry, fgp, bgp, scalar)["pred"]
        fts, ep, ctr_c, fg_c, bg_c, centre, den = saved
        # `fts` is f5[:, :S] itself (read in place) or a dense copy of it; everything goes through save_for_backward so that
        # an in-place write to the encoder output or to ctr between forward and backward raises instead of corrupting grads
        ctx.save_for_backward(f5, fts, ctr_c, fg_c, bg_c, centre, den, fgp, bgp)
        ctx.ep, ctx.S, ctx.scalar = ep, S, scalar
        return pred

    @staticmethod
    def backward(ctx, g_pred):
        f5, fts, ctr, fg, bg, centre, den, fgp, bgp = ctx.saved_tensors
        S = ctx.S
        B = f5.shape[0]
        d_f5 = torch.empty(f5.shape, dtype=torch.float32, device=f5.device)
        _, d_fg, d_bg = ops.cosine_match_bwd(f5[:, S:], fgp, bgp, g_pred.contiguous(), ctx.scalar, out=d_f5[:, S:])
        _, d_ctr = ops.meta_proto_attn_bwd((fts, ctx.ep, ctr, fg, bg, centre, den), d_fg, d_bg, B, S, out=d_f5[:, :S])
        return d_f5, d_ctr, None, None, None, None, None


def pemp_head(features, sup_mask_low, ctr, B, S, Q, scalar=20.0, eps=1e-6):
    """features [B*(S+Q), c, h, w] (encoder output), sup_mask_low [B*S, 2, h*w], ctr [c, 2p] -> pred [B*Q, 2, h, w];
    differentiable in features and ctr (`PEMPStage1.forward` in training mode, pemp_stage1.py:140-162 without the
    up-sampling)."""
    _, c, h, w = features.shape
    f5 = features.view(B, S + Q, c, h, w)
    return _PempHead.apply(f5, ctr, sup_mask_low[:, 0], sup_mask_low[:, 1], S, scalar, eps).view(B * Q, 2, h, w)


def upsample_ce(pred, target, weight=None):
    """loss = CrossEntropyLoss(ignore_index=255)(F.interpolate(pred, target.shape[-2:], bilinear, align_corners=True),
    target) in one pass, gradient included (K13).  With `weight` [N, H, W]: sum(w * ce) / sum(w)."""
    return _UpsampleCE.apply(pred, target, weight)


def ce_loss_dt(pred, target, sigma=5.0):
    """`CELossDT(sigma)` (core/losses.py:17-43, `loss=cedt`) on the up-sampled prediction: the boundary distance-transform
    weights are computed on the device (K14; the reference goes through scipy on the host every step)."""
    return upsample_ce(pred, target, ops.boundary_weight(target, sigma))


def meta_proto_attn(sup_fts, ctr, sup_fg, sup_bg, eps=1e-6):
    """sup_fts [B, S, c, h, w] (may be a slice of the encoder output, read in place), ctr [c, 2p] (the module's
    `self.ctr` viewed as [c, 2p]), sup_fg / sup_bg [B*S, h*w] -> fg_proto, bg_proto [B, c, p]; differentiable in
    sup_fts and ctr."""
    if sup_fts.dim() != 5:
        raise ValueError("sup_fts must be [B, S, c, h, w]")
    return _MetaProtoAttn.apply(sup_fts, ctr, sup_fg, sup_bg, eps)


def cosine_match(qry_fts, fg_proto, bg_proto, scalar=20.0):
    """qry_fts [B, Q, c, h, w]; prototypes [B, c, p] (or [B, c]) -> pred [B*Q, 2, h, w]; differentiable in all three."""
    if qry_fts.dim() != 5:
        raise ValueError("qry_fts must be [B, Q, c, h, w]")
    B, Q, _, h, w = qry_fts.shape
    return _CosineMatch.apply(qry_fts, fg_proto, bg_proto, scalar).view(B * Q, 2, h, w)


def pemp_head_loss(features, sup_mask_low, ctr, B, S, Q, target, out_shape=None, scalar=20.0):
    """One training step of the head as `entry/pemp_stage1.py:57-65` runs it: features [B*(S+Q), c, h, w] from the encoder,
    sup_mask_low [B*S, 2, h*w] (K0 output), target [B*Q, H, W] int64 with 255 = ignore -> (loss, pred)."""
    pred = pemp_head(features, sup_mask_low, ctr, B, S, Q, scalar)
    if out_shape is not None and tuple(out_shape) != tuple(target.shape[-2:]):
        raise ValueError("the loss is taken at the size of the target")
    return upsample_ce(pred, target), pred
