"""Torch-tensor front end of the C ABI: argument validation, workspace and stream plumbing.

Each function takes CUDA tensors, passes raw device pointers + the current stream to
`libpemp_b200.so`, and returns freshly allocated CUDA tensors.  Nothing here computes on the host and
nothing falls back to PyTorch ops.
"""
import functools

import torch

from . import _cabi

_launches = 0      # kernels launched through this module (bench.py's `gpu_launches` claim)


def launch_count():
    return _launches


def _count(n):
    global _launches
    _launches += n


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _tensors(args):
    for a in args:
        if isinstance(a, torch.Tensor):
            yield a
        elif isinstance(a, (tuple, list)):
            yield from _tensors(a)


def _on_device(fn):
    """Every operand of a call must live on ONE CUDA device, and the call runs with that device current: raw pointers of
    one GPU are never launched on another GPU's stream (e.g. `FewShotMetric(device='cuda:1')` while cuda:0 is current).
    `_stream()` inside the call is then the caller's current stream of that device."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for t in _tensors(args + tuple(kwargs.values())):
            if not t.is_cuda:
                continue                      # reported by the per-argument checks with the argument's name
            if dev is None:
                dev = t.device
            elif t.device != dev:
                raise ValueError(f"{fn.__name__}: operands live on different devices ({dev} and {t.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _need(t, dtype, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device (pemp_b200 has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _episodes(t, B, S, name):
    """Feature operand of B*S images.  Accepts image-major [B*S, c, hw] / [B*S, c, h, w] or the 5-D episode
    view [B, S, c, h, w] - typically a slice `features.view(B, S+Q, c, h, w)[:, :S]` of the encoder output,
    which is read in place through the episode stride (no copy).  Returns (tensor, episode_stride, c, hw)."""
    t = _need_loose(t, name)
    if t.dim() == 5:
        if t.shape[0] != B or t.shape[1] != S:
            raise ValueError(f"{name} must be [B={B}, S={S}, c, h, w], got {tuple(t.shape)}")
        c, h, w = t.shape[2:]
        hw = h * w
        inner_ok = t.stride(4) == 1 and t.stride(3) == w and t.stride(2) == hw and (S == 1 or t.stride(1) == c * hw)
        if not inner_ok or (B > 1 and t.stride(0) < S * c * hw):
            t = t.contiguous()
        return t, (t.stride(0) if B > 1 else S * c * hw), c, hw
    if t.dim() not in (3, 4) or t.shape[0] != B * S:
        raise ValueError(f"{name} must be [B*S={B * S}, c, hw], [B*S, c, h, w] or [B, S, c, h, w], got {tuple(t.shape)}")
    t = t if t.is_contiguous() else t.contiguous()
    c = t.shape[1]
    return t, S * c * (t.numel() // (B * S * c)), c, t.numel() // (B * S * c)


def _mask_pair(fg, bg, n_img, hw):
    """fg / bg may be two views of one [n_img, 2, hw] tensor (the K0 output) or separate [n_img, hw] tensors.
    Returns (fg_ptr, bg_ptr, stride_in_floats, keepalive)."""
    if fg.dim() != 2 or fg.shape != (n_img, hw):
        raise ValueError(f"mask must be [{n_img}, {hw}], got {tuple(fg.shape)}")
    if bg is not None and bg.shape != fg.shape:
        raise ValueError("fg and bg masks differ in shape")
    ok = fg.stride(1) == 1 and (bg is None or (bg.stride(1) == 1 and bg.stride(0) == fg.stride(0)))
    if n_img > 1 and fg.stride(0) < hw:
        ok = False
    if not ok:
        fg = fg.contiguous()
        bg = None if bg is None else bg.contiguous()
    stride = fg.stride(0) if n_img > 1 else hw
    return fg.data_ptr(), (0 if bg is None else bg.data_ptr()), stride, (fg, bg)


# ------------------------------------------------------------------------------------------------ K0
@_on_device
def mask_nearest(mask, h, w):
    """[..., H, W] float32 -> [..., h, w] (`F.interpolate(mode='nearest')`, pemp_stage1.py:147)."""
    mask = _need(mask, torch.float32, "mask")
    H, W = mask.shape[-2:]
    planes = mask.numel() // (H * W)
    out = torch.empty(*mask.shape[:-2], h, w, dtype=torch.float32, device=mask.device)
    _cabi.check(_cabi.lib().pemp_mask_nearest(mask.data_ptr(), planes, H, W, h, w, out.data_ptr(), _stream()),
                "pemp_mask_nearest")
    _count(1)
    return out


@_on_device
def mask_nearest_labels(labels, h, w):
    """labels [..., H, W] uint8 (1 = object, 0 = background, 255 = boundary; the map `pascal_voc.py:209-210` expands into the
    float `sup_mask`) -> [..., 2, h, w] float32 (fg, bg): the same low-res masks as `mask_nearest` on the expansion."""
    labels = _need(labels, torch.uint8, "labels")
    H, W = labels.shape[-2:]
    planes = labels.numel() // (H * W)
    out = torch.empty(*labels.shape[:-2], 2, h, w, dtype=torch.float32, device=labels.device)
    _cabi.check(_cabi.lib().pemp_mask_nearest_labels(labels.data_ptr(), planes, H, W, h, w, out.data_ptr(), _stream()),
                "pemp_mask_nearest_labels")
    _count(1)
    return out


# ------------------------------------------------------------------------------------------------ K1 / K8
@_on_device
def map_pool_lowres(fts, fg, bg, B, S, eps=1e-5):
    """fts [B*S, c, hw]; fg, bg [B*S, hw] -> (fg_proto [B, c], bg_proto [B, c])  (pemp_stage1.py:223-227)."""
    fts, ep, c, hw = _episodes(fts, B, S, "fts")
    n_img = B * S
    fg = _need_loose(fg, "fg").reshape(n_img, hw) if fg.dim() != 2 else _need_loose(fg, "fg")
    bg = None if bg is None else (_need_loose(bg, "bg").reshape(n_img, hw) if bg.dim() != 2 else _need_loose(bg, "bg"))
    fgp, bgp, stride, keep = _mask_pair(fg, bg, n_img, hw)
    L = _cabi.lib()
    ws = _ws(L.pemp_map_pool_workspace_bytes(B, S, c, hw), fts.device)
    out_f = torch.empty(B, c, dtype=torch.float32, device=fts.device)
    out_b = torch.empty(B, c, dtype=torch.float32, device=fts.device) if bg is not None else None
    _cabi.check(L.pemp_map_pool_lowres(fts.data_ptr(), ep, fgp, bgp, stride, B, S, c, hw, float(eps), out_f.data_ptr(),
                                       _ptr(out_b), ws.data_ptr(), ws.numel(), _stream()), "pemp_map_pool_lowres")
    _count(2)
    del keep
    return out_f, out_b


def _need_loose(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device (pemp_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {t.dtype}")
    return t


@_on_device
def weighted_gap(supp_feat, mask):
    """PFENet `Weighted_GAP(supp_feat [B,c,h,w], mask [B,1,h,w]) -> [B,c,1,1]` (pfenet.py:15-20)."""
    supp_feat = _need(supp_feat, torch.float32, "supp_feat")
    mask = _need(mask, torch.float32, "mask")
    B, c, h, w = supp_feat.shape
    if mask.numel() != B * h * w:
        raise ValueError(f"mask must be [B,1,h,w] = [{B},1,{h},{w}], got {tuple(mask.shape)}")
    L = _cabi.lib()
    ws = _ws(L.pemp_map_pool_workspace_bytes(B, 1, c, h * w), supp_feat.device)
    out = torch.empty(B, c, 1, 1, dtype=torch.float32, device=supp_feat.device)
    _cabi.check(L.pemp_weighted_gap(supp_feat.data_ptr(), mask.data_ptr(), B, c, h * w, out.data_ptr(), ws.data_ptr(),
                                    ws.numel(), _stream()), "pemp_weighted_gap")
    _count(2)
    return out


# ------------------------------------------------------------------------------------------------ K11
@_on_device
def comm_module(x, mask, weight, bias, spq, stride=2):
    """`ResNetCM.comm` / `VGG16CM.comm` (backbones.py:208-222, 469-479): x [N,c,h,w], mask [N,1,Hm,Wm], nn.Linear
    weight [n,2c] / bias [n] -> (feat broadcast [N,n,h,w], pooled mask [N,1,h,w])."""
    x = _need(x, torch.float32, "x")
    mask = _need(mask, torch.float32, "mask")
    weight = _need(weight, torch.float32, "weight")
    bias = None if bias is None else _need(bias, torch.float32, "bias")
    N, c, h, w = x.shape
    if mask.dim() != 4 or mask.shape[0] != N or mask.shape[1] != 1:
        raise ValueError(f"mask must be [N={N}, 1, Hm, Wm], got {tuple(mask.shape)}")
    Hm, Wm = mask.shape[-2:]
    n_out = weight.shape[0]
    if weight.dim() != 2 or weight.shape[1] != 2 * c:
        raise ValueError(f"weight must be [n, 2c = {2 * c}], got {tuple(weight.shape)}")
    if N % spq:
        raise ValueError(f"N = {N} is not a multiple of spq = {spq}")
    L = _cabi.lib()
    ws = _ws(L.pemp_comm_workspace_bytes(N, c, spq, n_out), x.device)
    mask_out = torch.empty(N, 1, h, w, dtype=torch.float32, device=x.device)
    out = torch.empty(N, n_out, h, w, dtype=torch.float32, device=x.device)
    _cabi.check(L.pemp_comm_module(x.data_ptr(), mask.data_ptr(), N, c, h, w, Hm, Wm, int(stride), int(spq), weight.data_ptr(),
                                   _ptr(bias), n_out, mask_out.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                "pemp_comm_module")
    _count(4)
    return out, mask_out


# ------------------------------------------------------------------------------------------------ K2
@_on_device
def meta_proto_attn(fts, ctr, fg, bg, B, S, eps=1e-6, want_adaptive=True):
    """fts [B*S, c, hw]; ctr [c, 2p]; fg, bg [B*S, hw] -> fg_proto [B,c,p], bg_proto [B,c,p], adaptive_p [B,c,2p]
    (pemp_stage1.py:202-213, pemp_stage2.py:174-186)."""
    fts, ep, c, hw = _episodes(fts, B, S, "fts")
    ctr = _need(ctr, torch.float32, "ctr")
    n_img = B * S
    if ctr.dim() != 2 or ctr.shape[0] != c or ctr.shape[1] % 2:
        raise ValueError(f"ctr must be [c, 2p] with c = {c}, got {tuple(ctr.shape)}")
    p = ctr.shape[1] // 2
    fg = _need_loose(fg, "fg")
    bg = _need_loose(bg, "bg")
    fg = fg if fg.dim() == 2 else fg.reshape(n_img, hw)
    bg = bg if bg.dim() == 2 else bg.reshape(n_img, hw)
    fgp, bgp, stride, keep = _mask_pair(fg, bg, n_img, hw)
    L = _cabi.lib()
    ws = _ws(L.pemp_meta_proto_attn_workspace_bytes(B, S, c, hw, p), fts.device)
    out_f = torch.empty(B, c, p, dtype=torch.float32, device=fts.device)
    out_b = torch.empty(B, c, p, dtype=torch.float32, device=fts.device)
    adaptive = torch.empty(B, c, 2 * p, dtype=torch.float32, device=fts.device) if want_adaptive else None
    _cabi.check(L.pemp_meta_proto_attn(fts.data_ptr(), ep, ctr.data_ptr(), fgp, bgp, stride, B, S, c, hw, p, float(eps),
                                       out_f.data_ptr(), out_b.data_ptr(), _ptr(adaptive), ws.data_ptr(), ws.numel(),
                                       _stream()), "pemp_meta_proto_attn")
    # launches: TMA path (c = 512, p = 3, hw >= 32) = main + finalize; generic = prepare (p > 1) + main + finalize
    _count(2 if (c == 512 and p == 3 and hw >= 32) or p == 1 else 3)
    del keep
    return out_f, out_b, adaptive


# ------------------------------------------------------------------------------------------------ K3
@_on_device
def cosine_match(qry, fg_proto, bg_proto, scalar=20.0, want_sim=False, want_pred=True, want_response=False):
    """qry [N, c, hw] / [N, c, h, w] or the episode view [Bp, Q, c, h, w]; protos [Bp, c] or [Bp, c, P]
    -> dict(sim [N,2,P,hw], pred [N,2,hw], response [N,hw] int64)
    (compute_similarity + max over prototypes, pemp_stage1.py:214-222,233-261)."""
    fg_proto = _need(fg_proto, torch.float32, "fg_proto")
    bg_proto = _need(bg_proto, torch.float32, "bg_proto")
    if fg_proto.shape != bg_proto.shape or fg_proto.dim() not in (2, 3):
        raise ValueError(f"prototypes must be [Bp, c(, P)], got {tuple(fg_proto.shape)} / {tuple(bg_proto.shape)}")
    Bp = fg_proto.shape[0]
    P = 1 if fg_proto.dim() == 2 else fg_proto.shape[2]
    n_maps = qry.shape[0] * qry.shape[1] if qry.dim() == 5 else qry.shape[0]
    if n_maps % Bp:
        raise ValueError(f"{n_maps} query maps cannot be split over {Bp} prototype sets")
    qry, ep, c, hw = _episodes(qry, Bp, n_maps // Bp, "qry")
    N = n_maps
    if fg_proto.shape[1] != c:
        raise ValueError(f"prototypes have {fg_proto.shape[1]} channels, query features {c}")
    dev = qry.device
    sim = torch.empty(N, 2, P, hw, dtype=torch.float32, device=dev) if want_sim else None
    pred = torch.empty(N, 2, hw, dtype=torch.float32, device=dev) if want_pred else None
    resp = torch.empty(N, hw, dtype=torch.int64, device=dev) if want_response else None
    _cabi.check(_cabi.lib().pemp_cosine_match(qry.data_ptr(), ep, fg_proto.data_ptr(), bg_proto.data_ptr(), N, Bp, c, hw, P,
                                              float(scalar), _ptr(sim), _ptr(pred), _ptr(resp), _stream()),
                "pemp_cosine_match")
    _count(1)
    return {"sim": sim, "pred": pred, "response": resp}


# ------------------------------------------------------------------------------------------------ K4 / K5
@_on_device
def upsample_argmax(pred, out_hw, want_logits=False, want_mask8=True, want_mask64=False):
    """pred [N, 2, h, w] -> dict(logits [N,2,H,W], mask8 [N,H,W] uint8, mask64 [N,H,W] int64)
    (F.interpolate bilinear align_corners + argmax, pemp_stage1.py:157-162, entry/pemp_stage1.py:52)."""
    pred = _need(pred, torch.float32, "pred")
    N, two, h, w = pred.shape
    if two != 2:
        raise ValueError("pred must be [N, 2, h, w]")
    H, W = int(out_hw[0]), int(out_hw[1])
    dev = pred.device
    logits = torch.empty(N, 2, H, W, dtype=torch.float32, device=dev) if want_logits else None
    m8 = torch.empty(N, H, W, dtype=torch.uint8, device=dev) if want_mask8 else None
    m64 = torch.empty(N, H, W, dtype=torch.int64, device=dev) if want_mask64 else None
    _cabi.check(_cabi.lib().pemp_upsample_argmax(pred.data_ptr(), N, h, w, H, W, _ptr(logits), _ptr(m8), _ptr(m64),
                                                 _stream()), "pemp_upsample_argmax")
    _count(1)
    return {"logits": logits, "mask8": m8, "mask64": m64}


@_on_device
def bilinear_resize(x, out_hw):
    """[..., h, w] float32 -> [..., H, W], bilinear align_corners=True (pfenet.py:191,205)."""
    x = _need(x, torch.float32, "x")
    h, w = x.shape[-2:]
    H, W = int(out_hw[0]), int(out_hw[1])
    out = torch.empty(*x.shape[:-2], H, W, dtype=torch.float32, device=x.device)
    _cabi.check(_cabi.lib().pemp_bilinear_resize(x.data_ptr(), x.numel() // (h * w), h, w, H, W, out.data_ptr(),
                                                 _stream()), "pemp_bilinear_resize")
    _count(1)
    return out


@_on_device
def nearest_resize_labels(lab, out_hw):
    """[..., h, w] int64 -> [..., H, W] nearest (response map, pemp_stage1.py:158-159)."""
    lab = _need(lab, torch.int64, "labels")
    h, w = lab.shape[-2:]
    H, W = int(out_hw[0]), int(out_hw[1])
    out = torch.empty(*lab.shape[:-2], H, W, dtype=torch.int64, device=lab.device)
    _cabi.check(_cabi.lib().pemp_nearest_resize_i64(lab.data_ptr(), lab.numel() // (h * w), h, w, H, W, out.data_ptr(),
                                                    _stream()), "pemp_nearest_resize_i64")
    _count(1)
    return out


# ------------------------------------------------------------------------------------------------ K6
@_on_device
def map_pool_fullres(fts, sup_mask, B, S, eps=1e-5):
    """fts [B*S, c, h, w]; sup_mask [B*S, 2, H, W] float32 (fg, bg) - or the uint8 label map [B*S, H, W] (1 object / 0 background /
    255 boundary) the loader expands it from - -> (fg_proto [B,c], bg_proto [B,c])  (baseline.py:100-110)."""
    if fts.dim() not in (4, 5):
        raise ValueError("fts must be [B*S, c, h, w] or [B, S, c, h, w]")
    h, w = fts.shape[-2:]
    fts, ep, c, _ = _episodes(fts, B, S, "fts")
    n_img = B * S
    labels = isinstance(sup_mask, torch.Tensor) and sup_mask.dtype == torch.uint8
    if labels:
        sup_mask = _need(sup_mask, torch.uint8, "sup_mask")
        if sup_mask.dim() < 3 or sup_mask.numel() // (sup_mask.shape[-2] * sup_mask.shape[-1]) != n_img:
            raise ValueError(f"expected labels [{n_img},H,W], got {tuple(sup_mask.shape)}")
    else:
        sup_mask = _need(sup_mask, torch.float32, "sup_mask")
        if sup_mask.shape[:2] != (n_img, 2):
            raise ValueError(f"expected sup_mask [{B * S},2,H,W], got {tuple(sup_mask.shape)}")
    H, W = sup_mask.shape[-2:]
    L = _cabi.lib()
    out_f = torch.empty(B, c, dtype=torch.float32, device=fts.device)
    out_b = torch.empty(B, c, dtype=torch.float32, device=fts.device)
    if labels:
        ws = _ws(L.pemp_map_pool_fullres_labels_workspace_bytes(B, S, c, h, w, H, W), fts.device)
        _cabi.check(L.pemp_map_pool_fullres_labels(fts.data_ptr(), ep, sup_mask.data_ptr(), B, S, c, h, w, H, W, float(eps),
                                                   out_f.data_ptr(), out_b.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                    "pemp_map_pool_fullres_labels")
    else:
        ws = _ws(L.pemp_map_pool_fullres_workspace_bytes(B, S, c, h, w), fts.device)
        _cabi.check(L.pemp_map_pool_fullres(fts.data_ptr(), ep, sup_mask.data_ptr(), B, S, c, h, w, H, W, float(eps),
                                            out_f.data_ptr(), out_b.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                    "pemp_map_pool_fullres")
    _count(5)
    return out_f, out_b


@_on_device
def bilinear_adjoint(mask, out_hw, want_sum=True):
    """mask [..., H, W] -> (U^T mask [..., h, w], plane sums [...])."""
    mask = _need(mask, torch.float32, "mask")
    H, W = mask.shape[-2:]
    h, w = int(out_hw[0]), int(out_hw[1])
    planes = mask.numel() // (H * W)
    wt = torch.empty(*mask.shape[:-2], h, w, dtype=torch.float32, device=mask.device)
    ms = torch.empty(mask.shape[:-2], dtype=torch.float32, device=mask.device) if want_sum else None
    _cabi.check(_cabi.lib().pemp_bilinear_adjoint(mask.data_ptr(), planes, H, W, h, w, wt.data_ptr(), _ptr(ms), _stream()),
                "pemp_bilinear_adjoint")
    _count(2 if want_sum else 1)
    return wt, ms


# ------------------------------------------------------------------------------------------------ K7
@_on_device
def panet_align(qry_fts, pred, sup_fts, sup_mask_fg, Q, scalar=20.0):
    """`PANet.alignLoss(qry_fts [BQ,c,h,w], pred [BQ,2,h,w], sup_fts [BS,c,h,w], sup_mask_fg [BS,1,H,W], Q)`
    -> 0-dim loss tensor (panet.py:158-194).  `sup_mask_fg` may also be the uint8 label map [BS,H,W] (1 = object) that the float
    foreground plane was expanded from."""
    pred = _need(pred, torch.float32, "pred")
    labels = isinstance(sup_mask_fg, torch.Tensor) and sup_mask_fg.dtype == torch.uint8
    sup_mask_fg = _need(sup_mask_fg, torch.uint8, "sup_mask_fg") if labels else _need_loose(sup_mask_fg, "sup_mask_fg")
    h, w = qry_fts.shape[-2:]
    BQ = qry_fts.shape[0] * qry_fts.shape[1] if qry_fts.dim() == 5 else qry_fts.shape[0]
    BS = sup_fts.shape[0] * sup_fts.shape[1] if sup_fts.dim() == 5 else sup_fts.shape[0]
    if BQ % Q:
        raise ValueError("qry_fts batch is not a multiple of Q")
    B = BQ // Q
    if BS % B:
        raise ValueError("sup_fts batch is not a multiple of B")
    S = BS // B
    qry_fts, qep, c, _ = _episodes(qry_fts, B, Q, "qry_fts")
    sup_fts, sep, _, _ = _episodes(sup_fts, B, S, "sup_fts")
    H, W = sup_mask_fg.shape[-2:]
    m = sup_mask_fg.reshape(BS, H * W) if sup_mask_fg.dim() != 2 else sup_mask_fg
    if m.stride(1) != 1 or (BS > 1 and m.stride(0) < H * W):
        m = m.contiguous()
    stride = m.stride(0) if BS > 1 else H * W
    L = _cabi.lib()
    ws = _ws(L.pemp_panet_align_workspace_bytes(B, S, Q, c, h, w, H, W), qry_fts.device)
    loss = torch.empty((), dtype=torch.float32, device=qry_fts.device)
    fn = L.pemp_panet_align_labels if labels else L.pemp_panet_align
    _cabi.check(fn(qry_fts.data_ptr(), qep, pred.data_ptr(), sup_fts.data_ptr(), sep, m.data_ptr(), stride, B, S, Q,
                   c, h, w, H, W, float(scalar), loss.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                "pemp_panet_align_labels" if labels else "pemp_panet_align")
    _count(6)
    return loss


# ------------------------------------------------------------------------------------------------ K9
PRIOR_BF16, PRIOR_FP32, PRIOR_BF16X3 = 0, 1, 2
PRIOR_DEFAULT = PRIOR_BF16X3      # tensor-core path with fp32-grade cosines (three bf16 products of a hi/lo split)


@_on_device
def prior_mask(q4, s4, smask, precision=PRIOR_DEFAULT, want_rowmax=False):
    """q4 [B, C, hq, wq]; s4 [S, B, C, hs, ws]; smask [S, B, hs, ws] (support mask at feature size)
    -> prior [B, 1, hq, wq] (and rowmax [S, B, hq*wq])  (pfenet.py:201-231)."""
    q4 = _need(q4, torch.float32, "q4")
    s4 = _need(s4, torch.float32, "s4")
    smask = _need(smask, torch.float32, "smask")
    B, C, hq, wq = q4.shape
    S = s4.shape[0]
    hs, ws_ = s4.shape[-2:]
    if s4.shape[:3] != (S, B, C) or smask.numel() != S * B * hs * ws_:
        raise ValueError("expected s4 [S,B,C,h,w] and smask [S,B,h,w]")
    L = _cabi.lib()
    ws = _ws(L.pemp_prior_mask_workspace_bytes(B, S, C, hs * ws_, hq * wq, precision), q4.device)
    prior = torch.empty(B, 1, hq, wq, dtype=torch.float32, device=q4.device)
    rowmax = torch.empty(S, B, hq * wq, dtype=torch.float32, device=q4.device) if want_rowmax else None
    _cabi.check(L.pemp_prior_mask(q4.data_ptr(), s4.data_ptr(), smask.data_ptr(), B, S, C, hs * ws_, hq * wq, precision,
                                  prior.data_ptr(), _ptr(rowmax), ws.data_ptr(), ws.numel(), _stream()),
                "pemp_prior_mask")
    _count(4)
    return (prior, rowmax) if want_rowmax else prior


# ------------------------------------------------------------------------------------------------ K10
@_on_device
def upsample_argmax_hist(pred, out_hw, ref, cls, stat):
    """K4 + K10 in one launch: pred [N,2,h,w] -> uint8 argmax mask [N,H,W] at `out_hw`, and the FewShotMetric counts of
    (mask, ref [N,H,W] uint8, cls [N]) added to stat [(C+1),3] (entry/pemp_stage2.py:63-65, core/metrics.py:9-23)."""
    pred = _need(pred, torch.float32, "pred")
    ref = _need(ref, torch.uint8, "ref")
    cls = _need(cls, torch.int64, "cls")
    stat = _need(stat, torch.int64, "stat")
    N, two, h, w = pred.shape
    H, W = int(out_hw[0]), int(out_hw[1])
    if two != 2 or ref.numel() != N * H * W or cls.numel() != N:
        raise ValueError("pred must be [N,2,h,w], ref [N,H,W] at the output size, cls [N]")
    if stat.dim() != 2 or stat.shape[1] != 3:
        raise ValueError("stat must be [(classes+1), 3]")
    m8 = torch.empty(N, H, W, dtype=torch.uint8, device=pred.device)
    _cabi.check(_cabi.lib().pemp_upsample_argmax_hist(pred.data_ptr(), N, h, w, H, W, m8.data_ptr(), ref.data_ptr(), cls.data_ptr(),
                                                      stat.shape[0] - 1, stat.data_ptr(), _stream()), "pemp_upsample_argmax_hist")
    _count(1)
    return m8


@_on_device
def iou_hist(pred, ref, cls, stat):
    """Accumulate `FewShotMetric.update` counts on the device.  pred, ref [N, ...] uint8 (same shape);
    cls [N] int64; stat [(C+1), 3] int64 is updated in place (core/metrics.py:9-23)."""
    pred = _need(pred, torch.uint8, "pred")
    ref = _need(ref, torch.uint8, "ref")
    cls = _need(cls, torch.int64, "cls")
    stat = _need(stat, torch.int64, "stat")
    N = pred.shape[0]
    if ref.numel() != pred.numel() or cls.numel() != N:
        raise ValueError("pred / ref / cls disagree in size")
    if stat.dim() != 2 or stat.shape[1] != 3:
        raise ValueError("stat must be [(classes+1), 3]")
    _cabi.check(_cabi.lib().pemp_iou_hist(pred.data_ptr(), ref.data_ptr(), cls.data_ptr(), N, pred.numel() // N,
                                          stat.shape[0] - 1, stat.data_ptr(), _stream()), "pemp_iou_hist")
    _count(1)
    return stat


# ------------------------------------------------------------------------------------------------ K12 (training path)
@_on_device
def meta_proto_attn_train(fts, ctr, fg, bg, B, S, eps=1e-6):
    """K2 forward for training: -> (fg_proto [B,c,p], bg_proto [B,c,p], saved) where `saved` is what
    `meta_proto_attn_bwd` needs (per-shot centres [BS,c,2p] and denominators [BS,2p])."""
    fts, ep, c, hw = _episodes(fts, B, S, "fts")
    ctr = _need(ctr, torch.float32, "ctr")
    n_img = B * S
    if ctr.dim() != 2 or ctr.shape[0] != c or ctr.shape[1] % 2:
        raise ValueError(f"ctr must be [c, 2p] with c = {c}, got {tuple(ctr.shape)}")
    p = ctr.shape[1] // 2
    fg = _need_loose(fg, "fg").reshape(n_img, hw)
    bg = _need_loose(bg, "bg").reshape(n_img, hw)
    fgp, bgp, stride, keep = _mask_pair(fg, bg, n_img, hw)
    L = _cabi.lib()
    dev = fts.device
    ws = _ws(L.pemp_meta_proto_attn_workspace_bytes(B, S, c, hw, p), dev)
    out_f = torch.empty(B, c, p, dtype=torch.float32, device=dev)
    out_b = torch.empty(B, c, p, dtype=torch.float32, device=dev)
    centre = torch.empty(n_img, c, 2 * p, dtype=torch.float32, device=dev)
    den = torch.empty(n_img, 2 * p, dtype=torch.float32, device=dev)
    _cabi.check(L.pemp_meta_proto_attn_train(fts.data_ptr(), ep, ctr.data_ptr(), fgp, bgp, stride, B, S, c, hw, p, float(eps),
                                             out_f.data_ptr(), out_b.data_ptr(), centre.data_ptr(), den.data_ptr(),
                                             ws.data_ptr(), ws.numel(), _stream()), "pemp_meta_proto_attn_train")
    _count(2 if (c == 512 and p == 3 and hw >= 32) or p == 1 else 3)
    return out_f, out_b, (fts, ep, ctr, keep[0], keep[1] if keep[1] is not None else keep[0], centre, den)


def _grad_out(out, B, S, c, hw, device):
    """Gradient destination: a fresh dense [B*S, c, hw] tensor, or `out` = a [B, S, c, h, w] slice of the gradient of the
    encoder output (written in place through its episode stride).  -> (tensor, episode_stride)"""
    if out is None:
        return torch.empty(B * S, c, hw, dtype=torch.float32, device=device), 0
    if out.dtype != torch.float32 or not out.is_cuda or out.dim() != 5 or tuple(out.shape[:3]) != (B, S, c) or \
            out.shape[3] * out.shape[4] != hw:
        raise ValueError(f"out must be a CUDA float32 [{B}, {S}, {c}, h, w] tensor")
    if not (out.stride(4) == 1 and out.stride(3) == out.shape[4] and out.stride(2) == hw and (S == 1 or out.stride(1) == c * hw)):
        raise ValueError("out must be contiguous inside an episode")
    return out, (out.stride(0) if B > 1 else S * c * hw)


@_on_device
def meta_proto_attn_bwd(saved, g_fg, g_bg, B, S, out=None):
    """-> (d_fts [B*S, c, hw] or `out`, d_ctr [c, 2p]) from the gradients of fg_proto / bg_proto [B, c, p]."""
    fts, ep, ctr, fg, bg, centre, den = saved
    c, p = ctr.shape[0], ctr.shape[1] // 2
    hw = fg.shape[1]
    g_fg = _need(g_fg, torch.float32, "g_fg")
    g_bg = _need(g_bg, torch.float32, "g_bg")
    if tuple(g_fg.shape) != (B, c, p) or tuple(g_bg.shape) != (B, c, p):
        raise ValueError(f"gradients must be [{B}, {c}, {p}]")
    fgp, bgp, stride, keep = _mask_pair(fg, bg, B * S, hw)
    L = _cabi.lib()
    dev = fts.device
    ws = _ws(L.pemp_meta_proto_attn_bwd_workspace_bytes(B, S, c, hw, p), dev)
    d_fts, d_ep = _grad_out(out, B, S, c, hw, dev)
    d_ctr = torch.empty(c, 2 * p, dtype=torch.float32, device=dev)
    _cabi.check(L.pemp_meta_proto_attn_bwd(fts.data_ptr(), ep, ctr.data_ptr(), fgp, bgp, stride, centre.data_ptr(), den.data_ptr(),
                                           g_fg.data_ptr(), g_bg.data_ptr(), B, S, c, hw, p, d_fts.data_ptr(), d_ep, d_ctr.data_ptr(),
                                           ws.data_ptr(), ws.numel(), _stream()), "pemp_meta_proto_attn_bwd")
    _count(3)
    del keep
    return d_fts, d_ctr


@_on_device
def cosine_match_bwd(qry, fg_proto, bg_proto, g_pred, scalar=20.0, out=None, dense=False):
    """Backward of `cosine_match(...)["pred"]`: qry as in the forward, g_pred [N, 2, hw] -> (d_qry [N, c, hw] or `out`,
    d_fg, d_bg shaped like the prototypes).  dense=True: g_pred is the gradient [N, 2, P, hw] of `["sim"]`, the per-prototype
    maps `compute_similarity` returns (no arg-max)."""
    fg_proto = _need(fg_proto, torch.float32, "fg_proto")
    bg_proto = _need(bg_proto, torch.float32, "bg_proto")
    g_pred = _need(g_pred, torch.float32, "g_pred")
    Bp = fg_proto.shape[0]
    P = 1 if fg_proto.dim() == 2 else fg_proto.shape[2]
    n_maps = qry.shape[0] * qry.shape[1] if qry.dim() == 5 else qry.shape[0]
    qry, ep, c, hw = _episodes(qry, Bp, n_maps // Bp, "qry")
    want = (n_maps, 2, P, hw) if dense else (n_maps, 2, hw)
    if tuple(g_pred.shape) != want:
        raise ValueError(f"g_pred must be {list(want)}, got {tuple(g_pred.shape)}")
    L = _cabi.lib()
    dev = qry.device
    ws = _ws(L.pemp_cosine_match_bwd_workspace_bytes(n_maps, Bp, c, hw, P), dev)
    d_qry, d_ep = _grad_out(out, Bp, n_maps // Bp, c, hw, dev)
    d_fg, d_bg = torch.empty_like(fg_proto), torch.empty_like(bg_proto)
    fn = L.pemp_cosine_sim_bwd if dense else L.pemp_cosine_match_bwd
    _cabi.check(fn(qry.data_ptr(), ep, fg_proto.data_ptr(), bg_proto.data_ptr(), g_pred.data_ptr(), n_maps, Bp,
                   c, hw, P, float(scalar), d_qry.data_ptr(), d_ep, d_fg.data_ptr(), d_bg.data_ptr(),
                   ws.data_ptr(), ws.numel(), _stream()), "pemp_cosine_sim_bwd" if dense else "pemp_cosine_match_bwd")
    _count(3)
    return d_qry, d_fg, d_bg


def _labels(target):
    if not isinstance(target, torch.Tensor) or not target.is_cuda or target.dtype not in (torch.int64, torch.uint8):
        raise ValueError("target must be a CUDA int64 or uint8 tensor")
    if target.dim() != 3:
        raise ValueError("target must be [N, H, W]")
    return target.contiguous()


@_on_device
def boundary_weight(target, sigma):
    """`CELossDT.boundary2weight` on the device (core/losses.py:23-40): target [N,H,W] -> weight [N,H,W] float32."""
    target = _labels(target)
    N, H, W = target.shape
    L = _cabi.lib()
    ws = _ws(L.pemp_boundary_weight_workspace_bytes(N, H, W), target.device)
    weight = torch.empty(N, H, W, dtype=torch.float32, device=target.device)
    _cabi.check(L.pemp_boundary_weight(target.data_ptr(), int(target.dtype == torch.uint8), N, H, W, float(sigma),
                                       weight.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "pemp_boundary_weight")
    _count(3)
    return weight


@_on_device
def upsample_ce(pred, target, want_grad=True, weight=None):
    """Cross entropy (255 ignored) of the bilinear align_corners up-sampling of pred [N,2,h,w] to target [N,H,W]
    (int64 or uint8) -> (loss [1], d_pred [N,2,h,w] or None).  weight None: mean over the valid pixels
    (entry/pemp_stage1.py:51,57-60); weight [N,H,W]: sum(w * ce) / sum(w), CELossDT (core/losses.py:33-43)."""
    pred = _need(pred, torch.float32, "pred")
    target = _labels(target)
    N, two, h, w = pred.shape
    if two != 2 or target.dim() != 3 or target.shape[0] != N:
        raise ValueError("pred must be [N,2,h,w] and target [N,H,W]")
    H, W = target.shape[1:]
    if weight is not None:
        weight = _need(weight, torch.float32, "weight")
        if tuple(weight.shape) != (N, H, W):
            raise ValueError("weight must have the shape of target")
    L = _cabi.lib()
    ws = _ws(L.pemp_upsample_ce_workspace_bytes(N, h, w, H, W), pred.device)
    loss = torch.empty(1, dtype=torch.float32, device=pred.device)
    d_pred = torch.empty_like(pred) if want_grad else None
    _cabi.check(L.pemp_upsample_ce(pred.data_ptr(), target.data_ptr(), int(target.dtype == torch.uint8), _ptr(weight), N, h, w, H, W,
                                   loss.data_ptr(), _ptr(d_pred), ws.data_ptr(), ws.numel(), _stream()), "pemp_upsample_ce")
    _count(4 if want_grad else 2)
    return loss, d_pred


@_on_device
def map_pool_lowres_bwd(fg, bg, g_fg, g_bg, B, S, c, eps=1e-5, out=None):
    """Backward of `map_pool_lowres`: masks [B*S, hw], g_fg / g_bg [B, c] -> d_fts [B*S, c, hw] (or `out`, a
    [B, S, c, h, w] slice of the gradient of the encoder output)."""
    n_img = B * S
    fg = _need_loose(fg, "fg")
    hw = fg.numel() // n_img
    fg = fg.reshape(n_img, hw)
    bg = None if bg is None else _need_loose(bg, "bg").reshape(n_img, hw)
    g_fg = _need(g_fg, torch.float32, "g_fg")
    g_bg = None if bg is None else _need(g_bg, torch.float32, "g_bg")
    if tuple(g_fg.shape) != (B, c):
        raise ValueError(f"g_fg must be [{B}, {c}]")
    fgp, bgp, stride, keep = _mask_pair(fg, bg, n_img, hw)
    d_fts, d_ep = _grad_out(out, B, S, c, hw, fg.device)
    _cabi.check(_cabi.lib().pemp_map_pool_lowres_bwd(fgp, bgp, stride, g_fg.data_ptr(), _ptr(g_bg), B, S, c, hw, float(eps),
                                                     d_fts.data_ptr(), d_ep, _stream()), "pemp_map_pool_lowres_bwd")
    _count(1)
    del keep
    return d_fts


# ------------------------------------------------------------------------------------------------ K15 (CaNet)
@_on_device
def canet_map_tile(features, sup_mask, B, S, Q):
    """CaNet's dense-comparison input (networks/canet.py:172-180): features [B*(S+Q), c, h, w] (encoder output, read in
    place), sup_mask [B, S, 2, H, W] -> out [B*Q, 2c, h, w] = cat(query features, foreground prototype tiled)."""
    features = _need(features, torch.float32, "features")
    sup_mask = _need(sup_mask, torch.float32, "sup_mask")
    _, c, h, w = features.shape
    f5 = features.view(B, S + Q, c, h, w)
    H, W = sup_mask.shape[-2:]
    low = mask_nearest(sup_mask.view(B * S, 2, H, W), h, w).view(B * S, 2, h * w)
    z, _ = map_pool_lowres(f5[:, :S], low[:, 0], None, B, S, eps=1e-5)
    qry, ep, _, hw = _episodes(f5[:, S:], B, Q, "qry")
    out = torch.empty(B * Q, 2 * c, h, w, dtype=torch.float32, device=features.device)
    _cabi.check(_cabi.lib().pemp_canet_concat(qry.data_ptr(), ep, z.data_ptr(), B * Q, B, c, hw, out.data_ptr(), _stream()),
                "pemp_canet_concat")
    _count(1)
    return out
