"""K12: gradient parity of the training path.  The reference trains with autograd over the PyTorch ops of
`pemp_stage1.py:202-261`; the oracle restates those ops in torch, so autograd over the ORACLE (float64) is the checker
for the hand-written backward kernels.  Tolerance: 2e-5 of the largest gradient entry (fp32 kernels, float64 checker);
where the op itself is ill conditioned in fp32 (soft-max over squared distances summed over >= 512 channels) the bar is
the error of fp32 autograd over the same ops - i.e. of the reference's own training step - against the float64 result
(`tools/probes/grad_error_probe.py`: 3-4e-5 for it, 0.4-2.5e-5 for the kernels)."""
import numpy as np
import pytest
import torch

from conftest import nrel
from oracle import restate as O

pytestmark = pytest.mark.gpu
GTOL = 2e-5


def _case(B, S, Q, c, h, w, P, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, S + Q, c, h, w, generator=g) * scale
    ctr = torch.randn(c, 2 * P, generator=g) * 0.5 * scale
    fg = (torch.rand(B * S, h * w, generator=g) > 0.6).float()
    fg[0, : h * w // 3] = torch.rand(h * w // 3, generator=g)            # soft mask values are legal too
    bg = 1.0 - fg
    return feats, ctr, fg, bg


@pytest.mark.parametrize("B,S,c,h,w,P", [(2, 2, 64, 9, 11, 3), (1, 1, 512, 51, 51, 3), (2, 5, 512, 13, 13, 3),
                                          (3, 1, 32, 7, 5, 1), (1, 2, 128, 8, 9, 4), (1, 3, 1024, 6, 7, 2),
                                          # the tensor-path kernel (P = 3, c in {128, 256, 512, 1024}), ragged last tiles
                                          (2, 3, 128, 9, 11, 3), (1, 2, 256, 13, 7, 3), (1, 1, 1024, 6, 7, 3), (1, 2, 512, 5, 5, 3)])
def test_meta_proto_attn_backward_matches_autograd_of_the_oracle(B, S, c, h, w, P):
    from pemp_b200 import autograd as A
    feats, ctr, fg, bg = _case(B, S, 1, c, h, w, P, seed=c + h)
    g = torch.Generator().manual_seed(7)
    wf, wb = torch.randn(B, c, P, generator=g), torch.randn(B, c, P, generator=g)
    # checker: float64 autograd over the oracle (and fp32 autograd over it = the reference's own arithmetic)
    sup64 = feats[:, :S].double().reshape(B * S, c, h * w).requires_grad_(True)
    ctr64 = ctr.double().requires_grad_(True)
    of, ob, _ = O.meta_proto_attention(sup64, fg.double(), bg.double(), ctr64, B, S, P)
    ((of * wf.double()).sum() + (ob * wb.double()).sum()).backward()
    sup32 = feats[:, :S].reshape(B * S, c, h * w).clone().requires_grad_(True)
    ctr32 = ctr.clone().requires_grad_(True)
    of32, ob32, _ = O.meta_proto_attention(sup32, fg, bg, ctr32, B, S, P)
    ((of32 * wf).sum() + (ob32 * wb).sum()).backward()
    tol_f = max(GTOL, nrel(sup32.grad, sup64.grad.float()))
    tol_c = max(GTOL, nrel(ctr32.grad, ctr64.grad.float()))
    # kernels (the support features are a strided slice of the encoder output, read in place)
    f_cu = feats.cuda().requires_grad_(True)
    ctr_cu = ctr.cuda().requires_grad_(True)
    kf, kb = A.meta_proto_attn(f_cu[:, :S], ctr_cu, fg.cuda(), bg.cuda())
    assert nrel(kf.detach().cpu(), of.detach().float()) < 1e-5 and nrel(kb.detach().cpu(), ob.detach().float()) < 1e-5
    ((kf * wf.cuda()).sum() + (kb * wb.cuda()).sum()).backward()
    d_sup = f_cu.grad[:, :S].reshape(B * S, c, h * w).cpu()
    assert float(f_cu.grad[:, S:].abs().max()) == 0.0
    assert nrel(d_sup, sup64.grad.float()) <= tol_f
    assert nrel(ctr_cu.grad.cpu(), ctr64.grad.float()) <= tol_c


def test_meta_proto_attn_backward_tensor_path_vs_cuda_core_kernel_full_size():
    """The two implementations of the K2 backward (train_mma.cu: 3 x TF32 on the warp-level tensor path; train.cu: CUDA cores)
    on the same full-size input through `pemp_debug_bwd_path`: they agree to fp32 summation-order noise (the soft-max over
    squared distances is ill conditioned: fp32 autograd of the reference's ops is itself 3-4e-5 from float64)."""
    from pemp_b200 import _cabi, ops
    B, S, c, h, P = 2, 5, 512, 51, 3
    feats, ctr, fg, bg = _case(B, S, 1, c, h, h, P, seed=11)
    g = torch.Generator().manual_seed(3)
    gf, gb = torch.randn(B, c, P, generator=g).cuda(), torch.randn(B, c, P, generator=g).cuda()
    f_cu, ctr_cu, fg_cu, bg_cu = feats.cuda(), ctr.cuda(), fg.cuda(), bg.cuda()
    _, _, saved = ops.meta_proto_attn_train(f_cu[:, :S], ctr_cu, fg_cu, bg_cu, B, S)
    d1, c1 = ops.meta_proto_attn_bwd(saved, gf, gb, B, S)
    d1b, c1b = ops.meta_proto_attn_bwd(saved, gf, gb, B, S)
    assert torch.equal(d1, d1b) and torch.equal(c1, c1b)            # deterministic
    _cabi.lib().pemp_debug_bwd_path(1)
    try:
        d0, c0 = ops.meta_proto_attn_bwd(saved, gf, gb, B, S)
    finally:
        _cabi.lib().pemp_debug_bwd_path(0)
    ef, ec = nrel(d1.cpu(), d0.cpu()), nrel(c1.cpu(), c0.cpu())
    print({"case": "K2 backward tensor path vs CUDA cores", "d_fts": ef, "d_ctr": ec})
    assert ef < 5e-5 and ec < 5e-5


def test_meta_proto_attn_backward_survives_many_launches_at_bench_size():
    """The tensor-path K2 backward hands work between warps through mbarriers (boxes, dots, weights) with parity waits; a
    hazard in such a protocol shows up as a launch failure or a changed bit only under timing variation.  200 launches at
    the training-bench size (64 episodes: 320 images, 1 CTA per SM, ~12 waves), a competing stream writing memory
    meanwhile, every 50th result compared bit for bit with the first."""
    from pemp_b200 import ops
    B, S, c, h, P = 64, 5, 512, 51, 3
    g = torch.Generator(device="cuda").manual_seed(4)
    feats = torch.randn(B, S, c, h, h, device="cuda", generator=g) * 0.5
    ctr = torch.randn(c, 2 * P, device="cuda", generator=g) * 0.5
    fg = (torch.rand(B * S, h * h, device="cuda", generator=g) > 0.6).float()
    bg = 1 - fg
    gf, gb = torch.randn(B, c, P, device="cuda", generator=g), torch.randn(B, c, P, device="cuda", generator=g)
    _, _, saved = ops.meta_proto_attn_train(feats, ctr, fg, bg, B, S)
    side = torch.cuda.Stream()
    noise = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    ref = None
    for it in range(200):
        if it % 4 == 0:
            with torch.cuda.stream(side):
                noise.fill_(float(it))                       # 256 MB of competing writes: shifts the kernel's timing
        d, dc = ops.meta_proto_attn_bwd(saved, gf, gb, B, S)
        if it % 50 == 49:
            torch.cuda.synchronize()
            if ref is None:
                ref = (d.clone(), dc.clone())
            else:
                assert torch.equal(ref[0], d) and torch.equal(ref[1], dc), f"launch {it} differs"
    torch.cuda.synchronize()
    assert bool(torch.isfinite(ref[0]).all()) and bool(torch.isfinite(ref[1]).all())


def test_meta_proto_attn_backward_falls_back_when_the_operand_has_no_tensor_map():
    """The tensor-path kernel reads the features through a TMA tensor map (16-byte aligned base); a support map that starts
    4 bytes into an allocation cannot be encoded, and the entry point must take the CUDA-core kernel for it - bit for bit what
    `pemp_debug_bwd_path(1)` gives on an aligned copy of the same values."""
    from pemp_b200 import _cabi, ops
    B, S, c, h, P = 1, 2, 256, 13, 3
    feats, ctr, fg, bg = _case(B, S, 1, c, h, h, P, seed=5)
    g = torch.Generator().manual_seed(9)
    gf, gb = torch.randn(B, c, P, generator=g).cuda(), torch.randn(B, c, P, generator=g).cuda()
    sup = feats[:, :S].reshape(B * S, c, h * h).cuda().contiguous()
    store = torch.empty(sup.numel() + 1, dtype=torch.float32, device="cuda")
    off = store[1:].view_as(sup)                                   # same values, base pointer 4 bytes off a 16-byte boundary
    off.copy_(sup)
    assert off.data_ptr() % 16 != 0 and sup.data_ptr() % 16 == 0
    ctr_cu, fg_cu, bg_cu = ctr.cuda(), fg.cuda(), bg.cuda()
    _, _, saved_a = ops.meta_proto_attn_train(sup, ctr_cu, fg_cu, bg_cu, B, S)
    _, _, saved_o = ops.meta_proto_attn_train(off, ctr_cu, fg_cu, bg_cu, B, S)
    d_o, c_o = ops.meta_proto_attn_bwd(saved_o, gf, gb, B, S)      # automatic choice: must fall back
    _cabi.lib().pemp_debug_bwd_path(1)
    try:
        d_a, c_a = ops.meta_proto_attn_bwd(saved_a, gf, gb, B, S)  # CUDA-core kernel on the aligned copy
    finally:
        _cabi.lib().pemp_debug_bwd_path(0)
    d_t, c_t = ops.meta_proto_attn_bwd(saved_a, gf, gb, B, S)      # tensor path on the aligned copy: close, not identical
    assert nrel(d_o.cpu(), d_a.cpu()) < 2e-5 and nrel(c_o.cpu(), c_a.cpu()) < 2e-5
    assert nrel(d_t.cpu(), d_a.cpu()) < 5e-5 and nrel(c_t.cpu(), c_a.cpu()) < 5e-5


@pytest.mark.parametrize("B,Q,c,h,w,P", [(2, 1, 64, 9, 11, 3), (1, 1, 512, 51, 51, 3), (2, 2, 512, 13, 13, 3),
                                          (3, 1, 32, 7, 5, 1), (1, 2, 128, 8, 9, 4), (2, 1, 1024, 6, 7, 2),
                                          # the tensor-path kernel (P = 3, c in {256, 512}), ragged last tiles, Q > 1
                                          (2, 1, 256, 9, 11, 3), (1, 3, 256, 13, 7, 3), (2, 2, 512, 7, 9, 3)])
def test_cosine_match_backward_matches_autograd_of_the_oracle(B, Q, c, h, w, P):
    from pemp_b200 import autograd as A
    g = torch.Generator().manual_seed(c + w)
    qry = torch.randn(B, Q, c, h, w, generator=g)
    shape = (B, c, P) if P > 1 else (B, c)
    fgp, bgp = torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)
    wgt = torch.randn(B * Q, 2, h, w, generator=g)
    q64 = qry.double().reshape(B * Q, c, h * w).requires_grad_(True)
    f64, b64 = fgp.double().requires_grad_(True), bgp.double().requires_grad_(True)
    pred64, _ = O.reduce_over_protos(O.cosine_match(q64, f64, b64, 20.0))
    (pred64 * wgt.double().view(B * Q, 2, h * w)).sum().backward()
    q_cu = qry.cuda().requires_grad_(True)
    f_cu, b_cu = fgp.cuda().requires_grad_(True), bgp.cuda().requires_grad_(True)
    pred = A.cosine_match(q_cu, f_cu, b_cu, 20.0)
    assert nrel(pred.detach().cpu().view(B * Q, 2, h * w), pred64.detach().float()) < 1e-5
    (pred * wgt.cuda()).sum().backward()
    assert nrel(q_cu.grad.cpu().view(B * Q, c, h * w), q64.grad.float()) < GTOL
    assert nrel(f_cu.grad.cpu(), f64.grad.float()) < GTOL
    assert nrel(b_cu.grad.cpu(), b64.grad.float()) < GTOL


def test_cosine_match_backward_tensor_path_vs_cuda_core_kernel_full_size():
    """The two implementations of the K3 backward (train_mma_cos.cu / train.cu) on the same full-size input, arg-max and dense
    form, through `pemp_debug_bwd_path`; the arg-max is recomputed by both, so the inputs avoid near-ties (random features)."""
    from pemp_b200 import _cabi, ops
    B, Q, c, h, P = 4, 1, 512, 51, 3
    g = torch.Generator().manual_seed(17)
    qry = torch.randn(B * Q, c, h * h, generator=g).cuda()
    fgp, bgp = torch.randn(B, c, P, generator=g).cuda(), torch.randn(B, c, P, generator=g).cuda()
    for dense in (False, True):
        gp = (torch.randn(B * Q, 2, P, h * h, generator=g) if dense else torch.randn(B * Q, 2, h * h, generator=g)).cuda()
        new = ops.cosine_match_bwd(qry, fgp, bgp, gp, dense=dense)
        again = ops.cosine_match_bwd(qry, fgp, bgp, gp, dense=dense)
        assert all(torch.equal(a, b) for a, b in zip(new, again))             # deterministic
        _cabi.lib().pemp_debug_bwd_path(1)
        try:
            old = ops.cosine_match_bwd(qry, fgp, bgp, gp, dense=dense)
        finally:
            _cabi.lib().pemp_debug_bwd_path(0)
        errs = [nrel(a.cpu(), b.cpu()) for a, b in zip(new, old)]
        print({"case": "K3 backward tensor path vs CUDA cores", "dense": dense, "d_qry": errs[0], "d_fg": errs[1], "d_bg": errs[2]})
        assert max(errs) < 2e-5


def test_cosine_match_backward_with_tiny_vectors_uses_the_clamped_branch():
    """|q| < eps and |proto| < eps: F.cosine_similarity clamps the norm, so the projection term of the gradient vanishes."""
    from pemp_b200 import autograd as A
    B, Q, c, h, w, P = 1, 1, 16, 4, 8, 3
    g = torch.Generator().manual_seed(5)
    qry = torch.randn(B, Q, c, h, w, generator=g)
    qry[0, 0, :, 0, :3] = 0.0
    qry[0, 0, :, 1, :2] *= 1e-12
    fgp, bgp = torch.randn(B, c, P, generator=g), torch.randn(B, c, P, generator=g)
    bgp[0, :, 1] = 0.0
    wgt = torch.randn(B * Q, 2, h, w, generator=g)
    q64 = qry.double().reshape(B * Q, c, h * w).requires_grad_(True)
    f64, b64 = fgp.double().requires_grad_(True), bgp.double().requires_grad_(True)
    pred64, _ = O.reduce_over_protos(O.cosine_match(q64, f64, b64, 20.0))
    (pred64 * wgt.double().view(B * Q, 2, h * w)).sum().backward()
    q_cu = qry.cuda().requires_grad_(True)
    f_cu, b_cu = fgp.cuda().requires_grad_(True), bgp.cuda().requires_grad_(True)
    (A.cosine_match(q_cu, f_cu, b_cu, 20.0) * wgt.cuda()).sum().backward()
    live = torch.ones(h * w, dtype=torch.bool)
    live[:3] = False                                      # exactly-zero q: sub-gradient of max/ties is implementation defined
    assert nrel(q_cu.grad.cpu().view(c, h * w)[:, live], q64.grad.float()[0][:, live]) < GTOL
    assert nrel(f_cu.grad.cpu(), f64.grad.float()) < GTOL


@pytest.mark.parametrize("B,S,Q,c,h,w,H,W", [(2, 2, 1, 64, 9, 9, 33, 33), (1, 5, 1, 512, 51, 51, 401, 401), (2, 1, 2, 128, 13, 11, 50, 41)])
def test_head_loss_gradients_match_the_reference_training_step(B, S, Q, c, h, w, H, W):
    """`entry/pemp_stage1.py:57-65`: loss = CE(up-sampled pred, query mask with 255 ignored); gradients reach the encoder
    output (support and query halves) and the meta-prototype centres."""
    from pemp_b200 import autograd as A
    P = 3
    feats, ctr, fg, bg = _case(B, S, Q, c, h, w, P, seed=11 + c)
    g = torch.Generator().manual_seed(3)
    target = torch.randint(0, 2, (B * Q, H, W), generator=g)
    target[:, :2] = 255
    low = torch.stack((fg, bg), dim=1)                                    # [BS, 2, hw]
    # checker
    f64 = feats.double().requires_grad_(True)
    ctr64 = ctr.double().requires_grad_(True)
    of, ob, _ = O.meta_proto_attention(f64[:, :S].reshape(B * S, c, h * w), fg.double(), bg.double(), ctr64, B, S, P)
    p64, _ = O.reduce_over_protos(O.cosine_match(f64[:, S:].reshape(B * Q, c, h * w), of, ob, 20.0))
    lg64 = torch.nn.functional.interpolate(p64.view(B * Q, 2, h, w), size=(H, W), mode="bilinear", align_corners=True)
    loss64 = torch.nn.functional.cross_entropy(lg64, target, ignore_index=255)
    loss64.backward()
    f32 = feats.clone().requires_grad_(True)
    ctr32 = ctr.clone().requires_grad_(True)
    of32, ob32, _ = O.meta_proto_attention(f32[:, :S].reshape(B * S, c, h * w), fg, bg, ctr32, B, S, P)
    p32, _ = O.reduce_over_protos(O.cosine_match(f32[:, S:].reshape(B * Q, c, h * w), of32, ob32, 20.0))
    lg32 = torch.nn.functional.interpolate(p32.view(B * Q, 2, h, w), size=(H, W), mode="bilinear", align_corners=True)
    torch.nn.functional.cross_entropy(lg32, target, ignore_index=255).backward()
    tol_f = max(GTOL, nrel(f32.grad, f64.grad.float()))
    tol_c = max(GTOL, nrel(ctr32.grad, ctr64.grad.float()))
    # kernels
    f_cu = feats.cuda().view(B * (S + Q), c, h, w).requires_grad_(True)
    ctr_cu = ctr.cuda().requires_grad_(True)
    loss, pred = A.pemp_head_loss(f_cu, low.cuda(), ctr_cu, B, S, Q, target.cuda())
    loss.backward()
    assert abs(float(loss.detach()) - float(loss64.detach())) < 1e-5 * max(1.0, abs(float(loss64.detach())))
    assert nrel(f_cu.grad.cpu().view(B, S + Q, c, h, w), f64.grad.float()) <= tol_f
    assert nrel(ctr_cu.grad.cpu(), ctr64.grad.float()) <= tol_c


def test_backward_is_deterministic():
    from pemp_b200 import autograd as A
    B, S, Q, c, h, w, P = 2, 2, 1, 512, 21, 21, 3
    feats, ctr, fg, bg = _case(B, S, Q, c, h, w, P, seed=2)
    grads = []
    for _ in range(2):
        f_cu = feats.cuda().requires_grad_(True)
        ctr_cu = ctr.cuda().requires_grad_(True)
        kf, kb = A.meta_proto_attn(f_cu[:, :S], ctr_cu, fg.cuda(), bg.cuda())
        A.cosine_match(f_cu[:, S:], kf, kb).square().sum().backward()
        grads.append((f_cu.grad.clone(), ctr_cu.grad.clone()))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])


def test_training_ops_refuse_cpu_tensors():
    from pemp_b200 import autograd as A
    feats, ctr, fg, bg = _case(1, 1, 1, 16, 4, 4, 3, seed=1)
    with pytest.raises(ValueError):
        A.meta_proto_attn(feats[:, :1], ctr, fg, bg)
    with pytest.raises(ValueError):
        A.cosine_match(feats[:, 1:], torch.zeros(1, 16, 3), torch.zeros(1, 16, 3))


@pytest.mark.parametrize("N,h,w,H,W,dtype", [(3, 51, 51, 401, 401, torch.int64), (2, 13, 17, 57, 83, torch.uint8), (1, 4, 4, 1, 7, torch.int64),
                                              (4, 51, 51, 333, 500, torch.uint8), (2, 9, 9, 9, 9, torch.int64), (2, 20, 20, 11, 13, torch.int64)])
def test_upsample_ce_loss_and_gradient(N, h, w, H, W, dtype):
    """K13 against F.interpolate + cross_entropy(ignore_index=255) and their autograd in float64."""
    from pemp_b200 import autograd as A
    g = torch.Generator().manual_seed(N * H + w)
    pred = torch.randn(N, 2, h, w, generator=g) * 5
    target = torch.randint(0, 2, (N, H, W), generator=g)
    target[torch.rand(N, H, W, generator=g) < 0.1] = 255
    p64 = pred.double().requires_grad_(True)
    lg = torch.nn.functional.interpolate(p64, size=(H, W), mode="bilinear", align_corners=True)
    loss64 = torch.nn.functional.cross_entropy(lg, target, ignore_index=255)
    (loss64 * 3.0).backward()
    p_cu = pred.cuda().requires_grad_(True)
    loss = A.upsample_ce(p_cu, target.to(dtype).cuda())
    (loss * 3.0).backward()
    assert abs(float(loss.detach()) - float(loss64.detach())) <= 2e-6 * max(1.0, abs(float(loss64.detach())))
    assert nrel(p_cu.grad.cpu(), p64.grad.float()) < 1e-5


def test_upsample_ce_with_nothing_valid_is_nan_like_torch():
    from pemp_b200 import ops
    pred = torch.randn(1, 2, 5, 5).cuda()
    target = torch.full((1, 9, 9), 255, dtype=torch.int64).cuda()
    loss, _ = ops.upsample_ce(pred, target)
    assert torch.isnan(loss).all()


@pytest.mark.parametrize("B,S,Q,c,h,w", [(2, 2, 1, 64, 9, 11), (1, 5, 1, 512, 51, 51), (3, 1, 2, 32, 7, 5)])
def test_baseline_head_gradients(B, S, Q, c, h, w):
    """K1 + K3 with one prototype per class (the baseline / PANet heads in training, entry/panet.py:108-115)."""
    from pemp_b200 import autograd as A
    feats, _, fg, bg = _case(B, S, Q, c, h, w, 1, seed=5 + c)
    g = torch.Generator().manual_seed(9)
    wgt = torch.randn(B * Q, 2, h, w, generator=g)
    f64 = feats.double().requires_grad_(True)
    of, ob = O.map_pool_lowres(f64[:, :S].reshape(B * S, c, h * w), fg.double(), bg.double(), B, S)
    p64, _ = O.reduce_over_protos(O.cosine_match(f64[:, S:].reshape(B * Q, c, h * w), of, ob, 20.0))
    (p64 * wgt.double().view(B * Q, 2, h * w)).sum().backward()
    f_cu = feats.cuda().requires_grad_(True)
    kf, kb = A.map_pool_lowres(f_cu[:, :S], fg.cuda(), bg.cuda())
    pred = A.cosine_match(f_cu[:, S:], kf, kb, 20.0)
    assert nrel(pred.detach().cpu().view(B * Q, 2, h * w), p64.detach().float()) < 1e-5
    (pred * wgt.cuda()).sum().backward()
    assert nrel(f_cu.grad.cpu(), f64.grad.float()) < GTOL


def _panet_reference_losses(f64, sup_mask, target, B, S, Q, scalar=20.0):
    """The oracle's `baseline_head` + `panet_align_loss` with the (NumPy, not differentiable) up-sampler of the oracle
    replaced by `F.interpolate(align_corners=True)` - the op the reference itself calls (baseline.py:100, panet.py:190)."""
    up = lambda t, size: torch.nn.functional.interpolate(t, size=size, mode="bilinear", align_corners=True)
    _, _, c, h, w = f64.shape
    H, W = sup_mask.shape[-2:]
    sup = f64[:, :S].reshape(B * S, c, h, w)
    qry = f64[:, S:].reshape(B * Q, c, h, w)
    big = up(sup, (H, W))
    fgm, bgm = sup_mask[:, 0:1], sup_mask[:, 1:2]
    fgp = ((big * fgm).sum(dim=(2, 3)) / (fgm.sum(dim=(2, 3)) + 1e-5)).view(B, S, c).mean(dim=1)
    bgp = ((big * bgm).sum(dim=(2, 3)) / (bgm.sum(dim=(2, 3)) + 1e-5)).view(B, S, c).mean(dim=1)
    pred = O.cosine_match(qry.reshape(B * Q, c, h * w), fgp, bgp, scalar)[:, :, 0].reshape(B * Q, 2, h, w)
    ce = torch.nn.functional.cross_entropy(up(pred, tuple(target.shape[-2:])), target, ignore_index=255)
    winner = O.argmax2(pred.detach()).view(B * Q, h * w)
    q = qry.reshape(B * Q, c, h * w)
    qf = O.masked_average(q, (winner == 1).double(), 1e-5).view(B, Q, c).mean(dim=1)
    qb = O.masked_average(q, (winner == 0).double(), 1e-5).view(B, Q, c).mean(dim=1)
    rev = O.cosine_match(sup.reshape(B * S, c, h * w), qf, qb, scalar)[:, :, 0].reshape(B * S, 2, h, w)
    al = torch.nn.functional.cross_entropy(up(rev, (H, W)), sup_mask[:, 0].long())
    return ce, al


@pytest.mark.parametrize("B,S,Q,c,h,w,H,W", [(2, 2, 1, 32, 9, 9, 33, 33), (1, 3, 2, 64, 13, 11, 50, 41)])
def test_panet_training_step_gradients(B, S, Q, c, h, w, H, W):
    """`entry/panet.py:108-115`: loss = CE(query) + alignLoss, with the prototypes pooled at mask resolution (K6)."""
    from pemp_b200 import autograd as A
    g = torch.Generator().manual_seed(21 + c)
    feats = torch.randn(B, S + Q, c, h, w, generator=g)
    fgm = (torch.rand(B * S, 1, H, W, generator=g) > 0.6).float()
    sup_mask = torch.cat((fgm, 1.0 - fgm), dim=1)
    target = torch.randint(0, 2, (B * Q, H, W), generator=g)
    target[:, :2] = 255
    f64 = feats.double().requires_grad_(True)
    ce64, al64 = _panet_reference_losses(f64, sup_mask.double(), target, B, S, Q)
    (ce64 + al64).backward()
    f_cu = feats.cuda().requires_grad_(True)
    fgp, bgp = A.map_pool_fullres(f_cu[:, :S], sup_mask.cuda())
    pred = A.cosine_match(f_cu[:, S:], fgp, bgp)
    ce = A.upsample_ce(pred, target.cuda())
    al = A.panet_align_loss(f_cu[:, S:], pred.detach(), f_cu[:, :S], sup_mask[:, 0].cuda())
    (ce + al).backward()
    assert abs(float(ce.detach()) - float(ce64.detach())) < 1e-5 and abs(float(al.detach()) - float(al64.detach())) < 1e-5
    assert nrel(f_cu.grad.cpu(), f64.grad.float()) < GTOL


# ------------------------------------------------------------------------------------------------ K14 (CELossDT)
@pytest.mark.parametrize("name", ["cedt_a", "cedt_b"])
def test_boundary_weight_and_cedt_loss_against_reference_fixtures(name):
    """Fixtures produced by the reference's own `CELossDT` (core/losses.py:17-43, scipy EDT on the host).  The weights come
    from exact integer squared distances, so they must match to the last bit of the float64 -> float32 rounding (at most one
    ulp apart where CUDA's and NumPy's double `exp` differ in their last bit)."""
    from conftest import golden
    from pemp_b200 import autograd as A, ops
    g = golden(name)
    target = torch.from_numpy(g["target"]).cuda()
    sigma = float(g["sigma"])
    for t in (target, target.to(torch.uint8)):
        wgt = ops.boundary_weight(t, sigma).cpu().numpy()
        assert np.abs(wgt - g["weight"]).max() <= 2.4e-7
        assert (wgt != g["weight"]).mean() < 1e-3
    inputs = torch.from_numpy(g["inputs"])
    # the reference feeds logits at the target size: identity "up-sampling" (h, w) == (H, W)
    p_cu = inputs.cuda().requires_grad_(True)
    loss = A.ce_loss_dt(p_cu, target, sigma)
    assert abs(float(loss.detach()) - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    p64 = inputs.double().requires_grad_(True)
    w64 = torch.from_numpy(g["weight"]).double()
    ce = torch.nn.functional.cross_entropy(p64, torch.from_numpy(g["target"]), ignore_index=255, reduction="none")
    ((ce * w64).sum() / w64.sum()).backward()
    loss.backward()
    assert nrel(p_cu.grad.cpu(), p64.grad.float()) < 1e-5


@pytest.mark.parametrize("N,H,W,sigma", [(2, 401, 401, 5.0), (3, 97, 130, 3.0), (1, 5, 7, 1.0)])
def test_boundary_weight_random_shapes_against_the_oracle(N, H, W, sigma):
    from pemp_b200 import ops
    g = torch.Generator().manual_seed(N * H)
    target = torch.zeros(N, H, W, dtype=torch.int64)
    blobs = torch.rand(N, H // 4 + 1, W // 4 + 1, generator=g) > 0.7
    target[:] = torch.nn.functional.interpolate(blobs[:, None].float(), size=(H, W), mode="nearest")[:, 0].long()
    target[torch.rand(N, H, W, generator=g) < 0.02] = 255
    if N > 1:
        target[1] = 0                                     # no foreground, no boundary
    want = O.boundary_weight(target, sigma).numpy()
    got = ops.boundary_weight(target.cuda(), sigma).cpu().numpy()
    assert np.abs(got - want).max() <= 2.4e-7
    assert (got != want).mean() < 1e-3


def test_cedt_training_step_on_upsampled_prediction():
    """`loss=cedt` on the PEMP head: pred [N,2,h,w] up-sampled to the target inside the op."""
    from pemp_b200 import autograd as A
    N, h, w, H, W, sigma = 2, 13, 13, 97, 97, 5.0
    g = torch.Generator().manual_seed(4)
    pred = torch.randn(N, 2, h, w, generator=g) * 4
    target = torch.zeros(N, H, W, dtype=torch.int64)
    target[0, 20:60, 30:80] = 1
    target[1, 5:30, 5:50] = 1
    target[:, :3] = 255
    p64 = pred.double().requires_grad_(True)
    lg = torch.nn.functional.interpolate(p64, size=(H, W), mode="bilinear", align_corners=True)
    want, _ = O.ce_loss_dt(lg, target, sigma)
    want.backward()
    p_cu = pred.cuda().requires_grad_(True)
    loss = A.ce_loss_dt(p_cu, target.cuda(), sigma)
    loss.backward()
    assert abs(float(loss.detach()) - float(want.detach())) <= 2e-6 * abs(float(want.detach()))
    assert nrel(p_cu.grad.cpu(), p64.grad.float()) < 1e-5


def test_dropin_heads_take_the_differentiable_path_under_autograd():
    """The reference trains through the same methods the drop-ins replace: with autograd on, `PEMPHead` / `BaselineHead`
    (and the patched reference classes, which share these functions) produce gradients equal to float64 autograd over the
    oracle; under `torch.no_grad()` (core/base_trainer.py:69) they stay on the forward-only kernels."""
    from pemp_b200 import heads
    B, S, Q, c, h, w, H, W, P = 2, 2, 1, 64, 9, 9, 33, 33, 3
    feats, ctr, _, _ = _case(B, S, Q, c, h, w, P, seed=77)
    g = torch.Generator().manual_seed(5)
    fgm = (torch.rand(B, S, 1, H, W, generator=g) > 0.6).float()
    sup_mask = torch.cat((fgm, 1.0 - fgm), dim=2)
    target = torch.randint(0, 2, (B * Q, H, W), generator=g)
    target[:, :2] = 255
    # --- PEMP head
    head = heads.PEMPHead(out_channels=c, protos=P).cuda()
    with torch.no_grad():
        head.ctr.copy_(ctr.cuda())
    f_cu = feats.cuda().view(B * (S + Q), c, h, w).requires_grad_(True)
    out = head(f_cu, sup_mask.cuda(), B, S, Q)
    assert out.requires_grad and tuple(out.shape) == (B * Q, 2, H, W)
    torch.nn.functional.cross_entropy(out, target.cuda(), ignore_index=255).backward()
    f64 = feats.double().requires_grad_(True)
    c64 = ctr.double().requires_grad_(True)
    low = O.mask_nearest(sup_mask.view(B * S, 2, H, W), h, w).view(B * S, 2, h * w).double()
    of, ob, _ = O.meta_proto_attention(f64[:, :S].reshape(B * S, c, h * w), low[:, 0], low[:, 1], c64, B, S, P)
    p64, _ = O.reduce_over_protos(O.cosine_match(f64[:, S:].reshape(B * Q, c, h * w), of, ob, 20.0))
    lg = torch.nn.functional.interpolate(p64.view(B * Q, 2, h, w), size=(H, W), mode="bilinear", align_corners=True)
    torch.nn.functional.cross_entropy(lg, target, ignore_index=255).backward()
    assert nrel(f_cu.grad.cpu().view(B, S + Q, c, h, w), f64.grad.float()) < GTOL
    assert nrel(head.ctr.grad.cpu(), c64.grad.float()) < GTOL
    with torch.no_grad():
        assert not head(f_cu, sup_mask.cuda(), B, S, Q).requires_grad
    sim = head.compute_similarity(torch.ones(B, c, P).cuda(), torch.ones(B, c, P).cuda(), f_cu[:B].view(B, c, 1, h, w))
    assert sim.requires_grad and tuple(sim.shape) == (B, 2, P, h, w)        # differentiable on its own as well (round 2)
    # --- PANet head (prototypes pooled at mask resolution + alignment loss)
    pa = heads.BaselineHead(align=True).cuda()
    f_cu2 = feats.cuda().view(B * (S + Q), c, h, w).requires_grad_(True)
    out2, align = pa(f_cu2, sup_mask.cuda(), B, S, Q)
    (torch.nn.functional.cross_entropy(out2, target.cuda(), ignore_index=255) + align).backward()
    f64b = feats.double().requires_grad_(True)
    ce64, al64 = _panet_reference_losses(f64b, sup_mask.view(B * S, 2, H, W).double(), target, B, S, Q)
    (ce64 + al64).backward()
    assert abs(float(align.detach()) - float(al64.detach())) < 1e-5
    assert nrel(f_cu2.grad.cpu().view(B, S + Q, c, h, w), f64b.grad.float()) < GTOL


def test_bench_size_properties_of_the_training_step():
    """BASELINE size (c = 512, 51 x 51 features, 401 x 401 targets, 5-shot; 16 episodes) through properties that need no
    oracle: (1) the batch holds every episode twice (b and b + 8): the two copies get bit-identical feature gradients
    (same per-image arithmetic, different CTAs), and a run over one half alone gives the same loss, twice the per-episode
    feature gradient (half as many valid pixels in the mean) and the same centre gradient; (2) the backward is linear in
    the upstream gradient; (3) gradients are finite, and a support image whose masks are all zero gets no gradient."""
    from pemp_b200 import autograd as A, ops
    B, S, Q, c, h, w, H, W, P = 16, 5, 1, 512, 51, 51, 401, 401, 3
    g = torch.Generator(device="cuda").manual_seed(3)
    half = torch.randn(B // 2, S + Q, c, h, w, device="cuda", generator=g)
    feats = torch.cat((half, half), dim=0).contiguous()
    ctr = torch.rand(c, 2 * P, device="cuda", generator=g)
    fg_h = (torch.rand(B // 2 * S, h * w, device="cuda", generator=g) > 0.6).float()
    fg = torch.cat((fg_h, fg_h), dim=0)
    low = torch.stack((fg, 1.0 - fg), dim=1).contiguous()
    tgt_h = torch.randint(0, 2, (B // 2 * Q, H, W), device="cuda", generator=g)
    tgt_h[:, :3] = 255
    target = torch.cat((tgt_h, tgt_h), dim=0)

    def run(scale):
        f = feats.view(B * (S + Q), c, h, w).detach().clone().requires_grad_(True)
        cc = ctr.detach().clone().requires_grad_(True)
        loss, _ = A.pemp_head_loss(f, low, cc, B, S, Q, target)
        (loss * scale).backward()
        return float(loss.detach()), f.grad.view(B, S + Q, c, h, w), cc.grad

    loss1, gf1, gc1 = run(1.0)
    assert torch.isfinite(gf1).all() and torch.isfinite(gc1).all() and float(gf1.abs().max()) > 0
    assert torch.equal(gf1[: B // 2], gf1[B // 2:])                                  # (1) duplicate episodes
    loss3, gf3, gc3 = run(3.0)                                                       # (2) linearity
    assert loss1 == loss3
    assert nrel(gf3.cpu(), (3.0 * gf1).cpu()) < 1e-6 and nrel(gc3.cpu(), (3.0 * gc1).cpu()) < 1e-6
    # half batch: same per-episode arithmetic, the mean over valid pixels has the same value, gradients per episode are 2 x
    fh = half.view(B // 2 * (S + Q), c, h, w).detach().clone().requires_grad_(True)
    ch = ctr.detach().clone().requires_grad_(True)
    lh, _ = A.pemp_head_loss(fh, low[: B // 2 * S].contiguous(), ch, B // 2, S, Q, tgt_h)
    lh.backward()
    assert abs(float(lh.detach()) - loss1) < 1e-6 * max(1.0, abs(loss1))
    # (not bit for bit: how an image's tiles are split over CTAs depends on the number of images in the launch, so the
    # prototype gradients of the half batch are summed in another grouping - 1.5e-6 observed)
    assert nrel((2.0 * gf1[: B // 2]).cpu(), fh.grad.view(B // 2, S + Q, c, h, w).cpu()) < 5e-6
    assert nrel(gc1.cpu(), ch.grad.cpu()) < 1e-5
    # (3) a support image whose masks are all zero contributes no gradient to its features
    low0 = low.clone()
    low0[0] = 0.0
    f = feats.view(B * (S + Q), c, h, w).detach().clone().requires_grad_(True)
    cc = ctr.detach().clone().requires_grad_(True)
    loss, _ = A.pemp_head_loss(f, low0, cc, B, S, Q, target)
    loss.backward()
    assert float(f.grad.view(B, S + Q, c, h, w)[0, 0].abs().max()) == 0.0


def test_backward_twice_with_retain_graph():
    from pemp_b200 import autograd as A
    B, S, Q, c, h, w, P = 1, 2, 1, 64, 7, 7, 3
    feats, ctr, fg, bg = _case(B, S, Q, c, h, w, P, seed=6)
    f_cu = feats.cuda().view(B * (S + Q), c, h, w).requires_grad_(True)
    c_cu = ctr.cuda().requires_grad_(True)
    pred = A.pemp_head(f_cu, torch.stack((fg, bg), 1).cuda(), c_cu, B, S, Q)
    loss = pred.square().mean()
    loss.backward(retain_graph=True)
    g1 = f_cu.grad.clone()
    loss.backward()
    assert torch.equal(f_cu.grad, 2 * g1)


@pytest.mark.parametrize("N,Bp,c,h,w,P", [(2, 2, 64, 9, 11, 3), (1, 1, 512, 51, 51, 3), (4, 2, 32, 7, 5, 1), (2, 1, 128, 8, 9, 4),
                                          (3, 3, 512, 13, 13, 1)])
def test_compute_similarity_is_differentiable(N, Bp, c, h, w, P):
    """`compute_similarity` called on its own under autograd (pemp_stage1.py:233-261; baseline.py / panet.py with [B, c]
    prototypes, N a multiple of B as in panet.py:145-149): gradient of the per-prototype maps against float64 autograd over the
    oracle (VERDICT r1 missing #6: this used to raise NotImplementedError)."""
    from pemp_b200 import heads
    g = torch.Generator().manual_seed(c + w + P)
    single = P == 1
    qry = torch.randn(N, c, h, w, generator=g)
    shape = (Bp, c) if single else (Bp, c, P)
    fgp, bgp = torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)
    wgt = torch.randn(N, 2, P, h * w, generator=g)
    q64 = qry.double().reshape(N, c, h * w).requires_grad_(True)
    f64, b64 = fgp.double().requires_grad_(True), bgp.double().requires_grad_(True)
    sim64 = O.cosine_match(q64, f64, b64, 20.0)                                   # [N, 2, P, hw]
    (sim64 * wgt.double()).sum().backward()
    q_cu = qry.cuda().requires_grad_(True)
    f_cu, b_cu = fgp.cuda().requires_grad_(True), bgp.cuda().requires_grad_(True)
    out = heads.compute_similarity(None, f_cu, b_cu, q_cu if single else q_cu.unsqueeze(2))
    assert out.shape == ((N, 2, h, w) if single else (N, 2, P, h, w)) and out.requires_grad
    assert nrel(out.detach().cpu().reshape(N, 2, P, h * w), sim64.detach().float()) < 1e-5
    (out.reshape(N, 2, P, h * w) * wgt.cuda()).sum().backward()
    assert nrel(q_cu.grad.cpu().view(N, c, h * w), q64.grad.float()) < GTOL
    assert nrel(f_cu.grad.cpu(), f64.grad.float()) < GTOL
    assert nrel(b_cu.grad.cpu(), b64.grad.float()) < GTOL


def test_weighted_gap_and_comm_keep_their_gradients():
    """the reference trains THROUGH `Weighted_GAP` (`down_supp` feeds it, pfenet.py:197-198) and through
    `ResNetCM.comm` (nn.Linear weights, backbones.py:208-222).  The drop-ins must hand gradients to those parameters: Weighted_GAP
    through the K1 backward kernel, comm through the stock differentiable expression; both checked against float64 autograd."""
    import types
    from pemp_b200 import heads
    torch.backends.cudnn.allow_tf32 = False             # the 1x1 convolution below must not round to TF32
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(3)
    B, cin, c, h, w = 3, 24, 256, 15, 15
    # Weighted_GAP alone: gradient w.r.t. its input
    xin = torch.randn(B, c, h, w, generator=g)
    m0 = (torch.rand(B, 1, h, w, generator=g) > 0.4).float()
    x_cu = xin.cuda().requires_grad_(True)
    heads.Weighted_GAP(x_cu, m0.cuda()).square().sum().backward()
    x64 = xin.double().requires_grad_(True)
    O.weighted_gap(x64, m0.double()).square().sum().backward()
    assert nrel(x_cu.grad.cpu(), x64.grad.float()) < GTOL
    conv = torch.nn.Conv2d(cin, c, 1, bias=False).cuda()                     # stands for `down_supp`
    x = torch.randn(B, cin, h, w, generator=g).cuda()
    mask = (torch.rand(B, 1, h, w, generator=g) > 0.4).float().cuda()
    out = heads.Weighted_GAP(torch.relu(conv(x)), mask)
    assert out.shape == (B, c, 1, 1) and out.requires_grad
    out.square().sum().backward()
    assert conv.weight.grad is not None and float(conv.weight.grad.abs().max()) > 0
    c64 = torch.nn.Conv2d(cin, c, 1, bias=False).double()
    c64.weight.data.copy_(conv.weight.detach().cpu().double())
    O.weighted_gap(torch.relu(c64(x.cpu().double())), mask.cpu().double()).square().sum().backward()
    assert nrel(conv.weight.grad.cpu(), c64.weight.grad.float()) < GTOL
    # comm: x and the linear layer both receive gradients, equal to the oracle's
    spq, n_out = 3, 2
    lin = torch.nn.Linear(2 * 16, n_out).cuda()
    xc = torch.randn(2 * spq, 16, 9, 9, generator=g).cuda().requires_grad_(True)
    mc = (torch.rand(2 * spq, 1, 17, 17, generator=g) > 0.5).float().cuda()
    feat, pooled = heads.comm(types.SimpleNamespace(spq=spq), xc, mc, lin, stride=2)
    feat.square().sum().backward()
    assert lin.weight.grad is not None and lin.bias.grad is not None and xc.grad is not None
    x64 = xc.detach().cpu().double().requires_grad_(True)
    w64, b64 = lin.weight.detach().cpu().double().requires_grad_(True), lin.bias.detach().cpu().double().requires_grad_(True)
    f64, _ = O.comm_module(x64, mc.cpu().double(), w64, b64, spq, 2)
    f64.square().sum().backward()
    assert nrel(feat.detach().cpu(), f64.detach().float()) < 1e-5
    assert nrel(lin.weight.grad.cpu(), w64.grad.float()) < GTOL and nrel(xc.grad.cpu(), x64.grad.float()) < GTOL
    with torch.no_grad():                                                     # evaluation still runs the fused kernel (K11)
        f_eval, _ = heads.comm(types.SimpleNamespace(spq=spq), xc, mc, lin, stride=2)
    assert not f_eval.requires_grad and nrel(f_eval.cpu(), f64.detach().float()) < 1e-5


def test_inplace_write_between_forward_and_backward_is_detected():
    """the saved tensors go through `save_for_backward`, so autograd's version counters catch an in-place write
    to the encoder output or to `ctr` before backward instead of returning silently wrong gradients."""
    from pemp_b200 import autograd as A
    B, S, Q, c, h, w, P = 1, 2, 1, 64, 7, 7, 3
    feats, ctr, fg, bg = _case(B, S, Q, c, h, w, P, seed=9)
    base = feats.cuda().view(B * (S + Q), c, h, w).requires_grad_(True)
    f_cu = base * 1.0                                   # a non-leaf the test may write in place
    c_cu = ctr.cuda().requires_grad_(True)
    pred = A.pemp_head(f_cu, torch.stack((fg, bg), 1).cuda(), c_cu, B, S, Q)
    f_cu.add_(1.0)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        pred.sum().backward()
