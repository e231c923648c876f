"""How far are the fp32 backward kernels from float64 autograd of the oracle, and how far is fp32 autograd of the oracle
(= what the reference's training step computes) from the same float64 result?  Prints nrel for both."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
from conftest import nrel
from oracle import restate as O
from pemp_b200 import autograd as A
from test_gpu_train import _case

for (B, S, c, h, w, P) in [(2, 2, 64, 9, 11, 3), (1, 1, 512, 51, 51, 3), (2, 5, 512, 13, 13, 3), (1, 3, 1024, 6, 7, 2)]:
    feats, ctr, fg, bg = _case(B, S, 1, c, h, w, P, seed=c + h)
    g = torch.Generator().manual_seed(7)
    wf, wb = torch.randn(B, c, P, generator=g), torch.randn(B, c, P, generator=g)
    res = {}
    for dt in (torch.float64, torch.float32):
        sup = feats[:, :S].to(dt).reshape(B * S, c, h * w).clone().requires_grad_(True)
        cc = ctr.to(dt).clone().requires_grad_(True)
        of, ob, _ = O.meta_proto_attention(sup, fg.to(dt), bg.to(dt), cc, B, S, P)
        ((of * wf.to(dt)).sum() + (ob * wb.to(dt)).sum()).backward()
        res[dt] = (sup.grad.double(), cc.grad.double())
    f_cu = feats.cuda().requires_grad_(True)
    ctr_cu = ctr.cuda().requires_grad_(True)
    kf, kb = A.meta_proto_attn(f_cu[:, :S], ctr_cu, fg.cuda(), bg.cuda())
    ((kf * wf.cuda()).sum() + (kb * wb.cuda()).sum()).backward()
    d_sup = f_cu.grad[:, :S].reshape(B * S, c, h * w).cpu().double()
    print((B, S, c, h, w, P), "kernel vs f64: d_sup %.2e d_ctr %.2e | torch fp32 vs f64: d_sup %.2e d_ctr %.2e" % (
        nrel(d_sup, res[torch.float64][0]), nrel(ctr_cu.grad.cpu().double(), res[torch.float64][1]),
        nrel(res[torch.float32][0], res[torch.float64][0]), nrel(res[torch.float32][1], res[torch.float64][1])))
