import sys, json, torch, numpy as np
sys.path.insert(0, "/root/repo")
from oracle import ref_import as R
from pemp_b200 import dropin
torch.backends.cudnn.deterministic = True
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
S, H = 1, 225
net = R.full_model("pemp_stage1", seed=5).cuda()
g = torch.Generator().manual_seed(21)
sup_img = torch.randn(1, S, 3, H, H, generator=g).cuda(); qry_img = torch.randn(1, 1, 3, H, H, generator=g).cuda()
fg = torch.zeros(1, S, H, H); fg[:, :, 60:170, 50:180] = 1.0
sup_mask = torch.stack((fg, 1 - fg), dim=2).cuda()
target = torch.zeros(1, H, H, dtype=torch.int64); target[:, 80:150, 70:200] = 1; target = target.cuda()
def step(n, dt=torch.float32):
    n.zero_grad(set_to_none=True)
    out = n(sup_img.to(dt), sup_mask.to(dt), qry_img.to(dt), (H, H))
    loss = torch.nn.functional.cross_entropy(out, target, ignore_index=255); loss.backward()
    return float(loss), {k: p.grad.detach().double().cpu() for k, p in n.named_parameters() if p.grad is not None}
l32, g32 = step(net)
dropin.patch(); lo, go = step(net); dropin.unpatch()
import copy
net64 = copy.deepcopy(net).double()
l64, g64 = step(net64, torch.float64)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
rows = sorted(((rel(go[k], g32[k]), rel(go[k], g64[k]), rel(g32[k], g64[k]), k) for k in g32), reverse=True)[:6]
print(l32, lo, l64)
for r in rows: print("ours-vs-ref32 %.2e  ours-vs-64 %.2e  ref32-vs-64 %.2e  %s" % r)
print("ctr:", rel(go["ctr"], g32["ctr"]), rel(go["ctr"], g64["ctr"]), rel(g32["ctr"], g64["ctr"]))
