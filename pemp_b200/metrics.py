"""`FewShotMetric` with the reference's interface (`core/metrics.py:4-35`), counted on the GPU.

`update` accepts what the reference's `test_step` hands over (NumPy arrays) as well as CUDA tensors that never
left the device; counts are exact int64 accumulated by `pemp_iou_hist`.  `stat` and `mIoU` return float64
NumPy values computed exactly as the reference does (K11 stays on the host: <= 243 numbers).
"""
import numpy as np
import torch

from . import ops


class FewShotMetric(object):
    def __init__(self, classes, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("pemp_b200.metrics.FewShotMetric counts on a CUDA device; none is available")
        self.classes = classes
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.stat_dev = torch.zeros(classes + 1, 3, dtype=torch.int64, device=self.device)

    def _as_u8(self, x):
        if isinstance(x, torch.Tensor):
            t = x if x.dtype == torch.uint8 else x.to(torch.uint8)       # np.asarray(pred, np.uint8), metrics.py:10
            return t.to(self.device, non_blocking=True)
        return torch.from_numpy(np.ascontiguousarray(np.asarray(x, np.uint8))).to(self.device, non_blocking=True)

    def update(self, pred, ref, cls, verbose=0):
        pred = self._as_u8(pred)
        ref = self._as_u8(ref)
        n = pred.shape[0]
        ref = ref.reshape(n, -1)
        pred = pred.reshape(n, -1)
        if isinstance(cls, torch.Tensor):
            cls_t = cls.to(self.device, dtype=torch.int64).reshape(-1)
        else:
            cls_t = torch.as_tensor([int(c) for c in cls], dtype=torch.int64, device=self.device)
        if verbose:
            before = self.stat_dev.clone()
        ops.iou_hist(pred, ref, cls_t, self.stat_dev)
        if verbose:                                   # the reference prints the per-class IoU of this call
            d = (self.stat_dev - before).cpu().numpy().astype(np.float64)
            for row in d[d.sum(axis=1) > 0]:
                print(row[0] / row.sum())

    @property
    def stat(self):
        """[(classes+1), 3] float64 (tp, fp, fn), like `self.stat` of the reference."""
        return self.stat_dev.cpu().numpy().astype(np.float64)

    def all_reduce(self):
        """Sum the counts over all ranks (episodes are sharded over GPUs; NCCL all-reduce of (C+1)*3 int64)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.stat_dev, op=dist.ReduceOp.SUM)
        return self

    def mIoU(self, labels, binary=False):
        return miou_from_stat(self.stat, labels, binary)


def miou_from_stat(stat, labels, binary=False):
    """`FewShotMetric.mIoU` (core/metrics.py:25-35) on a float64 stat array."""
    stat = np.asarray(stat, np.float64)
    if binary:
        stat = np.c_[stat[0], stat[1:].sum(axis=0)].T
    else:
        stat = stat[labels]
    tp, fp, fn = stat.T
    mIoU_class = tp / (tp + fp + fn)
    return mIoU_class, mIoU_class.mean()
