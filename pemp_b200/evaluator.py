"""Device-resident evaluator glue for PEMP (`entry/pemp_stage2.py:58-65`, `core/base_trainer.py:76-85`).

The reference runs one episode per call and synchronises with the host twice per episode (`.item()`,
`.cpu().numpy()`).  Here a batch of episodes goes through

    K0 nearest masks -> stage-1 head (K2, K3, K4) -> prior mask           (qry_prior of pemp_stage2.py:133-138)
                     -> [stage-2 encoder: stock PyTorch, not part of this library]
                     -> stage-2 head (K2, K3, K4) -> argmax mask -> K10 confusion counts

without leaving the GPU; only the (C+1) x 3 count table is read back, once per round.
"""
import torch

from . import ops


class KernelTimer:
    """Optional CUDA-event bracket around one kernel family (bench.py uses it for the roofline of K2)."""

    def __init__(self):
        self.pairs = []

    def bracket(self, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        self.pairs.append((a, b))
        return out

    def mean_ms(self):
        if not self.pairs:
            return None
        return sum(a.elapsed_time(b) for a, b in self.pairs) / len(self.pairs)

    def count(self):
        return len(self.pairs)


class GraphedStep:
    """`PEMPStage2Pipeline.step` captured once in a CUDA graph ("CUDA-graph the head").

    The tensors given to `capture` are the graph's static buffers: write the next batch of features / masks / labels
    into them (e.g. let the encoder produce into `out=`-style views of them, or `copy_`) and call `replay()`; `stat`
    keeps accumulating, `prior` and `mask` are overwritten by every replay.  TMA descriptors, workspace addresses and
    launch shapes are frozen in the graph, so the shapes are fixed for the life of the object."""

    def __init__(self, pipe, args):
        self.args = args
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        stat = args[7]
        keep = stat.clone()
        with torch.cuda.stream(side):                # warm-up outside the capture (function attributes, lazy module load)
            pipe.step(*args)
        cur.wait_stream(side)
        n0 = ops.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) may touch CUDA
            self.prior, self.mask = pipe.step(*args)
        self.launches = ops.launch_count() - n0      # kernels recorded in the graph (capturing does not run them)
        ops._count(-self.launches)
        stat.copy_(keep)                             # the warm-up step must not be counted

    def replay(self):
        self.graph.replay()
        ops._count(self.launches)
        return self.prior, self.mask


class PEMPStage2Pipeline:
    def __init__(self, ctr1, ctr2, classes=20, dist_scalar=20):
        self.ctr1, self.ctr2 = ctr1, ctr2
        self.classes = classes
        self.dist_scalar = dist_scalar

    def _head(self, feats, low, ctr, B, S, Q, out_shape, timer, hist=None):
        _, c, h, w = feats.shape
        f5 = feats.view(B, S + Q, c, h, w)
        sup, qry = f5[:, :S], f5[:, S:]               # read in place through the episode stride
        run = (lambda: ops.meta_proto_attn(sup, ctr, low[:, 0], low[:, 1], B, S, want_adaptive=False))
        fgp, bgp, _ = timer.bracket(run) if timer is not None else run()
        pred = ops.cosine_match(qry, fgp, bgp, self.dist_scalar)["pred"].view(B * Q, 2, h, w)
        if hist is not None:                          # (qry_msk, cls, stat): mask and FewShotMetric counts in one launch
            ref, cls, stat = hist
            return ops.upsample_argmax_hist(pred, out_shape, ref.view(B * Q, *out_shape),
                                            cls.repeat_interleave(Q) if Q > 1 else cls, stat)
        return ops.upsample_argmax(pred, out_shape, want_mask8=True)["mask8"]

    def stage1_prior(self, feats1, low, B, S, Q, HW, timer=None):
        """-> qry_prior [BQ, H, W] uint8 (the reference builds an int64 [BQ,1,H,W] and `.float()`s it)."""
        return self._head(feats1, low, self.ctr1, B, S, Q, HW, timer)

    def stage2_mask(self, feats2, low, B, S, Q, out_shape, timer=None, hist=None):
        return self._head(feats2, low, self.ctr2, B, S, Q, out_shape, timer, hist)

    def step(self, sup_feats1, qry_feats1, sup_feats2, qry_feats2, sup_mask, qry_msk, cls, stat, timer=None):
        """One batch of episodes with support and query features stored separately:
        sup_feats* [B, S, c, h, w], qry_feats* [B, Q, c, h, w], sup_mask [B, S, 2, H, W] float32 as the loader emits it -
        or the uint8 label map [B, S, H, W] (1 object / 0 background / 255 boundary) it was expanded from, an eighth of the
        bytes -, qry_msk [B, Q, H', W'] uint8, cls [B] int64; `stat` [(C+1), 3] int64 is accumulated.
        Returns (prior [BQ,H,W] uint8, mask [BQ,H',W'] uint8)."""
        B, S, c, h, w = sup_feats1.shape
        Q = qry_feats1.shape[1]
        H, W = sup_mask.shape[-2:]
        if sup_mask.dtype == torch.uint8:
            low = ops.mask_nearest_labels(sup_mask.view(B * S, H, W), h, w).view(B * S, 2, h * w)
        else:
            low = ops.mask_nearest(sup_mask.view(B * S, 2, H, W), h, w).view(B * S, 2, h * w)
        out = []
        for sup, qry, ctr, shape in ((sup_feats1, qry_feats1, self.ctr1, (H, W)),
                                     (sup_feats2, qry_feats2, self.ctr2, tuple(qry_msk.shape[-2:]))):
            run = (lambda sup=sup, ctr=ctr: ops.meta_proto_attn(sup, ctr, low[:, 0], low[:, 1], B, S,
                                                                want_adaptive=False))
            fgp, bgp, _ = timer.bracket(run) if timer is not None else run()
            pred = ops.cosine_match(qry, fgp, bgp, self.dist_scalar)["pred"].view(B * Q, 2, h, w)
            if len(out) == 0:
                out.append(ops.upsample_argmax(pred, shape, want_mask8=True)["mask8"])       # stage 1: the prior mask
            else:                                                                             # stage 2: mask + counts, one launch
                out.append(ops.upsample_argmax_hist(pred, shape, qry_msk.view(B * Q, *shape),
                                                    cls.repeat_interleave(Q) if Q > 1 else cls, stat))
        return out[0], out[1]

    def capture(self, sup_feats1, qry_feats1, sup_feats2, qry_feats2, sup_mask, qry_msk, cls, stat):
        """-> GraphedStep replaying `step` on these (static) tensors with one graph launch instead of nine."""
        return GraphedStep(self, (sup_feats1, qry_feats1, sup_feats2, qry_feats2, sup_mask, qry_msk, cls, stat))
