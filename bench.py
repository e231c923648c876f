#!/usr/bin/env python
"""Benchmark of the prototype-matching head (BASELINE.json metric: episodes/s of the prototype head).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload stage2_5shot]

One "step" = one batch of synthetic episodes through the whole head path of the workload:
  stage2_5shot (headline, north_star target): PEMP Stage-2 ResNet-50 5-shot head = K0 nearest masks, stage-1 head
      (K2 meta-prototype attention, K3 cosine matching, K4 up-sample+argmax -> prior), stage-2 head (K2, K3, K4) and
      the K10 IoU counts; c=512, 51x51 features, 401x401 masks (entry/pemp_stage2.py:58-65).
  stage1_1shot: PEMP Stage-1 1-shot head (K0, K2, K3, K4, K10).
Episodes are independent, so N ranks each process `--batch` episodes per step (weak scaling); the only collective is
the all-reduce of the (C+1)x3 int64 count table, once per step.

The line printed by rank 0 follows the driver contract: `value` = whole-job episodes/s with inputs resident in HBM,
`e2e` = the same through the host-facing call with pinned HOST buffers (H2D of every input and D2H of the count table
inside the timed region), `roofline` for the dominant kernel (K2) timed with CUDA events inside the timed region,
`cpu_baseline` = the oracle port of the reference timed on this box's cores (N=1 only).
`--impl reference` times that CPU path alone on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

_JSON_OUT = sys.stdout

WORKLOADS = {
    "stage2_5shot": dict(shot=5, stages=2, desc="PEMP Stage-2 ResNet-50 5-shot prototype head (stage-1 head -> prior -> stage-2 head -> IoU)"),
    "stage1_1shot": dict(shot=1, stages=1, desc="PEMP Stage-1 ResNet-50 1-shot prototype head"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="stage2_5shot", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=64, help="episodes per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay of the step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def window(self, t0, t1):
        return [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 7]

    def stop(self, t0, t1, extra=None):
        """Samples taken inside [t0, t1] (the timed region); `extra` = (t0, t1) of an untimed load loop used when the
        timed region was shorter than the sampling period."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows, where = self.window(t0, t1), "timed region"
        if not rows and extra is not None:
            rows, where = self.window(*extra), "untimed repeat of the timed loop (timed region shorter than the sampling period)"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "sampled_during": where, "reasons": reasons}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_port_rate(spec, stages, budget_s, threads):
    """Oracle port of the reference head (oracle/restate.py) on the host cores: episodes/s at the reference's own
    test batch size 1 (`data_kits/datasets.py:23`).  Bounded sample: 1 warm-up episode + as many as fit `budget_s`."""
    from oracle import restate as O
    from pemp_b200 import episodes as E
    torch.set_num_threads(threads)
    ctr1, ctr2 = E.make_ctr(spec, 1), E.make_ctr(spec, 2)

    def one(i):
        b = E.make_batch(spec, [i])
        t0 = time.perf_counter()
        with torch.no_grad():
            if stages == 2:
                O.stage2_episode_batch(b["feats1"], b["feats2"], b["sup_mask"], ctr1, ctr2, 1, spec.shot, spec.query,
                                       b["qry_msk"].numpy(), b["cls"].numpy(), spec.classes)
            else:
                s1 = O.pemp_head(b["feats1"], b["sup_mask"], ctr1, 1, spec.shot, spec.query)
                O.few_shot_stat(O.argmax2(s1["logits"]).numpy(), b["qry_msk"].numpy(), b["cls"].numpy(), spec.classes)
        return time.perf_counter() - t0

    one(0)
    times, i, t_start = [], 1, time.perf_counter()
    while (time.perf_counter() - t_start < budget_s and len(times) < 64) or len(times) < 2:
        times.append(one(i))
        i += 1
    times.sort()
    med = times[len(times) // 2]
    return 1.0 / med, len(times)


def make_spec(args):
    from pemp_b200 import episodes as E
    w = WORKLOADS[args.workload]
    return E.EpisodeSpec(shot=w["shot"], stages=w["stages"]), w


def config_of(args, spec, w, world):
    return {"workload": f"{args.workload}: {w['desc']}", "shot": spec.shot, "query": spec.query, "channels": spec.channels,
            "feature_hw": [spec.h, spec.w], "image_hw": [spec.H, spec.W], "protos": spec.protos, "classes": spec.classes,
            "episodes_per_gpu_per_step": args.batch, "global_episodes_per_step": args.batch * world,
            "parallelism": f"episode-sharded dp{world}", "l2_policy": "inputs larger than L2 (>= 1 GB per step)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    spec, w = make_spec(args)
    threads = os.cpu_count() or 1
    from oracle import restate as O  # noqa: F401  (the reference arm is the one other place that may run the oracle)
    from pemp_b200 import episodes as E
    torch.set_num_threads(threads)
    ctr1, ctr2 = E.make_ctr(spec, 1), E.make_ctr(spec, 2)

    def step(i):
        b = E.make_batch(spec, [i])
        with torch.no_grad():
            if w["stages"] == 2:
                O.stage2_episode_batch(b["feats1"], b["feats2"], b["sup_mask"], ctr1, ctr2, 1, spec.shot, spec.query,
                                       b["qry_msk"].numpy(), b["cls"].numpy(), spec.classes)
            else:
                s1 = O.pemp_head(b["feats1"], b["sup_mask"], ctr1, 1, spec.shot, spec.query)
                O.few_shot_stat(O.argmax2(s1["logits"]).numpy(), b["qry_msk"].numpy(), b["cls"].numpy(), spec.classes)

    steps = max(1, min(args.steps, 12))
    for i in range(min(args.warmup, 2)):
        step(i)
    batches = [E.make_batch(spec, [100 + i]) for i in range(steps)]
    t0 = time.perf_counter()
    for b in batches:
        with torch.no_grad():
            if w["stages"] == 2:
                O.stage2_episode_batch(b["feats1"], b["feats2"], b["sup_mask"], ctr1, ctr2, 1, spec.shot, spec.query,
                                       b["qry_msk"].numpy(), b["cls"].numpy(), spec.classes)
            else:
                s1 = O.pemp_head(b["feats1"], b["sup_mask"], ctr1, 1, spec.shot, spec.query)
                O.few_shot_stat(O.argmax2(s1["logits"]).numpy(), b["qry_msk"].numpy(), b["cls"].numpy(), spec.classes)
    dt = time.perf_counter() - t0
    value = steps / dt
    sample = f"{steps} steps of 1 episode each (the reference's test batch size), inputs resident in host memory"
    cfg = config_of(args, spec, w, 1)
    cfg["episodes_per_gpu_per_step"] = cfg["global_episodes_per_step"] = 1
    cfg["parallelism"] = "single process, CPU"
    print(json.dumps({
        "impl": "reference", "metric": "episodes/sec of prototype head", "value": value, "unit": "episodes/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "episodes/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "Python reference cannot travel to the GPU box: this is oracle/restate.py (bit-exact CPU restatement of the "
                "reference head, pinned by tests/test_oracle_pinned.py) on torch CPU with all host threads"}),
          file=_JSON_OUT, flush=True)


def run_ours(args):
    from pemp_b200 import _cabi, dist as pdist, episodes as E, ops
    from pemp_b200.evaluator import KernelTimer, PEMPStage2Pipeline

    rank, local_rank, world = pdist.init()
    sampler = ClockSampler(local_rank) if rank == 0 else None      # started early: nvidia-smi needs ~0.5 s to emit
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _cabi.check(_cabi.lib().pemp_check_device(), "pemp_check_device")
    spec, w = make_spec(args)
    B, S, Q, c, h, wd = args.batch, spec.shot, spec.query, spec.channels, spec.h, spec.w
    stages = w["stages"]

    batch = E.device_batch(spec, B, dev, seed=E.REFERENCE_SEED + rank)
    ctr1, ctr2 = E.make_ctr(spec, 1).to(dev), E.make_ctr(spec, 2).to(dev)
    pipe = PEMPStage2Pipeline(ctr1, ctr2, spec.classes)
    stat = torch.zeros(spec.classes + 1, 3, dtype=torch.int64, device=dev)
    f1 = batch["feats1"].view(B, S + Q, c, h, wd)
    f2 = batch["feats2"].view(B, S + Q, c, h, wd) if stages == 2 else None

    def step(fa, fb, sup_mask, qry_msk, cls, st, timer=None, Bn=B):
        if stages == 2:
            pipe.step(fa[:, :S], fa[:, S:], fb[:, :S], fb[:, S:], sup_mask, qry_msk, cls, st, timer)
        else:
            H, W = sup_mask.shape[-2:]
            low = ops.mask_nearest(sup_mask.view(Bn * S, 2, H, W), h, wd).view(Bn * S, 2, h * wd)
            pipe.stage2_mask(fa.view(Bn * (S + Q), c, h, wd), low, Bn, S, Q, tuple(qry_msk.shape[-2:]), timer,
                             hist=(qry_msk, cls, st))
        pdist.all_reduce_stat(st)

    # ---------------- value: device-resident inputs ------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step(f1, f2, batch["sup_mask"], batch["qry_msk"], batch["cls"], stat)
    torch.cuda.synchronize()
    pdist.barrier()
    timer = KernelTimer()
    launches0 = ops.launch_count()
    stat.zero_()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(f1, f2, batch["sup_mask"], batch["qry_msk"], batch["cls"], stat, timer)
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    launches = ops.launch_count() - launches0
    pdist.barrier()
    ms_total = pdist.max_over_ranks(ev0.elapsed_time(ev1), dev)
    extra = None
    if t_wall1 - t_wall0 < 0.25:          # keep the GPU under the same load long enough for a few clock samples
        scratch = torch.zeros_like(stat)
        e0 = time.time()
        while time.time() - e0 < 0.4:
            for _ in range(10):
                step(f1, f2, batch["sup_mask"], batch["qry_msk"], batch["cls"], scratch)
            torch.cuda.synchronize()
        extra = (e0, time.time())
        pdist.barrier()
    clocks = sampler.stop(t_wall0, t_wall1, extra) if sampler else None
    ms_per_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total / 1e3)

    # roofline of the dominant kernel (K2): algorithmic bytes per launch / mean launch duration
    peaks, how = measured_peaks()
    k2_ms = timer.mean_ms()
    k2_bytes = B * (S * (c * h * wd + 2 * h * wd) * 4 + 2 * c * 2 * spec.protos * 4)
    achieved = k2_bytes / (k2_ms / 1e3) / 1e9
    traffic = None      # dram read+write bytes per K2 launch from the committed ncu --set full capture of this shape
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"mpa_tma_kernel:B{B}:S{S}:c{c}:hw{h * wd}")
    roofline = {"kernel": "mpa_tma_kernel (+ prepare, finalize; K2 meta_proto_attn)", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": how,
                "algorithmic_bytes_per_launch": k2_bytes, "launch_ms": k2_ms, "launches_timed": timer.count(),
                "share_of_step": k2_ms * (stages if stages == 2 else 1) / ms_per_step}

    # ---------------- the same step replayed from a CUDA graph (PEMPStage2Pipeline.capture) ---------------
    graphed = None
    if stages == 2 and not args.no_graph:
        gstat = torch.zeros_like(stat)
        g = pipe.capture(f1[:, :S], f1[:, S:], f2[:, :S], f2[:, S:], batch["sup_mask"], batch["qry_msk"], batch["cls"], gstat)
        for _ in range(3):
            g.replay()
            pdist.all_reduce_stat(gstat)
        torch.cuda.synchronize()
        pdist.barrier()
        gstat.zero_()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            g.replay()
            pdist.all_reduce_stat(gstat)
        g1.record()
        torch.cuda.synchronize()
        pdist.barrier()
        ms_g = pdist.max_over_ranks(g0.elapsed_time(g1), dev)
        graphed = {"value": B * world * args.steps / (ms_g / 1e3), "unit": "episodes/s", "ms_per_step": ms_g / args.steps,
                   "kernels_per_graph": g.launches, "same_counts_as_eager": bool(torch.equal(gstat, stat))}
        del g

    # ---------------- e2e: pinned host inputs, H2D + D2H inside the timed region -------------------------
    e2e = None
    if not args.no_e2e:
        host = {k: batch[k].cpu().pin_memory() for k in (["feats1", "feats2"] if stages == 2 else ["feats1"]) + ["sup_mask", "qry_msk", "cls"]}
        chunk = 8 if B % 8 == 0 else B
        nchunks = B // chunk
        copy_stream = torch.cuda.Stream(device=dev)
        rows = {k: (chunk * (S + Q) if k.startswith("feats") else chunk) for k in host}
        bufs = [{k: torch.empty((rows[k],) + tuple(v.shape[1:]), dtype=v.dtype, device=dev) for k, v in host.items()}
                for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        host_stat = torch.zeros(spec.classes + 1, 3, dtype=torch.int64).pin_memory()
        h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

        def e2e_step():
            st = torch.zeros(spec.classes + 1, 3, dtype=torch.int64, device=dev)
            cur = torch.cuda.current_stream()
            for ci in range(nchunks):
                b = bufs[ci % 2]
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[ci % 2])
                    for k, v in host.items():
                        n = b[k].shape[0]
                        b[k].copy_(v[ci * n:(ci + 1) * n], non_blocking=True)
                    ready[ci % 2].record(copy_stream)
                cur.wait_event(ready[ci % 2])
                fa = b["feats1"].view(chunk, S + Q, c, h, wd)
                fb = b["feats2"].view(chunk, S + Q, c, h, wd) if stages == 2 else None
                if stages == 2:
                    pipe.step(fa[:, :S], fa[:, S:], fb[:, :S], fb[:, S:], b["sup_mask"], b["qry_msk"], b["cls"], st)
                else:
                    step(fa, None, b["sup_mask"], b["qry_msk"], b["cls"], st, None, chunk)
                free[ci % 2].record(cur)
            pdist.all_reduce_stat(st)
            host_stat.copy_(st, non_blocking=True)
            cur.synchronize()
            return host_stat

        for f in free:
            f.record(torch.cuda.current_stream())
        e2e_step()
        torch.cuda.synchronize()
        pdist.barrier()
        a, bq = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.e2e_steps):
            e2e_step()
        bq.record()
        torch.cuda.synchronize()
        ms_e2e = pdist.max_over_ranks(a.elapsed_time(bq), dev)
        e2e = {"value": B * world * args.e2e_steps / (ms_e2e / 1e3), "unit": "episodes/s", "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": host_stat.numel() * 8, "steps": args.e2e_steps, "ms_per_step": ms_e2e / args.e2e_steps,
               "path": "pinned host features/masks -> double-buffered cudaMemcpyAsync (8-episode chunks) -> head kernels -> count table D2H"}
        del host, bufs

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, n = cpu_port_rate(spec, stages, args.cpu_seconds, threads)
        cpu = {"value": rate, "unit": "episodes/s", "cores": threads, "kind": "port",
               "sample": f"median of {n} single-episode calls (reference test batch size 1) of the same workload, after 1 warm-up"}

    if rank == 0:
        out = {"metric": "episodes/sec of prototype head", "value": value, "unit": "episodes/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": config_of(args, spec, w, world), "clocks": clocks, "e2e": e2e, "graphed": graphed,
               "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
               "episode_roofline": {"algorithmic_MB_per_episode": (stages * (k2_bytes / B + Q * c * h * wd * 4 + 2 * h * wd * 4 + spec.H * spec.W)
                                                                  + 2 * spec.out_h * spec.out_w) / 1e6}}
        er = out["episode_roofline"]
        er["episodes_per_s_at_hbm_peak"] = peaks["hbm_gbs"] * 1e3 / er["algorithmic_MB_per_episode"]
        er["frac"] = (value / world) / er["episodes_per_s_at_hbm_peak"]
        print(json.dumps(out), file=_JSON_OUT, flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly one JSON line.  Native libraries write there too (NCCL prints its "NCCL version ..." banner on
    # stdout at every debug level the image or the caller may set), so file descriptor 1 is pointed at stderr for the whole
    # run and the JSON line goes out through a private copy of the original stdout.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
