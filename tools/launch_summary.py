"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into per-kernel shares.
    python tools/launch_summary.py gpurun_out/launches.csv [bench.json] > profiles/rNN_launches_summary.txt"""
import collections
import csv
import json
import sys

OURS = ("mpa_", "cosine", "upsample", "mask_nearest", "iou_hist", "pool_", "adjoint", "comm_", "prior", "prep_", "eps_flag")


def main(path, bench=None):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= iv or not any(t in r[ik] for t in OURS):
            continue
        k = r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print("ncu --metrics gpu__time_duration.sum --clock-control none -c 400 : python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline "
          "--no-graph --no-extra")
    print("(first 400 launches of the command: warm-up, timed and sustained steps of 64 five-shot episodes; per-launch times are "
          "cold-cache and serialised - compare SHARES; full list beside this file)")
    print("kernel, launches, total_us, share of our kernels")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:45s} {a[0]:4d} {a[1] / 1e3:9.1f} {100 * a[1] / tot:5.1f}%")
    k2 = sum(a[1] for k, a in agg.items() if k.startswith("mpa_"))
    line = f"K2 kernels (mpa_tma + finalize) = {100 * k2 / tot:.1f} % of our kernels' GPU time"
    if bench:
        b = json.load(open(bench))
        line += f"; bench.py (same build, same box) reported roofline.share_of_step = {b['roofline']['share_of_step']:.3f}"
    print(line)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
