"""Per-source-line instruction / stall / shared-wavefront totals from `ncu --page source --csv --print-source sass,cuda`
(one table per source file; all of them are summarised).
    ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda > src.csv; python tools/ncu_source_lines.py src.csv [N]"""
import collections
import csv
import sys

STALLS = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_math", "stall_lg",
          "stall_membar", "stall_not_selected"]


def num(v):
    try:
        return int(v)
    except (ValueError, TypeError):
        return 0


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    for si, s in enumerate(starts):
        hdr = rows[s]
        end = starts[si + 1] - 2 if si + 1 < len(starts) else len(rows)
        fname = rows[s - 2][1] if s >= 2 and len(rows[s - 2]) > 1 else "?"
        i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        cols = {n: hdr.index(n) for n in STALLS + ["L1 Wavefronts Shared"] if n in hdr}
        per = collections.OrderedDict()
        tot = ts = 0
        for r in rows[s + 1:end]:
            if len(r) < len(hdr) or r[0] == "":
                continue
            d = per.setdefault(r[0], dict(src=r[1], inst=0, samp=0, **{k: 0 for k in cols}))
            d["inst"] += num(r[i_inst])
            d["samp"] += num(r[i_samp])
            for k, c in cols.items():
                d[k] += num(r[c])
            tot += num(r[i_inst])
            ts += num(r[i_samp])
        if ts == 0:
            continue
        print("==", fname, "warp instructions", tot, "samples", ts, {k: sum(d[k] for d in per.values()) for k in cols if k != "L1 Wavefronts Shared"})
        for line, d in sorted(per.items(), key=lambda kv: -kv[1]["samp"])[:top]:
            if d["samp"] * 200 < ts:
                break
            st = " ".join(f"{k[6:9]}={d[k]:5d}" for k in cols if k != "L1 Wavefronts Shared")
            wf = d.get("L1 Wavefronts Shared", 0) / 1e6
            print(f"{line:>4} inst={d['inst'] / 1e6:7.2f}M samp={d['samp']:6d} {st} wf={wf:6.2f}M | {d['src'][:80]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
