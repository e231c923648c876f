"""Per-source-line instruction / stall / shared-wavefront totals from `ncu --page source --csv --print-source sass,cuda`.
    ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda > src.csv; python tools/ncu_source_lines.py src.csv [N]"""
import collections
import csv
import sys


def num(v):
    try:
        return int(v)
    except (ValueError, TypeError):
        return 0


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    hdr = rows[starts[0]]
    end = starts[1] if len(starts) > 1 else len(rows)
    i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    cols = {n: hdr.index(n) for n in ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio",
                                      "stall_math", "stall_not_selected", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"]}
    per = collections.OrderedDict()
    tot = 0
    for r in rows[starts[0] + 1:end]:
        if len(r) < len(hdr):
            continue
        d = per.setdefault(r[0], dict(src=r[1], inst=0, samp=0, **{k: 0 for k in cols}))
        d["inst"] += num(r[i_inst])
        d["samp"] += num(r[i_samp])
        for k, c in cols.items():
            d[k] += num(r[c])
        tot += num(r[i_inst])
    print("total warp instructions", tot, " total samples", sum(d["samp"] for d in per.values()))
    for line, d in sorted(per.items(), key=lambda kv: -kv[1]["samp"])[:top]:
        print(f"{line:>4} inst={d['inst'] / 1e6:7.2f}M samp={d['samp']:6d} bar={d['stall_barrier']:5d} lsb={d['stall_long_sb']:5d} "
              f"ssb={d['stall_short_sb']:5d} wait={d['stall_wait']:5d} mio={d['stall_mio']:4d} math={d['stall_math']:4d} "
              f"wf={d['L1 Wavefronts Shared'] / 1e6:6.2f}/{d['L1 Wavefronts Shared Ideal'] / 1e6:6.2f} | {d['src'][:80]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
