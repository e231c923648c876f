"""Executed-instruction histogram by opcode over every table of a `--page source --print-source sass,cuda --csv` export.
    python tools/ncu_opcode_hist.py src.csv [warp_tiles]"""
import collections
import csv
import sys


def num(v):
    try:
        return int(v)
    except (ValueError, TypeError):
        return 0


def main(path, units):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Line No"] + [len(rows)]
    seen, hist, samp = set(), collections.Counter(), collections.Counter()
    for a, b in zip(starts[:-1], starts[1:]):
        hdr = rows[a]
        i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        for r in rows[a + 1:b]:
            if len(r) < len(hdr) or r[2] in ("", "-") or r[2] in seen:
                continue
            seen.add(r[2])
            toks = [x for x in r[3].split() if not x.startswith("@")]
            op = toks[0].split(".")[0]
            if op in ("LDS", "STS", "LDG", "STG"):
                op = ".".join(toks[0].split(".")[:1] + [x for x in toks[0].split(".")[1:] if x in ("64", "128")])
            hist[op] += num(r[i_inst])
            samp[op] += num(r[i_samp])
    tot = sum(hist.values())
    print(f"total {tot / 1e6:.1f} M warp instructions, {tot / units:.1f} per unit; samples {sum(samp.values())}")
    for k, v in hist.most_common(40):
        print(f"{k:12s} {v / 1e6:8.2f}M {v / units:7.1f}/unit   samples {samp[k]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
