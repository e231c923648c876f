"""Print the numbers DESIGN.md / README.md quote from a set of bench JSON lines:  python tools/bench_summary.py gpurun_out/r02m_bench_*.json"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.load(open(path))
    except Exception as e:          # noqa: BLE001
        print(path, "ERR", e)
        continue
    if d.get("impl") == "reference":
        print(f"{path}: reference arm {d['value']:.2f} eps/s ({d['cpu_baseline']['kind']}, {d['cpu_baseline']['cores']} cores)")
        continue
    r, er = d["roofline"], d["episode_roofline"]
    print(f"{path}\n  N={d['n_gpus']} value {d['value']:.0f} eps/s  ms/step {d['ms_per_step']:.4f}  sustained {d['sustained']['value']:.0f}"
          f"  graphed {d['graphed'] and round(d['graphed']['value'])}  launches {d['gpu_launches']}  clocks {d['clocks'] and (d['clocks']['sm_mhz'], d['clocks']['reasons'])}")
    print(f"  dominant: {r['kernel'][:40]} frac {r['frac']:.3f} ({r['achieved']:.0f} {r['unit']}) launch_ms {r['launch_ms']:.4f} share {r['share_of_step']:.2f}"
          f"  exec-frac {r.get('tensor_pipe_frac_executed')}")
    print(f"  whole step frac {er.get('frac', er.get('frac_of_bf16_peak')):.3f}   e2e {d['e2e'] and round(d['e2e']['value'])} eps/s "
          f"({d['e2e'] and round(d['e2e']['h2d_GBps_aggregate'], 1)} GB/s aggregate)   cpu {d['cpu_baseline'] and (round(d['cpu_baseline']['value'], 2), d['cpu_baseline']['kind'])}")
    print(f"  parity {d['parity']}")
    if d.get("allreduce"):
        print(f"  allreduce {d['allreduce']['ms_on_rank0']:.4f} ms")
    for x in d.get("roofline_extra") or []:
        print(f"    extra {x['kernel'][:52]:52s} {x['ms']:.4f} ms  {x['achieved']:.0f} {x['unit']}  frac {x['frac']}  exec {x.get('tensor_pipe_frac_executed')}")
