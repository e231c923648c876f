"""Condense an Nsight Compute report into the few numbers DESIGN.md / bench.py quote.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"))
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k} = {d[k]} {u.get(k, '')}")
        stalls = {k: float(v) for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v}
        for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:6]:
            print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]} = {v:.3f} warps per issue-active cycle")
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        if rd and wr:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tot = float(rd) * scale.get(u["dram__bytes_read.sum"], 1) + float(wr) * scale.get(u["dram__bytes_write.sum"], 1)
            print(f"  traffic (dram read + write) = {tot / 1e6:.2f} MB per launch")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
