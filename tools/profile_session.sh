#!/bin/bash
# One gpurun session that produces the ncu evidence kept under profiles/ (launch list of the bench command + one
# `--set full` capture per hot kernel).  Every ncu run follows a plain run of the same command (B200_PROFILING.md).
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-extra"
$B > gpurun_out/${TAG}_plain_bench.json 2> gpurun_out/${TAG}_plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv $B > /dev/null 2>&1
prof() {   # prof <op> <kernel regex> <name>
  python tools/prof_op.py --op $1 > gpurun_out/plain_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o gpurun_out/${TAG}_prof_$3 python tools/prof_op.py --op $1 > gpurun_out/ncu_$3.log 2>&1
}
prof K2 mpa_tma_kernel mpa_tma
prof K2s1 mpa_tma_kernel mpa_tma_64img
prof K3 cosine_tma_kernel cosine_tma
prof K6 adjoint_rows2 adjoint_rows2
prof K7 upsample_ce_band2 ce_band2
prof K11 comm_pool comm_pool
prof K9 prep_kmajor prep_kmajor
prof K9 prior_tc_kernel prior_tc
prof K9x3 prior_tc_kernel prior_tc_x3
ls -la gpurun_out/${TAG}_prof_*.ncu-rep
