#!/bin/bash
# The round-end evidence run: full GPU test suite, smoke, every bench workload, kernel / training benches, the ncu capture of the
# tensor-path backward kernel and the launch list of the bench command.  Logs land in gpurun_out/ (merged back).
TAG=${1:-r02final}
tools/gpu_session.sh $TAG
TB_B=16 timeout 300 python tools/train_bench.py > gpurun_out/${TAG}_train_bench.txt 2>&1
TB_B=64 timeout 300 python tools/train_bench.py >> gpurun_out/${TAG}_train_bench.txt 2>&1
python tools/prof_op.py --op K2b --B 16 > gpurun_out/plain_bwd.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:mpa_bwd_mma_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_mpa_bwd_mma python tools/prof_op.py --op K2b --B 16 > gpurun_out/ncu_bwd.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_bwd.csv python tools/prof_op.py --op K2b --B 16 > /dev/null 2>&1
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-extra"
$B > gpurun_out/${TAG}_plain_bench.json 2> gpurun_out/${TAG}_plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv $B > /dev/null 2>&1
tail -3 gpurun_out/${TAG}_train_bench.txt
