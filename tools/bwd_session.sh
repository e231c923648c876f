#!/bin/bash
# One gpurun session for the backward kernels: parity tests of the training path, the kernel timings, one ncu capture.
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_train.py -m gpu -x -q -rP -k "${KEXPR:-meta_proto_attn_backward or cosine_match_backward or deterministic or head_loss or bench_size or retain or similarity}" 2>&1 | tail -${TAILN:-12} | tee gpurun_out/${TAG}_train_tests.txt
TB_B=16 timeout 200 python tools/train_bench.py 2>&1 | head -3 | tee gpurun_out/${TAG}_train_bench16.txt
TB_B=64 timeout 200 python tools/train_bench.py 2>&1 | head -3 | tee gpurun_out/${TAG}_train_bench64.txt
if [ "${NCU:-1}" = "1" ]; then
  python tools/prof_op.py --op ${NCU_OP:-K2b} --B 16 > gpurun_out/plain_bwd.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:${NCU_K:-mpa_bwd_mma_kernel} -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_bwd python tools/prof_op.py --op ${NCU_OP:-K2b} --B 16 > gpurun_out/ncu_bwd.log 2>&1
  tail -2 gpurun_out/ncu_bwd.log
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_bwd.csv python tools/prof_op.py --op ${NCU_OP:-K2b} --B 16 > /dev/null 2>&1
fi
