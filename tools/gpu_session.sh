#!/bin/bash
# One gpurun session: GPU tests, then the bench lines of every workload.  Logs land in gpurun_out/ (merged back).
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  python -m pytest tests -m gpu -q --timeout 1200 -rfEP ${PYTEST_ARGS:-} > gpurun_out/${TAG}_tests.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
  tail -5 gpurun_out/${TAG}_tests.log
fi
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/${TAG}_bench_stage2_5shot.json 2> gpurun_out/${TAG}_bench_stage2_5shot.err; echo "bench rc=$?"
for w in ${WORKLOADS:-stage1_1shot baseline_1shot panet_5shot_coco pfenet_5shot}; do
  python bench.py --workload $w > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "bench $w rc=$?"
done
python bench.py --workload pfenet_5shot --prior-precision bf16 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_pfenet_bf16.json 2> gpurun_out/${TAG}_bench_pfenet_bf16.err
python bench.py --impl reference --steps 4 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
python tools/kernel_bench.py > gpurun_out/${TAG}_kernel_bench.txt 2>&1
tail -n 3 gpurun_out/${TAG}_bench_*.err | tail -40
