#!/bin/bash
# Round-end style validation on one B200: GPU tests, smoke, both bench arms, and the final fused kernel under ncu.
#   gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh [ncu]'
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench exit $?" >> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref exit $?" >> gpurun_out/r02_bench_ref.err
if [ "$1" = "ncu" ]; then
python tools/ncu_targets.py fused > gpurun_out/plain_fused.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_fused.csv \
      python tools/ncu_targets.py fused > gpurun_out/ncu_l_fused.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:blend_skin_fused -s 1 -c 1 -o gpurun_out/r02_fused \
      python tools/ncu_targets.py fused > gpurun_out/ncu_f_fused.log 2>&1
fi
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; tail -c 300 gpurun_out/r02_bench.err; tail -c 300 gpurun_out/r02_bench_ref.err
