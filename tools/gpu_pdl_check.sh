# One-GPU check of a build: forward / fitting-step A/B of the `pdl` option, GPU tests, ncu of the pose and fitting-step kernels.
set -x
python tools/pdl_ab.py > gpurun_out/r02_pdl_ab.txt 2>&1
python tools/fit_ab.py 1024 pdl >> gpurun_out/r02_pdl_ab.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python tools/ncu_targets.py fused > gpurun_out/plain_fused.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:pose_forward -s 1 -c 1 -o gpurun_out/r02_pose_fwd \
      python tools/ncu_targets.py fused > gpurun_out/ncu_f_pose.log 2>&1
python tools/ncu_targets.py fit > gpurun_out/plain_fit.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_fit.csv \
      python tools/ncu_targets.py fit > gpurun_out/ncu_l_fit.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'dA_seg|skin_fit_l2|pose_|reduce_splits' -s 5 -c 5 -o gpurun_out/r02_fit \
      python tools/ncu_targets.py fit > gpurun_out/ncu_f_fit.log 2>&1
cat gpurun_out/r02_pdl_ab.txt; tail -5 gpurun_out/pytest_gpu.log
