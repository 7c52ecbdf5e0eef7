set -x
L=3d-human-body-reconstruction_b200
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu2.log
SMPLK_LIB=$PWD/$L/var_ab_nof2.so SMPLK_SKIN_G8=0 python tools/skin_bench.py > gpurun_out/skin_nof2_g4.log 2>&1
SMPLK_LIB=$PWD/$L/var_ab_nof2.so SMPLK_SKIN_G8=1 python tools/skin_bench.py > gpurun_out/skin_nof2_g8.log 2>&1
SMPLK_LIB=$PWD/$L/var_ab_f2.so SMPLK_SKIN_G8=0 python tools/skin_bench.py > gpurun_out/skin_f2_g4.log 2>&1
SMPLK_LIB=$PWD/$L/var_ab_f2.so SMPLK_SKIN_G8=1 python tools/skin_bench.py > gpurun_out/skin_f2_g8.log 2>&1
python tools/ncu_targets.py fit > gpurun_out/plain_fit3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'pose_' -s 2 -c 2 -o gpurun_out/r02_pose python tools/ncu_targets.py fit > gpurun_out/ncu_f_pose.log 2>&1
python tools/ncu_targets.py lbs > gpurun_out/plain_lbs3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:skin_ -s 1 -c 1 -o gpurun_out/r02_lbs_g8 python tools/ncu_targets.py lbs > gpurun_out/ncu_f_lbs8.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu2.log; cat gpurun_out/skin_*.log
