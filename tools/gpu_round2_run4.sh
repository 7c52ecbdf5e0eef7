set -x
L=$PWD/3d-human-body-reconstruction_b200
python tools/skin_bench.py > gpurun_out/skin_nobar.log 2>&1
SMPLK_LIB=$L/var_bar.so python tools/skin_bench.py > gpurun_out/skin_bar.log 2>&1
SMPLK_LIB=$L/var_pref.so python tools/skin_bench.py > gpurun_out/skin_pref.log 2>&1
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu4.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu4.log
cat gpurun_out/skin_nobar.log gpurun_out/skin_bar.log gpurun_out/skin_pref.log; tail -4 gpurun_out/r02_pytest_gpu4.log
