set -x
L=$PWD/3d-human-body-reconstruction_b200
for v in var_fz12a var_fz12c var_fz8; do
  SMPLK_LIB=$L/$v.so python tools/kbench.py 4096 > gpurun_out/k_$v.log 2>&1
  SMPLK_LIB=$L/$v.so python tools/kbench.py 16384 > gpurun_out/k16_$v.log 2>&1
done
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu3.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu3.log
tail -qn1 gpurun_out/k_var*.log gpurun_out/k16_var*.log; tail -4 gpurun_out/r02_pytest_gpu3.log
