set -x
L=$PWD/3d-human-body-reconstruction_b200
python tools/kbench.py 1024 fit > gpurun_out/k_fit_seg.log 2>&1
SMPLK_LIB=$L/var_ab.so SMPLK_DA_V1=1 python tools/kbench.py 1024 fit > gpurun_out/k_fit_v1.log 2>&1
SMPLK_LIB=$L/var_ab.so python tools/kbench.py 1024 fit > gpurun_out/k_fit_seg_ab.log 2>&1
python tools/kbench.py 4096 fit > gpurun_out/k_fit_seg_4096.log 2>&1
SMPLK_LIB=$L/var_ab.so SMPLK_DA_V1=1 python tools/kbench.py 4096 fit > gpurun_out/k_fit_v1_4096.log 2>&1
python tools/kbench.py 64 fit > gpurun_out/k_fit_seg_64.log 2>&1
SMPLK_LIB=$L/var_ab.so SMPLK_DA_V1=1 python tools/kbench.py 64 fit > gpurun_out/k_fit_v1_64.log 2>&1
python -m pytest tests -m gpu -q -x -k "fit or fitting or loss or backward" > gpurun_out/r02_pytest_gpu8.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu8.log
tail -qn1 gpurun_out/k_fit_*.log; tail -4 gpurun_out/r02_pytest_gpu8.log
