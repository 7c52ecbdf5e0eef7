set -x
L=$PWD/3d-human-body-reconstruction_b200
python -m pytest tests -m gpu -q -x -k "fit or fitting or loss or backward or smoke" > gpurun_out/r02_pytest_gpu10.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu10.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke3.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02_smoke3.log
for B in 1024 4096 64; do
python tools/kbench.py $B fit > gpurun_out/k_fitp_$B.log 2>&1
SMPLK_LIB=$L/var_ab.so SMPLK_FIT_PLANAR=0 python tools/kbench.py $B fit > gpurun_out/k_fiti_$B.log 2>&1
done
tail -qn1 gpurun_out/k_fitp_*.log gpurun_out/k_fiti_*.log; tail -4 gpurun_out/r02_pytest_gpu10.log; tail -2 gpurun_out/r02_smoke3.log
