set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu6.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu6.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke2.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02_smoke2.log
python tools/ncu_targets.py fit > gpurun_out/plain_fit4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'dA_seg' -s 4 -c 2 -o gpurun_out/r02_daseg python tools/ncu_targets.py fit > gpurun_out/ncu_f_daseg.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu6.log; tail -2 gpurun_out/r02_smoke2.log
