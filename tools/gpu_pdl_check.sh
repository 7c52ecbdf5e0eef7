set -x
python tools/pdl_ab.py > gpurun_out/r02_pdl_ab.txt 2>&1
python tools/fit_ab.py 1024 pdl >> gpurun_out/r02_pdl_ab.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
cat gpurun_out/r02_pdl_ab.txt; tail -5 gpurun_out/pytest_gpu.log
