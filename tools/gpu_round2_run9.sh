set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu9.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu9.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench exit $?" >> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
python tools/ncu_targets.py fit > gpurun_out/plain_fit.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_fit.csv python tools/ncu_targets.py fit > gpurun_out/ncu_l_fit.log 2>&1
python tools/ncu_targets.py fit > gpurun_out/plain_fit2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'dA_seg|skin_fit_l2|pose_' -s 4 -c 4 -o gpurun_out/r02_fit python tools/ncu_targets.py fit > gpurun_out/ncu_f_fit.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu9.log; tail -c 300 gpurun_out/r02_bench.err
