set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02_smoke.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench exit $?" >> gpurun_out/r02_bench_a.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
for t in fused twokernel fit lbs; do
  python tools/ncu_targets.py $t > gpurun_out/plain_$t.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_$t.csv python tools/ncu_targets.py $t > gpurun_out/ncu_l_$t.log 2>&1
done
python tools/ncu_targets.py fused > gpurun_out/plain_fused2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:blend_skin_fused -s 1 -c 1 -o gpurun_out/r02_fused python tools/ncu_targets.py fused > gpurun_out/ncu_f_fused.log 2>&1
python tools/ncu_targets.py twokernel > gpurun_out/plain_two2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:skin_grouped -s 1 -c 1 -o gpurun_out/r02_skin python tools/ncu_targets.py twokernel > gpurun_out/ncu_f_skin.log 2>&1
python tools/ncu_targets.py fit > gpurun_out/plain_fit2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'dA_kernel|skin_fit_l2' -s 2 -c 2 -o gpurun_out/r02_fit python tools/ncu_targets.py fit > gpurun_out/ncu_f_fit.log 2>&1
python tools/ncu_targets.py lbs > gpurun_out/plain_lbs2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:skin_ -s 1 -c 1 -o gpurun_out/r02_lbs python tools/ncu_targets.py lbs > gpurun_out/ncu_f_lbs.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/r02_pytest_gpu.log; tail -2 gpurun_out/r02_smoke.log
