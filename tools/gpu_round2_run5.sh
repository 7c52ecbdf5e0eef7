set -x
python tools/hbm_probe.py 4 > gpurun_out/r02_hbm_probe.json 2> gpurun_out/hbm_probe.err
python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench exit $?" >> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
python bench.py --steps 30 --warmup 5 --no-extras > gpurun_out/plain_bench_ne.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 30 --warmup 5 --no-extras > gpurun_out/ncu_bench_ne.log 2>&1
for t in fused twokernel fit lbs; do
  python tools/ncu_targets.py $t > gpurun_out/plain_$t.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_$t.csv python tools/ncu_targets.py $t > gpurun_out/ncu_l_$t.log 2>&1
done
python tools/ncu_targets.py fused > gpurun_out/plain_fused2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:blend_skin_fused -s 1 -c 1 -o gpurun_out/r02_fused python tools/ncu_targets.py fused > gpurun_out/ncu_f_fused.log 2>&1
python tools/ncu_targets.py twokernel > gpurun_out/plain_two2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'skin_grouped|blend_tcgen05' -s 2 -c 2 -o gpurun_out/r02_twokernel python tools/ncu_targets.py twokernel > gpurun_out/ncu_f_two.log 2>&1
python tools/ncu_targets.py fit > gpurun_out/plain_fit2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'dA_kernel|skin_fit_l2|pose_' -s 5 -c 5 -o gpurun_out/r02_fit python tools/ncu_targets.py fit > gpurun_out/ncu_f_fit.log 2>&1
python tools/ncu_targets.py lbs > gpurun_out/plain_lbs2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:skin_ -s 1 -c 1 -o gpurun_out/r02_lbs python tools/ncu_targets.py lbs > gpurun_out/ncu_f_lbs.log 2>&1
python tools/ncu_targets.py lbs 200000 > gpurun_out/plain_lbs3.log 2>&1 && ncu --set full --clock-control none -k regex:skin_ -s 1 -c 1 -o gpurun_out/r02_lbs200k python tools/ncu_targets.py lbs 200000 > gpurun_out/ncu_f_lbs200k.log 2>&1
cat gpurun_out/r02_hbm_probe.json; tail -c 600 gpurun_out/r02_bench.err
