#!/bin/bash
# Reproduces the round-2 evidence under profiles/ on a B200 box (run through gpurun; everything lands in gpurun_out/):
#   gpurun --timeout 1800 -- 'bash tools/gpu_evidence.sh'
# then, here:  python tools/ncu_summary.py gpurun_out/r02_fused.ncu-rep   (etc.; see profiles/README.md)
set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
python tools/hbm_probe.py 4 > gpurun_out/r02_hbm_probe.json
python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
# launch lists (cold-cache, serialised: compare shares) -- each only after the plain run exited 0
python bench.py --steps 30 --warmup 5 --no-extras > gpurun_out/plain_bench.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv \
      python bench.py --steps 30 --warmup 5 --no-extras > gpurun_out/ncu_bench.log 2>&1
for t in fused twokernel fit lbs; do
  python tools/ncu_targets.py $t > gpurun_out/plain_$t.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_$t.csv \
      python tools/ncu_targets.py $t > gpurun_out/ncu_l_$t.log 2>&1
done
# one --set full capture per kernel family
cap() {  # name, target args, kernel regex, skip, count
  python tools/ncu_targets.py $2 > gpurun_out/plain_cap_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c $5 -o gpurun_out/r02_$1 \
      python tools/ncu_targets.py $2 > gpurun_out/ncu_f_$1.log 2>&1
}
cap fused fused blend_skin_fused 1 1
cap twokernel twokernel 'skin_grouped|blend_tcgen05' 2 2
cap fit fit 'dA_seg|skin_fit_l2|pose_' 4 4
cap lbs lbs 'lbs_replay_gemm|skin_' 1 1
cap lbs200k "lbs 200000" 'lbs_replay_gemm|skin_' 1 1
# probes behind the round-2 store-path decisions (tools/micro/*.cu are built by the caller: see their headers)
if [ -x tools/micro/tma_store_probe ]; then
  for a in "4 0 36" "4 4 36" "4 2 36" "4 1 36" "4 20670 36" "8 1 18" "8 10335 18" "4 2 32" "4 20672 36"; do
    timeout 30 tools/micro/tma_store_probe $a; done > gpurun_out/r02_tma_store_probe.txt 2>&1
fi
if [ -x tools/micro/store_pattern_probe ]; then
  { timeout 60 tools/micro/store_pattern_probe 6890 16384; timeout 60 tools/micro/store_pattern_probe 50000 2048; } > gpurun_out/r02_store_pattern_probe.txt 2>&1
fi
python tools/replay_ab.py > gpurun_out/r02_replay_gemm_ab.txt 2>&1
python tools/skin_gemm_ab.py > gpurun_out/r02_skin_gemm_ab.txt 2>&1
python tools/fz_ab.py > gpurun_out/r02_fused_tma_out_ab.txt 2>&1
# programmatic dependent launch on / off (forward sizes, fitting step) and the fitting step's per-kernel times
{ python tools/pdl_ab.py; python tools/fit_ab.py 1024 pdl; python tools/fit_profile.py 1024; } > gpurun_out/r02_pdl_ab.txt 2>&1
