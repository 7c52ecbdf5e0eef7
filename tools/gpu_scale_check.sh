# bench.py at 2 and 4 GPUs of one box (run with: gpurun --gpus 4 -- 'bash tools/gpu_scale_check.sh')
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
$TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err
$TR --nproc-per-node 4 --master-port 29523 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_4gpu.json 2> gpurun_out/r02_bench_ref_4gpu.err
python -m pytest tests/test_sharding_gloo.py -m gpu -q > gpurun_out/pytest_shard_multi.log 2>&1
tail -c 400 gpurun_out/r02_bench_4gpu.err; tail -2 gpurun_out/pytest_shard_multi.log
