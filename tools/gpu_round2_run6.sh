set -x
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
(lscpu | head -40; echo; cat /sys/devices/system/node/online 2>/dev/null; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/class); fi; done; free -g) >> gpurun_out/r02_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 tools/d2h_probe.py > gpurun_out/r02_d2h_probe_8.json 2> gpurun_out/d2h8.err
$TR --nproc-per-node 4 --master-port 29512 tools/d2h_probe.py > gpurun_out/r02_d2h_probe_4.json 2> gpurun_out/d2h4.err
$TR --nproc-per-node 2 --master-port 29513 tools/d2h_probe.py > gpurun_out/r02_d2h_probe_2.json 2> gpurun_out/d2h2.err
$TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
cat gpurun_out/r02_d2h_probe_*.json; tail -c 1500 gpurun_out/r02_bench_8gpu.json; tail -3 gpurun_out/r02_bench_8gpu.err
