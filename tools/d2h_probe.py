"""Where does the end-to-end step go at N GPUs?  Device -> pinned-host copy bandwidth per GPU, (a) every rank
copying at once (what a sharded e2e step does), (b) one rank at a time.  Launch under torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/d2h_probe.py
Prints one JSON line from rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from bench import bind_to_gpu_numa_node


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if os.environ.get("SMPLK_BENCH_NUMA", "1") != "0" else {"bound": False}
    if world > 1:
        dist.init_process_group("gloo")
    nbytes = 4096 * 6890 * 12
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    stream = torch.cuda.current_stream(dev)

    def timed(n=10):
        for _ in range(2):
            h.copy_(d, non_blocking=True)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            h.copy_(d, non_blocking=True)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        return nbytes * n / (e0.elapsed_time(e1) * 1e-3) / 1e9

    def gather(x):
        if world == 1:
            return [x]
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    if world > 1:
        dist.barrier()
    together = gather(timed())
    solo = []
    for r in range(world):
        if world > 1:
            dist.barrier()
        v = timed() if r == rank else None
        if world > 1:
            dist.barrier()
        solo.append(v)
    solo_all = gather(solo[rank])
    numas = gather(numa)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bytes": nbytes, "d2h_gbs_all_ranks_at_once": together,
                          "aggregate_gbs": sum(together), "d2h_gbs_one_rank_at_a_time": solo_all, "numa": numas,
                          "cpus": os.cpu_count()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
