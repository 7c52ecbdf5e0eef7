"""CPU, world_size 2 over gloo: the multi-GPU path shards bodies with no data-path collective;
the only communication is the harness' barrier + max-reduce of the step time."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, out_dir):
    sys.path.insert(0, ROOT)
    from smplk.sharding import shard_bounds, max_over_ranks_ms
    from smplk import synthetic
    from oracle import smpl_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(total, world, rank)
    model = synthetic.make_model("smpl", seed=1, num_verts=200)
    om = O.TorchOracleModel(model, dtype=torch.float64)
    betas, pose, transl = synthetic.make_inputs(model, total, seed=3, dtype=np.float64)
    # each rank evaluates only its slice (stand-in for the per-GPU C-ABI call)
    out = om.forward_full_pose(torch.tensor(betas[lo:hi]), torch.tensor(pose[lo:hi]), torch.tensor(transl[lo:hi]))
    np.save(os.path.join(out_dir, "shard%d.npy" % rank), out.vertices.numpy())
    dist.barrier()
    ms = max_over_ranks_ms(10.0 + rank, dist)
    assert ms == 10.0 + world - 1
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(tmp_path):
    from smplk import synthetic
    from smplk.sharding import shard_bounds
    from oracle import smpl_oracle as O
    total, world = 7, 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    model = synthetic.make_model("smpl", seed=1, num_verts=200)
    om = O.TorchOracleModel(model, dtype=torch.float64)
    betas, pose, transl = synthetic.make_inputs(model, total, seed=3, dtype=np.float64)
    full = om.forward_full_pose(torch.tensor(betas), torch.tensor(pose), torch.tensor(transl)).vertices.numpy()
    got = np.concatenate([np.load(tmp_path / ("shard%d.npy" % r)) for r in range(world)])
    assert got.shape == full.shape and np.abs(got - full).max() < 1e-12  # CPU BLAS is not batch-invariant
    spans = [shard_bounds(total, world, r) for r in range(world)]
    assert spans == [(0, 4), (4, 7)]


def test_shard_bounds_cover_everything():
    from smplk.sharding import shard_bounds
    for total in (0, 1, 5, 4096, 1000003):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
