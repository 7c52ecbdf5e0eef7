"""CPU, world_size 2 over gloo: the multi-GPU path shards bodies with no data-path collective;
the only communication is the harness' barrier + max-reduce of the step time."""
import os
import sys

import numpy as np
import pytest
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


def _gpu_worker(rank, world, port, total, out_dir):
    """One process per rank; ranks map onto the visible GPUs round-robin (a 1-GPU box runs both ranks
    on cuda:0).  The rank's slice goes through the PRODUCT path: DeviceModel -> smplk_forward."""
    sys.path.insert(0, ROOT)
    import smplk
    from smplk import synthetic
    from smplk.sharding import forward_shard, max_over_ranks_ms
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    device = rank % torch.cuda.device_count()
    model = synthetic.make_model("smplh", seed=21)
    dm = smplk.DeviceModel(model, device=device)
    betas, pose, transl = synthetic.make_inputs(model, total, seed=5)
    before = smplk._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.device(device):
        dist.barrier()
        e0.record()
        lo, hi, v, j = forward_shard(dm, betas, pose, transl, world_size=world, rank=rank)
        e1.record()
        torch.cuda.synchronize()
    assert smplk._lib.launch_count() > before, "the shard did not run through libsmplk.so"
    np.save(os.path.join(out_dir, "gshard%d.npy" % rank), v.cpu().numpy())
    np.save(os.path.join(out_dir, "gbounds%d.npy" % rank), np.array([lo, hi]))
    ms = max_over_ranks_ms(e0.elapsed_time(e1), dist)
    assert ms >= e0.elapsed_time(e1) - 1e-9
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_rank_sharding_runs_the_cuda_path(tmp_path):
    """world_size 2, one process per rank, each with its own model handle: the concatenated shard
    results equal the single-call result BITWISE and the float64 oracle within 1e-5 m."""
    import smplk
    from smplk import synthetic
    from smplk.body_models import body_model_apply
    from oracle import smpl_oracle as O
    total, world = 301, 2
    port = 29500 + (os.getpid() + 7) % 2000
    mp.spawn(_gpu_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / ("gshard%d.npy" % r)) for r in range(world)])
    bounds = [tuple(np.load(tmp_path / ("gbounds%d.npy" % r))) for r in range(world)]
    assert bounds == [(0, 151), (151, 301)]
    model = synthetic.make_model("smplh", seed=21)
    betas, pose, transl = synthetic.make_inputs(model, total, seed=5)
    dm = smplk.DeviceModel(model, device=0)
    t = lambda a: torch.tensor(a, device="cuda:0")
    single = body_model_apply(dm, t(betas), t(pose), transl=t(transl))[0].cpu().numpy()
    assert got.shape == single.shape and np.array_equal(got, single)
    ref = O.TorchOracleModel(model, dtype=torch.float64).forward_full_pose(
        *[torch.tensor(x, dtype=torch.float64) for x in (betas, pose, transl)]).vertices.numpy()
    assert np.abs(got - ref).max() <= 1e-5
