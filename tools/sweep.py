"""BASELINE.json configs 4 and 5 on the GPU box (one JSON line per measurement).

  config 4: SMPL-H forward, B in {1K, 4K, 16K, 64K, 256K, 1M} bodies, sharded evenly over the
            ranks (torchrun) -- no collective on the data path; per-rank slices, MAX time over ranks.
            A rank keeps the vertices of its whole slice in HBM (82,680 B/body: 1M bodies on one
            GPU = 82.7 GB); the forward runs in 8192-body chunks through a bounded workspace.
  config 5: 100k-frame motion sequence, full hand pose, ONE betas row broadcast, no render:
            frames/s of the full SMPL-H forward, and of the LBS-only rigged-mesh replay
            (lib/model2video.py RecoverModel) for Nv in {6890, 50k, 200k} with the output of each
            frame chunk overwritten (100k x 200k x 12 B would be 240 GB).

  python tools/sweep.py [--config 4|5|all] [--max-bodies N]
  python -m torch.distributed.run --nproc-per-node G ... tools/sweep.py --config 4
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="all")
    ap.add_argument("--max-bodies", type=int, default=1 << 20)
    ap.add_argument("--frames", type=int, default=100000)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import smplk
    from smplk import _lib, synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream(dev)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def forward_call(dm, betas, pose, transl, verts, joints, ws):
        a = _lib.ForwardArgs()
        a.batch, a.flags = pose.shape[0], 0
        a.betas, a.betas_batch = ctypes.c_void_p(betas.data_ptr()), betas.shape[0]
        a.pose = ctypes.c_void_p(pose.data_ptr())
        a.transl = ctypes.c_void_p(transl.data_ptr())
        a.verts = ctypes.c_void_p(verts.data_ptr())
        a.joints = ctypes.c_void_p(joints.data_ptr()) if joints is not None else None
        a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), ws.numel()
        a.stream = ctypes.c_void_p(stream.cuda_stream)
        dm.forward(a)

    def timed(fn, reps):
        fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        sync_all()
        return max_over_ranks(e0.elapsed_time(e1) / reps)

    model = synthetic.make_model("smplh", seed=0)
    dm = smplk.DeviceModel(model, device=local)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)

    if args.config in ("4", "all"):
        for total in (1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20):
            if total > args.max_bodies:
                break
            n = total // world                      # this rank's slice (contiguous, smplk.sharding)
            if n < 1:
                continue
            free, _ = torch.cuda.mem_get_info(dev)
            need = n * (dm.V * 12 + 700 + dm.J * 12) + dm.workspace_bytes(n, 0)
            if need > free * 0.9:
                if rank == 0:
                    print(json.dumps({"config": 4, "bodies": total, "n_gpus": world, "skipped": "needs %.1f GB per GPU" % (need / 1e9)}), flush=True)
                continue
            betas = torch.randn(n, 16, device=dev, generator=gen)
            pose = torch.randn(n, 156, device=dev, generator=gen) * 0.3
            transl = torch.randn(n, 3, device=dev, generator=gen)
            verts = torch.empty(n, dm.V, 3, device=dev)
            joints = torch.empty(n, dm.J, 3, device=dev)
            ws = torch.empty(dm.workspace_bytes(n, 0), device=dev, dtype=torch.uint8)
            reps = 20 if total <= (1 << 14) else (5 if total <= (1 << 18) else 2)
            ms = timed(lambda: forward_call(dm, betas, pose, transl, verts, joints, ws), reps)
            if rank == 0:
                print(json.dumps({"config": 4, "metric": "smplh_posed_meshes_per_sec_fwd", "bodies": total,
                                  "n_gpus": world, "bodies_per_gpu": n, "ms": ms,
                                  "value": total / (ms * 1e-3), "unit": "meshes/s",
                                  "hbm_gbs_fused_equiv_per_gpu": 84004 * n / (ms * 1e-3) / 1e9,
                                  "note": "vertices of the whole slice kept in HBM (%.1f GB per GPU)" % (n * dm.V * 12 / 1e9)}),
                      flush=True)
            del betas, pose, transl, verts, joints, ws
            torch.cuda.empty_cache()

    if args.config in ("5", "all") and rank == 0:
        N = args.frames
        rng = np.random.default_rng(5)
        # smooth synthetic motion: a 2,689-frame clip (the AMASS clip length of the reference's data/)
        # of low-pass filtered noise, tiled to N frames; translation relative to frame 0
        clip = np.cumsum(rng.standard_normal((2689, 156)) * 0.02, axis=0).astype(np.float32)
        clip = np.clip(clip, -1.5, 1.5)
        pose = torch.tensor(np.tile(clip, (N // 2689 + 1, 1))[:N], device=dev)
        tr = np.cumsum(rng.standard_normal((N, 3)) * 0.01, axis=0).astype(np.float32)
        transl = torch.tensor(tr - tr[0], device=dev)
        betas = torch.randn(1, 16, device=dev, generator=gen)
        chunk = 16384                               # frames whose vertices are resident at once
        verts = torch.empty(chunk, dm.V, 3, device=dev)
        joints = torch.empty(chunk, dm.J, 3, device=dev)
        ws = torch.empty(dm.workspace_bytes(chunk, 0), device=dev, dtype=torch.uint8)

        def run_smplh():
            for f0 in range(0, N, chunk):
                f1 = min(N, f0 + chunk)
                forward_call(dm, betas, pose[f0:f1], transl[f0:f1], verts, joints, ws)
        ms = timed(run_smplh, 2)
        print(json.dumps({"config": 5, "model": "SMPL-H full forward (hands, 16 betas broadcast)", "frames": N,
                          "ms": ms, "value": N / (ms * 1e-3), "unit": "frames/s",
                          "note": "vertices of each %d-frame chunk overwritten by the next" % chunk}), flush=True)
        del verts, joints, ws
        torch.cuda.empty_cache()

        pose24 = pose[:, :72].contiguous()
        for nv in (6890, 50000, 200000):
            rig = synthetic.make_rigged_mesh(nv, seed=13)
            rdm = smplk.DeviceModel(rig, device=local, lbs_only=True)
            ck = max(256, min(16384, int(8e9 // (nv * 12))))
            verts = torch.empty(ck, nv, 3, device=dev)
            ws = torch.empty(rdm.workspace_bytes(ck, 0), device=dev, dtype=torch.uint8)
            p24 = pose24.clone()
            p24.view(N, 24, 3)[:, [13, 14, 22, 23]] = 0.0           # lib/model2video.py:44-45
            nob = torch.zeros(1, 1, device=dev)

            def run_rig():
                for f0 in range(0, N, ck):
                    f1 = min(N, f0 + ck)
                    a = _lib.ForwardArgs()
                    a.batch, a.flags = f1 - f0, 0
                    a.betas, a.betas_batch = None, 1
                    a.pose = ctypes.c_void_p(p24[f0:f1].data_ptr())
                    a.transl = ctypes.c_void_p(transl[f0:f1].data_ptr())
                    a.verts = ctypes.c_void_p(verts.data_ptr())
                    a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), ws.numel()
                    a.stream = ctypes.c_void_p(stream.cuda_stream)
                    rdm.forward(a)
            ms = timed(run_rig, 1 if nv > 6890 else 2)
            gb = N * (2 * nv * 12) / 1e9             # template read (L2) not counted; out write + in-place
            print(json.dumps({"config": 5, "model": "LBS-only rigged mesh (RecoverModel), 24 joints", "verts": nv,
                              "frames": N, "ms": ms, "value": N / (ms * 1e-3), "unit": "frames/s",
                              "hbm_write_gbs": N * nv * 12 / (ms * 1e-3) / 1e9,
                              "note": "output of each %d-frame chunk overwritten by the next" % ck}), flush=True)
            del verts, ws, rdm
            torch.cuda.empty_cache()

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
