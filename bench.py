#!/usr/bin/env python
"""Benchmark of the SMPL-H body-model hot path (BASELINE.json metric: posed meshes/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One "step" = one forward pass (pose/FK -> blend GEMM -> skinning) over a batch of B synthetic
bodies PER GPU (weak scaling; bodies are independent, no data-path collective).  Workload at N=1:
BASELINE.json configs[1] -- SMPL-H forward+LBS, batch 4096, fp32, 52 joints, 16 betas, 459 posedirs.

Printed JSON line (rank 0): value = whole-job posed meshes/s with inputs resident in HBM;
e2e = same metric through the C-ABI host-buffer call (H2D of inputs + D2H of vertices inside the
timed region); roofline = dominant kernel (fused tcgen05 blend GEMM + skinning epilogue) against the
tensor roofline, with its HBM figure and the stand-alone GEMM / skinning kernels reported beside it; cpu_baseline = the oracle port timed on the host
cores.  `--impl reference` times the CPU restatement of the reference path (oracle port; the
reference is pure Python and /root/reference does not exist on the GPU box).
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_BYTES_SKIN = 167856          # v_posed in + A in + verts out, per body
FWD_BYTES_FUSED = 84004          # inputs + verts + FK joints, per body
GEMM_FLOPS_PER_BODY = 2 * 20670 * (459 + 16)   # fp32-equivalent algorithmic FLOPs, K = 459 + 16
METRIC = "smplh_posed_meshes_per_sec_fwd"
UNIT = "meshes/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Polls SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.001)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "sm_mhz_min": float(min(self.samples)),
                "note": "sampled through NVML during the timed steps and an untimed continuation of the same steps"}


def _host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 for its children)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return max(1, os.cpu_count() or 1)


def _median_ms(fn, n):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


def cpu_baseline(model, sample_bodies, budget_s=12.0):
    """The reference's CPU path restated (oracle/) and timed on this box's host cores, bounded samples:

      value (C4)  torch-CPU fp32 restatement of upstream smplx.lbs, all host threads, passes of the SAME
                  step as the GPU arm
      c1 / c2     numpy float64 restatement of models/smpl_np.py / models/smplh_np.py set_params, one body
                  per call, median of 20 calls (BASELINE config 1; C1 / C2)
      c3          the per-frame replay loop of lib/model2video.py:514-518 over 1,000 frames of the AMASS
                  clip fixture: LBS-only rigged mesh (RecoverModel math, 24 joints, 6,890 vertices), and the
                  full SMPL-H forward over 200 frames (C3)
    """
    import torch
    from oracle import smpl_oracle as O
    from smplk import clips, synthetic
    torch.set_num_threads(_host_threads())
    om = O.TorchOracleModel(model, dtype=torch.float32)
    betas, pose, transl = synthetic.make_inputs(model, sample_bodies, seed=123)
    tb, tp, tt = torch.tensor(betas), torch.tensor(pose), torch.tensor(transl)
    with torch.no_grad():
        om.forward_full_pose(tb, tp, tt)
        n, t0 = 0, time.perf_counter()
        while True:
            om.forward_full_pose(tb, tp, tt)
            n += 1
            el = time.perf_counter() - t0
            if el > budget_s or n >= 200:
                break
    out = {"value": sample_bodies * n / el, "unit": UNIT, "cores": torch.get_num_threads(),
           "kind": "port",
           "sample": "%d passes of %d bodies (torch-CPU fp32 restatement of smplx.lbs, oracle/smpl_oracle.py), %.1f s"
                     % (n, sample_bodies, el)}
    try:
        rng = np.random.default_rng(0)
        ms_ = synthetic.make_model("smpl", num_betas=10, seed=8)
        mh_ = synthetic.make_model("smplh", num_betas=10, seed=7)
        p24, p52 = rng.standard_normal((24, 3)) * 0.3, rng.standard_normal((52, 3)) * 0.3
        be, tr = rng.standard_normal(10), rng.standard_normal(3)
        c1 = _median_ms(lambda: O.np_forward(ms_, p24, be, tr), 20)
        c2 = _median_ms(lambda: O.np_forward(mh_, p52, be, tr), 20)
        out["c1"] = {"ms_per_body": c1, "value": 1e3 / c1, "unit": UNIT, "cores": "numpy BLAS default",
                     "what": "SMPL B=1 float64, oracle np_forward (models/smpl_np.py:158-206 set_params), median of 20"}
        out["c2"] = {"ms_per_body": c2, "value": 1e3 / c2, "unit": UNIT, "cores": "numpy BLAS default",
                     "what": "SMPL-H B=1 float64, oracle np_forward (models/smplh_np.py:39-86 set_params), median of 20"}
        clip = clips.read_amsass(os.path.join(ROOT, "tests", "golden", "amass_clip_09_05.npz"), full=True)
        n0 = clip.poses.shape[0]
        rig = synthetic.make_rigged_mesh(6890, seed=13)
        frames = 1000
        t0 = time.perf_counter()
        for i in range(frames):                      # lib/model2video.py:514-518: one set_params per frame
            O.np_lbs_only(rig, clip.poses[i % n0, :72], clip.trans[i % n0])
        el_rig = time.perf_counter() - t0
        fh = 200
        t0 = time.perf_counter()
        for i in range(fh):
            O.np_forward(model, clip.poses[i % n0], clip.betas[:model["shapedirs"].shape[2]], clip.trans[i % n0])
        el_h = time.perf_counter() - t0
        out["c3"] = {"lbs_only_frames_per_s": frames / el_rig, "lbs_only_sample": "%d frames, 6,890-vertex rig, %.1f s" % (frames, el_rig),
                     "lbs_only_100k_frames_s": 1e5 * el_rig / frames,
                     "smplh_frames_per_s": fh / el_h, "smplh_sample": "%d frames, %.1f s" % (fh, el_h),
                     "smplh_100k_frames_s": 1e5 * el_h / fh,
                     "what": "per-frame replay loop (lib/model2video.py:514-518) over the tiled AMASS clip, numpy float64 oracle"}
    except Exception as e:       # the extra baselines never take the bench line down
        out["c_error"] = repr(e)
    return out


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (oracle port: the reference is pure
    Python over un-vendored smplx, and /root/reference does not exist on the GPU box), all host threads,
    each step one pass over the SAME batch as the GPU arm's step."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(_host_threads())
    from smplk import synthetic
    model = synthetic.make_model("smplh", seed=0)
    from oracle import smpl_oracle as O
    om = O.TorchOracleModel(model, dtype=torch.float32)
    B = args.batch
    betas, pose, transl = synthetic.make_inputs(model, B, seed=1)
    tb, tp, tt = torch.tensor(betas), torch.tensor(pose), torch.tensor(transl)
    sub = 512          # bodies are independent: the step is evaluated in slices to bound host memory (T is B x V x 16 floats)

    def step():
        for c0 in range(0, B, sub):
            om.forward_full_pose(tb[c0:c0 + sub], tp[c0:c0 + sub], tt[c0:c0 + sub])
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        el = time.perf_counter() - t0
    val = B * args.steps / el
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(B, args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d steps x %d bodies (one GPU's share of a step; in slices of %d), torch-CPU fp32 "
                                       "restatement of the reference torch path" % (args.steps, B, sub)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


NSETS = 4


def workload_config(B, world):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": "SMPL-H forward+LBS batch %d per GPU, fp32, 52 joints, 16 betas, 459 posedirs "
                        "(BASELINE.json configs[1])" % B,
            "global_batch": world * B, "parallelism": "batch-sharded x%d, no collective" % world,
            "l2": "per-step output %.0f MB (verts) > 126 MB L2; %d rotating input sets" % (B * 82680 / 1e6, NSETS),
            "weights": "canonically sparse LBS weights (<=4 per vertex)"}


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and so its pinned staging buffers, first-touch) to the NUMA node the
    rank's GPU hangs off: with 8 ranks copying 339 MB per step each, buffers that all sit on one node
    make that node's memory controllers / the socket interconnect the bottleneck."""
    info = {"bound": False}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        pci = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % pci).read().strip())
        info.update(pci=pci, numa_node=node)
        if node < 0:
            return info
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as e:
        info["error"] = repr(e)
    return info


import contextlib


@contextlib.contextmanager
def _stdout_to_stderr():
    """Point file descriptor 1 at stderr for the duration (covers native writes, not just sys.stdout)."""
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4096, help="bodies per GPU per step")
    ap.add_argument("--bwd-batch", type=int, default=1024)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip fwd+bwd / e2e / per-kernel passes")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import smplk
    from smplk import _lib, synthetic
    from smplk.body_models import body_model_apply

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if os.environ.get("SMPLK_BENCH_NUMA", "1") != "0" else {"bound": False}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries the one JSON line: whatever the communicator set-up prints there (NCCL's
        # "NCCL version ..." banner is written to file descriptor 1 by native code) goes to stderr
        with _stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            try:                                      # the communicator is created here, inside the redirection
                dist.barrier()
                torch.cuda.synchronize(dev)
            except Exception as e:                    # only the banner's destination depends on it
                sys.stderr.write("early barrier skipped: %r\n" % (e,))

    B = args.batch
    model = synthetic.make_model("smplh", seed=0)
    # SMPLK_BLEND=tf32 in the bench's environment runs the whole line on 3xTF32 operands (handle option blend_tf32)
    dm = smplk.DeviceModel(model, device=local, options={"blend_tf32": 1} if os.environ.get("SMPLK_BLEND") == "tf32" else None)
    # NSETS rotating input sets; the per-step output (339 MB of vertices) already exceeds the 126 MB L2
    sets = []
    for s in range(NSETS):
        b, p, t = synthetic.make_inputs(model, B, seed=10 * rank + s)
        sets.append(tuple(torch.tensor(x, device=dev) for x in (b, p, t)))
    verts = torch.empty(B, dm.V, 3, device=dev)
    joints = torch.empty(B, dm.J, 3, device=dev)
    ws_bytes = dm.workspace_bytes(B, 0)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    stream = torch.cuda.current_stream(dev)

    def step(i):
        b, p, t = sets[i % NSETS]
        a = _lib.ForwardArgs()
        a.batch, a.flags = B, 0
        a.betas, a.betas_batch = ctypes.c_void_p(b.data_ptr()), B
        a.pose, a.transl = ctypes.c_void_p(p.data_ptr()), ctypes.c_void_p(t.data_ptr())
        a.verts, a.joints = ctypes.c_void_p(verts.data_ptr()), ctypes.c_void_p(joints.data_ptr())
        a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), ws_bytes
        a.stream = ctypes.c_void_p(stream.cuda_stream)
        dm.forward(a)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    barrier()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record(stream)
        for i in range(args.steps):
            step(i)
        e1.record(stream)
        barrier()
        launches = _lib.launch_count() - launches0
        n_timed = len(clocks.samples)
        # the timed region lasts a few milliseconds, an NVML query about one: keep the same step running (untimed)
        # until the sampler has seen the clocks UNDER THIS LOAD at least 20 times
        t_probe = time.perf_counter()
        while len(clocks.samples) < n_timed + 20 and time.perf_counter() - t_probe < 2.0:
            for i in range(20):
                step(i)
            torch.cuda.synchronize(dev)
    ms_local = e0.elapsed_time(e1)
    ms = ms_local
    if world > 1:
        t = torch.tensor([ms_local], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    extras = {}
    peaks = load_peaks()
    if not args.no_extras:
        # ---- per-kernel device time (CUDA events on the launching stream, inside libsmplk)
        dm.profile_enable(True)
        for i in range(min(args.steps, 50)):
            step(i)
        torch.cuda.synchronize(dev)
        prof = dm.profile_read(reset=True)
        dm.profile_enable(False)
        kern = {k: (v[0] / v[1]) for k, v in prof.items() if v[1] > 0}
        # The dominant kernel's AVERAGE LAUNCH DURATION for the roofline: `steps` back-to-back launches of it alone between
        # two CUDA events on the launching stream (handle option skip_pose: no pose kernel, the workspace rows of the
        # previous call are reused; same inputs, same 339 MB of output per launch).  The per-launch event brackets above
        # add the event-to-kernel latency to every launch (kernel_ms: 5-10 us over this figure) and break the programmatic
        # dependent launch, so they are kept for the kernels' SHARES of the step only.
        kern_b2b = None
        if "blend_skin_fused" in kern:
            step(0)
            dm.set_option("skip_pose", 1)
            for _ in range(3):
                step(0)
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nrep = max(args.steps, 20)
            torch.cuda.synchronize(dev)
            k0.record(stream)
            for _ in range(nrep):
                step(0)
            k1.record(stream)
            torch.cuda.synchronize(dev)
            dm.set_option("skip_pose", 0)
            kern_b2b = k0.elapsed_time(k1) / nrep
        # TF32 tensor peak measured here with a cuBLAS 8192^3 TF32 GEMM (MEASURED_PEAKS.json has bf16 only)
        tf32_peak = None
        if rank == 0:
            torch.backends.cuda.matmul.allow_tf32 = True
            x = torch.randn(8192, 8192, device=dev)
            y = torch.randn(8192, 8192, device=dev)
            for _ in range(3):
                x @ y
            torch.cuda.synchronize(dev)
            best = 1e9
            for _ in range(8):
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(); x @ y; a1.record(); torch.cuda.synchronize(dev)
                best = min(best, a0.elapsed_time(a1))
            tf32_peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
            torch.backends.cuda.matmul.allow_tf32 = False
            del x, y
        fmt = os.environ.get("SMPLK_BLEND", "f16")
        # fp32-accurate contraction = 3 tensor-core passes over two-term-split operands:
        #   default: fp16 split, kind::f16 -> ceiling = measured bf16/fp16 dense peak / 3
        #   SMPLK_BLEND=tf32: 3xTF32   -> ceiling = TF32 dense peak (cuBLAS, measured here) / 3
        dense_peak = peaks["bf16"] if fmt != "tf32" else (tf32_peak or peaks["bf16"] / 2)
        ncu_traffic = None
        for name in ("r02_ncu_fused_traffic.json", "r01_ncu_fused_traffic.json"):
            tp = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tp):
                ncu_traffic = json.load(open(tp))
                break

        def tensor_roofline(name, g_ms, npad, note_extra=""):
            alg_tflops = GEMM_FLOPS_PER_BODY * B / (g_ms * 1e-3) / 1e12
            issued = 3 * 2 * B * npad * 480 / (g_ms * 1e-3) / 1e12
            return {"kernel": name, "bound": "tensor", "achieved": alg_tflops, "peak": dense_peak / 3.0,
                    "unit": "TFLOP/s", "frac": alg_tflops / (dense_peak / 3.0), "traffic": None,
                    "ms_per_launch": g_ms,
                    "note": "achieved = 2*B*20670*475 fp32-equivalent FLOPs / CUDA-event time on the launching stream; "
                            "peak = dense tensor peak / 3 passes (%s: %.0f TFLOP/s, %s burst figure; sustained %.0f); "
                            "issued tensor FLOPs = 3*2*B*%d*480%s" % (
                                "bf16/fp16" if fmt != "tf32" else "tf32 cuBLAS 8192^3 measured in this run",
                                dense_peak, peaks["source"], peaks["bf16_sustained"], npad, note_extra),
                    "tf32_tflops_measured": tf32_peak, "tensor_tflops_issued": issued,
                    "tensor_frac_issued": issued / dense_peak}

        if "blend_skin_fused" in kern:
            # the forward's dominant kernel: blend GEMM + skinning epilogue in one launch
            f_ms = kern_b2b or kern["blend_skin_fused"]
            r = tensor_roofline("blend_skin_fused_kernel", f_ms, 83 * 256,
                                "; the epilogue (LBS skinning from TMEM, CUDA cores) runs under the MMAs; ms_per_launch = "
                                "average of back-to-back launches of this kernel alone between two CUDA events "
                                "(ms_per_launch_event_bracketed = one event pair around every launch, inside libsmplk)")
            r["ms_per_launch_event_bracketed"] = kern["blend_skin_fused"]
            if ncu_traffic:
                r["traffic"] = ncu_traffic.get("dram_bytes_per_launch")
                r["traffic_note"] = ncu_traffic.get("note")
            extras["roofline"] = r
            gbs = FWD_BYTES_FUSED * B / (f_ms * 1e-3) / 1e9
            extras["roofline_hbm_fused"] = {
                "kernel": "blend_skin_fused_kernel", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm"],
                "unit": "GB/s", "frac": gbs / peaks["hbm"], "traffic": r["traffic"], "ms_per_launch": f_ms,
                "note": "84,004 algorithmic B/body (inputs + verts + FK joints; v_posed never leaves the SM); "
                        "this kernel is tensor-bound, the HBM figure shows how far below the memory roofline it sits"}
        elif "blend_tcgen05" in kern:
            extras["roofline"] = tensor_roofline("blend_tcgen05_2cta_kernel<%s>" % ("f16" if fmt != "tf32" else "tf32"),
                                                 kern["blend_tcgen05"], 20736)
        # the stand-alone kernels (forward with SAVE_FOR_BACKWARD, LBS-only models, dense weights):
        # timed through a second handle created with the option fused=0
        if "blend_skin_fused" in kern and rank == 0:
            dm_u = smplk.DeviceModel(model, device=local, options={"fused": 0})
            dm_u.profile_enable(True)
            for i in range(20):
                b_, p_, t_ = sets[i % NSETS]
                a = _lib.ForwardArgs()
                a.batch, a.flags = B, 0
                a.betas, a.betas_batch = ctypes.c_void_p(b_.data_ptr()), B
                a.pose, a.transl = ctypes.c_void_p(p_.data_ptr()), ctypes.c_void_p(t_.data_ptr())
                a.verts, a.joints = ctypes.c_void_p(verts.data_ptr()), ctypes.c_void_p(joints.data_ptr())
                wsb = dm_u.workspace_bytes(B, 0)
                if wsb > ws.numel():
                    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
                a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), ws.numel()
                a.stream = ctypes.c_void_p(stream.cuda_stream)
                dm_u.forward(a)
            torch.cuda.synchronize(dev)
            pu = {k: (v[0] / v[1]) for k, v in dm_u.profile_read(reset=True).items() if v[1] > 0}
            kern.update({"unfused_" + k: v for k, v in pu.items()})
            if "blend_tcgen05" in pu:
                extras["roofline_blend_gemm"] = tensor_roofline("blend_tcgen05_2cta_kernel<f16>", pu["blend_tcgen05"], 20736)
            if "skin" in pu:
                kern["skin"] = pu["skin"]
        if "skin" in kern:
            s_ms = kern["skin"]
            gbs = FWD_BYTES_SKIN * B / (s_ms * 1e-3) / 1e9
            extras["roofline_skinning"] = {"kernel": "skin_grouped_kernel (two-kernel forward: SAVE_FOR_BACKWARD / handle option fused=0)",
                                           "bound": "hbm", "achieved": gbs,
                                           "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                                           "traffic": None, "ms_per_launch": s_ms,
                                           "note": "167,856 algorithmic B/body (v_posed in, A in, verts out); peak %s" % peaks["source"]}
        extras["kernel_ms"] = kern

        # ---- forward+backward fitting step (config 3: vertex L2 loss, batch 1024)
        Bb = args.bwd_batch
        bb, pb, tb_ = (torch.tensor(x, device=dev, requires_grad=True) for x in synthetic.make_inputs(model, Bb, seed=99))
        target = torch.randn(Bb, dm.V, 3, device=dev)

        from smplk.body_models import vertex_l2_loss

        from smplk.body_models import fit_vertex_l2

        def fb_step(fused):
            for t_ in (bb, pb, tb_):
                t_.grad = None
            if fused == "node":       # body model + loss as one autograd node (smplk.fit_vertex_l2)
                loss = fit_vertex_l2(dm, bb, pb, target, transl=tb_, reduce="sum")
            else:
                v, _, _, _ = body_model_apply(dm, bb, pb, transl=tb_)
                loss = vertex_l2_loss(v, target).sum() if fused else ((v - target) ** 2).sum()
            loss.backward()
            return loss

        def time_fb(fused):
            for _ in range(3):
                fb_step(fused)
            torch.cuda.synchronize(dev)
            nfb = max(5, min(args.steps, 30))
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            for _ in range(nfb):
                fb_step(fused)
            f1.record(stream)
            torch.cuda.synchronize(dev)
            return f0.elapsed_time(f1) / nfb
        fb_ms, fb_torch_ms, fb_two_ms = time_fb("node"), time_fb(False), time_fb(True)
        # the same step captured once into a CUDA graph and replayed (the ~10 short kernels of a
        # batch-1024 step are launch-bound from Python)
        fb_graph_ms = None
        try:
            gstream = torch.cuda.Stream(dev)
            gstream.wait_stream(stream)
            with torch.cuda.stream(gstream):
                for _ in range(3):
                    fb_step("node")
            stream.wait_stream(gstream)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fb_step("node")
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize(dev)
            nfb = max(5, min(args.steps, 30))
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            for _ in range(nfb):
                graph.replay()
            f1.record(stream)
            torch.cuda.synchronize(dev)
            fb_graph_ms = f0.elapsed_time(f1) / nfb
        except Exception as e:   # graph capture is an optimisation of the harness, not of the path
            fb_graph_ms = None
            sys.stderr.write("fwd+bwd CUDA-graph variant skipped: %r\n" % (e,))
        extras["fwd_bwd"] = {"metric": "smplh_fitting_steps_meshes_per_sec_fwd_bwd", "batch": Bb,
                             "value": world * Bb / (fb_ms * 1e-3), "unit": UNIT, "ms_per_step": fb_ms,
                             "loss": "sum ||V - V*||^2, body model + fused loss/gradient kernel as one autograd node "
                                     "(smplk.fit_vertex_l2), grads w.r.t. betas, pose, transl",
                             "ms_per_step_two_nodes": fb_two_ms,
                             "value_two_nodes": world * Bb / (fb_two_ms * 1e-3),
                             "two_nodes": "body_model_apply(...) then smplk.vertex_l2_loss(v, target) as separate autograd nodes",
                             "ms_per_step_torch_loss": fb_torch_ms,
                             "value_torch_loss": world * Bb / (fb_torch_ms * 1e-3),
                             "ms_per_step_cuda_graph": fb_graph_ms,
                             "value_cuda_graph": (world * Bb / (fb_graph_ms * 1e-3)) if fb_graph_ms else None}

        dm.profile_enable(True)
        dm.profile_read()
        for _ in range(10):
            fb_step("node")
        torch.cuda.synchronize(dev)
        fbk = {k: v[0] / v[1] for k, v in dm.profile_read(reset=True).items() if v[1] > 0}
        dm.profile_enable(False)
        extras["fwd_bwd"]["kernel_ms"] = fbk
        extras["fwd_bwd"]["kernel_ms_note"] = ("CUDA events around each launch of the one-node step: pose_fwd, blend_tcgen05 "
                                               "(forward GEMM), skin (= skin_fit_l2: skinning + loss + gradient + skinning backward), "
                                               "dA, blend_bwd (backward GEMM), pose_bwd")
        best_fb = min(fb_ms, fb_graph_ms) if fb_graph_ms else fb_ms
        fb_tflops = 2 * GEMM_FLOPS_PER_BODY * Bb / (best_fb * 1e-3) / 1e12
        fb_gbs = 167384 * Bb / (best_fb * 1e-3) / 1e9
        extras["roofline_fwd_bwd"] = {
            "kernel": "fitting step (config 3, batch %d): %d launches" % (Bb, len(fbk)), "bound": "tensor",
            "achieved": fb_tflops, "peak": dense_peak / 3.0, "unit": "TFLOP/s", "frac": fb_tflops / (dense_peak / 3.0),
            "traffic": None, "ms_per_step": best_fb,
            "hbm": {"achieved": fb_gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": fb_gbs / peaks["hbm"]},
            "note": "achieved = 2 GEMMs x 2*B*20670*475 fp32-equivalent FLOPs / step time (CUDA graph replay when captured); "
                    "peak = dense tensor peak / 3 passes; hbm = 167,384 algorithmic B/body (SURVEY 8d fwd+bwd fused minimum) / step time"}

        # ---- batch-1 fitting closure (what lib/Gen_SMPLH/fit_single_frame.py runs: batch_size == 1)
        try:
            b1, p1, t1 = (torch.tensor(x, device=dev, requires_grad=True) for x in synthetic.make_inputs(model, 1, seed=5))
            tgt1 = torch.randn(1, dm.V, 3, device=dev)

            def closure():
                for t_ in (b1, p1, t1):
                    t_.grad = None
                fit_vertex_l2(dm, b1, p1, tgt1, transl=t1, reduce="sum").backward()

            def timeit(fn, n=50):
                for _ in range(5):
                    fn()
                torch.cuda.synchronize(dev)
                q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                q0.record(stream)
                for _ in range(n):
                    fn()
                q1.record(stream)
                torch.cuda.synchronize(dev)
                return q0.elapsed_time(q1) / n
            with torch.no_grad():
                fwd1 = timeit(lambda: body_model_apply(dm, b1.detach(), p1.detach(), transl=t1.detach()))
            eager1 = timeit(closure)
            gs = torch.cuda.Stream(dev)
            gs.wait_stream(stream)
            with torch.cuda.stream(gs):
                for _ in range(3):
                    closure()
            stream.wait_stream(gs)
            torch.cuda.synchronize(dev)
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                closure()
            graph1 = timeit(g1.replay)
            extras["batch1_fitting_closure"] = {
                "forward_ms": fwd1, "fwd_bwd_ms_eager": eager1, "fwd_bwd_ms_cuda_graph": graph1,
                "note": "one body: forward, vertex-L2 loss, backward w.r.t. betas / pose / transl; eager = "
                        "launched from Python (launch-bound), cuda_graph = the same closure captured once and replayed"}
        except Exception as e:
            sys.stderr.write("batch-1 closure timing skipped: %r\n" % (e,))

        # ---- BASELINE configs[0]: the numpy twin, one body per call, host numpy in / out (drop-in for
        # models/smpl_np.py SMPLModel.set_params, which takes 16 ms per call on the build container's
        # CPU: profiles/r01_reference_cpu_timing.json)
        try:
            from smplk import SMPLModel
            twin = SMPLModel(synthetic.make_model("smpl", num_betas=10, seed=8), device=local)
            rng1 = np.random.default_rng(0)
            tp_, tb_, tt_ = rng1.standard_normal((24, 3)) * 0.3, rng1.standard_normal(10), rng1.standard_normal(3)
            for _ in range(5):
                twin.set_params(pose=tp_.copy(), beta=tb_.copy(), trans=tt_.copy())
            t0 = time.perf_counter()
            for _ in range(200):
                twin.set_params(pose=tp_.copy(), beta=tb_.copy(), trans=tt_.copy())
            extras["config1_numpy_twin_batch1"] = {
                "ms_per_call": (time.perf_counter() - t0) / 200 * 1e3,
                "note": "smplk.SMPLModel.set_params(pose, beta, trans) -> (6890,3) numpy array: H2D, pose kernel, "
                        "tcgen05 blend, skinning, D2H, host wall clock"}
        except Exception as e:
            sys.stderr.write("config-1 twin timing skipped: %r\n" % (e,))

        def timed_ms(fn, reps, warm=2):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize(dev)
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record(stream)
            for _ in range(reps):
                fn()
            q1.record(stream)
            torch.cuda.synchronize(dev)
            return q0.elapsed_time(q1) / reps

        def fwd_call(m_, b_, p_, t_, v_, ws_, j_=None):
            a = _lib.ForwardArgs()
            a.batch, a.flags = p_.shape[0], 0
            a.betas, a.betas_batch = (ctypes.c_void_p(b_.data_ptr()) if b_ is not None else None), (b_.shape[0] if b_ is not None else 1)
            a.pose, a.transl = ctypes.c_void_p(p_.data_ptr()), ctypes.c_void_p(t_.data_ptr())
            a.verts = ctypes.c_void_p(v_.data_ptr())
            a.joints = ctypes.c_void_p(j_.data_ptr()) if j_ is not None else None
            a.workspace, a.workspace_bytes = ctypes.c_void_p(ws_.data_ptr()), ws_.numel()
            a.stream = ctypes.c_void_p(stream.cuda_stream)
            m_.forward(a)

        # ---- the operand format north_star names (3xTF32) next to the fp16 two-term default
        if rank == 0:
            try:
                dm_t = smplk.DeviceModel(model, device=local, options={"blend_tf32": 1})
                wst = torch.empty(dm_t.workspace_bytes(B, 0), device=dev, dtype=torch.uint8)
                dm_t.profile_enable(True)
                for i in range(10):
                    fwd_call(dm_t, *sets[i % NSETS], verts, wst, joints)
                torch.cuda.synchronize(dev)
                pt = {k: v[0] / v[1] for k, v in dm_t.profile_read(reset=True).items() if v[1] > 0}
                dm_t.profile_enable(False)
                g_ms = pt.get("blend_tcgen05")
                extras["blend_tf32_variant"] = {
                    "kernel_ms": pt, "forward_ms": sum(pt.values()),
                    "blend_gemm_tflops_fp32_equiv": (GEMM_FLOPS_PER_BODY * B / (g_ms * 1e-3) / 1e12) if g_ms else None,
                    "frac_of_tf32_ceiling": (GEMM_FLOPS_PER_BODY * B / (g_ms * 1e-3) / 1e12) / ((tf32_peak or peaks["bf16"] / 2) / 3.0) if g_ms else None,
                    "note": "handle option blend_tf32: 3xTF32 blend GEMM (tcgen05 kind::tf32) + skinning kernel, same batch; the default is "
                            "the fp16 hi+lo split fused with the skinning epilogue (equal 11+11 mantissa bits, half the tensor time)"}
                del dm_t, wst
            except Exception as e:
                sys.stderr.write("tf32 variant skipped: %r\n" % (e,))

        # ---- BASELINE config 4: one point of the batch sweep, 64K bodies on this GPU (vertices of the whole
        # slice resident: 5.4 GB), forward in 8192-body chunks through a bounded workspace
        try:
            n4 = 1 << 16
            gen = torch.Generator(device=dev).manual_seed(100 + rank)
            b4 = torch.randn(n4, 16, device=dev, generator=gen)
            p4 = torch.randn(n4, 156, device=dev, generator=gen) * 0.3
            t4 = torch.randn(n4, 3, device=dev, generator=gen)
            v4 = torch.empty(n4, dm.V, 3, device=dev)
            j4 = torch.empty(n4, dm.J, 3, device=dev)
            ws4 = torch.empty(dm.workspace_bytes(n4, 0), device=dev, dtype=torch.uint8)
            ms4 = timed_ms(lambda: fwd_call(dm, b4, p4, t4, v4, ws4, j4), 5)
            if world > 1:
                t = torch.tensor([ms4], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms4 = float(t.item())
            extras["config4_64k_per_gpu"] = {"metric": METRIC, "bodies_per_gpu": n4, "n_gpus": world, "ms": ms4,
                                             "value": world * n4 / (ms4 * 1e-3), "unit": UNIT,
                                             "note": "BASELINE configs[3], the 64K-bodies-per-GPU point; tools/sweep.py runs the whole sweep"}
            del b4, p4, t4, v4, j4, ws4
            torch.cuda.empty_cache()
        except Exception as e:
            sys.stderr.write("config-4 point skipped: %r\n" % (e,))

        # ---- BASELINE config 5: 100k-frame sequence tiled from the AMASS clip fixture, one betas row, no render
        if rank == 0:
            try:
                from smplk import clips
                clip = clips.read_amsass(os.path.join(ROOT, "tests", "golden", "amass_clip_09_05.npz"), full=True)
                N5, n0 = 100000, clip.poses.shape[0]
                reps5 = N5 // n0 + 1
                pose5 = torch.tensor(np.tile(clip.poses, (reps5, 1))[:N5].astype(np.float32), device=dev)
                tr5 = torch.tensor(np.tile(clip.trans, (reps5, 1))[:N5].astype(np.float32), device=dev)
                be5 = torch.tensor(clip.betas.reshape(1, 16).astype(np.float32), device=dev)
                ck = 16384                     # frames whose vertices are resident at once (1.35 GB)
                v5 = torch.empty(ck, dm.V, 3, device=dev)
                ws5 = torch.empty(dm.workspace_bytes(ck, 0), device=dev, dtype=torch.uint8)

                def run_smplh():
                    for f0 in range(0, N5, ck):
                        f1 = min(N5, f0 + ck)
                        fwd_call(dm, be5, pose5[f0:f1], tr5[f0:f1], v5, ws5)
                ms5 = timed_ms(run_smplh, 3, warm=1)
                c5 = {"frames": N5, "smplh": {"ms": ms5, "value": N5 / (ms5 * 1e-3), "unit": "frames/s",
                                              "note": "full SMPL-H forward (hands, 16 betas broadcast); each %d-frame chunk overwrites the last" % ck},
                      "lbs_only": []}
                del v5, ws5
                p24 = pose5[:, :72].contiguous()
                p24.view(N5, 24, 3)[:, [13, 14, 22, 23]] = 0.0           # lib/model2video.py:44-45
                for nv in (6890, 50000, 200000):
                    rig = synthetic.make_rigged_mesh(nv, seed=13)
                    rdm = smplk.DeviceModel(rig, device=local, lbs_only=True)
                    # the whole sequence in ONE call when its vertices fit (8.3 GB at 6,890 vertices, 60 GB at 50,000;
                    # the library walks it in 8192-frame chunks), else chunks of ~8 GB whose output is overwritten
                    ckr = N5 if N5 * nv * 12 <= 64e9 else max(256, min(16384, int(8e9 // (nv * 12))))
                    vr = torch.empty(ckr, nv, 3, device=dev)
                    wsr = torch.empty(rdm.workspace_bytes(ckr, 0), device=dev, dtype=torch.uint8)

                    def run_rig():
                        for f0 in range(0, N5, ckr):
                            f1 = min(N5, f0 + ckr)
                            fwd_call(rdm, None, p24[f0:f1], tr5[f0:f1], vr, wsr)
                    msr = timed_ms(run_rig, 2, warm=1)
                    wgbs = N5 * nv * 12 / (msr * 1e-3) / 1e9
                    c5["lbs_only"].append({"verts": nv, "ms": msr, "value": N5 / (msr * 1e-3), "unit": "frames/s",
                                           "hbm_write_gbs": wgbs, "frac_of_hbm_peak": wgbs / peaks["hbm"],
                                           "kernel": "lbs_replay_gemm_kernel (tcgen05: verts = P . T^T, K = 4 J = 96, three fp16 two-term passes; "
                                                     "lane = vertex, 12-byte stores straight from the TMEM registers) + replay_operand_kernel",
                                           "note": "rigged-mesh replay (RecoverModel, 24 joints): the only HBM stream is the vertex write "
                                                   "(the P operand and the frames' transform rows are L2-resident); %d frames per call" % ckr})
                    del vr, wsr, rdm
                    torch.cuda.empty_cache()
                extras["config5_sequence_100k"] = c5
                del pose5, tr5, p24
                torch.cuda.empty_cache()
            except Exception as e:
                sys.stderr.write("config-5 timing skipped: %r\n" % (e,))

        # ---- e2e: C-ABI host-buffer call (pinned host memory, H2D + D2H inside the timed region)
        lib = smplk.load()
        hb, hp, ht = (torch.tensor(x).pin_memory() for x in synthetic.make_inputs(model, B, seed=7))
        hv = torch.empty(B, dm.V, 3).pin_memory()
        def e2e_step():
            _lib.check(lib.smplk_forward_host(dm.handle, B, 0, ctypes.c_void_p(hb.data_ptr()), B,
                                              ctypes.c_void_p(hp.data_ptr()), ctypes.c_void_p(ht.data_ptr()),
                                              ctypes.c_void_p(hv.data_ptr()), None,
                                              ctypes.c_void_p(stream.cuda_stream)))
        for _ in range(3):
            e2e_step()
        ne = max(5, min(args.steps, 20))
        barrier()
        t0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(ne):
            e2e_step()
        g1.record(stream)
        torch.cuda.synchronize(dev)
        e2e_ms = g0.elapsed_time(g1) / ne
        if world > 1:
            t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        # copy-only ceiling of the same step on this box, all ranks at once: ONE cudaMemcpyAsync per buffer
        # (inputs H2D, then the vertices D2H from a resident device buffer into the same pinned host buffer)
        def copy_only():
            for h_, d_ in ((hb, sets[0][0]), (hp, sets[0][1]), (ht, sets[0][2])):
                d_.copy_(h_, non_blocking=True)
            hv.copy_(verts, non_blocking=True)
        for _ in range(2):
            copy_only()
        barrier()
        c0_, c1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0_.record(stream)
        for _ in range(ne):
            copy_only()
        c1_.record(stream)
        torch.cuda.synchronize(dev)
        copy_ms = c0_.elapsed_time(c1_) / ne
        if world > 1:
            t = torch.tensor([copy_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            copy_ms = float(t.item())
        d2h_bytes = int(hv.numel()) * 4
        extras["e2e"] = {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT,
                         "h2d_bytes_per_step": int(hb.numel() + hp.numel() + ht.numel()) * 4,
                         "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms,
                         "api": "smplk_forward_host (C ABI, pinned host buffers; numpy-twin path): 1024-body chunks, "
                                "kernels on the caller's stream overlapped with the D2H of the previous chunk on a copy stream",
                         "copy_only_ms_per_step": copy_ms,
                         "copy_only_d2h_gbs_per_gpu": d2h_bytes / (copy_ms * 1e-3) / 1e9,
                         "frac_of_copy_ceiling": copy_ms / e2e_ms,
                         "numa": numa,
                         "note": "copy_only = the same buffers moved by plain cudaMemcpyAsync with no kernels, all ranks concurrently "
                                 "(max over ranks): the PCIe / host-memory ceiling of this box for the step"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (blend contraction on tensor cores with two-term split operands, %s; fp32 accumulate and fp32 everywhere else)" % ("3xTF32" if os.environ.get("SMPLK_BLEND") == "tf32" else "fp16 hi+lo"),
                "data": "synthetic",
                "config": workload_config(B, world),
                "clocks": clocks.summary(), "gpu_launches": int(launches),
                "hbm_gbs_fused_equiv": FWD_BYTES_FUSED * world * B * args.steps / (ms * 1e-3) / 1e9}
        line.update(extras)
        if "e2e" not in line:
            line["e2e"] = None
        if not args.no_cpu_baseline and not args.no_extras:
            line["cpu_baseline"] = cpu_baseline(model, 512)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
