#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark of the RetinaNet inference hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One JSON line on rank 0.  A "step" is one pass of the hot path over one batch of synthetic
input.  `value` is device-timed whole-job throughput with inputs resident in HBM; `e2e` is the
same metric through the public API with HOST (pinned) buffers and the host<->device copies
inside the timed region; `roofline` describes the dominant kernel (live CUDA-event timing on the
launching stream); `cpu_baseline` is the CPU oracle timed on this box's host cores (rank 0, N=1).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor": d["bf16_tflops"], "tensor_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm": 6650.0, "tensor": 1590.0, "tensor_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for (t, line) in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = float(f[1])
                if t0 - 0.05 <= t <= t1 + 0.05:
                    sm.append(float(f[0]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                pass
        if not sm:   # region shorter than the sampling period: use every sample we have
            for (t, line) in self.rows:
                try:
                    sm.append(float(line.split(",")[0]))
                except ValueError:
                    pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# =================================================================================================
# workload: decode + NMS only (BASELINE.json configs[1]: RN50FPN head shapes, batch 8 per GPU)
# =================================================================================================
class PostprocWorkload:
    name = "decode+nms only, ResNet50FPN head shapes 3x800x1280, 80 classes, 9 anchors (BASELINE configs[1])"
    metric = "decode+NMS images/sec (3x800x1280 head outputs, fp32 NCHW entry point)"
    dtype = "f32"

    def __init__(self, batch, rank, device):
        import torch
        from retinanet_examples_b200 import box, synth
        self.torch, self.batch, self.device = torch, batch, device
        self.box = box
        cls, deltas = synth.head_outputs(batch, seed=rank)
        self.host = [(c.pin_memory(), d.pin_memory()) for c, d in zip(cls, deltas)]
        self.dev = [(c.to(device), d.to(device)) for c, d in self.host]
        self.anchors = [box.generate_anchors(s, box.DEFAULT_RATIOS, box.DEFAULT_SCALES).reshape(-1).tolist()
                        for s in synth.LEVEL_STRIDES]
        self.strides = synth.LEVEL_STRIDES
        self.top_n, self.det = 1000, 100
        self.h2d_bytes = sum(c.numel() * 4 + d.numel() * 4 for c, d in self.host)
        self.d2h_bytes = batch * self.det * 6 * 4
        self.out = None
        self.host_out = torch.empty((batch, self.det, 6), dtype=torch.float32).pin_memory()
        self.launches_per_step = 3 + 1   # filter, gather, select+decode (all levels), nms
        # algorithmic bytes of the dominant kernel (score filter at P3): scores read once
        self.dominant = {"tag": 0, "name": "score_filter_kernel", "bound": "hbm"}
        self.level_score_bytes = [c.numel() * 4 for c, _ in self.host]

    def _run(self, tensors):
        from retinanet_examples_b200 import _C
        torch = self.torch
        scores, boxes, classes = _C.decode_levels([c for c, _ in tensors], [d for _, d in tensors], self.anchors,
                                                  self.strides, 0.05, self.top_n, False)
        return _C.nms(scores, boxes, classes, 0.5, self.det, False)

    def step(self):
        self.out = self._run(self.dev)

    def step_e2e(self):
        torch = self.torch
        dev = [(c.to(self.device, non_blocking=True), d.to(self.device, non_blocking=True)) for c, d in self.host]
        s, b, c = self._run(dev)
        packed = torch.cat([s[..., None], b, c[..., None]], dim=2)
        self.host_out.copy_(packed, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def units_per_step(self):
        return self.batch

    def roofline(self, lib, peaks):
        import ctypes
        ms, n = ctypes.c_double(0), ctypes.c_longlong(0)
        lib.odtk_prof_get(self.dominant["tag"], ctypes.byref(ms), ctypes.byref(n))
        if n.value == 0:
            return None
        # one launch per step covers all five levels; algorithmic bytes = every score read once
        bytes_per_launch = float(sum(self.level_score_bytes))
        avg_ms = ms.value / n.value
        achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        return {"kernel": self.dominant["name"], "bound": "hbm", "achieved": round(achieved, 1),
                "peak": peaks["hbm"], "unit": "GB/s", "frac": round(achieved / peaks["hbm"], 4),
                "traffic": None, "peak_source": peaks["source"] + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                "avg_launch_ms": round(avg_ms, 5), "launches_timed": n.value,
                "algorithmic_bytes_per_launch": int(bytes_per_launch),
                "note": "one launch per step streams the scores of all 5 levels of the batch"}

    # ---- CPU legs (oracle; rank 0 only) --------------------------------------------------------
    def cpu_once(self, nimg):
        import numpy as np
        from oracle import oracle
        outs = []
        t0 = time.perf_counter()
        for lvl, (c, d) in enumerate(self.host):
            outs.append(oracle.decode(c[:nimg].numpy(), d[:nimg].numpy(), np.asarray(self.anchors[lvl], np.float32),
                                      self.strides[lvl], 0.05, self.top_n))
        cat = [np.concatenate(t, 1) for t in zip(*outs)]
        oracle.nms(cat[0], cat[1], cat[2], 0.5, self.det)
        return time.perf_counter() - t0

    def cpu_baseline(self):
        nimg = min(self.batch, 4)
        self.cpu_once(1)
        dt = self.cpu_once(nimg)
        return {"value": round(nimg / dt, 3), "unit": "images/sec", "cores": 1, "kind": "port",
                "sample": "%d images of the same batch, all 5 levels, oracle/odtk_oracle.c decode+nms, 1 thread" % nimg}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="postproc", choices=["postproc"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    from retinanet_examples_b200 import _lib
    lib = _lib.lib()
    peaks = _peaks()

    wl = PostprocWorkload(args.batch, rank, device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile):
        lib.odtk_prof_reset()
        lib.odtk_prof_enable(1 if profile else 0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t1 = time.time()
        lib.odtk_prof_enable(0)
        ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    for _ in range(args.warmup):
        wl.step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms, t0, t1 = timed(wl.step, args.steps, profile=True)
    roof = wl.roofline(lib, peaks)
    clocks = sampler.stop(t0, t1) if rank == 0 else None

    for _ in range(2):
        wl.step_e2e()
    ms_e2e, _, _ = timed(wl.step_e2e, args.steps, profile=False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    units = wl.units_per_step() * world
    out = {
        "metric": wl.metric, "value": round(units * args.steps / (ms * 1e-3), 2), "unit": "images/sec",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": {"workload": wl.name, "images_per_gpu": args.batch, "global_batch": units, "top_n": 1000,
                   "detections": 100, "threshold": 0.05, "nms": 0.5, "l2": "inputs (%.0f MB per step per GPU) exceed the 126 MB L2"
                   % (wl.h2d_bytes / 1e6), "sharding": "image-wise, no data-path collective" if world > 1 else "single GPU"},
        "us_per_image": round(ms * 1e3 / args.steps / wl.units_per_step(), 3),
        "clocks": clocks,
        "e2e": {"value": round(units * args.steps / (ms_e2e * 1e-3), 2), "unit": "images/sec",
                "h2d_bytes_per_step": wl.h2d_bytes, "d2h_bytes_per_step": wl.d2h_bytes,
                "ms_per_step": round(ms_e2e / args.steps, 4)},
        "gpu_launches": wl.launches_per_step * args.steps,
        "roofline": roof,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = wl.cpu_baseline()
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def reference_arm(args, rank, world):
    """The reference's CPU implementation of the path on this box's host cores (the oracle port:
    /root/reference does not exist on the GPU box).  Rank 0 only; other ranks exit 0."""
    if rank != 0:
        return
    import torch
    wl = PostprocWorkload.__new__(PostprocWorkload)
    from retinanet_examples_b200 import box, synth
    nimg = 2
    cls, deltas = synth.head_outputs(nimg, seed=0)
    wl.torch, wl.batch, wl.box = torch, nimg, box
    wl.host = list(zip(cls, deltas))
    wl.anchors = [box.generate_anchors(s, box.DEFAULT_RATIOS, box.DEFAULT_SCALES).reshape(-1).tolist() for s in synth.LEVEL_STRIDES]
    wl.strides, wl.top_n, wl.det = synth.LEVEL_STRIDES, 1000, 100
    for _ in range(min(args.warmup, 1)):
        wl.cpu_once(nimg)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        wl.cpu_once(nimg)
    dt = time.perf_counter() - t0
    v = round(nimg * steps / dt, 3)
    sample = "%d steps x %d images, all 5 levels, oracle decode+nms (port of csrc/cuda semantics), 1 thread" % (steps, nimg)
    print(json.dumps({
        "impl": "reference", "metric": PostprocWorkload.metric, "value": v, "unit": "images/sec", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": round(dt / steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": PostprocWorkload.name, "images_per_step": nimg},
        "cpu_baseline": {"value": v, "unit": "images/sec", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


if __name__ == "__main__":
    main()
