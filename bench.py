#!/usr/bin/env python
"""Benchmark of the FastMOT per-frame hot path on B200 (BASELINE.json metric: frames/sec/stream @1080p, 200 tracks).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # CPU baseline arm (oracle port on the host cores)

A "step" is one `MOT.step(frame)` on the next 1920x1080 frame of a deterministic synthetic stream with 200 objects
(BASELINE.json configs[2]: YOLOv4-csp 640 letterbox + OSNet x1.0, KLT on, detector every 5th frame).  Weights are
synthetic (no trained weights offline), so the detector's OUTPUT rows are replaced by the scripted ground-truth boxes
AFTER the whole detector pipeline (letterbox, 177-layer conv stack, decode, DIoU-NMS) has run at full cost — random
weights cannot detect, and the tracker must see 200 tracks.  Everything else is real data flow: ReID crops come
from the frame, OSNet embeddings feed the association kernels.

`value`  : frames already resident in HBM (ring of distinct frames, 373 MB > L2) when the timed region starts.
`e2e`    : same steps through the public API with frames in pinned HOST memory (H2D of 6.2 MB inside every step,
           track boxes/ids read back every step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec/stream @1080p, 200 tracks"
N_OBJECTS = 200
WORKLOAD = ("configs[2]: single 1080p stream per GPU, YOLOv4-csp 640 letterbox + OSNet x1.0, KLT on, "
            "detector every 5th frame, 200 tracks")
FRAME_SKIP = 5


def _cfg():
    from types import SimpleNamespace as NS
    from fastmot_b200.config import default_tracker_cfg
    t = default_tracker_cfg()
    return dict(detector_type='YOLO', detector_frame_skip=FRAME_SKIP, class_ids=(0,),
                yolo_detector_cfg=NS(model='YOLOv4CSP', conf_thresh=0.25, nms_thresh=0.5, max_area=800000,
                                     min_aspect_ratio=1.2),
                feature_extractor_cfgs=(NS(model='OSNet10', batch_size=16),),
                tracker_cfg=NS(**t))


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_frames(seed, n):
    from fastmot_b200.synth import SyntheticScene
    # objects bounce inside their grid cell (+-16 px): the stream holds 200 separate tracks for any number of steps
    scene = SyntheticScene(N_OBJECTS, seed=seed, label=0, dropout_frames=(), bounce_radius=16)
    return scene, [scene.frame(t) for t in range(n)]


def det_override(scene):
    from fastmot_b200.detector import DET_DTYPE

    def f(t):
        tl, lb, cf, _ = scene.detections(t)
        d = np.zeros(len(tl), DET_DTYPE)
        d['tlbr'], d['label'], d['conf'] = tl, lb, cf
        return d.view(np.recarray)
    return f


def reduce_max_ms(ms, world, device):
    """Timing of a multi-rank run = max over ranks (one all_reduce on the given device; NCCL on GPUs, gloo in the
    CPU test)."""
    if world <= 1:
        return float(ms)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(ms)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fastmot_b200 import MOT, _lib
    from fastmot_b200 import engine as eng_mod
    from fastmot_b200.utils import Profiler
    _lib.require_device()
    K, W = args.steps, args.warmup
    total = W + K
    scene, frames = make_frames(rank, 2 * total + 1)     # value pass then e2e pass continue the same stream
    cfg = _cfg()
    mot = MOT(scene.size, detections_override=det_override(scene), **cfg)
    mot.reset(1 / 30.)
    mot.extractors[0]._engine(N_OBJECTS)      # build + calibrate the ReID engine outside the timed region
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # stage timers (CUDA events on the launching streams) for the roofline of the dominant kernels
    prof = eng_mod.enable_profiling()
    sampler = ClockSampler(local)
    sampler.start()          # nvidia-smi needs ~1 s before its first sample: start before the warm-up

    def timed(step_inputs):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for f in step_inputs:
            mot.step(f)
            n_vis = sum(1 for _ in mot.visible_tracks())
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = reduce_max_ms(e0.elapsed_time(e1), world, dev)
        return ms, wall, n_vis

    # ---- pass 1: frames resident in HBM ----
    dev_frames = [torch.as_tensor(f).to(dev) for f in frames[:total]]
    for f in dev_frames[:W]:
        mot.step(f)
    prof.reset()
    launches0 = _lib.launch_count()
    sampler.rows.clear()     # keep only samples taken during the timed region
    ms_dev, wall_dev, n_vis = timed(dev_frames[W:])
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    stage = prof.summary()
    del dev_frames
    # ---- pass 2: end to end from pinned host memory through the public API ----
    pinned = [torch.as_tensor(f).pin_memory() for f in frames[total:2 * total]]
    host_frames = [p.numpy() for p in pinned]
    for f in host_frames[:W]:
        mot.step(f)
    ms_e2e, wall_e2e, n_vis2 = timed(host_frames[W:])

    if rank == 0:
        peaks, peak_src = _peaks()
        fps = world * K / (ms_dev / 1e3)
        fps_e2e = world * K / (ms_e2e / 1e3)
        conv_ms = stage.get("yolo_ms", 0.0) + stage.get("osnet_ms", 0.0)
        conv_flops = stage.get("yolo_flops", 0.0) + stage.get("osnet_flops", 0.0)
        ach = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
        peak_bw = peaks.get("hbm_gbs")
        os_ms = stage.get("osnet_ms", 0.0)
        os_gbs = stage.get("osnet_bytes", 0.0) / (os_ms * 1e-3) / 1e9 if os_ms > 0 else 0.0
        # DRAM bytes of one OSNet forward from the committed ncu capture (dram__bytes_read + dram__bytes_write summed
        # over the forward's kernels); null when the capture is not in the tree
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_osnet_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("osnet_forward_dram_bytes")
        out = {
            "metric": METRIC, "value": round(fps, 2), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms_dev / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 conv (fp32 accumulate; depthwise 3x3 sums each window row in fp16, rows in fp32), "
                     "u8/fixed-point KLT, fp64 Kalman/assignment",
            "data": "synthetic 1920x1080 stream, 200 moving textured objects, synthetic (seeded, BN-calibrated) weights",
            "config": {"workload": WORKLOAD,
                       "streams": world, "parallelism": f"{world} independent streams, one per GPU, no collective",
                       "l2": "ring of distinct frames (6.2 MB each, > 126 MB L2 in total) — inputs larger than L2",
                       "detections": "scripted ground-truth boxes replace the detector output rows after the full "
                                     "detector pipeline ran (random weights cannot detect)",
                       "visible_tracks_last_step": int(n_vis), "conv_path": stage.get("conv_path"),
                       "yolo_candidates_last_frame": int(getattr(mot.detector, "last_num_candidates", -1))},
            "e2e": {"value": round(fps_e2e, 2), "unit": "frames/s", "h2d_bytes_per_step": int(frames[0].nbytes),
                    "d2h_bytes_per_step": int(n_vis2 * 33 + 128), "ms_per_step": round(ms_e2e / K, 4)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            # dominant stage of the step: the OSNet x1.0 stack on the 200 crops (1x1 tcgen05 convs + depthwise 3x3),
            # HBM bound: algorithmic bytes = every conv / depthwise layer's input + output + weights moved once
            "roofline": {"bound": "hbm", "achieved": round(os_gbs, 1), "peak": peak_bw, "unit": "GB/s",
                         "frac": round(os_gbs / peak_bw, 4) if peak_bw else None, "traffic": traffic,
                         "kernel": "OSNet x1.0 stack (conv_tc_kernel 1x1 + dwconv3_tile), batch 224 crops",
                         "peak_source": peak_src,
                         "bytes_per_launch": stage.get("osnet_bytes", 0.0) / max(stage.get("osnet_calls", 1), 1),
                         "ms_per_launch": stage.get("osnet_ms", 0.0) / max(stage.get("osnet_calls", 1), 1)},
            # the detector stack is the tensor-core view: batch-1 YOLOv4-csp 640 (117 convs of 20-80 us each)
            "roofline_tensor": {"bound": "tensor", "achieved": round(ach, 2), "peak": peak_tf, "unit": "TFLOP/s",
                                "frac": round(ach / peak_tf, 4) if peak_tf else None,
                                "kernel": "implicit-GEMM conv, both stacks (YOLOv4-csp + OSNet x1.0)",
                                "flops_per_detector_frame": conv_flops / max(stage.get("detector_frames", 1), 1),
                                "conv_ms_per_detector_frame": conv_ms / max(stage.get("detector_frames", 1), 1),
                                "yolo_tflops": round(stage.get("yolo_flops", 0.0) / max(stage.get("yolo_ms", 0.0), 1e-9)
                                                     / 1e9, 2)},
            "stages_ms_per_step": {k: round(v / K, 4) for k, v in stage.items() if k.endswith("_ms")},
            "wall_ms_per_step": round(wall_dev * 1e3 / K, 4),
            # reference stage names (mot.py:138-163), host wall clock per call over the whole run incl. warm-up
            "stage_wall_ms_per_call": {k: round(Profiler.get_avg_millis(k), 3)
                                       for k in ("track", "preproc", "detect", "extract", "assoc")},
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args.cpu_steps)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
def cpu_baseline(steps):
    """The oracle port of the same path on the host cores, bounded sample (kind "port": the reference is Python +
    TensorRT and cannot be installed here; its CPU tracker path is restated in oracle/ and pinned to it)."""
    import torch
    from oracle.pipeline import OraclePipeline
    scene, frames = make_frames(0, steps)
    tl0 = det_override(scene)

    def ov(t):
        d = tl0(t)
        return d.tlbr, d.label, d.conf
    pipe = OraclePipeline(scene.size, detections_override=ov)
    pipe.step(frames[0])            # frame 0 (init) is warm-up, like the GPU arm's warm-up steps
    pipe.stage_s.clear()
    t0 = time.perf_counter()
    for f in frames[1:]:
        pipe.step(f)
    dt = time.perf_counter() - t0
    n = len(frames) - 1
    trk_s = sum(pipe.stage_s.get(k, 0.0) for k in ("flow", "kalman", "assoc"))
    return {"value": round(n / dt, 4), "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
            "torch_threads": torch.get_num_threads(),
            "sample": f"frames 1..{n} of the same stream ({sum(1 for t in range(1, n + 1) if t % FRAME_SKIP == 0)} "
                      f"detector frames): conv stacks in fp32 PyTorch-CPU, KLT via OpenCV, numpy Kalman/assignment",
            "tracker_only_fps": round(n / trk_s, 3) if trk_s > 0 else None,
            "stages_s": {k: round(v, 3) for k, v in pipe.stage_s.items()}}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    steps = min(max(args.steps, FRAME_SKIP + 1), 16)
    cb = cpu_baseline(steps)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "frames/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)), "steps": steps - 1, "warmup": 1,
           "ms_per_step": round(1e3 / cb["value"], 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "fp32 conv, u8/fixed-point KLT (OpenCV), fp64 Kalman/assignment", "data": "synthetic",
           "config": {"workload": WORKLOAD, "arm": "reference CPU path (oracle port) on the host cores, bounded sample",
                      "sample": cb["sample"], "requested_steps": args.steps, "requested_warmup": args.warmup},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)   # 0.5 s timed region: enough nvidia-smi clock samples
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=11)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
