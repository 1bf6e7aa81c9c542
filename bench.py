#!/usr/bin/env python
"""bench.py -- frames/s of tag36h11 detect + per-tag pose on 1080p frames (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's B200 path
  python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU reference arm (oracle port, all host threads)
  torchrun ... bench.py --gpus N ...                              # one rank per GPU, frames sharded, no collective

Workload = BASELINE.json configs[2] ("C3"): 1920x1080 gray frames, batch 1024 per GPU, quad_decimate = 1,
refine_edges = 1, ~50 tag36h11 tags per frame.  A step = one pass of detect+pose over the batch.
`value`  : device-resident uint8 [B,H,W] batch in -> host-visible detection + pose lists out.
`e2e`    : the same call with HOST (pinned) frames: H2D of the frames inside the timed region.
`roofline`: HBM fraction of the dominant stage (+ per-stage table), algorithmic bytes from SURVEY.md 8(d).
`cpu_baseline`: the restated CPU detector (oracle/) + cv2.solvePnP per tag on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

W, H = 1920, 1080
GRID = (10, 5)          # ~50 tags / frame
TAG_SIZE = 0.2
CAP = 64
METRIC = "frames/s (tag36h11 detect+pose, 1080p)"


def _render_one(i):
    from aprilslam_b200 import synth
    return synth.render(synth.grid_scene(W, H, i, GRID))


def make_frames(distinct: int, total: int, seed0: int = 0) -> np.ndarray:
    """`distinct` seeded synthetic frames (default_rng(1000 + frame)), replicated to `total`
    with a per-replica gray offset so that no two frames of the batch are byte-identical."""
    from multiprocessing import Pool
    nproc = min(os.cpu_count() or 1, 16, distinct)
    idx = [seed0 + i for i in range(distinct)]
    if nproc > 1:
        with Pool(nproc) as p:
            pool = p.map(_render_one, idx)
    else:
        pool = [_render_one(i) for i in idx]
    pool = np.stack(pool)
    out = np.empty((total, H, W), np.uint8)
    for r in range(0, total, distinct):
        n = min(distinct, total - r)
        off = (r // distinct) % 8
        out[r:r + n] = np.minimum(pool[:n].astype(np.int16) + off, 255).astype(np.uint8)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_run(frames: np.ndarray, K, steps: int, warmup: int, threads: int):
    """The reference's CPU path restated: detector (oracle, frames spread over `threads` host threads) +
    cv2.solvePnP per detection exactly as tag_detector.py:30-43 calls it, the per-frame pose loops spread over the same
    number of threads (cv2 releases the GIL inside solvePnP).  Returns (seconds per step, tags per frame,
    seconds of the detector alone, seconds of the pose loop alone)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.binding import OracleDetector, reference_pose
    det = OracleDetector("tag36h11", decimate=1.0, refine_edges=True)
    dist = np.zeros((4, 1))

    def poses_of(recs):
        return [reference_pose(r["p"], K, dist, TAG_SIZE) for r in recs]

    times, t_det, t_pose, ndet = [], [], [], 0
    with ThreadPoolExecutor(max_workers=max(1, threads)) as pool:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            lists = det.detect_batch(frames, nthreads=threads, cap=CAP)
            t1 = time.perf_counter()
            if threads > 1:
                list(pool.map(poses_of, lists))
            else:
                for recs in lists:
                    poses_of(recs)
            t2 = time.perf_counter()
            if it >= warmup:
                times.append(t2 - t0); t_det.append(t1 - t0); t_pose.append(t2 - t1)
                ndet += sum(len(x) for x in lists)
    return float(np.mean(times)), ndet / max(1, steps * len(frames)), float(np.mean(t_det)), float(np.mean(t_pose))


def run_reference_arm(args):
    """CPU arm: the oracle port of the reference's detector + the reference's own cv2.solvePnP call, all host threads.
    Loads nothing of the product: only oracle/ is built and dlopen-ed here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    from aprilslam_b200 import synth                      # (pure numpy scene / frame generator, no native code)
    threads = len(os.sched_getaffinity(0)) or 1
    nframes = args.ref_frames or max(16, min(128, 8 * threads))
    # the FIRST nframes frames of the GPU arm's batch: same seeds (default_rng(1000 + frame index)), all distinct
    frames = make_frames(nframes, nframes)
    K = synth.intrinsics(W, H, 45.0)
    sec, dpf, sec_det, sec_pose = cpu_reference_run(frames, K, args.steps, args.warmup, threads)
    n1 = min(4, nframes)
    sec1, _, _, _ = cpu_reference_run(frames[:n1], K, 1, 0, 1)
    val = nframes / sec
    sample = ("frames 0..%d of the workload's 1024-frame batch per step (same seeds as the GPU arm), %d host threads: "
              "detector %.2f s + cv2.solvePnP per tag %.2f s per step; 1 thread (the reference's own setting, "
              "tag_detector.py:18): %.2f frames/s" % (nframes - 1, threads, sec_det, sec_pose, n1 / sec1))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
            "config": {"workload": "C3: 1920x1080 gray, quad_decimate=1, refine_edges=1, ~50 tag36h11/frame "
                                   "(bounded sample of the 1024-frame batch)", "frames_per_step": nframes,
                       "tags_per_frame": dpf, "same_frames_as_gpu_arm": True},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample,
                             "one_thread_value": n1 / sec1},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "restated CPU detector (oracle/), not the upstream apriltag binary (unavailable offline)"}
    print(json.dumps(line))
    return 0


def run_b200(args):
    import torch
    import __graft_entry__ as ge
    from aprilslam_b200 import synth
    from aprilslam_b200.detector import Detector

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if args.gpus > 1 and not distributed:
        print("bench.py: --gpus %d needs torchrun (one rank per GPU); see the module docstring" % args.gpus,
              file=sys.stderr)
        return 2
    if rank == 0:
        ge.build()
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one process per GPU: stay on the CPU cores / NUMA node next to this GPU before the pinned frame buffers exist
    from aprilslam_b200.shard import bind_to_gpu_numa
    visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    phys = int(visible.split(",")[local_rank]) if visible and all(v.strip().isdigit() for v in visible.split(",")) else local_rank
    cpus_before = os.sched_getaffinity(0)
    bound_cpus = bind_to_gpu_numa(phys)
    B = args.batch
    K = synth.intrinsics(W, H, 45.0)
    det = Detector("tag36h11", decimate=1.0, refine_edges=True, device=local_rank, chunk_frames=args.chunk,
                   pipeline_slots=args.slots)
    t_gen = time.time()
    if args.distinct <= 0:
        # every frame of the batch is a distinct seeded scene (default_rng(1000 + global frame index)), rendered
        # straight into HBM by the GPU restatement of the reference's renderer (bit-identical to synth.render)
        from aprilslam_b200.render import render_batch
        scenes = [synth.grid_scene(W, H, rank * B + i, GRID) for i in range(B)]
        frames_dev = render_batch(det, scenes)
        torch.cuda.synchronize()
        pinned = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
        pinned.copy_(frames_dev)
        frames_host = pinned.numpy()
        distinct = B
    else:
        frames_host = make_frames(args.distinct, B, seed0=rank * args.distinct)
        pinned = torch.from_numpy(frames_host).pin_memory()
        frames_dev = pinned.to(dev, non_blocking=False)
        distinct = args.distinct
    t_gen = time.time() - t_gen

    def step_dev():
        return det.detect_pose_batch(frames_dev, K, None, TAG_SIZE, cap_per_frame=CAP)

    def step_host():
        return det.detect_pose_batch(pinned.numpy(), K, None, TAG_SIZE, cap_per_frame=CAP)

    def barrier():
        if distributed:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, d=None):
        """-> (seconds for `steps` steps by CUDA events, launches, per-stage ms summed, last result)"""
        d = d or det
        stage = {}
        launches = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = None
        for _ in range(steps):
            res = fn()
            launches += d.launch_count()
            for k, v in d.stage_ms().items():
                stage[k] = stage.get(k, 0.0) + v
        e1.record()
        barrier()
        sec = e0.elapsed_time(e1) / 1e3
        if distributed:
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            sec = float(t.item())
        return sec, launches, stage, res

    for _ in range(max(args.warmup, 3)):
        step_dev()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sec, launches, _, res = timed(step_dev, args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    dets, poses = res
    tags_per_frame = float(np.mean([len(x) for x in dets]))
    pose_ok = float(np.mean([p["ok"].mean() if len(p) else 1.0 for p in poses]))
    # end to end: host (pinned) frames through the same public call
    step_host()
    e2e_steps = max(1, min(args.steps, 3))
    sec_e2e, _, _, _ = timed(step_host, e2e_steps)
    # per-stage times for the roofline: the same batch through a detector with ONE chunk in flight (no overlap of
    # chunks, so every stage interval on the stream is that stage alone), CUDA events on the library's stream
    det_prof = Detector("tag36h11", decimate=1.0, refine_edges=True, device=local_rank, chunk_frames=args.chunk,
                        pipeline_slots=1)
    det_prof.set_profiling(True)

    def step_prof():
        return det_prof.detect_pose_batch(frames_dev, K, None, TAG_SIZE, cap_per_frame=CAP)

    step_prof()
    prof_steps = max(1, min(args.steps, 2))
    cc_local_ms = [0.0]

    def step_prof_k():
        r = step_prof()
        cc_local_ms[0] += det_prof.kernel_ms("k_cc_local")
        return r

    sec_prof, _, stage, _ = timed(step_prof_k, prof_steps, det_prof)
    det_prof.close()

    if rank != 0:
        if distributed:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        return 0

    frames_per_step = B * world
    value = frames_per_step * args.steps / sec
    e2e_value = frames_per_step * e2e_steps / sec_e2e
    peak, peak_src = measured_peak_gbs()
    N = W * H
    # algorithmic bytes per frame (decimate = 1): image stage R_src + N_d = 2N;
    # CC: read threshold N + write labels 4N = 5N; edges: threshold N + labels 4N = 5N; dense pipeline 12N
    alg = {"image": 2 * N, "cc": 5 * N, "edges": 5 * N}
    nsteps = prof_steps
    stages = {}
    for k, ms in stage.items():
        per_frame_us = ms * 1e3 / (B * nsteps)
        ent = {"ms_per_step": ms / nsteps, "us_per_frame": per_frame_us, "share": ms / max(1e-9, sum(stage.values()))}
        if k in alg:
            gbs = alg[k] * B * nsteps / (ms / 1e3) / 1e9
            ent.update({"algorithmic_bytes_per_frame": alg[k], "achieved_gbs": gbs, "frac": gbs / peak})
        stages[k] = ent
    pipe_ms = sum(v for k, v in stage.items() if k not in ("h2d",))
    pipe_gbs = 12 * N * B * nsteps / (pipe_ms / 1e3) / 1e9
    dom = max((k for k in stage if k in alg), key=lambda k: stage[k])
    chunk_frames = min(B, det_chunk(det, args))
    nchunks = -(-B // max(1, chunk_frames))
    kernels_per_stage = {"image": ["k_decimate_threshold<1,4>"],
                         "cc": ["k_cc_local", "k_cc_boundary<8>", "k_cc_sizes", "k_cc_dense"], "edges": ["k_edges<2>"]}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        tr = None

    def traffic_per_launch(ent):
        # DRAM bytes per launch from the committed ncu capture, scaled from its chunk size to this run's
        if not ent:
            return None
        return (ent["dram_read_mb"] + ent["dram_write_mb"]) * 1e6 / tr["frames_per_launch"] * chunk_frames

    # the roofline line is for ONE kernel: the dominant kernel of the dominant dense stage, timed by its own pair of
    # CUDA events on the library's stream (cc: k_cc_local; image / edges are single-kernel stages)
    if dom == "cc":
        k_ms = cc_local_ms[0]
        k_name = "k_cc_local"
        k_traffic = traffic_per_launch(tr["cc"]["dominant_kernel"]) if tr else None
    else:
        k_ms = stage[dom]
        k_name = kernels_per_stage[dom][0]
        k_traffic = traffic_per_launch(tr[dom]) if tr else None
    k_gbs = alg[dom] * B * nsteps / (k_ms / 1e3) / 1e9
    roofline = {"kernel": k_name, "stage": dom, "bound": "hbm", "achieved": k_gbs, "peak": peak, "unit": "GB/s",
                "frac": k_gbs / peak, "traffic": k_traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom] * chunk_frames,
                "launches_per_step": nchunks, "us_per_launch": k_ms * 1e3 / (nchunks * nsteps),
                "note": "achieved = SURVEY 8(d) algorithmic bytes of the %s stage (%d B/frame) / CUDA-event time of %s; the "
                        "kernel moves fewer bytes than that (traffic): no u32 label image is materialised" % (
                            dom, alg[dom], k_name),
                "stage_all_kernels": {"kernels": kernels_per_stage[dom], "achieved": stages[dom]["achieved_gbs"],
                                      "frac": stages[dom]["frac"],
                                      "traffic": traffic_per_launch(tr[dom]) if tr else None},
                "image_stage": {"kernel": "k_decimate_threshold<1,4>", "achieved": stages["image"]["achieved_gbs"],
                                "frac": stages["image"]["frac"], "algorithmic_bytes_per_frame": alg["image"],
                                "traffic": traffic_per_launch(tr["image"]) if tr else None},
                "edges_stage": {"kernel": "k_edges<2>", "achieved": stages["edges"]["achieved_gbs"],
                                "frac": stages["edges"]["frac"], "algorithmic_bytes_per_frame": alg["edges"]},
                "dense_pipeline": {"achieved": pipe_gbs, "frac": pipe_gbs / peak, "algorithmic_bytes_per_frame": 12 * N}}
    # CPU baseline on this box's host cores (bounded sample of the same workload)
    os.sched_setaffinity(0, cpus_before)          # the CPU baseline may use every host core again
    threads = len(os.sched_getaffinity(0)) or 1
    nref = max(16, min(256, 16 * threads, B))     # ~10-30 core-seconds of CPU work
    t0 = time.time()
    sec_cpu, _, sec_cpu_det, sec_cpu_pose = cpu_reference_run(frames_host[:nref], K, 1, 0, threads)
    sec_cpu1, _, _, _ = cpu_reference_run(frames_host[:4], K, 1, 0, 1)
    cpu_val = nref / sec_cpu
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
        "config": {"workload": "C3: 1920x1080 gray, batch %d per GPU, quad_decimate=1, refine_edges=1, "
                               "~50 tag36h11/frame, detect + per-tag pose" % B,
                   "batch_per_gpu": B, "distinct_frames": distinct, "tags_per_frame": tags_per_frame,
                   "pose_ok_fraction": pose_ok, "l2": "inputs (%.1f GB per GPU) larger than L2" % (B * N / 1e9),
                   "chunk_frames": det_chunk(det, args), "pipeline_slots": args.slots or 3,
                   "stages_note": "stage times / roofline measured with pipeline_slots=1 (%.1f frames/s in that mode)" % (
                       B * prof_steps / sec_prof),
                   "parallelism": "frames sharded, %d rank(s), no collective" % world,
                   "cpus_bound_rank0": bound_cpus,
                   "frame_generation_s": t_gen},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(B * N),
                "d2h_bytes_per_step": int(B * CAP * (168 + 136) + 4 * (16 + 22 * B)), "steps": e2e_steps,
                "h2d_gbs": B * N * e2e_steps / sec_e2e / 1e9,
                "note": "PCIe-bound: the frames cross the host link once, chunk copies overlap the kernels of other chunks"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "stages": stages,
        "cpu_baseline": {"value": cpu_val, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "frames 0..%d of the same batch, oracle detector on %d host threads (%.2f s) + cv2.solvePnP "
                                   "per tag on the same threads (%.2f s); 1 thread (the reference's setting, "
                                   "tag_detector.py:18): %.2f frames/s" % (nref - 1, threads, sec_cpu_det, sec_cpu_pose,
                                                                           4 / sec_cpu1),
                         "one_thread_value": 4 / sec_cpu1,
                         "seconds": time.time() - t0},
    }
    print(json.dumps(line))
    if distributed:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


def det_chunk(det, args):
    if args.chunk > 0:
        return args.chunk
    plane = ((W + 15) // 16 * 16) * H
    chunk = max(1, min(256, (256 << 20) // plane))
    if args.batch >= 48:
        chunk = min(chunk, (args.batch + 2) // 3)
    return min(chunk, args.batch)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="frames per GPU per step (BASELINE config C3: 1024)")
    ap.add_argument("--distinct", type=int, default=0,
                    help="0: every frame distinct, rendered on the GPU; N>0: N host-rendered frames replicated")
    ap.add_argument("--chunk", type=int, default=0, help="frames per pipeline pass (0 = library default)")
    ap.add_argument("--slots", type=int, default=0, help="chunks in flight (0 = library default: 3)")
    ap.add_argument("--ref-frames", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
