#!/usr/bin/env python
"""bench.py -- frames/s of tag36h11 detect + per-tag pose on 1080p frames (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's B200 path
  python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU reference arm (oracle port, all host threads)
  torchrun ... bench.py --gpus N ...                              # one rank per GPU, frames sharded, no collective

Workload (default) = BASELINE.json configs[2] ("C3"): 1920x1080 gray frames, batch 1024 per GPU, quad_decimate = 1,
refine_edges = 1, ~50 tag36h11 tags per frame.  A step = one pass of detect+pose over the batch.
--config C1|C2|C4|C5 selects the other named shapes of BASELINE.json (same JSON line, same keys).
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

METRIC = "frames/s (tag36h11 detect+pose, 1080p)"
TAG36 = (("tag36h11", range(587)),)
MIXED = (("tag25h9", range(35)), ("tagStandard41h12", range(5)))
# BASELINE.json configs[0..4] (C1..C5).  batch = frames per GPU per step.
CONFIGS = {
    "C1": dict(W=640, H=480, batch=256, per_call=1, decimate=2.0, families="tag36h11", scene="sim", cap=16, tag_size=10.0,
               label="C1: 640x480 gray, batch 1 per call (256 calls per step), reference defaults (quad_decimate=2, "
                     "refine_edges=1), config/sim_settings.json scene with tag36h11 textures, detect + per-tag pose"),
    "C2": dict(W=1280, H=720, batch=256, decimate=2.0, families="tag36h11", famspec=TAG36, grid=(5, 2), cap=64, tag_size=0.2,
               label="C2: 1280x720 gray, batch 256 per GPU, quad_decimate=2, ~10 tag36h11/frame, detect + per-tag pose"),
    "C3": dict(W=1920, H=1080, batch=1024, decimate=1.0, families="tag36h11", famspec=TAG36, grid=(10, 5), cap=64, tag_size=0.2,
               label="C3: 1920x1080 gray, batch 1024 per GPU, quad_decimate=1, refine_edges=1, ~50 tag36h11/frame, "
                     "detect + per-tag pose"),
    "C4": dict(W=3840, H=2160, batch=64, decimate=2.0, families="tag36h11", famspec=TAG36, grid=(20, 10), cap=256, tag_size=0.2,
               label="C4: 3840x2160 gray, batch 64 per GPU (camera streams sharded over the GPUs), quad_decimate=2, "
                     "~200 tag36h11/frame, detect + per-tag pose"),
    "C5": dict(W=1920, H=1080, batch=256, decimate=1.0, families="tag25h9 tagStandard41h12", famspec=MIXED, grid=(10, 5),
               cap=64, tag_size=0.2, augment=True,
               label="C5: 1920x1080 gray, batch 256 per GPU, quad_decimate=1, mixed tag25h9 + tagStandard41h12 (ids 0-4), "
                     "seeded blur / gain / illumination ramp / sensor noise (synth.augment), detect + per-tag pose"),
}
METRICS = {"C3": METRIC}
W, H = 1920, 1080       # (C3, the reference arm's workload)
GRID = (10, 5)
TAG_SIZE = 0.2
CAP = 64
NOMINAL_PEAK_GBS = 8000.0   # BASELINE.json north_star: "B200 peak (~8 TB/s)"


def metric_of(cfg_name):
    return METRIC if cfg_name == "C3" else "frames/s (detect+pose, %s)" % cfg_name


def csrc_sha():
    """Hash of the kernel sources: ties profiles/ncu_traffic.json (an ncu capture) to the code it was taken from."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "aprilslam_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def _render_one(i):
    from aprilslam_b200 import synth
    return synth.render(synth.grid_scene(W, H, i, GRID))


def _augment_one(a):
    from aprilslam_b200 import synth
    return synth.augment(a[0], a[1])

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
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML in-process (a
    sample every few milliseconds: the timed region of a default run is well under a second, less than nvidia-smi needs
    to start up); `nvidia-smi -lms` as the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"),
               (0x80, "hw_power_brake_slowdown"))

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []      # (sm MHz, reasons bit mask)
        self.stop_flag = False

    def _nvml_sample(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        try:
            rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        self.samples.append((float(sm), int(rs)))

    def _nvml_loop(self):
        while not self.stop_flag:
            try:
                self._nvml_sample()
            except Exception:
                break
            time.sleep(0.004)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
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
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            sm = [s[0] for s in self.samples]
            mask = 0
            for _, r in self.samples:
                mask |= r
            reasons = sorted(name for bit, name in self.REASONS if mask & bit)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_run(frames: np.ndarray, K, steps: int, warmup: int, threads: int, fams: str = "tag36h11",
                      decimate: float = 1.0, tag_size: float = TAG_SIZE, cap: int = CAP):
    """The reference's CPU path restated: detector (oracle, frames spread over `threads` host threads) +
    cv2.solvePnP per detection exactly as tag_detector.py:30-43 calls it, the per-frame pose loops spread over the same
    number of threads (cv2 releases the GIL inside solvePnP).  Returns (seconds per step, tags per frame,
    seconds of the detector alone, seconds of the pose loop alone)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.binding import OracleDetector, reference_pose
    det = OracleDetector(fams, decimate=decimate, refine_edges=True)
    dist = np.zeros((4, 1))

    def poses_of(recs):
        return [reference_pose(r["p"], K, dist, tag_size) for r in recs]

    times, t_det, t_pose, ndet = [], [], [], 0
    with ThreadPoolExecutor(max_workers=max(1, threads)) as pool:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            lists = det.detect_batch(frames, nthreads=threads, cap=cap)
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


def build_frames(cfg, det, rank, B, torch):
    """-> (frames_dev uint8 [B,H,W] CUDA, scenes or None, K).  Every frame is a distinct seeded scene (SURVEY.md 8d:
    default_rng(1000 + global frame index)), rendered straight into HBM by the GPU restatement of the reference's
    renderer (bit-identical to synth.render); C5 adds the seeded augmentation on the host."""
    from aprilslam_b200 import synth
    from aprilslam_b200.render import render_batch
    Wc, Hc = cfg["W"], cfg["H"]
    if cfg.get("scene") == "sim":
        # C1: the reference's own scene (config/sim_settings.json) at the webcam path's 640x480
        # (video_detection.py:106-107), camera poses = seeded Monte-Carlo in the engine's bounds (simulation_engine.py:92)
        rng = np.random.default_rng(rank)
        scenes = []
        for _ in range(B):
            cam = (rng.uniform(-15, 50), rng.uniform(-5, 5), rng.uniform(-1.25, 15))
            scenes.append(synth.sim_settings_scene(Wc, Hc, cam_pos=cam, family="tag36h11"))
    else:
        scenes = [synth.grid_scene(Wc, Hc, rank * B + i, cfg["grid"], families=cfg["famspec"]) for i in range(B)]
    frames_dev = render_batch(det, scenes)
    torch.cuda.synchronize()
    if cfg.get("augment"):
        from multiprocessing import Pool
        host = frames_dev.cpu().numpy()
        with Pool(min(os.cpu_count() or 1, 16)) as pool:
            host = np.stack(pool.map(_augment_one, [(host[i], 7000 + rank * B + i) for i in range(B)], chunksize=4))
        frames_dev = torch.from_numpy(host).to(frames_dev.device)
    return frames_dev, scenes, scenes[0].K


def check_against_scenes(cfg, dets, scenes):
    """Ids of EVERY frame of the timed batch against the scene that was rendered: no id that is not in the scene, and the
    fraction of the scene's tags that were found.  (Corner / pose parity against the oracle lives in tests/.)"""
    fams = cfg["families"].split()
    false_pos = found = total = 0
    for recs, sc in zip(dets, scenes):
        truth = {(t.family, t.tag_id) for t in sc.tags}
        got = {(fams[int(r["family"])], int(r["id"])) for r in recs}
        false_pos += len(got - truth)
        found += len(got & truth)
        total += len(truth)
    return {"frames_checked": len(scenes), "ids_not_in_scene": false_pos, "scene_tags_found_fraction": found / max(1, total)}


def run_b200(args):
    import torch
    import __graft_entry__ as ge
    from aprilslam_b200.detector import Detector

    cfg = CONFIGS[args.config]
    Wc, Hc, d = cfg["W"], cfg["H"], cfg["decimate"]
    cap, tag_size, fams = cfg["cap"], cfg["tag_size"], cfg["families"]
    per_call = cfg.get("per_call", 0)          # C1: batch 1 per call
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
    B = args.batch or cfg["batch"]
    det = Detector(fams, decimate=d, refine_edges=True, device=local_rank, chunk_frames=args.chunk,
                   pipeline_slots=args.slots)
    t_gen = time.time()
    frames_dev, scenes, K = build_frames(cfg, det, rank, B, torch)
    pinned = torch.empty((B, Hc, Wc), dtype=torch.uint8).pin_memory()
    pinned.copy_(frames_dev)
    frames_host = pinned.numpy()
    t_gen = time.time() - t_gen

    def run(dd, frames):
        if not per_call:
            return dd.detect_pose_batch(frames, K, None, tag_size, cap_per_frame=cap)
        out_d, out_p = [], []
        for i in range(0, B, per_call):        # C1: one synchronous call per frame, like the reference's loop
            a, b = dd.detect_pose_batch(frames[i:i + per_call], K, None, tag_size, cap_per_frame=cap)
            out_d.extend(np.array(x) for x in a)
            out_p.extend(np.array(x) for x in b)
        return out_d, out_p

    def step_dev():
        return run(det, frames_dev)

    def step_host():
        return run(det, frames_host)

    def barrier():
        if distributed:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, dd=None):
        """-> (seconds for `steps` steps by CUDA events, launches, per-stage ms, per-kernel {name: [ms, launches]}, last result)"""
        dd = dd or det
        stage, kern = {}, {}
        launches = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = None
        for _ in range(steps):
            res = fn()
            launches += dd.launch_count() * (B // per_call if per_call else 1)
            if not per_call:
                for k, v in dd.stage_ms().items():
                    stage[k] = stage.get(k, 0.0) + v
                for k, (ms, cnt) in dd.kernel_table().items():
                    e = kern.setdefault(k, [0.0, 0])
                    e[0] += ms
                    e[1] += cnt
        e1.record()
        barrier()
        sec = e0.elapsed_time(e1) / 1e3
        if distributed:
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            sec = float(t.item())
        return sec, launches, stage, kern, res

    for _ in range(max(args.warmup, 3)):
        step_dev()
    sampler = ClockSampler(phys)
    if rank == 0:
        sampler.start()
    sec, launches, _, _, res = timed(step_dev, args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    dets, poses = res
    tags_per_frame = float(np.mean([len(x) for x in dets]))
    pose_ok = float(np.mean([p["ok"].mean() if len(p) else 1.0 for p in poses]))
    checked = check_against_scenes(cfg, dets, scenes)
    counters = det.counters()
    # end to end: host (pinned) frames through the same public call
    step_host()
    e2e_steps = max(1, min(args.steps, 3))
    sec_e2e, _, _, _, _ = timed(step_host, e2e_steps)
    # N > 1: the host-side gather SURVEY 8(e) names -- rank 0 collects every rank's per-frame lists in frame order over the
    # process group (NCCL here), timed together with the device-resident step it follows
    gather = None
    if distributed:
        from aprilslam_b200._lib import DET_DTYPE
        from aprilslam_b200.shard import gather_lists

        def step_gather():
            dd, pp = step_dev()
            return dd, gather_lists([np.asarray(x) for x in dd], DET_DTYPE)

        step_gather()
        sec_g, _, _, _, (own, allists) = timed(step_gather, e2e_steps)
        ok_g = None
        if rank == 0:
            ok_g = len(allists) == B * world and all(
                np.array_equal(allists[i]["id"], np.asarray(own[i])["id"]) for i in range(0, B, max(1, B // 64)))
        gather = {"value": B * world * e2e_steps / sec_g, "unit": "frames/s", "lists_on_rank0": B * world,
                  "rank0_lists_in_frame_order": ok_g,
                  "bytes_gathered_per_step": int(sum(len(x) for x in own) * DET_DTYPE.itemsize * world),
                  "note": "device-resident step + shard.gather_lists (two all_gathers: sizes, then the ragged payload) on the "
                          "bench's process group; max over ranks"}
    # end to end from BGR frames, the reference's real input (tag_detector.py:25): 3 bytes per pixel over the host link
    e2e_bgr = None
    if not per_call and not args.no_bgr:
        nb = min(B, max(1, int(3.3e9 // (3 * Wc * Hc))))
        bgr_pinned = torch.empty((nb, Hc, Wc, 3), dtype=torch.uint8).pin_memory()
        bgr_pinned.copy_(frames_dev[:nb].unsqueeze(-1).expand(nb, Hc, Wc, 3))
        bgr_host = bgr_pinned.numpy()

        def step_bgr():
            return det.detect_pose_batch(bgr_host, K, None, tag_size, cap_per_frame=cap, bgr=True)

        rb, _ = step_bgr()
        same = all(np.array_equal(rb[i]["id"], dets[i]["id"]) for i in range(nb))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e2e_steps):
            step_bgr()
        e1.record()
        barrier()
        sec_bgr = e0.elapsed_time(e1) / 1e3
        if distributed:
            t = torch.tensor([sec_bgr], dtype=torch.float64, device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            sec_bgr = float(t.item())
        e2e_bgr = {"value": nb * world * e2e_steps / sec_bgr, "unit": "frames/s", "frames_per_step": nb,
                   "h2d_bytes_per_step": int(nb * 3 * Wc * Hc), "h2d_gbs": nb * 3 * Wc * Hc * e2e_steps / sec_bgr / 1e9,
                   "ids_equal_gray_path": bool(same),
                   "note": "uint8 [B,H,W,3] BGR host frames (the reference's input, tag_detector.py:25), gray conversion on the GPU"}
        del bgr_pinned, bgr_host
    # per-stage and per-kernel times: the same batch through a detector with ONE chunk in flight and the quad-fit tiers
    # serialised, every kernel launch bracketed by its own pair of CUDA events on the library's stream
    stage, kern, sec_prof, prof_steps, tier = {}, {}, 0.0, 0, None
    if not per_call:
        det_prof = Detector(fams, decimate=d, refine_edges=True, device=local_rank, chunk_frames=args.chunk,
                            pipeline_slots=1)
        det_prof.set_profiling(True)

        def step_prof():
            return det_prof.detect_pose_batch(frames_dev, K, None, tag_size, cap_per_frame=cap)

        step_prof()
        prof_steps = max(1, min(args.steps, 2))
        sec_prof, _, stage, kern, _ = timed(step_prof, prof_steps, det_prof)
        tier = det_prof.tier_stats()
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
    N = Wc * Hc
    f = int(d)
    wd, hd = 1 + (Wc - 1) // f, 1 + (Hc - 1) // f
    Nd = wd * hd
    # algorithmic bytes per frame: image stages R_src + N_d (R_src = N at decimate 1, N/d otherwise);
    # CC: threshold N_d + labels 4 N_d; edge points: threshold N_d + labels 4 N_d; dense pipeline A_img + 10 N_d
    a_img = (N if f == 1 else N // f) + Nd
    alg_stage = {"image": a_img, "cc": 5 * Nd, "edges": 5 * Nd}
    a_pipe = a_img + 10 * Nd
    roofline, stages, kernels = None, {}, {}
    chunk_frames = min(B, det_chunk(cfg, args, B))
    nchunks = -(-B // max(1, chunk_frames))
    if not per_call:
        nsteps = prof_steps
        tot_ms = max(1e-9, sum(stage.values()))
        for k, ms in stage.items():
            ent = {"ms_per_step": ms / nsteps, "us_per_frame": ms * 1e3 / (B * nsteps), "share": ms / tot_ms}
            if k in alg_stage:
                gbs = alg_stage[k] * B * nsteps / (ms / 1e3) / 1e9
                ent.update({"algorithmic_bytes_per_frame": alg_stage[k], "achieved_gbs": gbs, "frac": gbs / peak,
                            "frac_of_nominal_8TBs": gbs / NOMINAL_PEAK_GBS})
            stages[k] = ent
        # per kernel: its OWN algorithmic bytes per frame (what the kernel has to read and write for the work it does),
        # None where the kernel is a gather / compute kernel without a byte floor
        tiers_pf = [r / (B * 1.0) for r in tier["records"]]
        own = {"k_decimate_threshold": a_img, "k_cc_local": Nd + Nd // 4, "k_cc_boundary": Nd // 4,
               "k_fit_quads<1>": 8 * tiers_pf[0], "k_fit_quads<2>": 8 * tiers_pf[1], "k_fit_quads<4>": 8 * tiers_pf[2],
               "k_sort_scatter": 16 * sum(tiers_pf), "k_edges": Nd // 4 + 8 * sum(tiers_pf)}   # (records of clusters >= 24 points)
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            tr = None
        sha = csrc_sha()
        tr_ok = bool(tr) and tr.get("csrc_sha") == sha and tr.get("config", "C3") == args.config
        ksum = max(1e-9, sum(v[0] for v in kern.values()))
        for name, (ms, cnt) in sorted(kern.items(), key=lambda kv: -kv[1][0]):
            ent = {"ms_per_step": ms / nsteps, "us_per_launch": ms * 1e3 / max(1, cnt), "launches_per_step": cnt // nsteps,
                   "share_of_kernel_time": ms / ksum}
            if own.get(name):
                gbs = own[name] * B * nsteps / (ms / 1e3) / 1e9
                ent.update({"algorithmic_bytes_per_frame": own[name], "achieved_gbs": gbs, "frac": gbs / peak,
                            "frac_of_nominal_8TBs": gbs / NOMINAL_PEAK_GBS})
            if tr_ok and name in tr.get("kernels", {}):
                e = tr["kernels"][name]
                per_launch = (e["dram_read_mb"] + e["dram_write_mb"]) * 1e6 / max(1, e["launches"]) / tr["frames_per_launch"] * chunk_frames
                ent["dram_traffic_per_launch"] = per_launch
                ent["dram_gbs"] = per_launch / (ent["us_per_launch"] * 1e-6) / 1e9
            kernels[name] = ent
        dom = max(kern, key=lambda k: kern[k][0])          # the longest kernel over ALL stages
        kd = kernels[dom]
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kd.get("achieved_gbs"), "peak": peak, "unit": "GB/s",
                    "frac": kd.get("frac"), "frac_of_nominal_8TBs": kd.get("frac_of_nominal_8TBs"),
                    "traffic": kd.get("dram_traffic_per_launch"), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": (own.get(dom) or 0) * chunk_frames or None,
                    "us_per_launch": kd["us_per_launch"], "launches_per_step": kd["launches_per_step"],
                    "share_of_kernel_time": kd["share_of_kernel_time"],
                    "traffic_source": ("profiles/ncu_traffic.json (ncu --set full, kernel sources %s = this build)" % sha) if tr_ok
                                      else "none: profiles/ncu_traffic.json was captured from other kernel sources or another config",
                    "note": "the longest kernel of the step by its own CUDA-event pair (pipeline_slots=1, quad-fit tiers "
                            "serialised). achieved = the kernel's own algorithmic bytes (for k_fit_quads: 8 B per edge-point "
                            "record of its clusters) / its time: an irregular, latency- and issue-bound kernel, far from the "
                            "HBM roof by construction; the bandwidth-bound kernels are in image_stage / kernels",
                    "image_stage": {"kernel": "k_decimate_threshold", "achieved": stages["image"]["achieved_gbs"],
                                    "frac": stages["image"]["frac"],
                                    "frac_of_nominal_8TBs": stages["image"]["frac_of_nominal_8TBs"],
                                    "algorithmic_bytes_per_frame": a_img,
                                    "traffic": kernels.get("k_decimate_threshold", {}).get("dram_traffic_per_launch")},
                    "cc_stage": {"kernels": ["k_cc_local", "k_cc_boundary", "k_cc_sizes", "k_cc_dense"],
                                 "achieved": stages["cc"]["achieved_gbs"], "frac": stages["cc"]["frac"],
                                 "algorithmic_bytes_per_frame": alg_stage["cc"]},
                    "edges_stage": {"kernel": "k_edges", "achieved": stages["edges"]["achieved_gbs"],
                                    "frac": stages["edges"]["frac"], "algorithmic_bytes_per_frame": alg_stage["edges"]},
                    "dense_pipeline": {"achieved": a_pipe * value / world / 1e9, "frac": a_pipe * value / world / 1e9 / peak,
                                       "frac_of_nominal_8TBs": a_pipe * value / world / 1e9 / NOMINAL_PEAK_GBS,
                                       "algorithmic_bytes_per_frame": a_pipe,
                                       "note": "A_pipe x the PIPELINED per-GPU frame rate (`value` / n_gpus)"},
                    "kernels": kernels}
    # CPU baseline on this box's host cores: a bounded sample of the same batch (the first frames; sized from a short
    # probe so that the sample is ~10 s of wall time whatever the shape costs on the CPU)
    cpu_baseline = None
    if rank == 0 and world == 1:
        os.sched_setaffinity(0, cpus_before)          # the CPU baseline may use every host core again
        threads = len(os.sched_getaffinity(0)) or 1
        t0 = time.time()
        kw = dict(fams=fams if isinstance(fams, str) else " ".join(fams), decimate=d, tag_size=tag_size, cap=cap)
        nprobe = min(B, threads)
        sec_probe, _, _, _ = cpu_reference_run(frames_host[:nprobe], K, 1, 0, threads, **kw)
        nref = int(max(nprobe, min(B, 256 if args.config == "C3" else 512, nprobe * 10.0 / max(sec_probe, 1e-3))))
        sec_cpu, _, sec_cpu_det, sec_cpu_pose = cpu_reference_run(frames_host[:nref], K, 1, 0, threads, **kw)
        n1 = min(4, B)
        sec_cpu1, _, _, _ = cpu_reference_run(frames_host[:n1], K, 1, 0, 1, **kw)
        cpu_baseline = {"value": nref / sec_cpu, "unit": "frames/s", "cores": threads, "kind": "port",
                        "sample": "frames 0..%d of the same batch, oracle detector on %d host threads (%.2f s) + cv2.solvePnP "
                                  "per tag on the same threads (%.2f s); 1 thread (the reference's setting, "
                                  "tag_detector.py:18): %.2f frames/s" % (nref - 1, threads, sec_cpu_det, sec_cpu_pose,
                                                                          n1 / sec_cpu1),
                        "one_thread_value": n1 / sec_cpu1, "seconds": time.time() - t0}
    line = {
        "metric": metric_of(args.config), "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
        "config": {"workload": cfg["label"], "name": args.config,
                   "batch_per_gpu": B, "distinct_frames": B, "tags_per_frame": tags_per_frame,
                   "pose_ok_fraction": pose_ok, "checked_against_scenes": checked,
                   "edge_points_per_frame": counters["edge_points"] / B, "clusters_per_frame": counters["clusters"] / B,
                   "clusters_over_upstream_size_limit": counters["oversize_clusters"],
                   "l2": "inputs (%.2f GB per GPU) larger than L2" % (B * N / 1e9) if B * N > 2.6e8 else
                         "inputs %.0f MB per GPU; every step re-reads all frames and re-writes all work buffers (> L2 in total)" % (B * N / 1e6),
                   "chunk_frames": chunk_frames, "pipeline_slots": args.slots or 3,
                   "stages_note": ("stage / kernel times measured with pipeline_slots=1, tiers serialised (%.1f frames/s in that mode)" % (
                       B * prof_steps / sec_prof)) if prof_steps else "batch-1 path: no per-stage table",
                   "parallelism": "frames sharded, %d rank(s), no collective" % world,
                   "cpus_bound_rank0": bound_cpus,
                   "frame_generation_s": t_gen},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(B * N),
                "d2h_bytes_per_step": int(B * cap * (168 + 136) + 4 * (16 + 22 * B)), "steps": e2e_steps,
                "h2d_gbs": B * N * e2e_steps / sec_e2e / 1e9,
                "note": "gray uint8 host (pinned) frames through the public call; the frames cross the host link once, chunk "
                        "copies overlap the kernels of other chunks"},
        "e2e_bgr": e2e_bgr,
        "gather": gather,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "stages": stages,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))
    if distributed:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


def det_chunk(cfg, args, B):
    """The library's automatic chunk size (aprilgpu.cu detect_run) for device-resident frames."""
    if args.chunk > 0:
        return args.chunk
    f = int(cfg["decimate"])
    wd, hd = 1 + (cfg["W"] - 1) // f, 1 + (cfg["H"] - 1) // f
    plane = ((wd + 15) // 16 * 16) * hd
    chunk = max(1, min(256, (256 << 20) // plane))
    if B >= 48:
        chunk = min(chunk, (B + 2) // 3)
    return min(chunk, B)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS), help="BASELINE.json configs[0..4] = C1..C5 (default C3)")
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step (0 = the config's: C3 1024)")
    ap.add_argument("--chunk", type=int, default=0, help="frames per pipeline pass (0 = library default)")
    ap.add_argument("--slots", type=int, default=0, help="chunks in flight (0 = library default: 3)")
    ap.add_argument("--no-bgr", action="store_true", help="skip the BGR end-to-end leg")
    ap.add_argument("--ref-frames", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
