#!/usr/bin/env python3
"""bench.py — tracking frames/s of the orbx hot path on EuRoC-shaped synthetic stereo streams.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  A *step* is one pass of the hot path over one batch: S independent 752x480 stereo
streams advance one frame each (2*S images per GPU) through the whole per-frame chain — ORB extraction L+R,
ComputeStereoMatches, SearchByProjection(last frame), PoseOptimization, SearchByProjection(local map),
PoseOptimization (orbx_tracker_step).  Multi-GPU = replicas only (independent streams,
no data-path collective; NCCL is used for the barrier and the max-over-ranks reduction).

  value  : stereo frames/s, inputs resident in HBM (device API), CUDA-event timed on the launch stream
  e2e    : the same through the host-buffer C ABI (orbx_tracker_submit / orbx_tracker_collect: page-locked H2D of
           every image on a copy stream + D2H of every step's poses and statistics, all inside the timed region, two
           steps in flight), wall clock between synchronisations
  roofline: dominant kernel = the per-stage CUDA-event time measured live (C ABI stage timers)
  cpu_baseline: the CPU oracle (a port of the reference's CPU path; the reference itself cannot be
           built here) on a bounded sample, one thread

`--impl reference` times the CPU oracle on all host cores (one stream per thread).
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
sys.path.insert(0, os.path.join(ROOT, "awesome-orb-slam3-3dvisioncraft-version_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

W, H, NFEAT, NLEVELS = 752, 480, 1000, 8
METRIC = "tracking frames/sec (extract+match+poseopt) EuRoC 752x480"
UNIT = "stereo frames/s"
MONO = False
INERTIAL = False
IMU_KF_PERIOD = 10          # c3: every 10th step optimises against the last keyframe (mbMapUpdated), the others against the last frame
CHAIN_INERTIAL = ("ORB extraction L+R, ComputeStereoMatches, SearchByProjection(last frame), PoseOptimization, isInFrustum + "
                  "SearchByProjection(local map), PoseInertialOptimizationLastFrame (LastKeyFrame every 10th step)")
CHAIN_MONO = ("ORB extraction (one image), SearchByProjection(last frame, th 15), PoseOptimization (monocular edges), "
              "isInFrustum + SearchByProjection(local map), PoseOptimization")
CHAIN = ("ORB extraction L+R, ComputeStereoMatches, SearchByProjection(last frame), PoseOptimization, "
         "isInFrustum + SearchByProjection(local map), PoseOptimization")
# BASELINE.json configs[1] (the configuration the metric is quoted on) and configs[3]
CONFIGS = {
    "c3": dict(w=752, h=480, nfeat=1000, streams=444, n_map=1500, inertial=True,
               name="EuRoC V1_02-shaped stereo-inertial 752x480, 1000 feat, 8 levels (BASELINE config 3)"),
    "c1": dict(w=752, h=480, nfeat=1000, streams=888, n_map=1500, mono=True,
               name="EuRoC MH01-shaped monocular 752x480, 1000 feat, 8 levels (BASELINE config 1)"),
    # streams per GPU: a multiple of the 148 SMs (one PoseOptimization CTA per stream, two resident per SM); measured
    # 256 -> 42.0 k, 296 -> 43.8 k, 444 -> 44.8 k, 592 -> 44.3 k frames/s (profiles/r02v_streams_sweep.md)
    "c2": dict(w=752, h=480, nfeat=1000, streams=444, n_map=1500,
               name="EuRoC MH01-shaped stereo 752x480, 1000 feat, 8 levels"),
    "c4": dict(w=1920, h=1080, nfeat=2000, streams=32, n_map=3000,
               name="synthetic 1920x1080 stereo, 2000 feat, 8 levels (BASELINE config 4)"),
}
WORKLOAD = CONFIGS["c2"]["name"] + ": " + CHAIN
MAP_NOTE = ("SURVEY 8(d) map per stream: every keypoint as a MapPoint (descriptor with Binomial(mean 0..40) bit flips, pixel "
            "noise 0.6 px x scale, 20 % gross outliers displaced 3-6 px x scale) + ORBvoc distractors up to n_map points; "
            "60 % tracked by the last frame; a ring of 3 different image sets, the prior of step t+1 = dT x pose(step t) "
            "composed on the device")


def set_config(name):
    """Select the image geometry / feature budget the module-level helpers below work on."""
    global W, H, NFEAT, WORKLOAD, MONO, UNIT, INERTIAL
    c = CONFIGS[name]
    W, H, NFEAT = c["w"], c["h"], c["nfeat"]
    MONO = bool(c.get("mono"))
    INERTIAL = bool(c.get("inertial"))
    UNIT = "monocular frames/s" if MONO else "stereo frames/s"
    WORKLOAD = c["name"] + ": " + (CHAIN_MONO if MONO else CHAIN_INERTIAL if INERTIAL else CHAIN)
    return c


def level_sizes(w=None, h=None, nlevels=NLEVELS, sf=1.2):
    w, h = w or W, h or H
    out, s = [], np.float32(1.0)
    for l in range(nlevels):
        inv = np.float32(1.0) / s
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
        s = np.float32(float(s) * float(np.float32(sf)))
    return out


def algorithmic_bytes(k_per_image=None):
    k_per_image = k_per_image or NFEAT
    """SURVEY.md §8(d): per-image algorithmic bytes of the extractor, split per stage."""
    P = [w * h for w, h in level_sizes()]
    resize = sum(P[l - 1] + P[l] for l in range(1, len(P)))
    fast = sum(P)
    blur = 2 * sum(P)
    patch = k_per_image * (961 + 961)
    outb = k_per_image * 60
    return {"pyramid": resize, "fast": fast, "blur": blur, "quadtree": 0, "describe": patch + outb,
            "total": resize + fast + blur + patch + outb}


def make_streams(n_streams, seed0=100):
    """n_streams stereo pairs from a pool of distinct synthetic scenes (rolled copies beyond the pool)."""
    from orbx import synth
    pool = [synth.stereo_pair(seed0 + i, W, H) for i in range(min(n_streams, 12))]
    imgs = []
    for s in range(n_streams):
        l, r = pool[s % len(pool)]
        sh = (s // len(pool)) * 7
        imgs.append(np.roll(l, sh, axis=1))
        imgs.append(np.roll(r, sh, axis=1))
    return imgs  # [L0, R0, L1, R1, ...]


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def _rot(rng, deg):
    w = rng.normal(0, 1, 3)
    w = w / np.linalg.norm(w) * np.deg2rad(deg)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def make_poses(n_streams, seed=7):
    """Ground-truth pose per stream and the motion-model guess tracking starts from (0.4 deg, 1 cm off)."""
    rng = np.random.default_rng(seed)
    Tt = np.zeros((n_streams, 4, 4), np.float32)
    Tp = np.zeros((n_streams, 4, 4), np.float32)
    for s in range(n_streams):
        R, t = _rot(rng, rng.uniform(0.1, 10)), rng.uniform(-0.5, 0.5, 3)
        Rp = _rot(rng, 0.4)
        Tt[s] = np.eye(4)
        Tt[s][:3, :3], Tt[s][:3, 3] = R, t
        Tp[s] = np.eye(4)
        Tp[s][:3, :3], Tp[s][:3, 3] = Rp @ R, Rp @ t + rng.normal(0, 0.01, 3)
    return Tt, Tp


def make_sequence(n_streams, ring, seed=7):
    """A ring of `ring` frames per stream: true poses Tt[r][s], the relative motion dT[r][s] that takes the pose of frame
    r-1 to the motion-model guess of frame r (true relative motion x a 0.4 deg / 1 cm model error), and the absolute
    prior of the very first step."""
    rng = np.random.default_rng(seed)
    Tt = np.zeros((ring, n_streams, 4, 4), np.float64)
    for r in range(ring):
        for s in range(n_streams):
            Tt[r, s] = np.eye(4)
            Tt[r, s][:3, :3], Tt[r, s][:3, 3] = _rot(rng, rng.uniform(0.1, 10)), rng.uniform(-0.5, 0.5, 3)
    dT = np.zeros_like(Tt)
    for r in range(ring):
        for s in range(n_streams):
            E = np.eye(4)
            E[:3, :3], E[:3, 3] = _rot(rng, 0.4), rng.normal(0, 0.01, 3)
            dT[r, s] = E @ Tt[r, s] @ np.linalg.inv(Tt[(r - 1) % ring, s])
    T_init = Tt[ring - 1].copy()          # the pose "before" the first step is the last frame of the ring
    return Tt.astype(np.float32), dT.astype(np.float32), T_init.astype(np.float32)


def build_maps(frames, Tt, n_map, m_cap, seed0):
    """frames[s] = (kL, dL, uright, depth) of stream s -> stacked host arrays of orbx_track_map (tests/scenarios.py generator)."""
    import scenarios as sc
    maps = [sc.track_map_scenario(seed0 + s, f[0], f[1], f[2], f[3], Tt[s], m_cap=m_cap, n_map=min(n_map, m_cap), W=W, H=H)
            for s, f in enumerate(frames)]
    host = sc.stack_track_maps(maps)
    gt = {"true_points": float(np.mean([m["gt"]["n_true"] for m in maps])),
          "gross_outliers": float(np.mean([m["gt"]["is_outlier"].sum() for m in maps])),
          "mean_bit_flips": float(np.mean([m["gt"]["mean_flips"] for m in maps])),
          "map_points": int(min(n_map, m_cap))}
    return host, gt


# ---------------------------------------------------------------------------------------------------------------------
# CPU side (the ONLY places bench.py executes oracle/): cpu_baseline of the default run and `--impl reference`
# ---------------------------------------------------------------------------------------------------------------------
def cv2_anchor_ms(img, reps=3):
    """Per-image time of the OpenCV primitives the reference's extractor spends its time in (cv2 4.13, one thread):
    7 resizes, per-30-px-cell FAST at 20 with the fall-back to 7, 8 Gaussian blurs.  An ANCHOR for the oracle's speed, not
    the reference: quadtree, orientation and descriptors are not included."""
    try:
        import cv2
    except Exception:
        return None
    cv2.setNumThreads(1)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        pyr = [img]
        for l in range(1, NLEVELS):
            w, h = level_sizes(img.shape[1], img.shape[0])[l]
            pyr.append(cv2.resize(pyr[-1], (w, h), interpolation=cv2.INTER_LINEAR))
        t1 = time.perf_counter()
        f20, f7 = cv2.FastFeatureDetector_create(20, True), cv2.FastFeatureDetector_create(7, True)
        for im in pyr:
            h, w = im.shape
            minb, maxbx, maxby = 16, w - 16, h - 16
            nc, nr = int((maxbx - minb) / 30), int((maxby - minb) / 30)
            wc, hc = int(np.ceil((maxbx - minb) / nc)), int(np.ceil((maxby - minb) / nr))
            for i in range(nr):
                y0 = minb + i * hc
                if y0 >= maxby - 3:
                    continue
                y1 = min(y0 + hc + 6, maxby)
                for j in range(nc):
                    x0 = minb + j * wc
                    if x0 >= maxbx - 6:
                        continue
                    cell = im[y0:y1, x0:min(x0 + wc + 6, maxbx)]
                    if not f20.detect(cell):
                        f7.detect(cell)
        t2 = time.perf_counter()
        for im in pyr:
            cv2.GaussianBlur(im, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        t3 = time.perf_counter()
        cur = {"pyramid": 1e3 * (t1 - t0), "fast_cells": 1e3 * (t2 - t1), "blur": 1e3 * (t3 - t2), "sum": 1e3 * (t3 - t0)}
        if best is None or cur["sum"] < best["sum"]:
            best = cur
    return best


class _CpuStreams:
    """The CPU arm's inputs: a pool of stereo pairs, their poses and 8(d) maps (generated from the oracle's own features)."""

    def __init__(self, npool=8, n_map=1500):
        import oracle
        import scenarios as sc
        from orbx import abi
        self.oracle, self.cam = oracle, abi.make_camera()
        self.imgs = make_streams(npool)
        self.Tt, self.Tp = make_poses(npool)
        self.maps = []
        ex = (oracle.Extractor(NFEAT, 1.2, NLEVELS, 20, 7), oracle.Extractor(NFEAT, 1.2, NLEVELS, 20, 7))
        for j in range(npool):
            _, kL, dL, _ = ex[0](self.imgs[2 * j])
            _, kR, dR, _ = ex[1](self.imgs[2 * j + 1])
            ur, dp = oracle.stereo_match([ex[0].pyramid_level(l) for l in range(NLEVELS)], [ex[1].pyramid_level(l) for l in range(NLEVELS)],
                                         kL, dL, kR, dR, ex[0].scale, ex[0].inv_scale, sc.BF, sc.BF / sc.FX)
            self.maps.append(sc.track_map_scenario(900 + j, kL, dL, ur, dp, self.Tt[j], n_map=n_map, W=W, H=H))
        self.imu = [[sc.track_imu_scenario(7000 + 10 * j + m, self.Tt[j], m) for m in (1, 2)] for j in range(npool)] if INERTIAL else None
        self.n = npool

    def frame(self, j, extractors, two_threads=None):
        from replay_reference import track_frame_map
        i, j = j, j % self.n
        mode = 0
        if INERTIAL:
            mode = 1 if i % IMU_KF_PERIOD == 0 else 2
        return track_frame_map(self.oracle, self.cam, self.imgs[2 * j], self.imgs[2 * j + 1], self.maps[j], self.Tp[j],
                               nfeatures=NFEAT, extractors=extractors, mono=MONO, imu=self.imu[j][mode - 1] if mode else None,
                               imu_mode=mode)


def cpu_baseline(n_frames, warm=20):
    """The CPU oracle chain on ONE thread over a bounded sample: per-frame latencies after `warm` warm-up frames, the
    oracle's extraction time next to the cv2 anchor, and a 2-thread variant that extracts left and right concurrently as
    the reference's Frame constructor does (src/Frame.cc:111-114)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    cs = _CpuStreams()
    ex = (oracle.Extractor(NFEAT, 1.2, NLEVELS, 20, 7), oracle.Extractor(NFEAT, 1.2, NLEVELS, 20, 7))
    for i in range(warm):
        cs.frame(i, ex)
    lat = []
    t_all = time.perf_counter()
    for i in range(n_frames):
        t0 = time.perf_counter()
        cs.frame(i, ex)
        lat.append(time.perf_counter() - t0)
    total = time.perf_counter() - t_all
    lat = np.array(lat)
    # extraction alone (oracle) vs the cv2 primitives it restates
    t0 = time.perf_counter()
    for i in range(10):
        ex[0](cs.imgs[(2 * i) % (2 * cs.n)])
    ext_ms = 1e3 * (time.perf_counter() - t0) / 10
    anchor = cv2_anchor_ms(cs.imgs[0])
    # reference-like: left and right extraction on two threads (ctypes releases the GIL), the rest on the caller's thread
    pool = ThreadPoolExecutor(2)
    m = min(n_frames, 40)
    t0 = time.perf_counter()
    for i in range(m):
        j = i % cs.n
        fl, fr = pool.submit(ex[0], cs.imgs[2 * j]), pool.submit(ex[1], cs.imgs[2 * j + 1])
        fl.result(), fr.result()
    two_thread_ext = (time.perf_counter() - t0) / m
    one_thread_ext = 2 * ext_ms * 1e-3
    if MONO:                                   # one image per frame: nothing to run side by side
        two_thread_ext = one_thread_ext = 0.0
    med = float(np.median(lat))
    pool.shutdown()
    return {"value": 1.0 / med, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d frames of the same chain and 8(d) map workload after %d warm-up frames, %.1f s, 1 thread "
                      "(host has %d cores); value = 1 / median frame time" % (n_frames, warm, total, os.cpu_count() or 0),
            "frame_ms": {"median": 1e3 * med, "p95": 1e3 * float(np.percentile(lat, 95)), "mean": 1e3 * float(lat.mean())},
            "oracle_extraction_ms_per_image": ext_ms,
            "cv2_anchor_ms_per_image": anchor,
            "oracle_over_cv2_primitives": (ext_ms / anchor["sum"]) if anchor else None,
            "reference_like_two_thread": {"value": 1.0 / (med - one_thread_ext + two_thread_ext), "unit": UNIT, "cores": 2,
                                          "note": "L/R extraction on two threads (src/Frame.cc:111-114), rest of the chain on one"},
            "note": "the oracle is a scalar port; the cv2 anchor shows how much faster OpenCV's own primitives are than the "
                    "oracle's (divide GPU/CPU ratios by oracle_over_cv2_primitives for an OpenCV-backed estimate)"}


def _ref_worker(job):
    tid, count = job
    import oracle
    g = _ref_worker.__dict__
    if "cs" not in g:
        g["cs"] = _CpuStreams()
        g["ex"] = (oracle.Extractor(NFEAT, 1.2, NLEVELS, 20, 7), oracle.Extractor(NFEAT, 1.2, NLEVELS, 20, 7))
    for i in range(count):
        g["cs"].frame(tid * 5 + i, g["ex"])
    return count


def run_reference(args):
    """The reference arm: the reference's CMake build cannot run here (OpenCV 3 / Eigen / Boost / Pangolin are not
    installed; oracle/_ref holds the extractor, matcher and DBoW2 compiled from the reference sources but g2o needs
    Eigen, so the chain cannot be closed with reference code), so this times the CPU oracle — the restatement pinned to
    those sources — on all host cores, one independent stream per process, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    set_config(args.config)
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    per_proc = 2
    frames_per_step = procs * per_proc
    import oracle
    oracle.lib()   # build once before forking
    with mp.get_context("fork").Pool(procs) as pool:
        jobs = [(t, per_proc) for t in range(procs)]
        for _ in range(max(args.warmup, 1)):
            pool.map(_ref_worker, jobs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_worker, jobs)
        t_total = time.perf_counter() - t0
    n_total = frames_per_step * args.steps
    value = n_total / t_total
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8+f64", "data": "synthetic", "impl": "reference",
           "config": {"workload": WORKLOAD, "map": MAP_NOTE, "sample": "CPU oracle, %d frames/step" % frames_per_step},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                            "sample": "%d stereo frames per step x %d steps on %d processes (host has %d cores)"
                                      % (frames_per_step, args.steps, procs, cores)},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def reduce_over_ranks(dist, times_ms, device=None):
    """Replicas only: the job's time for each timed region is the MAX over ranks."""
    import torch
    t = torch.tensor(times_ms, dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def run_dry(args):
    """Multi-rank plumbing on CPU (gloo): same rendezvous / barrier / reduction / JSON path as the GPU run, with the
    GPU step replaced by a deterministic synthetic timing (rank r takes (10 + r) ms per step)."""
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    d = None
    if world > 1:
        dist.init_process_group("gloo")
        d = dist
        dist.barrier()
    dev_ms, e2e_ms = reduce_over_ranks(d, [(10.0 + rank) * args.steps, (20.0 + rank) * 3])
    if d:
        dist.barrier()
    frames = args.streams * world
    if rank == 0:
        print(json.dumps({"metric": METRIC, "value": frames * args.steps / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "u8+f64", "data": "synthetic",
                          "config": {"workload": "dry-run of the multi-rank plumbing (no GPU work)",
                                     "streams_per_gpu": args.streams, "parallelism": "replicas x%d" % world},
                          "e2e": {"value": frames * 3 / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0, "dry_run": True}))
    if d:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="orbx", choices=["orbx", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="c1: EuRoC 752x480 monocular (BASELINE config 1); "
                    "c2: EuRoC 752x480 stereo (the metric's configuration); c3: stereo-inertial (BASELINE config 3); "
                    "c4: 1920x1080 stereo, 2000 features (BASELINE config 4)")
    ap.add_argument("--streams", type=int, default=0, help="independent stereo streams per GPU per step (0: the config's default)")
    ap.add_argument("--workload", default="sec8d", choices=["sec8d", "selfmap"],
                    help="sec8d: SURVEY 8(d) map per stream, temporal pose chain; selfmap: the round-1 best-case harness")
    ap.add_argument("--ring", type=int, default=3, help="different image sets cycled through (sec8d)")
    ap.add_argument("--kf-period", type=int, default=10, help="keyframe-rate work (10 x SearchForTriangulation + LocalBA per stream) "
                    "every this many steps in the with_keyframe_step region; 0 skips the region")
    ap.add_argument("--overlap", type=int, default=1,
                    help="1: pose/matching of step t overlap extraction of step t+1 (two streams, double-buffered)")
    ap.add_argument("--cpu-frames", type=int, default=200, help="stereo frames of the single-thread CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--dry-run", action="store_true",
                    help="CPU-only check of the multi-rank plumbing (gloo): rendezvous, barrier, MAX-over-ranks "
                         "reduction and rank-0 JSON with synthetic per-rank timings; no GPU work")
    args = ap.parse_args()
    cfg = set_config(args.config)
    if not args.streams:
        args.streams = cfg["streams"]
    if MONO and args.workload == "selfmap":
        raise SystemExit("bench.py: the monocular tracker has no self-map harness (no stereo depth); use --workload sec8d")
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    if args.dry_run:
        return run_dry(args)

    import torch
    import orbx
    import scenarios as sc
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the orbx path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        # exactly ONE JSON line on stdout: NCCL prints its version banner to stdout when its debug level is VERSION
        # (this image's default) -- send NCCL's log to stderr, and keep fd 1 pointed at stderr while the communicator is
        # created (process-group init + the first collective), whatever else chats during initialisation
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # NCCL_DEBUG_FILE is only honoured above VERSION
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    S = args.streams
    IPS = 1 if MONO else 2
    B = IPS * S
    RING = max(1, args.ring) if args.workload == "sec8d" else 1
    cam = orbx.make_camera()
    ctx = orbx.Context(local)
    ex = orbx.ORBextractor(ctx, NFEAT, 1.2, NLEVELS, 20, 7, max_w=W, max_h=H, max_batch=2 * S)
    trk = orbx.Tracker(ctx, ex, S, cam, th_frame=None, th_map=1.0, nnratio_map=0.8, mono=MONO)
    mcap = trk.map_capacity
    Tt, dT, T_init = make_sequence(S, RING, seed=7 + rank)
    Tt_abs, Tp_abs = make_poses(S, seed=7 + rank)     # the self-map harness: one true pose and one absolute prior per stream
    sets = []
    map_gt = None
    for r in range(RING):
        imgs = make_streams(S, seed0=100 + 1000 * rank + 37 * r)
        pin = orbx.host_array((B, H, W), np.uint8)          # page-locked inputs for the e2e arm
        pin[:] = np.stack(imgs[0::2] if MONO else imgs)     # monocular: the left images only (the right ones only build the map)
        entry = dict(pin=pin, imgs=[pin[i] for i in range(B)], d_img=torch.from_numpy(pin).cuda(),
                     Tt=Tt[r], dT=dT[r], d_true=torch.from_numpy(Tt[r].reshape(S, 16)).cuda(),
                     d_dT=torch.from_numpy(dT[r].reshape(S, 16)).cuda())
        if args.workload == "sec8d":
            # the map is built from THIS pipeline's own features (device extraction + device stereo matching, untimed)
            feats = ex.extract_batch(imgs)
            frames = []
            for s in range(S):
                (_, kL, dL), (_, kR, dR) = feats[2 * s], feats[2 * s + 1]
                ur, dp = orbx.stereo_match(ctx, ex, 2 * s, ex, 2 * s + 1, kL, dL, kR, dR, cam.bf, cam.b)
                frames.append((kL, dL, ur, dp))
            host, gt = build_maps(frames, Tt[r], cfg["n_map"], mcap, 5000 + 100 * r + 100000 * rank)
            map_gt = gt
            pm = {}
            for name, dtp in orbx.abi.TrackMap.FIELDS:      # page-locked copies for the e2e arm, device copies for the resident arm
                a = orbx.host_array(host[name].shape, dtp)
                a[...] = host[name]
                pm[name] = a
            entry["map_host"] = pm
            entry["map_dev_t"] = {k: torch.from_numpy(v).cuda() for k, v in pm.items()}
            entry["map_dev"] = {k: v.data_ptr() for k, v in entry["map_dev_t"].items()}
        if INERTIAL:
            # inertial inputs of both optimisers for this frame of every stream (tests/scenarios.track_imu_scenario)
            entry["imu_host"], entry["imu_dev_t"], entry["imu_dev"] = {}, {}, {}
            for m in (1, 2):
                himu = sc.stack_track_imu([sc.track_imu_scenario(90000 + 1000 * r + 10 * s + m + 100000 * rank, Tt[r][s], m) for s in range(S)])
                pin_imu = {}
                for k, v in himu.items():
                    a = orbx.host_array(v.shape, v.dtype)
                    a[...] = v
                    pin_imu[k] = a
                entry["imu_host"][m] = pin_imu
                entry["imu_dev_t"][m] = {k: torch.from_numpy(v).cuda() for k, v in pin_imu.items()}
                entry["imu_dev"][m] = {k: v.data_ptr() for k, v in entry["imu_dev_t"][m].items()}
        sets.append(entry)
    d_prior_abs = torch.from_numpy(Tp_abs.reshape(S, 16)).cuda()
    d_true_abs = torch.from_numpy(Tt_abs.reshape(S, 16)).cuda()
    d_init = torch.from_numpy(T_init.reshape(S, 16).copy()).cuda()
    d_out = torch.zeros((S, 16), dtype=torch.float32, device="cuda")
    d_stats = torch.zeros((S, 8), dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.ExternalStream(ex.stream, device=local)
    torch.cuda.synchronize()
    step_no = [0]

    def step_device():
        E = sets[step_no[0] % RING]
        if args.workload == "sec8d":
            trk.set_map(E["map_dev"])
            if INERTIAL:
                m = 1 if step_no[0] % IMU_KF_PERIOD == 0 else 2
                trk.set_inertial(m, E["imu_dev"][m])
            trk.step_device(E["d_img"].data_ptr(), W, H, W, E["d_true"].data_ptr(), E["d_dT"].data_ptr(), d_out.data_ptr(),
                            d_stats.data_ptr())
        else:
            trk.step_device(E["d_img"].data_ptr(), W, H, W, d_true_abs.data_ptr(), d_prior_abs.data_ptr(), d_out.data_ptr(),
                            d_stats.data_ptr())
        step_no[0] += 1

    def restart_chain():
        trk.synchronize()
        step_no[0] = 0
        if args.workload == "sec8d":
            trk.set_chain(True, d_init.data_ptr())

    def timed_region(nsteps, extra_streams=()):
        """EXACTLY nsteps steps back to back, CUDA-event timed on the launch stream, ending when every stream is done."""
        main = torch.cuda.current_stream()
        rstream = torch.cuda.ExternalStream(trk.result_stream, device=local)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        start = torch.cuda.Event(enable_timing=True)
        start.record(main)
        stream.wait_event(start)
        for _ in range(nsteps):
            step_device()
        ends = []
        for st in (stream, rstream) + tuple(extra_streams):
            e = torch.cuda.Event(enable_timing=True)
            e.record(st)
            ends.append(e)
        for e in ends:
            e.synchronize()
        ms = max(start.elapsed_time(e) for e in ends)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        return ms

    # ---------------- resident arm: inputs already in HBM ----------------
    restart_chain()
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    # pass 1 — serial steps (each synchronised, L2 flushed in between) with the C ABI's stage timers on: this is
    # where the per-kernel durations for the roofline come from (a kernel timed without anything overlapping it)
    ex.set_profiling(True)
    trk.set_profiling(True)
    ext_sum = np.zeros(len(ex.STAGES))
    trk_sum = np.zeros(len(trk.STAGES))
    serial_ms = 0.0
    main_stream = torch.cuda.current_stream()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between iterations (not timed)
        start = torch.cuda.Event(enable_timing=True)
        start.record(main_stream)
        stream.wait_event(start)
        step_device()
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        e.synchronize()
        serial_ms += start.elapsed_time(e)
        ext_sum += ex.stage_ms()[0]
        trk_sum += trk.stage_ms()
    ex.set_profiling(False)
    trk.set_profiling(False)
    torch.cuda.synchronize()
    # pass 2 — the timed region: EXACTLY `steps` steps back to back.  With --overlap (default) the tracker runs
    # extraction+stereo of step t+1 on one CUDA stream while matching+pose optimisation of step t finish on a
    # second one (double-buffered); every step re-reads its 2*S images and rebuilds the whole pyramid, far more than the
    # 126 MB L2, so no explicit flush is needed inside the region.
    trk.set_overlap(bool(args.overlap))
    restart_chain()
    for _ in range(2):
        step_device()
    trk.synchronize()
    launches0 = ctx.launches
    dev_ms = timed_region(args.steps)
    launches = ctx.launches - launches0
    stats = d_stats.cpu().numpy()
    Tout = d_out.cpu().numpy().reshape(-1, 4, 4)
    last_true = sets[(step_no[0] - 1) % RING]["Tt"] if args.workload == "sec8d" else Tt_abs
    pose_err = float(np.abs(Tout[:, :3, 3] - last_true[:, :3, 3]).max())

    # ---------------- the same with the keyframe-rate work of every stream beside it (BASELINE config 5) ----------------
    kf = None
    kf_ms = 0.0
    if args.kf_period > 0:
        kps, desc = sc.synthetic_keypoints(1, NFEAT, W, H)
        ur = np.where(np.arange(NFEAT) % 2 == 0, kps["x"] - 10.0, -1.0).astype(np.float32)
        pairs = []
        for q in range(10):                      # nn = 10 covisible neighbours (src/LocalMapping.cc:506-509)
            t = sc.tri_scenario(200 + q, kps, desc, ur)
            pairs.append(dict(KF1=orbx.Frame(t["k1"], t["d1"], t["ur1"], bounds=(0, 0, W, H)), KF2=orbx.Frame(t["k2"], t["d2"], t["ur2"], bounds=(0, 0, W, H)),
                              has1=t["has1"], has2=t["has2"], fv1=t["fv1"], fv2=t["fv2"], cam1=cam, cam2=cam, R1w=t["R1w"], t1w=t["t1w"],
                              R2w=t["R2w"], t2w=t["t2w"]))
        lbas = [sc.lba_scenario(i, K=20, M=3000, n_fixed=3) for i in range(4)]
        tri_b = orbx.TriangulationBatch(ctx, [pairs[q % 10] for q in range(10 * S)], t["sigma2"], t["scaleFactors"], True)
        lba_b = orbx.LocalBABatch(ctx, [lbas[p % 4] for p in range(S)], cam)
        trk.set_keyframe_work(tri_b, lba_b, args.kf_period)
        restart_chain()
        for _ in range(args.kf_period):          # warm-up incl. one keyframe round
            step_device()
        trk.synchronize()
        kstream = torch.cuda.ExternalStream(trk.keyframe_stream, device=local)
        kf_steps = 2 * args.kf_period
        runs0 = trk.keyframe_runs
        kf_ms = timed_region(kf_steps, (kstream,))
        kf = dict(steps=kf_steps, period=args.kf_period, rounds=int(trk.keyframe_runs - runs0),
                  work_per_round="per stream: 10 x SearchForTriangulation (1000 features) + 1 LocalBundleAdjustment "
                                 "(20 KF / 3000 MP / %d observations), prepared plans on a third, low-priority stream" % len(lbas[0]["e_kf"]),
                  lba_pool_GB=lba_b.device_bytes / 1e9)
        trk.set_keyframe_work(None, None, 0)

    # ---------------- continuity with round 1: the best-case self-map harness, resident ----------------
    self_ms = None
    if args.workload == "sec8d" and not MONO:
        trk.synchronize()
        trk.set_chain(False)
        trk.set_map(None)
        trk.set_inertial(0)
        wl = args.workload
        args.workload = "selfmap"
        for _ in range(3):
            step_device()
        trk.synchronize()
        self_steps = min(args.steps, 10)
        self_ms = timed_region(self_steps)
        self_stats = d_stats.cpu().numpy()
        args.workload = wl

    # ---------------- e2e arm: host buffers through the C ABI ----------------
    # Every step copies its 2*S images AND its flattened local map from page-locked host memory to the device and reads
    # its poses + statistics back, all inside the timed region, through orbx_tracker_upload_map + orbx_tracker_submit /
    # orbx_tracker_collect: the H2D of step t+1 runs on a copy stream under the kernels of step t, and matching + pose
    # optimisation of step t overlap the extraction of step t+1, so two steps are in flight.
    e2e_steps = max(3, min(args.steps, 40))      # >= 2 map uploads at the default 20 steps; 10-step samples were too noisy
    e2e_ms, e2e_same = 0.0, None
    if not args.no_e2e:
        trk.set_overlap(True)
        for E in sets:
            E["prep"] = orbx.prepare_images(E["imgs"])
        eno = [0]
        map_period = max(1, args.kf_period) if args.kf_period > 0 else 10

        def e2e_submit():
            E = sets[eno[0] % RING]
            if args.workload == "sec8d":
                # the local map is state that lives beside the tracker and changes at keyframe rate (Tracking::UpdateLocalMap
                # rebuilds the point LIST per frame, but MapPoint positions / descriptors only change when LocalMapping
                # inserts a keyframe): its flat arrays cross PCIe every map_period-th step, the steps in between track
                # against the device-resident copy
                if eno[0] % map_period == 0:
                    trk.upload_map(E["map_host"])
                else:
                    trk.set_map(E["map_dev"])
                if INERTIAL:   # the IMU pre-integration of the frame crosses PCIe with it, every step
                    m = 1 if eno[0] % IMU_KF_PERIOD == 0 else 2
                    trk.upload_inertial(m, E["imu_host"][m])
                trk.submit(E["prep"], E["Tt"], E["dT"])
            else:
                trk.submit(E["prep"], Tt_abs, Tp_abs)
            eno[0] += 1

        def e2e_run(nsteps):
            last = None
            e2e_submit()
            for _ in range(nsteps - 1):
                e2e_submit()
                last = trk.collect()
            return trk.collect() if nsteps else last

        trk.synchronize()
        if args.workload == "sec8d":
            trk.set_chain(True, d_init.data_ptr())
        e2e_run(RING)
        trk.synchronize()
        if args.workload == "sec8d":
            trk.set_chain(True, d_init.data_ptr())
        eno[0] = 0
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_last = e2e_run(e2e_steps)
        torch.cuda.synchronize()
        e2e_ms = 1e3 * (time.perf_counter() - t0)
        # same inputs as resident step number e2e_steps of a fresh chain -> compare against a resident replay
        trk.synchronize()
        restart_chain()
        for _ in range(e2e_steps):
            step_device()
        trk.synchronize()
        e2e_same = bool(np.array_equal(e2e_last[1], d_stats.cpu().numpy()))
    map_period = max(1, args.kf_period) if args.kf_period > 0 else 10
    map_uploads = len([k for k in range(e2e_steps) if k % map_period == 0]) if args.workload == "sec8d" else 0
    map_bytes = trk.map_bytes * map_uploads // max(e2e_steps, 1)     # amortised over the timed steps
    imu_bytes = S * (12 + 24 + 8 * (21 + 16 + 45 + 6 + 81 + 9 + 9 + 21 + 225)) + 128 if INERTIAL else 0   # mode-2 upload (the common step)
    h2d = B * W * H + 2 * S * 64 + map_bytes + imu_bytes
    d2h = S * 64 + S * 8 * 4
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- reduce over ranks (max time) ----------------
    dev_ms, kf_ms, e2e_ms, self_ms_r = reduce_over_ranks(dist, [dev_ms, kf_ms, e2e_ms, self_ms or 0.0], device="cuda")
    frames = S * world
    value = frames * args.steps / (dev_ms / 1e3)
    e2e_value = frames * e2e_steps / (e2e_ms / 1e3) if e2e_ms > 0 else None

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (live CUDA-event stage times) ----------------
    ext_ms = ext_sum / args.steps          # kernels inside stage "extract"
    trk_ms = trk_sum / args.steps
    kp_mean = float(stats[:, 0].mean())
    alg = algorithmic_bytes(int(round(kp_mean)))
    kernel_ms = {("extract." + n): float(v) for n, v in zip(ex.STAGES, ext_ms)}
    kernel_ms.update({n: float(v) for n, v in zip(trk.STAGES[1:], trk_ms[1:])})
    # The roofline block describes the dominant kernel of the HBM-streaming part of the path, the extractor (north_star:
    # ">= 60 % HBM roofline on the extractor kernel").  The matcher / optimiser stages are latency-bound (ordered replay,
    # serial fp64 LM chains: DESIGN.md §4); they run on the second stream under the next step's extraction and are listed
    # with their times in stage_ms / latency_bound_stages instead of being given a meaningless HBM fraction.
    ext_names = ["extract." + n for n in ex.STAGES if alg.get(n, 0) > 0]
    dom_name = max(ext_names, key=kernel_ms.get)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    short = dom_name.split(".")[-1]
    n_launch = (NLEVELS - 1) if short == "pyramid" else 1
    dom_bytes = alg[short] * B
    achieved = dom_bytes / (kernel_ms[dom_name] * 1e-3) / 1e9 if kernel_ms[dom_name] > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and args.config == "c2":
        try:
            traffic = json.load(open(tp)).get(short)
        except Exception:
            traffic = None
    ext_total_ms = float(ext_ms.sum())
    per_stage = {}
    for n in ex.STAGES:
        if alg.get(n, 0) > 0 and kernel_ms["extract." + n] > 0:
            gbs = alg[n] * B / (kernel_ms["extract." + n] * 1e-3) / 1e9
            per_stage[n] = {"ms": kernel_ms["extract." + n], "GB/s": gbs, "frac": gbs / peak}
    roof = {"bound": "hbm", "kernel": dom_name, "launches_per_step": n_launch, "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
            "algorithmic_bytes_per_launch": dom_bytes // n_launch, "stage_ms": kernel_ms,
            "stage_ms_note": "CUDA-event time per stage in the serial, L2-flushed pass (nothing overlapping the kernel)",
            "extractor_stages": per_stage,
            "latency_bound_stages": {n: kernel_ms[n] for n in trk.STAGES[1:]},
            "longest_stage_overall": max(kernel_ms, key=kernel_ms.get),
            "extractor_total": {"achieved": alg["total"] * B / (ext_total_ms * 1e-3) / 1e9,
                                "frac": alg["total"] * B / (ext_total_ms * 1e-3) / 1e9 / peak,
                                "bytes_per_image": alg["total"], "ms": ext_total_ms}}

    cpu = None
    if not args.no_cpu and world == 1 and args.config in ("c1", "c2", "c3"):
        cpu = cpu_baseline(args.cpu_frames)

    mean_stats = {n: float(v) for n, v in zip(trk.STATS, stats.mean(0))}
    mean_stats["outliers_1"] = mean_stats["matches_frame"] - mean_stats["inliers_1"]
    cfg_out = {"workload": WORKLOAD, "config": args.config, "streams_per_gpu": S, "images_per_step_per_gpu": B,
               "parallelism": "replicas x%d" % world,
               "l2": "inputs larger than L2: each step streams %d*S fresh images (%d MB) and rebuilds %d MB of pyramid; a ring of %d "
                     "different image sets" % (IPS, B * W * H >> 20, int(B * sum(w * h for w, h in level_sizes()) * 2) >> 20, RING),
               "overlap_steps": bool(args.overlap), "ms_per_step_serial_flushed": serial_ms / args.steps,
               "mean_per_stream": mean_stats, "max_translation_error_m": pose_err}
    if args.workload == "sec8d":
        cfg_out["map"] = MAP_NOTE
        cfg_out["map_generator"] = map_gt
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8+f64", "data": "synthetic", "config": cfg_out,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "steps": e2e_steps, "api": "orbx_tracker_submit/collect (+ orbx_tracker_upload_map every %d-th step: %d map uploads of %d bytes "
                                              "inside the timed region)" % (map_period, map_uploads, trk.map_bytes if args.workload == "sec8d" else 0),
                   "results_equal_resident_arm": e2e_same},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    if kf:
        kf["value"] = frames * kf["steps"] / (kf_ms / 1e3)
        kf["unit"] = UNIT
        kf["ms_per_step"] = kf_ms / kf["steps"]
        out["with_keyframe_step"] = kf
    if self_ms:
        out["selfmap_round1_workload"] = {"value": frames * min(args.steps, 10) / (self_ms_r / 1e3), "unit": UNIT,
                                          "ms_per_step": self_ms_r / min(args.steps, 10),
                                          "mean_per_stream": {n: float(v) for n, v in zip(trk.STATS, self_stats.mean(0))},
                                          "note": "round-1 harness (map = the frame's own stereo points and descriptors), kept for continuity"}
    print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
