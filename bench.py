#!/usr/bin/env python3
"""bench.py — tracking frames/s of the orbx hot path on EuRoC-shaped synthetic stereo streams.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  A *step* is one pass of the hot path over one batch: S independent 752x480 stereo
streams, one stereo frame each (2*S images per GPU).  Multi-GPU = replicas only (independent streams,
no data-path collective; NCCL is used for the barrier and the max-over-ranks reduction).

  value  : stereo frames/s, inputs resident in HBM (device API), CUDA-event timed on the launch stream
  e2e    : the same through the host-buffer C ABI (pinned H2D of every image + D2H of keypoints and
           descriptors inside the timed region), wall clock between synchronisations
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

import numpy as np  # noqa: E402

W, H, NFEAT, NLEVELS = 752, 480, 1000, 8
METRIC = "tracking frames/sec (extract+match+poseopt) EuRoC 752x480"
UNIT = "stereo frames/s"


def level_sizes(w=W, h=H, nlevels=NLEVELS, sf=1.2):
    out, s = [], np.float32(1.0)
    for l in range(nlevels):
        inv = np.float32(1.0) / s
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
        s = np.float32(float(s) * float(np.float32(sf)))
    return out


def algorithmic_bytes(k_per_image=NFEAT):
    """per-image algorithmic bytes of the extractor, split per stage."""
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


def cpu_oracle_frames_per_s(n_frames, threads=1):
    """Time the CPU oracle on n_frames stereo frames (2 extractions each)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    imgs = make_streams(min(n_frames, 12))
    oracle.lib()

    def work(tid, count):
        ex = oracle.Extractor(NFEAT, 1.2, NLEVELS, 20, 7)
        for i in range(count):
            j = (tid * 5 + i) % (len(imgs) // 2)
            ex(imgs[2 * j])
            ex(imgs[2 * j + 1])
        return count

    per = [n_frames // threads + (1 if t < n_frames % threads else 0) for t in range(threads)]
    work(0, 1)  # warm caches / page in
    t0 = time.perf_counter()
    if threads == 1:
        work(0, per[0])
    else:
        with ThreadPoolExecutor(threads) as pool:
            list(pool.map(lambda a: work(*a), enumerate(per)))
    dt = time.perf_counter() - t0
    return n_frames / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    frames_per_step = max(threads, 16)
    for _ in range(args.warmup):
        cpu_oracle_frames_per_s(threads, threads)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        fps, dt = cpu_oracle_frames_per_s(frames_per_step, threads)
        t_total += dt
        n_total += frames_per_step
    value = n_total / t_total
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "impl": "reference",
           "config": {"workload": "EuRoC-shaped 752x480 stereo, 1000 feat, 8 levels: ORB extraction L+R "
                                  "(CPU oracle, %d frames/step)" % frames_per_step},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": "%d stereo frames per step x %d steps, %d threads" % (frames_per_step, args.steps, threads)},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="orbx", choices=["orbx", "reference"])
    ap.add_argument("--streams", type=int, default=256, help="independent stereo streams per GPU per step")
    ap.add_argument("--cpu-frames", type=int, default=150, help="stereo frames of the single-thread CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import orbx
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the orbx path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    S = args.streams
    B = 2 * S
    ctx = orbx.Context(local)
    ex = orbx.ORBextractor(ctx, NFEAT, 1.2, NLEVELS, 20, 7, max_w=W, max_h=H, max_batch=B)
    imgs = make_streams(S, seed0=100 + 1000 * rank)
    cap = ex.cap
    stream = torch.cuda.ExternalStream(ex.stream, device=local)

    # ---------------- resident arm: inputs already in HBM ----------------
    host = torch.from_numpy(np.stack(imgs)).pin_memory()
    d_img = host.cuda(non_blocking=False)
    d_kps = torch.empty((B, cap, 6), dtype=torch.float32, device="cuda")
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(B, dtype=torch.int32, device="cuda")
    d_mono = torch.zeros(B, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    torch.cuda.synchronize()

    def step_device():
        ex.extract_batch_device(d_img.data_ptr(), B, W, H, W, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                                d_n.data_ptr(), d_mono.data_ptr())

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    ex.set_profiling(True)
    launches0 = ctx.launches
    stage_sum = np.zeros(len(ex.STAGES))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                      # L2 flush between timed iterations (not timed)
            ev[k][0].record(stream)
            step_device()
            ev[k][1].record(stream)
        ev[k][1].synchronize()
        ms, ln = ex.stage_ms()
        stage_sum += ms
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    ex.set_profiling(False)
    kp_mean = float(d_n.float().mean().item())

    # ---------------- e2e arm: host buffers through the C ABI ----------------
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        ex.extract_batch(imgs)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = ex.extract_batch(imgs)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = B * W * H
    d2h = int(sum(len(r[1]) for r in res) * (24 + 32) + 2 * 4 * B)

    # ---------------- reduce over ranks (max time) ----------------
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    frames = S * world
    value = frames * args.steps / (dev_ms / 1e3)
    e2e_value = frames * e2e_steps / e2e_s

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (live CUDA-event stage times) ----------------
    stage_ms = stage_sum / args.steps
    alg = algorithmic_bytes(int(round(kp_mean)))
    dom = int(np.argmax(stage_ms))
    dom_name = ex.STAGES[dom]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom_bytes = alg[dom_name] * B
    n_launch = (NLEVELS - 1) if dom_name == "pyramid" else 1
    achieved = dom_bytes / n_launch / (stage_ms[dom] / n_launch * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom_name)
        except Exception:
            traffic = None
    roof = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic,
            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s",
            "stage_ms": {n: float(v) for n, v in zip(ex.STAGES, stage_ms)},
            "extractor_total": {"achieved": alg["total"] * B / (stage_ms.sum() * 1e-3) / 1e9,
                                "frac": alg["total"] * B / (stage_ms.sum() * 1e-3) / 1e9 / peak,
                                "bytes_per_image": alg["total"]}}

    cpu = None
    if not args.no_cpu and world >= 1:
        fps, dt = cpu_oracle_frames_per_s(args.cpu_frames, 1)
        cpu = {"value": fps, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d stereo frames (2 extractions each) of the same workload, %.1f s, 1 thread of %d cores"
                         % (args.cpu_frames, dt, os.cpu_count() or 0)}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": {"workload": "EuRoC MH01-shaped stereo 752x480, 1000 feat, 8 levels, extractor (L+R) only "
                                  "[matcher/pose-opt stages join as they land]",
                      "streams_per_gpu": S, "images_per_step_per_gpu": B, "parallelism": "replicas x%d" % world,
                      "l2": "256 MiB flush between timed steps + working set > L2",
                      "mean_keypoints_per_image": kp_mean},
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "steps": e2e_steps},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
