"""Throughput of the many-stream PoseInertialOptimizationLastFrame launch vs the CPU oracle (tools, not a test).
usage: python tools/f3_probe.py [E]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "awesome-orb-slam3-3dvisioncraft-version_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import orbx  # noqa: E402
import oracle  # noqa: E402
import scenarios as sc  # noqa: E402


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    ctx = orbx.Context(0)
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    base = [sc.inertial_lf_scenario(100 + i, E, 0.6) for i in range(37)]
    t0 = time.perf_counter()
    for s in base[:8]:
        oracle.pose_inertial_optimization_last_frame(s, cam)
    cpu = (time.perf_counter() - t0) / 8
    print("E=%d  CPU oracle %.3f ms/problem (%.0f problems/s/thread)" % (E, cpu * 1e3, 1 / cpu))
    for P in (1, 37, 148, 296, 592, 1184):
        probs = opt.pack_inertial_lf([base[i % len(base)] for i in range(P)])
        opt.PoseInertialOptimizationLastFrameBatch(probs, cam)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            opt.PoseInertialOptimizationLastFrameBatch(probs, cam)
            ts.append(time.perf_counter() - t0)
        t = float(np.median(ts))
        print("P=%5d  %.3f ms/launch incl. H2D/D2H of every argument and result  ->  %.0f problems/s  (x%.1f one CPU thread)" % (P, t * 1e3, P / t, P / t * cpu))


if __name__ == "__main__":
    main()
