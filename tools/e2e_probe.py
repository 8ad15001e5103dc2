#!/usr/bin/env python3
"""Where does the end-to-end arm lose time against the resident arm?  Times, for NP pipelines of S/NP streams each:
  A  host threads calling step_device + synchronize (inputs resident: no H2D)
  B  host threads calling the host-buffer step (pinned H2D inside)
  C  one host thread issuing step_device on all pipelines, one synchronize per round."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "awesome-orb-slam3-3dvisioncraft-version_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
import orbx  # noqa: E402

S, W, H = 256, bench.W, bench.H
ctx = orbx.Context(0)
cam = orbx.make_camera()
imgs = bench.make_streams(S)
Tt, Tp = bench.make_poses(S)
for NP in (1, 2, 4, 8):
    bounds = [S * k // NP for k in range(NP + 1)]
    pipes = []
    for k in range(NP):
        s0, s1 = bounds[k], bounds[k + 1]
        n = s1 - s0
        ex = orbx.ORBextractor(ctx, bench.NFEAT, 1.2, bench.NLEVELS, 20, 7, max_w=W, max_h=H, max_batch=2 * n)
        trk = orbx.Tracker(ctx, ex, n, cam, th_frame=7.0, th_map=1.0, nnratio_map=0.8)
        pin = orbx.host_array((2 * n, H, W), np.uint8)
        pin[:] = np.stack(imgs[2 * s0:2 * s1])
        pipes.append(dict(n=n, ex=ex, trk=trk, imgs=[pin[i] for i in range(2 * n)], Tt=Tt[s0:s1], Tp=Tp[s0:s1],
                          d_img=torch.from_numpy(np.stack(imgs[2 * s0:2 * s1])).cuda(),
                          d_true=torch.from_numpy(Tt[s0:s1].reshape(n, 16)).cuda(), d_prior=torch.from_numpy(Tp[s0:s1].reshape(n, 16)).cuda(),
                          d_out=torch.zeros((n, 16), dtype=torch.float32, device="cuda"),
                          d_stats=torch.zeros((n, 8), dtype=torch.int32, device="cuda")))
    torch.cuda.synchronize()

    def dev(P):
        P["trk"].step_device(P["d_img"].data_ptr(), W, H, W, P["d_true"].data_ptr(), P["d_prior"].data_ptr(), P["d_out"].data_ptr(),
                             P["d_stats"].data_ptr())

    def A(P):
        dev(P)
        P["trk"].synchronize()

    def B(P):
        P["trk"].step(P["imgs"], P["Tt"], P["Tp"])

    steps = 10
    res = {}
    if os.environ.get("PROBE_OVERLAP", "0") == "1":
        for P in pipes:
            P["trk"].set_overlap(True)     # stage B (matching + pose) of every pipeline on its own high-priority stream
    with ThreadPoolExecutor(NP) as pool:
        for name, fn in (("A", A), ("B", B)):
            for _ in range(2):
                list(pool.map(fn, pipes))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                list(pool.map(fn, pipes))
            torch.cuda.synchronize()
            res[name] = S * steps / (time.perf_counter() - t0)
    for _ in range(2):
        for P in pipes:
            dev(P)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for P in pipes:
            dev(P)
        for P in pipes:
            P["trk"].synchronize()
    torch.cuda.synchronize()
    res["C"] = S * steps / (time.perf_counter() - t0)
    print("pipelines %d: A threads+resident %.0f  B threads+H2D %.0f  C one thread resident %.0f frames/s" % (NP, res["A"], res["B"], res["C"]), flush=True)
    del pipes
