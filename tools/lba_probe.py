#!/usr/bin/env python3
"""One LocalBundleAdjustment call on the BASELINE config-4 shape (20 KF / ~3000 MP / ~16 k edges), for ncu."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "awesome-orb-slam3-3dvisioncraft-version_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import orbx  # noqa: E402
import scenarios as sc  # noqa: E402

ctx = orbx.Context(0)
cam = orbx.make_camera()
opt = orbx.Optimizer(ctx)
l = sc.lba_scenario(0)
a5 = (l["kf_T"], l["kf_fixed"], l["mp_xyz"], l["e_kf"], l["e_mp"], l["e_obs"], l["e_inv_sigma2"], cam)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    t0 = time.perf_counter()
    r = opt.LocalBundleAdjustment(*a5)
    print("LBA %.3f ms, iters %s" % ((time.perf_counter() - t0) * 1e3, r[3] if len(r) > 3 else ""), flush=True)
