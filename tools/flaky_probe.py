#!/usr/bin/env python3
"""Repeat the inertial tracker parity check on freshly created trackers (80 trials) and print every mismatch.  Written while
chasing a rare wrong result on a cold box (a lazily allocated staging buffer zeroed by a late legacy-stream cudaMemset,
DESIGN.md section 7); kept as a regression probe.  GPU box: python tools/flaky_probe.py"""
import sys, os
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awesome-orb-slam3-3dvisioncraft-version_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, oracle as ork, scenarios as sc, orbx
import importlib.util
spec = importlib.util.spec_from_file_location('tg', os.path.join(ROOT, 'tests/test_tracker_gpu.py')); tg = importlib.util.module_from_spec(spec); spec.loader.exec_module(tg)
from replay_reference import track_frame_map
ctx = orbx.Context(0)
cam = orbx.make_camera()
S = 2
imgs, Tt, Tp, maps = tg._map_setup(ork, S, 180)
host = sc.stack_track_maps(maps)
bad = 0
for mode in (1, 2):
    imus = [sc.track_imu_scenario(900 + s, Tt[s], mode) for s in range(S)]
    himu = sc.stack_track_imu(imus)
    want = [track_frame_map(ork, cam, imgs[2 * s], imgs[2 * s + 1], maps[s], Tp[s], imu=imus[s], imu_mode=mode, want_inertial=True) for s in range(S)]
    for trial in range(40):
        ex = orbx.ORBextractor(ctx, max_batch=2 * S)
        trk = orbx.Tracker(ctx, ex, S, cam)
        for rep in range(2):
            trk.upload_map(host)
            trk.upload_inertial(mode, himu)
            Tout, stats = trk.step(imgs, Tt, Tp)
            state, H = trk.inertial_result()
            for s in range(S):
                T2, st, res = want[s]
                if not np.array_equal(stats[s], st) or np.abs(state[s] - res["state"]).max() > 1e-9:
                    bad += 1
                    print("MISMATCH mode", mode, "trial", trial, "rep", rep, "stream", s, stats[s], st, np.abs(state[s] - res["state"]).max(), flush=True)
        trk.close(); ex.close()
print("done, mismatches:", bad)
