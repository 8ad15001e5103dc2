#!/usr/bin/env python3
"""Per-phase cycle breakdown of pose_opt_kernel on the tracker's 8(d) workload (orbx_debug_pose_opt_profile).
python tools/pose_probe.py [S]   (GPU box)"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "awesome-orb-slam3-3dvisioncraft-version_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

NAMES = ["build pass", "reduction+unpack", "solve+exp (warp 0)", "trial residual pass", "trial reduction", "accept/reject logic",
         "chi2 classification", "whole kernel"]


def main():
    import orbx
    import oracle as ork
    import scenarios as sc
    from orbx import synth
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    U = 8
    rng = np.random.default_rng(5)
    cam = orbx.make_camera()
    imgs, maps, Tt, Tp = [], [], [], []
    for u in range(U):
        L, R = synth.stereo_pair(300 + u)
        exL, kL, dL = sc.extract_frame(ork, L)
        exR, kR, dR = sc.extract_frame(ork, R)
        ur, dp = ork.stereo_match([exL.pyramid_level(l) for l in range(8)], [exR.pyramid_level(l) for l in range(8)], kL, dL, kR, dR,
                                  exL.scale, exL.inv_scale, sc.BF, sc.BF / sc.FX)
        T = sc.se3_matrix(sc.rot_small(rng, 5.0), rng.uniform(-0.5, 0.5, 3)).astype(np.float32)
        P = (sc.se3_matrix(sc.rot_small(rng, 0.4), rng.normal(0, 0.01, 3)) @ T.astype(np.float64)).astype(np.float32)
        imgs.append((L, R)); Tt.append(T); Tp.append(P)
        maps.append(sc.track_map_scenario(400 + u, kL, dL, ur, dp, T, n_map=1500))
    images = []
    for s in range(S):
        images += list(imgs[s % U])
    TtS = np.stack([Tt[s % U] for s in range(S)])
    TpS = np.stack([Tp[s % U] for s in range(S)])
    host = sc.stack_track_maps([maps[s % U] for s in range(S)])
    ctx = orbx.Context(0)
    ex = orbx.ORBextractor(ctx, max_batch=2 * S)
    trk = orbx.Tracker(ctx, ex, S, cam)
    lib = orbx.load_library()
    lib.orbx_debug_pose_opt_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    for _ in range(2):
        trk.upload_map(host)
        out, st = trk.step(images, TtS, TpS)
    lib.orbx_debug_pose_opt_profile(ctx.h, 1, None)
    trk.set_profiling(True)
    n = 5
    ms = np.zeros(len(trk.STAGES))
    for _ in range(n):
        trk.upload_map(host)
        out, st = trk.step(images, TtS, TpS)
        ms += trk.stage_ms()
    prof = (C.c_ulonglong * 16)()
    lib.orbx_debug_pose_opt_profile(ctx.h, 0, prof)
    p = np.array(list(prof), np.float64)
    ctas = max(p[11], 1)
    print("S=%d  stage ms (profiling on):" % S, dict(zip(trk.STAGES, np.round(ms / n, 3))))
    print("mean stats:", dict(zip(trk.STATS, st.mean(0).round(1))))
    print("per CTA (one PoseOptimization call): %.0f cycles = %.1f us at 1.965 GHz; builds %.1f, solved trials %.1f, replayed %.1f"
          % (p[7] / ctas, p[7] / ctas / 1965, p[8] / ctas, p[9] / ctas, p[10] / ctas))
    for k in range(7):
        print("  %-24s %9.0f cycles  %5.1f %%" % (NAMES[k], p[k] / ctas, 100 * p[k] / max(p[7], 1)))
    print("  per build pass %.0f cycles, per reduction %.0f, per solve+exp %.0f, per trial pass %.0f, per trial reduction %.0f"
          % (p[0] / max(p[8], 1), p[1] / max(p[8], 1), p[2] / max(p[9], 1), p[3] / max(p[9], 1), p[4] / max(p[9], 1)))


if __name__ == "__main__":
    main()
