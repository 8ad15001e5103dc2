#!/usr/bin/env python3
"""Single-call latency of every C-ABI entry point (host buffers in, host buffers out: what Tracking.cc /
LocalMapping.cc would observe through the shim) next to the CPU oracle on the same host, one thread.
Covers BASELINE.json configs 1-4 at the call level.  Writes profiles/<tag>_latency.md.

    python tools/latency_table.py r01          (on the GPU box)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "awesome-orb-slam3-3dvisioncraft-version_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import oracle  # noqa: E402
import orbx  # noqa: E402
import scenarios as sc  # noqa: E402
from orbx import synth  # noqa: E402


def med_ms(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(t)), float(np.percentile(t, 95))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    ctx = orbx.Context(0)
    cam = orbx.make_camera()
    rows = []

    def add(name, ref, gpu_fn, cpu_fn, reps=30, cpu_reps=10):
        g, g95 = med_ms(gpu_fn, reps)
        c, c95 = med_ms(cpu_fn, cpu_reps, warm=1)
        rows.append((name, ref, g, g95, c, c / g))
        print("%-58s gpu %8.3f ms  cpu %9.3f ms  x%.1f" % (name, g, c, c / g), flush=True)

    # --- extractor ---
    img = synth.scene_image(0)
    ex = orbx.ORBextractor(ctx)
    orc = oracle.Extractor()
    add("ORBextractor::operator() 752x480, 1000 feat", "src/ORBextractor.cc:1074", lambda: ex(img), lambda: orc(img))
    ex5, orc5 = orbx.ORBextractor(ctx, nfeatures=5000), oracle.Extractor(5000)
    add("ORBextractor::operator() 752x480, 5000 feat (mono init)", "src/Tracking.cc:233", lambda: ex5(img, (0, 1000)),
        lambda: orc5(img, (0, 1000)))
    big = synth.scene_image(14, 1920, 1080)
    exb, orcb = orbx.ORBextractor(ctx, nfeatures=2000, max_w=1920, max_h=1080), oracle.Extractor(2000)
    add("ORBextractor::operator() 1920x1080, 2000 feat", "BASELINE config 4", lambda: exb(big), lambda: orcb(big), 20, 5)
    # --- stereo + matchers on a real stereo frame ---
    L, R = synth.stereo_pair(3)
    exL, exR = orbx.ORBextractor(ctx), orbx.ORBextractor(ctx)
    _, kL, dL = exL(L)
    _, kR, dR = exR(R)
    oL, oR = oracle.Extractor(), oracle.Extractor()
    oL(L), oR(R)
    pyrL, pyrR = [oL.pyramid_level(l) for l in range(8)], [oR.pyramid_level(l) for l in range(8)]
    bf, b = sc.BF, sc.BF / sc.FX
    add("Frame::ComputeStereoMatches (1000 x 1000 kp)", "src/Frame.cc:955",
        lambda: orbx.stereo_match(ctx, exL, 0, exR, 0, kL, dL, kR, dR, bf, b),
        lambda: oracle.stereo_match(pyrL, pyrR, kL, dL, kR, dR, oL.scale, oL.inv_scale, bf, b))
    ur, dp = oracle.stereo_match(pyrL, pyrR, kL, dL, kR, dR, oL.scale, oL.inv_scale, bf, b)
    F = orbx.Frame(kL, dL, ur)
    s = sc.sbp_map_scenario(1, kL, dL, ur)
    a = (F, s["kp_blocked"], s["projX"], s["projY"], s["projXR"], s["level"], s["viewCos"], s["mpDesc"], s["flags"], 3.0)
    m = orbx.ORBmatcher(ctx, 0.8)
    add("ORBmatcher::SearchByProjection(F, 1500 MapPoints, th=3)", "src/ORBmatcher.cc:59",
        lambda: m.SearchByProjectionMap(*a, s["scaleFactors"]),
        lambda: oracle.search_by_projection_map(*a, 0.8, s["scaleFactors"]))
    f = sc.sbp_frame_scenario(2, kL, dL, ur, dp)
    a2 = (F, f["cur_blocked"], cam, f["Tcw_cur"], f["Tcw_last"], f["flags"], f["xw"], f["octave"], f["angle"], f["mpDesc"], 7.0, False)
    m9 = orbx.ORBmatcher(ctx, 0.9, True)
    add("ORBmatcher::SearchByProjection(Cur, Last, th=7)", "src/ORBmatcher.cc:2244",
        lambda: m9.SearchByProjectionFrame(*a2, f["scaleFactors"]),
        lambda: oracle.search_by_projection_frame(*a2, True, f["scaleFactors"]))
    # the same two searches on a device-resident Frame (orbx_frame_upload once per Frame: no per-call frame upload, no grid build)
    RF = orbx.ResidentFrame(ctx, F)
    add("  ... SearchByProjection(F, MapPoints) on a resident Frame", "orbx_frame_upload", lambda: m.SearchByProjectionMap(*a, s["scaleFactors"]),
        lambda: oracle.search_by_projection_map(*a, 0.8, s["scaleFactors"]))
    add("  ... SearchByProjection(Cur, Last) on a resident Frame", "orbx_frame_upload", lambda: m9.SearchByProjectionFrame(*a2, f["scaleFactors"]),
        lambda: oracle.search_by_projection_frame(*a2, True, f["scaleFactors"]))
    RF.release()
    # one stereo frame through the whole per-frame chain in ONE call (S = 1 tracker, 8(d) map uploaded with the call)
    from replay_reference import track_frame_map
    ex1 = orbx.ORBextractor(ctx, max_batch=2)
    trk1 = orbx.Tracker(ctx, ex1, 1, cam)
    Tt1 = np.eye(4, dtype=np.float32)[None]
    Tp1 = Tt1.copy()
    Tp1[0, :3, 3] = (0.01, -0.01, 0.005)
    mp1 = sc.track_map_scenario(5, kL, dL, ur, dp, Tt1[0])
    host1 = sc.stack_track_maps([mp1])
    oex = (oracle.Extractor(), oracle.Extractor())

    def chain_gpu():
        trk1.upload_map(host1)
        return trk1.step([L, R], Tt1, Tp1)
    add("whole per-frame chain of ONE stereo frame (extract L+R .. PoseOptimization #2), 1500-point map",
        "src/Tracking.cc:1793-2480", chain_gpu, lambda: track_frame_map(oracle, cam, L, R, mp1, Tp1[0], extractors=oex), 30, 5)
    # the same with the caller's frame buffers and map arrays in page-locked memory (what a capture pipeline that feeds a
    # GPU keeps them in): the images are DMA'd in place instead of being staged, the ten map arrays are true async copies
    pin = orbx.host_array((2,) + L.shape, np.uint8)
    pin[0], pin[1] = L, R
    host1p = {}
    for k, v in host1.items():
        host1p[k] = orbx.host_array(v.shape, v.dtype)
        host1p[k][...] = v

    def chain_gpu_pinned():
        trk1.upload_map(host1p)
        return trk1.step([pin[0], pin[1]], Tt1, Tp1)
    assert np.array_equal(chain_gpu()[1], chain_gpu_pinned()[1])
    add("  ... the same, frame buffers and map arrays page-locked", "orbx_host_alloc", chain_gpu_pinned,
        lambda: track_frame_map(oracle, cam, L, R, mp1, Tp1[0], extractors=oex), 30, 5)
    trk1.set_graph(True)
    assert np.array_equal(chain_gpu_pinned()[1], chain_gpu()[1])
    add("  ... the same, replayed as ONE CUDA graph (orbx_tracker_set_graph)", "orbx_tracker_set_graph", chain_gpu_pinned,
        lambda: track_frame_map(oracle, cam, L, R, mp1, Tp1[0], extractors=oex), 30, 5)
    assert trk1.graph_launches > 20
    trk1.set_graph(False)
    trk1.close()
    ex1.close()
    q = sc.tri_scenario(3, kL, dL, ur)
    K1, K2 = orbx.Frame(q["k1"], q["d1"], q["ur1"]), orbx.Frame(q["k2"], q["d2"], q["ur2"])
    a3 = (K1, K2, q["has1"], q["has2"], q["fv1"], q["fv2"], cam, cam, q["R1w"], q["t1w"], q["R2w"], q["t2w"], q["sigma2"], q["scaleFactors"])
    m6 = orbx.ORBmatcher(ctx, 0.6, True)
    add("ORBmatcher::SearchForTriangulation (2 x ~1000 kp)", "src/ORBmatcher.cc:1138",
        lambda: m6.SearchForTriangulation(*a3), lambda: oracle.search_for_triangulation(*a3))
    # --- the "next" rows ---
    import voc_util as vu
    vb = vu.make_vocabulary(3, 10, 5, ragged=False, p_early_leaf=0.0)       # 111 110 nodes, the reference's k with one level less
    V = vu.parse(vb)
    qd = vu.query_descriptors(V, 1, 1000, 30)
    dv, ov = orbx.ORBVocabulary(ctx, vb), oracle.Vocabulary(vb)
    add("ORBVocabulary::transform, 1000 descriptors (k=10, L=5 synthetic tree)", "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1140",
        lambda: dv.transform(qd, 3), lambda: ov.transform(qd, 3))
    bs = sc.bow_scenario(1)
    KFb, Fb = orbx.Frame(bs["kK"], bs["dK"]), orbx.Frame(bs["kF"], bs["dF"])
    add("ORBmatcher::SearchByBoW(KF, F) (900 x 950 kp)", "src/ORBmatcher.cc:323",
        lambda: orbx.search_by_bow(ctx, KFb, Fb, bs["has"], bs["fvK"], bs["fvF"], 0.7, True),
        lambda: oracle.search_by_bow(KFb, Fb, bs["has"], bs["fvK"], bs["fvF"], 0.7, True))
    fs = sc.fuse_scenario(1, 1000, 1500)
    KFf = orbx.Frame(fs["kK"], fs["dK"], fs["ur"])
    fa = (KFf, cam, fs["R"], fs["t"], fs["Ow"], fs["flags"], fs["xw"], fs["maxd"], fs["mind"], fs["normal"], fs["desc"], 3.0, fs["scale"],
          fs["inv_sigma2"], fs["log_sf"])
    add("ORBmatcher::Fuse(KF, 1500 MapPoints) search", "src/ORBmatcher.cc:1630", lambda: orbx.fuse(ctx, *fa), lambda: oracle.fuse(*fa))
    fr = (cam, fs["R"], fs["t"], fs["Ow"], (0.0, 752.0, 0.0, 480.0), 0.5, 8, fs["log_sf"], fs["xw"], fs["maxd"], fs["mind"], fs["normal"])
    add("Frame::isInFrustum x 1500 MapPoints", "src/Frame.cc:571", lambda: orbx.is_in_frustum(ctx, *fr), lambda: oracle.is_in_frustum(*fr))
    xy = np.stack([kL["x"], kL["y"]], 1)
    dist = [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]
    add("Frame::UndistortKeyPoints (1000 kp)", "src/Frame.cc:874", lambda: orbx.undistort_keypoints(ctx, xy, cam, dist),
        lambda: oracle.undistort_points(xy, cam, dist))
    # --- optimisers ---
    opt = orbx.Optimizer(ctx)
    for E in (150, 300, 500):
        p = sc.pose_opt_scenario(E, E=E)
        a4 = (p["xw"], p["obs"], p["inv_sigma2"], cam, p["Tcw"])
        add("Optimizer::PoseOptimization, %d edges" % E, "src/Optimizer.cc:907", lambda: opt.PoseOptimization(*a4),
            lambda: oracle.pose_optimization(*a4))
    l = sc.lba_scenario(0)
    a5 = (l["kf_T"], l["kf_fixed"], l["mp_xyz"], l["e_kf"], l["e_mp"], l["e_obs"], l["e_inv_sigma2"], cam)
    add("Optimizer::LocalBundleAdjustment 20 KF / %d MP / %d edges" % (len(l["mp_xyz"]), len(l["e_kf"])),
        "src/Optimizer.cc:1811", lambda: opt.LocalBundleAdjustment(*a5), lambda: oracle.local_ba(*a5), 10, 5)
    for E in (300, 1000):
        i = sc.inertial_scenario(E, E, 0.6)
        a6 = (i["xw"], i["obs"], i["isg"], i["close"], cam, i["Tcw"], i["Tcb"], i["Tbc"], i["state"], i["kf"], i["preint"], i["infoI"],
              i["infoG"], i["infoA"])
        add("Optimizer::PoseInertialOptimizationLastKeyFrame, %d edges" % E, "src/Optimizer.cc:7665",
            lambda: opt.PoseInertialOptimizationLastKeyFrame(*a6), lambda: oracle.pose_inertial_optimization_last_keyframe(i, cam))
    for E in (300, 1000):
        j = sc.inertial_lf_scenario(E, E, 0.6)
        a7 = (j["xw"], j["obs"], j["isg"], j["close"], cam, j["Tcw"], j["Tcb"], j["Tbc"], j["state"], j["prev"], j["preint"],
              j["preint_jac"], j["preint_bias"], j["infoI"], j["infoG"], j["infoA"], j["prior_state"], j["prior_H"])
        add("Optimizer::PoseInertialOptimizationLastFrame, %d edges" % E, "src/Optimizer.cc:8068",
            lambda: opt.PoseInertialOptimizationLastFrame(*a7), lambda: oracle.pose_inertial_optimization_last_frame(j, cam))
    out = ["# Single-call latency through the C ABI vs the CPU oracle (%s)" % tag, "",
           "Host buffers in, host buffers out, median of repeated synchronous calls (p95 in brackets); CPU = the oracle on "
           "one thread of the same host (%d cores).  This is the call-level view of BASELINE.json configs 1-4; the "
           "many-stream throughput is bench.py's number." % (os.cpu_count() or 0), "",
           "| call | replaces | GPU ms (p95) | CPU oracle ms | CPU/GPU |", "|---|---|---|---|---|"]
    for name, ref, g, g95, c, r in rows:
        out.append("| %s | %s | %.3f (%.3f) | %.3f | %.1fx |" % (name, ref, g, g95, c, r))
    open(os.path.join(ROOT, "profiles", "%s_latency.md" % tag), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
