"""GPU parity: liborbx matchers (C ABI) vs the CPU oracle on identical flat buffers — bit-exact index
outputs."""
import numpy as np
import pytest

import scenarios as sc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stereo_frames(ork):
    from orbx import synth
    out = []
    for seed in (3, 8):
        L, R = synth.stereo_pair(seed)
        exL, kL, dL = sc.extract_frame(ork, L)
        exR, kR, dR = sc.extract_frame(ork, R)
        pyrL = [exL.pyramid_level(l) for l in range(8)]
        pyrR = [exR.pyramid_level(l) for l in range(8)]
        ur, dp = ork.stereo_match(pyrL, pyrR, kL, dL, kR, dR, exL.scale, exL.inv_scale, sc.BF, sc.BF / sc.FX)
        out.append(dict(L=L, R=R, kL=kL, dL=dL, kR=kR, dR=dR, ur=ur, dp=dp, scale=exL.scale))
    return out


def test_descriptor_distance(ctx, ork):
    import orbx
    voc = sc.orbvoc()
    rng = np.random.default_rng(0)
    for _ in range(200):
        a, b = voc[rng.integers(len(voc))], voc[rng.integers(len(voc))]
        ref = int(np.unpackbits(a ^ b).sum())
        assert orbx.ORBmatcher.DescriptorDistance(a, b) == ref == ork.descriptor_distance(a, b)


def test_features_in_area(ctx, ork, stereo_frames):
    import orbx
    f = stereo_frames[0]
    F = orbx.Frame(f["kL"], f["dL"], f["ur"])
    rng = np.random.default_rng(1)
    nq = 600
    x = rng.uniform(-40, 800, nq).astype(np.float32)
    y = rng.uniform(-40, 520, nq).astype(np.float32)
    r = rng.uniform(1, 80, nq).astype(np.float32)
    mn = rng.integers(-1, 6, nq).astype(np.int32)
    mx = np.where(rng.random(nq) < 0.2, -1, mn + rng.integers(0, 3, nq)).astype(np.int32)
    go, gn = orbx.features_in_area(ctx, F, x, y, r, mn, mx)
    oo, on = ork.features_in_area(F, x, y, r, mn, mx)
    assert np.array_equal(gn, on)
    for q in range(nq):
        assert np.array_equal(go[q, :min(gn[q], 256)], oo[q, :min(on[q], 256)]), q


def test_stereo_match(ctx, ork):
    import orbx
    from orbx import synth
    for seed in (3, 8, 21):
        L, R = synth.stereo_pair(seed)
        exL, exR = orbx.ORBextractor(ctx), orbx.ORBextractor(ctx)
        _, kL, dL = exL(L)
        _, kR, dR = exR(R)
        oL, oR = ork.Extractor(), ork.Extractor()
        oL(L), oR(R)
        pyrL = [oL.pyramid_level(l) for l in range(8)]
        pyrR = [oR.pyramid_level(l) for l in range(8)]
        rur, rdp = ork.stereo_match(pyrL, pyrR, kL, dL, kR, dR, oL.scale, oL.inv_scale, sc.BF, sc.BF / sc.FX)
        gur, gdp = orbx.stereo_match(ctx, exL, 0, exR, 0, kL, dL, kR, dR, sc.BF, sc.BF / sc.FX)
        assert (rur >= 0).sum() > 200
        assert np.array_equal(rur, gur) and np.array_equal(rdp, gdp), seed
        exL.close(), exR.close()


def test_stereo_match_batched_images_in_one_extractor(ctx, ork):
    """Left and right images as images 0/1 of one batched extractor call (the many-stream layout)."""
    import orbx
    from orbx import synth
    L, R = synth.stereo_pair(5)
    ex = orbx.ORBextractor(ctx, max_batch=2)
    (_, kL, dL), (_, kR, dR) = ex.extract_batch([L, R])
    gur, gdp = orbx.stereo_match(ctx, ex, 0, ex, 1, kL, dL, kR, dR, sc.BF, sc.BF / sc.FX)
    oL, oR = ork.Extractor(), ork.Extractor()
    oL(L), oR(R)
    rur, rdp = ork.stereo_match([oL.pyramid_level(l) for l in range(8)], [oR.pyramid_level(l) for l in range(8)],
                                kL, dL, kR, dR, oL.scale, oL.inv_scale, sc.BF, sc.BF / sc.FX)
    assert np.array_equal(rur, gur) and np.array_equal(rdp, gdp)


@pytest.mark.parametrize("th,nnratio,mono", [(1.0, 0.8, False), (3.0, 0.8, False), (5.0, 0.6, True), (15.0, 0.9, False)])
def test_search_by_projection_map(ctx, ork, stereo_frames, th, nnratio, mono):
    import orbx
    for fi, f in enumerate(stereo_frames):
        ur = None if mono else f["ur"]
        F = orbx.Frame(f["kL"], f["dL"], ur)
        s = sc.sbp_map_scenario(10 + fi, f["kL"], f["dL"], ur)
        args = (F, s["kp_blocked"], s["projX"], s["projY"], None if mono else s["projXR"], s["level"], s["viewCos"],
                s["mpDesc"], s["flags"], th)
        rn, rbest = ork.search_by_projection_map(*args, nnratio, s["scaleFactors"])
        gn, gbest = orbx.ORBmatcher(ctx, nnratio).SearchByProjectionMap(*args, s["scaleFactors"])
        assert rn > 100
        assert gn == rn and np.array_equal(gbest, rbest), (th, fi, int((gbest != rbest).sum()))


@pytest.mark.parametrize("th,bMono,forward", [(7.0, False, 0.0), (15.0, True, 0.0), (7.0, False, 0.4),
                                              (7.0, False, -0.4), (30.0, True, 0.0)])
def test_search_by_projection_frame(ctx, ork, stereo_frames, th, bMono, forward):
    import orbx
    cam = orbx.make_camera()
    for fi, f in enumerate(stereo_frames):
        ur = None if bMono else f["ur"]
        F = orbx.Frame(f["kL"], f["dL"], ur)
        s = sc.sbp_frame_scenario(20 + fi, f["kL"], f["dL"], f["ur"], f["dp"], forward=forward)
        a = (F, s["cur_blocked"], cam, s["Tcw_cur"], s["Tcw_last"], s["flags"], s["xw"], s["octave"], s["angle"],
             s["mpDesc"], th, bMono)
        for ori in (True, False):
            rn, rm, rk, rc = ork.search_by_projection_frame(*a, ori, s["scaleFactors"])
            gn, gm, gk, gc = orbx.ORBmatcher(ctx, 0.9, ori).SearchByProjectionFrame(*a, s["scaleFactors"])
            assert rn > 50
            assert gn == rn and np.array_equal(gm, rm) and np.array_equal(gk, rk) and np.array_equal(gc, rc), \
                (th, bMono, forward, ori)


@pytest.mark.parametrize("only_stereo,coarse,wrong_pose", [(False, False, False), (False, True, False),
                                                          (True, False, False), (False, False, True)])
def test_search_for_triangulation(ctx, ork, stereo_frames, only_stereo, coarse, wrong_pose):
    import orbx
    cam = orbx.make_camera()
    for fi, f in enumerate(stereo_frames):
        q = sc.tri_scenario(30 + fi, f["kL"], f["dL"], f["ur"])
        if wrong_pose:   # epipolar gate must now reject many descriptor-consistent pairs
            q["t2w"] = (q["t2w"] + np.array([0.0, 0.25, 0.1], np.float32)).astype(np.float32)
        K1, K2 = orbx.Frame(q["k1"], q["d1"], q["ur1"]), orbx.Frame(q["k2"], q["d2"], q["ur2"])
        a = (K1, K2, q["has1"], q["has2"], q["fv1"], q["fv2"], cam, cam, q["R1w"], q["t1w"], q["R2w"], q["t2w"],
             q["sigma2"], q["scaleFactors"], only_stereo, coarse)
        rn, rm = ork.search_for_triangulation(*a, True)
        gn, gm = orbx.ORBmatcher(ctx, 0.6, True).SearchForTriangulation(*a)
        assert gn == rn and np.array_equal(gm, rm), (only_stereo, coarse, wrong_pose, int((gm != rm).sum()))
        if not only_stereo and not wrong_pose:
            assert rn > 100


def test_matchers_handle_empty_inputs(ctx, ork):
    import orbx
    cam = orbx.make_camera()
    kp0 = np.zeros(0, orbx.KP_DTYPE)
    F0 = orbx.Frame(kp0, np.zeros((0, 32), np.uint8), None)
    m = orbx.ORBmatcher(ctx, 0.8)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    z = np.zeros(0, np.float32)
    n, best = m.SearchByProjectionMap(F0, None, z, z, None, np.zeros(0, np.int32), z, np.zeros((0, 32), np.uint8),
                                      np.zeros(0, np.uint8), 1.0, sf)
    assert n == 0 and len(best) == 0
    n, mt, kp, cm = m.SearchByProjectionFrame(F0, None, cam, np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32),
                                              np.zeros(0, np.uint8), np.zeros((0, 3), np.float32),
                                              np.zeros(0, np.int32), z, np.zeros((0, 32), np.uint8), 7.0, False, sf)
    assert n == 0


def test_projection_searches_on_a_frame_above_48k_keypoints(ctx, ork):
    """The ordered-replay kernels keep one byte per keypoint in dynamic shared memory: above 49 136 keypoints that needs the
    opt-in (and every launch is checked) — the result must still be the oracle's, not untouched output buffers."""
    import orbx
    n = 52000
    kps, desc = sc.synthetic_keypoints(5, n)
    F = orbx.Frame(kps, desc, None)
    s = sc.sbp_map_scenario(77, kps[:1000], desc[:1000], None, nq=300)
    args = (F, None, s["projX"], s["projY"], None, s["level"], s["viewCos"], s["mpDesc"], s["flags"], 3.0)
    rn, rbest = ork.search_by_projection_map(*args, 0.8, s["scaleFactors"])
    gn, gbest = orbx.ORBmatcher(ctx, 0.8).SearchByProjectionMap(*args, s["scaleFactors"])
    assert rn > 20 and gn == rn and np.array_equal(gbest, rbest)


def test_resident_frame_gives_identical_results_without_reuploading(ctx, ork, stereo_frames):
    """orbx_frame_upload: a Frame made device-resident once is found by every later matcher call over the same host
    arrays (no signature change); results are identical to the per-call upload path, before and after release."""
    import orbx
    cam = orbx.make_camera()
    f = stereo_frames[0]
    F = orbx.Frame(f["kL"], f["dL"], f["ur"])
    s1 = sc.sbp_map_scenario(10, f["kL"], f["dL"], f["ur"])
    s2 = sc.sbp_frame_scenario(20, f["kL"], f["dL"], f["ur"], f["dp"])
    m = orbx.ORBmatcher(ctx, 0.8)

    def run():
        a = m.SearchByProjectionMap(F, s1["kp_blocked"], s1["projX"], s1["projY"], s1["projXR"], s1["level"], s1["viewCos"], s1["mpDesc"],
                                    s1["flags"], 3.0, s1["scaleFactors"])
        b = orbx.ORBmatcher(ctx, 0.9, True).SearchByProjectionFrame(F, s2["cur_blocked"], cam, s2["Tcw_cur"], s2["Tcw_last"], s2["flags"],
                                                                     s2["xw"], s2["octave"], s2["angle"], s2["mpDesc"], 7.0, False,
                                                                     s2["scaleFactors"])
        x = np.array([100.0, 400.0, 700.0], np.float32)
        c = orbx.features_in_area(ctx, F, x, x * 0.5, np.full(3, 30.0, np.float32), np.zeros(3, np.int32), np.full(3, 7, np.int32))
        return a, b, c

    before = run()
    n0 = orbx.load_library().orbx_frame_count(ctx.h)
    R = orbx.ResidentFrame(ctx, F)
    assert orbx.load_library().orbx_frame_count(ctx.h) == n0 + 1
    launches0 = ctx.launches
    during = run()
    used = ctx.launches - launches0
    R.release()
    assert orbx.load_library().orbx_frame_count(ctx.h) == n0
    launches0 = ctx.launches
    after = run()
    assert used == (ctx.launches - launches0) - 3          # three grid builds saved
    for x, y in ((before, during), (before, after)):
        assert x[0][0] == y[0][0] and np.array_equal(x[0][1], y[0][1])
        assert x[1][0] == y[1][0] and all(np.array_equal(p, q) for p, q in zip(x[1][1:], y[1][1:]))
        assert np.array_equal(x[2][0], y[2][0]) and np.array_equal(x[2][1], y[2][1])
