"""CPU: the oracle's matchers (oracle/ork_matcher.cpp) against the reference's OWN src/ORBmatcher.cc, compiled UNMODIFIED
into oracle/_ref/libref_matcher.so behind the stand-in Frame / KeyFrame / MapPoint of oracle/ref_stub/matcher_prelude.h.

Same flat inputs to both sides (the scenarios the -m gpu parity tests use), outputs compared index for index:
rows a10-a13 and f2 of SURVEY.md §8 are thereby pinned to reference source, not to a second restatement.
"""
import numpy as np
import pytest

import scenarios as sc

ref = pytest.importorskip("oracle.ref")
pytestmark = pytest.mark.skipif(not ref.available("libref_matcher.so"), reason="oracle/_ref/libref_matcher.so not built")


@pytest.fixture(scope="module")
def frames(ork):
    from orbx import synth
    out = []
    for seed in (3, 8):
        L, R = synth.stereo_pair(seed)
        exL, kL, dL = sc.extract_frame(ork, L)
        exR, kR, dR = sc.extract_frame(ork, R)
        ur, dp = ork.stereo_match([exL.pyramid_level(l) for l in range(8)], [exR.pyramid_level(l) for l in range(8)], kL, dL,
                                  kR, dR, exL.scale, exL.inv_scale, sc.BF, sc.BF / sc.FX)
        out.append(dict(kL=kL, dL=dL, ur=ur, dp=dp))
    return out


def test_descriptor_distance_and_three_maxima(ork):
    voc = sc.orbvoc()
    rng = np.random.default_rng(0)
    for _ in range(300):
        a, b = voc[rng.integers(len(voc))], voc[rng.integers(len(voc))]
        assert ref.descriptor_distance(a, b) == ork.descriptor_distance(a, b) == int(np.unpackbits(a ^ b).sum())

    def three(h):   # the oracle's statement of ORBmatcher::ComputeThreeMaxima, in numpy
        m1 = m2 = m3 = 0
        i1 = i2 = i3 = -1
        for i, s in enumerate(h):
            if s > m1:
                m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
            elif s > m2:
                m3, m2, i3, i2 = m2, s, i2, i
            elif s > m3:
                m3, i3 = s, i
        if m2 < np.float32(0.1) * np.float32(m1):
            i2 = i3 = -1
        elif m3 < np.float32(0.1) * np.float32(m1):
            i3 = -1
        return i1, i2, i3
    for t in range(500):
        h = rng.integers(0, 40, 30) if t % 3 else rng.integers(0, 4, 30)
        if t % 7 == 0:
            h[rng.integers(30)] = 400          # one dominant bin: the 10 % rule fires
        assert ref.three_maxima(h) == three(h.tolist())


def _replay(best, n, blocked):
    cur = np.full(n, -1, np.int32)
    for q, b in enumerate(best):            # "the caller replays F.mvpMapPoints[best_idx[q]] = pMP_q in increasing q"
        if b >= 0:
            cur[b] = q
    cur[(np.asarray(blocked) > 0) & (cur < 0)] = -2
    return cur


@pytest.mark.parametrize("th,nnratio,mono", [(1.0, 0.8, False), (3.0, 0.8, False), (5.0, 0.6, True), (15.0, 0.9, False)])
def test_search_by_projection_map(ork, frames, th, nnratio, mono):
    import orbx
    for fi, f in enumerate(frames):
        ur = None if mono else f["ur"]
        F = orbx.Frame(f["kL"], f["dL"], ur)
        s = sc.sbp_map_scenario(10 + fi, f["kL"], f["dL"], ur)
        args = (F, s["kp_blocked"], s["projX"], s["projY"], None if mono else s["projXR"], s["level"], s["viewCos"], s["mpDesc"],
                s["flags"], th)
        on, obest = ork.search_by_projection_map(*args, nnratio, s["scaleFactors"])
        rn, rcur = ref.search_by_projection_map(*args, nnratio, s["scaleFactors"])
        assert on == rn > 100
        assert np.array_equal(_replay(obest, F.n, s["kp_blocked"]), rcur)        # final F.mvpMapPoints, keypoint by keypoint


@pytest.mark.parametrize("th,bMono,forward", [(7.0, False, 0.0), (15.0, True, 0.0), (7.0, False, 0.4), (7.0, False, -0.4),
                                              (30.0, True, 0.0)])
def test_search_by_projection_frame(ork, frames, th, bMono, forward):
    import orbx
    cam = orbx.make_camera()
    for fi, f in enumerate(frames):
        ur = None if bMono else f["ur"]
        F = orbx.Frame(f["kL"], f["dL"], ur)
        s = sc.sbp_frame_scenario(20 + fi, f["kL"], f["dL"], f["ur"], f["dp"], forward=forward)
        a = (F, s["cur_blocked"], cam, s["Tcw_cur"], s["Tcw_last"], s["flags"], s["xw"], s["octave"], s["angle"], s["mpDesc"], th, bMono)
        for ori in (True, False):
            on, _, _, oc = ork.search_by_projection_frame(*a, ori, s["scaleFactors"])
            rn, rc = ref.search_by_projection_frame(*a, ori, s["scaleFactors"])
            oc = oc.copy()
            oc[(s["cur_blocked"] > 0) & (oc < 0)] = -2
            assert on == rn > 50
            assert np.array_equal(oc, rc), (th, bMono, forward, ori)              # final CurrentFrame.mvpMapPoints


@pytest.mark.parametrize("only_stereo,coarse,wrong_pose", [(False, False, False), (False, True, False), (True, False, False),
                                                          (False, False, True)])
def test_search_for_triangulation(ork, frames, only_stereo, coarse, wrong_pose):
    import orbx
    cam = orbx.make_camera()
    for fi, f in enumerate(frames):
        q = sc.tri_scenario(30 + fi, f["kL"], f["dL"], f["ur"])
        if wrong_pose:
            q["t2w"] = (q["t2w"] + np.array([0.0, 0.25, 0.1], np.float32)).astype(np.float32)
        K1, K2 = orbx.Frame(q["k1"], q["d1"], q["ur1"]), orbx.Frame(q["k2"], q["d2"], q["ur2"])
        a = (K1, K2, q["has1"], q["has2"], q["fv1"], q["fv2"], cam, cam, q["R1w"], q["t1w"], q["R2w"], q["t2w"], q["sigma2"],
             q["scaleFactors"], only_stereo, coarse)
        on, om = ork.search_for_triangulation(*a, True)
        rn, rm = ref.search_for_triangulation(*a, True)
        assert on == rn and np.array_equal(om, rm), (only_stereo, coarse, wrong_pose)
        if not only_stereo and not wrong_pose:
            assert rn > 100


@pytest.mark.parametrize("seed,nnratio,ori,nk,nf,nw", [(1, 0.7, True, 900, 950, 60), (2, 0.9, True, 900, 950, 60), (3, 0.6, False, 900, 950, 60),
                                                      (4, 0.75, True, 2000, 1900, 100), (5, 0.7, True, 300, 1200, 8), (6, 0.7, True, 40, 30, 4)])
def test_search_by_bow(ork, seed, nnratio, ori, nk, nf, nw):
    from orbx import abi
    s = sc.bow_scenario(seed, nk, nf, nw)
    KF, F = abi.Frame(s["kK"], s["dK"]), abi.Frame(s["kF"], s["dF"])
    on, om = ork.search_by_bow(KF, F, s["has"], s["fvK"], s["fvF"], nnratio, ori)
    rn, rm = ref.search_by_bow(KF, F, s["has"], s["fvK"], s["fvF"], nnratio, ori)
    assert on == rn and np.array_equal(om, rm)


@pytest.mark.parametrize("seed,stereo,th,nkp,nmp", [(1, True, 3.0, 900, 700), (2, False, 3.0, 900, 700), (3, True, 4.0, 2000, 3000),
                                                    (4, False, 2.5, 100, 50), (5, True, 3.0, 1000, 1)])
def test_fuse(ork, seed, stereo, th, nkp, nmp):
    """Equal except on the knife edge of MapPoint::PredictScale (src/MapPoint.cc:586): `ceil(log(ratio)/mfLogScaleFactor)` with
    a float `ratio` resolves to glibc's logf in the reference build (TemplatedVocabulary.h:36 leaks `using namespace std`),
    while the oracle and the device take the logarithm in double.  The scenario generator places map
    points at distances of exactly maxDistance / 1.2^k, where the two can land on different sides of the integer; every
    disagreement must be such a point, and there must be few."""
    from orbx import abi
    s = sc.fuse_scenario(seed, nkp, nmp, stereo, th)
    KF = abi.Frame(s["kK"], s["dK"], s["ur"])
    args = (KF, abi.make_camera(), s["R"], s["t"], s["Ow"], s["flags"], s["xw"], s["maxd"], s["mind"], s["normal"], s["desc"], s["th"],
            s["scale"], s["inv_sigma2"], s["log_sf"])
    on, ob = ork.fuse(*args)
    rn, rb = ref.fuse(*args)
    bad = np.flatnonzero(ob != rb)
    dist = np.linalg.norm(s["xw"].astype(np.float64) - s["Ow"].astype(np.float64), axis=1)
    frac = np.log(s["maxd"].astype(np.float64) / dist) / float(s["log_sf"])
    edge = np.abs(frac - np.rint(frac)) < 1e-5
    assert np.all(edge[bad]), "a disagreement away from the PredictScale knife edge: %s" % bad[~edge[bad]]
    assert len(bad) <= max(2, 0.004 * nmp)
    assert abs(on - rn) <= len(bad)
    assert np.array_equal(ob[~edge], rb[~edge])
