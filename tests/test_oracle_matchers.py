"""CPU: the oracle matchers against independent brute-force numpy restatements and structural properties
(the reference has no tests for them)."""
import numpy as np
import pytest

import scenarios as sc


@pytest.fixture(scope="module")
def frame(ork):
    from orbx import synth, abi
    L, R = synth.stereo_pair(3)
    exL, kL, dL = sc.extract_frame(ork, L)
    exR, kR, dR = sc.extract_frame(ork, R)
    pyrL = [exL.pyramid_level(l) for l in range(8)]
    pyrR = [exR.pyramid_level(l) for l in range(8)]
    ur, dp = ork.stereo_match(pyrL, pyrR, kL, dL, kR, dR, exL.scale, exL.inv_scale, sc.BF, sc.BF / sc.FX)
    return dict(kL=kL, dL=dL, kR=kR, dR=dR, ur=ur, dp=dp, pyrL=pyrL, pyrR=pyrR, scale=exL.scale, abi=abi)


def ham(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def test_descriptor_distance_is_popcount(ork):
    voc = sc.orbvoc()
    rng = np.random.default_rng(0)
    for _ in range(300):
        a, b = voc[rng.integers(len(voc))], voc[rng.integers(len(voc))]
        assert ork.descriptor_distance(a, b) == ham(a, b)
    assert ork.descriptor_distance(voc[0], voc[0]) == 0
    assert ork.descriptor_distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_features_in_area_matches_bruteforce_set_and_grid_order(ork, frame):
    F = frame["abi"].Frame(frame["kL"], frame["dL"], frame["ur"])
    k = frame["kL"]
    rng = np.random.default_rng(1)
    nq = 300
    x = rng.uniform(0, 752, nq).astype(np.float32)
    y = rng.uniform(0, 480, nq).astype(np.float32)
    r = rng.uniform(2, 70, nq).astype(np.float32)
    mn = rng.integers(-1, 5, nq).astype(np.int32)
    mx = (mn + rng.integers(0, 3, nq)).astype(np.int32)
    out, n = ork.features_in_area(F, x, y, r, mn, mx)
    wInv, hInv = np.float32(64) / np.float32(752), np.float32(48) / np.float32(480)
    cx = np.floor((k["x"] * wInv).astype(np.float32) + np.float32(0.5)).astype(int)   # round() for x >= 0
    cy = np.floor((k["y"] * hInv).astype(np.float32) + np.float32(0.5)).astype(int)
    for q in range(nq):
        inwin = (np.abs(k["x"] - x[q]) < r[q]) & (np.abs(k["y"] - y[q]) < r[q]) & (cx < 64) & (cy < 48)
        # GetFeaturesInArea only visits cells floor((x-r)*inv)..ceil((x+r)*inv): every in-window keypoint is inside
        chk = (mn[q] > 0) or (mx[q] >= 0)
        if chk:
            inwin &= k["octave"] >= mn[q]
            if mx[q] >= 0:
                inwin &= k["octave"] <= mx[q]
        got = out[q, :n[q]]
        assert set(got.tolist()) == set(np.flatnonzero(inwin).tolist()), q
        # order: (cell column, cell row, index) lexicographic
        key = [(cx[i], cy[i], i) for i in got]
        assert key == sorted(key), q


def _sbp_map_python(F_k, F_d, F_ur, s, th, nnratio):
    """Straight Python restatement of src/ORBmatcher.cc:59-255 using brute-force window search."""
    k = F_k
    wInv, hInv = np.float32(64) / np.float32(752), np.float32(48) / np.float32(480)
    cx = np.floor((k["x"] * wInv).astype(np.float32) + np.float32(0.5)).astype(int)
    cy = np.floor((k["y"] * hInv).astype(np.float32) + np.float32(0.5)).astype(int)
    blocked = s["kp_blocked"].copy()
    best_out = np.full(len(s["projX"]), -1, np.int32)
    n = 0
    for q in range(len(s["projX"])):
        if not (s["flags"][q] & 1):
            continue
        lvl = int(s["level"][q])
        r = np.float32(2.5) if float(s["viewCos"][q]) > 0.998 else np.float32(4.0)
        if th != 1.0:
            r = np.float32(r * np.float32(th))
        rs = np.float32(r * s["scaleFactors"][lvl])
        x, y = s["projX"][q], s["projY"][q]
        c = np.flatnonzero((np.abs(k["x"] - x) < rs) & (np.abs(k["y"] - y) < rs) & (k["octave"] >= lvl - 1)
                           & (k["octave"] <= lvl))
        c = sorted(c.tolist(), key=lambda i: (cx[i], cy[i], i))
        bd, bl, bd2, bl2, bi = 256, -1, 256, -1, -1
        for i in c:
            if blocked[i]:
                continue
            if F_ur is not None and F_ur[i] > 0 and abs(np.float32(s["projXR"][q] - F_ur[i])) > rs:
                continue
            d = ham(s["mpDesc"][q], F_d[i])
            if d < bd:
                bd2, bd, bl2, bl, bi = bd, d, bl, int(k["octave"][i]), i
            elif d < bd2:
                bl2, bd2 = int(k["octave"][i]), d
        if bd <= 100:
            if bl == bl2 and np.float32(bd) > np.float32(nnratio) * np.float32(bd2):
                continue
            best_out[q] = bi
            blocked[bi] = 1 if (s["flags"][q] & 2) else 0
            n += 1
    return n, best_out


@pytest.mark.parametrize("th,mono", [(1.0, False), (3.0, True)])
def test_search_by_projection_map_matches_python_restatement(ork, frame, th, mono):
    ur = None if mono else frame["ur"]
    F = frame["abi"].Frame(frame["kL"], frame["dL"], ur)
    s = sc.sbp_map_scenario(5, frame["kL"], frame["dL"], ur, nq=700)
    rn, rbest = ork.search_by_projection_map(F, s["kp_blocked"], s["projX"], s["projY"], s["projXR"], s["level"],
                                             s["viewCos"], s["mpDesc"], s["flags"], th, 0.8, s["scaleFactors"])
    pn, pbest = _sbp_map_python(frame["kL"], frame["dL"], ur, s, th, 0.8)
    assert rn == pn and np.array_equal(rbest, pbest)
    # structural: a keypoint claimed by a MapPoint with observations is never claimed again later
    claimed = {}
    for q, b in enumerate(rbest):
        if b >= 0:
            if b in claimed:
                assert not (s["flags"][claimed[b]] & 2)
            claimed[b] = q


def test_search_by_projection_frame_properties(ork, frame):
    abi = frame["abi"]
    cam = abi.make_camera()
    F = abi.Frame(frame["kL"], frame["dL"], frame["ur"])
    s = sc.sbp_frame_scenario(7, frame["kL"], frame["dL"], frame["ur"], frame["dp"])
    a = (F, s["cur_blocked"], cam, s["Tcw_cur"], s["Tcw_last"], s["flags"], s["xw"], s["octave"], s["angle"],
         s["mpDesc"], 7.0, False)
    n1, m1, k1, c1 = ork.search_by_projection_frame(*a, True, s["scaleFactors"])
    n0, m0, k0, c0 = ork.search_by_projection_frame(*a, False, s["scaleFactors"])
    assert np.array_equal(m0, m1)                       # the rotation filter only removes matches
    assert n0 == (m0 >= 0).sum() and n1 == k1.sum() and n1 < n0
    assert ((m1 >= 0) | (k1 == 0)).all() and not (s["flags"][m1 >= 0] & 1 == 0).any()
    assert not s["cur_blocked"][m1[m1 >= 0]].any()      # initially blocked keypoints are never taken
    for q in np.flatnonzero(m1 >= 0):                   # every match respects TH_HIGH
        assert ham(s["mpDesc"][q], frame["dL"][m1[q]]) <= 100
    # the kept matches concentrate in <= 3 rotation bins
    rot = (s["angle"][k1 == 1] - frame["kL"]["angle"][m1[k1 == 1]]) % 360
    bins = np.floor(rot * np.float32(1 / 30.0) + 0.5).astype(int) % 30
    assert len(np.unique(bins)) <= 3
    # final assignment array is consistent with the per-query outputs
    for idx in np.flatnonzero(c1 >= 0):
        assert m1[c1[idx]] == idx and k1[c1[idx]] == 1


def _tri_python(q, cam, coarse):
    k1, k2 = q["k1"], q["k2"]
    # F12 in float64 (tolerance-level cross-check of the fp32 oracle): count disagreements, must be tiny
    R1, t1, R2, t2 = q["R1w"].reshape(3, 3).astype(float), q["t1w"].astype(float), q["R2w"].reshape(3, 3).astype(float), q["t2w"].astype(float)
    R12 = R1 @ R2.T
    t12 = -R12 @ t2 + t1
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    K = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1]])
    F12 = np.linalg.inv(K).T @ tx @ R12 @ np.linalg.inv(K)
    Cw = -R1.T @ t1
    C2 = R2 @ Cw + t2
    ep = np.array([cam.fx * C2[0] / C2[2] + cam.cx, cam.fy * C2[1] / C2[2] + cam.cy])
    m12 = np.full(len(k1), -1, np.int32)
    ids1, off1, idx1 = q["fv1"]
    ids2, off2, idx2 = q["fv2"]
    pos2 = {int(n): j for j, n in enumerate(ids2)}
    for a, nid in enumerate(ids1):
        if int(nid) not in pos2:
            continue
        b = pos2[int(nid)]
        for i1 in idx1[off1[a]:off1[a + 1]]:
            if q["has1"][i1]:
                continue
            st1 = q["ur1"] is not None and q["ur1"][i1] >= 0
            best, bi = 50, -1
            for i2 in idx2[off2[b]:off2[b + 1]]:
                if q["has2"][i2]:
                    continue
                st2 = q["ur2"] is not None and q["ur2"][i2] >= 0
                d = ham(q["d1"][i1], q["d2"][i2])
                if d > 50 or d > best:
                    continue
                if not st1 and not st2:
                    de = ep - np.array([k2["x"][i2], k2["y"][i2]])
                    if de @ de < 100 * q["scaleFactors"][k2["octave"][i2]]:
                        continue
                ok = coarse
                if not ok:
                    l = np.array([k1["x"][i1], k1["y"][i1], 1.0]) @ F12
                    den = l[0] ** 2 + l[1] ** 2
                    if den != 0:
                        num = l[0] * k2["x"][i2] + l[1] * k2["y"][i2] + l[2]
                        ok = num * num / den < 3.84 * q["sigma2"][k2["octave"][i2]]
                if ok:
                    best, bi = d, i2
            if bi >= 0:
                m12[i1] = bi
    return m12


@pytest.mark.parametrize("coarse", [False, True])
def test_search_for_triangulation_matches_python_restatement(ork, frame, coarse):
    abi = frame["abi"]
    cam = abi.make_camera()
    q = sc.tri_scenario(9, frame["kL"], frame["dL"], frame["ur"])
    K1, K2 = abi.Frame(q["k1"], q["d1"], q["ur1"]), abi.Frame(q["k2"], q["d2"], q["ur2"])
    n, m12 = ork.search_for_triangulation(K1, K2, q["has1"], q["has2"], q["fv1"], q["fv2"], cam, cam, q["R1w"],
                                          q["t1w"], q["R2w"], q["t2w"], q["sigma2"], q["scaleFactors"], False, coarse,
                                          checkOri=False)
    py = _tri_python(q, cam, coarse)
    assert n == (m12 >= 0).sum() and n > 100
    # fp32 vs fp64 fundamental matrix: gate decisions may differ only for pairs on the 3.84*sigma2 boundary
    assert (m12 != py).sum() <= 2
    assert not q["has1"][m12 >= 0].any() and not q["has2"][m12[m12 >= 0]].any()
    n2, m2 = ork.search_for_triangulation(K1, K2, q["has1"], q["has2"], q["fv1"], q["fv2"], cam, cam, q["R1w"],
                                          q["t1w"], q["R2w"], q["t2w"], q["sigma2"], q["scaleFactors"], False, coarse,
                                          checkOri=True)
    assert n2 <= n and ((m2 == m12) | (m2 == -1)).all()


def test_stereo_match_properties(ork, frame):
    ur, dp, kL, kR = frame["ur"], frame["dp"], frame["kL"], frame["kR"]
    ok = ur >= 0
    assert ok.sum() > 300
    assert (dp[ok] > 0).all() and (dp[~ok] == -1).all() and (ur[~ok] == -1).all()
    disp = kL["x"][ok] - ur[ok]
    assert (disp >= 0).all() and (disp < sc.FX).all()
    assert np.allclose(dp[ok], sc.BF / np.maximum(disp, 0.01), rtol=1e-5)
    # the synthetic right image is the left one shifted by a per-band disparity bf/z in [3, 48] px
    assert (disp > 1.0).mean() > 0.95 and (disp < 52).all()
    # swapping left/right (negative disparities) yields (almost) no matches
    ur2, _ = ork.stereo_match(frame["pyrR"], frame["pyrL"], kR, frame["dR"], kL, frame["dL"], frame["scale"],
                              (1.0 / frame["scale"]).astype(np.float32), sc.BF, sc.BF / sc.FX)
    assert (ur2 >= 0).sum() < 0.1 * ok.sum()
