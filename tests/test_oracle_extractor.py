"""CPU: the C++ oracle extractor against (a) the committed golden vectors and (b) where cv2 is
importable, the independent cv2-composed restatement on further images; plus host-logic checks."""
import numpy as np
import pytest

from golden_util import extractor_cases, assert_kp_equal


def test_oracle_reproduces_golden_vectors(ork):
    n = 0
    for name, img, lap, nf, gk, gd, gm in extractor_cases():
        rc, k, d, m = ork.Extractor(nf)(img, lap)
        assert rc == 0
        assert_kp_equal(k, d, m, gk, gd, gm, name)
        n += 1
    assert n >= 4


def test_oracle_matches_cv2_composition(ork):
    pytest.importorskip("cv2")
    import os
    import cv2_compose
    from orbx import synth
    pat = cv2_compose.load_pattern(os.path.join(os.path.dirname(__file__), "..", "oracle", "orb_pattern.inc"))
    cases = [(synth.scene_image(31, 480, 360), (0, 0), 700), (synth.stereo_pair(4)[1], (0, 1000), 1000),
             (synth.constant_image(17), (0, 0), 1000)]
    for img, lap, nf in cases:
        k2, d2, m2 = cv2_compose.extract(img, pat, nfeatures=nf, lap=lap)
        rc, k, d, m = ork.Extractor(nf)(img, lap)
        assert rc == 0
        assert_kp_equal(k, d, m, k2, d2, m2, "cv2-compose")


def test_constructor_tables(ork):
    """mnFeaturesPerLevel / scale tables / umax for the EuRoC settings (SURVEY.md §8 derived sizes)."""
    e = ork.Extractor(1000, 1.2, 8, 20, 7)
    assert e.features_per_level.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert ork.Extractor(2000).features_per_level.tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert ork.Extractor(5000).features_per_level.tolist() == [1086, 905, 754, 628, 524, 436, 364, 303]
    assert e.umax.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert np.allclose(e.scale, 1.2 ** np.arange(8), rtol=1e-6)
    from orbx import synth
    e(synth.scene_image(0))
    assert [e.pyramid_level(l).shape[::-1] for l in range(8)] == [(752, 480), (627, 400), (522, 333), (435, 278),
                                                                 (363, 231), (302, 193), (252, 161), (210, 134)]


def test_edge_cases(ork):
    from orbx import synth
    e = ork.Extractor()
    rc, k, d, m = e(np.empty((0, 0), np.uint8))
    assert rc == -1 and len(k) == 0
    rc, k, d, m = e(synth.constant_image(0))
    assert rc == 0 and len(k) == 0 and m == 0
    rc, k, d, m = e(synth.scene_image(1, 200, 150))   # last level narrower than one 30-px FAST cell
    assert rc == -2
    # octree never returns more than N+2 per level, always >= 1 keypoint per occupied root
    rc, k, d, m = e(synth.noise_image(9))
    assert rc == 0
    per = np.bincount(k["octave"], minlength=8)
    assert (per <= e.features_per_level + 2).all() and per.sum() == len(k)
    # keypoints stay EDGE_THRESHOLD(19) pixels inside their level
    for l in range(8):
        kl = k[k["octave"] == l]
        w, h = e.pyramid_level(l).shape[::-1]
        x, y = np.rint(kl["x"] / e.scale[l]), np.rint(kl["y"] / e.scale[l])
        assert (x >= 19).all() and (x < w - 19).all() and (y >= 19).all() and (y < h - 19).all()


def test_cosf_vs_double_rotation_is_documented_choice(ork):
    """The reference's `cos(angle)` resolves to glibc cosf (src/ORBextractor.cc:110-111); the oracle and
    the device round the double result instead.  Quantify the consequence: the two
    rotations differ by 1 ulp for ~1 % of angles, and that flips a sampled pixel coordinate
    (cvRound of x*b+y*a) for a vanishing fraction of (angle, pattern point) pairs."""
    import ctypes
    import os
    import cv2_compose
    libm = ctypes.CDLL("libm.so.6")
    for f in (libm.cosf, libm.sinf):
        f.restype, f.argtypes = ctypes.c_float, [ctypes.c_float]
    ang = np.linspace(0, 360, 20001, dtype=np.float32) * np.float32(np.pi / 180.0)
    c_f = np.array([libm.cosf(float(a)) for a in ang], np.float32)
    s_f = np.array([libm.sinf(float(a)) for a in ang], np.float32)
    c_d = np.cos(ang.astype(np.float64)).astype(np.float32)
    s_d = np.sin(ang.astype(np.float64)).astype(np.float32)
    assert (c_f != c_d).mean() < 0.03 and (s_f != s_d).mean() < 0.03
    assert np.abs(c_f.astype(np.float64) - c_d).max() <= 1.2e-7
    pat = cv2_compose.load_pattern(os.path.join(os.path.dirname(__file__), "..", "oracle", "orb_pattern.inc"))
    px, py = pat[:, 0].astype(np.float32)[None, :], pat[:, 1].astype(np.float32)[None, :]
    flips = 0
    for (c, s) in ((c_f, s_f),):
        r1 = np.rint(px * s[:, None] + py * c[:, None])
        r2 = np.rint(px * s_d[:, None] + py * c_d[:, None])
        q1 = np.rint(px * c[:, None] - py * s[:, None])
        q2 = np.rint(px * c_d[:, None] - py * s_d[:, None])
        flips += int((r1 != r2).sum() + (q1 != q2).sum())
    assert flips <= 20, flips   # out of 2 * 20001 * 512 coordinates
