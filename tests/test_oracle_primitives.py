"""CPU: pin the oracle's restated OpenCV primitives to cv2 4.13, bit for bit.
cv2 only exists in the build container; on a box without it these tests skip (the golden-vector
tests still pin the oracle there)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def imgs():
    from orbx import synth
    return [synth.scene_image(1), synth.noise_image(3), synth.scene_image(4, 333, 211)]


def test_resize_chain_matches_cv2(ork, imgs):
    for img in imgs:
        prev = img
        h, w = img.shape
        for l in range(1, 8):
            s = np.float32(1.0) / np.float32(1.2 ** l)
            dw, dh = int(np.rint(np.float32(w) * s)), int(np.rint(np.float32(h) * s))
            ref = cv2.resize(prev, (dw, dh), interpolation=cv2.INTER_LINEAR)
            got = ork.resize_linear(prev, dw, dh)
            assert np.array_equal(ref, got), "level %d" % l
            prev = ref


def test_resize_upscale_and_odd_ratios(ork):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53)).astype(np.uint8)
    for dw, dh in [(53, 37), (60, 40), (106, 74), (17, 11), (52, 36), (200, 5)]:
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ref, ork.resize_linear(img, dw, dh)), (dw, dh)


def test_gaussian_blur_matches_cv2(ork, imgs):
    for img in imgs + [imgs[0][:9, :9].copy(), imgs[1][:7, :40].copy()]:
        ref = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(ref, ork.gaussian_blur7(img))


def _cv_fast(img, th):
    f = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                       type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kp = f.detect(img)
    return np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in kp], np.int32).reshape(-1, 3)


def test_fast_whole_image_matches_cv2(ork, imgs):
    for img in imgs:
        for th in (20, 7, 40):
            assert np.array_equal(_cv_fast(img, th), ork.fast(img, th)), th


def test_fast_small_rois_match_cv2(ork, imgs):
    """The reference calls cv::FAST on ~36x36 ROIs (src/ORBextractor.cc:808); ROI semantics matter."""
    rng = np.random.default_rng(1)
    img = imgs[0]
    for _ in range(400):
        x0, y0 = int(rng.integers(0, 700)), int(rng.integers(0, 430))
        w, h = int(rng.integers(4, 46)), int(rng.integers(4, 46))
        roi = np.ascontiguousarray(img[y0:y0 + h, x0:x0 + w])
        th = int(rng.choice([7, 20]))
        assert np.array_equal(_cv_fast(roi, th), ork.fast(roi, th)), (x0, y0, w, h, th)


def test_fast_atan2_matches_cv2_scalar(ork):
    rng = np.random.default_rng(2)
    y = rng.normal(0, 3000, 20000).astype(np.float32)
    x = rng.normal(0, 3000, 20000).astype(np.float32)
    y[:64] = 0
    x[32:96] = 0
    y[96:128] = x[96:128]
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert np.array_equal(ref, ork.fast_atan2(y, x))
