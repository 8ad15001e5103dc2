import glob
import hashlib
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def extractor_cases():
    from orbx import synth
    for path in sorted(glob.glob(os.path.join(GOLDEN, "extractor_*.npz"))):
        g = np.load(path)
        img = getattr(synth, str(g["gen"]))(*[int(v) for v in g["args"]])
        assert hashlib.sha256(img.tobytes()).hexdigest() == str(g["image_sha256"]), \
            "synthetic generator drifted from the fixture: " + path
        yield (os.path.basename(path), img, tuple(int(v) for v in g["lap"]), int(g["nfeatures"]), g["keypoints"],
               g["descriptors"], int(g["mono_index"]))


def assert_kp_equal(gk, gd, gm, rk, rd, rm, tag):
    assert len(gk) == len(rk), "%s: %d keypoints vs golden %d" % (tag, len(gk), len(rk))
    assert gm == rm, "%s: monoIndex %d vs %d" % (tag, gm, rm)
    for f in rk.dtype.names:
        assert np.array_equal(gk[f], rk[f]), "%s: field %s differs" % (tag, f)
    assert np.array_equal(gd, rd), "%s: descriptors differ" % tag
