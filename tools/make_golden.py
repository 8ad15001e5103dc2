#!/usr/bin/env python3
"""Mint tests/golden/extractor_*.npz: known-answer vectors for the extractor.

The reference has no tests or golden vectors for this path, and it cannot be built
here, so the vectors come from tests/cv2_compose.py — ORBextractor::operator() re-composed in
Python from the *real* OpenCV primitives of cv2 4.13 (the only executable OpenCV available).
The C++ oracle and the CUDA path must both reproduce them bit for bit.

    python tools/make_golden.py        (run in the build container; needs cv2)

Round 2: the same vectors are now also REFERENCE OUTPUT.  `--source ref` mints them from oracle/_ref — the reference's
own src/ORBextractor.cc compiled unmodified against the OpenCV stand-in (oracle/Makefile, target _ref) — and
`--check` verifies that the committed files equal both sources bit for bit (they do; the fixtures were not changed).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "awesome-orb-slam3-3dvisioncraft-version_b200"))
import cv2_compose  # noqa: E402
from orbx import synth  # noqa: E402

CASES = {
    "scene0_stereo": dict(gen="scene_image", args=(0, 752, 480), lap=(0, 0), nfeatures=1000),
    "scene1_mono": dict(gen="scene_image", args=(1, 752, 480), lap=(0, 1000), nfeatures=1000),
    "noise_400x300": dict(gen="noise_image", args=(5, 400, 300), lap=(0, 0), nfeatures=1000),
    "scene_640x480_500": dict(gen="scene_image", args=(12, 640, 480), lap=(100, 300), nfeatures=500),
}


def from_ref(img, nfeatures, lap):
    sys.path.insert(0, ROOT)
    from oracle import ref
    rc, k, d, mono = ref.Extractor(nfeatures)(img, lap)
    assert rc == 0
    return k, d, mono


def main():
    source = "ref" if "--source" in sys.argv and sys.argv[sys.argv.index("--source") + 1] == "ref" else "cv2"
    check = "--check" in sys.argv
    pat = cv2_compose.load_pattern(os.path.join(ROOT, "oracle", "orb_pattern.inc"))
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, c in CASES.items():
        img = getattr(synth, c["gen"])(*c["args"])
        if check:
            g = np.load(os.path.join(out_dir, "extractor_%s.npz" % name))
            for src in ("cv2", "ref"):
                k, d, mono = (cv2_compose.extract(img, pat, nfeatures=c["nfeatures"], lap=c["lap"]) if src == "cv2"
                              else from_ref(img, c["nfeatures"], c["lap"]))
                same = (len(k) == len(g["keypoints"]) and all(np.array_equal(k[f], g["keypoints"][f]) for f in k.dtype.names)
                        and np.array_equal(d, g["descriptors"]) and mono == int(g["mono_index"]))
                print(name, src, "equals the committed fixture:", same)
                assert same
            continue
        if source == "ref":
            k, d, mono = from_ref(img, c["nfeatures"], c["lap"])
        else:
            k, d, mono = cv2_compose.extract(img, pat, nfeatures=c["nfeatures"], lap=c["lap"])
        np.savez_compressed(os.path.join(out_dir, "extractor_%s.npz" % name), gen=c["gen"], args=np.array(c["args"]),
                            lap=np.array(c["lap"]), nfeatures=c["nfeatures"],
                            image_sha256=hashlib.sha256(img.tobytes()).hexdigest(), keypoints=k, descriptors=d,
                            mono_index=mono)
        print(name, len(k), mono)


if __name__ == "__main__":
    main()
