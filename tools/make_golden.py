#!/usr/bin/env python3
"""Mint tests/golden/extractor_*.npz: known-answer vectors for the extractor.

The reference has no tests or golden vectors for this path, and it cannot be built
here, so the vectors come from tests/cv2_compose.py — ORBextractor::operator() re-composed in
Python from the *real* OpenCV primitives of cv2 4.13 (the only executable OpenCV available).
The C++ oracle and the CUDA path must both reproduce them bit for bit.

    python tools/make_golden.py        (run in the build container; needs cv2)
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


def main():
    pat = cv2_compose.load_pattern(os.path.join(ROOT, "oracle", "orb_pattern.inc"))
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, c in CASES.items():
        img = getattr(synth, c["gen"])(*c["args"])
        k, d, mono = cv2_compose.extract(img, pat, nfeatures=c["nfeatures"], lap=c["lap"])
        np.savez_compressed(os.path.join(out_dir, "extractor_%s.npz" % name), gen=c["gen"], args=np.array(c["args"]),
                            lap=np.array(c["lap"]), nfeatures=c["nfeatures"],
                            image_sha256=hashlib.sha256(img.tobytes()).hexdigest(), keypoints=k, descriptors=d,
                            mono_index=mono)
        print(name, len(k), mono)


if __name__ == "__main__":
    main()
