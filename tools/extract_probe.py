#!/usr/bin/env python3
"""Minimal extractor workload for ncu: B synthetic 752x480 images, two extract_batch_device-equivalent calls.
python tools/extract_probe.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "awesome-orb-slam3-3dvisioncraft-version_b200"))
import numpy as np  # noqa: E402


def main():
    import orbx
    from orbx import synth
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    base = []
    for s in range(8):
        L, R = synth.stereo_pair(100 + s)
        base += [L, R]
    imgs = [base[i % 16] for i in range(B)]
    ctx = orbx.Context(0)
    ex = orbx.ORBextractor(ctx, 1000, 1.2, 8, 20, 7, max_w=752, max_h=480, max_batch=B)
    for _ in range(2):
        out = ex.extract_batch(imgs)
    print("keypoints of image 0:", len(out[0][1]))


if __name__ == "__main__":
    main()
