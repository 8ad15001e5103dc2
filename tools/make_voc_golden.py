#!/usr/bin/env python3
"""Golden vectors of DBoW2 `transform` on the reference's real Vocabulary/ORBvoc.bin, computed with the numpy
brute-force restatement in tests/voc_util.py (not with the oracle).  Only runnable where /root/reference is mounted;
the result is committed as tests/golden/orbvoc_transform_golden.npz."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import voc_util as vu  # noqa: E402

V = vu.parse(open("/root/reference/Vocabulary/ORBvoc.bin", "rb").read())
sample = np.load(os.path.join(ROOT, "tests", "golden", "orbvoc_sample.npy"))
rng = np.random.default_rng(0)
q = sample[rng.choice(len(sample), 300, replace=False)].copy()
for r in range(100, 300):
    for b in rng.integers(0, 256, int(rng.integers(1, 51))):
        q[r, b >> 3] ^= np.uint8(1 << (b & 7))
out = vu.brute_transform(V, q, 4)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "orbvoc_transform_golden.npz"), desc=q, **out)
print({k: v.shape for k, v in out.items()})
