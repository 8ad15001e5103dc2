#!/usr/bin/env python3
"""One call of every SURVEY §8 f-row entry point (f1 vocabulary transform, f2 SearchByBoW + Fuse, f4 frustum + undistort)
and of LocalBundleAdjustment at realistic sizes — a target for `ncu -k regex:...`."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "awesome-orb-slam3-3dvisioncraft-version_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import orbx  # noqa: E402
import scenarios as sc  # noqa: E402
import voc_util as vu  # noqa: E402

ctx = orbx.Context(0)
cam = orbx.make_camera()
vb = vu.make_vocabulary(3, 10, 5, ragged=False, p_early_leaf=0.0)
V = vu.parse(vb)
dv = orbx.ORBVocabulary(ctx, vb)
q = vu.query_descriptors(V, 1, 1000, 30)
for _ in range(2):
    dv.transform(q, 3)
bs = sc.bow_scenario(1)
KF, F = orbx.Frame(bs["kK"], bs["dK"]), orbx.Frame(bs["kF"], bs["dF"])
for _ in range(2):
    orbx.search_by_bow(ctx, KF, F, bs["has"], bs["fvK"], bs["fvF"], 0.7, True)
fs = sc.fuse_scenario(1, 1000, 1500)
KFf = orbx.Frame(fs["kK"], fs["dK"], fs["ur"])
for _ in range(2):
    orbx.fuse(ctx, KFf, cam, fs["R"], fs["t"], fs["Ow"], fs["flags"], fs["xw"], fs["maxd"], fs["mind"], fs["normal"], fs["desc"], 3.0,
              fs["scale"], fs["inv_sigma2"], fs["log_sf"])
    orbx.is_in_frustum(ctx, cam, fs["R"], fs["t"], fs["Ow"], (0.0, 752.0, 0.0, 480.0), 0.5, 8, fs["log_sf"], fs["xw"], fs["maxd"],
                       fs["mind"], fs["normal"])
    orbx.undistort_keypoints(ctx, np.stack([fs["kK"]["x"], fs["kK"]["y"]], 1), cam, [-0.2834, 0.0739, 0.00019, 1.76e-05])
l = sc.lba_scenario(0)
opt = orbx.Optimizer(ctx)
for _ in range(2):
    opt.LocalBundleAdjustment(l["kf_T"], l["kf_fixed"], l["mp_xyz"], l["e_kf"], l["e_mp"], l["e_obs"], l["e_inv_sigma2"], cam)
print("ok")
