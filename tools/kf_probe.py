#!/usr/bin/env python3
"""Device time of the keyframe-rate step for S streams: 10 x SearchForTriangulation + 1 LocalBundleAdjustment (20 KF /
3000 MP) per stream, as prepared many-problem plans (orbx_tri_batch_*, orbx_lba_batch_*).  CUDA-event timed.

    python tools/kf_probe.py [S]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "awesome-orb-slam3-3dvisioncraft-version_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import orbx  # noqa: E402
import scenarios as sc  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ctx = orbx.Context(0)
    cam = orbx.make_camera()
    kps, desc = sc.synthetic_keypoints(1, 1000)
    ur = np.where(np.arange(1000) % 2 == 0, kps["x"] - 10.0, -1.0).astype(np.float32)
    tri = []
    for q in range(10):
        s = sc.tri_scenario(200 + q, kps, desc, ur)
        tri.append(dict(KF1=orbx.Frame(s["k1"], s["d1"], s["ur1"]), KF2=orbx.Frame(s["k2"], s["d2"], s["ur2"]), has1=s["has1"], has2=s["has2"],
                        fv1=s["fv1"], fv2=s["fv2"], cam1=cam, cam2=cam, R1w=s["R1w"], t1w=s["t1w"], R2w=s["R2w"], t2w=s["t2w"]))
    lba = [sc.lba_scenario(i, K=20, M=3000, n_fixed=3) for i in range(4)]
    t0 = time.time()
    tb = orbx.TriangulationBatch(ctx, [tri[q % 10] for q in range(10 * S)], s["sigma2"], s["scaleFactors"], True)
    lb = orbx.LocalBABatch(ctx, [lba[p % 4] for p in range(S)], cam)
    prep = time.time() - t0
    st = torch.cuda.Stream()
    out = {"streams": S, "prepare_s": prep, "lba_pool_GB": lb.device_bytes / 1e9, "edges_per_lba": int(len(lba[0]["e_kf"]))}
    for name, plan in (("tri_x10", tb), ("local_ba", lb)):
        ms = []
        for it in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            plan.run(st.cuda_stream)
            b.record(st)
            b.synchronize()
            ms.append(a.elapsed_time(b))
        out[name + "_ms"] = ms
    r = lb.fetch()
    out["lba_iters"] = [r[i][3].tolist() for i in range(4)]
    out["lba_status"] = [r[i][4] for i in range(4)]
    out["tri_matches"] = [m[0] for m in tb.fetch()[:10]]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
