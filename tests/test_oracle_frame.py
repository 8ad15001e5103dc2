"""CPU: oracle isInFrustum batch / UndistortKeyPoints.  Undistortion is pinned to Python cv2 4.13's
cv2.undistortPoints (bit for bit); the frustum test to a straight numpy-float32 restatement of src/Frame.cc:571-650."""
import numpy as np
import pytest

import scenarios as sc
from orbx import abi

F32 = np.float32
EUROC_D = [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]   # Examples/Monocular/EuRoC.yaml:14-17


@pytest.mark.parametrize("dist", [EUROC_D, EUROC_D + [0.01], [0.3, -0.2, 0.001, -0.002, 0.05], [0.0, 0.1, 0.0, 0.0]])
def test_undistort_matches_cv2(ork, dist):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    xy = np.stack([rng.uniform(0, 752, 4000), rng.uniform(0, 480, 4000)], 1).astype(F32)
    cam = abi.make_camera()
    got = ork.undistort_points(xy, cam, dist)
    if dist[0] == 0.0:
        assert np.array_equal(got, xy)       # the reference's early return (src/Frame.cc:877-881)
        return
    K = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1]], F32)
    want = cv2.undistortPoints(xy.reshape(-1, 1, 2), K, np.array(dist, F32).reshape(-1, 1), None, None, K).reshape(-1, 2)
    assert np.array_equal(got, want)


def frustum_python(s, cos_limit, stale):
    R, t, Ow = s["R"], s["t"], s["Ow"]
    fx, fy, cx, cy, bf = map(F32, (sc.FX, sc.FY, sc.CX, sc.CY, sc.BF))
    n = len(s["maxd"])
    out = dict(in_view=np.zeros(n, np.uint8), proj_x=np.full(n, -1, F32), proj_y=np.full(n, -1, F32),
               proj_xr=stale["proj_xr"].copy(), depth=stale["depth"].copy(), level=stale["level"].copy(), view_cos=stale["view_cos"].copy())
    for i in range(n):
        X, Y, Z = s["xw"][i]
        xc = F32(F32(F32(F32(R[0] * X) + F32(R[1] * Y)) + F32(R[2] * Z)) + t[0])
        yc = F32(F32(F32(F32(R[3] * X) + F32(R[4] * Y)) + F32(R[5] * Z)) + t[1])
        zc = F32(F32(F32(F32(R[6] * X) + F32(R[7] * Y)) + F32(R[8] * Z)) + t[2])
        pcd = F32(np.sqrt(float(xc) * float(xc) + float(yc) * float(yc) + float(zc) * float(zc)))
        if zc < 0:
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            invz = F32(1) / zc
            u = F32(F32(F32(fx * xc) / zc) + cx)
            v = F32(F32(F32(fy * yc) / zc) + cy)
        if u < 0 or u > 752 or v < 0 or v > 480 or np.isnan(u) or np.isnan(v):
            continue
        out["proj_x"][i], out["proj_y"][i] = u, v
        PO = [F32(X - Ow[0]), F32(Y - Ow[1]), F32(Z - Ow[2])]
        d = F32(np.sqrt(float(PO[0]) ** 2 + float(PO[1]) ** 2 + float(PO[2]) ** 2))
        if d < F32(F32(0.8) * s["mind"][i]) or d > F32(F32(1.2) * s["maxd"][i]):
            continue
        nr = s["normal"][i]
        dot = float(PO[0]) * float(nr[0]) + float(PO[1]) * float(nr[1]) + float(PO[2]) * float(nr[2])
        vc = F32(dot / float(d))
        if vc < F32(cos_limit):
            continue
        lvl = int(np.ceil(np.log(float(F32(s["maxd"][i] / d))) / float(F32(s["log_sf"]))))
        lvl = 0 if lvl < 0 else (7 if lvl >= 8 else lvl)
        out["in_view"][i] = 1
        out["proj_xr"][i] = F32(u - F32(bf * invz))
        out["depth"][i], out["level"][i], out["view_cos"][i] = pcd, lvl, vc
    out["n"] = int(out["in_view"].sum())
    return out


@pytest.mark.parametrize("seed,cos_limit", [(1, 0.5), (2, 0.5), (3, 0.9)])
def test_is_in_frustum_matches_python(ork, seed, cos_limit):
    s = sc.fuse_scenario(seed, 400, 1500)
    rng = np.random.default_rng(seed)
    n = 1500
    stale = dict(proj_xr=rng.uniform(0, 700, n).astype(F32), depth=rng.uniform(1, 9, n).astype(F32),
                 level=rng.integers(0, 8, n).astype(np.int32), view_cos=rng.uniform(0, 1, n).astype(F32))
    got = ork.is_in_frustum(abi.make_camera(), s["R"], s["t"], s["Ow"], (0.0, 752.0, 0.0, 480.0), cos_limit, 8, s["log_sf"], s["xw"],
                            s["maxd"], s["mind"], s["normal"], stale)
    want = frustum_python(s, cos_limit, stale)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    assert 100 < got["n"] < n
