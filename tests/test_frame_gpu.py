"""GPU parity of the per-frame glue: isInFrustum batch and UndistortKeyPoints == oracle, bit for bit."""
import numpy as np
import pytest

import orbx
import scenarios as sc
from orbx import abi

pytestmark = pytest.mark.gpu
EUROC_D = [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]


@pytest.mark.parametrize("dist", [EUROC_D, EUROC_D + [0.01], [0.3, -0.2, 0.001, -0.002, 0.05], [0.0, 0.1, 0.0, 0.0]])
@pytest.mark.parametrize("n", [0, 1, 1000, 5000])
def test_undistort_matches_oracle(ctx, ork, dist, n):
    rng = np.random.default_rng(n)
    xy = np.stack([rng.uniform(-50, 800, n), rng.uniform(-50, 530, n)], 1).astype(np.float32)
    cam = abi.make_camera()
    assert np.array_equal(orbx.undistort_keypoints(ctx, xy, cam, dist), ork.undistort_points(xy, cam, dist))


@pytest.mark.parametrize("seed,cos_limit,nmp", [(1, 0.5, 1500), (2, 0.5, 3000), (3, 0.9, 1500), (4, 0.5, 1)])
def test_is_in_frustum_matches_oracle(ctx, ork, seed, cos_limit, nmp):
    s = sc.fuse_scenario(seed, 300, nmp)
    rng = np.random.default_rng(seed)
    stale = dict(proj_xr=rng.uniform(0, 700, nmp).astype(np.float32), depth=rng.uniform(1, 9, nmp).astype(np.float32),
                 level=rng.integers(0, 8, nmp).astype(np.int32), view_cos=rng.uniform(0, 1, nmp).astype(np.float32))
    args = (abi.make_camera(), s["R"], s["t"], s["Ow"], (0.0, 752.0, 0.0, 480.0), cos_limit, 8, s["log_sf"], s["xw"], s["maxd"],
            s["mind"], s["normal"], stale)
    want = ork.is_in_frustum(*args)
    got = orbx.is_in_frustum(ctx, *args)
    for k in want:
        assert np.array_equal(got[k], want[k]), k


def test_frustum_feeds_search_by_projection(ctx, ork):
    """f4 -> a11 chained: the track fields computed on the device are exactly what SearchByProjection(F, MapPoints) consumes."""
    s = sc.fuse_scenario(5, 900, 1200)
    cam = abi.make_camera()
    fr = orbx.is_in_frustum(ctx, cam, s["R"], s["t"], s["Ow"], (0.0, 752.0, 0.0, 480.0), 0.5, 8, s["log_sf"], s["xw"], s["maxd"],
                            s["mind"], s["normal"])
    F = abi.Frame(s["kK"], s["dK"], s["ur"])
    flags = (fr["in_view"] | 2).astype(np.uint8)
    m = orbx.ORBmatcher(ctx, 0.8)
    blocked = np.zeros(len(s["kK"]), np.uint8)
    gn, gbest = m.SearchByProjectionMap(F, blocked, fr["proj_x"], fr["proj_y"], fr["proj_xr"], fr["level"], fr["view_cos"], s["desc"],
                                     flags, 3.0, s["scale"])
    wn, wbest = ork.search_by_projection_map(F, blocked, fr["proj_x"], fr["proj_y"], fr["proj_xr"], fr["level"], fr["view_cos"],
                                             s["desc"], flags, 3.0, 0.8, s["scale"])
    assert gn == wn and np.array_equal(gbest, wbest) and gn > 50
