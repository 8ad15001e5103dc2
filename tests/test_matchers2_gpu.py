"""GPU parity of SearchByBoW / Fuse: device == oracle, bit for bit, through the C ABI."""
import numpy as np
import pytest

import orbx
import scenarios as sc
from orbx import abi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,nnratio,ori,nk,nf,nw", [(1, 0.7, True, 900, 950, 60), (2, 0.9, True, 900, 950, 60), (3, 0.6, False, 900, 950, 60),
                                                      (4, 0.75, True, 2000, 1900, 100), (5, 0.7, True, 300, 1200, 8), (6, 0.7, True, 40, 30, 4)])
def test_search_by_bow_matches_oracle(ctx, ork, seed, nnratio, ori, nk, nf, nw):
    s = sc.bow_scenario(seed, nk, nf, nw)
    KF, F = abi.Frame(s["kK"], s["dK"]), abi.Frame(s["kF"], s["dF"])
    wn, wm = ork.search_by_bow(KF, F, s["has"], s["fvK"], s["fvF"], nnratio, ori)
    gn, gm = orbx.search_by_bow(ctx, KF, F, s["has"], s["fvK"], s["fvF"], nnratio, ori)
    assert gn == wn and np.array_equal(gm, wm)


def test_search_by_bow_empty_inputs(ctx):
    s = sc.bow_scenario(9, n_kf=50, n_f=40, nwords=6)
    KF, F = abi.Frame(s["kK"], s["dK"]), abi.Frame(s["kF"], s["dF"])
    empty = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    assert orbx.search_by_bow(ctx, KF, F, s["has"], empty, s["fvF"])[0] == 0
    assert orbx.search_by_bow(ctx, KF, F, s["has"], s["fvK"], empty)[0] == 0
    n, m = orbx.search_by_bow(ctx, KF, F, np.zeros_like(s["has"]), s["fvK"], s["fvF"])
    assert n == 0 and np.all(m == -1)


def test_bow_pipeline_vocabulary_to_matches(ctx, ork):
    """f1 -> f2 chained on the device side of the ABI: FeatureVectors from orbx_vocabulary_transform feed
    orbx_search_by_bow; every stage equals the oracle's."""
    import oracle
    import voc_util as vu
    vb = vu.make_vocabulary(41, 10, 4, p_stop=0.02)
    V = vu.parse(vb)
    rng = np.random.default_rng(0)
    kK, _ = sc.synthetic_keypoints(71, 800)
    kF, _ = sc.synthetic_keypoints(72, 800)
    dK = vu.query_descriptors(V, 1, 800, 30)
    dF = dK[rng.permutation(800)].copy()
    for r in range(800):
        for b in rng.integers(0, 256, int(rng.integers(0, 12))):
            dF[r, b >> 3] ^= np.uint8(1 << (b & 7))
    dv, ov = orbx.ORBVocabulary(ctx, vb), oracle.Vocabulary(vb)
    gK, gF = dv.transform(dK, 2), dv.transform(dF, 2)
    wK, wF = ov.transform(dK, 2), ov.transform(dF, 2)
    for g, w in ((gK, wK), (gF, wF)):
        assert all(np.array_equal(a, w[k]) for a, k in zip(g[2:], ("fv_node", "fv_off", "fv_idx")))
    KF, F = abi.Frame(kK, dK), abi.Frame(kF, dF)
    has = np.ones(800, np.uint8)
    gn, gm = orbx.search_by_bow(ctx, KF, F, has, gK[2:], gF[2:], 0.8, False)
    wn, wm = ork.search_by_bow(KF, F, has, (wK["fv_node"], wK["fv_off"], wK["fv_idx"]), (wF["fv_node"], wF["fv_off"], wF["fv_idx"]), 0.8, False)
    assert gn == wn and np.array_equal(gm, wm) and gn > 300


@pytest.mark.parametrize("seed,stereo,th,nkp,nmp", [(1, True, 3.0, 900, 700), (2, False, 3.0, 900, 700), (3, True, 4.0, 2000, 3000),
                                                    (4, False, 2.5, 100, 50), (5, True, 3.0, 1000, 1)])
def test_fuse_matches_oracle(ctx, ork, seed, stereo, th, nkp, nmp):
    s = sc.fuse_scenario(seed, nkp, nmp, stereo, th)
    KF = abi.Frame(s["kK"], s["dK"], s["ur"])
    cam = abi.make_camera()
    args = (KF, cam, s["R"], s["t"], s["Ow"], s["flags"], s["xw"], s["maxd"], s["mind"], s["normal"], s["desc"], s["th"],
            s["scale"], s["inv_sigma2"], s["log_sf"])
    wn, wb = ork.fuse(*args)
    gn, gb = orbx.fuse(ctx, *args)
    assert gn == wn and np.array_equal(gb, wb)


def test_fuse_no_candidates(ctx):
    s = sc.fuse_scenario(7, 200, 100)
    KF = abi.Frame(s["kK"], s["dK"], s["ur"])
    n, b = orbx.fuse(ctx, KF, abi.make_camera(), s["R"], s["t"], s["Ow"], np.zeros(100, np.uint8), s["xw"], s["maxd"], s["mind"],
                     s["normal"], s["desc"], 3.0, s["scale"], s["inv_sigma2"], s["log_sf"])
    assert n == 0 and np.all(b == -1)
