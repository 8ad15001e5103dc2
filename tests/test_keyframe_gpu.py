"""GPU parity of the keyframe-rate half of the batched mode (BASELINE.json config 5; SURVEY.md §8 a13 + a16 run for many
streams at once): the prepared many-problem plans orbx_tri_batch_* / orbx_lba_batch_* return, problem by problem,
exactly what the single-call entry points return (which test_matchers_gpu.py / test_optimizers_gpu.py compare with
the oracle) and what the oracle returns."""
import numpy as np
import pytest

import scenarios as sc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kf_frames(ork):
    from orbx import synth
    out = []
    for seed in (3, 8):
        L, R = synth.stereo_pair(seed)
        exL, kL, dL = sc.extract_frame(ork, L)
        exR, kR, dR = sc.extract_frame(ork, R)
        ur, dp = ork.stereo_match([exL.pyramid_level(l) for l in range(8)], [exR.pyramid_level(l) for l in range(8)], kL, dL,
                                  kR, dR, exL.scale, exL.inv_scale, sc.BF, sc.BF / sc.FX)
        out.append(dict(kL=kL, dL=dL, ur=ur))
    return out


def _tri_problems(kf_frames, n):
    import orbx
    cam = orbx.make_camera()
    probs, meta = [], None
    for q in range(n):
        f = kf_frames[q % len(kf_frames)]
        s = sc.tri_scenario(100 + q, f["kL"], f["dL"], f["ur"], baseline=0.15 + 0.05 * (q % 4), rot_deg=1.0 + (q % 3))
        if q % 5 == 4:    # an inconsistent pose: the epipolar gate rejects most pairs
            s["t2w"] = (s["t2w"] + np.array([0.0, 0.25, 0.1], np.float32)).astype(np.float32)
        probs.append(dict(KF1=orbx.Frame(s["k1"], s["d1"], s["ur1"]), KF2=orbx.Frame(s["k2"], s["d2"], s["ur2"]), has1=s["has1"],
                          has2=s["has2"], fv1=s["fv1"], fv2=s["fv2"], cam1=cam, cam2=cam, R1w=s["R1w"], t1w=s["t1w"], R2w=s["R2w"],
                          t2w=s["t2w"], only_stereo=(q % 7 == 6), coarse=(q % 6 == 5)))
        meta = (s["sigma2"], s["scaleFactors"])
    return probs, meta


def test_triangulation_batch_equals_single_calls_and_oracle(ctx, ork, kf_frames):
    import orbx
    probs, (sigma2, sf) = _tri_problems(kf_frames, 12)
    batch = orbx.TriangulationBatch(ctx, probs, sigma2, sf, True)
    batch.run()
    got = batch.fetch()
    batch.run()                                   # a prepared plan is re-runnable and deterministic
    again = batch.fetch()
    m = orbx.ORBmatcher(ctx, 0.6, True)
    total = 0
    for q, pr in enumerate(probs):
        a = (pr["KF1"], pr["KF2"], pr["has1"], pr["has2"], pr["fv1"], pr["fv2"], pr["cam1"], pr["cam2"], pr["R1w"], pr["t1w"], pr["R2w"],
             pr["t2w"], sigma2, sf, pr["only_stereo"], pr["coarse"])
        sn, sm = m.SearchForTriangulation(*a)
        on, om = ork.search_for_triangulation(*a, True)
        assert got[q][0] == sn == on, q
        assert np.array_equal(got[q][1], sm) and np.array_equal(sm, om), q
        assert again[q][0] == sn and np.array_equal(again[q][1], sm)
        total += sn
    assert total > 1000
    batch.close()


def test_triangulation_batch_handles_empty_feature_vectors(ctx, kf_frames):
    import orbx
    probs, (sigma2, sf) = _tri_problems(kf_frames, 2)
    empty = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    probs[1]["fv2"] = empty
    batch = orbx.TriangulationBatch(ctx, probs, sigma2, sf, True)
    batch.run()
    got = batch.fetch()
    assert got[0][0] > 50 and got[1][0] == 0 and np.all(got[1][1] == -1)
    batch.close()


@pytest.mark.parametrize("shapes", [[(6, 300, 2), (8, 500, 2), (6, 300, 2), (5, 120, 1)], [(20, 3000, 3), (20, 3000, 3)]])
def test_local_ba_batch_equals_single_calls(ctx, ork, shapes):
    import orbx
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    scen = [sc.lba_scenario(40 + i, K=K, M=M, n_fixed=nf) for i, (K, M, nf) in enumerate(shapes)]
    if len(shapes) == 4:      # one problem the sanity check rejects (>= 50 % bad): status 2, nothing written
        scen[2] = sc.lba_scenario(4, K=6, M=300, n_fixed=2, outlier_frac=0.9)
    batch = orbx.LocalBABatch(ctx, scen, cam)
    batch.run()
    got = batch.fetch()
    for q, s in enumerate(scen):
        a = (s["kf_T"], s["kf_fixed"], s["mp_xyz"], s["e_kf"], s["e_mp"], s["e_obs"], s["e_inv_sigma2"], cam)
        sT, sX, sbad, sit, sst = opt.LocalBundleAdjustment(*a)
        gT, gX, gbad, git, gst = got[q]
        assert gst == sst, q
        assert np.array_equal(git, sit), (q, git, sit)
        # the batch runs the single-CTA body, the single call the cooperative kernel: same canonical summation order,
        # hence the same bits
        assert np.array_equal(gT, sT) and np.array_equal(gX, sX) and np.array_equal(gbad, sbad), q
        if q == 0 and len(shapes) == 4:
            rT, rX, rbad, rit, rst = ork.local_ba(*a)
            assert rst == gst and np.array_equal(rit, git)
            assert np.abs(rT - gT).max() < 1e-5
    if len(shapes) == 4:
        assert got[2][4] == 2 and np.array_equal(got[2][0].reshape(-1, 16), scen[2]["kf_T"].reshape(-1, 16))
    batch.run()                                   # re-run from the pristine inputs: identical results
    again = batch.fetch()
    for q in range(len(scen)):
        assert np.array_equal(again[q][0], got[q][0]) and np.array_equal(again[q][1], got[q][1]) and np.array_equal(again[q][3], got[q][3])
    batch.close()
