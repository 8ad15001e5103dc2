"""GPU parity: liborbx PoseOptimization / LocalBundleAdjustment (C ABI) vs the CPU oracle.

fp64 on both sides with identical expressions; only the summation order differs (device: fixed-shape
trees; oracle: edge order), so poses must agree far inside the 1e-4 rad / 1e-3 m target, with the same
LM iteration counts and the same inlier/outlier classification."""
import numpy as np
import pytest

import scenarios as sc

pytestmark = pytest.mark.gpu

ROT_TOL_RAD, T_TOL_M = 1e-4, 1e-3          # north_star tolerance
TIGHT = 2e-6                               # what identical arithmetic actually delivers (float32 outputs)


def _pose_delta(Ta, Tb):
    Ra, Rb = Ta[:3, :3].astype(float), Tb[:3, :3].astype(float)
    # small-angle rotation distance ||Ra - Rb||_F / sqrt(2) (arccos of the trace is ill-conditioned near 0
    # on float32 matrices: a 1e-7 rounding of the trace already reads as 4e-4 rad)
    ang = np.linalg.norm(Ra - Rb) / np.sqrt(2.0)
    return ang, np.linalg.norm(Ta[:3, 3].astype(float) - Tb[:3, 3].astype(float))


@pytest.mark.parametrize("E,stereo_frac", [(150, 0.7), (300, 0.7), (500, 0.7), (300, 0.0), (300, 1.0)])
def test_pose_optimization_matches_oracle(ctx, ork, E, stereo_frac):
    import orbx
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    flips = 0
    for seed in range(12):
        s = sc.pose_opt_scenario(100 * E + seed, E=E, stereo_frac=stereo_frac)
        rT, rout, rn, rit = ork.pose_optimization(s["xw"], s["obs"], s["inv_sigma2"], cam, s["Tcw"])
        gT, gout, gn, git = opt.PoseOptimization(s["xw"], s["obs"], s["inv_sigma2"], cam, s["Tcw"])
        assert np.array_equal(git, rit), (seed, git, rit)            # same iteration count, every round
        ang, dt = _pose_delta(gT, rT)
        assert ang < ROT_TOL_RAD and dt < T_TOL_M
        assert ang < TIGHT and dt < TIGHT, (seed, ang, dt)
        flips += int((gout != rout).sum())
        assert abs(gn - rn) <= int((gout != rout).sum())
        # and the optimiser did its job: close to the ground truth, gross outliers rejected
        ang_gt, dt_gt = _pose_delta(gT, s["Tgt"].astype(np.float32))
        assert ang_gt < np.radians(0.5) and dt_gt < 0.03
        assert not (s["is_outlier"] & (gout == 0)).any()
    assert flips == 0, "chi2-threshold decisions flipped for %d edges" % flips


def test_pose_optimization_edge_cases(ctx, ork):
    import orbx
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    s = sc.pose_opt_scenario(1, E=300)
    for E in (0, 2, 3, 9, 10):        # <3: untouched pose, returns 0 ; <10 edges: a single round
        sl = slice(0, E)
        rT, rout, rn, rit = ork.pose_optimization(s["xw"][sl], s["obs"][sl], s["inv_sigma2"][sl], cam, s["Tcw"])
        gT, gout, gn, git = opt.PoseOptimization(s["xw"][sl], s["obs"][sl], s["inv_sigma2"][sl], cam, s["Tcw"])
        assert gn == rn and np.array_equal(git, rit) and np.array_equal(gout, rout), E
        assert np.allclose(gT, rT, atol=1e-5)
        if E < 3:
            assert gn == 0 and np.array_equal(gT, s["Tcw"].reshape(4, 4))
        if 3 <= E < 10:
            assert git[1:].sum() == 0


@pytest.mark.parametrize("K,M,nfixed", [(6, 300, 2), (20, 3000, 3)])
def test_local_ba_matches_oracle(ctx, ork, K, M, nfixed):
    import orbx
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    for seed in range(2):
        s = sc.lba_scenario(seed, K=K, M=M, n_fixed=nfixed)
        a = (s["kf_T"], s["kf_fixed"], s["mp_xyz"], s["e_kf"], s["e_mp"], s["e_obs"], s["e_inv_sigma2"], cam)
        rT, rX, rbad, rit, rst = ork.local_ba(*a)
        gT, gX, gbad, git, gst = opt.LocalBundleAdjustment(*a)
        assert gst == rst == 0
        assert np.array_equal(git, rit), (git, rit)
        for k in range(K):
            ang, dt = _pose_delta(gT[k], rT[k])
            assert ang < ROT_TOL_RAD and dt < T_TOL_M
            assert ang < 1e-5 and dt < 1e-5, (k, ang, dt)
        # points: compare the well-constrained ones (a monocular point seen under a tiny parallax has a
        # near-singular 3x3 block; its update is rounding noise on both sides)
        good = np.linalg.norm(rX - s["Pgt"], axis=1) < 1.0
        assert good.mean() > 0.9
        assert np.abs(gX - rX)[good].max() < 1e-3
        assert np.median(np.abs(gX - rX)) < 1e-5
        assert (gbad != rbad).mean() < 1e-3
        # fixed keyframes are untouched, free ones moved towards the ground truth
        Tin = s["kf_T"].reshape(-1, 4, 4)
        assert np.array_equal(gT[:nfixed], Tin[:nfixed])
        before = np.mean([_pose_delta(Tin[k], s["Tgt"][k].astype(np.float32))[0] for k in range(nfixed, K)])
        after = np.mean([_pose_delta(gT[k], s["Tgt"][k].astype(np.float32))[0] for k in range(nfixed, K)])
        assert after < 0.25 * before


def test_local_ba_stop_flag_and_sanity_abort(ctx, ork):
    import orbx
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    s = sc.lba_scenario(3, K=6, M=300, n_fixed=2)
    a = (s["kf_T"], s["kf_fixed"], s["mp_xyz"], s["e_kf"], s["e_mp"], s["e_obs"], s["e_inv_sigma2"], cam)
    stop = np.ones(1, np.uint8)                     # mbAbortBA already set: nothing happens
    gT, gX, gbad, git, gst = opt.LocalBundleAdjustment(*a, stop=stop)
    rT, rX, rbad, rit, rst = ork.local_ba(*a, stop=stop)
    assert gst == rst == 1 and git.sum() == 0
    assert np.array_equal(gT.reshape(-1, 16), s["kf_T"]) and np.array_equal(gX, s["mp_xyz"])
    stop[0] = 0                                     # flag present but clear: identical to no flag
    gT2, gX2, _, git2, gst2 = opt.LocalBundleAdjustment(*a, stop=stop)
    gT3, gX3, _, git3, gst3 = opt.LocalBundleAdjustment(*a)
    assert gst2 == gst3 == 0 and np.array_equal(git2, git3) and np.array_equal(gT2, gT3) and np.array_equal(gX2, gX3)
    # >= 50 % bad observations: the reference returns without writing anything back
    bad = sc.lba_scenario(4, K=6, M=300, n_fixed=2, outlier_frac=0.9)
    b = (bad["kf_T"], bad["kf_fixed"], bad["mp_xyz"], bad["e_kf"], bad["e_mp"], bad["e_obs"], bad["e_inv_sigma2"], cam)
    gT, gX, gbad, git, gst = opt.LocalBundleAdjustment(*b)
    rT, rX, rbad, rit, rst = ork.local_ba(*b)
    assert gst == rst == 2
    assert np.array_equal(gT.reshape(-1, 16), bad["kf_T"]) and np.array_equal(gX, bad["mp_xyz"])


def test_lba_is_deterministic(ctx):
    import orbx
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    s = sc.lba_scenario(5, K=8, M=500, n_fixed=2)
    a = (s["kf_T"], s["kf_fixed"], s["mp_xyz"], s["e_kf"], s["e_mp"], s["e_obs"], s["e_inv_sigma2"], cam)
    r1 = opt.LocalBundleAdjustment(*a)
    r2 = opt.LocalBundleAdjustment(*a)
    assert np.array_equal(r1[0], r2[0]) and np.array_equal(r1[1], r2[1]) and np.array_equal(r1[2], r2[2])


def _inertial_call(opt, s, cam, rec_init=False):
    return opt.PoseInertialOptimizationLastKeyFrame(s["xw"], s["obs"], s["isg"], s["close"], cam, s["Tcw"], s["Tcb"], s["Tbc"],
                                                    s["state"], s["kf"], s["preint"], s["infoI"], s["infoG"], s["infoA"],
                                                    rec_init=rec_init)


@pytest.mark.parametrize("E,stereo_frac", [(300, 0.6), (150, 0.0), (500, 1.0), (60, 0.5), (1000, 0.7)])
def test_pose_inertial_optimization_matches_oracle(ctx, ork, E, stereo_frac):
    """same expressions, same summation trees, same LDL^T pivoting on both sides -> the 21-value state and
    the 15x15 prior Hessian agree to fp64 rounding (1e-9 relative), the classification exactly."""
    import orbx
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    for seed in range(6):
        s = sc.inertial_scenario(1000 * E + seed, E, stereo_frac)
        r = ork.pose_inertial_optimization_last_keyframe(s, cam)
        g = _inertial_call(opt, s, cam)
        assert np.array_equal(g["iters"], r["iters"])
        assert np.array_equal(g["outlier"], r["outlier"]) and g["n"] == r["n"]
        assert np.abs(g["state"] - r["state"]).max() < 1e-9, (seed, np.abs(g["state"] - r["state"]).max())
        assert np.abs(g["H"] - r["H"]).max() <= 1e-9 * np.abs(r["H"]).max()
        # north_star tolerance w.r.t. the oracle, and the job done w.r.t. the truth
        assert np.linalg.norm(g["state"][:9] - r["state"][:9]) / np.sqrt(2) < ROT_TOL_RAD
        assert np.abs(g["state"][9:12] - s["truth"][9:12]).max() < 1e-2


def test_pose_inertial_optimization_edge_cases(ctx, ork):
    import orbx
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    # fewer than 30 inliers: the recovery branch and bRecInit (src/Optimizer.cc:7990-8020)
    s = sc.inertial_scenario(7, 24, 0.5, outlier_frac=0.3)
    for rec in (False, True):
        r = ork.pose_inertial_optimization_last_keyframe(s, cam, rec_init=rec)
        g = _inertial_call(opt, s, cam, rec_init=rec)
        assert np.array_equal(g["outlier"], r["outlier"]) and g["n"] == r["n"] and np.array_equal(g["iters"], r["iters"])
        assert np.abs(g["state"] - r["state"]).max() < 1e-9
    # fewer than 10 graph edges: one round only (:7893)
    s = sc.inertial_scenario(8, 5, 0.5, outlier_frac=0.0)
    r = ork.pose_inertial_optimization_last_keyframe(s, cam)
    g = _inertial_call(opt, s, cam)
    assert list(g["iters"]) == list(r["iters"]) == [10, 0, 0, 0]
    assert np.abs(g["state"] - r["state"]).max() < 1e-9 and np.array_equal(g["outlier"], r["outlier"])
    # no visual edges at all: the inertial and random-walk edges alone
    s = sc.inertial_scenario(9, 0, 0.5)
    r = ork.pose_inertial_optimization_last_keyframe(s, cam)
    g = _inertial_call(opt, s, cam)
    assert np.abs(g["state"] - r["state"]).max() < 1e-9 and g["n"] == r["n"] == 0
    assert np.abs(g["H"] - r["H"]).max() <= 1e-9 * np.abs(r["H"]).max()


def _inertial_lf_call(opt, s, cam, rec_init=False):
    return opt.PoseInertialOptimizationLastFrame(s["xw"], s["obs"], s["isg"], s["close"], cam, s["Tcw"], s["Tcb"], s["Tbc"], s["state"],
                                                 s["prev"], s["preint"], s["preint_jac"], s["preint_bias"], s["infoI"], s["infoG"],
                                                 s["infoA"], s["prior_state"], s["prior_H"], rec_init=rec_init)


@pytest.mark.parametrize("E,stereo_frac", [(300, 0.6), (150, 0.0), (500, 1.0), (60, 0.5), (1000, 0.7)])
def test_pose_inertial_optimization_last_frame_matches_oracle(ctx, ork, E, stereo_frac):
    """SURVEY.md §8 f3, second function (30 unknowns, prior edge, marginalisation): state and marginalised 15x15 prior to fp64
    rounding, classification and iteration counts exactly."""
    import orbx
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    for seed in range(6):
        s = sc.inertial_lf_scenario(2000 * E + seed, E, stereo_frac)
        r = ork.pose_inertial_optimization_last_frame(s, cam)
        g = _inertial_lf_call(opt, s, cam)
        assert np.array_equal(g["iters"], r["iters"])
        assert np.array_equal(g["outlier"], r["outlier"]) and g["n"] == r["n"]
        assert np.abs(g["state"] - r["state"]).max() < 1e-9, (seed, np.abs(g["state"] - r["state"]).max())
        assert np.abs(g["H"] - r["H"]).max() <= 1e-9 * np.abs(r["H"]).max(), (seed, np.abs(g["H"] - r["H"]).max(), np.abs(r["H"]).max())
        assert np.linalg.norm(g["state"][:9] - r["state"][:9]) / np.sqrt(2) < ROT_TOL_RAD
        assert np.abs(g["state"][9:12] - s["truth"][9:12]).max() < 1e-2


def test_pose_inertial_optimization_last_frame_edge_cases(ctx, ork):
    import orbx
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    s = sc.inertial_lf_scenario(7, 24, 0.5, outlier_frac=0.3)           # < 30 inliers: recovery branch / bRecInit
    for rec in (False, True):
        r = ork.pose_inertial_optimization_last_frame(s, cam, rec_init=rec)
        g = _inertial_lf_call(opt, s, cam, rec_init=rec)
        assert np.array_equal(g["outlier"], r["outlier"]) and g["n"] == r["n"] and np.array_equal(g["iters"], r["iters"])
        assert np.abs(g["state"] - r["state"]).max() < 1e-9
    s = sc.inertial_lf_scenario(8, 5, 0.5, outlier_frac=0.0)            # 5 + 4 graph edges < 10: one round only
    r = ork.pose_inertial_optimization_last_frame(s, cam)
    g = _inertial_lf_call(opt, s, cam)
    assert list(g["iters"]) == list(r["iters"]) == [10, 0, 0, 0]
    assert np.abs(g["state"] - r["state"]).max() < 1e-9 and np.array_equal(g["outlier"], r["outlier"])
    s = sc.inertial_lf_scenario(9, 0, 0.5)                              # no visual edges: IMU + prior alone
    r = ork.pose_inertial_optimization_last_frame(s, cam)
    g = _inertial_lf_call(opt, s, cam)
    assert np.abs(g["state"] - r["state"]).max() < 1e-9 and g["n"] == r["n"] == 0
    assert np.abs(g["H"] - r["H"]).max() <= 1e-9 * np.abs(r["H"]).max()
    # a stiff and a rank-deficient prior (pseudo-inverse path of the marginalisation)
    s = sc.inertial_lf_scenario(10, 200, 0.6)
    Hp = s["prior_H"].copy()
    Hp[9:, :] = 0
    Hp[:, 9:] = 0
    s["prior_H"] = Hp
    r = ork.pose_inertial_optimization_last_frame(s, cam)
    g = _inertial_lf_call(opt, s, cam)
    assert np.array_equal(g["iters"], r["iters"]) and np.array_equal(g["outlier"], r["outlier"])
    assert np.abs(g["state"] - r["state"]).max() < 1e-8
    assert np.abs(g["H"] - r["H"]).max() <= 1e-8 * np.abs(r["H"]).max()


def test_pose_inertial_last_frame_batch_equals_single_calls(ctx, ork):
    """Many-stream launch (one CTA per problem): bit-identical to the single calls, ragged sizes including an empty problem."""
    import orbx
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    sizes = [300, 0, 57, 1000, 150, 24, 5, 411] * 5
    probs = [sc.inertial_lf_scenario(9000 + i, E, 0.6) for i, E in enumerate(sizes)]
    got = opt.PoseInertialOptimizationLastFrameBatch(probs, cam)
    for i, (s, g) in enumerate(zip(probs, got)):
        r = _inertial_lf_call(opt, s, cam)
        assert np.array_equal(g["state"], r["state"]) and np.array_equal(g["H"], r["H"]), i
        assert np.array_equal(g["outlier"], r["outlier"]) and g["n"] == r["n"] and np.array_equal(g["iters"], r["iters"])
    o = ork.pose_inertial_optimization_last_frame(probs[3], cam)
    assert np.abs(got[3]["state"] - o["state"]).max() < 1e-9


def test_pose_inertial_last_keyframe_batch_equals_single_calls(ctx):
    import orbx
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    sizes = [300, 0, 57, 1000, 150, 24, 5, 411] * 5
    probs = [sc.inertial_scenario(9500 + i, E, 0.6) for i, E in enumerate(sizes)]
    got = opt.PoseInertialOptimizationLastKeyFrameBatch(probs, cam)
    for i, (s, g) in enumerate(zip(probs, got)):
        r = _inertial_call(opt, s, cam)
        assert np.array_equal(g["state"], r["state"]) and np.array_equal(g["H"], r["H"]), i
        assert np.array_equal(g["outlier"], r["outlier"]) and g["n"] == r["n"] and np.array_equal(g["iters"], r["iters"])


def test_pose_inertial_argument_errors(ctx):
    """Error behaviour of the f3 entry points: invalid offsets / missing arrays are refused with ORBX_EINVAL, not launched."""
    import ctypes as C
    import orbx
    from orbx import api
    L = api.load_library()
    cam = orbx.make_camera()
    s = sc.inertial_lf_scenario(1, 40, 0.5)
    k = orbx.Optimizer.pack_inertial_lf([s, s])
    st = k["state"].copy()
    outl = np.zeros(80, np.uint8); H = np.zeros((2, 225)); n = np.zeros(2, np.int32); it = np.zeros((2, 4), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731

    def call(ofs, xw=k["xw"], prior_H=k["prior_H"]):
        return L.orbx_pose_inertial_optimization_last_frame_batch(
            ctx.h, 2, p(ofs), p(xw) if xw is not None else None, p(k["obs"]), p(k["isg"]), p(k["close"]), C.byref(cam), p(k["Tcw"]),
            p(k["Tcb"]), p(k["Tbc"]), p(st), p(k["prev"]), p(k["preint"]), p(k["preint_jac"]), p(k["preint_bias"]), p(k["infoI"]),
            p(k["infoG"]), p(k["infoA"]), p(k["prior_state"]), p(prior_H) if prior_H is not None else None, 0, p(outl), p(H), p(n), p(it))

    assert call(k["ofs"]) == 0
    assert call(np.array([0, 50, 40], np.int32)) != 0          # decreasing offsets
    assert call(np.array([1, 40, 80], np.int32)) != 0          # does not start at 0
    assert call(k["ofs"], xw=None) != 0                        # edges without coordinates
    assert call(k["ofs"], prior_H=None) != 0                   # missing prior
    with pytest.raises(orbx.OrbxError):
        api._check(call(np.array([0, 50, 40], np.int32)), "orbx_pose_inertial_optimization_last_frame_batch")


def test_local_ba_stop_flag_raised_during_the_run(ctx, ork):
    """Tracking::InterruptBA sets mbAbortBA while LocalBundleAdjustment runs (g2o polls forceStopFlag between iterations,
    sparse_optimizer.cpp:369-376; the reference then skips the second optimize(), src/Optimizer.cc:2201-2290).  The flag
    is raised from another thread while the kernel is running: the call must come back early, consistently (status 0,
    fewer iterations than an undisturbed run, poses written), and never later than the undisturbed run's count."""
    import threading
    import time
    import orbx
    cam = orbx.make_camera()
    opt = orbx.Optimizer(ctx)
    s = sc.lba_scenario(1, K=20, M=3000, n_fixed=3)
    a = (s["kf_T"], s["kf_fixed"], s["mp_xyz"], s["e_kf"], s["e_mp"], s["e_obs"], s["e_inv_sigma2"], cam)
    _, _, _, full_it, full_st = opt.LocalBundleAdjustment(*a)
    assert full_st == 0 and full_it.sum() >= 6
    early = 0
    for delay in (0.0002, 0.0006, 0.0012, 0.002):
        stop = np.zeros(1, np.uint8)

        def raise_flag():
            time.sleep(delay)
            stop[0] = 1
        th = threading.Thread(target=raise_flag)
        th.start()
        gT, gX, gbad, git, gst = opt.LocalBundleAdjustment(*a, stop=stop)
        th.join()
        assert gst in (0, 1)
        assert git[0] <= full_it[0] and git[1] <= full_it[1]
        if git.sum() < full_it.sum():
            early += 1
            assert np.isfinite(gT).all() and np.isfinite(gX).all()
            if gst == 0 and git.sum() > 0:      # interrupted after some iterations: the partial result is written back
                assert not np.array_equal(gT.reshape(-1, 16), s["kf_T"].reshape(-1, 16))
    assert early >= 1, "the stop flag never took effect during a run"
