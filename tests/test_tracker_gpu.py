"""GPU parity of the whole per-frame chain (orbx_tracker_step: extract -> stereo match -> SearchByProjection(last)
-> PoseOptimization -> SearchByProjection(local map) -> PoseOptimization) against the same chain composed from the
CPU oracle's functions."""
import numpy as np
import pytest

import scenarios as sc
from replay_reference import track_frame

pytestmark = pytest.mark.gpu


def _poses(rng, S):
    Tt, Tp = [], []
    for _ in range(S):
        R = sc.rot_small(rng, rng.uniform(0, 10))
        t = rng.uniform(-0.5, 0.5, 3)
        T = sc.se3_matrix(R, t)
        Rp = sc.rot_small(rng, 0.4)
        P = sc.se3_matrix(Rp @ R, Rp @ t + rng.normal(0, 0.01, 3))
        Tt.append(T.astype(np.float32))
        Tp.append(P.astype(np.float32))
    return np.array(Tt), np.array(Tp)


def test_tracker_chain_matches_oracle_chain(ctx, ork):
    import orbx
    from orbx import synth
    S = 3
    cam = orbx.make_camera()
    rng = np.random.default_rng(11)
    imgs = []
    for s in range(S):
        L, R = synth.stereo_pair(40 + s)
        imgs += [L, R]
    Tt, Tp = _poses(rng, S)
    ex = orbx.ORBextractor(ctx, max_batch=2 * S)
    trk = orbx.Tracker(ctx, ex, S, cam)
    for rep in range(2):   # second call: same buffers, same answer
        Tout, stats = trk.step(imgs, Tt, Tp)
        for s in range(S):
            T2, st = track_frame(ork, cam, imgs[2 * s], imgs[2 * s + 1], Tt[s], Tp[s])
            assert np.array_equal(stats[s], st), (s, stats[s], st)
            assert np.abs(Tout[s] - T2).max() < 2e-6, (s, np.abs(Tout[s] - T2).max())
            # the replay converges back to the pose the map was built at
            assert np.abs(Tout[s][:3, 3] - Tt[s][:3, 3]).max() < 5e-3
            assert st[2] > 300 and st[3] > 150 and st[6] > 150
    trk.close()
    ex.close()


def test_tracker_overlap_mode_gives_identical_results(ctx, ork):
    """Two-stream, double-buffered overlap mode is pure scheduling: same poses and statistics, step after step."""
    import orbx
    from orbx import synth
    S = 2
    cam = orbx.make_camera()
    rng = np.random.default_rng(5)
    imgsA, imgsB = [], []
    for s in range(S):
        L, R = synth.stereo_pair(60 + s)
        imgsA += [L, R]
        L, R = synth.stereo_pair(70 + s)
        imgsB += [L, R]
    Tt, Tp = _poses(rng, S)
    ex = orbx.ORBextractor(ctx, max_batch=2 * S)
    trk = orbx.Tracker(ctx, ex, S, cam)
    ref = [trk.step(im, Tt, Tp) for im in (imgsA, imgsB, imgsA)]
    trk.set_overlap(True)
    got = [trk.step(im, Tt, Tp) for im in (imgsA, imgsB, imgsA, imgsB)]
    for k in range(3):
        assert np.array_equal(ref[k][0], got[k][0]) and np.array_equal(ref[k][1], got[k][1]), k
    assert np.array_equal(got[3][0], got[1][0])
    trk.close()
    ex.close()


def test_tracker_submit_collect_matches_step(ctx):
    """The asynchronous host pipeline (copy stream + overlap mode, two steps in flight) returns, step for step, exactly
    what the synchronous orbx_tracker_step returns — from pageable and from page-locked image memory."""
    import orbx
    from orbx import synth
    S = 2
    cam = orbx.make_camera()
    rng = np.random.default_rng(9)
    batches = []
    for k in range(3):
        imgs = []
        for s in range(S):
            L, R = synth.stereo_pair(80 + 10 * k + s)
            imgs += [L, R]
        batches.append(imgs)
    Tt, Tp = _poses(rng, S)
    ex = orbx.ORBextractor(ctx, max_batch=2 * S)
    trk = orbx.Tracker(ctx, ex, S, cam)
    ref = [trk.step(im, Tt, Tp) for im in batches]
    pinned = []
    for im in batches:
        buf = orbx.host_array((2 * S,) + im[0].shape, np.uint8)
        buf[:] = np.stack(im)
        pinned.append([buf[i] for i in range(2 * S)])
    for source in (batches, pinned):
        prepared = [orbx.prepare_images(im) for im in source]
        order = [0, 1, 2, 0, 1]
        got = []
        trk.submit(prepared[order[0]], Tt, Tp)
        for j in range(1, len(order)):
            trk.submit(prepared[order[j]], Tt, Tp)       # two in flight
            got.append(trk.collect())
        got.append(trk.collect())
        for j, k in enumerate(order):
            assert np.array_equal(got[j][0], ref[k][0]) and np.array_equal(got[j][1], ref[k][1]), (j, k)
    with pytest.raises(orbx.OrbxError):
        trk.collect()                                      # nothing outstanding
    trk.submit(prepared[0], Tt, Tp)
    trk.submit(prepared[1], Tt, Tp)
    with pytest.raises(orbx.OrbxError):
        trk.submit(prepared[2], Tt, Tp)                    # a third outstanding step is refused
    trk.collect()
    trk.collect()
    out, st = trk.step(batches[2], Tt, Tp)                 # the synchronous entry point still works afterwards
    assert np.array_equal(out, ref[2][0]) and np.array_equal(st, ref[2][1])
    trk.close()
    ex.close()


def _map_setup(ork, S, seed0):
    """S stereo pairs, their true poses / priors and the §8(d) map of every stream (generated from the oracle's features)."""
    from orbx import synth
    rng = np.random.default_rng(seed0)
    imgs, maps = [], []
    Tt, Tp = _poses(rng, S)
    for s in range(S):
        L, R = synth.stereo_pair(seed0 + s)
        imgs += [L, R]
        exL, kL, dL = sc.extract_frame(ork, L)
        exR, kR, dR = sc.extract_frame(ork, R)
        ur, dp = ork.stereo_match([exL.pyramid_level(l) for l in range(8)], [exR.pyramid_level(l) for l in range(8)], kL, dL, kR, dR,
                                  exL.scale, exL.inv_scale, sc.BF, sc.BF / sc.FX)
        maps.append(sc.track_map_scenario(seed0 + 100 + s, kL, dL, ur, dp, Tt[s]))
    return imgs, Tt, Tp, maps


def test_tracker_given_map_matches_oracle_chain(ctx, ork):
    """§8(d) workload: bit-flipped descriptors, distractors, pixel noise, 20 % gross outliers, full isInFrustum test."""
    import orbx
    from replay_reference import track_frame_map
    S = 2
    cam = orbx.make_camera()
    imgs, Tt, Tp, maps = _map_setup(ork, S, 140)
    ex = orbx.ORBextractor(ctx, max_batch=2 * S)
    trk = orbx.Tracker(ctx, ex, S, cam)
    assert trk.map_capacity == 2048
    host = sc.stack_track_maps(maps)
    for rep in range(2):
        trk.upload_map(host)
        Tout, stats = trk.step(imgs, Tt, Tp)
        for s in range(S):
            T2, st = track_frame_map(ork, cam, imgs[2 * s], imgs[2 * s + 1], maps[s], Tp[s])
            assert np.array_equal(stats[s], st), (s, stats[s], st)
            assert np.abs(Tout[s] - T2).max() < 2e-6, (s, np.abs(Tout[s] - T2).max())
            # the workload is not a best case any more: outliers are found, and the pose still comes back
            assert st[3] - st[4] > 20 and st[5] > 50, st           # PoseOptimization #1 rejected the gross outliers
            assert np.abs(Tout[s][:3, 3] - Tt[s][:3, 3]).max() < 2e-2
    # back to the self-map harness: identical to a tracker that never saw a map
    trk.set_map(None)
    a = trk.step(imgs, Tt, Tp)
    trk2 = orbx.Tracker(ctx, ex, S, cam)
    b = trk2.step(imgs, Tt, Tp)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    trk.close()
    trk2.close()
    ex.close()


def test_monocular_tracker_matches_oracle_chain(ctx, ork):
    """BASELINE config 1: ONE image per stream, no stereo matching, monocular edges only (chi2 5.991), th = 15."""
    import orbx
    from replay_reference import track_frame_map
    S = 3
    cam = orbx.make_camera()
    imgs, Tt, Tp, maps = _map_setup(ork, S, 170)
    left = imgs[0::2]
    ex = orbx.ORBextractor(ctx, max_batch=S)
    trk = orbx.Tracker(ctx, ex, S, cam, mono=True)
    assert trk.ips == 1
    with pytest.raises(orbx.OrbxError):                             # no depth -> no self-map harness
        trk.step(left, Tt, Tp)
    host = sc.stack_track_maps(maps)
    for rep in range(2):
        trk.upload_map(host)
        Tout, stats = trk.step(left, Tt, Tp)
        for s in range(S):
            T2, st = track_frame_map(ork, cam, left[s], None, maps[s], Tp[s], mono=True)
            assert st[1] == 0 and st[2] == 0
            assert np.array_equal(stats[s], st), (s, stats[s], st)
            assert np.abs(Tout[s] - T2).max() < 2e-6, (s, np.abs(Tout[s] - T2).max())
            assert st[6] > 100 and np.abs(Tout[s][:3, 3] - Tt[s][:3, 3]).max() < 3e-2, st
    # the asynchronous host path carries one image per stream as well
    trk.upload_map(host)
    trk.submit(left, Tt, Tp)
    got = trk.collect()
    assert np.array_equal(got[0], Tout) and np.array_equal(got[1], stats)
    trk.close()
    ex.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_inertial_tracker_matches_oracle_chain(ctx, ork, mode):
    """BASELINE config 3: the second pose optimisation of the step is PoseInertialOptimizationLastKeyFrame (mode 1) /
    LastFrame (mode 2), fed on the device from the pose the first PoseOptimization left (src/Tracking.cc:2466-2490)."""
    import orbx
    from replay_reference import track_frame_map
    S = 2
    cam = orbx.make_camera()
    imgs, Tt, Tp, maps = _map_setup(ork, S, 180)
    imus = [sc.track_imu_scenario(900 + s, Tt[s], mode) for s in range(S)]
    ex = orbx.ORBextractor(ctx, max_batch=2 * S)
    trk = orbx.Tracker(ctx, ex, S, cam)
    host = sc.stack_track_maps(maps)
    himu = sc.stack_track_imu(imus)
    want = [track_frame_map(ork, cam, imgs[2 * s], imgs[2 * s + 1], maps[s], Tp[s], imu=imus[s], imu_mode=mode, want_inertial=True)
            for s in range(S)]
    for rep in range(2):
        trk.upload_map(host)
        trk.upload_inertial(mode, himu)
        Tout, stats = trk.step(imgs, Tt, Tp)
        state, H = trk.inertial_result()
        for s in range(S):
            T2, st, res = want[s]
            assert np.array_equal(stats[s], st), (s, stats[s], st)
            assert np.abs(state[s] - res["state"]).max() < 1e-9, (s, np.abs(state[s] - res["state"]).max())
            assert np.abs(H[s] - res["H"]).max() <= 1e-9 * np.abs(res["H"]).max()
            assert np.abs(Tout[s] - T2).max() < 2e-6, (s, np.abs(Tout[s] - T2).max())
            assert st[6] > 500 and np.abs(state[s][9:12] - imus[s]["truth"][9:12]).max() < 2e-2
    # the asynchronous host path carries the inertial inputs too
    trk.upload_map(host)
    trk.upload_inertial(mode, himu)
    trk.submit(imgs, Tt, Tp)
    got = trk.collect()
    assert np.array_equal(got[0], Tout) and np.array_equal(got[1], stats)
    # mode 0 detaches: the visual chain again
    trk.set_inertial(0)
    trk.upload_map(host)
    Tv, sv = trk.step(imgs, Tt, Tp)
    for s in range(S):
        T2, st = track_frame_map(ork, cam, imgs[2 * s], imgs[2 * s + 1], maps[s], Tp[s])
        assert np.array_equal(sv[s], st) and np.abs(Tv[s] - T2).max() < 2e-6
    trk.close()
    ex.close()


def test_inertial_tracker_chains_the_prior_on_the_device(ctx, ork):
    """Mode 2 without ref_state / prior: step t+1 reads the state and marginalised Hessian step t left on the device."""
    import orbx
    from replay_reference import track_frame_map
    S = 2
    cam = orbx.make_camera()
    imgs, Tt, Tp, maps = _map_setup(ork, S, 190)
    ex = orbx.ORBextractor(ctx, max_batch=2 * S)
    trk = orbx.Tracker(ctx, ex, S, cam)
    host = sc.stack_track_maps(maps)
    # step 1: LastKeyFrame seeds state + H15
    imu1 = [sc.track_imu_scenario(950 + s, Tt[s], 1) for s in range(S)]
    trk.upload_map(host)
    trk.upload_inertial(1, sc.stack_track_imu(imu1))
    trk.step(imgs, Tt, Tp)
    st1, H1 = trk.inertial_result()
    # step 2: LastFrame on the same images, previous frame = step 1's result (a stationary rig: zero relative motion)
    imu2 = [sc.track_imu_scenario(970 + s, Tt[s], 2) for s in range(S)]
    h2 = sc.stack_track_imu(imu2)
    chained = {k: v for k, v in h2.items() if k not in ("ref_state", "prior_state", "prior_H")}
    trk.upload_map(host)
    trk.upload_inertial(2, chained)
    Tout, stats = trk.step(imgs, Tt, Tp)
    st2, H2 = trk.inertial_result()
    for s in range(S):
        imu = dict(imu2[s], ref_state=st1[s], prior_state=st1[s], prior_H=H1[s].ravel())
        T2, st, res = track_frame_map(ork, cam, imgs[2 * s], imgs[2 * s + 1], maps[s], Tp[s], imu=imu, imu_mode=2, want_inertial=True)
        assert np.array_equal(stats[s], st), (s, stats[s], st)
        assert np.abs(st2[s] - res["state"]).max() < 1e-9 and np.abs(H2[s] - res["H"]).max() <= 1e-9 * np.abs(res["H"]).max()
        assert np.abs(Tout[s] - T2).max() < 2e-6
    trk.close()
    ex.close()


def test_tracker_graph_replay_equals_eager(ctx, ork):
    """orbx_tracker_set_graph: the step captured as one CUDA graph gives the results of the eager launches, follows new
    image CONTENT (same buffers) and falls back to eager launches when an argument changes."""
    import orbx
    S = 2
    cam = orbx.make_camera()
    imgs, Tt, Tp, maps = _map_setup(ork, S, 220)
    imgs2, _, _, _ = _map_setup(ork, S, 230)
    host = sc.stack_track_maps(maps)
    ex = orbx.ORBextractor(ctx, max_batch=2 * S)
    ref = orbx.Tracker(ctx, ex, S, cam)
    want = []
    for im in (imgs, imgs2, imgs):
        ref.upload_map(host)
        want.append(ref.step(im, Tt, Tp))
    ref.close()
    trk = orbx.Tracker(ctx, ex, S, cam)
    trk.set_graph(True)
    got = []
    for im in (imgs, imgs, imgs2, imgs):               # eager (binds), capture + launch, replay, replay
        trk.upload_map(host)
        got.append(trk.step(im, Tt, Tp))
    assert trk.graph_launches == 3
    for g, w in zip(got, (want[0], want[0], want[1], want[2])):
        assert np.array_equal(g[0], w[0]) and np.array_equal(g[1], w[1])
    assert got[0][1][0][6] > 100
    # a different prior array of the same content is the same staging buffer: still a replay; a different image size is not
    trk.upload_map(host)
    a = trk.step(imgs, Tt, Tp.copy())
    assert trk.graph_launches == 4 and np.array_equal(a[0], want[0][0])
    small = [np.ascontiguousarray(im[:400, :640]) for im in imgs]
    trk.upload_map(host)
    trk.step(small, Tt, Tp)
    assert trk.graph_launches == 4
    trk.set_graph(False)
    trk.upload_map(host)
    b = trk.step(imgs, Tt, Tp)
    assert np.array_equal(b[0], want[0][0]) and np.array_equal(b[1], want[0][1]) and trk.graph_launches == 4
    trk.close()
    ex.close()


def test_tracker_blank_images_in_every_mode(ctx, ork):
    """A stream whose images hold no corner beside a normal one: no keypoints, no matches, PoseOptimization returns the
    prior untouched (nInitialCorrespondences < 3, src/Optimizer.cc:1134); the inertial optimiser runs on the IMU edge alone."""
    import orbx
    from replay_reference import track_frame_map
    S = 2
    cam = orbx.make_camera()
    imgs, Tt, Tp, maps = _map_setup(ork, S, 210)
    blank = np.full_like(imgs[0], 127)
    imgs = [blank, blank, imgs[2], imgs[3]]
    host = sc.stack_track_maps(maps)
    ex = orbx.ORBextractor(ctx, max_batch=2 * S)
    # stereo
    trk = orbx.Tracker(ctx, ex, S, cam)
    trk.upload_map(host)
    Tout, stats = trk.step(imgs, Tt, Tp)
    for s in range(S):
        T2, st = track_frame_map(ork, cam, imgs[2 * s], imgs[2 * s + 1], maps[s], Tp[s])
        assert np.array_equal(stats[s], st) and np.abs(Tout[s] - T2).max() < 2e-6
    assert not stats[0].any() and np.array_equal(Tout[0], Tp[0]) and stats[1][6] > 100
    # stereo-inertial, both optimisers
    for mode in (1, 2):
        imus = [sc.track_imu_scenario(990 + s, Tt[s], mode) for s in range(S)]
        trk.upload_map(host)
        trk.upload_inertial(mode, sc.stack_track_imu(imus))
        Tout, stats = trk.step(imgs, Tt, Tp)
        state, H = trk.inertial_result()
        for s in range(S):
            T2, st, res = track_frame_map(ork, cam, imgs[2 * s], imgs[2 * s + 1], maps[s], Tp[s], imu=imus[s], imu_mode=mode,
                                          want_inertial=True)
            assert np.array_equal(stats[s], st), (mode, s, stats[s], st)
            assert np.abs(state[s] - res["state"]).max() < 1e-9 and np.abs(Tout[s] - T2).max() < 2e-6
            assert np.abs(H[s] - res["H"]).max() <= 1e-9 * max(np.abs(res["H"]).max(), 1.0)
        assert stats[0][0] == 0 and stats[0][6] == 0
    trk.close()
    ex.close()
    # monocular
    ex1 = orbx.ORBextractor(ctx, max_batch=S)
    trk = orbx.Tracker(ctx, ex1, S, cam, mono=True)
    left = imgs[0::2]
    trk.upload_map(host)
    Tout, stats = trk.step(left, Tt, Tp)
    for s in range(S):
        T2, st = track_frame_map(ork, cam, left[s], None, maps[s], Tp[s], mono=True)
        assert np.array_equal(stats[s], st) and np.abs(Tout[s] - T2).max() < 2e-6
    assert not stats[0].any() and np.array_equal(Tout[0], Tp[0])
    trk.close()
    ex1.close()


def test_tracker_chain_mode_composes_the_prior_on_the_device(ctx, ork):
    """Motion-model chaining: step t+1 starts from dT * (pose step t produced).  Equal to feeding that product from the host."""
    import orbx
    S = 2
    cam = orbx.make_camera()
    imgs, Tt, Tp, maps = _map_setup(ork, S, 150)
    host = sc.stack_track_maps(maps)
    ex = orbx.ORBextractor(ctx, max_batch=2 * S)
    trk = orbx.Tracker(ctx, ex, S, cam)
    trk.upload_map(host)
    T1, st1 = trk.step(imgs, Tt, Tp)                               # step 1 (absolute prior)
    rng = np.random.default_rng(3)
    dT = np.stack([sc.se3_matrix(sc.rot_small(rng, 0.2), rng.normal(0, 0.004, 3)).astype(np.float32) for _ in range(S)])
    # host-side product in the kernel's order: ((a0*b0 + a1*b1) + a2*b2) + a3*b3, fp32
    prior = np.zeros((S, 4, 4), np.float32)
    for s in range(S):
        for i in range(4):
            for j in range(4):
                acc = np.float32(dT[s, i, 0] * T1[s, 0, j])
                for k in range(1, 4):
                    acc = np.float32(acc + np.float32(dT[s, i, k] * T1[s, k, j]))
                prior[s, i, j] = acc
    trk.upload_map(host)
    want = trk.step(imgs, Tt, prior)                               # step 2 with the prior composed on the host
    init = np.ascontiguousarray(T1.reshape(S, 16))
    trk.set_chain(True, init.ctypes.data)
    trk.upload_map(host)
    got = trk.step(imgs, Tt, dT)                                   # the same, composed on the device
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    trk.upload_map(host)
    third = trk.step(imgs, Tt, np.stack([np.eye(4, dtype=np.float32)] * S))   # dT = I: starts from step 2's output
    assert np.abs(third[0] - got[0]).max() < 1e-3 and third[1][0][6] > 100
    trk.set_chain(False)
    trk.close()
    ex.close()


def test_tracker_keyframe_work_runs_beside_the_frame_chain(ctx, ork):
    """The keyframe-rate plans attached to a tracker run every `period`-th step on their own stream: same per-frame
    results as without them, and the plans' results are those of a stand-alone run."""
    import orbx
    S = 2
    cam = orbx.make_camera()
    imgs, Tt, Tp, maps = _map_setup(ork, S, 160)
    ex = orbx.ORBextractor(ctx, max_batch=2 * S)
    trk = orbx.Tracker(ctx, ex, S, cam)
    ref = trk.step(imgs, Tt, Tp)
    scen = [sc.lba_scenario(70 + i, K=6, M=300, n_fixed=2) for i in range(S)]
    lba = orbx.LocalBABatch(ctx, scen, cam)
    lba.run()
    alone = lba.fetch()
    trk.set_overlap(True)
    trk.set_keyframe_work(None, lba, 2)
    outs = [trk.step(imgs, Tt, Tp) for _ in range(5)]
    trk.synchronize()
    assert trk.keyframe_runs == 3          # the tracker had done 1 step before: steps 2, 4 and 6 carry keyframe work
    for o in outs:
        assert np.array_equal(o[0], ref[0]) and np.array_equal(o[1], ref[1])
    beside = lba.fetch()
    for q in range(S):
        assert np.array_equal(beside[q][0], alone[q][0]) and np.array_equal(beside[q][3], alone[q][3])
    trk.set_keyframe_work(None, None, 0)
    trk.close()
    lba.close()
    ex.close()
