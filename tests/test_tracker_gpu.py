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
