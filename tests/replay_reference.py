"""The tracking-replay chain of orbx_tracker_step, composed on the CPU from the oracle's functions plus numpy
float32 glue that mirrors the harness kernels' expression order (back-projection, frustum projection).
Used by the GPU parity test of the whole chain and by bench.py's CPU baseline."""
import numpy as np

import scenarios as sc

F32 = np.float32


def track_frame(ork, cam, L, R, Tcw_true, Tcw_prior, th_frame=7.0, th_map=1.0, nn_map=0.8, nfeatures=1000,
                extractors=None):
    from orbx import abi
    exL, exR = extractors if extractors else (ork.Extractor(nfeatures), ork.Extractor(nfeatures))
    _, kL, dL, _ = exL(L)
    _, kR, dR, _ = exR(R)
    scale, inv_scale = exL.scale, exL.inv_scale
    isg = exL.inv_sigma2
    pyrL = [exL.pyramid_level(l) for l in range(exL.nlevels)]
    pyrR = [exR.pyramid_level(l) for l in range(exR.nlevels)]
    bf, b = F32(cam.bf), F32(cam.b)
    ur, dp = ork.stereo_match(pyrL, pyrR, kL, dL, kR, dR, scale, inv_scale, float(bf), float(b))
    n = len(kL)
    fx, fy, cx, cy = F32(cam.fx), F32(cam.fy), F32(cam.cx), F32(cam.cy)
    # --- back-projection at the true pose (backproject_kernel) ---
    T = np.asarray(Tcw_true, F32).reshape(4, 4)
    has = dp > 0
    z = dp.astype(F32)
    xc = (kL["x"] - cx) * z / fx
    yc = (kL["y"] - cy) * z / fy
    dx, dy, dz = xc - T[0, 3], yc - T[1, 3], z - T[2, 3]
    xw = np.stack([T[0, 0] * dx + T[1, 0] * dy + T[2, 0] * dz, T[0, 1] * dx + T[1, 1] * dy + T[2, 1] * dz,
                   T[0, 2] * dx + T[1, 2] * dy + T[2, 2] * dz], 1).astype(F32)
    flags = np.where(has, 3, 0).astype(np.uint8)
    in_last = (((np.arange(n, dtype=np.uint64) * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)) >> np.uint64(16)) % np.uint64(5) < 3
    last_flags = np.where(has & in_last, 3, 0).astype(np.uint8)   # the last frame tracked ~60 % of the local map
    H, W = L.shape
    Fr = abi.Frame(kL, dL, ur, bounds=(0, 0, W, H))
    # --- SearchByProjection(Cur, Last) ---
    Tp = np.asarray(Tcw_prior, F32).reshape(4, 4)
    nm1, match, kept, cur = ork.search_by_projection_frame(Fr, None, cam, Tp, np.eye(4, dtype=F32), last_flags, xw,
                                                           kL["octave"].astype(np.int32), kL["angle"], dL, th_frame,
                                                           False, True, scale)
    # (the device harness fixes the octave-window mode to "neither forward nor backward"; with Tcw_last = I and a
    #  prior close to the true pose |tlc.z| < mb holds as long as the test keeps the translation small)

    def edges(assign):
        idx = np.flatnonzero(assign >= 0)
        q = assign[idx]
        obs = np.stack([kL["x"][idx], kL["y"][idx], ur[idx]], 1).astype(F32)
        return idx, xw[q], obs, isg[kL["octave"][idx]].astype(F32)

    idx1, exw, eobs, eisg = edges(cur)
    T1, out1, nin1, it1 = ork.pose_optimization(exw, eobs, eisg, cam, Tp)
    # --- drop outliers, SearchLocalPoints + SearchByProjection(F, local map) ---
    cur2 = cur.copy()
    cur2[idx1[out1 == 1]] = -1
    blocked = (cur2 >= 0).astype(np.uint8)
    taken = np.zeros(n, bool)
    taken[cur2[cur2 >= 0]] = True
    X, Y, Z = xw[:, 0], xw[:, 1], xw[:, 2]
    xcm = T1[0, 0] * X + T1[0, 1] * Y + T1[0, 2] * Z + T1[0, 3]
    ycm = T1[1, 0] * X + T1[1, 1] * Y + T1[1, 2] * Z + T1[1, 3]
    zcm = T1[2, 0] * X + T1[2, 1] * Y + T1[2, 2] * Z + T1[2, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        invz = F32(1.0) / zcm
        u = fx * xcm / zcm + cx
        v = fy * ycm / zcm + cy
    vis = has & ~taken & (zcm > 0) & (u >= 0) & (u <= W) & (v >= 0) & (v <= H)
    mflags = np.where(vis, 3, 0).astype(np.uint8)
    projXR = (u - bf * invz).astype(F32)
    nm2, best = ork.search_by_projection_map(Fr, blocked, np.where(vis, u, 0).astype(F32), np.where(vis, v, 0).astype(F32),
                                             np.where(vis, projXR, 0).astype(F32), kL["octave"].astype(np.int32),
                                             np.ones(n, F32), dL, mflags, th_map, nn_map, scale)
    kpmp = cur2.copy()
    for q in range(n):
        if best[q] >= 0:
            kpmp[best[q]] = q
    idx2, exw2, eobs2, eisg2 = edges(kpmp)
    T2, out2, nin2, it2 = ork.pose_optimization(exw2, eobs2, eisg2, cam, T1)
    stats = np.array([n, len(kR), int((ur >= 0).sum()), nm1, nin1, nm2, nin2, int(it1.sum() + it2.sum())], np.int32)
    return T2, stats


def _f32_dot3(a0, b0, a1, b1, a2, b2):
    """((a0*b0 + a1*b1) + a2*b2) in float32, the order of cv::gemm's small-matrix path"""
    return F32(F32(F32(a0) * F32(b0)) + F32(F32(a1) * F32(b1))) + F32(F32(a2) * F32(b2))


def camera_center(T):
    """Frame::UpdatePoseMatrices: mOw = -mRcw.t()*mtcw (src/Frame.cc:543).  cv::MatExpr hands the transposed operand to
    cv::gemm as GEMM_1_T, which takes the general path: products and sum in double, one rounding to float."""
    T = np.asarray(T, F32).reshape(4, 4).astype(np.float64)
    return np.array([-(T[0, i] * T[0, 3] + T[1, i] * T[1, 3] + T[2, i] * T[2, 3]) for i in range(3)]).astype(F32)


def imu_state_from_pose(T1, Tcb, velocity, bias):
    """ImuCamPose(Frame*) inputs: Frame::GetImuRotation / GetImuPosition on the pose T1 (src/Frame.cc:534-554), float32"""
    T1 = np.asarray(T1, F32).reshape(4, 4)
    Tcb = np.asarray(Tcb, F32).reshape(4, 4)
    Rwc = T1[:3, :3].T
    tcw = T1[:3, 3]
    Ow = camera_center(T1)
    Rwb = np.array([[_f32_dot3(Rwc[i, 0], Tcb[0, j], Rwc[i, 1], Tcb[1, j], Rwc[i, 2], Tcb[2, j]) for j in range(3)] for i in range(3)], F32)
    twb = np.array([F32(_f32_dot3(Rwc[i, 0], Tcb[0, 3], Rwc[i, 1], Tcb[1, 3], Rwc[i, 2], Tcb[2, 3]) + Ow[i]) for i in range(3)], F32)
    return np.concatenate([Rwb.astype(np.float64).ravel(), twb.astype(np.float64), np.asarray(velocity, F32).astype(np.float64),
                           np.asarray(bias, F32).astype(np.float64)])


def pose_from_imu_state(state, Tcb):
    """Frame::SetImuPoseVelocity (src/Frame.cc:520-530): Tcw = Tcb * [Rwb^T | -Rwb^T twb], float32"""
    Tcb = np.asarray(Tcb, F32).reshape(4, 4)
    Rbw = np.asarray(state[:9], np.float64).reshape(3, 3).astype(F32).T
    twb = np.asarray(state[9:12], np.float64).astype(F32)
    Tbw = np.eye(4, dtype=F32)
    Tbw[:3, :3] = Rbw
    for i in range(3):
        Tbw[i, 3] = -_f32_dot3(Rbw[i, 0], twb[0], Rbw[i, 1], twb[1], Rbw[i, 2], twb[2])
    T = np.zeros((4, 4), F32)
    for i in range(4):
        for j in range(4):
            T[i, j] = F32(_f32_dot3(Tcb[i, 0], Tbw[0, j], Tcb[i, 1], Tbw[1, j], Tcb[i, 2], Tbw[2, j]) + F32(Tcb[i, 3] * Tbw[3, j]))
    return T


def track_frame_map(ork, cam, L, R, mp, Tcw_prior, th_frame=None, th_map=1.0, nn_map=0.8, nfeatures=1000, extractors=None,
                    log_sf=None, mono=False, imu=None, imu_mode=0, want_inertial=False):
    """The chain of orbx_tracker_step with a GIVEN map (orbx_tracker_set_map; mp = one stream's arrays as produced by
    scenarios.track_map_scenario), composed from the oracle's functions: extract L+R, ComputeStereoMatches,
    SearchByProjection(Cur, Last) over the last-frame entries, PoseOptimization, outliers dropped, isInFrustum over the
    unmatched local map, SearchByProjection(F, local map), PoseOptimization."""
    from orbx import abi
    if th_frame is None:
        th_frame = 15.0 if mono else 7.0                  # src/Tracking.cc:2364-2368
    exL, exR = extractors if extractors else (ork.Extractor(nfeatures), ork.Extractor(nfeatures))
    _, kL, dL, _ = exL(L)
    scale, inv_scale, isg = exL.scale, exL.inv_scale, exL.inv_sigma2
    if mono:
        # Frame::Frame(mono): one image, mvuRight = mvDepth = -1 (src/Frame.cc:330-331); R is ignored
        kR = ()
        ur = np.full(len(kL), -1, F32)
    else:
        _, kR, dR, _ = exR(R)
        pyrL = [exL.pyramid_level(l) for l in range(exL.nlevels)]
        pyrR = [exR.pyramid_level(l) for l in range(exR.nlevels)]
        ur, dp = ork.stereo_match(pyrL, pyrR, kL, dL, kR, dR, scale, inv_scale, float(F32(cam.bf)), float(F32(cam.b)))
    n, M = len(kL), int(mp["n_map"])
    H, W = L.shape
    Fr = abi.Frame(kL, dL, ur, bounds=(0, 0, W, H))
    xw = mp["xw"][:M]
    Tp = np.asarray(Tcw_prior, F32).reshape(4, 4)
    # (the device harness fixes the octave-window mode to "neither forward nor backward", i.e. |tlc.z| <= mb: the last
    #  frame's pose is taken equal to the prior here, so tlc = 0)
    nm1, match, kept, cur = ork.search_by_projection_frame(Fr, None, cam, Tp, Tp, mp["last_flags"][:M], xw,
                                                           mp["last_octave"][:M], mp["last_angle"][:M], mp["desc"][:M], th_frame,
                                                           mono, True, scale)

    def edges(assign):
        idx = np.flatnonzero(assign >= 0)
        q = assign[idx]
        obs = np.stack([kL["x"][idx], kL["y"][idx], ur[idx]], 1).astype(F32)
        return idx, xw[q], obs, isg[kL["octave"][idx]].astype(F32)

    idx1, exw, eobs, eisg = edges(cur)
    T1, out1, nin1, it1 = ork.pose_optimization(exw, eobs, eisg, cam, Tp)
    cur2 = cur.copy()
    cur2[idx1[out1 == 1]] = -1
    blocked = (cur2 >= 0).astype(np.uint8)
    taken = np.zeros(M, bool)
    taken[cur2[cur2 >= 0]] = True
    Ow = camera_center(T1)
    if log_sf is None:
        log_sf = float(F32(np.log(np.float64(scale[1]))))
    fr = ork.is_in_frustum(cam, T1[:3, :3], T1[:3, 3], Ow, (0.0, float(W), 0.0, float(H)), 0.5, exL.nlevels, log_sf, xw,
                           mp["max_dist"][:M], mp["min_dist"][:M], mp["normal"][:M])
    vis = (fr["in_view"][:M] > 0) & ((mp["map_flags"][:M] & 1) > 0) & ~taken
    mflags = np.where(vis, 1 | (mp["map_flags"][:M] & 2), 0).astype(np.uint8)
    nm2, best = ork.search_by_projection_map(Fr, blocked, fr["proj_x"][:M], fr["proj_y"][:M], fr["proj_xr"][:M], fr["level"][:M],
                                             fr["view_cos"][:M], mp["desc"][:M], mflags, th_map, nn_map, scale)
    kpmp = cur2.copy()
    for q in range(M):
        if best[q] >= 0:
            kpmp[best[q]] = q
    idx2, exw2, eobs2, eisg2 = edges(kpmp)
    inertial = None
    if imu_mode:
        # visual-inertial TrackLocalMap (src/Tracking.cc:2466-2490) on the same edges
        close = ((mp["map_flags"][:M][kpmp[idx2]] >> 2) & 1).astype(np.uint8)
        sc_in = dict(xw=exw2, obs=eobs2, isg=eisg2, close=close, Tcw=T1, Tcb=imu["Tcb"], Tbc=imu["Tbc"],
                     state=imu_state_from_pose(T1, imu["Tcb"], imu["velocity"], imu["bias"]), preint=imu["preint"],
                     infoI=imu["info_inertial"], infoG=imu["info_gyro"], infoA=imu["info_acc"])
        if imu_mode == 1:
            sc_in["kf"] = imu["ref_state"]
            res = ork.pose_inertial_optimization_last_keyframe(sc_in, cam)
        else:
            sc_in.update(prev=imu["ref_state"], preint_jac=imu["preint_jac"], preint_bias=imu["preint_bias"],
                         prior_state=imu["prior_state"], prior_H=imu["prior_H"])
            res = ork.pose_inertial_optimization_last_frame(sc_in, cam)
        T2, nin2, it2 = pose_from_imu_state(res["state"], imu["Tcb"]), res["n"], res["iters"]
        inertial = res
    else:
        T2, out2, nin2, it2 = ork.pose_optimization(exw2, eobs2, eisg2, cam, T1)
    stats = np.array([n, len(kR), int((ur >= 0).sum()), nm1, nin1, nm2, nin2, int(it1.sum() + it2.sum())], np.int32)
    if want_inertial:
        return T2, stats, inertial
    return T2, stats
