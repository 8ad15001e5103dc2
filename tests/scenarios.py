"""Seeded synthetic matcher / optimiser scenarios, shared by CPU and GPU tests.
Real descriptors come from the extractor run on synthetic images and from tests/golden/orbvoc_sample.npy
(4096 genuine ORB descriptors sampled from the reference's Vocabulary/ORBvoc.bin)."""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FX, FY, CX, CY, BF = 458.654, 457.296, 367.215, 248.375, 47.9   # Examples/Monocular/EuRoC.yaml:9-12


def orbvoc():
    return np.load(os.path.join(HERE, "golden", "orbvoc_sample.npy"))


def flip_bits(rng, desc, nflip):
    d = np.unpackbits(desc.copy())
    pos = rng.choice(256, size=nflip, replace=False)
    d[pos] ^= 1
    return np.packbits(d)


def extract_frame(ork, img, nfeatures=1000, lap=(0, 0)):
    ex = ork.Extractor(nfeatures)
    rc, k, d, m = ex(img, lap)
    assert rc == 0
    return ex, k, d


def sbp_map_scenario(seed, kps, desc, uright=None, nq=1500, th=1.0):
    rng = np.random.default_rng(seed)
    voc = orbvoc()
    n = len(kps)
    scale = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    projX, projY, projXR = np.zeros(nq, np.float32), np.zeros(nq, np.float32), np.zeros(nq, np.float32)
    level, viewCos = np.zeros(nq, np.int32), np.zeros(nq, np.float32)
    mpDesc, flags = np.zeros((nq, 32), np.uint8), np.zeros(nq, np.uint8)
    for q in range(nq):
        if rng.random() < 0.7 and n > 0:
            i = int(rng.integers(n))
            mpDesc[q] = flip_bits(rng, desc[i], int(rng.integers(0, 41)))
            projX[q] = kps["x"][i] + rng.normal(0, 2.0)
            projY[q] = kps["y"][i] + rng.normal(0, 2.0)
            level[q] = int(np.clip(kps["octave"][i] + rng.choice([0, 0, 0, 1]), 0, 7))
            if uright is not None and uright[i] > 0:
                projXR[q] = uright[i] + rng.normal(0, 1.5)
            else:
                projXR[q] = projX[q] - rng.uniform(3, 40)
        else:
            mpDesc[q] = voc[int(rng.integers(len(voc)))]
            projX[q], projY[q] = rng.uniform(0, 752), rng.uniform(0, 480)
            level[q] = int(rng.integers(0, 8))
            projXR[q] = projX[q] - rng.uniform(3, 40)
        viewCos[q] = rng.choice([0.9995, 0.9985, 0.95, 0.998])
        flags[q] = (1 if rng.random() < 0.9 else 0) | (2 if rng.random() < 0.85 else 0)
    blocked = (rng.random(n) < 0.05).astype(np.uint8)
    return dict(kp_blocked=blocked, projX=projX, projY=projY, projXR=projXR, level=level, viewCos=viewCos,
                mpDesc=mpDesc, flags=flags, th=th, scaleFactors=scale)


def rot_small(rng, deg):
    w = rng.normal(0, 1, 3)
    w = w / np.linalg.norm(w) * np.deg2rad(deg)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def sbp_frame_scenario(seed, kps, desc, uright=None, depth=None, rot_deg=0.6, trans=0.03, forward=0.0):
    """Last frame == the given keypoints observed at identity pose; current pose slightly perturbed."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    z = np.where((depth is not None) & (np.asarray(depth if depth is not None else np.zeros(n)) > 0),
                 depth if depth is not None else 0, rng.uniform(1.5, 12.0, n)).astype(np.float64)
    X = np.stack([(kps["x"] - CX) / FX * z, (kps["y"] - CY) / FY * z, z], 1)
    Tl = np.eye(4, dtype=np.float32)
    R = rot_small(rng, rot_deg)
    t = rng.normal(0, trans, 3)
    t[2] += forward
    Tc = np.eye(4)
    Tc[:3, :3], Tc[:3, 3] = R, t
    flags = ((rng.random(n) < 0.7).astype(np.uint8)) | ((rng.random(n) < 0.9).astype(np.uint8) << 1)
    voc = orbvoc()
    mpDesc = np.stack([flip_bits(rng, desc[i], int(rng.integers(0, 31))) if rng.random() < 0.9
                       else voc[int(rng.integers(len(voc)))] for i in range(n)]) if n else np.zeros((0, 32), np.uint8)
    angle = (kps["angle"] + np.where(rng.random(n) < 0.85, rng.normal(3.0, 2.0, n), rng.uniform(0, 360, n))) % 360.0
    blocked = (rng.random(n) < 0.03).astype(np.uint8)
    return dict(cur_blocked=blocked, Tcw_cur=Tc.astype(np.float32), Tcw_last=Tl, flags=flags,
                xw=X.astype(np.float32), octave=kps["octave"].astype(np.int32), angle=angle.astype(np.float32),
                mpDesc=mpDesc, scaleFactors=(np.float32(1.2) ** np.arange(8)).astype(np.float32))


def feature_vector(desc, words, id_stride=7):
    """Stand-in for DBoW2::FeatureVector: bucket = nearest of `words` (Hamming); CSR (node ids, offsets, idx)."""
    if len(desc) == 0:
        return np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32)
    d = np.unpackbits(desc, axis=1).astype(np.int16)
    w = np.unpackbits(words, axis=1).astype(np.int16)
    ham = (d[:, None, :] != w[None, :, :]).sum(2)
    node = ham.argmin(1)
    ids = np.unique(node)
    off, idx = [0], []
    for nid in ids:
        members = np.flatnonzero(node == nid)
        idx.extend(members.tolist())
        off.append(len(idx))
    return (ids * id_stride + 3).astype(np.int32), np.array(off, np.int32), np.array(idx, np.int32)


def tri_scenario(seed, kps, desc, uright=None, baseline=0.25, rot_deg=2.0, nwords=48):
    """KF1 = given keypoints at identity; KF2 = their reprojection under a known relative pose (+noise,
    +distractors, shuffled)."""
    rng = np.random.default_rng(seed)
    voc = orbvoc()
    n = len(kps)
    z = rng.uniform(2.0, 12.0, n)
    X = np.stack([(kps["x"] - CX) / FX * z, (kps["y"] - CY) / FY * z, z], 1)
    R2 = rot_small(rng, rot_deg)
    t2 = np.array([-baseline, rng.normal(0, 0.02), rng.normal(0, 0.05)])
    Xc2 = X @ R2.T + t2
    u2 = FX * Xc2[:, 0] / Xc2[:, 2] + CX + rng.normal(0, 0.6, n)
    v2 = FY * Xc2[:, 1] / Xc2[:, 2] + CY + rng.normal(0, 0.6, n)
    ok = (u2 > 20) & (u2 < 732) & (v2 > 20) & (v2 < 460) & (rng.random(n) < 0.85)
    k2 = kps[ok].copy()
    k2["x"], k2["y"] = u2[ok].astype(np.float32), v2[ok].astype(np.float32)
    k2["angle"] = ((k2["angle"] + np.where(rng.random(len(k2)) < 0.9, rng.normal(-4, 2, len(k2)),
                                            rng.uniform(0, 360, len(k2)))) % 360).astype(np.float32)
    d2 = np.stack([flip_bits(rng, d, int(rng.integers(0, 36))) for d in desc[ok]]) if ok.any() else np.zeros((0, 32), np.uint8)
    nd = 150
    kd = np.zeros(nd, kps.dtype)
    kd["x"], kd["y"] = rng.uniform(20, 732, nd), rng.uniform(20, 460, nd)
    kd["octave"], kd["angle"] = rng.integers(0, 8, nd), rng.uniform(0, 360, nd)
    dd = voc[rng.integers(0, len(voc), nd)]
    k2 = np.concatenate([k2, kd])
    d2 = np.concatenate([d2, dd])
    perm = rng.permutation(len(k2))
    k2, d2 = k2[perm], np.ascontiguousarray(d2[perm])
    words = voc[rng.choice(len(voc), nwords, replace=False)]
    ur1 = uright
    ur2 = np.where(rng.random(len(k2)) < 0.3, k2["x"] - rng.uniform(2, 30, len(k2)), -1).astype(np.float32) \
        if uright is not None else None
    return dict(k1=kps, d1=desc, ur1=ur1, k2=k2, d2=d2, ur2=ur2,
                has1=(rng.random(n) < 0.3).astype(np.uint8), has2=(rng.random(len(k2)) < 0.3).astype(np.uint8),
                fv1=feature_vector(desc, words), fv2=feature_vector(d2, words),
                R1w=np.eye(3, dtype=np.float32).ravel(), t1w=np.zeros(3, np.float32),
                R2w=R2.astype(np.float32).ravel(), t2w=t2.astype(np.float32),
                sigma2=((np.float32(1.2) ** np.arange(8)) ** 2).astype(np.float32),
                scaleFactors=(np.float32(1.2) ** np.arange(8)).astype(np.float32))


# ------------------------------------------------------------------------------------------------
# optimiser scenes
# ------------------------------------------------------------------------------------------------
def se3_matrix(R, t):
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    return T


def rot_err_deg(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return np.degrees(np.arccos(np.clip(c, -1, 1)))


def pose_opt_scenario(seed, E=300, stereo_frac=0.7, outlier_frac=0.2, rot_deg=2.0, trans=0.05):
    """E 3-D points seen from a ground-truth pose; pixel noise sigma = scale[octave]; gross outliers;
    initial pose perturbed by (rot_deg, trans)."""
    rng = np.random.default_rng(seed)
    scale = 1.2 ** np.arange(8)
    R = rot_small(rng, rng.uniform(0, 30))
    t = rng.uniform(-1, 1, 3)
    Tgt = se3_matrix(R, t)
    z = rng.uniform(1.5, 15, E)
    u = rng.uniform(30, 722, E)
    v = rng.uniform(30, 450, E)
    Xc = np.stack([(u - CX) / FX * z, (v - CY) / FY * z, z], 1)
    Xw = (Xc - t) @ R          # R^T (Xc - t)
    octv = np.minimum(rng.geometric(0.35, E) - 1, 7)
    sig = scale[octv]
    obs = np.stack([u + rng.normal(0, 1, E) * sig, v + rng.normal(0, 1, E) * sig, np.full(E, -1.0)], 1)
    st = rng.random(E) < stereo_frac
    obs[st, 2] = obs[st, 0] - BF / z[st] + rng.normal(0, 1, st.sum()) * sig[st]
    out = rng.random(E) < outlier_frac
    obs[out, 0] += rng.choice([-1, 1], out.sum()) * rng.uniform(20, 80, out.sum())
    obs[out, 1] += rng.choice([-1, 1], out.sum()) * rng.uniform(20, 80, out.sum())
    Rp = rot_small(rng, rot_deg)
    Tinit = se3_matrix(Rp @ R, Rp @ t + rng.normal(0, trans, 3))
    return dict(xw=Xw.astype(np.float32), obs=obs.astype(np.float32),
                inv_sigma2=(1.0 / sig ** 2).astype(np.float32), Tcw=Tinit.astype(np.float32), Tgt=Tgt,
                is_outlier=out)


def lba_scenario(seed, K=20, M=3000, n_fixed=3, obs_range=(3, 10), outlier_frac=0.05, stereo_frac=0.6,
                 rot_deg=1.0, trans=0.03, pt_sigma=0.02):
    rng = np.random.default_rng(seed)
    scale = 1.2 ** np.arange(8)
    # cameras on a gentle arc looking at a point cloud 4-12 m ahead
    Tgt = []
    for k in range(K):
        R = rot_small(rng, 4.0) @ np.array([[np.cos(0.03 * k), 0, np.sin(0.03 * k)], [0, 1, 0],
                                             [-np.sin(0.03 * k), 0, np.cos(0.03 * k)]])
        c = np.array([0.25 * k, rng.normal(0, 0.05), rng.normal(0, 0.1)])
        Tgt.append(se3_matrix(R, -R @ c))
    Tgt = np.array(Tgt)
    P = np.stack([rng.uniform(-4, 9, M), rng.uniform(-2.5, 2.5, M), rng.uniform(4, 12, M)], 1)
    ekf, emp, obs, isg = [], [], [], []
    for m in range(M):
        want = int(rng.integers(obs_range[0], obs_range[1] + 1))
        ks = rng.permutation(K)
        got = 0
        for k in ks:
            pc = Tgt[k, :3, :3] @ P[m] + Tgt[k, :3, 3]
            if pc[2] < 0.5:
                continue
            u, v = FX * pc[0] / pc[2] + CX, FY * pc[1] / pc[2] + CY
            if not (0 < u < 752 and 0 < v < 480):
                continue
            o = min(int(rng.geometric(0.35)) - 1, 7)
            s = scale[o]
            uo, vo = u + rng.normal(0, 1) * s, v + rng.normal(0, 1) * s
            ur = -1.0
            if rng.random() < stereo_frac:
                ur = uo - BF / pc[2] + rng.normal(0, 1) * s
            if rng.random() < outlier_frac:
                uo += rng.choice([-1, 1]) * rng.uniform(15, 60)
            ekf.append(k), emp.append(m), obs.append([uo, vo, ur]), isg.append(1.0 / s ** 2)
            got += 1
            if got >= want:
                break
    # keep only points with >= 2 observations (a MapPoint in the reference's local map always has them;
    # an unobserved point would make its 3x3 Hessian block singular) and re-index
    ekf, emp = np.array(ekf, np.int32), np.array(emp, np.int32)
    obs, isg = np.array(obs, np.float32), np.array(isg, np.float32)
    cnt = np.bincount(emp, minlength=M)
    keep_pt = cnt >= 2
    remap = np.cumsum(keep_pt) - 1
    ke = keep_pt[emp]
    ekf, emp, obs, isg = ekf[ke], remap[emp[ke]].astype(np.int32), obs[ke], isg[ke]
    P = P[keep_pt]
    M = len(P)
    fixed = np.zeros(K, np.uint8)
    fixed[:n_fixed] = 1
    Tin = Tgt.copy()
    for k in range(n_fixed, K):
        Rp = rot_small(rng, rot_deg)
        Tin[k] = se3_matrix(Rp @ Tgt[k, :3, :3], Rp @ Tgt[k, :3, 3] + rng.normal(0, trans, 3))
    Pin = P + rng.normal(0, pt_sigma, P.shape)
    return dict(kf_T=Tin.astype(np.float32).reshape(K, 16), kf_fixed=fixed, mp_xyz=Pin.astype(np.float32),
                e_kf=ekf, e_mp=emp, e_obs=obs, e_inv_sigma2=isg, Tgt=Tgt, Pgt=P)


# ------------------------------------------------------------------------------------------------
# SURVEY.md §8 f2 scenarios: SearchByBoW(KeyFrame, Frame) and Fuse(KeyFrame, MapPoints)
# ------------------------------------------------------------------------------------------------
KP_DT = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])


def synthetic_keypoints(seed, n, W=752, H=480):
    """n keypoints with real ORB descriptors (orbvoc sample), geometric octave distribution, random angles."""
    rng = np.random.default_rng(seed)
    k = np.zeros(n, KP_DT)
    k["octave"] = np.minimum(rng.geometric(0.35, n) - 1, 7)
    sc_ = (np.float32(1.2) ** k["octave"]).astype(np.float32)
    k["x"] = (rng.uniform(20, W - 20, n)).astype(np.float32)
    k["y"] = (rng.uniform(20, H - 20, n)).astype(np.float32)
    k["size"] = np.floor(31 * sc_)
    k["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    k["response"] = rng.integers(20, 120, n)
    voc = orbvoc()
    d = voc[rng.choice(len(voc), n, replace=n > len(voc))].copy()
    return k, d


def bow_scenario(seed, n_kf=900, n_f=950, nwords=60):
    """A keyframe and a frame that sees ~70 % of its features again (bit-flipped descriptors, rotated by a common
    angle + noise) plus distractors; FeatureVectors from a shared stand-in vocabulary level."""
    rng = np.random.default_rng(seed)
    kK, dK = synthetic_keypoints(seed * 7 + 1, n_kf)
    kF, dF = synthetic_keypoints(seed * 7 + 2, n_f)
    common = rng.permutation(min(n_kf, n_f))[:int(0.7 * min(n_kf, n_f))]
    rot = rng.uniform(0, 40)
    for j, i in enumerate(common):
        dF[j] = flip_bits(rng, dK[i], int(rng.integers(0, 30)))
        kF["angle"][j] = (kK["angle"][i] - rot + (rng.normal(0, 2.0) if rng.random() < 0.85 else rng.uniform(0, 360))) % 360.0
        kF["octave"][j] = kK["octave"][i]
    perm = rng.permutation(n_f)
    kF, dF = kF[perm], dF[perm]
    # a few exact duplicates so that best == second best and ties occur
    for _ in range(10):
        a, b = rng.integers(n_f, size=2)
        dF[a] = dF[b]
    words = orbvoc()[rng.choice(4096, nwords, replace=False)]
    fvK = feature_vector(dK, words)
    fn, fo, fi = feature_vector(dF, words)
    keep = rng.random(len(fn)) < 0.9                      # some nodes exist in only one of the two FeatureVectors
    off, idx = [0], []
    for j in np.flatnonzero(keep):
        idx += fi[fo[j]:fo[j + 1]].tolist()
        off.append(len(idx))
    fvF = (fn[keep].astype(np.int32), np.array(off, np.int32), np.array(idx, np.int32))
    has = (rng.random(n_kf) < 0.8).astype(np.uint8)
    return dict(kK=kK, dK=dK, kF=kF, dF=dF, fvK=fvK, fvF=fvF, has=has)


def fuse_scenario(seed, n_kp=900, n_mp=700, stereo=True, th=3.0):
    rng = np.random.default_rng(seed)
    kK, dK = synthetic_keypoints(seed * 11 + 3, n_kp)
    scale = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    inv_sigma2 = (np.float32(1.0) / (scale * scale)).astype(np.float32)
    R = rot_small(rng, rng.uniform(0.5, 25.0))
    t = rng.uniform(-0.6, 0.6, 3)
    Ow = -R.T @ t
    ur = np.full(n_kp, -1.0, np.float32)
    voc = orbvoc()
    xw = np.zeros((n_mp, 3), np.float32)
    maxd, mind = np.zeros(n_mp, np.float32), np.zeros(n_mp, np.float32)
    normal = np.zeros((n_mp, 3), np.float32)
    desc = np.zeros((n_mp, 32), np.uint8)
    flags = (rng.random(n_mp) < 0.9).astype(np.uint8)
    for i in range(n_mp):
        kind = rng.random()
        if kind < 0.75:
            j = int(rng.integers(n_kp))
            z = rng.uniform(1.5, 12.0)
            u, v = kK["x"][j] + rng.normal(0, 1.5), kK["y"][j] + rng.normal(0, 1.5)
            Xc = np.array([(u - CX) / FX * z, (v - CY) / FY * z, z])
            desc[i] = flip_bits(rng, dK[j], int(rng.integers(0, 40))) if rng.random() < 0.85 else voc[int(rng.integers(len(voc)))]
            lvl = int(np.clip(kK["octave"][j] + rng.choice([0, 0, 0, 1, -1]), 0, 7))
            if stereo and ur[j] < 0 and rng.random() < 0.6:
                ur[j] = np.float32(kK["x"][j] - BF / z + rng.normal(0, 0.5))
        else:
            z = rng.uniform(-3.0, 14.0)                       # some behind the camera
            Xc = np.array([rng.uniform(-8, 8), rng.uniform(-5, 5), z])
            desc[i] = voc[int(rng.integers(len(voc)))]
            lvl = int(rng.integers(0, 8))
        Xw = R.T @ (Xc - t)
        xw[i] = Xw
        dist = np.linalg.norm(Xw - Ow)
        maxd[i] = dist * float(scale[lvl]) * rng.choice([1.0, 1.0, 1.0, 0.7, 1.6])
        mind[i] = maxd[i] / float(scale[7]) * rng.choice([1.0, 1.0, 2.5])
        nrm = (Xw - Ow) / max(dist, 1e-9)
        nrm = nrm + rng.normal(0, 0.35, 3) * (rng.random() < 0.4)
        normal[i] = nrm / np.linalg.norm(nrm) * rng.choice([1.0, 1.0, 1.0, -1.0])
    return dict(kK=kK, dK=dK, ur=ur if stereo else None, R=R.astype(np.float32).ravel(), t=t.astype(np.float32),
                Ow=Ow.astype(np.float32), flags=flags, xw=xw, maxd=maxd, mind=mind, normal=normal, desc=desc, th=th,
                scale=scale, inv_sigma2=inv_sigma2, log_sf=float(np.float32(np.log(np.float32(1.2)))))


# ------------------------------------------------------------------------------------------------
# SURVEY.md §8 f3 scenario: PoseInertialOptimizationLastKeyFrame
# ------------------------------------------------------------------------------------------------
def _exp_so3(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def inertial_scenario(seed, E=300, stereo_frac=0.6, outlier_frac=0.1, dt=0.25, noise_px=0.7, perturb=True):
    """A keyframe and a frame dt seconds later with known body states; the pre-integrated IMU deltas are the exact relative
    motion (so the inertial residual vanishes at the truth), E map points observed by the frame's left camera with pixel
    noise and a few gross outliers.  The frame's initial state is the truth perturbed by (1 deg, 3 cm, 5 cm/s, small biases)."""
    rng = np.random.default_rng(seed)
    g = np.array([0.0, 0.0, -9.81])
    # EuRoC-like body-camera extrinsics
    Rbc = _exp_so3(np.array([0.01, -0.02, 1.55]))
    tbc = np.array([-0.02, -0.06, 0.01])
    Tbc = np.eye(4); Tbc[:3, :3] = Rbc; Tbc[:3, 3] = tbc
    Tcb = np.linalg.inv(Tbc)
    Rwb1 = _exp_so3(rng.normal(0, 0.3, 3)); twb1 = rng.uniform(-1, 1, 3); v1 = rng.normal(0, 0.5, 3)
    bg1 = rng.normal(0, 0.002, 3); ba1 = rng.normal(0, 0.02, 3)
    Rwb2 = Rwb1 @ _exp_so3(rng.normal(0, 0.08, 3)); v2 = v1 + rng.normal(0, 0.3, 3); twb2 = twb1 + v1 * dt + rng.normal(0, 0.03, 3)
    dR = Rwb1.T @ Rwb2
    dV = Rwb1.T @ (v2 - v1 - g * dt)
    dP = Rwb1.T @ (twb2 - twb1 - v1 * dt - 0.5 * g * dt * dt)
    # camera pose of the frame
    Rbw2 = Rwb2.T
    Rcw = Tcb[:3, :3] @ Rbw2
    tcw = Tcb[:3, :3] @ (-Rbw2 @ twb2) + Tcb[:3, 3]
    xw = np.zeros((E, 3), np.float32); obs = np.zeros((E, 3), np.float32); isg = np.zeros(E, np.float32); close = np.zeros(E, np.uint8)
    scale = 1.2 ** np.arange(8)
    for e in range(E):
        z = rng.uniform(1.5, 18.0)
        u, v = rng.uniform(30, 720), rng.uniform(30, 450)
        Xc = np.array([(u - CX) / FX * z, (v - CY) / FY * z, z])
        Xw = Rcw.T @ (Xc - tcw)
        xw[e] = Xw
        Xc = Rcw @ xw[e].astype(np.float64) + tcw
        lvl = min(int(rng.geometric(0.4)) - 1, 7)
        s = noise_px * scale[lvl]
        uu = FX * Xc[0] / Xc[2] + CX + rng.normal(0, s)
        vv = FY * Xc[1] / Xc[2] + CY + rng.normal(0, s)
        ur = -1.0
        if rng.random() < stereo_frac:
            ur = uu - BF / Xc[2] + rng.normal(0, s)
        if rng.random() < outlier_frac:
            uu += rng.choice([-1, 1]) * rng.uniform(15, 60)
        obs[e] = (uu, vv, ur)
        isg[e] = 1.0 / (scale[lvl] ** 2)
        close[e] = Xc[2] < 10.0
    Tcw = np.eye(4); Tcw[:3, :3] = Rcw; Tcw[:3, 3] = tcw
    truth = np.concatenate([Rwb2.ravel(), twb2, v2, bg1, ba1])
    state = truth.copy()
    if perturb:
        Rp = Rwb2 @ _exp_so3(rng.normal(0, 0.01, 3))
        tp = twb2 + rng.normal(0, 0.03, 3)
        state = np.concatenate([Rp.ravel(), tp, v2 + rng.normal(0, 0.05, 3), bg1 + rng.normal(0, 1e-4, 3), ba1 + rng.normal(0, 1e-3, 3)])
        Rbw = Rp.T
        Tcw[:3, :3] = Tcb[:3, :3] @ Rbw
        Tcw[:3, 3] = Tcb[:3, :3] @ (-Rbw @ tp) + Tcb[:3, 3]
    kf = np.concatenate([Rwb1.ravel(), twb1, v1, bg1, ba1])
    preint = np.concatenate([dR.ravel(), dV, dP, [dt]])
    A = rng.normal(0, 1, (9, 9))
    infoI = A @ A.T + np.diag([4e4] * 3 + [2e3] * 3 + [8e3] * 3)      # SPD, magnitudes of a 0.25 s pre-integration covariance inverse
    infoG = np.eye(3) * 1e7 + 0.0
    infoA = np.eye(3) * 1e4 + 0.0
    return dict(xw=xw, obs=obs, isg=isg, close=close, Tcw=Tcw.astype(np.float32), Tcb=Tcb.astype(np.float32), Tbc=Tbc.astype(np.float32),
                state=state, truth=truth, kf=kf, preint=preint, infoI=infoI, infoG=infoG, infoA=infoA)


def inertial_lf_scenario(seed, E=300, stereo_frac=0.6, outlier_frac=0.1, dt=0.05, noise_px=0.7):
    """PoseInertialOptimizationLastFrame: inertial_scenario's frame pair read as (previous frame, frame), plus what the second
    function needs — the RAW pre-integration (taken at a bias that differs slightly from the previous frame's, with bias
    Jacobians, so the bias-corrected deltas explain the motion exactly at the truth) and the previous frame's
    ConstraintPoseImu (mean = its current estimate, SPD information)."""
    s = inertial_scenario(seed, E, stereo_frac, outlier_frac, dt, noise_px)
    rng = np.random.default_rng(seed + 77777)
    prev_truth = s["kf"].copy()
    dRc, dVc, dPc = s["preint"][:9].reshape(3, 3), s["preint"][9:12], s["preint"][12:15]
    dg, da = rng.normal(0, 3e-4, 3), rng.normal(0, 3e-3, 3)                 # previous-frame bias - pre-integration bias
    bpre = np.concatenate([prev_truth[15:18] - dg, prev_truth[18:21] - da])
    JRg = -dt * (np.eye(3) + rng.normal(0, 0.02, (3, 3)))
    JVg = rng.normal(0, 0.02, (3, 3))
    JVa = -dt * (np.eye(3) + rng.normal(0, 0.05, (3, 3)))
    JPg = rng.normal(0, 0.002, (3, 3))
    JPa = -0.5 * dt * dt * (np.eye(3) + rng.normal(0, 0.05, (3, 3)))
    dR0 = dRc @ _exp_so3(JRg @ dg).T
    dV0 = dVc - JVg @ dg - JVa @ da
    dP0 = dPc - JPg @ dg - JPa @ da
    # previous frame: estimate = prior mean = truth perturbed a little
    R1 = prev_truth[:9].reshape(3, 3) @ _exp_so3(rng.normal(0, 0.002, 3))
    prev = np.concatenate([R1.ravel(), prev_truth[9:12] + rng.normal(0, 0.005, 3), prev_truth[12:15] + rng.normal(0, 0.01, 3),
                           prev_truth[15:18] + rng.normal(0, 5e-5, 3), prev_truth[18:21] + rng.normal(0, 5e-4, 3)])
    sc_ = np.sqrt(np.array([3e5] * 3 + [2e5] * 3 + [5e3] * 3 + [1e8] * 3 + [1e5] * 3))
    Q = rng.normal(0, 1, (15, 15))
    Hp = (sc_[:, None] * (np.eye(15) + 0.02 * (Q @ Q.T) / 15) * sc_[None, :])
    Hp = 0.5 * (Hp + Hp.T)
    s.update(prev=prev, prev_truth=prev_truth, preint=np.concatenate([dR0.ravel(), dV0, dP0, [dt]]),
             preint_jac=np.concatenate([JRg.ravel(), JVg.ravel(), JVa.ravel(), JPg.ravel(), JPa.ravel()]), preint_bias=bpre,
             prior_state=prev.copy(), prior_H=Hp)
    del s["kf"]
    return s


# ------------------------------------------------------------------------------------------------------------------
# The local map of one stream for the many-stream tracker (orbx_track_map; SURVEY.md §8(d) workload): what Tracking
# would hold when the frame with keypoints (kps, desc, uright, depth) arrives at the true pose Tcw_true.
#   * "true" MapPoints: every keypoint, back-projected at its stereo depth (monocular keypoints get a random depth) from
#     a pixel position jittered by N(0, noise_px * scale[octave]); descriptor = the keypoint's with 0..flip_max random
#     bit flips; `outlier_frac` of them are GROSS outliers (displaced by 3..6 px * scale[octave], still inside the search
#     windows, so they are matched by descriptor and PoseOptimization has to reject them);
#   * distractors up to n_map points: genuine ORB descriptors from the reference's vocabulary at random positions;
#   * ~60 % of the true points (and 30 % of the distractors) were tracked by the last frame;
#   * mfMaxDistance / mfMinDistance / normal as MapPoint::UpdateNormalAndDepth (src/MapPoint.cc:496-560) leaves them,
#     seen from a reference keyframe near the true pose (the max distance is spread by +-12 % so that PredictScale does
#     not sit on an integer).
# Returned arrays are padded to m_cap; "gt" carries generator-side truth for assertions.
# ------------------------------------------------------------------------------------------------------------------
def track_map_scenario(seed, kps, desc, uright, depth, Tcw_true, m_cap=2048, n_map=1500, noise_px=0.6, outlier_frac=0.2,
                       flip_max=40, W=752, H=480):
    rng = np.random.default_rng(seed)
    voc = orbvoc()
    n = min(len(kps), n_map)
    nd = n_map - n
    scale = (np.float32(1.2) ** np.arange(8)).astype(np.float64)
    T = np.asarray(Tcw_true, np.float64).reshape(4, 4)
    R, t = T[:3, :3], T[:3, 3]
    Ow = -R.T @ t
    ref_center = Ow + rng.normal(0, 0.05, 3)
    # ---- true points (vectorised over the n keypoints) ----
    o = np.asarray(kps["octave"][:n], np.int64)
    z = rng.uniform(1.0, 15.0, n)
    if depth is not None:
        d = np.asarray(depth[:n], np.float64)
        z = np.where(d > 0, d, z)
    is_out = rng.random(n) < outlier_frac
    r, a = rng.uniform(3.0, 6.0, n) * scale[o], rng.uniform(0, 2 * np.pi, n)
    nx, ny = rng.normal(0, 1, n) * noise_px * scale[o], rng.normal(0, 1, n) * noise_px * scale[o]
    px = kps["x"][:n].astype(np.float64) + np.where(is_out, r * np.cos(a), nx)
    py = kps["y"][:n].astype(np.float64) + np.where(is_out, r * np.sin(a), ny)
    Pc = np.stack([(px - CX) * z / FX, (py - CY) * z / FY, z], 1)
    # each bit flips independently with probability k/256, k ~ U{0..flip_max}: the flip count is Binomial with mean k
    k = rng.integers(0, flip_max + 1, n)
    flips = rng.random((n, 256)) < (k[:, None] / 256.0)
    tdesc = np.packbits(np.unpackbits(np.ascontiguousarray(desc[:n]), axis=1) ^ flips.astype(np.uint8), axis=1)
    tang = kps["angle"][:n].astype(np.float64) + rng.normal(0, 2.0, n)
    tlast = rng.random(n) < 0.6
    # ---- distractors ----
    zd = rng.uniform(1.0, 15.0, nd)
    Pd = np.stack([(rng.uniform(0, W, nd) - CX) * zd / FX, (rng.uniform(0, H, nd) - CY) * zd / FY, zd], 1)
    od = rng.integers(0, 8, nd)
    ddesc = voc[rng.integers(0, len(voc), nd)]
    dang = rng.uniform(0, 360, nd)
    dlast = rng.random(nd) < 0.3
    # ---- assemble, padded to m_cap ----
    Pw = (np.concatenate([Pc, Pd]) - t) @ R                       # R^T (Pc - t), row-wise
    octv = np.concatenate([o, od])
    dist = np.linalg.norm(Pw - Ow, axis=1)
    md = dist * scale[octv] * rng.uniform(0.88, 1.12, n_map)
    v = Pw - ref_center
    xw = np.zeros((m_cap, 3), np.float32)
    xw[:n_map] = Pw
    mdesc = np.zeros((m_cap, 32), np.uint8)
    mdesc[:n] = tdesc
    mdesc[n:n_map] = ddesc
    last_flags, map_flags = np.zeros(m_cap, np.uint8), np.zeros(m_cap, np.uint8)
    map_flags[:n_map] = 3 | np.where(dist < 10.0, 4, 0)              # bit 2: mTrackDepth < 10 (inertial optimisers)
    last_flags[:n_map] = np.where(np.concatenate([tlast, dlast]), 3, 0)
    last_octave, last_angle = np.zeros(m_cap, np.int32), np.zeros(m_cap, np.float32)
    last_octave[:n_map] = octv
    last_angle[:n_map] = np.concatenate([tang, dang])
    maxd, mind = np.ones(m_cap, np.float32), np.ones(m_cap, np.float32)
    maxd[:n_map] = md
    mind[:n_map] = md / scale[7]
    normal = np.zeros((m_cap, 3), np.float32)
    normal[:, 2] = 1
    normal[:n_map] = v / np.linalg.norm(v, axis=1, keepdims=True)
    is_outlier = np.zeros(m_cap, bool)
    is_outlier[:n] = is_out
    return dict(n_map=np.int32(n_map), xw=xw, desc=mdesc, last_flags=last_flags, last_octave=last_octave, last_angle=last_angle,
                map_flags=map_flags, max_dist=maxd, min_dist=mind, normal=normal,
                gt=dict(n_true=n, is_outlier=is_outlier, mean_flips=float(flips.sum(1).mean()) if n else 0.0))


def stack_track_maps(maps):
    """list of per-stream track_map_scenario dicts -> the [S, m_cap, ...] host arrays of orbx_track_map"""
    out = {}
    for k in ("xw", "desc", "last_flags", "last_octave", "last_angle", "map_flags", "max_dist", "min_dist", "normal"):
        out[k] = np.ascontiguousarray(np.stack([m[k] for m in maps]))
    out["n_map"] = np.array([m["n_map"] for m in maps], np.int32)
    return out


# ------------------------------------------------------------------------------------------------------------------
# Inertial inputs of one stream for the tracker's visual-inertial TrackLocalMap (orbx_track_imu): the frame's TRUE body
# state follows from its true camera pose and the rig, the reference state (last keyframe, mode 1 / previous frame,
# mode 2) lies dt seconds earlier, and the pre-integrated deltas are the exact relative motion, so the inertial residual
# vanishes at the truth.  Velocity and bias handed to the tracker are the truth perturbed like PredictStateIMU would leave
# them.  Mode 2 adds the raw pre-integration with bias Jacobians and the previous frame's ConstraintPoseImu.
# ------------------------------------------------------------------------------------------------------------------
def track_imu_scenario(seed, Tcw_true, mode, dt=None):
    rng = np.random.default_rng(seed)
    dt = dt if dt is not None else (0.25 if mode == 1 else 0.05)
    g = np.array([0.0, 0.0, -9.81])
    Rbc = _exp_so3(np.array([0.01, -0.02, 1.55]))
    tbc = np.array([-0.02, -0.06, 0.01])
    Tbc = np.eye(4); Tbc[:3, :3] = Rbc; Tbc[:3, 3] = tbc
    Tcb = np.linalg.inv(Tbc)
    T = np.asarray(Tcw_true, np.float64).reshape(4, 4)
    Rwc = T[:3, :3].T
    Ow = -Rwc @ T[:3, 3]
    Rwb2 = Rwc @ Tcb[:3, :3]
    twb2 = Rwc @ Tcb[:3, 3] + Ow
    v2 = rng.normal(0, 0.5, 3)
    bg, ba = rng.normal(0, 0.002, 3), rng.normal(0, 0.02, 3)
    Rwb1 = Rwb2 @ _exp_so3(rng.normal(0, 0.08 if mode == 1 else 0.02, 3)).T
    v1 = v2 - rng.normal(0, 0.3 if mode == 1 else 0.05, 3)
    twb1 = twb2 - v1 * dt - rng.normal(0, 0.03 if mode == 1 else 0.003, 3)
    dR = Rwb1.T @ Rwb2
    dV = Rwb1.T @ (v2 - v1 - g * dt)
    dP = Rwb1.T @ (twb2 - twb1 - v1 * dt - 0.5 * g * dt * dt)
    ref_truth = np.concatenate([Rwb1.ravel(), twb1, v1, bg, ba])
    A = rng.normal(0, 1, (9, 9))
    out = dict(Tcb=Tcb.astype(np.float32), Tbc=Tbc.astype(np.float32),
               velocity=(v2 + rng.normal(0, 0.05, 3)).astype(np.float32),
               bias=np.concatenate([bg + rng.normal(0, 1e-4, 3), ba + rng.normal(0, 1e-3, 3)]).astype(np.float32),
               info_inertial=(A @ A.T + np.diag([4e4] * 3 + [2e3] * 3 + [8e3] * 3)).ravel(),
               info_gyro=(np.eye(3) * 1e7).ravel(), info_acc=(np.eye(3) * 1e4).ravel(),
               truth=np.concatenate([Rwb2.ravel(), twb2, v2, bg, ba]))
    if mode == 1:
        out.update(ref_state=ref_truth, preint=np.concatenate([dR.ravel(), dV, dP, [dt]]))
        return out
    dg, da = rng.normal(0, 3e-4, 3), rng.normal(0, 3e-3, 3)                 # previous-frame bias - pre-integration bias
    bpre = np.concatenate([bg - dg, ba - da])
    JRg = -dt * (np.eye(3) + rng.normal(0, 0.02, (3, 3)))
    JVg = rng.normal(0, 0.02, (3, 3))
    JVa = -dt * (np.eye(3) + rng.normal(0, 0.05, (3, 3)))
    JPg = rng.normal(0, 0.002, (3, 3))
    JPa = -0.5 * dt * dt * (np.eye(3) + rng.normal(0, 0.05, (3, 3)))
    dR0 = dR @ _exp_so3(JRg @ dg).T
    dV0 = dV - JVg @ dg - JVa @ da
    dP0 = dP - JPg @ dg - JPa @ da
    R1 = Rwb1 @ _exp_so3(rng.normal(0, 0.002, 3))
    prev = np.concatenate([R1.ravel(), twb1 + rng.normal(0, 0.005, 3), v1 + rng.normal(0, 0.01, 3), bg + rng.normal(0, 5e-5, 3),
                           ba + rng.normal(0, 5e-4, 3)])
    sc_ = np.sqrt(np.array([3e5] * 3 + [2e5] * 3 + [5e3] * 3 + [1e8] * 3 + [1e5] * 3))
    Q = rng.normal(0, 1, (15, 15))
    Hp = (sc_[:, None] * (np.eye(15) + 0.02 * (Q @ Q.T) / 15) * sc_[None, :])
    Hp = 0.5 * (Hp + Hp.T)
    out.update(ref_state=prev, preint=np.concatenate([dR0.ravel(), dV0, dP0, [dt]]),
               preint_jac=np.concatenate([JRg.ravel(), JVg.ravel(), JVa.ravel(), JPg.ravel(), JPa.ravel()]), preint_bias=bpre,
               prior_state=prev.copy(), prior_H=Hp.ravel())
    return out


def stack_track_imu(imus):
    """list of per-stream track_imu_scenario dicts -> host arrays of orbx_track_imu ([S, ...]; Tcb / Tbc shared)"""
    out = {"Tcb": np.ascontiguousarray(imus[0]["Tcb"], np.float32), "Tbc": np.ascontiguousarray(imus[0]["Tbc"], np.float32)}
    for k in ("velocity", "bias"):
        out[k] = np.ascontiguousarray(np.stack([m[k] for m in imus]), np.float32)
    for k in ("ref_state", "preint", "preint_jac", "preint_bias", "info_inertial", "info_gyro", "info_acc", "prior_state", "prior_H"):
        if k in imus[0]:
            out[k] = np.ascontiguousarray(np.stack([np.asarray(m[k], np.float64).ravel() for m in imus]), np.float64)
    return out
